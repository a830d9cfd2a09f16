"""Parameter containers with the reference's module tree (baseline/models/CNN.py).

Same sub-module names and classes (``conv{i}`` nn.Conv2d, ``batchnorm{i}`` nn.BatchNorm2d(eps=1e-3,
momentum=0.99), ``glu{i}.linear`` nn.Linear, ``dropout{i}``, ``pooling{i}``) so ``named_parameters()`` order,
``state_dict`` keys, ``print(model)`` and ``module.apply(weights_init)`` (utils/utils.py:205-224) behave as in the
reference.  The arithmetic does not run through these torch modules: ``CRNN.forward`` hands the flat parameter
slab to the sm_100a kernels.  Only the configuration ``cfg.crnn_kwargs`` selects (config.py:53-58) is in scope;
the relu / leakyrelu / ContextGating branches (CNN.py:50-57) raise.
"""
from collections import OrderedDict

import torch
import torch.nn as nn


class GLU(nn.Module):
    """CNN.py:5-16 -- ``Linear(x.permute(0,2,3,1)) * sigmoid(x)`` (gate from the un-projected input)."""

    def __init__(self, input_num):
        super(GLU, self).__init__()
        self.sigmoid = nn.Sigmoid()
        self.linear = nn.Linear(input_num, input_num)

    def forward(self, x):
        raise NotImplementedError("GLU runs fused inside CRNN.forward (dcase_crnn_forward)")


class CNN(nn.Module):

    def __init__(self, n_in_channel, activation="Relu", conv_dropout=0,
                 kernel_size=[3, 3, 3], padding=[1, 1, 1], stride=[1, 1, 1], nb_filters=[64, 64, 64],
                 pooling=[(1, 4), (1, 4), (1, 4)]):
        super(CNN, self).__init__()
        if activation.lower() != "glu":
            raise NotImplementedError("only activation='glu' (cfg.crnn_kwargs) is built; got %r" % (activation,))
        if (list(kernel_size) != [3, 3, 3] or list(padding) != [1, 1, 1] or list(stride) != [1, 1, 1]
                or list(nb_filters) != [64, 64, 64] or [tuple(p) for p in pooling] != [(2, 4)] * 3
                or n_in_channel != 1):
            raise NotImplementedError("only the CNN geometry of cfg.crnn_kwargs (config.py:53-58) is built")
        if conv_dropout not in (0, 0.5):
            raise NotImplementedError("dropout must be 0 or 0.5 (one Philox bit per element)")
        self.nb_filters = nb_filters
        self.conv_dropout = conv_dropout
        # the reference's child names, in its registration order (CNN.py:40-62): they are the state_dict keys
        layers = OrderedDict()
        widths = [n_in_channel] + list(nb_filters)
        for i, (c_in, c_out) in enumerate(zip(widths[:-1], widths[1:])):
            layers["conv%d" % i] = nn.Conv2d(c_in, c_out, kernel_size[i], stride[i], padding[i])
            layers["batchnorm%d" % i] = nn.BatchNorm2d(c_out, eps=1e-3, momentum=0.99)
            layers["glu%d" % i] = GLU(c_out)
            if conv_dropout is not None:
                layers["dropout%d" % i] = nn.Dropout(conv_dropout)
            layers["pooling%d" % i] = nn.AvgPool2d(pooling[i])
        self.cnn = nn.Sequential(layers)

    def load(self, filename=None, parameters=None):
        if filename is not None:
            self.cnn.load_state_dict(torch.load(filename, weights_only=False))
        elif parameters is not None:
            self.cnn.load_state_dict(parameters)
        else:
            raise NotImplementedError("load is a filename or a list of parameters (state_dict)")

    def state_dict(self, destination=None, prefix='', keep_vars=False):
        return self.cnn.state_dict(destination=destination, prefix=prefix, keep_vars=keep_vars)

    def save(self, filename):
        torch.save(self.cnn.state_dict(), filename)

    def forward(self, x):
        raise NotImplementedError("the CNN stack runs fused inside CRNN.forward (dcase_crnn_forward)")
