"""Parameter container with the reference's module tree (baseline/models/RNN.py:7-16).

``self.rnn`` is an ``nn.GRU`` so the 16 weight / bias names, their shapes, PyTorch's default initialisation and
``weights_init``'s orthogonal init (class name contains 'GRU') are the reference's.  The recurrence itself runs in
csrc/gru.cu.  ``BidirectionalLSTM`` (RNN.py:19-45) is never instantiated by the reference and is not built."""
from torch import nn as nn


class BidirectionalGRU(nn.Module):

    def __init__(self, n_in, n_hidden, dropout=0, num_layers=1):
        super(BidirectionalGRU, self).__init__()
        if dropout != 0:
            raise NotImplementedError("dropout_recurrent != 0 is not selected by cfg.crnn_kwargs")
        self.rnn = nn.GRU(n_in, n_hidden, bidirectional=True, dropout=dropout, batch_first=True,
                          num_layers=num_layers)

    def forward(self, input_feat):
        raise NotImplementedError("the BiGRU runs fused inside CRNN.forward (dcase_crnn_forward)")
