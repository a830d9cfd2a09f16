"""``models.CRNN.CRNN`` of the reference (baseline/models/CRNN.py) on the B200-native kernels.

Same constructor keywords, sub-module tree, ``named_parameters()`` order, nested ``state_dict`` / ``load`` /
``save`` checkpoint format (CRNN.py:39-57, incl. the quirk that ``dense_softmax`` is not saved), and the same
``forward(x [B,1,T,64]) -> (strong [B,T/8,nclass], weak [B,nclass])``.  All 38 parameters are views into ONE flat
fp32 slab (``flat_parameters()``), which is what the C ABI consumes and what lets Adam + EMA and the gradient
all-reduce be single launches.  There is no CPU path: ``forward`` raises without the CUDA library / a GPU.
"""
import warnings

import torch
import torch.nn as nn

from .. import kernels as K
from .CNN import CNN
from .RNN import BidirectionalGRU


class _CrnnFunction(torch.autograd.Function):
    """dcase_crnn_forward / dcase_crnn_backward as one autograd node over all parameters."""

    @staticmethod
    def forward(ctx, module, x, flags, seed, step, keep, *params):
        B, T = x.shape[0], x.shape[-2]
        # ``keep``: grad mode of the CALLER (inside Function.forward grad mode is always off, and needs_input_grad stays
        # True under torch.no_grad()); only a grad-enabled TRAINING forward keeps its workspace for the backward
        need_grad = bool(keep) and any(ctx.needs_input_grad[6:]) and bool(flags & K.FLAG_BN_BATCH_STATS)
        ws = module._take_workspace(B, T, x.device, keep=need_grad)
        strong, weak = K.crnn_forward(x, module._flat, module._bn_flat, flags, ws, n_class=module.nclass,
                                      seed=seed, step=step, model_id=module.model_id)
        ctx.module = module
        ctx.ws = ws if need_grad else None
        ctx.cfg = (flags, seed, step)
        if need_grad:
            ctx.save_for_backward(x, weak)
        return strong, weak

    @staticmethod
    def backward(ctx, d_strong, d_weak):
        module = ctx.module
        x, weak = ctx.saved_tensors
        flags, seed, step = ctx.cfg
        if not (flags & K.FLAG_BN_BATCH_STATS) or ctx.ws is None:
            raise NotImplementedError("backward through eval-mode BatchNorm is not on the reference's path")
        grads = K.crnn_backward(x, module._flat, flags, ctx.ws, d_strong.contiguous(), d_weak.contiguous(), weak,
                                n_class=module.nclass, seed=seed, step=step, model_id=module.model_id)
        module._give_workspace(ctx.ws, x.shape[0], x.shape[-2])
        ctx.ws = None
        out = []
        for p, (off, n) in zip(module._param_list, module._param_slices):
            out.append(grads[off:off + n].view(p.shape) if p.requires_grad else None)
        return (None, None, None, None, None, None) + tuple(out)


class CRNN(nn.Module):

    def __init__(self, n_in_channel, nclass, attention=False, activation="Relu", dropout=0,
                 train_cnn=True, rnn_type='BGRU', n_RNN_cell=64, n_layers_RNN=1, dropout_recurrent=0, **kwargs):
        super(CRNN, self).__init__()
        if not attention:
            raise NotImplementedError("attention=False (mean pooling) is not selected by cfg.crnn_kwargs")
        if rnn_type != 'BGRU' or n_RNN_cell != 64 or n_layers_RNN != 2:
            raise NotImplementedError("only the 2-layer 64-cell BGRU of cfg.crnn_kwargs is built")
        if not 1 <= nclass <= 16:
            raise NotImplementedError("nclass must be in [1, 16]")
        self.attention = attention
        self.nclass = nclass
        self.cnn = CNN(n_in_channel, activation, dropout, **kwargs)
        if not train_cnn:
            for param in self.cnn.parameters():
                param.requires_grad = False
        self.train_cnn = train_cnn
        self.rnn = BidirectionalGRU(self.cnn.nb_filters[-1], n_RNN_cell, dropout=dropout_recurrent,
                                    num_layers=n_layers_RNN)
        self.dropout = nn.Dropout(dropout)
        self.dense = nn.Linear(n_RNN_cell * 2, nclass)
        self.sigmoid = nn.Sigmoid()
        self.dense_softmax = nn.Linear(n_RNN_cell * 2, nclass)
        self.softmax = nn.Softmax(dim=-1)
        # device RNG contract (include/dcase_b200.h): key drawn once from torch's generator, counter = call index
        self.model_id = 0
        self._rng_seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        self._rng_step = 0
        self._nbt_pending = 0
        self._ws_pool = {}
        self._flat = None
        self._bn_flat = None
        self._flatten()

    # ---- flat slabs --------------------------------------------------------------------------------
    def _flatten(self):
        """(Re)build the flat parameter and BN-statistic slabs on the parameters' current device."""
        params = list(self.parameters())
        dev = params[0].device
        n_total = sum(p.numel() for p in params)
        if self._slabs_current(params, dev, n_total):
            return                                  # .cuda() / .to() / .float() that changed nothing: addresses stay stable
        # nn.GRU._apply re-packs its weights into a buffer of its own on every .cuda() (flatten_parameters), so a no-op
        # .cuda() still lands here: copy back INTO the existing slabs so their addresses (held by captured CUDA graphs,
        # engines and the optimizer state views) stay stable
        reuse = self._flat is not None and self._flat.device == dev and self._flat.numel() == n_total
        flat = self._flat if reuse else torch.empty(n_total, device=dev, dtype=torch.float32)
        slices = []
        off = 0
        with torch.no_grad():
            for p in params:
                n = p.numel()
                view = flat[off:off + n].view(p.shape)
                if not (p.dtype == torch.float32 and p.data_ptr() == view.data_ptr() and p.is_contiguous()):
                    view.copy_(p.detach())
                    p.data = view
                slices.append((off, n))
                off += n
        self._flat = flat
        self._param_list = params
        self._param_slices = slices
        reuse_bn = self._bn_flat is not None and self._bn_flat.device == dev
        bn = self._bn_flat if reuse_bn else torch.empty(3 * 2 * 64, device=dev, dtype=torch.float32)
        with torch.no_grad():
            for i in range(3):
                m = getattr(self.cnn.cnn, "batchnorm%d" % i)
                for j, name in enumerate(("running_mean", "running_var")):
                    view = bn[(2 * i + j) * 64:(2 * i + j + 1) * 64]
                    t = getattr(m, name)
                    if not (t.dtype == torch.float32 and t.data_ptr() == view.data_ptr()):
                        view.copy_(t)
                        t.data = view
        self._bn_flat = bn
        if not (reuse and reuse_bn):
            self._ws_pool = {}

    def _slabs_current(self, params, dev, n_total):
        """True when every parameter and BN statistic is still the fp32 view into the current slabs that _flatten
        made (captured CUDA graphs and the engines hold these addresses; main.py:316 re-applies .cuda() every epoch)."""
        flat, bn = self._flat, self._bn_flat
        if flat is None or bn is None or flat.device != dev or flat.numel() != n_total or bn.device != dev:
            return False
        if len(params) != len(getattr(self, "_param_list", ())):
            return False
        base, off = flat.data_ptr(), 0
        for p, q in zip(params, self._param_list):
            if p is not q or p.dtype != torch.float32 or p.data_ptr() != base + 4 * off or not p.is_contiguous():
                return False
            off += p.numel()
        for i in range(3):
            m = getattr(self.cnn.cnn, "batchnorm%d" % i)
            for j, name in enumerate(("running_mean", "running_var")):
                t = getattr(m, name)
                if t.dtype != torch.float32 or t.data_ptr() != bn.data_ptr() + 4 * (2 * i + j) * 64:
                    return False
        return True

    def _apply(self, fn, recurse=True):
        out = super(CRNN, self)._apply(fn, recurse)
        self._flatten()
        return out

    def flat_parameters(self):
        """The [214356] fp32 slab all parameters are views of (named_parameters() order)."""
        return self._flat

    def flat_bn_running(self):
        return self._bn_flat

    def _take_workspace(self, B, T, device, keep):
        key = (B, T, str(device))
        pool = self._ws_pool.setdefault(key, [])
        if not keep:
            if not pool:
                pool.append(K.new_workspace(B, T, self.nclass, device))
            return pool[0]
        return pool.pop() if pool else K.new_workspace(B, T, self.nclass, device)

    def _give_workspace(self, ws, B, T):
        """Return a workspace taken with keep=True to the pool of ITS OWN shape (at most two are kept per shape)."""
        pool = self._ws_pool.setdefault((B, T, str(ws.device)), [])
        if len(pool) < 2 and ws.numel() == K.workspace_bytes(B, T, self.nclass):
            pool.append(ws)

    def _flush_counters(self):
        if self._nbt_pending:
            with torch.no_grad():
                for i in range(3):
                    getattr(self.cnn.cnn, "batchnorm%d" % i).num_batches_tracked += self._nbt_pending
            self._nbt_pending = 0

    # ---- reference checkpoint surface (CRNN.py:33-57) ------------------------------------------------
    def load_cnn(self, parameters):
        self.cnn.load(parameters=parameters)
        if not self.train_cnn:
            for param in self.cnn.parameters():
                param.requires_grad = False

    def load(self, filename=None, parameters=None):
        if filename is not None:
            parameters = torch.load(filename, weights_only=False)
        if parameters is None:
            raise NotImplementedError("load is a filename or a list of parameters (state_dict)")
        self.cnn.load(parameters=parameters["cnn"])
        self.rnn.load_state_dict(parameters["rnn"])
        self.dense.load_state_dict(parameters["dense"])

    def state_dict(self, destination=None, prefix='', keep_vars=False):
        self._flush_counters()
        state_dict = {"cnn": self.cnn.state_dict(destination=destination, prefix=prefix, keep_vars=keep_vars),
                      "rnn": self.rnn.state_dict(destination=destination, prefix=prefix, keep_vars=keep_vars),
                      'dense': self.dense.state_dict(destination=destination, prefix=prefix, keep_vars=keep_vars)}
        return state_dict

    def save(self, filename):
        self._flush_counters()
        parameters = {'cnn': self.cnn.state_dict(), 'rnn': self.rnn.state_dict(), 'dense': self.dense.state_dict()}
        torch.save(parameters, filename)

    # ---- forward (CRNN.py:59-84) ---------------------------------------------------------------------
    def forward_flags(self):
        flags = 0
        if self.training:
            flags |= K.FLAG_BN_BATCH_STATS
            if self.dropout.p > 0:
                flags |= K.FLAG_DROPOUT
        return flags

    def next_rng(self):
        """(seed, step) of the next training forward; the step counter advances per call."""
        step = self._rng_step
        self._rng_step = (self._rng_step + 1) & 0xFFFFFFFF
        return self._rng_seed, step

    def forward(self, x):
        # input size : (batch_size, n_channels, n_frames, n_freq)
        if not x.is_cuda:
            raise RuntimeError("dcase2019_task4_b200.models.CRNN runs on a B200 only (no CPU fallback); "
                               "move the model and the input to cuda")
        if x.dim() != 4 or x.shape[1] != 1 or x.shape[3] != 64:
            raise ValueError("expected input [batch, 1, frames, 64], got %s" % (tuple(x.shape),))
        if x.shape[2] % 8 != 0:
            warnings.warn("frames not a multiple of pooling_time_ratio=8")
            raise ValueError("frames must be a multiple of 8")
        flags = self.forward_flags()
        seed, step = self.next_rng() if self.training else (self._rng_seed, 0)
        if self.training:
            self._nbt_pending += 1
        x = x.contiguous().float()
        strong, weak = _CrnnFunction.apply(self, x, flags, seed, step, torch.is_grad_enabled(), *self._param_list)
        return strong, weak
