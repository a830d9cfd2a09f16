"""Plain-torch fp32 restatement of the reference CRNN (TEST ORACLE).

Functional form over a flat ``{name: tensor}`` dict that uses the reference's
``named_parameters()`` / buffer names, with injectable dropout
masks so that the CUDA path (counter-based Philox masks) can be compared in
train mode.  Gradients come from torch autograd.

Follows:
* ``baseline/models/CNN.py:5-16``   GLU: ``Linear(x.permute(0,2,3,1)) * sigmoid(x)``
* ``baseline/models/CNN.py:42-67``  conv3x3 -> BatchNorm2d(eps=1e-3, momentum=0.99)
                                    -> GLU -> Dropout -> AvgPool2d((2,4)), x3
* ``baseline/models/RNN.py:7-16``   ``nn.GRU(bidirectional, batch_first)``; PyTorch gate
                                    order (r, z, n), h0 = 0, no inter-layer dropout
* ``baseline/models/CRNN.py:59-84`` squeeze/permute, dropout, sigmoid head, class-softmax
                                    attention pooling with clamp(1e-7, 1)

Pinned against the reference's own ``models.CRNN`` in
``tests/test_oracle_vs_reference.py`` (needs /root/reference) and through the
committed fixtures ``tests/golden/crnn_*.npz``.
"""
import torch
import torch.nn.functional as F

BN_EPS = 1e-3        # CNN.py:49
BN_MOMENTUM = 0.99   # CNN.py:49  (running = 0.01 * old + 0.99 * new)
N_LAYERS_CNN = 3
POOL = (2, 4)        # config.py:58


def param_shapes(n_class=10, n_ch=64, hidden=64, n_in=1):
    """Ordered {name: shape} exactly as ``CRNN(**cfg.crnn_kwargs).named_parameters()``."""
    s = {}
    for i in range(N_LAYERS_CNN):
        cin = n_in if i == 0 else n_ch
        s[f"cnn.cnn.conv{i}.weight"] = (n_ch, cin, 3, 3)
        s[f"cnn.cnn.conv{i}.bias"] = (n_ch,)
        s[f"cnn.cnn.batchnorm{i}.weight"] = (n_ch,)
        s[f"cnn.cnn.batchnorm{i}.bias"] = (n_ch,)
        s[f"cnn.cnn.glu{i}.linear.weight"] = (n_ch, n_ch)
        s[f"cnn.cnn.glu{i}.linear.bias"] = (n_ch,)
    for layer in range(2):
        nin = n_ch if layer == 0 else 2 * hidden
        for suf in ("", "_reverse"):
            s[f"rnn.rnn.weight_ih_l{layer}{suf}"] = (3 * hidden, nin)
            s[f"rnn.rnn.weight_hh_l{layer}{suf}"] = (3 * hidden, hidden)
            s[f"rnn.rnn.bias_ih_l{layer}{suf}"] = (3 * hidden,)
            s[f"rnn.rnn.bias_hh_l{layer}{suf}"] = (3 * hidden,)
    s["dense.weight"] = (n_class, 2 * hidden)
    s["dense.bias"] = (n_class,)
    s["dense_softmax.weight"] = (n_class, 2 * hidden)
    s["dense_softmax.bias"] = (n_class,)
    return s


def init_buffers(n_ch=64):
    b = {}
    for i in range(N_LAYERS_CNN):
        b[f"cnn.cnn.batchnorm{i}.running_mean"] = torch.zeros(n_ch)
        b[f"cnn.cnn.batchnorm{i}.running_var"] = torch.ones(n_ch)
        b[f"cnn.cnn.batchnorm{i}.num_batches_tracked"] = torch.zeros((), dtype=torch.long)
    return b


def cnn_block(x, p, buf, i, training, mask=None, p_drop=0.5):
    """One conv/BN/GLU/dropout/pool block.  x: [B, Cin, T, F] NCHW.
    mask: bool/float [B, 64, T, F] keep-mask (NCHW) or None (no dropout)."""
    pre = f"cnn.cnn."
    y = F.conv2d(x, p[pre + f"conv{i}.weight"], p[pre + f"conv{i}.bias"], stride=1, padding=1)
    if training:
        mean = y.mean(dim=(0, 2, 3))
        var = y.var(dim=(0, 2, 3), unbiased=False)
        if buf is not None:
            n = y.numel() // y.shape[1]
            with torch.no_grad():
                rm = buf[pre + f"batchnorm{i}.running_mean"]
                rv = buf[pre + f"batchnorm{i}.running_var"]
                rm.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * mean.detach())
                rv.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * var.detach() * n / (n - 1))
                buf[pre + f"batchnorm{i}.num_batches_tracked"] += 1
    else:
        mean = buf[pre + f"batchnorm{i}.running_mean"]
        var = buf[pre + f"batchnorm{i}.running_var"]
    g = p[pre + f"batchnorm{i}.weight"]
    b = p[pre + f"batchnorm{i}.bias"]
    y = (y - mean[None, :, None, None]) / torch.sqrt(var[None, :, None, None] + BN_EPS)
    y = y * g[None, :, None, None] + b[None, :, None, None]
    lin = F.linear(y.permute(0, 2, 3, 1), p[pre + f"glu{i}.linear.weight"], p[pre + f"glu{i}.linear.bias"])
    z = lin.permute(0, 3, 1, 2) * torch.sigmoid(y)
    if training and mask is not None:
        z = z * mask.to(z.dtype) / (1.0 - p_drop)
    return F.avg_pool2d(z, POOL)


def gru_direction(x, w_ih, w_hh, b_ih, b_hh, reverse):
    """Explicit GRU recurrence, PyTorch gate order (r, z, n). x: [B, T, In] -> [B, T, H]."""
    B, T, _ = x.shape
    H = w_hh.shape[1]
    gi = x @ w_ih.t() + b_ih
    h = x.new_zeros(B, H)
    outs = [None] * T
    order = range(T - 1, -1, -1) if reverse else range(T)
    for t in order:
        gh = h @ w_hh.t() + b_hh
        r = torch.sigmoid(gi[:, t, :H] + gh[:, :H])
        z = torch.sigmoid(gi[:, t, H:2 * H] + gh[:, H:2 * H])
        n = torch.tanh(gi[:, t, 2 * H:] + r * gh[:, 2 * H:])
        h = (1 - z) * n + z * h
        outs[t] = h
    return torch.stack(outs, dim=1)


FUSED_GRU = False     # bench.py's CPU arm sets this: same recurrence through torch's library GRU (as nn.GRU runs it)


def bigru_fused(x, p, n_layers=2):
    """The whole bidirectional stack through ``torch._VF.gru`` -- the routine ``nn.GRU.forward`` calls (RNN.py:12-15):
    same weights, same gate order, h0 = 0; used where speed matters (the CPU baseline), checked against the explicit
    loop below in tests/test_oracle_golden.py."""
    flat = [p[f"rnn.rnn.{n}_l{layer}{suf}"] for layer in range(n_layers) for suf in ("", "_reverse")
            for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
    h0 = x.new_zeros(2 * n_layers, x.shape[0], flat[1].shape[1])
    out, _ = torch._VF.gru(x, h0, flat, True, n_layers, 0.0, False, True, True)
    return out


def bigru(x, p, n_layers=2):
    if FUSED_GRU:
        return bigru_fused(x, p, n_layers)
    for layer in range(n_layers):
        outs = []
        for suf, rev in (("", False), ("_reverse", True)):
            outs.append(gru_direction(
                x, p[f"rnn.rnn.weight_ih_l{layer}{suf}"], p[f"rnn.rnn.weight_hh_l{layer}{suf}"],
                p[f"rnn.rnn.bias_ih_l{layer}{suf}"], p[f"rnn.rnn.bias_hh_l{layer}{suf}"], rev))
        x = torch.cat(outs, dim=-1)
    return x


def head(x, p, training, mask=None, p_drop=0.5):
    """CRNN.py:74-81.  x: [B, T, 2H]; mask: keep-mask [B, T, 2H] or None."""
    if training and mask is not None:
        x = x * mask.to(x.dtype) / (1.0 - p_drop)
    strong = torch.sigmoid(F.linear(x, p["dense.weight"], p["dense.bias"]))
    sof = torch.softmax(F.linear(x, p["dense_softmax.weight"], p["dense_softmax.bias"]), dim=-1)
    sof = torch.clamp(sof, min=1e-7, max=1)
    weak = (strong * sof).sum(1) / sof.sum(1)
    return strong, weak


def crnn_forward(x, p, buf, training, masks=None, p_drop=0.5, return_intermediates=False):
    """x: [B, 1, T, 64] -> (strong [B, T/8, n_class], weak [B, n_class]).

    masks: None (no dropout) or dict {"cnn0","cnn1","cnn2": [B,64,T_i,F_i], "head": [B,T/8,128]}
    keep-masks.  ``training`` selects BN batch statistics (and running-stat update in ``buf``)."""
    inter = {}
    h = x
    for i in range(N_LAYERS_CNN):
        m = None if masks is None else masks.get(f"cnn{i}")
        h = cnn_block(h, p, buf, i, training, m, p_drop)
        inter[f"cnn{i}"] = h
    h = h.squeeze(-1).permute(0, 2, 1)          # [B, T/8, 64]   CRNN.py:69-70
    h = bigru(h, p)
    inter["rnn"] = h
    m = None if masks is None else masks.get("head")
    strong, weak = head(h, p, training, m, p_drop)
    if return_intermediates:
        return strong, weak, inter
    return strong, weak


def to_reference_state_dict(p, buf):
    """Nested {cnn, rnn, dense} dict in the reference checkpoint format (CRNN.py:49-53)."""
    sd = {"cnn": {}, "rnn": {}, "dense": {}}
    for k, v in list(p.items()) + list(buf.items()):
        if k.startswith("cnn.cnn."):
            sd["cnn"][k[len("cnn.cnn."):]] = v.detach().clone()
        elif k.startswith("rnn."):
            sd["rnn"][k[len("rnn."):]] = v.detach().clone()
        elif k.startswith("dense."):
            sd["dense"][k[len("dense."):]] = v.detach().clone()
    return sd


def init_params(seed=0, n_class=10):
    """Random parameters in the spirit of utils.weights_init (utils/utils.py:205-224),
    but with widened heads / BN affine so parity tests are not trivially near 0.5."""
    g = torch.Generator().manual_seed(seed)
    p = {}
    for name, shape in param_shapes(n_class).items():
        if name.endswith("conv0.weight") or ".conv" in name and name.endswith("weight"):
            fan_in = shape[1] * 9
            fan_out = shape[0] * 9
            bound = (2.0 ** 0.5) * (6.0 / (fan_in + fan_out)) ** 0.5
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif "batchnorm" in name and name.endswith("weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif "batchnorm" in name:
            t = 0.1 * torch.randn(shape, generator=g)
        elif "rnn." in name and len(shape) > 1:
            t = torch.randn(shape, generator=g) * (1.0 / shape[1]) ** 0.5
        elif "rnn." in name:
            t = (torch.rand(shape, generator=g) * 2 - 1) / 8.0
        elif name.endswith("weight"):
            t = torch.randn(shape, generator=g) * (0.3 if name.startswith("dense") else 0.125)
        else:
            t = 0.1 * torch.randn(shape, generator=g)
        p[name] = t.float()
    return p
