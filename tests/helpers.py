"""Shared helpers for the parity tests (oracle <-> CUDA path)."""
import numpy as np
import torch

from oracle import crnn as ocrnn
from oracle import philox


def flat_params(p, n_class=10):
    """Oracle {name: tensor} dict -> flat slab in named_parameters() order."""
    return torch.cat([p[k].reshape(-1) for k in ocrnn.param_shapes(n_class)]).float()


def unflat_params(flat, n_class=10):
    out = {}
    off = 0
    for k, shp in ocrnn.param_shapes(n_class).items():
        n = int(np.prod(shp))
        out[k] = flat[off:off + n].reshape(shp)
        off += n
    return out


def bn_running_flat(buf):
    return torch.cat([torch.stack([buf[f"cnn.cnn.batchnorm{i}.running_mean"],
                                   buf[f"cnn.cnn.batchnorm{i}.running_var"]]) for i in range(3)]).reshape(-1).float()


def oracle_masks(B, T, seed, step, model_id):
    """Dropout keep-masks of the CUDA RNG contract in the oracle's NCHW / [B,To,128] layouts."""
    masks = {}
    t, f = T, 64
    for i in range(3):
        m = philox.dropout_mask(B * t * f, 64, seed, 8 * model_id + i, step)
        masks[f"cnn{i}"] = torch.from_numpy(m.reshape(B, t, f, 64)).permute(0, 3, 1, 2).contiguous()
        t, f = t // 2, f // 4
    m = philox.dropout_mask(B * (T // 8), 128, seed, 8 * model_id + 3, step)
    masks["head"] = torch.from_numpy(m.reshape(B, T // 8, 128))
    return masks


def nchw_to_cl(x):
    """[B, C, T, F] -> channels-last [B, T, F, C]."""
    return x.permute(0, 2, 3, 1).contiguous()


def maxerr(a, b):
    return float((a.double() - b.double()).abs().max())


def relerr(a, b):
    return maxerr(a, b) / max(float(b.double().abs().max()), 1e-30)
