"""GPU parity of the stand-alone BiGRU entry (dcase_bigru_forward; BidirectionalGRU, models/RNN.py:7-16) against
``torch.nn.GRU(64, 64, num_layers=2, bidirectional=True, batch_first=True)`` on the CPU in fp32 -- the module the
reference wraps.  The recurrence is fp32 FMA with ex2 / rcp approximations (abs error ~1e-7 per gate), the input
projections are fp32 FMA GEMMs: tolerance 2e-5 on outputs in (-1, 1)."""
import os

import numpy as np
import pytest
import torch

from oracle import crnn as ocrnn
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _torch_gru(p):
    gru = torch.nn.GRU(64, 64, num_layers=2, bidirectional=True, batch_first=True)
    with torch.no_grad():
        for k, v in gru.named_parameters():
            v.copy_(p["rnn.rnn." + k])
    return gru


@pytest.mark.parametrize("B,To", [(3, 108), (1, 1), (24, 136), (256, 108)])
def test_bigru_forward_matches_torch_gru(cuda_device, B, To):
    from dcase2019_task4_b200 import kernels as K
    p = ocrnn.init_params(seed=61)
    for k in p:
        if k.startswith("rnn.rnn.bias"):
            p[k] = p[k] + 0.05 * torch.randn(p[k].shape, generator=torch.Generator().manual_seed(len(k)))
    flat = H.flat_params(p).to(cuda_device)
    off = K.param_offset("rnn.rnn.weight_ih_l0")
    assert K.param_offset("dense.weight") - off == K.GRU_PARAM_COUNT
    x = torch.randn(B, To, 64, generator=torch.Generator().manual_seed(B * 1000 + To))
    out = K.bigru_forward(x.to(cuda_device), flat[off:off + K.GRU_PARAM_COUNT])
    with torch.no_grad():
        ref, _ = _torch_gru(p)(x)
    assert tuple(out.shape) == (B, To, 128)
    assert H.maxerr(out.cpu(), ref) <= 2e-5


def test_bigru_length_limit_fails_loudly(cuda_device):
    from dcase2019_task4_b200 import kernels as K
    from dcase2019_task4_b200._lib import DcaseError
    flat = torch.zeros(K.GRU_PARAM_COUNT, device=cuda_device)
    with pytest.raises(DcaseError):
        K.bigru_forward(torch.zeros(1, 137, 64, device=cuda_device), flat)


@pytest.mark.parametrize("hidden,B,To", [(128, 3, 108), (256, 3, 108), (128, 24, 20), (256, 9, 7), (256, 1, 1)])
def test_cluster_bigru_hidden_128_256_matches_torch_gru(cuda_device, hidden, B, To):
    """BASELINE.json configs[4], hidden 128 / 256: the thread-block-cluster recurrence (dcase_bigru_forward_h) against
    torch.nn.GRU on the CPU (fp32); batch sizes that are not a multiple of the 4-clip group exercise the tail."""
    from dcase2019_task4_b200 import kernels as K
    torch.manual_seed(hidden + B)
    gru = torch.nn.GRU(64, hidden, num_layers=2, bidirectional=True, batch_first=True)
    flat = torch.cat([p.detach().reshape(-1) for _, p in gru.named_parameters()])
    x = torch.randn(B, To, 64)
    out = K.bigru_forward_h(x.to(cuda_device), flat.to(cuda_device), hidden)
    with torch.no_grad():
        ref, _ = gru(x)
    assert tuple(out.shape) == (B, To, 2 * hidden)
    assert H.maxerr(out.cpu(), ref) <= 2e-5
