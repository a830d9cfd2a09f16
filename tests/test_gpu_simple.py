"""GPU parity of ``main_simple_CRNN.train`` (main_simple_CRNN.py:31-82): three batches
through the drop-in ``train`` against the restated loop body (oracle.train_step.train_batch with teacher_p=None:
student forward, weak / strong BCE on slice masks, backward, Adam).  dropout=0, so no mask injection.

Tolerances as in tests/test_gpu_api.py: Adam's first steps move every element by ~lr * sign(g), so an element whose
gradient sits at the tf32 noise level may flip direction; >= 99.5 % of the parameters within 1e-4 and all within the
travel bound 2 * 3 * lr."""
import numpy as np
import pytest
import torch

from oracle import crnn as ocrnn
from oracle import train_step as otrain

pytestmark = pytest.mark.gpu


class _Loader(list):
    pass


@pytest.mark.parametrize("no_weak", [False, True])
def test_simple_train_three_steps_match_oracle(cuda_device, no_weak):
    import dcase2019_task4_b200.config as cfg
    from dcase2019_task4_b200 import main_simple_CRNN as simple
    from dcase2019_task4_b200.models.CRNN import CRNN
    kw = dict(cfg.crnn_kwargs)
    kw["dropout"] = 0
    B, T = 8, 64
    ps = ocrnn.init_params(seed=51)
    model = CRNN(**kw)
    with torch.no_grad():
        for k, v in model.named_parameters():
            v.copy_(ps[k])
    model = model.train().cuda()
    opt = torch.optim.Adam(filter(lambda q: q.requires_grad, model.parameters()), lr=0.001, betas=(0.9, 0.999))
    g = torch.Generator().manual_seed(9)
    batches = _Loader()
    for _ in range(3):
        x = torch.randn(B, 1, T, 64, generator=g)
        tgt = (torch.rand(B, T // 8, 10, generator=g) < 0.2).float()
        batches.append((x, tgt))
    wm, sm = simple.masks_for(B, no_weak=no_weak)                    # main_simple_CRNN.py:186-192
    meters = simple.train(batches, model, opt, 0, weak_mask=wm, strong_mask=sm)

    sbuf = ocrnn.init_buffers()
    adam = otrain.new_adam_state(ps)
    last = None
    for i, (x, tgt) in enumerate(batches):
        last, _ = otrain.train_batch(ps, sbuf, adam, x, tgt, i, len(batches), weak_mask=wm, strong_mask=sm)
    got = {k: v.detach().cpu() for k, v in model.named_parameters()}
    n_tot = n_bad = 0
    for k in ps:
        if ".conv" in k and k.endswith("bias"):
            continue      # zero gradient behind BatchNorm
        d = (got[k].double() - ps[k].double()).abs()
        assert float(d.max()) <= 2 * 3 * 1e-3 + 1e-6, k
        n_tot += d.numel()
        n_bad += int((d > 1e-4).sum())
    assert n_bad <= 0.005 * n_tot, (n_bad, n_tot)
    names = ["Loss", "Strong loss"] + ([] if no_weak else ["Weak loss"])
    ref = {"Loss": last["Loss"], "Strong loss": last["Strong loss"], "Weak loss": last.get("weak_class_loss")}
    for name in names:
        assert abs(meters[name].val - ref[name]) <= 1e-4 * max(1.0, abs(ref[name])), name
    assert set(meters.meters) == set(names) | {"lr"}                  # reference meter names, nothing teacher-related
    st = opt.state_dict()["state"]
    assert len(st) == 38 and float(st[0]["step"]) == 3.0
    assert int(model.state_dict()["cnn"]["batchnorm2.num_batches_tracked"]) == 3
    assert float(np.abs(model.cnn.cnn.batchnorm1.running_var.cpu().numpy()
                        - sbuf["cnn.cnn.batchnorm1.running_var"].numpy()).max()) <= 3e-4
