"""Drop-in boundary on the GPU: the reference's Python surface (models.CRNN.CRNN, utils.get_transforms,
main.train, torch.optim.Adam state) driven the way baseline/main.py drives it, checked against the oracle."""
import copy

import numpy as np
import pytest
import torch

from oracle import crnn as ocrnn
from oracle import mel as omel
from oracle import philox
from oracle import train_step as otrain
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg(cuda_device):
    import dcase2019_task4_b200.config as cfg
    from dcase2019_task4_b200 import main as bmain
    from dcase2019_task4_b200.models.CRNN import CRNN
    from dcase2019_task4_b200.utils.utils import get_transforms, weights_init
    from dcase2019_task4_b200.utils.Scaler import Scaler
    return dict(cfg=cfg, main=bmain, CRNN=CRNN, get_transforms=get_transforms, weights_init=weights_init,
                Scaler=Scaler)


def _load(model, p):
    with torch.no_grad():
        for k, v in model.named_parameters():
            v.copy_(p[k].to(v.device))


def test_module_surface_and_autograd(pkg, cuda_device):
    cfg, CRNN = pkg["cfg"], pkg["CRNN"]
    model = CRNN(**cfg.crnn_kwargs)
    model.apply(pkg["weights_init"])                       # main.py:282
    assert [k for k, _ in model.named_parameters()] == list(ocrnn.param_shapes(10).keys())
    model = model.cuda()
    flat = model.flat_parameters()
    assert all(p.is_cuda and p.untyped_storage().data_ptr() == flat.untyped_storage().data_ptr()
               for p in model.parameters())
    p = ocrnn.init_params(seed=21)
    _load(model, p)
    x = torch.randn(3, 1, 64, 64)
    # eval mode, batch 1 as evaluation_measures.py:207
    model.eval()
    with torch.no_grad():
        s1, _ = model(x[:1].cuda())
        s_ref, _ = ocrnn.crnn_forward(x[:1], p, ocrnn.init_buffers(), training=False)
    assert H.maxerr(s1.cpu(), s_ref) <= 1e-3
    # train mode with dropout=0 kwargs: autograd through the fused Function vs torch autograd on the oracle
    kw = dict(cfg.crnn_kwargs)
    kw["dropout"] = 0
    m2 = CRNN(**kw).cuda().train()
    _load(m2, p)
    strong, weak = m2(x.cuda())
    loss = (strong ** 2).mean() + weak.sum()
    loss.backward()
    sp = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    s_ref, w_ref = ocrnn.crnn_forward(x, sp, ocrnn.init_buffers(), training=True)
    gref = torch.autograd.grad((s_ref ** 2).mean() + w_ref.sum(), list(sp.values()))
    for (k, v), gr in zip(m2.named_parameters(), gref):
        if ".conv" in k and k.endswith("bias"):
            continue
        assert H.maxerr(v.grad.cpu(), gr) <= 1e-2 * max(float(gr.abs().max()), 1e-12), k   # tf32 GEMMs
    # torch's own optimizer works on the views, checkpoint keeps the nested reference format
    opt = torch.optim.Adam(filter(lambda q: q.requires_grad, m2.parameters()), lr=0.001, betas=(0.9, 0.999))
    before = m2.flat_parameters().clone()
    opt.step()
    assert not torch.equal(before, m2.flat_parameters())
    sd = m2.state_dict()
    assert set(sd.keys()) == {"cnn", "rnn", "dense"} and len(sd["cnn"]) == 27 and len(sd["rnn"]) == 16
    assert int(sd["cnn"]["batchnorm0.num_batches_tracked"]) == 1
    m3 = CRNN(**kw).cuda()
    m3.load(parameters=sd)
    assert torch.equal(m3.cnn.cnn.conv1.weight, m2.cnn.cnn.conv1.weight)


def test_get_transforms_per_sample_matches_oracle(pkg, cuda_device):
    Scaler = pkg["Scaler"]
    rng = np.random.default_rng(0)
    feats = np.abs(rng.normal(0, 1, (87, 64))).astype(np.float32) * 30
    sc = Scaler()
    sc.load_state_dict({"mean_": rng.normal(-5, 2, 64).tolist(),
                        "mean_of_square_": (rng.uniform(150, 300, 64)).tolist()})
    label = np.zeros((108, 10))
    label[3:9, 2] = 1
    tf = pkg["get_transforms"](864, sc, augment_type="noise")
    x, x_noisy, y = tf((feats, label))
    assert tuple(x.shape) == (1, 864, 64) and tuple(x_noisy.shape) == (1, 864, 64) and y.dtype == torch.float32
    ref = omel.transform_chain(feats, sc.mean_, sc.std_, frames=864)[0]
    assert np.abs(x.cpu().numpy() - ref).max() <= 2e-5
    nz = philox.teacher_noise(87, tf._seed, 0)
    refn = omel.transform_chain(feats, sc.mean_, sc.std_, noise=nz.astype(np.float64), frames=864)[1]
    assert np.abs(x_noisy.cpu().numpy() - refn).max() <= 1e-4
    # validation transform (no noise, no scaler) returns [x, label]
    out = pkg["get_transforms"](864)((feats, label))
    assert len(out) == 2 and tuple(out[0].shape) == (1, 864, 64)


class _Loader(list):
    pass


def test_train_three_steps_match_oracle(pkg, cuda_device):
    """main.train on a 3-batch loader vs the restated loop body: Adam state, EMA alpha schedule (0.5 first),
    BN running statistics and the meters.  dropout=0 so no mask injection is needed here."""
    cfg, CRNN, bmain = pkg["cfg"], pkg["CRNN"], pkg["main"]
    kw = dict(cfg.crnn_kwargs)
    kw["dropout"] = 0
    B, T = 8, 64
    ps, pt = ocrnn.init_params(seed=31), ocrnn.init_params(seed=32)
    student, teacher = CRNN(**kw), CRNN(**kw)
    _load(student, ps)
    _load(teacher, pt)
    for q in teacher.parameters():
        q.detach_()
    student, teacher = student.train().cuda(), teacher.train().cuda()
    opt = torch.optim.Adam(filter(lambda q: q.requires_grad, student.parameters()), lr=0.001, betas=(0.9, 0.999))
    g = torch.Generator().manual_seed(5)
    batches = _Loader()
    for _ in range(3):
        x = torch.randn(B, 1, T, 64, generator=g)
        xe = x + 0.1 * torch.randn(B, 1, T, 64, generator=g)
        tgt = (torch.rand(B, T // 8, 10, generator=g) < 0.2).float()
        tgt[2:6] = -1
        batches.append((x, xe, tgt))
    wm, sm = slice(2), slice(6, 8)
    meters = bmain.train(batches, student, opt, 0, ema_model=teacher, weak_mask=wm, strong_mask=sm)

    sbuf, tbuf = ocrnn.init_buffers(), ocrnn.init_buffers()
    adam = otrain.new_adam_state(ps)
    last = None
    for i, (x, xe, tgt) in enumerate(batches):
        last, _ = otrain.train_batch(ps, sbuf, adam, x, tgt, i, len(batches), teacher_p=pt, teacher_buf=tbuf,
                                     x_ema=xe, weak_mask=wm, strong_mask=sm)
    got_s = {k: v.detach().cpu() for k, v in student.named_parameters()}
    got_t = {k: v.detach().cpu() for k, v in teacher.named_parameters()}
    # Adam's first steps move every element by ~lr * sign(g): an element whose gradient is at the tf32 noise level
    # can legitimately flip direction, so parity is asserted on the bulk (>= 99.5 % within 1e-4) and the travel bound
    n_tot = n_bad = 0
    for k in ps:
        if ".conv" in k and k.endswith("bias"):
            continue      # zero gradient behind BatchNorm: torch feeds Adam rounding noise, we feed exact zeros
        for got, ref in ((got_s[k], ps[k]), (got_t[k], pt[k])):
            d = (got.double() - ref.double()).abs()
            assert float(d.max()) <= 2 * 3 * 1e-3 + 1e-6, k
            n_tot += d.numel()
            n_bad += int((d > 1e-4).sum())
    print(f"parameters off by more than 1e-4 after 3 steps: {n_bad} of {n_tot}")
    assert n_bad <= 0.005 * n_tot
    for i in range(3):
        bn = getattr(student.cnn.cnn, f"batchnorm{i}")
        # running_mean tracks (mean of conv output) = (mean without bias) + conv bias; the reference's conv biases
        # random-walk by +-lr per step on rounding noise, ours stay put: compare bias-corrected
        rm = bn.running_mean.cpu() - got_s[f"cnn.cnn.conv{i}.bias"]
        rm_ref = sbuf[f"cnn.cnn.batchnorm{i}.running_mean"] - ps[f"cnn.cnn.conv{i}.bias"]
        # (the reference's last forward used the bias BEFORE its final +-lr update, hence the 1e-3 slack; a weight
        # whose Adam step flipped sign (the <= 0.5 % tolerated above, up to 6e-3 away) drags its output channel's
        # mean with it, so the bulk is held to 1.3e-3 and single channels to 3e-3)
        err = (rm.double() - rm_ref.double()).abs()
        assert float((err <= 1e-3 + 3e-4).double().mean()) >= 0.95 and float(err.max()) <= 3e-3, (i, float(err.max()))
        assert H.maxerr(bn.running_var.cpu(), sbuf[f"cnn.cnn.batchnorm{i}.running_var"]) <= 3e-4
    st = opt.state_dict()["state"]
    assert len(st) == 38 and float(st[0]["step"]) == 3.0
    assert H.maxerr(st[0]["exp_avg"].cpu(), adam["exp_avg"]["cnn.cnn.conv0.weight"]) <= 1e-4
    for name in ("Loss", "Strong loss", "weak_class_loss", "Consistency strong", "Consistency weak"):
        # (observed 1.0e-4 on "Strong loss": the third step's loss carries two tf32 Adam steps of parameter drift)
        assert abs(meters[name].val - last[name]) <= 3e-4 * max(1.0, abs(last[name])), name
    assert int(student.state_dict()["cnn"]["batchnorm2.num_batches_tracked"]) == 3


def test_graph_replay_matches_eager_steps(pkg, cuda_device):
    """MeanTeacherEngine.step_from_waveforms replayed from a CUDA graph (per-step scalars read from device memory)
    vs the same three iterations launched eagerly: same Philox seed / step, so the two runs differ only by the order of
    floating-point atomics."""
    cfg, CRNN, bmain = pkg["cfg"], pkg["CRNN"], pkg["main"]
    from dcase2019_task4_b200 import synth
    B, T, L = 8, 64, 511 * 64
    waves, events = synth.make_clips(3 * B, seed=7, n_samples=L)
    waves = torch.from_numpy(waves.reshape(3, B, L)).to(cuda_device)
    tgt = (torch.rand(3, B, T // 8, 10, generator=torch.Generator().manual_seed(3)) < 0.2).float()
    tgt[:, 2:6] = -1
    tgt = tgt.to(cuda_device)
    mean = torch.full((64,), -30.0, device=cuda_device)
    std = torch.full((64,), 12.0, device=cuda_device)
    ps, pt = ocrnn.init_params(seed=41), ocrnn.init_params(seed=42)
    results = []
    for use_graph in (False, True):
        student, teacher = CRNN(**cfg.crnn_kwargs), CRNN(**cfg.crnn_kwargs)
        _load(student, ps)
        _load(teacher, pt)
        for q in teacher.parameters():
            q.detach_()
        student, teacher = student.train().cuda(), teacher.train().cuda()
        student._rng_seed, student._rng_step = 1234567, 0
        opt = torch.optim.Adam(student.parameters(), lr=0.001, betas=(0.9, 0.999))
        eng = bmain.MeanTeacherEngine(student, opt, teacher, slice(2), slice(6, 8), B, T, use_graph=use_graph)
        losses = []
        for i in range(3):
            eng.step_from_waveforms(waves[i], tgt[i], mean, std, 0.5, i + 1, check=False)
            losses.append(eng.read_meters()["Loss"])
        if use_graph:
            assert len(eng._graphs) == 3 and eng.graph_launches > 0
        results.append((losses, student.flat_parameters().detach().cpu().clone(),
                        teacher.flat_parameters().detach().cpu().clone(), float(opt.state_dict()["state"][0]["step"])))
    (l0, s0, t0, n0), (l1, s1, t1, n1) = results
    assert n0 == n1 == 3.0
    for a, b in zip(l0, l1):
        assert abs(a - b) <= 1e-4 * max(1.0, abs(a)), (l0, l1)
    for a, b in ((s0, s1), (t0, t1)):
        d = (a.double() - b.double()).abs()
        assert float(d.max()) <= 6e-3 + 1e-6
        assert int((d > 1e-4).sum()) <= 0.005 * d.numel()


def test_get_predictions_batched_matches_clip_by_clip(pkg, cuda_device):
    """evaluation_measures.get_predictions (batched forward + vectorised threshold / median filter / region decoding)
    vs the reference's clip-by-clip recipe (evaluation_measures.py:203-231) on the same model."""
    import pandas as pd
    import scipy.ndimage
    from dcase2019_task4_b200 import evaluation_measures as em
    from dcase2019_task4_b200.utils.utils import ManyHotEncoder
    cfg, CRNN = pkg["cfg"], pkg["CRNN"]
    T = 136
    model = CRNN(**cfg.crnn_kwargs)
    _load(model, ocrnn.init_params(seed=8))
    with torch.no_grad():
        model.dense.bias.fill_(0.3)                     # posteriors around 0.5 so that events actually appear
        model.dense.weight.mul_(30.0)
    model = model.eval().cuda()
    g = torch.Generator().manual_seed(2)

    class _Set(list):
        pass

    ds = _Set([(torch.randn(1, T, 64, generator=g), None) for _ in range(7)])
    ds.filenames = pd.Series(["clip%d.wav" % i for i in range(7)])
    enc = ManyHotEncoder(cfg.classes, n_frames=T // cfg.pooling_time_ratio)
    got = em.get_predictions(model, ds, enc.decode_strong, pooling_time_ratio=cfg.pooling_time_ratio, batch_size=3)
    rows = []
    for i, (x, _) in enumerate(ds):
        strong, _ = model(x.cuda().unsqueeze(0))
        p = strong.squeeze(0).detach().cpu().numpy()
        p = scipy.ndimage.median_filter((p > 0.5).astype(float), (cfg.median_window, 1))
        for label, on, off in enc.decode_strong(p):
            k = cfg.pooling_time_ratio / (cfg.sample_rate / cfg.hop_length)
            rows.append((label, on * k, off * k, ds.filenames.iloc[i]))
    ref = pd.DataFrame(rows, columns=["event_label", "onset", "offset", "filename"])
    assert len(ref) > 0 and len(got) == len(ref)
    key = ["filename", "event_label", "onset"]
    a, b = got.sort_values(key).reset_index(drop=True), ref.sort_values(key).reset_index(drop=True)
    assert (a.event_label == b.event_label).all() and (a.filename == b.filename).all()
    assert np.allclose(a.onset.to_numpy(float), b.onset.to_numpy(float)) and np.allclose(a.offset.to_numpy(float), b.offset.to_numpy(float))
    f1 = em.compute_strong_metrics(got, ref).results()["overall"]["f_measure"]["f_measure"]
    assert f1 == 1.0


@pytest.mark.parametrize("use_graph", [False, True])
def test_pipelined_features_match_plain_steps(pkg, cuda_device, use_graph):
    """step_pipelined (features of batch i + 1 prepared on a side stream during iteration i) vs step_from_waveforms:
    same kernels, same Philox (seed, step) per batch, so four iterations agree up to the order of float atomics."""
    cfg, CRNN, bmain = pkg["cfg"], pkg["CRNN"], pkg["main"]
    from dcase2019_task4_b200 import synth
    B, T, L, N = 8, 64, 511 * 64, 4
    waves, _ = synth.make_clips(N * B, seed=17, n_samples=L)
    waves = torch.from_numpy(waves.reshape(N, B, L)).to(cuda_device)
    tgt = (torch.rand(N, B, T // 8, 10, generator=torch.Generator().manual_seed(13)) < 0.2).float()
    tgt[:, 2:6] = -1
    tgt = tgt.to(cuda_device)
    mean = torch.full((64,), -30.0, device=cuda_device)
    std = torch.full((64,), 12.0, device=cuda_device)
    ps, pt = ocrnn.init_params(seed=71), ocrnn.init_params(seed=72)
    results = []
    for pipelined in (False, True):
        student, teacher = CRNN(**cfg.crnn_kwargs), CRNN(**cfg.crnn_kwargs)
        _load(student, ps)
        _load(teacher, pt)
        for q in teacher.parameters():
            q.detach_()
        student, teacher = student.train().cuda(), teacher.train().cuda()
        student._rng_seed, student._rng_step = 7654321, 5
        opt = torch.optim.Adam(student.parameters(), lr=0.001, betas=(0.9, 0.999))
        eng = bmain.MeanTeacherEngine(student, opt, teacher, slice(2), slice(6, 8), B, T, use_graph=use_graph)
        losses = []
        if pipelined:
            eng.prime_features(waves[0], mean, std)
        for i in range(N):
            if pipelined:
                eng.step_pipelined(waves[(i + 1) % N], tgt[i], mean, std, 0.5, i + 1, check=False)
            else:
                eng.step_from_waveforms(waves[i], tgt[i], mean, std, 0.5, i + 1, check=False)
            losses.append(eng.read_meters()["Loss"])
        results.append((losses, student.flat_parameters().detach().cpu().clone(),
                        teacher.flat_parameters().detach().cpu().clone()))
    (l0, s0, t0), (l1, s1, t1) = results
    for a, b in zip(l0, l1):
        assert abs(a - b) <= 1e-4 * max(1.0, abs(a)), (l0, l1)
    for a, b in ((s0, s1), (t0, t1)):
        d = (a.double() - b.double()).abs()
        assert float(d.max()) <= 8e-3 + 1e-6
        assert int((d > 1e-4).sum()) <= 0.005 * d.numel()


def test_train_three_steps_match_reference_train_fixture(pkg, cuda_device):
    """main.train on the inputs of tests/golden/train_reference.npz against what the reference's OWN main.train left
    in the student / teacher parameters (the fixture is written by tests/golden/make_golden.py from the unmodified
    baseline/main.py on the CPU).  Same tolerances as the oracle comparison above, with the bulk threshold widened by
    the oracle-vs-reference residual (2.5e-5)."""
    import os
    cfg, CRNN, bmain = pkg["cfg"], pkg["CRNN"], pkg["main"]
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "train_reference.npz"))
    kw = dict(cfg.crnn_kwargs)
    kw["dropout"] = 0
    s_seed, t_seed, _ = (int(v) for v in z["seeds"])
    student, teacher = CRNN(**kw), CRNN(**kw)
    _load(student, ocrnn.init_params(seed=s_seed))
    _load(teacher, ocrnn.init_params(seed=t_seed))
    for q in teacher.parameters():
        q.detach_()
    student, teacher = student.train().cuda(), teacher.train().cuda()
    opt = torch.optim.Adam(filter(lambda q: q.requires_grad, student.parameters()), lr=0.001, betas=(0.9, 0.999))
    batches = _Loader((torch.from_numpy(z["x%d" % i]), torch.from_numpy(z["xe%d" % i]), torch.from_numpy(z["tgt%d" % i]))
                      for i in range(3))
    bmain.train(batches, student, opt, 0, ema_model=teacher, weak_mask=slice(2), strong_mask=slice(6, 8))
    stride = int(z["stride"])
    names = [k for k, _ in student.named_parameters()]
    keep = torch.cat([torch.full((v.numel(),), not (".conv" in k and k.endswith("bias")))
                      for k, v in student.named_parameters()])[::stride].numpy().astype(bool)
    assert str(z["first_name"]) == names[0]
    for model, key in ((student, "student_after"), (teacher, "teacher_after")):
        got = model.flat_parameters().detach().cpu().numpy()[::stride]
        d = np.abs(got.astype(np.float64) - z[key].astype(np.float64))[keep]
        assert d.max() <= 2 * 3 * 1e-3 + 1e-6
        assert (d > 1.25e-4).sum() <= 0.005 * d.size
    for i in range(3):
        bn = getattr(student.cnn.cnn, "batchnorm%d" % i)
        assert np.abs(bn.running_var.cpu().numpy() - z["running_var%d" % i]).max() <= 3e-4
