"""Pins the torch oracle (oracle/crnn.py, oracle/train_step.py) against the reference's own
``baseline/models/CRNN.py`` imported UNMODIFIED from /root/reference.  Runs only where the reference tree is
mounted (the build container); the GPU box relies on the committed fixtures in tests/golden/."""
import copy
import os
import sys

import numpy as np
import pytest
import torch

from oracle import crnn as ocrnn
from oracle import train_step as otrain

REF = "/root/reference/baseline"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")

CRNN_KWARGS = {"n_in_channel": 1, "nclass": 10, "attention": True, "n_RNN_cell": 64, "n_layers_RNN": 2,
               "activation": "glu", "dropout": 0.5, "kernel_size": 3 * [3], "padding": 3 * [1], "stride": 3 * [1],
               "nb_filters": [64, 64, 64], "pooling": list(3 * ((2, 4),))}   # config.py:53-58


def ref_crnn(**over):
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from models.CRNN import CRNN
    kw = dict(CRNN_KWARGS)
    kw.update(over)
    return CRNN(**kw)


def load_oracle_params(model, p, buf):
    sd = dict(model.named_parameters())
    with torch.no_grad():
        for k, v in p.items():
            sd[k].copy_(v)
        for k, v in buf.items():
            dict(model.named_buffers())[k].copy_(v)


def test_named_parameters_order_and_count():
    m = ref_crnn()
    names = [k for k, _ in m.named_parameters()]
    shapes = ocrnn.param_shapes(10)
    assert names == list(shapes.keys())
    for k, v in m.named_parameters():
        assert tuple(v.shape) == tuple(shapes[k])
    assert sum(v.numel() for v in m.parameters()) == 214356


def test_eval_forward_matches_reference():
    p = ocrnn.init_params(seed=1)
    buf = ocrnn.init_buffers()
    buf["cnn.cnn.batchnorm1.running_mean"] += 0.3
    buf["cnn.cnn.batchnorm2.running_var"] *= 1.7
    m = ref_crnn().eval()
    load_oracle_params(m, p, buf)
    x = torch.randn(3, 1, 64, 64)
    with torch.no_grad():
        s_ref, w_ref = m(x)
        s, w = ocrnn.crnn_forward(x, p, buf, training=False)
    assert float((s - s_ref).abs().max()) < 2e-6
    assert float((w - w_ref).abs().max()) < 2e-6


def test_train_forward_backward_matches_reference_without_dropout():
    p = ocrnn.init_params(seed=2)
    buf = ocrnn.init_buffers()
    m = ref_crnn(dropout=0).train()
    load_oracle_params(m, p, buf)
    x = torch.randn(4, 1, 72, 64)
    g = torch.Generator().manual_seed(0)
    strong_t, weak_t = torch.rand(4, 9, 10, generator=g), torch.rand(4, 10, generator=g)
    target = (torch.rand(4, 9, 10, generator=g) < 0.3).float()
    wm, sm = slice(0, 1), slice(3, 4)
    s_ref, w_ref = m(x)
    # main.py:95-145 verbatim structure with torch criteria
    bce, mse = torch.nn.BCELoss(), torch.nn.MSELoss()
    tw = target.max(-2)[0]
    loss_ref = bce(w_ref[wm], tw[wm]) + bce(s_ref[sm], target[sm]) + 0.8 * mse(s_ref, strong_t) + 0.8 * mse(w_ref, weak_t)
    loss_ref.backward()
    sp = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    buf2 = copy.deepcopy(buf)
    s, w = ocrnn.crnn_forward(x, sp, buf2, training=True, masks=None)
    loss, meters = otrain.mean_teacher_losses(s, w, strong_t, weak_t, target, wm, sm, 0.8)
    assert abs(float(loss) - float(loss_ref)) < 1e-6
    grads = torch.autograd.grad(loss, list(sp.values()))
    for (k, v), gmine in zip(m.named_parameters(), grads):
        scale = max(float(v.grad.abs().max()), 1e-8)
        if ".conv" in k and k.endswith("bias"):
            continue                          # analytically zero behind BatchNorm; rounding noise on both sides
        assert float((gmine - v.grad).abs().max()) <= 1e-4 * scale + 1e-9, k
    for k, v in m.named_buffers():            # running stats: momentum 0.99, unbiased variance
        assert float((buf2[k].float() - v.float()).abs().max()) < 1e-5, k


def test_adam_and_ema_match_torch_optim():
    torch.manual_seed(0)
    p = {"a": torch.randn(50), "b": torch.randn(3, 7)}
    ref = [torch.nn.Parameter(v.clone()) for v in p.values()]
    opt = torch.optim.Adam(ref, lr=0.001, betas=(0.9, 0.999))   # main.py:289-290
    st = otrain.new_adam_state(p)
    for _ in range(5):
        g = {k: torch.randn_like(v) for k, v in p.items()}
        for r, gg in zip(ref, g.values()):
            r.grad = gg.clone()
        opt.step()
        otrain.adam_update(p, g, st)
        for r, v in zip(ref, p.values()):
            assert float((r.detach() - v).abs().max()) < 1e-6


def test_rampup_matches_reference_ramps():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from utils import ramps
    for cur, length in [(0, 10500), (1, 10500), (5000, 10500), (10500, 10500), (20000, 10500), (3, 0)]:
        assert otrain.sigmoid_rampup(cur, length) == pytest.approx(ramps.sigmoid_rampup(cur, length), abs=1e-15)
    assert otrain.consistency_weight(0, 210) == pytest.approx(2 * np.exp(-5.0))
    assert otrain.consistency_weight(10500, 210) == 2.0


def test_state_dict_format_roundtrip():
    p = ocrnn.init_params(seed=3)
    buf = ocrnn.init_buffers()
    m = ref_crnn()
    m.load(parameters=ocrnn.to_reference_state_dict(p, buf))
    sd = m.state_dict()
    assert set(sd.keys()) == {"cnn", "rnn", "dense"}
    assert torch.equal(sd["cnn"]["conv1.weight"], p["cnn.cnn.conv1.weight"])
    assert torch.equal(sd["rnn"]["rnn.weight_hh_l1_reverse"], p["rnn.rnn.weight_hh_l1_reverse"])


def test_scaler_means_match_reference_scaler(tmp_path, monkeypatch):
    """oracle.mel.scaler_means / scaler_std against the reference's utils/Scaler.py run live (float32 [1,T,64]
    samples as the transform chain yields them), incl. the state_dict wire format."""
    from oracle import mel as omel
    if REF not in sys.path:
        sys.path.insert(0, REF)
    monkeypatch.chdir(tmp_path)                      # utils/Logger.py opens Baseline.log in the CWD on import
    from utils.Scaler import Scaler
    rng = np.random.default_rng(5)
    data = [(torch.from_numpy(rng.normal(-30, 12, (1, 37, 64)).astype(np.float32)), None) for _ in range(7)]
    sc = Scaler()
    mean, std = sc.calculate_scaler(data)
    m, m2 = omel.scaler_means([d[0].numpy() for d in data])
    assert np.abs(m - mean).max() < 1e-12 and np.abs(m2 - sc.mean_of_square_).max() < 1e-10
    assert np.abs(omel.scaler_std(m, m2) - std).max() < 1e-11
    sd = sc.state_dict()
    assert set(sd) == {"mean_", "mean_of_square_"} and isinstance(sd["mean_"], list) and len(sd["mean_"]) == 64


def test_our_checkpoint_loads_into_reference_modules(tmp_path, monkeypatch):
    """A checkpoint written by dcase2019_task4_b200.main.save_checkpoint is consumed by the reference's own
    TestModel.py:30-38 recipe (CRNN(**kwargs).load(parameters=...), Scaler.load_state_dict) and gives the same
    eval posteriors as the oracle with those parameters."""
    from dcase2019_task4_b200 import config as cfg, main as bmain
    from dcase2019_task4_b200.models.CRNN import CRNN as OurCRNN
    from dcase2019_task4_b200.utils.Scaler import Scaler as OurScaler
    from dcase2019_task4_b200.utils.utils import ManyHotEncoder
    monkeypatch.chdir(tmp_path)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from utils.Scaler import Scaler as RefScaler
    p = ocrnn.init_params(seed=9)
    ours = OurCRNN(**cfg.crnn_kwargs)
    with torch.no_grad():
        for k, v in ours.named_parameters():
            v.copy_(p[k])
    opt = torch.optim.Adam(ours.parameters(), lr=0.001, betas=(0.9, 0.999))
    sc = OurScaler()
    sc.load_state_dict({"mean_": np.linspace(-5, 5, 64).tolist(), "mean_of_square_": np.linspace(30, 90, 64).tolist()})
    enc = ManyHotEncoder(cfg.classes, n_frames=108)
    state = bmain.build_state(ours, opt, cfg.crnn_kwargs, {"lr": 0.001, "betas": (0.9, 0.999)}, 8, sc, enc)
    bmain.save_checkpoint(bmain.update_state(state, ours, opt, 0), tmp_path / "ck")
    back = torch.load(tmp_path / "ck", map_location="cpu", weights_only=False)      # TestModel.py:75
    ref = ref_crnn(**{k: v for k, v in back["model"]["kwargs"].items() if k not in CRNN_KWARGS})
    ref.load(parameters=back["model"]["state_dict"])
    rs = RefScaler()
    rs.load_state_dict(back["scaler"])
    assert np.array_equal(rs.std_, sc.std_)
    x = torch.randn(2, 1, 64, 64)
    ref.eval()
    with torch.no_grad():
        s_ref, _ = ref(x)
        pp = dict(p)
        pp["dense_softmax.weight"] = dict(ref.named_parameters())["dense_softmax.weight"].detach()   # not in the checkpoint
        pp["dense_softmax.bias"] = dict(ref.named_parameters())["dense_softmax.bias"].detach()
        s, _ = ocrnn.crnn_forward(x, pp, ocrnn.init_buffers(), training=False)
    assert float((s - s_ref).abs().max()) < 2e-6


def test_config_matches_reference_config():
    """Every public name of baseline/config.py exists in dcase2019_task4_b200.config with the same value."""
    import importlib.util
    from dcase2019_task4_b200 import config as ours
    spec = importlib.util.spec_from_file_location("_ref_config", os.path.join(REF, "config.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)                       # reads ../dataset/metadata/validation/validation.tsv
    names = [n for n in vars(ref) if not n.startswith("_") and n not in ("math", "os", "pd", "file_path")]
    assert len(names) >= 35
    for n in names:
        a, b = getattr(ref, n), getattr(ours, n)
        if n == "classes":
            assert list(a) == list(b)
        elif n == "crnn_kwargs":
            assert set(a) == set(b)
            for k in a:
                assert list(a[k]) == list(b[k]) if isinstance(a[k], (list, tuple)) else a[k] == b[k], k
        else:
            assert a == b and type(a) is type(b), n


def test_reference_scripts_import_against_this_package(tmp_path):
    """dropin.install() aliases this package under the names baseline/main.py:19-29 imports; the reference's OWN
    main.py, main_simple_CRNN.py and TestModel.py then import unchanged (every ``from X import a, b, c`` resolves), and
    the objects they bind are ours.  Runs in a subprocess: the aliases are process-global."""
    import subprocess
    code = (
        "import sys\n"
        "sys.path.insert(0, %r)\n"
        "from dcase2019_task4_b200 import dropin\n"
        "dropin.install()\n"
        "sys.path.append(%r)\n"
        "import main, main_simple_CRNN, TestModel\n"
        "assert main.CRNN.__module__ == 'dcase2019_task4_b200.models.CRNN'\n"
        "assert main.Scaler.__module__ == 'dcase2019_task4_b200.utils.Scaler'\n"
        "assert main.get_transforms.__module__ == 'dcase2019_task4_b200.utils.utils'\n"
        "assert main.MultiStreamBatchSampler.__module__ == 'dcase2019_task4_b200.DataLoad'\n"
        "assert TestModel.get_predictions.__module__ == 'dcase2019_task4_b200.evaluation_measures'\n"
        "assert main.__file__.startswith(%r) and callable(main.train) and callable(main_simple_CRNN.train)\n"
        "m = main.CRNN(**main.cfg.crnn_kwargs); m.apply(main.weights_init)\n"
        "opt = __import__('torch').optim.Adam(filter(lambda p: p.requires_grad, m.parameters()), lr=0.001)\n"
        "print('DROPIN-OK', len(list(m.parameters())))\n"
    ) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), REF, REF)
    r = subprocess.run([sys.executable, "-c", code], cwd=str(tmp_path), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "DROPIN-OK 38" in r.stdout, r.stdout + r.stderr
    assert not os.path.exists(tmp_path / "Baseline.log")


def test_config0_full_geometry_logmel_plus_reference_crnn_forward():
    """BASELINE.json configs[0] at the real geometry (10-s clips -> 864 frames -> 108 output frames), scaled to 8 clips
    to stay within seconds: restated float64 log-mel -> dB -> z-score, then the UNMODIFIED reference models.CRNN in eval
    mode against the torch oracle on the same features (the pin the GPU parity tests inherit at full size)."""
    from oracle import mel as omel
    from dcase2019_task4_b200 import synth
    waves, _ = synth.make_clips(8, seed=123)
    amps = [omel.calculate_mel_spec(w.astype(np.float64)) for w in waves]
    assert amps[0].shape == (864, 64)
    feats = [omel.transform_chain(a, None, None, frames=864)[0] for a in amps]
    m_, m2_ = omel.scaler_means(feats)
    x = np.stack([omel.transform_chain(a, m_, omel.scaler_std(m_, m2_), frames=864)[0] for a in amps])
    x = torch.from_numpy(x)                                          # [8, 1, 864, 64]
    p = ocrnn.init_params(seed=5)
    buf = ocrnn.init_buffers()
    m = ref_crnn().eval()
    load_oracle_params(m, p, buf)
    with torch.no_grad():
        s_ref, w_ref = m(x)
        s, w = ocrnn.crnn_forward(x, p, buf, training=False)
    assert tuple(s_ref.shape) == (8, 108, 10) and tuple(w_ref.shape) == (8, 10)
    assert float((s - s_ref).abs().max()) < 5e-6 and float((w - w_ref).abs().max()) < 5e-6


def test_step_oracle_matches_the_reference_train_functions(tmp_path):
    """The reference's OWN ``main.train`` (mean teacher) and ``main_simple_CRNN.train``, imported unmodified and run on
    the CPU with the reference's ``models.CRNN``, against ``oracle.train_step.train_batch`` for three batches
    (tests/scripts/ref_train_vs_oracle.py): loss composition, ramp-up, masks, Adam, EMA alpha schedule, BN statistics.
    Adam's first steps move every element by ~lr, so 1e-4 on parameters after three steps is a tight bound."""
    import subprocess
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "scripts", "ref_train_vs_oracle.py")
    r = subprocess.run([sys.executable, script], cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("REF-TRAIN-OK")]
    assert r.returncode == 0 and lines, r.stdout[-2000:] + r.stderr[-2000:]
    d_student, d_teacher, d_simple, d_var, d_mean = (float(v) for v in lines[0].split()[1:])
    assert d_student <= 1e-4 and d_teacher <= 1e-4 and d_simple <= 1e-4
    assert d_var <= 1e-5
    assert d_mean <= 2.5e-3          # conv biases random-walk by +-lr on rounding noise (see the script)


def test_host_logic_matches_reference_utils_and_dataload(tmp_path):
    """ManyHotEncoder, DataLoadDf, ConcatDataset, MultiStreamBatchSampler, the transform chain's order / noise / padding /
    normalisation, SaveBest and AverageMeterSet against the reference's OWN utils/utils.py and DataLoad.py, imported
    unmodified with stub modules for their absent third-party imports (tests/scripts/ref_hostlogic_vs_ours.py)."""
    import subprocess
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "scripts", "ref_hostlogic_vs_ours.py")
    r = subprocess.run([sys.executable, script], cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("REF-HOST-OK")]
    assert r.returncode == 0 and lines and int(lines[0].split()[1]) >= 35, r.stdout[-2000:] + r.stderr[-2000:]
