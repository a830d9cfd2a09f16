"""Generates the committed golden fixtures.  Run in the BUILD container (needs /root/reference):

    python tests/golden/make_golden.py

* crnn_reference.npz  -- outputs of the UNMODIFIED reference ``baseline/models/CRNN.py`` (imported from
  /root/reference): eval-mode posteriors, train-mode (dropout=0 kwargs) posteriors, loss and a checksum of every
  parameter gradient, for parameters ``oracle.crnn.init_params(seed=7)`` (stored) and a seeded input (stored).
* scaler_reference.npz -- ``mean_``, ``mean_of_square_``, ``std_`` computed by the UNMODIFIED reference
  ``baseline/utils/Scaler.py`` (``calculate_scaler``) over seeded amplitude mels sent through the oracle's
  ApplyLog / PadOrTrunc / ToTensor chain (DataLoad.py is not importable here: librosa), one set padded
  (frames = 48 > T), one truncated (frames = 32 < T), plus ``normalize`` of one sample.
* train_reference.npz -- inputs of three mean-teacher batches and (a strided subsample of) the student / teacher
  parameters after the reference's OWN ``baseline/main.py::train`` processed them on the CPU (dropout=0 kwargs).
* mel_oracle.npz      -- float64 oracle log-mel features of three short seeded synthetic clips (librosa itself is
  not installed: these pin the restatement against accidental edits, not against librosa; "parity unpinned").
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/baseline"

from oracle import crnn as ocrnn          # noqa: E402
from oracle import mel as omel            # noqa: E402
from dcase2019_task4_b200 import synth    # noqa: E402

CRNN_KWARGS = {"n_in_channel": 1, "nclass": 10, "attention": True, "n_RNN_cell": 64, "n_layers_RNN": 2,
               "activation": "glu", "dropout": 0.5, "kernel_size": 3 * [3], "padding": 3 * [1], "stride": 3 * [1],
               "nb_filters": [64, 64, 64], "pooling": list(3 * ((2, 4),))}


def main():
    sys.path.insert(0, REF)
    from models.CRNN import CRNN
    torch.manual_seed(0)
    p = ocrnn.init_params(seed=7)
    g = torch.Generator().manual_seed(70)
    x = (torch.randn(2, 1, 64, 64, generator=g) * 1.2 + 0.1)
    rm = {f"cnn.cnn.batchnorm{i}.running_mean": 0.1 * torch.randn(64, generator=g) for i in range(3)}
    rv = {f"cnn.cnn.batchnorm{i}.running_var": 0.5 + torch.rand(64, generator=g) for i in range(3)}

    def build(**over):
        kw = dict(CRNN_KWARGS)
        kw.update(over)
        m = CRNN(**kw)
        with torch.no_grad():
            for k, v in m.named_parameters():
                v.copy_(p[k])
            for k, v in m.named_buffers():
                if k in rm:
                    v.copy_(rm[k])
                if k in rv:
                    v.copy_(rv[k])
        return m

    m = build().eval()
    with torch.no_grad():
        s_eval, w_eval = m(x)
    m = build(dropout=0).train()
    s_tr, w_tr = m(x)
    target = (torch.rand(2, 8, 10, generator=g) < 0.3).float()
    loss = torch.nn.BCELoss()(s_tr, target) + torch.nn.BCELoss()(w_tr, target.max(-2)[0])
    loss.backward()
    out = {"x": x.numpy(), "target": target.numpy(), "strong_eval": s_eval.numpy(), "weak_eval": w_eval.numpy(),
           "strong_train": s_tr.detach().numpy(), "weak_train": w_tr.detach().numpy(), "loss": np.float64(loss.item())}
    for k, v in m.named_parameters():
        out["param/" + k] = p[k].numpy()
        out["gradsum/" + k] = np.float64(v.grad.double().sum().item())
        out["gradabs/" + k] = np.float64(v.grad.double().abs().sum().item())
    for k in rm:
        out["buf/" + k] = rm[k].numpy()
    for k in rv:
        out["buf/" + k] = rv[k].numpy()
    for k, v in m.named_buffers():
        if "running" in k:
            out["buf_after_train/" + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "crnn_reference.npz"), **out)

    waves, _ = synth.make_clips(8, seed=42, n_samples=22050)
    sel = waves[[0, 3, 7]]
    rng = np.random.default_rng(3)
    mean, std = rng.normal(-20, 3, 64), rng.uniform(5, 15, 64)
    mels, clean, noisy, noise = [], [], [], []
    for w in sel:
        amp = omel.calculate_mel_spec(w.astype(np.float64))
        nz = np.abs(rng.normal(0, 0.25, amp.shape)).astype(np.float32)
        c, n = omel.transform_chain(amp, mean, std, noise=nz.astype(np.float64), frames=48)
        mels.append(amp); clean.append(c); noisy.append(n); noise.append(nz)
    np.savez_compressed(os.path.join(HERE, "mel_oracle.npz"), wave=sel, mel_amp=np.stack(mels), noise=np.stack(noise),
                        mean=mean, std=std, clean=np.stack(clean), noisy=np.stack(noisy),
                        fb_sum=np.float64(omel.mel_filterbank().astype(np.float64).sum()),
                        fb_nnz=np.int64(np.count_nonzero(omel.mel_filterbank())))
    scaler_fixture()
    train_fixture()
    print("wrote fixtures to", HERE)


def train_fixture():
    """train_reference.npz: three batches through the reference's OWN main.train (tests/scripts/ref_train_vs_oracle.py
    runs it unmodified on the CPU): inputs + a strided subsample of the student / teacher slabs afterwards."""
    import subprocess
    import tempfile
    script = os.path.join(ROOT, "tests", "scripts", "ref_train_vs_oracle.py")
    subprocess.run([sys.executable, script, os.path.join(HERE, "train_reference.npz")], cwd=tempfile.mkdtemp(),
                   check=True)


def reference_scaler():
    """The reference's Scaler class; utils/Logger.py opens Baseline.log in the CWD on import -> scratch dir."""
    import tempfile
    if REF not in sys.path:
        sys.path.insert(0, REF)
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())
    try:
        from utils.Scaler import Scaler
    finally:
        os.chdir(cwd)
    return Scaler


def scaler_fixture():
    Scaler = reference_scaler()
    waves, _ = synth.make_clips(8, seed=77, n_samples=22050)          # 44 frames each; clip 7 is half silent
    amps = np.stack([omel.calculate_mel_spec(w.astype(np.float64)) for w in waves])
    out = {"mel_amp": amps}
    for frames in (48, 32):
        data = [(torch.from_numpy(omel.transform_chain(a, None, None, frames=frames)[0]), None) for a in amps]
        sc = Scaler()
        mean, std = sc.calculate_scaler(data)
        out["mean_%d" % frames] = np.asarray(mean)
        out["mean_of_square_%d" % frames] = np.asarray(sc.mean_of_square_)
        out["std_%d" % frames] = np.asarray(std)
        if frames == 48:
            out["normalized_48"] = sc.normalize(data[0][0]).numpy()
            assert set(sc.state_dict()) == {"mean_", "mean_of_square_"}
    np.savez_compressed(os.path.join(HERE, "scaler_reference.npz"), **out)


if __name__ == "__main__":
    main()
