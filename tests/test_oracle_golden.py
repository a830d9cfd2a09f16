"""CPU: the oracle against the committed golden fixtures (made from the unmodified reference by
tests/golden/make_golden.py) and against independent implementations (torch.stft, torchaudio filterbank)."""
import os

import numpy as np
import pytest
import torch

from oracle import crnn as ocrnn
from oracle import mel as omel
from oracle import philox
from oracle import train_step as otrain

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _crnn_fixture():
    z = np.load(os.path.join(GOLD, "crnn_reference.npz"))
    p = {k[len("param/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param/")}
    buf = ocrnn.init_buffers()
    for k in z.files:
        if k.startswith("buf/"):
            buf[k[len("buf/"):]] = torch.from_numpy(z[k]).clone()
    return z, p, buf


def test_oracle_crnn_matches_reference_fixture_eval():
    z, p, buf = _crnn_fixture()
    with torch.no_grad():
        s, w = ocrnn.crnn_forward(torch.from_numpy(z["x"]), p, buf, training=False)
    assert np.abs(s.numpy() - z["strong_eval"]).max() < 2e-6
    assert np.abs(w.numpy() - z["weak_eval"]).max() < 2e-6


def test_oracle_crnn_matches_reference_fixture_train_and_grads():
    z, p, buf = _crnn_fixture()
    sp = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    s, w = ocrnn.crnn_forward(torch.from_numpy(z["x"]), sp, buf, training=True)
    assert np.abs(s.detach().numpy() - z["strong_train"]).max() < 2e-6
    assert np.abs(w.detach().numpy() - z["weak_train"]).max() < 2e-6
    target = torch.from_numpy(z["target"])
    loss, _ = otrain.mean_teacher_losses(s, w, None, None, target, slice(0, 2), slice(0, 2), 0.0)
    assert abs(float(loss.detach()) - float(z["loss"])) < 1e-6
    grads = torch.autograd.grad(loss, list(sp.values()))
    for k, g in zip(sp.keys(), grads):
        ref_abs = float(z["gradabs/" + k])
        assert abs(float(g.double().abs().sum()) - ref_abs) <= 1e-4 * ref_abs + 1e-7, k
    for k in z.files:
        if k.startswith("buf_after_train/"):
            assert np.abs(buf[k[len("buf_after_train/"):]].numpy() - z[k]).max() < 1e-5, k


def test_oracle_mel_matches_fixture():
    z = np.load(os.path.join(GOLD, "mel_oracle.npz"))
    fb = omel.mel_filterbank()
    assert int(z["fb_nnz"]) == np.count_nonzero(fb) == 1983        # SURVEY: 1,983 of 65,600 weights non-zero
    assert abs(float(z["fb_sum"]) - fb.astype(np.float64).sum()) < 1e-9
    for i in range(3):
        amp = omel.calculate_mel_spec(z["wave"][i].astype(np.float64))
        assert np.array_equal(amp, z["mel_amp"][i])
        c, n = omel.transform_chain(amp, z["mean"], z["std"], noise=z["noise"][i].astype(np.float64), frames=48)
        assert np.array_equal(c, z["clean"][i]) and np.array_equal(n, z["noisy"][i])


def test_oracle_scaler_matches_reference_scaler_fixture():
    """oracle.mel.scaler_means / scaler_std against mean_ / mean_of_square_ / std_ produced by the UNMODIFIED
    reference utils/Scaler.py (tests/golden/make_golden.py), padded (48) and truncated (32) chains."""
    z = np.load(os.path.join(GOLD, "scaler_reference.npz"))
    for frames in (48, 32):
        feats = [omel.transform_chain(a, None, None, frames=frames)[0] for a in z["mel_amp"]]
        m, m2 = omel.scaler_means(feats)
        assert np.abs(m - z["mean_%d" % frames]).max() < 1e-12
        assert np.abs(m2 - z["mean_of_square_%d" % frames]).max() < 1e-10
        assert np.abs(omel.scaler_std(m, m2) - z["std_%d" % frames]).max() < 1e-11
    feats = omel.transform_chain(z["mel_amp"][0], z["mean_48"], z["std_48"], frames=48)[0]
    assert np.abs(feats - z["normalized_48"]).max() < 1e-6          # Scaler.normalize (Scaler.py:99-105)


def test_oracle_stft_matches_torch_stft():
    rng = np.random.default_rng(0)
    y = rng.standard_normal(20000)
    S = omel.stft_magnitude(y)
    win = torch.hamming_window(2048, periodic=False, dtype=torch.float64)
    ref = torch.stft(torch.from_numpy(y), n_fft=2048, hop_length=511, window=win, center=True, pad_mode="reflect",
                     return_complex=True).abs().numpy()
    assert S.shape == ref.shape == (1025, 1 + 20000 // 511)
    assert np.abs(S - ref).max() < 1e-10


def test_oracle_filterbank_matches_torchaudio():
    ta = pytest.importorskip("torchaudio")
    ref = ta.functional.melscale_fbanks(1025, 0.0, 22050.0, 64, 44100, norm=None, mel_scale="slaney").numpy().T
    assert np.abs(omel.mel_filterbank() - ref).max() < 1e-5       # torchaudio computes it in float32


def test_oracle_mel_chain_matches_transformers_audio_utils():
    """Third independent implementation: ``transformers.audio_utils`` (numpy, written upstream to reproduce librosa's
    stft / filters.mel / amplitude_to_db).  librosa itself is still absent, so the oracle stays "parity unpinned", but
    window / centering / reflect padding / frame count / Slaney filterbank / dB floor all agree."""
    au = pytest.importorskip("transformers.audio_utils")
    rng = np.random.default_rng(0)
    y = rng.standard_normal(30000) * 0.1
    y[20000:] = 0.0                                                  # exercises amin and the top_db floor
    fb = au.mel_filter_bank(num_frequency_bins=1025, num_mel_filters=64, min_frequency=0.0, max_frequency=22050.0,
                            sampling_rate=44100, norm=None, mel_scale="slaney")
    assert np.abs(fb.T - omel.mel_filterbank()).max() < 1e-6
    S = au.spectrogram(y, np.hamming(2048), frame_length=2048, hop_length=511, fft_length=2048, power=1.0, center=True,
                       pad_mode="reflect", onesided=True, mel_filters=fb, mel_floor=0.0)
    ref = omel.calculate_mel_spec(y)
    assert S.T.shape == ref.shape == (1 + 30000 // 511, 64)
    assert np.abs(S.T - ref).max() <= 2e-7 * ref.max() + 1e-6        # ours is rounded to float32 (DatasetDcase2019Task4.py:230)
    db = au.amplitude_to_db(ref.astype(np.float64).T, reference=1.0, min_value=1e-5, db_range=80.0)
    assert np.abs(db - omel.amplitude_to_db(ref.astype(np.float64).T)).max() < 1e-9


def test_amplitude_to_db_top_db_and_amin():
    S = np.array([[1.0, 1e-7, 0.0], [10.0, 1e-3, 1e-5]])
    L = omel.amplitude_to_db(S)
    assert L.max() == pytest.approx(20.0)
    assert L.min() == pytest.approx(20.0 - 80.0)                  # floored at max - top_db
    assert omel.amplitude_to_db(np.zeros((2, 2)))[0, 0] == pytest.approx(-100.0)   # amin = 1e-5
    assert omel.pad_trunc_seq(np.ones((3, 2)), 5).shape == (5, 2) and omel.pad_trunc_seq(np.ones((7, 2)), 5).shape == (5, 2)


def test_philox_known_answer_and_mask_statistics():
    # Random123 known-answer test for philox4x32-10: counter = key = 0
    out = philox.philox4x32(0, 0, 0, 0, 0, 0)
    assert [int(v) for v in out] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    out = philox.philox4x32(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff)
    assert [int(v) for v in out] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    m = philox.dropout_mask(4096, 64, seed=123, stream=1, step=7)
    assert 0.49 < m.mean() < 0.51
    assert not np.array_equal(m, philox.dropout_mask(4096, 64, seed=123, stream=1, step=8))


def test_ramp_and_ema_schedule():
    assert otrain.ema_alpha(1) == 0.5 and otrain.ema_alpha(999) == pytest.approx(0.999) and otrain.ema_alpha(5000) == 0.999
    assert otrain.consistency_weight(0, 210) == pytest.approx(2 * np.exp(-5.0))
    assert otrain.consistency_weight(10499, 210) < 2.0 and otrain.consistency_weight(10500, 210) == 2.0


def test_step_oracle_matches_reference_train_fixture():
    """oracle.train_step.train_batch on the inputs of tests/golden/train_reference.npz reproduces what the reference's
    OWN main.train left in the student / teacher parameters (strided subsample) and in the BN running variances."""
    z = np.load(os.path.join(GOLD, "train_reference.npz"))
    s_seed, t_seed, _ = (int(v) for v in z["seeds"])
    ps, pt = ocrnn.init_params(seed=s_seed), ocrnn.init_params(seed=t_seed)
    sbuf, tbuf = ocrnn.init_buffers(), ocrnn.init_buffers()
    adam = otrain.new_adam_state(ps)
    for i in range(3):
        otrain.train_batch(ps, sbuf, adam, torch.from_numpy(z["x%d" % i]), torch.from_numpy(z["tgt%d" % i]), i, 3,
                           teacher_p=pt, teacher_buf=tbuf, x_ema=torch.from_numpy(z["xe%d" % i]),
                           weak_mask=slice(2), strong_mask=slice(6, 8))
    stride = int(z["stride"])
    names = list(ocrnn.param_shapes(10).keys())
    keep = torch.cat([torch.full((int(np.prod(ocrnn.param_shapes(10)[k])),), not (".conv" in k and k.endswith("bias")))
                      for k in names])[::stride].numpy().astype(bool)   # conv biases: rounding-noise random walk
    for p, key in ((ps, "student_after"), (pt, "teacher_after")):
        flat = torch.cat([p[k].reshape(-1) for k in names]).numpy()
        assert flat.size == int(z["n_params"])
        assert np.abs(flat[::stride] - z[key])[keep].max() <= 1e-4
    for i in range(3):
        assert np.abs(sbuf["cnn.cnn.batchnorm%d.running_var" % i].numpy() - z["running_var%d" % i]).max() <= 1e-5


def test_fused_library_gru_matches_the_explicit_recurrence():
    """bench.py's CPU arm runs the BiGRU through torch's library routine (``oracle.crnn.bigru_fused``, what the
    reference's nn.GRU calls, RNN.py:12-15); it must be the same function as the explicit (r, z, n) loop the parity tests
    use -- outputs and gradients."""
    import torch
    from oracle import crnn as ocrnn
    p = ocrnn.init_params(seed=3)
    x = torch.randn(5, 27, 64, generator=torch.Generator().manual_seed(1))
    a, b = ocrnn.bigru(x, p), ocrnn.bigru_fused(x, p)
    assert float((a - b).abs().max()) <= 2e-6
    sp = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    names = [k for k in sp if k.startswith("rnn.")]
    g1 = torch.autograd.grad(ocrnn.bigru(x, sp).pow(2).sum(), [sp[k] for k in names])
    g2 = torch.autograd.grad(ocrnn.bigru_fused(x, sp).pow(2).sum(), [sp[k] for k in names])
    for k, u, v in zip(names, g1, g2):
        assert float((u - v).abs().max()) <= 1e-5 * max(1.0, float(u.abs().max())), k
