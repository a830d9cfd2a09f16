"""K1/K2 parity: CUDA log-mel path vs the float64 oracle (oracle/mel.py), through the C ABI."""
import numpy as np
import pytest
import torch

from oracle import mel as omel
from oracle import philox

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K(cuda_device):
    from dcase2019_task4_b200 import kernels
    return kernels


def _clips(n, L, seed=0):
    from dcase2019_task4_b200 import synth
    w, _ = synth.make_clips(n, seed=seed, n_samples=L)
    return w


def test_filterbank_matches_oracle(K):
    fb = K.mel_filterbank().numpy()
    ref = omel.mel_filterbank()
    assert fb.shape == ref.shape == (64, 1025)
    assert np.count_nonzero(ref) == np.count_nonzero(fb)
    np.testing.assert_allclose(fb, ref, rtol=0, atol=2e-7)


@pytest.mark.parametrize("L", [44100, 441000, 100000, 2048 + 511 * 3 + 7])
def test_calculate_mel_spec(K, cuda_device, L):
    B = 3 if L < 441000 else 2
    w = _clips(8, L)[[0, 3, 7][:B]]           # clip 7 has a silent second half
    got = K.logmel_fwd(torch.from_numpy(w).to(cuda_device)).cpu().numpy()
    assert got.shape == (B, 1 + L // 511, 64)
    for b in range(B):
        ref = omel.calculate_mel_spec(w[b].astype(np.float64))
        tol = 2e-5 * ref.max() + 1e-7          # fp32 FFT vs float64 oracle, relative to the clip peak
        err = np.abs(got[b] - ref).max()
        print(f"L={L} clip{b}: max|err|={err:.3e} peak={ref.max():.3e}")
        assert err <= tol


def test_pcm16_matches_float(K, cuda_device):
    w = _clips(2, 44100)
    pcm = np.clip(np.round(w * 32768.0), -32768, 32767).astype(np.int16)
    a = K.logmel_fwd(torch.from_numpy(pcm).to(cuda_device))
    b = K.logmel_fwd(torch.from_numpy(pcm.astype(np.float32) / 32768.0).to(cuda_device))
    assert torch.equal(a, b)


def test_silence_and_linearity(K, cuda_device):
    z = torch.zeros(1, 44100, device=cuda_device)
    assert float(K.logmel_fwd(z).abs().max()) == 0.0
    w = torch.from_numpy(_clips(1, 44100)).to(cuda_device)
    a = K.logmel_fwd(w)
    b = K.logmel_fwd(2.0 * w)                 # |STFT| and the mel projection are homogeneous of degree 1
    assert torch.equal(b, 2.0 * a)


@pytest.mark.parametrize("T_out", [864, 80, 100])
def test_transform_chain_with_injected_noise(K, cuda_device, T_out):
    L = 44100                                  # 87 frames: T_out 80 truncates, 100 / 864 pad with 0 dB rows
    w = _clips(8, L)[[1, 7]]
    rng = np.random.default_rng(5)
    mels = np.stack([omel.calculate_mel_spec(x.astype(np.float64)) for x in w])
    noise = np.abs(rng.normal(0, 0.25, mels.shape))
    mean = rng.normal(-20, 3, 64)
    std = rng.uniform(5, 15, 64)
    clean, noisy = K.logmel_finish(torch.from_numpy(mels).to(cuda_device),
                                   torch.from_numpy(mean.astype(np.float32)).to(cuda_device),
                                   torch.from_numpy(std.astype(np.float32)).to(cuda_device), T_out,
                                   noise=torch.from_numpy(noise.astype(np.float32)).to(cuda_device))
    for b in range(2):
        ref_c, ref_n = omel.transform_chain(mels[b], mean, std, noise=noise[b].astype(np.float32).astype(np.float64),
                                            frames=T_out)
        ec = np.abs(clean[b].cpu().numpy() - ref_c[0]).max()
        en = np.abs(noisy[b].cpu().numpy() - ref_n[0]).max()
        print(f"T_out={T_out} clip{b}: clean err {ec:.3e} noisy err {en:.3e}")
        assert ec <= 2e-5 and en <= 2e-5       # normalised dB units, fp32 log10 vs float64


def test_end_to_end_waveform_to_features(K, cuda_device):
    w = _clips(8, 441000)[[2, 7]]
    mean = np.full(64, -30.0)
    std = np.full(64, 12.0)
    amp = K.logmel_fwd(torch.from_numpy(w).to(cuda_device))
    clean = K.logmel_finish(amp, torch.from_numpy(mean.astype(np.float32)).to(cuda_device),
                            torch.from_numpy(std.astype(np.float32)).to(cuda_device), 864)
    for b in range(2):
        ref = omel.transform_chain(omel.calculate_mel_spec(w[b].astype(np.float64)), mean, std, frames=864)[0][0]
        err = np.abs(clean[b].cpu().numpy() - ref).max()
        print(f"clip{b}: log-mel err {err:.3e} (normalised units; x12 = dB)")
        assert err <= 1e-3                     # <= 0.012 dB on the quietest bins of a silent half


def test_philox_noise_matches_contract(K, cuda_device):
    B, T = 2, 87
    mels = torch.full((B, T, 64), 1.0, device=cuda_device)
    mean = torch.zeros(64, device=cuda_device)
    std = torch.ones(64, device=cuda_device)
    seed, step = 0x1234567855AA, 17
    _, noisy = K.logmel_finish(mels, mean, std, T, noisy=True, seed=seed, step=step)
    nz = philox.teacher_noise(B * T, seed, step).reshape(B, T, 64).astype(np.float64)
    L = 20 * np.log10(1.0 + nz)
    ref = np.maximum(L, L.reshape(B, -1).max(1)[:, None, None] - 80.0)
    err = np.abs(noisy.cpu().numpy() - ref).max()
    print(f"philox noise path err {err:.3e} dB")
    assert err <= 1e-4
    assert 0.15 < nz.mean() < 0.25             # E|N(0, 0.25^2)| = 0.25 * sqrt(2/pi) = 0.1995
