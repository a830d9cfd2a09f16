"""GPU parity of the device Scaler: ``Scaler.means`` / ``calculate_scaler``
(utils/Scaler.py:34-97) as ``dcase_scaler_accumulate`` + ``dcase_scaler_finalize`` against the fixture written by
the UNMODIFIED reference Scaler (tests/golden/scaler_reference.npz) and against the float64 oracle.

Tolerances: the device takes the dB in float32 with ``log10f`` (the reference: numpy float32 ``log10``), so single
features differ by ~1 ulp of |L| <= 100 (8e-6); per-bin means are compared at 2e-5 dB, mean squares at 2e-3 dB^2
(|L| up to 100) and std at 1e-4 dB.  The float64 reduction itself is checked at 1e-9 on finished features."""
import os

import numpy as np
import pandas as pd
import pytest
import torch

from oracle import mel as omel

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def pkg(cuda_device):
    from dcase2019_task4_b200 import DataLoad, kernels as K, synth
    from dcase2019_task4_b200.utils.Scaler import Scaler
    from dcase2019_task4_b200.utils.utils import get_transforms
    return dict(DataLoad=DataLoad, K=K, Scaler=Scaler, get_transforms=get_transforms, synth=synth)


def _dataset(pkg, amps, frames, names=None):
    names = names or ["clip%d.wav" % i for i in range(len(amps))]
    store = dict(zip(names, amps))
    df = pd.DataFrame({"filename": names})
    return pkg["DataLoad"].DataLoadDf(df, lambda f: store[f], None, transform=pkg["get_transforms"](frames))


@pytest.mark.parametrize("frames", [48, 32])
def test_calculate_scaler_on_dataset_matches_reference_fixture(pkg, frames):
    """main.py:249-250: Scaler().calculate_scaler(ConcatDataset([...])) over datasets carrying get_transforms(frames)."""
    z = np.load(os.path.join(GOLD, "scaler_reference.npz"))
    amps = z["mel_amp"]
    ds = pkg["DataLoad"].ConcatDataset([_dataset(pkg, amps[:3], frames), _dataset(pkg, amps[3:], frames)])
    sc = pkg["Scaler"]()
    mean, std = sc.calculate_scaler(ds)
    assert mean.dtype == np.float64 and mean.shape == (64,) and type(sc.mean_) is np.ndarray
    assert np.abs(mean - z["mean_%d" % frames]).max() <= 2e-5
    assert np.abs(sc.mean_of_square_ - z["mean_of_square_%d" % frames]).max() <= 2e-3
    assert np.abs(std - z["std_%d" % frames]).max() <= 1e-4
    sd = sc.state_dict()                                             # wire format, Scaler.py:107-113
    assert set(sd) == {"mean_", "mean_of_square_"} and len(sd["mean_"]) == 64
    # and the statistics feed the fused finish kernel: normalised features against Scaler.normalize of the reference
    if frames == 48:
        x = pkg["get_transforms"](48, sc)((amps[0], np.zeros((6, 10))))[0]
        assert np.abs(x.cpu().numpy() - z["normalized_48"]).max() <= 2e-5


def test_generic_iterable_float64_reduction(pkg, cuda_device):
    """Any iterable of (features, label) with [..., 64] samples is reduced as it is; same-shape rule as Scaler.py:57-61."""
    rng = np.random.default_rng(11)
    data = [(torch.from_numpy(rng.normal(-30, 12, (1, 37, 64)).astype(np.float32)), None) for _ in range(131)]
    sc = pkg["Scaler"]()
    sc.BATCH_CLIPS = 64                                              # 131 samples: three launches (64 + 64 + 3)
    mean, std = sc.calculate_scaler(data)
    m, m2 = omel.scaler_means([d[0].numpy() for d in data])
    assert np.abs(mean - m).max() <= 1e-9 and np.abs(sc.mean_of_square_ - m2).max() <= 1e-9 * 1e3
    assert np.abs(std - omel.scaler_std(m, m2)).max() <= 1e-9
    cuda_data = [d[0].cuda() for d in data[:5]]                      # bare CUDA tensors work too
    sc2 = pkg["Scaler"]().means(cuda_data)
    assert np.abs(sc2.mean_ - omel.scaler_means([d[0].numpy() for d in data[:5]])[0]).max() <= 1e-9
    with pytest.raises(NotImplementedError):
        pkg["Scaler"]().means(data[:2] + [(torch.zeros(1, 36, 64), None)])
    with pytest.raises(ValueError):
        pkg["Scaler"]().means([])


def test_ragged_clip_lengths_and_full_size(pkg, cuda_device):
    """Clips of different length (cached features are [T, 64] with T = 1 + len // 511) are grouped per shape and padded /
    truncated to 864 by the kernel; one is silent (amin floor) and one has a silent half (top_db floor)."""
    rng = np.random.default_rng(3)
    amps = []
    for T in (864, 864, 431, 900, 864, 17):
        a = np.abs(rng.normal(0, 1, (T, 64))).astype(np.float32) * rng.uniform(0.01, 50)
        amps.append(a)
    amps[1][400:] = 0.0
    amps[4][:] = 0.0
    sc = pkg["Scaler"]()
    mean, std = sc.calculate_scaler(_dataset(pkg, amps, 864))
    m, m2 = omel.scaler_means([omel.transform_chain(a, None, None, frames=864)[0] for a in amps])
    assert np.abs(mean - m).max() <= 2e-5 and np.abs(sc.mean_of_square_ - m2).max() <= 2e-3
    assert np.abs(std - omel.scaler_std(m, m2)).max() <= 1e-4


def test_means_from_waveforms_and_linearity(pkg, cuda_device):
    """Raw clips -> dcase_logmel_fwd -> reduction; and the size-independent property: statistics of the union of two
    sets are the count-weighted mean of the sets' statistics."""
    waves, _ = pkg["synth"].make_clips(6, seed=5, n_samples=44100)
    frames = 96                                                       # 87 frames per clip, padded to 96
    sc = pkg["Scaler"]()
    mean, std = sc.calculate_scaler_from_waveforms([torch.from_numpy(waves[:4]), torch.from_numpy(waves[4:])], frames)
    amps = [omel.calculate_mel_spec(w.astype(np.float64)) for w in waves]
    m, m2 = omel.scaler_means([omel.transform_chain(a, None, None, frames=frames)[0] for a in amps])
    assert np.abs(mean - m).max() <= 5e-3 and np.abs(std - omel.scaler_std(m, m2)).max() <= 5e-3   # fp32 STFT upstream
    a = pkg["Scaler"]().means_from_waveforms([torch.from_numpy(waves[:4])], frames)
    b = pkg["Scaler"]().means_from_waveforms([torch.from_numpy(waves[4:])], frames)
    assert np.abs((4 * a.mean_ + 2 * b.mean_) / 6 - sc.mean_).max() <= 1e-10
    assert np.abs((4 * a.mean_of_square_ + 2 * b.mean_of_square_) / 6 - sc.mean_of_square_).max() <= 1e-8
    pcm = torch.from_numpy((np.clip(waves, -1, 1) * 32767).astype(np.int16))
    c = pkg["Scaler"]().means_from_waveforms([pcm], frames)           # 16-bit PCM input (soundfile scaling 1 / 32768)
    d = pkg["Scaler"]().means_from_waveforms([pcm.float() / 32768.0], frames)
    assert np.abs(c.mean_ - d.mean_).max() <= 1e-6 and np.abs(c.mean_ - sc.mean_).max() <= 0.5


def test_read_audio_and_feature_cache_from_wav_files(pkg, cuda_device, tmp_path):
    """wav ingestion (read_audio, utils/utils.py:175-193: stereo mix-down by the mean,
    16-bit PCM scaled by 1 / 32768) and the reference's cache layout
    features/sr44100_win2048_hop511_mels64_nolog/features/<name>.npy holding float32 [T, 64] amplitude mels."""
    import scipy.io.wavfile
    from dcase2019_task4_b200.DatasetDcase2019Task4 import DatasetDcase2019Task4
    from dcase2019_task4_b200.utils.utils import read_audio
    waves, _ = pkg["synth"].make_clips(2, seed=9, n_samples=30000)
    left = (np.clip(waves[0], -1, 1) * 32767).astype(np.int16)
    right = (np.clip(waves[1], -1, 1) * 32767).astype(np.int16)
    scipy.io.wavfile.write(tmp_path / "stereo.wav", 44100, np.stack([left, right], axis=1))
    scipy.io.wavfile.write(tmp_path / "mono.wav", 44100, left)
    scipy.io.wavfile.write(tmp_path / "float.wav", 44100, np.stack([waves[0], waves[1]], axis=1))
    scipy.io.wavfile.write(tmp_path / "slow.wav", 16000, left)
    audio, fs = read_audio(tmp_path / "stereo.wav", 44100)
    ref = np.mean(np.stack([left, right], axis=1) / 32768.0, axis=1)          # soundfile.read + np.mean(axis=1)
    assert fs == 44100 and audio.shape == (30000,) and np.abs(audio - ref).max() <= 1e-7
    assert np.array_equal(read_audio(tmp_path / "mono.wav")[0], (left / 32768.0).astype(np.float32))
    assert np.abs(read_audio(tmp_path / "float.wav")[0] - waves.mean(axis=0)).max() <= 1e-7
    slow, fs = read_audio(tmp_path / "slow.wav", 44100)                        # librosa.resample leg (utils.py:190-192)
    from oracle import resample as oresample
    want_slow = oresample.resample(left / 32768.0, 16000, 44100)
    assert fs == 44100 and slow.shape == want_slow.shape == (int(np.ceil(30000 * 44100 / 16000)),)
    assert np.abs(slow - want_slow).max() <= 2e-6 * max(1.0, np.abs(want_slow).max())
    ds = DatasetDcase2019Task4(str(tmp_path), base_feature_dir=str(tmp_path / "features"), save_log_feature=False)
    assert ds.feature_dir.endswith(os.path.join("sr44100_win2048_hop511_mels64_nolog", "features"))
    done = ds.extract_features_from_files(str(tmp_path), ["stereo.wav", "mono.wav", "missing.wav"])
    assert done == ["stereo.wav", "mono.wav"]
    feat = ds.get_feature_file("stereo.wav")
    assert feat.dtype == np.float32 and feat.shape == (1 + 30000 // 511, 64)
    want = omel.calculate_mel_spec(ref)
    assert np.abs(feat - want).max() <= 2e-5 * want.max()                     # same bar as tests/test_gpu_logmel.py
    assert np.array_equal(np.load(os.path.join(ds.feature_dir, "mono.npy")), ds.calculate_mel_spec(left / 32768.0))


@pytest.mark.parametrize("sr_in,sr_out,n", [(16000, 44100, 16000), (48000, 44100, 48001), (22050, 44100, 5000),
                                            (44100, 16000, 44100), (32000, 44100, 777)])
def test_audio_resample_matches_oracle(pkg, cuda_device, sr_in, sr_out, n):
    """dcase_audio_resample (read_audio's librosa.resample step, kaiser_best band-limited sinc interpolation) against the
    float64 restatement in oracle/resample.py: up- and down-sampling, lengths that do not divide, the zero-padded tail
    of librosa's fix_length."""
    from oracle import resample as oresample
    from dcase2019_task4_b200 import kernels as K
    rng = np.random.default_rng(sr_in + n)
    t = np.arange(n) / sr_in
    x = (0.6 * np.sin(2 * np.pi * 440 * t) + 0.2 * np.sin(2 * np.pi * 0.3 * sr_in * t) + 0.05 * rng.standard_normal(n))
    got = K.audio_resample(torch.from_numpy(x.astype(np.float32)).to(cuda_device), sr_in, sr_out).cpu().numpy()
    want = oresample.resample(x.astype(np.float32), sr_in, sr_out)
    assert got.shape == want.shape == (int(np.ceil(n * sr_out / sr_in)),)
    err = np.abs(got - want).max()
    print(f"{sr_in} -> {sr_out}, n = {n}: max err {err:.3e}")
    assert err <= 2e-6 * max(1.0, np.abs(want).max())
