"""float64 numpy restatement of the reference's log-mel pipeline (TEST ORACLE).

Follows, stage by stage:

* ``baseline/DatasetDcase2019Task4.py:197-231``  ``calculate_mel_spec``
  (``np.hamming`` window, ``librosa.stft(center=True, pad_mode='reflect')``,
  ``librosa.feature.melspectrogram(S=|X|, htk=False, norm=None)``, ``.T``,
  ``astype(float32)``)
* ``baseline/DataLoad.py:274-287``  ``AugmentGaussianNoise``  (|N(0, 0.25)| on
  the amplitude mel, std hard-coded to ``0.5 ** 2``)
* ``baseline/DataLoad.py:192-207``  ``ApplyLog`` -> ``librosa.amplitude_to_db``
  (``ref=1, amin=1e-5, top_db=80``, clip-global max)
* ``baseline/DataLoad.py:210-259``  ``pad_trunc_seq`` / ``PadOrTrunc``
* ``baseline/DataLoad.py:302-321``  ``ToTensor(unsqueeze_axis=0)``
* ``baseline/utils/Scaler.py:99-105``  ``Scaler.normalize``
* order fixed by ``baseline/utils/utils.py:397-412``  ``get_transforms``

librosa (un-vendored, unpinned: ``environment.yml:17``, README ">=0.6.3") is
not installed here, so the four librosa calls are restated from its published
algorithm.  PARITY UNPINNED against librosa itself; cross-checked in
``tests/test_oracle_golden.py`` against ``torch.stft``,
``torchaudio.functional.melscale_fbanks`` and ``transformers.audio_utils``
(spectrogram / mel_filter_bank / amplitude_to_db).
"""
import numpy as np

# baseline/config.py:17-25
SAMPLE_RATE = 44100
N_WINDOW = 2048
HOP_LENGTH = 511
N_MELS = 64
F_MIN = 0.0
F_MAX = 22050.0
MAX_FRAMES = 864  # math.ceil(10. * 44100 / 511), config.py:22


# ---- librosa.core.convert.hz_to_mel / mel_to_hz (Slaney, htk=False) ---------
_F_SP = 200.0 / 3
_MIN_LOG_HZ = 1000.0
_MIN_LOG_MEL = _MIN_LOG_HZ / _F_SP
_LOGSTEP = np.log(6.4) / 27.0


def hz_to_mel(freq):
    freq = np.asarray(freq, dtype=np.float64)
    mels = freq / _F_SP
    log_t = freq >= _MIN_LOG_HZ
    safe = np.where(log_t, freq, _MIN_LOG_HZ)
    return np.where(log_t, _MIN_LOG_MEL + np.log(safe / _MIN_LOG_HZ) / _LOGSTEP, mels)


def mel_to_hz(mels):
    mels = np.asarray(mels, dtype=np.float64)
    freqs = _F_SP * mels
    log_t = mels >= _MIN_LOG_MEL
    return np.where(log_t, _MIN_LOG_HZ * np.exp(_LOGSTEP * (mels - _MIN_LOG_MEL)), freqs)


def mel_filterbank(sr=SAMPLE_RATE, n_fft=N_WINDOW, n_mels=N_MELS, fmin=F_MIN, fmax=F_MAX):
    """librosa.filters.mel(..., htk=False, norm=None): float32 [n_mels, 1+n_fft//2]."""
    n_bins = 1 + n_fft // 2
    fftfreqs = np.linspace(0, float(sr) / 2, n_bins, endpoint=True)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    weights = np.zeros((n_mels, n_bins), dtype=np.float32)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    return weights


def n_frames_for(n_samples, hop=HOP_LENGTH):
    """librosa.stft(center=True): 1 + len(y) // hop frames."""
    return 1 + n_samples // hop


def stft_magnitude(y, n_fft=N_WINDOW, hop=HOP_LENGTH):
    """|librosa.stft(y, n_fft, hop, window=np.hamming(n_fft), center=True, 'reflect')|.

    Returns float64 [1 + n_fft//2, n_frames]."""
    y = np.asarray(y, dtype=np.float64)
    win = np.hamming(n_fft)  # symmetric, DatasetDcase2019Task4.py:209
    yp = np.pad(y, n_fft // 2, mode="reflect")
    n_frames = n_frames_for(len(y), hop)
    idx = np.arange(n_fft)[None, :] + hop * np.arange(n_frames)[:, None]
    frames = yp[idx] * win[None, :]
    return np.abs(np.fft.rfft(frames, axis=1)).T


def calculate_mel_spec(y, mel_basis=None):
    """DatasetDcase2019Task4.calculate_mel_spec with save_log_feature=False
    (main.py:201): float32 [n_frames, 64] amplitude mel."""
    if mel_basis is None:
        mel_basis = mel_filterbank()
    S = stft_magnitude(y)
    mel = np.dot(mel_basis, S)  # float32 @ float64 -> float64
    return mel.T.astype(np.float32)


def amplitude_to_db(S, amin=1e-5, top_db=80.0):
    """librosa.amplitude_to_db(S, ref=1.0): 10*log10(max(amin^2, S^2)), then
    floor at (clip max - top_db).  dtype follows the input (f32 stays f32)."""
    S = np.asarray(S)
    magnitude = np.abs(S)
    power = np.square(magnitude, out=magnitude)
    log_spec = 10.0 * np.log10(np.maximum(amin ** 2, power))
    log_spec -= 10.0 * np.log10(np.maximum(amin ** 2, 1.0))
    if top_db is not None:
        log_spec = np.maximum(log_spec, log_spec.max() - top_db)
    return log_spec


def pad_trunc_seq(x, max_len):
    """DataLoad.py:210-229 (pads with 0.0, i.e. 0 dB rows AFTER the log)."""
    if len(x) < max_len:
        pad = np.zeros((max_len - len(x),) + x.shape[1:])
        return np.concatenate((x, pad), axis=0)
    return x[:max_len]


def scaler_std(mean_, mean_of_square_):
    """Scaler.py:31-32,89-97."""
    return np.sqrt(np.asarray(mean_of_square_) - np.asarray(mean_) ** 2)


def transform_chain(mel_amp, mean_, std_, noise=None, frames=MAX_FRAMES):
    """get_transforms(frames, scaler, augment_type='noise' if noise is given).

    mel_amp : float32 [T, 64] amplitude mel (the cached .npy feature)
    noise   : float64 [T, 64] = |N(0, 0.25)| sample (AugmentGaussianNoise draws it
              from numpy's global MT19937; tests inject it)
    returns : list of float32 [1, frames, 64] arrays: [clean] or [clean, noisy]
    """
    streams = [np.asarray(mel_amp)]
    if noise is not None:
        streams.append(streams[0] + noise)  # f32 + f64 -> f64, DataLoad.py:285
    out = []
    for s in streams:
        L = amplitude_to_db(s.T).T                      # ApplyLog
        L = pad_trunc_seq(L, frames)                    # PadOrTrunc
        L = L.astype(np.float32)[None]                  # ToTensor(.float(), unsqueeze 0)
        if mean_ is not None:
            L = ((L - mean_) / std_).astype(np.float32)  # Normalize -> torch.Tensor
        out.append(L)
    return out


def scaler_means(features):
    """Scaler.means (Scaler.py:34-87) over an iterable of [1,T,64] arrays:
    per-mel-bin mean and mean-of-square, averaged per sample then over samples."""
    m = None
    m2 = None
    n = 0
    for x in features:
        x = np.asarray(x)
        a = x
        while a.ndim != 1:
            a = np.mean(a, axis=0, dtype=np.float64)
        b = x ** 2
        while b.ndim != 1:
            b = np.mean(b, axis=0, dtype=np.float64)
        m = a if m is None else m + a
        m2 = b if m2 is None else m2 + b
        n += 1
    return m / n, m2 / n
