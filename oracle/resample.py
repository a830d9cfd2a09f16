"""numpy restatement of ``librosa.resample(audio, orig_sr, target_sr)`` as the reference calls it (TEST ORACLE).

Follows ``baseline/utils/utils.py:190-192``: ``librosa.resample`` with its default ``res_type='kaiser_best'`` of the librosa
versions the baseline was written for (< 0.10; ``environment.yml:17`` leaves librosa unpinned), i.e. resampy's
band-limited sinc interpolation (Smith, "Digital Audio Resampling Home Page"; resampy 0.2, ``resampy/core.py`` /
``resampy/interpn.py`` / ``resampy/filters.py``), followed by librosa's ``fix_length`` to ``ceil(n * ratio)`` samples.

resampy and librosa are NOT installed here (un-vendored third-party dependencies): PARITY UNPINNED against them.  The
published algorithm is restated:

* filter ``kaiser_best``: half of a Kaiser-windowed sinc, 64 zero crossings, 512 table samples per crossing, Kaiser
  beta = 14.769656459379492, roll-off 0.9475937167399596 x Nyquist (the constants resampy documents for this filter);
  scaled by the sample ratio when down-sampling;
* every output sample t at input time ``t / ratio`` sums the input samples left and right of it, weighted by the
  table linearly interpolated (``interp_win + eta * interp_delta``) at steps of ``int(scale * 512)`` table entries.
"""
import numpy as np

NUM_ZEROS = 64
NUM_TABLE = 512            # 2 ** precision, precision = 9
KAISER_BETA = 14.769656459379492
ROLLOFF = 0.9475937167399596


def kaiser_best_half_window():
    """resampy.filters.sinc_window(num_zeros=64, precision=9, window=kaiser(beta), rolloff): float64 [32769]."""
    n = NUM_TABLE * NUM_ZEROS
    sinc_win = ROLLOFF * np.sinc(ROLLOFF * np.linspace(0, NUM_ZEROS, num=n + 1, endpoint=True))
    taper = np.kaiser(2 * n + 1, KAISER_BETA)[n:]
    return taper * sinc_win


def resample(x, orig_sr, target_sr):
    """librosa.resample(x, orig_sr, target_sr) for a mono float signal -> array of ceil(len(x) * ratio) samples."""
    x = np.asarray(x, dtype=np.float64)
    ratio = float(target_sr) / float(orig_sr)
    n_out = int(len(x) * ratio)                          # resampy.resample's output length
    interp_win = kaiser_best_half_window()
    if ratio < 1:
        interp_win = interp_win * ratio
    interp_delta = np.zeros_like(interp_win)
    interp_delta[:-1] = np.diff(interp_win)
    scale = min(1.0, ratio)
    index_step = int(scale * NUM_TABLE)
    nwin, n_orig = len(interp_win), len(x)
    # resampy's clock: time_register = 0.0, then += 1 / ratio after every output sample (a serial float64 sum; its rounding
    # decides on which side of an integer the register falls, where the truncated table step makes the result jump)
    time_register = np.concatenate([[0.0], np.cumsum(np.full(max(n_out - 1, 0), 1.0 / ratio))])[:n_out]
    n = time_register.astype(np.int64)
    y = np.zeros(n_out)
    # left wing (samples n, n - 1, ...)
    frac = scale * (time_register - n)
    index_frac = frac * NUM_TABLE
    offset = index_frac.astype(np.int64)
    eta = index_frac - offset
    i_max = np.minimum(n + 1, (nwin - offset) // index_step)
    for i in range(int(i_max.max()) if n_out else 0):
        live = i < i_max
        idx = np.where(live, offset + i * index_step, 0)
        w = interp_win[idx] + eta * interp_delta[idx]
        y += np.where(live, w * x[np.where(live, n - i, 0)], 0.0)
    # right wing (samples n + 1, n + 2, ...)
    frac = scale - frac
    index_frac = frac * NUM_TABLE
    offset = index_frac.astype(np.int64)
    eta = index_frac - offset
    k_max = np.minimum(n_orig - n - 1, (nwin - offset) // index_step)
    for k in range(int(k_max.max()) if n_out else 0):
        live = k < k_max
        idx = np.where(live, offset + k * index_step, 0)
        w = interp_win[idx] + eta * interp_delta[idx]
        y += np.where(live, w * x[np.where(live, n + k + 1, 0)], 0.0)
    n_fix = int(np.ceil(len(x) * ratio))                 # librosa: util.fix_length(y_hat, ceil(n * ratio))
    if len(y) < n_fix:
        y = np.concatenate([y, np.zeros(n_fix - len(y))])
    return y[:n_fix]
