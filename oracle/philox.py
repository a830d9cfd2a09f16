"""numpy restatement of the device RNG contract (TEST ORACLE).

The reference draws dropout masks from torch's global generator and the teacher
noise from numpy's global MT19937 (``DataLoad.py:285``); neither stream can be
reproduced on a GPU, so the CUDA path defines its own counter-based contract
(``include/dcase_b200.h``, "RNG contract") and this file restates it so tests can
inject the very same masks / noise into the torch oracle.

Philox4x32-10 (Salmon et al., SC'11), key = (seed_lo, seed_hi),
counter = (row_lo, row_hi, stream, step).

* dropout keep-bit of element (row, col): bit ``col & 31`` of output word ``col >> 5``
  (p = 0.5 exactly; 64-wide tensors use words 0-1, the 128-wide head words 0-3).
* teacher noise for 4 consecutive mel bins of row ``r`` starting at ``4*q``:
  counter row = ``r * 16 + q``; u_i = (w_i + 0.5) * 2^-32;
  n0, n1 = BoxMuller(u0, u1); n2, n3 = BoxMuller(u2, u3); noise = 0.25 * |n|.
"""
import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)

STREAM_CNN = (0, 1, 2)      # + 8 * model  (model 0 = student, 1 = teacher)
STREAM_HEAD = 3
STREAM_NOISE = 4


def philox4x32(c0, c1, c2, c3, k0, k1, rounds=10):
    """Vectorised over uint32 arrays (broadcast). Returns 4 uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) & _MASK for c in np.broadcast_arrays(c0, c1, c2, c3)]
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for r in range(rounds):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return [c.astype(np.uint32) for c in (c0, c1, c2, c3)]


def dropout_mask(n_rows, n_cols, seed, stream, step):
    """bool [n_rows, n_cols] keep-mask (n_cols in {64, 128})."""
    rows = np.arange(n_rows, dtype=np.uint64)
    w = philox4x32(rows & _MASK, rows >> np.uint64(32), stream, step, seed & 0xFFFFFFFF, seed >> 32)
    cols = np.arange(n_cols)
    words = np.stack(w, axis=1)[:, cols >> 5]                    # [rows, cols]
    return ((words >> (cols & 31).astype(np.uint32)) & 1).astype(bool)


def teacher_noise(n_rows, seed, step, n_cols=64):
    """float32 [n_rows, 64] = 0.25 * |N(0,1)| (AugmentGaussianNoise std = 0.5 ** 2)."""
    q = n_cols // 4
    ctr = np.arange(n_rows * q, dtype=np.uint64)
    w = philox4x32(ctr & _MASK, ctr >> np.uint64(32), STREAM_NOISE, step, seed & 0xFFFFFFFF, seed >> 32)
    u = [(x.astype(np.float64) + 0.5) * 2.0 ** -32 for x in w]
    out = np.empty((n_rows * q, 4), dtype=np.float64)
    for j in (0, 2):
        r = np.sqrt(-2.0 * np.log(u[j]))
        th = 2.0 * np.pi * u[j + 1]
        out[:, j] = r * np.cos(th)
        out[:, j + 1] = r * np.sin(th)
    return (0.25 * np.abs(out)).reshape(n_rows, n_cols).astype(np.float32)
