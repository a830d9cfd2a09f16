#!/usr/bin/env python
"""HBM roofline of the Scaler reduction (dcase_scaler_accumulate) and of read_audio's
mix-down (dcase_audio_mixdown, rank 3) on one B200.

Scaler: batches of 64 (what Scaler.means launches, 14 MB) and 256 clips x 864 x 64 amplitude mels, rotating over a pool larger
than L2 so each launch's first pass (clip maxima) reads HBM; the second pass (dB + reduction) re-reads the batch from
L2.  Algorithmic bytes per clip = 864 * 64 * 4 = 221,184 B (one read; the 1 KB of float64 sums is noise).
Mix-down: stereo 16-bit PCM 10-s clips: 441000 * (2 * 2 read + 4 written) = 3,528,000 B per clip."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from dcase2019_task4_b200 import kernels as K  # noqa: E402


def timed(fn, n_pool, reps=5):
    best = None
    for _ in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for i in range(n_pool):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return best


def main():
    dev = torch.device("cuda", 0)
    peaks = bench.load_peaks()
    T = 864
    out = {}
    sums = torch.zeros(2, 64, dtype=torch.float64, device=dev)
    for B, n_pool in ((64, 24), (256, 8)):                                # 340 / 453 MB pools > 126 MB L2
        pool = (torch.rand(n_pool, B, T, 64, device=dev) * 40.0).contiguous()
        for name, log in (("scaler_accumulate(apply_log=1, B=%d)" % B, True),
                          ("scaler_accumulate(apply_log=0, B=%d)" % B, False)):
            ms = timed(lambda i: K.scaler_accumulate(pool[i], sums, frames=T, apply_log=log), n_pool)
            gbs = n_pool * B * T * 64 * 4 / (ms * 1e-3) / 1e9
            out[name] = {"ms_per_launch": ms / n_pool, "clips_per_s": n_pool * B / (ms * 1e-3), "achieved_GBs": gbs,
                         "frac_of_hbm": gbs / peaks["hbm_gbs"]}
        del pool
    n = 441000
    pcm = torch.randint(-32768, 32767, (48, n, 2), dtype=torch.int16, device=dev)   # 48 x 1.76 MB in, 85 + 85 MB total
    ms = timed(lambda i: K.audio_mixdown(pcm[i]), 48)
    gbs = 48 * n * 8 / (ms * 1e-3) / 1e9
    out["audio_mixdown(stereo pcm16)"] = {"ms_per_launch": ms / 48, "clips_per_s": 48 / (ms * 1e-3), "achieved_GBs": gbs,
                                          "frac_of_hbm": gbs / peaks["hbm_gbs"]}
    res = {"config": "Scaler reduction / mix-down roofline", "peak_hbm_GBs": peaks["hbm_gbs"],
           "peak_source": peaks["source"], "results": out}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "scaler_bench.json"), "w") as fh:
        json.dump(res, fh)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
