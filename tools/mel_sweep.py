#!/usr/bin/env python
"""BASELINE.json configs[3]: mel-feature-only sweep -- N synthetic 10-s clips through the fused STFT + mel kernel
(and the dB / normalise finish), HBM GB/s against the measured copy bandwidth.  Algorithmic bytes per clip:
441000 * 4 (f32 waveform read) + 864 * 64 * 4 (f32 mel write) = 1,985,184 B."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from dcase2019_task4_b200 import kernels as K  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=10000)
    ap.add_argument("--chunk", type=int, default=500)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    peaks = bench.load_peaks()
    waves, _ = bench.synthetic_batches(2, seed=5)                       # 48 distinct clips, tiled to the sweep size
    base = torch.from_numpy(waves.reshape(-1, bench.N_SAMPLES)).to(dev)
    n = args.clips
    big = base.repeat((n + base.shape[0] - 1) // base.shape[0], 1)[:n].contiguous()   # n x 441000 f32 (17.6 GB at 10k)
    mean = torch.full((64,), -30.0, device=dev)
    std = torch.full((64,), 12.0, device=dev)
    out = {}
    sampler = bench.ClockSampler(0)
    for name, fn in (("stft_mel", lambda w: K.logmel_fwd(w)),
                     ("stft_mel+finish", lambda w: K.logmel_finish(K.logmel_fwd(w), mean, std, 864))):
        best = None
        for _ in range(args.reps + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for i in range(0, n, args.chunk):
                fn(big[i:i + args.chunk])
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        gbs = n * bench.MEL_BYTES_PER_CLIP / (best * 1e-3) / 1e9
        out[name] = {"ms": best, "clips_per_s": n / (best * 1e-3), "achieved_GBs": gbs, "frac_of_hbm": gbs / peaks["hbm_gbs"]}
    print(json.dumps({"config": "mel-feature-only sweep (BASELINE.json configs[3])", "clips": n, "dtype": "f32",
                      "peak_hbm_GBs": peaks["hbm_gbs"], "peak_source": peaks["source"], "clocks": sampler.stop(),
                      "algorithmic_bytes_per_clip": bench.MEL_BYTES_PER_CLIP, "results": out}))


if __name__ == "__main__":
    main()
