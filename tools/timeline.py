#!/usr/bin/env python
"""Timeline of one eager mean-teacher iteration at the bench workload with the streams overlapping as in production:
every launch's (stream, start, end) from CUDA events (dcase_profile_timeline_*), printed in start order with a crude
occupancy lane per stream.  `python tools/timeline.py [--pipelined]`."""
import argparse
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from dcase2019_task4_b200 import _lib, config as cfg  # noqa: E402
from dcase2019_task4_b200.main import MeanTeacherEngine  # noqa: E402
from dcase2019_task4_b200.models.CRNN import CRNN  # noqa: E402
from dcase2019_task4_b200.utils.utils import weights_init  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pipelined", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    waves, targets = bench.synthetic_batches(2, seed=1)
    wave_dev = torch.from_numpy(waves).to(dev)
    target_dev = torch.from_numpy(targets).to(dev)
    mean = torch.full((64,), -30.0, device=dev)
    std = torch.full((64,), 12.0, device=dev)
    torch.manual_seed(0)
    crnn, crnn_ema = CRNN(**cfg.crnn_kwargs), CRNN(**cfg.crnn_kwargs)
    crnn.apply(weights_init)
    crnn_ema.apply(weights_init)
    for p in crnn_ema.parameters():
        p.detach_()
    crnn, crnn_ema = crnn.train().cuda(), crnn_ema.train().cuda()
    opt = torch.optim.Adam(crnn.parameters(), lr=0.001, betas=(0.9, 0.999))
    eng = MeanTeacherEngine(crnn, opt, crnn_ema, slice(6), slice(18, 24), 24, 864, use_graph=False)
    if args.pipelined:
        eng.prime_features(wave_dev[0], mean, std)

    def step(i):
        if args.pipelined:
            eng.step_pipelined(wave_dev[(i + 1) % 2], target_dev[i % 2], mean, std, 0.1, i + 1, check=False)
        else:
            eng.step_from_waveforms(wave_dev[i % 2], target_dev[i % 2], mean, std, 0.1, i + 1, check=False)

    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    L = _lib.lib()
    _lib.check(L.dcase_profile_timeline_begin())
    step(3)
    buf = ctypes.create_string_buffer(1 << 16)
    _lib.check(L.dcase_profile_timeline_end(buf, len(buf)))
    rows = [r.split(",") for r in buf.value.decode().strip().splitlines()]
    rows = [(n, int(s), float(a), float(b)) for n, s, a, b in rows]
    rows.sort(key=lambda r: r[2])
    end = max(r[3] for r in rows)
    print(f"{len(rows)} launches, {end:.1f} us from the first launch to the last completion")
    for n, s, a, b in rows:
        print(f"{a:8.1f} {b:8.1f} {b - a:7.1f}  s{s}  {'    ' * s}{n}")


if __name__ == "__main__":
    main()
