#!/usr/bin/env python
"""Runs N mean-teacher iterations of the bench workload (batch 24, 10-s clips, waveform in HBM) with no timing
code, as the target command for `ncu` (see /opt/skills/guides/B200_PROFILING.md and profiles/README.md)."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from dcase2019_task4_b200 import config as cfg, kernels as K  # noqa: E402
from dcase2019_task4_b200.main import MeanTeacherEngine  # noqa: E402
from dcase2019_task4_b200.models.CRNN import CRNN  # noqa: E402
from dcase2019_task4_b200.utils.utils import weights_init  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--mel-clips", type=int, default=0, help="also run the mel-only kernel on this many clips")
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
    eng = MeanTeacherEngine(crnn, opt, crnn_ema, slice(6), slice(18, 24), 24, 864,
                            use_graph=False)      # ncu needs the eager launches (it invalidates stream capture)
    for i in range(args.steps):
        eng.step_from_waveforms(wave_dev[i % 2], target_dev[i % 2], mean, std, 0.1, i + 1, check=False)
    if args.mel_clips:
        big = wave_dev.reshape(-1, wave_dev.shape[-1])[: args.mel_clips]
        K.logmel_fwd(big)
    torch.cuda.synchronize()
    print("loss", eng.read_meters()["Loss"])


if __name__ == "__main__":
    main()
