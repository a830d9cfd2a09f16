#!/usr/bin/env python
"""Cycles per tcgen05.mma kind::tf32 (M x N x 8, operands in shared memory) on this GPU: the table DESIGN.md uses to
bound the conv / GLU kernels (small-N tf32 MMAs are limited by the shared-memory operand fetch, not the math rate)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dcase2019_task4_b200 import _lib  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    lib, ctx = _lib.lib(), _lib.ctx(dev)
    out = torch.zeros(148, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    print("M    N   A-major B-major  cycles/MMA  floor(M*N/256)  operand bytes  bytes/cycle")
    cases = [(128, 256, 0, 0, 1024, 0), (128, 128, 0, 0, 1024, 0), (128, 64, 0, 0, 1024, 0), (128, 32, 0, 0, 1024, 0),
             (128, 16, 0, 0, 1024, 0), (64, 64, 0, 0, 1024, 0), (64, 64, 1, 1, 1024, 0), (64, 16, 1, 0, 1024, 0),
             (128, 16, 1, 1, 1024, 0), (64, 80, 1, 1, 1024, 0), (128, 64, 0, 1, 1024, 0),
             # the conv halo operand: 8-row groups one frame row (PITCH * 128 B) apart, start shifted by a tap offset
             (128, 64, 0, 0, 1280, 0), (128, 64, 0, 0, 1280, 128), (128, 64, 0, 0, 1280, 1408), (128, 64, 0, 0, 1024, 128),
             (128, 64, 0, 0, 2048, 0),
             # a tcgen05.commit after every 36 / 8 / 4 MMAs (the kernels commit 2-3 times per tile)
             (128, 64, 0, 0, 1024, 0, 36), (128, 64, 0, 0, 1024, 0, 8), (128, 64, 0, 0, 1024, 0, 4), (64, 16, 1, 0, 1024, 0, 16),
             # the conv kernels' issue loop (36 MMAs with compile-time offsets + commit)
             (128, 64, 0, 0, 1280, 0, -1)]
    for case in cases:
        M, N, am, bm, sbo, shift = case[:6]
        ce = case[6] if len(case) > 6 else 0
        for n_ctas in (148,):
            rc = lib.dcase_bench_umma(ctx, M, N, am, bm, 4608, n_ctas, sbo, shift, ce, out.data_ptr(), stream)
            assert rc == 0, _lib.lib().dcase_last_error()
            torch.cuda.synchronize()
            c = float(out[:n_ctas].mean())
            nbytes = (M + N) * 8 * 4
            print("%-4d %-4d %-7s %-7s %9.1f  %9.1f  %9d  %9.1f   (%d CTAs, A pitch %d shift %d, commit every %d)" % (
                M, N, "MN" if am else "K", "MN" if bm else "K", c, max(M, 128) * N / 256.0, nbytes, nbytes / c, n_ctas, sbo, shift, ce))


if __name__ == "__main__":
    main()
