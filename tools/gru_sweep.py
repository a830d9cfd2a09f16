"""BASELINE.json configs[4]: BiGRU sweep on one B200.  nn.GRU-equivalent (input 64, 2 layers, bidirectional),
T = 108 (cfg.max_frames // pooling_time_ratio), batch {24, 256}.

Hidden size 64 (the only one cfg.crnn_kwargs selects) runs through dcase_bigru_forward; every (H, B) point is also
timed through cuDNN (torch.nn.GRU on the same GPU, fp32, TF32 off) as the library baseline SURVEY.md section 8d asks
for.  H = 128 / 256 run through the thread-block-cluster kernel (csrc/gru_cluster.cu).  FLOPs (forward) =
2 B T [(In 3H + H 3H) + (2H 3H + H 3H)] 2.  Writes gpurun_out/gru_sweep.json and prints it.

    python tools/gru_sweep.py [--iters 200]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from dcase2019_task4_b200 import kernels as K  # noqa: E402

T = 108


def flops(B, H, In=64):
    return 2.0 * B * T * ((In * 3 * H + H * 3 * H) + (2 * H * 3 * H + H * 3 * H)) * 2


def time_ms(fn, iters, warmup=20):
    for _ in range(warmup):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=200)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    _sampler = bench.ClockSampler(0)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    rows = []
    for H in (64, 128, 256):
        for B in (24, 256):
            x = torch.randn(B, T, 64, device=dev)
            gru = torch.nn.GRU(64, H, num_layers=2, bidirectional=True, batch_first=True).to(dev).eval()
            with torch.no_grad():
                cudnn_ms = time_ms(lambda: gru(x), args.iters)
            ours_ms = None
            err = None
            if H == 64:
                flat = torch.cat([p.detach().reshape(-1) for _, p in gru.named_parameters()]).contiguous()
                assert flat.numel() == K.GRU_PARAM_COUNT       # nn.GRU's own order is the slab order
                out = torch.empty(B, T, 128, device=dev)
                ws = torch.empty(K.lib().dcase_bigru_workspace_bytes(B, T), dtype=torch.uint8, device=dev)
                ours_ms = time_ms(lambda: K.bigru_forward(x, flat, out=out, ws=ws), args.iters)
                with torch.no_grad():
                    err = float((out - gru(x)[0]).abs().max())
            else:                                                          # cluster BiGRU (csrc/gru_cluster.cu)
                flat = torch.cat([p.detach().reshape(-1) for _, p in gru.named_parameters()]).contiguous()
                out = torch.empty(B, T, 2 * H, device=dev)
                ws = torch.empty(K.lib().dcase_bigru_workspace_bytes_h(B, T, H), dtype=torch.uint8, device=dev)
                ours_ms = time_ms(lambda: K.bigru_forward_h(x, flat, H, out=out, ws=ws), args.iters)
                with torch.no_grad():
                    err = float((out - gru(x)[0]).abs().max())
            f = flops(B, H)
            rows.append({"hidden": H, "batch": B, "seq_len": T, "gflop_fwd": f / 1e9,
                         "ours_ms": ours_ms, "ours_tflops": None if ours_ms is None else f / ours_ms / 1e9,
                         "cudnn_ms": cudnn_ms, "cudnn_tflops": f / cudnn_ms / 1e9,
                         "speedup_vs_cudnn": None if ours_ms is None else cudnn_ms / ours_ms,
                         "max_abs_diff_vs_cudnn": err})
    out = {"config": "BiGRU sweep (BASELINE.json configs[4])", "dtype": "f32", "device": torch.cuda.get_device_name(0),
           "note": "hidden 128 / 256 run through the thread-block-cluster kernel csrc/gru_cluster.cu (cfg.crnn_kwargs selects 64); cuDNN rows are the library baseline",
           "rows": rows}
    out["clocks"] = _sampler.stop()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "gru_sweep.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
