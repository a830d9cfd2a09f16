#!/usr/bin/env python
"""DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) of every kernel of one training step, from a
per-kernel ncu summary csv (tools/ncu_summary.py output of an `ncu --set full` capture of tools/profile_step.py), keyed by
the names bench.py's live per-kernel timing uses (DCASE_PROF names).  bench.py reads the JSON at run time for
`roofline.traffic`.

    python tools/ncu_traffic.py profiles/r2_ncu_summary_final.csv > profiles/ncu_traffic.json
"""
import collections
import csv
import json
import sys


def prof_name(kernel, grid, order):
    """ncu kernel name (+ launch order within the step) -> DCASE_PROF name."""
    k = kernel
    if k.startswith("stft_mel"):
        return "stft_mel"
    if k.startswith("cnn0_fwd"):
        return "cnn0_fused_fwd"
    if k.startswith("cnn0_bwd_kernel"):
        return "cnn0_fused_bwd"
    if k.startswith("conv3x3_tma_kernel<"):          # <PITCH, HALF>: HALF = the fp16 forward pass, else the data gradient
        args = k[k.index("<") + 1:k.index(">")].replace("(int)", "").replace("(bool)", "").replace(" ", "").split(",")
        layer = "l1" if args[0] == "10" else "l2"
        if len(args) > 1:
            return ("conv3x3_fwd_" if args[1] in ("1", "true") else "conv3x3_dgrad_") + layer
        key = "conv10" if args[0] == "10" else "conv8"
        return ("conv3x3_fwd_" if order[key] <= 2 else "conv3x3_dgrad_") + layer
    if k.startswith("conv_wgrad_tma_kernel<10>"):
        return "conv3x3_wgrad_l1"
    if k.startswith("conv_wgrad_tma_kernel<8>"):
        return "conv3x3_wgrad_l2"
    if k.startswith("glu_pool_fwd"):
        return "glu_pool_fwd_l1" if grid.startswith("(296") else "glu_pool_fwd_l2"
    if k.startswith("glu_pool_bwd"):
        return "glu_pool_bwd_l2" if order["glu_bwd"] == 1 else "glu_pool_bwd_l1"
    return {"gru_fwd_kernel": "gru_fwd", "gru_bwd_kernel": "gru_bwd", "sgemm_batch_kernel": "sgemm",
            "head_fwd_kernel": "head_fwd", "head_bwd_kernel": "head_bwd", "mt_loss_kernel": "mt_loss",
            "adam_ema_kernel": "adam_ema", "bn_stats_kernel": "bn_stats", "bn_finalize_kernel": "bn_finalize",
            "cnn0_moments_kernel": "cnn0_moments", "bn0_finalize_kernel": "bn0_finalize",
            "bn_bwd_apply_kernel": "bn_bwd_apply", "finish_kernel": "logmel_finish", "clip_max_kernel": "logmel_finish",
            "colsum_batch_kernel": "colsum", "conv_w_image_kernel": "conv_w_prep",
            "cnn0_bwd_finalize_kernel": "cnn0_bwd_finalize", "gemm_tc_kernel": "gemm_tc",
            "syncbn_allreduce_kernel": "syncbn_allreduce"}.get(k.split("<")[0], k)


def main(path):
    order = collections.Counter()
    out = collections.OrderedDict()
    for r in csv.DictReader(open(path)):
        k = r["kernel"]
        if k.startswith("conv3x3_tma_kernel<10>"):
            order["conv10"] += 1
        if k.startswith("conv3x3_tma_kernel<8>"):
            order["conv8"] += 1
        if k.startswith("glu_pool_bwd"):
            order["glu_bwd"] += 1
        name = prof_name(k, r["grid"], order)
        b = (float(r["rd_MB"]) + float(r["wr_MB"])) * 1e6
        e = out.setdefault(name, {"bytes_per_launch": 0.0, "launches": 0})
        e["bytes_per_launch"] = max(e["bytes_per_launch"], b)      # several launches of one name: the largest
        e["launches"] += 1
    print(json.dumps({"source": path, "batch": 24, "kernels": out}, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])
