#!/usr/bin/env python
"""Per-layer operand-precision study: frame-posterior L-inf of the CRNN at the BASELINE size
(B = 24 x 864 frames, train mode with dropout, the masks of the device RNG contract) when the GEMM operands of selected
layers are rounded the way a tensor-core MMA kind would see them, with fp32 accumulation everywhere:

    tf32      10-bit mantissa, truncated (what tcgen05.mma kind::tf32 does to fp32 operands)
    bf16      8-bit mantissa, round to nearest even (kind::f16 with bf16 operands)
    bf16x2    hi + lo split of BOTH operands, products hi*hi + hi*lo + lo*hi (3 MMAs)
    bf16x2a   hi + lo split of the activation operand only, weights single bf16 (2 MMAs)

It is an EMULATION on the CPU oracle (oracle/crnn.py with operand rounding hooks), i.e. a measurement of the numerics a
kernel of that mode would have, not of its speed.  Output: profiles/r2_precision_modes.json."""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import crnn as ocrnn  # noqa: E402
from tests import helpers as H  # noqa: E402


def r_tf32(x):
    return (x.view(torch.int32) & -8192).view(torch.float32)          # truncate to 10 mantissa bits


def r_bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


def gemm_like(op, a, b, mode):
    """op(a, b) with the operands rounded per `mode` (op is linear in each argument)."""
    if mode == "fp32":
        return op(a, b)
    if mode == "tf32":
        return op(r_tf32(a), r_tf32(b))
    if mode == "bf16":
        return op(r_bf16(a), r_bf16(b))
    ah, bh = r_bf16(a), r_bf16(b)
    al, bl = r_bf16(a - ah), r_bf16(b - bh)
    if mode == "bf16x2":
        return op(ah, bh) + op(ah, bl) + op(al, bh)
    if mode == "bf16x2a":
        return op(ah, bh) + op(al, bh)
    raise ValueError(mode)


def forward(x, p, masks, modes):
    """oracle.crnn.crnn_forward (training, dropout masks injected) with per-layer operand modes:
    modes = {"conv0", "glu0", "conv1", "glu1", "conv2", "glu2"} -> mode name."""
    h = x
    for i in range(3):
        pre = "cnn.cnn."
        w, b = p[pre + f"conv{i}.weight"], p[pre + f"conv{i}.bias"]
        y = gemm_like(lambda a, ww: F.conv2d(a, ww, None, 1, 1), h, w, modes.get(f"conv{i}", "fp32")) + b[None, :, None, None]
        mean = y.mean(dim=(0, 2, 3))
        var = y.var(dim=(0, 2, 3), unbiased=False)
        y = (y - mean[None, :, None, None]) / torch.sqrt(var[None, :, None, None] + ocrnn.BN_EPS)
        y = y * p[pre + f"batchnorm{i}.weight"][None, :, None, None] + p[pre + f"batchnorm{i}.bias"][None, :, None, None]
        lin = gemm_like(lambda a, ww: F.linear(a, ww), y.permute(0, 2, 3, 1), p[pre + f"glu{i}.linear.weight"],
                        modes.get(f"glu{i}", "fp32")) + p[pre + f"glu{i}.linear.bias"]
        gate = torch.sigmoid(y)
        if modes.get(f"gate{i}") == "tanh.approx":      # sigmoid = 0.5 tanh(y / 2) + 0.5 with tanh.approx.f32 (max rel 2^-11)
            delta = (torch.rand(y.shape, generator=torch.Generator().manual_seed(99 + i)) * 2 - 1) * 2.0 ** -11
            gate = 0.5 * torch.tanh(0.5 * y) * (1 + delta) + 0.5
        z = lin.permute(0, 3, 1, 2) * gate
        z = z * masks[f"cnn{i}"].to(z.dtype) * 2.0
        h = F.avg_pool2d(z, ocrnn.POOL)
    h = h.squeeze(-1).permute(0, 2, 1)
    h = ocrnn.bigru(h, p)
    return ocrnn.head(h, p, True, masks["head"])


def main():
    B, T = (24, 864) if len(sys.argv) < 2 else (int(sys.argv[1]), int(sys.argv[2]))
    torch.manual_seed(0)
    p = ocrnn.init_params(seed=6)
    x = torch.randn(B, 1, T, 64, generator=torch.Generator().manual_seed(123)) * 1.2 + 0.1
    masks = H.oracle_masks(B, T, 0x5EED0000BEEF, 11, 0)
    with torch.no_grad():
        ref_s, ref_w = forward(x, p, masks, {})
        rows = []
        all6 = ["conv0", "glu0", "conv1", "glu1", "conv2", "glu2"]
        cases = [("tf32 everywhere (shipped)", {k: "tf32" for k in all6}),
                 ("bf16 on glu0 + conv1, tf32 elsewhere", dict({k: "tf32" for k in all6}, glu0="bf16", conv1="bf16")),
                 ("bf16 on conv1 only, tf32 elsewhere", dict({k: "tf32" for k in all6}, conv1="bf16")),
                 ("bf16 on glu0 only, tf32 elsewhere", dict({k: "tf32" for k in all6}, glu0="bf16")),
                 ("bf16 everywhere", {k: "bf16" for k in all6}),
                 ("bf16x2a (activation split, 2 MMAs) on glu0 + conv1", dict({k: "tf32" for k in all6}, glu0="bf16x2a", conv1="bf16x2a")),
                 ("bf16x2 (both split, 3 MMAs) on glu0 + conv1", dict({k: "tf32" for k in all6}, glu0="bf16x2", conv1="bf16x2")),
                 ("bf16x2 everywhere", {k: "bf16x2" for k in all6}),
                 ("tf32 everywhere + block-0 gate by tanh.approx.f32 (1 MUFU, rel. error up to 2^-11 modelled as uniform noise)",
                  dict({k: "tf32" for k in all6}, gate0="tanh.approx")),
                 ("fp32 operands + block-0 gate by tanh.approx.f32", {"gate0": "tanh.approx"})]
        if os.environ.get("ONLY_GATE"):
            cases = cases[-2:]
        for name, modes in cases:
            s, w = forward(x, p, masks, modes)
            rows.append({"mode": name, "strong_Linf": float((s - ref_s).abs().max()), "weak_Linf": float((w - ref_w).abs().max())})
            print(rows[-1], flush=True)
    out = {"what": "frame-posterior L-inf vs the fp32 oracle, train mode with dropout, operand rounding emulated on the CPU",
           "B": B, "T": T, "budget": 1e-3, "rows": rows}
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    if not os.environ.get("ONLY_GATE"):
        with open(os.path.join(ROOT, "profiles", "r2_precision_modes.json"), "w") as fh:
            json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
