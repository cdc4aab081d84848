#!/usr/bin/env python
"""Would a kind::tf32 tensor-core path meet the <= 1e-4 probability bound of BASELINE.json?  

CPU simulation of the numerics such a kernel would have: tcgen05.mma kind::tf32 reads fp32 operands and uses their top
19 bits (sign, 8 exponent, 10 mantissa bits); products are exact, accumulation is fp32.  Activations would stay fp32 in HBM
(so only the operand truncation differs from the exact path).  The oracle graph (oracle/unet.py) is run with every conv's
input and weight tensor truncated (RZ, what the hardware does) or rounded (RN, the best a pre-rounding epilogue could do) to
TF32, and compared with the float64 oracle.  bf16 and the bf16x3 split are simulated the same way for scale.

  python tools/tf32_error_sim.py [size]        # CPU only; prints max-abs probability error per format
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adipose_unet_b200 as A
from oracle import unet as U


def to_tf32(t: torch.Tensor, mode: str) -> torch.Tensor:
    i = t.contiguous().view(torch.int32)
    if mode == "rz":
        i = i & ~0x1FFF
    else:                                   # round to nearest even on the 13 dropped bits
        i = (i + 0xFFF + ((i >> 13) & 1)) & ~0x1FFF
    return i.view(torch.float32)


def quantiser(fmt):
    if fmt == "tf32_rz":
        return lambda t: to_tf32(t, "rz")
    if fmt == "tf32_rn":
        return lambda t: to_tf32(t, "rn")
    if fmt == "bf16":
        return lambda t: t.to(torch.bfloat16).to(torch.float32)
    return lambda t: t


def run(fmt, x, params):
    q = quantiser(fmt)
    orig = U._conv

    def conv(xx, p, name, act=True):
        w, b = p[name]
        d = U.DILATION[name]
        pad = d * (w.shape[-1] - 1) // 2
        if fmt == "bf16x3":                # hi/lo split of both operands, three products (lo*lo dropped)
            ah = xx.to(torch.bfloat16).to(torch.float32); al = (xx - ah).to(torch.bfloat16).to(torch.float32)
            wh = w.to(torch.bfloat16).to(torch.float32); wl = (w - wh).to(torch.bfloat16).to(torch.float32)
            y = F.conv2d(ah, wh, b, padding=pad, dilation=d) + F.conv2d(ah, wl, None, padding=pad, dilation=d) + \
                F.conv2d(al, wh, None, padding=pad, dilation=d)
        else:
            y = F.conv2d(q(xx), q(w), b, padding=pad, dilation=d)
        return F.relu(y) if act else y

    U._conv = conv
    try:
        with torch.no_grad():
            return U.forward(x, params).numpy()
    finally:
        U._conv = orig


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    torch.set_num_threads(os.cpu_count() or 1)
    torch.backends.mkldnn.enabled = True
    w = A.synth.init_weights()
    tile = A.synth.ecm_tile(S, seed=21).astype(np.float32)
    x = torch.from_numpy((tile - A.synth.DEFAULT_MEAN) / (A.synth.DEFAULT_STD + 1e-10)).float().unsqueeze(0)
    with torch.no_grad():
        ref = U.forward(x.double(), U.to_torch_params(w, torch.float64)).numpy()
    p32 = U.to_torch_params(w)
    print(f"max-abs probability error vs the float64 oracle, one {S}x{S} tile, random-init weights (seed 865):")
    for fmt in ("fp32", "tf32_rn", "tf32_rz", "bf16", "bf16x3"):
        out = run(fmt, x, p32)
        err = float(np.abs(out - ref).max())
        print(f"  {fmt:8s} {err:.3e}   {'meets' if err <= 1e-4 else 'MISSES'} the 1e-4 bound")


if __name__ == "__main__":
    main()
