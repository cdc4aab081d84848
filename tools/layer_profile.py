#!/usr/bin/env python
"""Per-layer device time of one chunk of forwards (event-bracketed launches)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adipose_unet_b200 as A
from adipose_unet_b200 import api

def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    prec = sys.argv[3] if len(sys.argv) > 3 else "bf16"
    ops = list(range(8)) if (len(sys.argv) > 4 and sys.argv[4] == "tta") else None   # n tiles -> n/8 tiles x 8 augmentations
    eng = api.Engine(precision=prec, max_forwards=n)
    eng.set_weights(A.synth.init_weights())
    if os.environ.get("ADP_FUSE_FIRST") is not None:      # experiment: first conv inside down1_conv2 on / off
        eng.set_option("fuse_first", int(os.environ["ADP_FUSE_FIRST"]))
    tiles = A.synth.ecm_tiles(min(n, 2), S)
    tiles = np.concatenate([tiles] * (n // len(tiles)))[:n]
    if ops:
        tiles = tiles[: max(1, n // 8)]
    import torch
    td = torch.from_numpy(tiles).cuda()
    out = torch.empty((len(tiles), S, S), dtype=torch.float32, device="cuda")
    for _ in range(2):
        eng.predict(td, 127.5, 50.0, ops, out)
    eng.profile(True)
    eng.predict(td, 127.5, 50.0, ops, out)
    rows = eng.profile_rows()
    eng.profile(False)
    tot = sum(r["ms"] for r in rows)
    print(f"{prec} S={S} forwards={n}: total {tot:.3f} ms -> {tot / n:.3f} ms/forward")
    print(f"{'kernel':36s} {'ms':>9s} {'share':>7s} {'TFLOP/s':>9s} {'GB/s':>9s}")
    order = ["first_conv"] + ["conv3x3_tcgen05/" + l for l in A.layers.LAYER_NAMES] + ["conv3x3_simt_fp32/" + l for l in A.layers.LAYER_NAMES]
    rows.sort(key=lambda r: order.index(r["name"]) if r["name"] in order else 999)
    for r in rows:
        tf = r["flops"] / (r["ms"] * 1e-3) / 1e12 if r["flops"] and r["ms"] > 0 else 0
        gb = r["bytes"] / (r["ms"] * 1e-3) / 1e9 if r["bytes"] and r["ms"] > 0 else 0
        print(f"{r['name']:36s} {r['ms']:9.3f} {r['ms'] / tot:7.2%} {tf:9.1f} {gb:9.1f}")

if __name__ == "__main__":
    main()
