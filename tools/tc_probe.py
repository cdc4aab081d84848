#!/usr/bin/env python
"""Layer-by-layer comparison of the tcgen05 conv path (bf16) with the CUDA-core bf16 path on the
same input — the first thing to run on a B200 after touching conv_tc.cuh."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adipose_unet_b200 as A
from adipose_unet_b200 import api

NAMES = ["down1_conv2", "pool1", "down2_conv2", "pool2", "down3_conv2", "pool3", "dilate1", "dilate2", "dilate3",
         "dilate4", "dilate5", "dilate6", "dilate_add", "up3_conv1", "up3_conv2", "up3_conv3", "up2_conv1",
         "up2_conv2", "up2_conv3", "up1_conv1", "up1_conv2", "up1_conv3", "prob"]


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    w = A.synth.init_weights()
    tiles = A.synth.ecm_tiles(n, S, seed=5)
    outs = {}
    for prec in ("bf16_simt", "bf16"):
        m = api.AdiposeUNet(precision=prec, max_forwards=max(8, n))
        m.build_model()
        m.set_weights(w)
        t0 = time.time()
        p = m.predict_batch(tiles, 127.5, 50.0)
        dt = time.time() - t0
        outs[prec] = {k: m.engine.debug_layer(k, n - 1) for k in NAMES}
        print(f"{prec}: predict ok in {dt * 1e3:.1f} ms, prob range [{p.min():.4f}, {p.max():.4f}]", flush=True)
    print(f"{'layer':14s} {'max|ref|':>10s} {'max|diff|':>10s} {'rel':>9s}  first bad (y,x,c)")
    for k in NAMES:
        a, b = outs["bf16_simt"][k], outs["bf16"][k]
        d = np.abs(a - b)
        rel = d.max() / max(np.abs(a).max(), 1e-9)
        bad = ""
        if rel > 2e-2:
            idx = np.argwhere(d > 2e-2 * np.abs(a).max())
            bad = f"{tuple(idx[0])} of {len(idx)} / {d.size}; got {b[tuple(idx[0])]:.4f} want {a[tuple(idx[0])]:.4f}"
        print(f"{k:14s} {np.abs(a).max():10.4f} {d.max():10.4f} {rel:9.2e}  {bad}")


if __name__ == "__main__":
    main()
