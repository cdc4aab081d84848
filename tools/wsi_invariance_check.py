#!/usr/bin/env python
"""Whole-slide masks must not depend on the number of strips: emulate world sizes 1, 2, 4, 8 one after another on ONE GPU at
the real tile size (1024) and row count of configs[2] (63 tile rows at 50 % overlap) on a narrow RGB slide, and compare
probabilities / masks / counts bit for bit.  (tests/test_gpu_wsi.py does this at tile size 128.)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adipose_unet_b200 as A
from adipose_unet_b200 import api, wsi as W

T = 1024


class LocalDist:
    def __init__(self):
        self.q = {}

    def send(self, t, dst):
        self.q.setdefault(dst, []).append(t.clone())

    def recv(self, t, src):
        t.copy_(self.q[self.me].pop(0))

    def get_backend(self):
        return "local"


def main():
    H = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
    Wd = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    overlap = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
    channels = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    tta = sys.argv[5] if len(sys.argv) > 5 else "full"
    blocks = {}
    for key in ((0, 0), (0, 1), (1, 0), (1, 1)):
        if channels == 3:
            blocks[key] = np.stack([A.synth.slide_block(key[0], key[1] + 2 * c, T) for c in range(3)], axis=-1)
        else:
            blocks[key] = A.synth.slide_block(*key, T)
    slide = np.empty((H, Wd) + ((3,) if channels == 3 else ()), np.uint8)
    for by in range(H // T):
        for bx in range(Wd // T):
            slide[by * T:(by + 1) * T, bx * T:(bx + 1) * T] = blocks[(by % 2, bx % 2)]
    gray = A.synth.rgb_to_gray_u8(slide) if channels == 3 else slide
    gt = (gray > 160).astype(np.uint8)
    eng = api.Engine(precision="bf16", max_forwards=16)
    eng.set_weights(A.synth.init_weights())
    win = api.GaussianBlender(T, engine=eng).weight_map
    ref = None
    for world in (1, 2, 4, 8):
        d = LocalDist()
        prob = np.zeros((H, Wd), np.float32); mask = np.zeros((H, Wd), np.uint8); counts = np.zeros(4, np.int64)
        for rank in range(world):
            d.me = rank
            r = W.reconstruct_wsi(eng, lambda y0, n: slide[y0:y0 + n], H, Wd, tile=T, overlap=overlap, blend_mode="gaussian", window=win,
                                  mean=A.synth.DEFAULT_MEAN, std=A.synth.DEFAULT_STD, tta_mode=None if tta == "none" else tta,
                                  gt_rows=lambda y0, n: gt[y0:y0 + n], rank=rank, world=world, dist=d,
                                  to_device=lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda())
            lo, hi = r["own"]
            if hi > lo:
                prob[lo:hi] = r["prob"]; mask[lo:hi] = r["mask"]; counts += np.array(r["counts"])
        if ref is None:
            ref = (prob, mask, counts)
        bad = np.argwhere(prob != ref[0])
        print(f"world {world}: counts {counts.tolist()} prob identical {bad.size == 0} mask diff px {int((mask != ref[1]).sum())}"
              + (f" first differing rows {sorted(set(bad[:, 0].tolist()))[:6]} ... n={len(bad)} max|d|={np.abs(prob - ref[0]).max():.3e}" if bad.size else ""),
              flush=True)


if __name__ == "__main__":
    main()
