#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/wsi_nccl_check.py [size] [overlap] [channels]
Whole-slide reconstruction over N real GPUs (NCCL boundary exchange) against the same slide on rank 0 alone: per-row mask
sums and the rows where they differ (which strip / zone they fall into)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adipose_unet_b200 as A
from adipose_unet_b200 import api, wsi as W

T = 1024


def main():
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
    overlap = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
    channels = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("cpu:gloo,cuda:nccl", device_id=torch.device("cuda", local))
    eng = api.Engine(precision="bf16", device=local, max_forwards=16)
    eng.set_weights(A.synth.init_weights())
    win = api.GaussianBlender(T, engine=eng).weight_map
    blocks = {}
    for key in ((0, 0), (0, 1), (1, 0), (1, 1)):
        blocks[key] = np.stack([A.synth.slide_block(key[0], key[1] + 2 * c, T) for c in range(3)], axis=-1) if channels == 3 \
            else A.synth.slide_block(*key, T)

    def rows_of(y0, rows):
        out = np.empty((rows, size) + ((3,) if channels == 3 else ()), np.uint8)
        for by in range(y0 // T, (y0 + rows - 1) // T + 1):
            a, b = max(by * T, y0), min((by + 1) * T, y0 + rows)
            for bx in range(size // T):
                out[a - y0:b - y0, bx * T:(bx + 1) * T] = blocks[(by % 2, bx % 2)][a - by * T:b - by * T]
        return out

    def run(r, w, d):
        res = W.reconstruct_wsi(eng, rows_of, size, size, tile=T, overlap=overlap, blend_mode="gaussian", window=win,
                                mean=A.synth.DEFAULT_MEAN, std=A.synth.DEFAULT_STD, tta_mode="full", rank=r, world=w, dist=d,
                                to_device=lambda a: torch.from_numpy(a).cuda(), want_prob=True, want_mask=True)
        lo, hi = res["own"]
        msum = res["mask"].sum(axis=1, dtype=np.int64) if hi > lo else np.zeros(0, np.int64)
        psum = res["prob"].astype(np.float64).sum(axis=1) if hi > lo else np.zeros(0)
        return lo, hi, msum, psum

    W.warmup_peer_channels(dist, rank, world, local)
    lo, hi, msum, psum = run(rank, world, dist)
    payload = [None] * world
    dist.all_gather_object(payload, (lo, hi, msum, psum))
    if rank == 0:
        full_m = np.zeros(size, np.int64); full_p = np.zeros(size)
        for (a, b, m, p) in payload:
            full_m[a:b] = m; full_p[a:b] = p
        lo1, hi1, m1, p1 = run(0, 1, None)
        bad = np.nonzero(full_m != m1)[0]
        badp = np.nonzero(full_p != p1)[0]
        stride = int(T * (1 - overlap))
        strips = W.plan_strips(size, size, T, stride, world)
        print(f"N={world}: mask rows differing {len(bad)} (net {int((full_m - m1).sum())} px), prob rows differing {len(badp)}")
        for s in strips:
            z = (s.own_lo, s.zone_hi)
            inz = [int(r_) for r_ in badp if s.own_lo <= r_ < s.own_hi]
            print(f"  strip {s.rank}: own [{s.own_lo},{s.own_hi}) zone_hi {s.zone_hi}: differing prob rows {len(inz)}"
                  + (f" first {inz[:3]} last {inz[-3:]}" if inz else ""))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
