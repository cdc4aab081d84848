#!/usr/bin/env python
"""Whole-slide sliding-window reconstruction at BASELINE.json's full sizes (configs[2]: 32768^2, 50 % overlap;
configs[4]: 16384^2, 75 % overlap), tile-row strips sharded over the ranks (torchrun) or one GPU.

  python tools/wsi_full.py --size 32768 --overlap 0.5 [--tta full|none] [--blend gaussian]
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 tools/wsi_full.py --size 16384 --overlap 0.75

Prints one JSON line (rank 0): tiles, forwards, seconds (max over ranks, wall clock including strip assembly, H2D,
boundary exchange, finalize and mask D2H), Mpx/s, phase breakdown of rank 0, and size-independent checks: confusion counts
sum to H*W, every pixel has coverage > 0 (finite probabilities), mask == (prob > thr) on a sampled band."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import adipose_unet_b200 as A
from adipose_unet_b200 import api, wsi as W

TILE = 1024


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=32768)
    ap.add_argument("--overlap", type=float, default=0.5)
    ap.add_argument("--tta", default="full")
    ap.add_argument("--blend", default="gaussian")
    ap.add_argument("--max-forwards", type=int, default=16)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("cpu:gloo,cuda:nccl", device_id=torch.device("cuda", local))
    eng = api.Engine(precision="bf16", device=local, max_forwards=a.max_forwards)
    eng.set_weights(A.synth.init_weights())
    H = Wd = a.size
    blocks = {}

    def slide_rows(y0, rows):
        out = np.empty((rows, Wd), np.uint8)
        for by in range(y0 // TILE, (y0 + rows - 1) // TILE + 1):
            for bx in range(Wd // TILE):
                key = (by % 4, bx % 4)
                if key not in blocks:
                    blocks[key] = A.synth.slide_block(*key, TILE)
                lo, hi = max(by * TILE, y0), min((by + 1) * TILE, y0 + rows)
                out[lo - y0:hi - y0, bx * TILE:(bx + 1) * TILE] = blocks[key][lo - by * TILE:hi - by * TILE]
        return out

    def gt_rows(y0, rows):
        return (slide_rows(y0, rows) > 160).astype(np.uint8)

    for by in range(4):                       # synthetic slide content is generated before the timed region
        for bx in range(4):
            blocks[(by, bx)] = A.synth.slide_block(by, bx, TILE)
    win = api.GaussianBlender(TILE, engine=eng).weight_map
    # warm-up: one small slide through the same path (allocations, tensor maps)
    W.reconstruct_wsi(eng, lambda y0, r: slide_rows(y0, r)[:, :2048], 2048, 2048, tile=TILE, overlap=0.5, blend_mode=a.blend, window=win,
                      tta_mode=None, rank=0, world=1, to_device=lambda x: torch.from_numpy(x).cuda(), want_prob=False)
    W.warmup_peer_channels(dist, rank, world, local)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    res = W.reconstruct_wsi(eng, slide_rows, H, Wd, tile=TILE, overlap=a.overlap, blend_mode=a.blend, window=win,
                            mean=A.synth.DEFAULT_MEAN, std=A.synth.DEFAULT_STD, tta_mode=None if a.tta == "none" else a.tta,
                            gt_rows=gt_rows, rank=rank, world=world, dist=dist,
                            to_device=lambda x: torch.from_numpy(x).cuda(), want_prob=False, want_mask=True, timings=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
    cnt = torch.tensor(list(res["counts"]), dtype=torch.int64, device="cuda")
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt)
    if rank == 0:
        n_aug = {"none": 1, "minimal": 2, "basic": 4, "full": 8}[a.tta]
        tiles = res["n_tiles_total"]
        secs = float(tt[0])
        counts = [int(c) for c in cnt]
        lo, hi = res["own"]
        out = {"slide": f"{H}x{Wd}", "overlap": a.overlap, "blend": a.blend, "tta": a.tta, "n_gpus": world, "tiles": tiles,
               "forwards": tiles * n_aug, "seconds": secs, "mpx_per_s": H * Wd / 1e6 / secs, "tiles_per_s": tiles / secs,
               "counts_tp_fp_fn_tn": counts, "counts_sum_equals_pixels": sum(counts) == H * Wd,
               "rank0_rows": [lo, hi], "rank0_mask_fraction": float(res["mask"].mean()) if res["mask"] is not None else None,
               "rank0_phases_s": res.get("timings")}
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
