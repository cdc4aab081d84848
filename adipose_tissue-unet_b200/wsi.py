"""Whole-slide sliding-window driver with tile-row strip sharding.

Replaces the loop of SlidingWindowInference.predict_with_sliding_window
(Segmentation/full_evaluation_enhanced.py:286-329) and reconstruct_slide
(Segmentation/reconstruct_full_images.py:334-417) for slides that are cut on the fly: the slide
strip lives in HBM as uint8, tiles are cut by the first conv kernel, predicted (TTA) and
blend-accumulated on the device; per-tile predictions are never materialised.

Multi-GPU: one process per GPU.  Rank g owns tile rows [g*n/G, (g+1)*n/G) and an accumulator over
the pixel rows those tiles touch.  Output rows are owned by the rank whose first tile row starts
them; a rank's partial sums that fall into the next rank's rows are shipped raw (acc, weight) — no
collective on the data path, only this boundary exchange and a final gather of disjoint strips.

Masks do not depend on the number of GPUs: float32 addition is order-dependent and the reference adds
tiles in row-major order (full_evaluation_enhanced.py:165-173), so in the boundary zone (the rows of a
strip that the strip above also reaches) the upper strip's contributions must come first.  The lower
strip therefore DEFERS the zone: it keeps the probabilities of its tiles that touch the zone, blends
only their rows below the zone while the upper strip is still working, receives the upper partial sums
into the untouched zone rows, and then replays its kept tiles over the zone in the original order
(adp_wsi_push_from_slide(defer_below_row) / adp_wsi_replay_deferred).  Every pixel sees exactly the
single-GPU sequence of float32 operations.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .api import TTA_OPCODES


def tile_positions(h: int, w: int, tile: int, stride: int) -> List[Tuple[int, int]]:
    """extract_tile_positions — full_evaluation_enhanced.py:237-265."""
    pos = []
    y_steps = max(1, math.ceil((h - tile) / stride) + 1)
    x_steps = max(1, math.ceil((w - tile) / stride) + 1)
    for yi in range(y_steps):
        for xi in range(x_steps):
            y = min(yi * stride, h - tile)
            x = min(xi * stride, w - tile)
            if y >= 0 and x >= 0 and y + tile <= h and x + tile <= w:
                pos.append((y, x))
    return pos


@dataclass
class Strip:
    rank: int
    tiles: List[Tuple[int, int]]     # (y, x) in list (row-major) order
    acc_y0: int                      # first pixel row of the accumulator
    acc_rows: int
    own_lo: int                      # output rows this rank normalises and returns
    own_hi: int
    zone_hi: int = 0                 # rows [own_lo, zone_hi) also receive the upper strip's partial sums (0 = none)


def _plan(h: int, w: int, tile: int, stride: int, world: int, active: int) -> List[Strip]:
    pos = tile_positions(h, w, tile, stride)
    rows = sorted({y for y, _ in pos})
    n = len(rows)
    strips = []
    bounds = [(g * n) // active for g in range(active + 1)] + [n] * (world - active)
    first_y = []
    for g in range(world):
        mine = rows[bounds[g]:bounds[g + 1]]
        first_y.append(mine[0] if mine else None)
    for g in range(world):
        mine = rows[bounds[g]:bounds[g + 1]]
        if not mine:
            strips.append(Strip(g, [], 0, 0, 0, 0))
            continue
        tiles = [(y, x) for (y, x) in pos if mine[0] <= y <= mine[-1]]
        nxt = next((first_y[k] for k in range(g + 1, world) if first_y[k] is not None), None)
        own_lo = 0 if all(first_y[k] is None for k in range(g)) else mine[0]
        own_hi = h if nxt is None else nxt
        strips.append(Strip(g, tiles, mine[0], mine[-1] + tile - mine[0], own_lo, own_hi))
    return strips


def plan_strips(h: int, w: int, tile: int, stride: int, world: int) -> List[Strip]:
    """Contiguous tile-row strips.  Ranks are left empty (fewer active strips than `world`) until every boundary zone
    involves exactly two neighbouring strips - the condition under which the deferred-zone exchange reproduces the
    single-GPU order of float32 additions (a strip needs at least ceil(tile/stride)-1 tile rows)."""
    for active in range(max(1, world), 0, -1):
        strips = _plan(h, w, tile, stride, world, active)
        live = [s for s in strips if s.tiles]
        ok = True
        for i, s in enumerate(live):
            end = s.acc_y0 + s.acc_rows
            if i + 1 < len(live) and end > live[i + 1].own_hi:      # reaches past the next strip's rows
                ok = False
        if ok:
            for i in range(1, len(live)):
                up_end = live[i - 1].acc_y0 + live[i - 1].acc_rows
                live[i].zone_hi = up_end if up_end > live[i].own_lo else 0
            return strips
    raise AssertionError("unreachable: a single strip has no boundary")


def boundary_transfers(strips: Sequence[Strip]) -> List[Tuple[int, int, int, int]]:
    """(src_rank, dst_rank, y, rows): raw partial sums of src that fall into rows dst owns."""
    out = []
    for s in strips:
        if not s.tiles:
            continue
        lo, hi = s.acc_y0, s.acc_y0 + s.acc_rows
        for d in strips:
            if d.rank == s.rank or not d.tiles:
                continue
            a, b = max(lo, d.own_lo), min(hi, d.own_hi)
            if b > a:
                out.append((s.rank, d.rank, a, b - a))
    return out


def reconstruct_wsi(engine, slide_rows: Callable[[int, int], np.ndarray], h: int, w: int, *, tile: int = 1024,
                    overlap: float = 0.5, blend_mode: str = "gaussian", window: Optional[np.ndarray] = None,
                    mean: float = 127.5, std: float = 50.0, tta_mode: Optional[str] = "full", threshold: float = 0.5,
                    gt_rows: Optional[Callable[[int, int], np.ndarray]] = None, rank: int = 0, world: int = 1,
                    dist=None, batch_tiles: int = 16, want_prob: bool = True, want_mask: bool = True,
                    to_device=None, timings: bool = False):
    """Run the sliding window on this rank's strip and exchange boundaries.

    slide_rows(y0, rows) -> uint8 [rows, w] gray rows of the slide, or [rows, w, 3] RGB rows (converted to gray on the
    device with OpenCV's 8-bit formula, as cv2.imread(IMREAD_GRAYSCALE) does for the reference), host.  to_device(host_u8) must return
    an object with a device pointer (`data_ptr()`), e.g. ``lambda a: torch.from_numpy(a).cuda()``.
    Returns dict(prob, mask, counts, own=(lo, hi), tiles, n_tiles_total) for THIS rank's owned rows; use
    gather_strips() for the full slide on rank 0.
    """
    ov = max(0.0, min(overlap, 0.75))
    stride = int(tile * (1 - ov))
    strips = plan_strips(h, w, tile, stride, world)
    me = strips[rank]
    ops = TTA_OPCODES[tta_mode] if tta_mode else None
    # 'gaussian' and the 'hann' extension are both window-weighted blends (the window is supplied by the caller)
    mode = _lib.BLEND_GAUSSIAN if blend_mode in ("gaussian", "hann") else _lib.BLEND_LINEAR
    result = dict(prob=None, mask=None, counts=(0, 0, 0, 0), own=(me.own_lo, me.own_hi), tiles=len(me.tiles),
                  n_tiles_total=sum(len(s.tiles) for s in strips))
    import time as _t
    ph = {}
    t_ = _t.perf_counter()

    def lap(name):
        nonlocal t_
        if timings:
            if hasattr(engine, "synchronize"):
                engine.synchronize()
            now = _t.perf_counter(); ph[name] = ph.get(name, 0.0) + now - t_; t_ = now

    if me.tiles:
        engine.wsi_begin(me.acc_rows, w, me.acc_y0, tile, mode, window if mode == _lib.BLEND_GAUSSIAN else None)
        lap("begin")
        strip_host = np.ascontiguousarray(slide_rows(me.acc_y0, me.acc_rows))
        assert strip_host.dtype == np.uint8 and strip_host.shape in ((me.acc_rows, w), (me.acc_rows, w, 3))
        channels = 3 if strip_host.ndim == 3 else 1
        lap("strip_assembly")
        strip_dev = to_device(strip_host)
        lap("strip_h2d")
        # tiles that touch the boundary zone first (they are the first tile rows of the strip): deferred blend of the zone
        deferred = [t for t in me.tiles if t[0] < me.zone_hi]
        rest = me.tiles[len(deferred):]
        assert deferred == me.tiles[:len(deferred)]
        if deferred:
            engine.wsi_push_from_slide(strip_dev, me.acc_y0, me.acc_rows, [p[0] for p in deferred], [p[1] for p in deferred],
                                       float(mean), float(std), ops, channels=channels, defer_below_row=me.zone_hi)
        for i in range(0, len(rest), batch_tiles):
            chunk = rest[i:i + batch_tiles]
            engine.wsi_push_from_slide(strip_dev, me.acc_y0, me.acc_rows, [p[0] for p in chunk], [p[1] for p in chunk],
                                       float(mean), float(std), ops, channels=channels)
        lap("tiles")
    # ---- boundary exchange: each strip hands its raw (acc, weight) overlap rows to the strip below, whose zone rows are
    # still untouched (0 + partial == partial exactly); the strip below then replays its deferred tiles over the zone
    # (only downwards => no cycle; one source per zone, plan_strips).  With an NCCL
    # group the rows travel device-to-device over NVLink straight out of / into the accumulators' staging tensors;
    # otherwise (gloo, CPU tests) through host buffers.
    transfers = boundary_transfers(strips)
    peer = _nccl_peer_path(dist, engine)
    if peer:
        import torch
        ops, incoming = [], []
        for (src, dst, y, rows) in transfers:
            if src == rank:
                buf = torch.empty((2, rows, w), dtype=torch.float32, device=f"cuda:{engine.device}")
                engine.wsi_export_into(y, rows, buf[0], buf[1])          # synchronises the engine stream
                ops.append(dist.P2POp(dist.isend, buf, dst))
            elif dst == rank:
                buf = torch.empty((2, rows, w), dtype=torch.float32, device=f"cuda:{engine.device}")
                ops.append(dist.P2POp(dist.irecv, buf, src))
                incoming.append((y, buf))
        if ops:
            for req in dist.batch_isend_irecv(ops):                       # one NCCL group: send and receive progress together
                req.wait()
            torch.cuda.synchronize()
        for y, buf in incoming:
            engine.wsi_import_add(y, buf[0], buf[1])
    else:
        for (src, dst, y, rows) in transfers:
            if src == rank:
                acc, wt = engine.wsi_export(y, rows, w)
                _send(dist, np.stack([acc, wt]), dst)
            elif dst == rank:
                buf = _recv(dist, (2, rows, w), src)
                engine.wsi_import_add(y, buf[0], buf[1])
    lap("boundary_exchange")
    if me.tiles and me.zone_hi > me.own_lo:
        engine.wsi_replay_deferred()
        lap("zone_replay")
    if me.tiles:
        rows = me.own_hi - me.own_lo
        gt = gt_rows(me.own_lo, rows) if gt_rows is not None else None
        lap("gt_assembly")
        prob, mask, counts = engine.wsi_finalize(me.own_lo, rows, w, threshold, gt, want_prob, want_mask)
        engine.wsi_end()
        lap("finalize")
        result.update(prob=prob, mask=mask, counts=counts)
    if timings:
        result["timings"] = {k: round(v, 4) for k, v in ph.items()}
    return result


def _nccl_peer_path(dist, engine) -> bool:
    """True when boundary rows can go GPU-to-GPU: an initialised torch.distributed group with an NCCL backend and a real
    device engine (the CPU tests drive a NumPy stand-in over gloo)."""
    if dist is None or not hasattr(engine, "wsi_export_into"):
        return False
    try:
        return "nccl" in str(dist.get_backend())
    except Exception:
        return False


def warmup_peer_channels(dist, rank: int, world: int, device: int = 0):
    """Open the NCCL point-to-point channels between neighbouring strips (communicator setup costs ~0.2 s the first
    time) so that a timed reconstruct_wsi only pays for the transfer itself."""
    if dist is None or world < 2 or "nccl" not in str(dist.get_backend()):
        return
    import torch
    t = torch.zeros(1, device=f"cuda:{device}")
    ops = []
    if rank + 1 < world:
        ops.append(dist.P2POp(dist.isend, t, rank + 1))
    if rank > 0:
        ops.append(dist.P2POp(dist.irecv, torch.zeros(1, device=f"cuda:{device}"), rank - 1))
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    torch.cuda.synchronize()


def _send(dist, arr: np.ndarray, dst: int):
    import torch
    dist.send(torch.from_numpy(np.ascontiguousarray(arr)), dst)


def _recv(dist, shape, src: int) -> np.ndarray:
    import torch
    t = torch.empty(shape, dtype=torch.float32)
    dist.recv(t, src)
    return t.numpy()


def gather_strips(result: dict, h: int, w: int, rank: int, world: int, dist=None):
    """Halo-free host gather: rank 0 concatenates the disjoint owned strips and sums the counts."""
    if world == 1:
        return result["prob"], result["mask"], result["counts"]
    payload = [None] * world
    dist.all_gather_object(payload, dict(own=result["own"], counts=result["counts"],
                                         prob=result["prob"], mask=result["mask"]))
    if rank != 0:
        return None, None, None
    prob = np.zeros((h, w), np.float32) if any(p["prob"] is not None for p in payload) else None
    mask = np.zeros((h, w), np.uint8) if any(p["mask"] is not None for p in payload) else None
    counts = [0, 0, 0, 0]
    for p in payload:
        lo, hi = p["own"]
        if hi <= lo:
            continue
        if prob is not None and p["prob"] is not None:
            prob[lo:hi] = p["prob"]
        if mask is not None and p["mask"] is not None:
            mask[lo:hi] = p["mask"]
        counts = [a + b for a, b in zip(counts, p["counts"])]
    return prob, mask, tuple(counts)
