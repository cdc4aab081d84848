"""ORACLE (test infrastructure, never shipped, never on the product path).

NumPy restatement of the reference's host-side tile geometry, blend windows,
blenders, dihedral test-time-augmentation transforms, threshold and pixel
metrics.  Every function cites the reference lines it follows (paths relative
to /root/reference).  These rows are PINNED:
tests/golden/make_golden.py imports the reference's own NumPy classes (with
TensorFlow & co. stubbed) and stores their outputs in tests/golden/*.npz;
tests/test_oracle_geometry.py checks this restatement against those vectors.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np


# --------------------------------------------------------------------- G1
def clamp_overlap(overlap: float) -> float:
    """Segmentation/full_evaluation_enhanced.py:223"""
    return max(0.0, min(overlap, 0.75))


def stride_for(tile: int, overlap: float) -> int:
    """Segmentation/full_evaluation_enhanced.py:224"""
    return int(tile * (1 - clamp_overlap(overlap)))


def tile_positions(h: int, w: int, tile: int, stride: int) -> List[Tuple[int, int]]:
    """Segmentation/full_evaluation_enhanced.py:237-265 (row-major, y outer)."""
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


def parse_tile_filename(filename: str) -> Tuple[str, int, int]:
    """Segmentation/reconstruct_full_images.py:114-146"""
    import os
    stem = os.path.splitext(os.path.basename(filename))[0]
    parts = stem.split("_")
    if len(parts) >= 2 and parts[-2].startswith("r") and parts[-1].startswith("c"):
        try:
            return "_".join(parts[:-2]), int(parts[-2][1:]), int(parts[-1][1:])
        except (ValueError, IndexError):
            pass
    raise ValueError(f"Cannot parse tile position from filename: {filename}")


def infer_full_dims(rc: Sequence[Tuple[int, int]], tile: int, stride: int) -> Tuple[int, int]:
    """Segmentation/reconstruct_full_images.py:240-271"""
    if not rc:
        return (0, 0)
    return (max(r for r, _ in rc) * stride + tile, max(c for _, c in rc) * stride + tile)


def recon_tile_origin(row: int, col: int, tile: int, stride: int, h: int, w: int) -> Tuple[int, int]:
    """Segmentation/reconstruct_full_images.py:399-400 (edge clamp)."""
    return (min(row * stride, h - tile), min(col * stride, w - tile))


# --------------------------------------------------------------------- G2
def gaussian_window(tile: int = 1024, sigma_factor: float = 0.25) -> np.ndarray:
    """Segmentation/full_evaluation_enhanced.py:133-147 (float64 -> /max -> float32;
    centre = tile/2, i.e. half-pixel asymmetric)."""
    sigma = tile * sigma_factor
    center = tile / 2
    y, x = np.ogrid[0:tile, 0:tile]
    dist_sq = (x - center) ** 2 + (y - center) ** 2
    w = np.exp(-dist_sq / (2 * sigma ** 2))
    w = w / w.max()
    return w.astype(np.float32)


# --------------------------------------------------------------------- G2x (extension, NOT in the reference)
HANN_FLOOR = 1e-3


def hann_window(tile: int = 1024) -> np.ndarray:
    """Hann blending window named by BASELINE.json's north_star; the reference has only Gaussian and linear
    blenders, so this is a labelled extension with its own formula:

        h[i] = max(0.5 - 0.5 * cos(2*pi*(i + 0.5) / tile), 1e-3),  w[y, x] = h[y] * h[x]   (float64 -> float32)

    The half-sample shift makes the window symmetric and keeps it positive (np.hanning is 0 at both ends); the floor keeps
    every weight >= 1e-6, i.e. above the 1e-8 clamp of the blender's normalisation `acc / max(weight_sum, 1e-8)`
    (full_evaluation_enhanced.py:176-181) - without it the ~60 pixels next to each slide corner, covered by one tile with
    weight < 1e-8, would be attenuated.  At 50 % overlap h[i] + h[i + tile/2] == 1 away from the floored ends, so the weight
    sum of interior pixels is 1 up to float32 rounding."""
    i = np.arange(tile, dtype=np.float64)
    h = np.maximum(0.5 - 0.5 * np.cos(2.0 * np.pi * (i + 0.5) / tile), HANN_FLOOR)
    return np.outer(h, h).astype(np.float32)


def hann_reconstruct(tiles, positions, output_shape, window: np.ndarray = None, return_parts: bool = False):
    """Weighted blend with the Hann window: same accumulation statement as GaussianBlender.reconstruct
    (full_evaluation_enhanced.py:149-183), different weights."""
    if window is None:
        window = hann_window(tiles[0].shape[0])
    return gaussian_reconstruct(tiles, positions, output_shape, window, return_parts)


# --------------------------------------------------------------------- G3
def gaussian_reconstruct(tiles, positions, output_shape, window: np.ndarray,
                         return_parts: bool = False):
    """Segmentation/full_evaluation_enhanced.py:149-183"""
    h, w = output_shape
    acc = np.zeros((h, w), dtype=np.float32)
    wsum = np.zeros((h, w), dtype=np.float32)
    for tile, (y, x) in zip(tiles, positions):
        th, tw = tile.shape[:2]
        ws = window[:th, :tw]
        acc[y:y + th, x:x + tw] += tile * ws
        wsum[y:y + th, x:x + tw] += ws
    wsum_c = np.maximum(wsum, 1e-8)
    res = (acc / wsum_c).astype(np.float32)
    return (res, acc, wsum) if return_parts else res


def linear_reconstruct(tiles, positions, output_shape):
    """Segmentation/full_evaluation_enhanced.py:189-204 (count is int32; the final
    division float32/int32 promotes to float64 in NumPy, then casts to float32)."""
    h, w = output_shape
    acc = np.zeros((h, w), dtype=np.float32)
    cnt = np.zeros((h, w), dtype=np.int32)
    for tile, (y, x) in zip(tiles, positions):
        th, tw = tile.shape[:2]
        acc[y:y + th, x:x + tw] += tile
        cnt[y:y + th, x:x + tw] += 1
    cnt = np.maximum(cnt, 1)
    return (acc / cnt).astype(np.float32)


# --------------------------------------------------------------------- M3
# out[i, j] = in[src(i, j)] for an N x N image; matches np.rot90 / np.flip as used in
# Segmentation/full_evaluation_enhanced.py:535-568 (same list in segmentation_inference.py:181-219)
def _ident(x): return x
def _rot90(x): return np.rot90(x, 1)
def _rot180(x): return np.rot90(x, 2)
def _rot270(x): return np.rot90(x, 3)
def _flip_h(x): return np.flip(x, axis=1)
def _flip_v(x): return np.flip(x, axis=0)


TTA_FULL = [
    (_ident, _ident),
    (_rot90, _rot270),
    (_rot180, _rot180),
    (_rot270, _rot90),
    (_flip_h, _flip_h),
    (_flip_v, _flip_v),
    (lambda x: _flip_h(_rot90(x)), lambda x: _rot270(_flip_h(x))),
    (lambda x: _flip_v(_rot90(x)), lambda x: _rot270(_flip_v(x))),
]
TTA_BASIC = [(_ident, _ident), (_flip_h, _flip_h), (_flip_v, _flip_v), (_rot90, _rot270)]
TTA_MINIMAL = [(_ident, _ident), (_flip_h, _flip_h)]
TTA_MODES = {"minimal": TTA_MINIMAL, "basic": TTA_BASIC, "full": TTA_FULL}

# Dihedral op codes used on the device (adipose_b200.h): aug index map out[i,j]=in[...]
#  0 ident (i,j)      1 rot90 (j,N-1-i)   2 rot180 (N-1-i,N-1-j)  3 rot270 (N-1-j,i)
#  4 flip_h (i,N-1-j) 5 flip_v (N-1-i,j)  6 anti-transpose (N-1-j,N-1-i)  7 transpose (j,i)
TTA_OPCODES = {"minimal": [0, 4], "basic": [0, 4, 5, 1], "full": [0, 1, 2, 3, 4, 5, 6, 7]}
# inverse op of each op code (ops 1 and 3 invert each other, the rest are involutions)
D4_INVERSE = [0, 3, 2, 1, 4, 5, 6, 7]


def d4_src_index(op: int, i, j, n: int):
    """Source (row, col) read by output (i, j) under op code `op`."""
    if op == 0: return i, j
    if op == 1: return j, n - 1 - i
    if op == 2: return n - 1 - i, n - 1 - j
    if op == 3: return n - 1 - j, i
    if op == 4: return i, n - 1 - j
    if op == 5: return n - 1 - i, j
    if op == 6: return n - 1 - j, n - 1 - i
    if op == 7: return j, i
    raise ValueError(op)


def d4_apply(op: int, x: np.ndarray) -> np.ndarray:
    n = x.shape[0]
    ii, jj = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    si, sj = d4_src_index(op, ii, jj, n)
    return x[si, sj]


def tta_mean(preds_deaug: Sequence[np.ndarray]) -> np.ndarray:
    """np.mean(preds, axis=0).astype(float32) — full_evaluation_enhanced.py:594.
    For float32 planes NumPy reduces axis 0 by sequential float32 adds, then divides."""
    return np.mean(list(preds_deaug), axis=0).astype(np.float32)


# --------------------------------------------------------------------- G4
def binarize(pred: np.ndarray, threshold: float = 0.5) -> np.ndarray:
    """Segmentation/full_evaluation_enhanced.py:716-718 (strict >)."""
    return (pred > threshold).astype(np.uint8)


def prob_to_u8(pred: np.ndarray) -> np.ndarray:
    """(p*255).astype(uint8) truncation — segmentation_inference.py:457,
    reconstruct_full_images.py:733."""
    return (pred * 255).astype(np.uint8)


def pixel_metrics(pred: np.ndarray, true: np.ndarray, threshold: float = 0.5) -> dict:
    """Segmentation/full_evaluation_enhanced.py:721-785"""
    pb = pred > threshold
    tb = true > 0.5
    if not tb.any() and not pb.any():
        return dict(dice_score=1.0, jaccard_index=1.0, sensitivity=1.0, specificity=1.0,
                    precision=1.0, f1_score=1.0, accuracy=1.0, tp=0, fp=0, fn=0, tn=int(tb.size))
    tp = np.sum(pb & tb); fp = np.sum(pb & ~tb); fn = np.sum(~pb & tb); tn = np.sum(~pb & ~tb)
    return metrics_from_counts(int(tp), int(fp), int(fn), int(tn), both_empty_rule=False)


def metrics_from_counts(tp: int, fp: int, fn: int, tn: int, both_empty_rule: bool = True) -> dict:
    """Ratios of full_evaluation_enhanced.py:761-771 from the four counts."""
    if both_empty_rule and tp == 0 and fp == 0 and fn == 0:
        return dict(dice_score=1.0, jaccard_index=1.0, sensitivity=1.0, specificity=1.0,
                    precision=1.0, f1_score=1.0, accuracy=1.0, tp=0, fp=0, fn=0, tn=int(tn))
    precision = tp / (tp + fp + 1e-10)
    sensitivity = tp / (tp + fn + 1e-10)
    specificity = tn / (tn + fp + 1e-10)
    accuracy = (tp + tn) / (tp + fp + fn + tn + 1e-10)
    f1 = 2 * tp / (2 * tp + fp + fn + 1e-10)
    jaccard = tp / (tp + fp + fn + 1e-10)
    return dict(dice_score=float(f1), jaccard_index=float(jaccard), sensitivity=float(sensitivity),
                specificity=float(specificity), precision=float(precision), f1_score=float(f1),
                accuracy=float(accuracy), tp=int(tp), fp=int(fp), fn=int(fn), tn=int(tn))


# --------------------------------------------------------------------- boundary metrics
def boundary_metrics_reference(pred: np.ndarray, true: np.ndarray, threshold: float = 0.5) -> dict:
    """Segmentation/full_evaluation_enhanced.py:788-844 statement by statement, with scipy.ndimage for the distance
    transform and for skimage.morphology.binary_erosion (skimage is not in this image; its published definition is
    ndimage.binary_erosion with the 3x3 cross footprint and border_value=True - unpinned).  Note :829-830: each mask's own
    distance transform is sampled on its own surface."""
    from scipy import ndimage
    pred_bin = pred > threshold
    true_bin = true > 0.5
    if not pred_bin.any() and not true_bin.any():
        return {"hausdorff95": 0.0, "assd": 0.0}
    if not pred_bin.any() or not true_bin.any():
        return {"hausdorff95": float("inf"), "assd": float("inf")}
    pred_dt = ndimage.distance_transform_edt(~pred_bin, sampling=(1.0, 1.0))
    true_dt = ndimage.distance_transform_edt(~true_bin, sampling=(1.0, 1.0))
    cross = ndimage.generate_binary_structure(2, 1)
    pred_surface = pred_bin & ~ndimage.binary_erosion(pred_bin, structure=cross, border_value=1)
    true_surface = true_bin & ~ndimage.binary_erosion(true_bin, structure=cross, border_value=1)
    if pred_surface.sum() > 0 and true_surface.sum() > 0:
        allv = np.concatenate([pred_dt[pred_surface], true_dt[true_surface]])
        return {"hausdorff95": float(np.percentile(allv, 95)), "assd": float(np.mean(allv))}
    return {"hausdorff95": float("inf"), "assd": float("inf")}
