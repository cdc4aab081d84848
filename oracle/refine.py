"""CPU restatement of BoundaryRefiner.refine (Segmentation/full_evaluation_enhanced.py:332-393) — TEST INFRASTRUCTURE ONLY
(imported by tests/ and nothing else; the product path is adp_boundary_refine in the CUDA library).

The arithmetic of the reference lives in a third-party dependency that is not part of /root/reference:
opencv-python==4.8.0.76 (requirements.txt:14) — cv2.getStructuringElement, cv2.erode, cv2.dilate, cv2.bilateralFilter,
cv2.morphologyEx.  Their published algorithms are restated here in NumPy:

* MORPH_ELLIPSE structuring element: row i holds ones in [c - dx, c + dx] with
  dx = round(c * sqrt((r^2 - dy^2) / r^2)), r = rows // 2, c = cols // 2, dy = i - r (0 rows where |dy| > r).
* erode / dilate: min / max over the element's ones around the anchor (centre); pixels outside the image are ignored
  (BORDER_CONSTANT with morphologyDefaultBorderValue).  MORPH_OPEN = dilate(erode), MORPH_CLOSE = erode(dilate).
* bilateralFilter, 8-bit single channel: radius = d // 2, taps (i, j) with sqrt(i^2 + j^2) <= radius in row-major order,
  space weight float(exp(-0.5 r^2 / sigma_space^2)), colour weight table float(exp(-0.5 k^2 / sigma_color^2)) for
  k = |v - v0| in 0..255, float32 running sums `sum += v * w; wsum += w` in tap order, result cvRound(sum / wsum)
  (round half to even), BORDER_REFLECT_101.

Pinning: tests/test_oracle_refine.py checks the element and the morphology bit-for-bit against the cv2 of this image and
the bilateral stage to within ONE grey level: the 8-bit bilateral path is not bit-stable across OpenCV builds (the
cv2 4.13 wheel of this image truncates in its vector loop and rounds only in the two scalar edge columns, measured with
this file), so the restatement follows the documented cvRound form and the final refined masks are compared by agreement."""
from __future__ import annotations

import math

import numpy as np


def ellipse_kernel(ksize: int) -> np.ndarray:
    """cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (ksize, ksize))  (full_evaluation_enhanced.py:354-355)."""
    k = np.zeros((ksize, ksize), np.uint8)
    r, c = ksize // 2, ksize // 2
    inv_r2 = 1.0 / (r * r) if r else 0.0
    for i in range(ksize):
        dy = i - r
        if abs(dy) <= r:
            dx = int(round(c * math.sqrt((r * r - dy * dy) * inv_r2)))
            j1, j2 = max(c - dx, 0), min(c + dx + 1, ksize)
            k[i, j1:j2] = 1
    return k


def se_offsets(kernel: np.ndarray):
    ay, ax = kernel.shape[0] // 2, kernel.shape[1] // 2
    return [(i - ay, j - ax) for i in range(kernel.shape[0]) for j in range(kernel.shape[1]) if kernel[i, j]]


def _morph(src: np.ndarray, kernel: np.ndarray, erode: bool) -> np.ndarray:
    h, w = src.shape
    pad = max(kernel.shape)
    fill = 255 if erode else 0
    t = np.full((h + 2 * pad, w + 2 * pad), fill, np.uint8)
    t[pad:pad + h, pad:pad + w] = src
    out = np.full((h, w), fill, np.uint8)
    for dy, dx in se_offsets(kernel):
        v = t[pad + dy:pad + dy + h, pad + dx:pad + dx + w]
        out = np.minimum(out, v) if erode else np.maximum(out, v)
    return out


def erode(src, kernel):
    return _morph(src, kernel, True)


def dilate(src, kernel):
    return _morph(src, kernel, False)


def bilateral_taps(d: int, sigma_space: float):
    radius = max(d // 2, 1)
    gs = -0.5 / (sigma_space * sigma_space)
    taps = []
    for i in range(-radius, radius + 1):
        for j in range(-radius, radius + 1):
            r = math.sqrt(i * i + j * j)
            if r > radius:
                continue
            taps.append((i, j, np.float32(math.exp(r * r * gs))))
    return radius, taps


def color_weights(sigma_color: float) -> np.ndarray:
    gc = -0.5 / (sigma_color * sigma_color)
    return np.array([np.float32(math.exp(i * i * gc)) for i in range(256)], np.float32)


def _reflect101(idx: np.ndarray, n: int) -> np.ndarray:
    idx = np.where(idx < 0, -idx, idx)
    return np.where(idx >= n, 2 * n - 2 - idx, idx)


def bilateral_u8(src: np.ndarray, d: int = 5, sigma_color: float = 50.0, sigma_space: float = 50.0) -> np.ndarray:
    h, w = src.shape
    _, taps = bilateral_taps(d, sigma_space)
    cw = color_weights(sigma_color)
    ys, xs = np.arange(h), np.arange(w)
    v0 = src.astype(np.int32)
    s = np.zeros((h, w), np.float32)
    ws = np.zeros((h, w), np.float32)
    for i, j, sw in taps:
        v = src[_reflect101(ys + i, h)][:, _reflect101(xs + j, w)]
        wgt = (sw * cw[np.abs(v.astype(np.int32) - v0)]).astype(np.float32)
        s = (s + (v.astype(np.float32) * wgt).astype(np.float32)).astype(np.float32)
        ws = (ws + wgt).astype(np.float32)
    return np.rint((s / ws).astype(np.float32)).astype(np.uint8)


def refine(mask: np.ndarray, kernel_size: int = 5, bilateral_d: int = 5, sigma_color: float = 50.0,
           sigma_space: float = 50.0, return_stages: bool = False):
    """BoundaryRefiner.refine (full_evaluation_enhanced.py:357-393); `image` is unused by the reference."""
    kernel = ellipse_kernel(kernel_size)
    mask_u8 = (mask * 255).astype(np.uint8)                                               # :369
    eroded, dilated = erode(mask_u8, kernel), dilate(mask_u8, kernel)                     # :372-373
    boundary = np.logical_xor(dilated > 0, eroded > 0)                                    # :374
    filtered = bilateral_u8(mask_u8, bilateral_d, sigma_color, sigma_space)               # :378-383
    refined = np.where(boundary, filtered, mask_u8)                                       # :386
    opened = dilate(erode(refined, kernel), kernel)                                       # :389
    closed = erode(dilate(opened, kernel), kernel)                                        # :390
    out = (closed / 255.0).astype(np.float32)                                             # :393
    return (out, dict(mask_u8=mask_u8, boundary=boundary, filtered=filtered, refined=refined)) if return_stages else out
