"""Synthetic inputs shared by tests and bench.

The reference ships no data, no checkpoint and no test vectors, so every
input of the parity and benchmark runs is generated here from seed 865
(the reference's global seed, /root/reference seed.csv, src/utils/seed_utils.py:11-50).
"""
from __future__ import annotations

import numpy as np

from .layers import conv_layers, INIT_NB

SEED = 865
# the reference's own fallback statistics (src/utils/data.py:454)
DEFAULT_MEAN = 127.5
DEFAULT_STD = 50.0


def _box_blur(a: np.ndarray, r: int) -> np.ndarray:
    """Separable running-mean blur with wrap-around (three passes ~ Gaussian)."""
    out = a.astype(np.float64)
    for axis in (0, 1):
        for _ in range(3):
            c = np.cumsum(np.concatenate([out.take(range(-r, 0), axis=axis), out,
                                          out.take(range(0, r + 1), axis=axis)], axis=axis), axis=axis)
            n = out.shape[axis]
            hi = c.take(range(2 * r + 1, 2 * r + 1 + n), axis=axis)
            lo = c.take(range(0, n), axis=axis)
            out = (hi - lo) / (2 * r + 1)
    return out


def ecm_tile(size: int = 1024, seed: int = SEED, sigma_px: int = 16) -> np.ndarray:
    """uint8 grayscale tile of smooth blobs (blurred noise rescaled to 0..255)."""
    rng = np.random.default_rng(seed)
    noise = rng.standard_normal((size, size))
    r = max(1, min(sigma_px, size // 8))
    f = _box_blur(noise, r)
    f = (f - f.min()) / max(f.max() - f.min(), 1e-12)
    return np.round(f * 255.0).astype(np.uint8)


def ecm_tiles(n: int, size: int = 1024, seed: int = SEED) -> np.ndarray:
    """(n,size,size) float32 tiles in 0..255, one independent field per tile."""
    return np.stack([ecm_tile(size, seed + 7919 * i).astype(np.float32) for i in range(n)])


def mask_from_tile(tile_u8: np.ndarray, thr: int = 160) -> np.ndarray:
    """{0,1} uint8 ground-truth mask: blob > thr."""
    return (tile_u8 > thr).astype(np.uint8)


def rgb_tile(size: int = 1024, seed: int = SEED) -> np.ndarray:
    """uint8 RGB pseudocolored tile (three independent blob fields), (size,size,3)."""
    return np.stack([ecm_tile(size, seed + 104729 * (c + 1)) for c in range(3)], axis=-1)


def rgb_to_gray_u8(rgb: np.ndarray) -> np.ndarray:
    """OpenCV 4.x 8-bit RGB->gray: Y=(9798 R + 19235 G + 3735 B + 16384)>>15, bit-identical to
    cv2.cvtColor(RGB2GRAY) (checked against the installed 4.13; the reference pins 4.8.0.76 and
    reads tiles with cv2.imread(..., IMREAD_GRAYSCALE), train_adipose_unet_v3.py:555,
    full_evaluation_enhanced.py:1383)."""
    r = rgb[..., 0].astype(np.int32)
    g = rgb[..., 1].astype(np.int32)
    b = rgb[..., 2].astype(np.int32)
    return ((9798 * r + 19235 * g + 3735 * b + 16384) >> 15).astype(np.uint8)


def slide_block(by: int, bx: int, block: int = 1024, seed: int = SEED) -> np.ndarray:
    """uint8 gray block (by,bx) of a synthetic whole slide; hashed seed per block so any
    shard can regenerate its strip without I/O.  Blocks are smooth inside, discontinuous
    across block borders (irrelevant for throughput and for blend parity)."""
    h = (by * 73856093) ^ (bx * 19349663) ^ seed
    return ecm_tile(block, seed=h & 0x7FFFFFFF)


def synthetic_slide(h: int, w: int, block: int = 1024, seed: int = SEED) -> np.ndarray:
    out = np.empty((h, w), dtype=np.uint8)
    for by in range((h + block - 1) // block):
        for bx in range((w + block - 1) // block):
            blk = slide_block(by, bx, block, seed)
            y0, x0 = by * block, bx * block
            out[y0:y0 + block, x0:x0 + block] = blk[: min(block, h - y0), : min(block, w - x0)]
    return out


def init_weights(seed: int = SEED, init_nb: int = INIT_NB, gain: float = np.sqrt(2.0),
                 bias_amp: float = 0.05, deep_supervision: bool = False):
    """Glorot-uniform x gain kernels (HWIO float32) + U(-bias_amp,bias_amp) biases.

    Plain Keras Glorot + zero bias gives p in [0.40,0.51] on random input (mask nearly
    empty), which would make mask-agreement tests vacuous; the
    x sqrt(2) gain keeps activations alive through 22 ReLU layers."""
    rng = np.random.default_rng(seed)
    w = {}
    layers = list(conv_layers(init_nb))
    if deep_supervision:
        layers += [("aux_out1", init_nb * 4, 1, 1, 1), ("aux_out2", init_nb * 2, 1, 1, 1)]
    for name, ci, co, k, _ in layers:
        fan_in, fan_out = k * k * ci, k * k * co
        lim = gain * np.sqrt(6.0 / (fan_in + fan_out))
        w[name + "/kernel"] = rng.uniform(-lim, lim, size=(k, k, ci, co)).astype(np.float32)
        w[name + "/bias"] = rng.uniform(-bias_amp, bias_amp, size=(co,)).astype(np.float32)
    return w
