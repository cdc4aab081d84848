"""Host-side helpers shared by the CLIs: checkpoint discovery, normalisation statistics, image files."""
from __future__ import annotations

import json
import os
from pathlib import Path
from typing import Dict, List, Optional, Tuple

import cv2
import numpy as np

# segmentation_inference.py:252-279 / full_evaluation_enhanced.py:456-487
WEIGHT_CANDIDATES_BEST = ["weights_best_overall.weights.h5", "phase2_best.weights.h5", "phase1_best.weights.h5",
                          "best_model.weights.h5", "model_best.weights.h5", "weights_best.weights.h5"]
WEIGHT_CANDIDATES_EMA = ["weights_ema.weights.h5", "ema_weights_phase2.weights.h5", "ema_weights.weights.h5"]
IMAGE_EXTS = {".jpg", ".jpeg", ".png", ".tif", ".tiff"}
OVERLAY_COLORS = {"cyan": (0, 255, 255), "yellow": (255, 255, 0), "magenta": (255, 0, 255), "green": (0, 255, 0),
                  "red": (255, 0, 0)}


def find_weights_file(weights_arg: str, use_ema: bool = False) -> Tuple[str, Path]:
    """File, or directory searched in the reference's candidate order -> (weights file, checkpoint dir)."""
    p = Path(weights_arg)
    if p.is_file():
        return str(p), p.parent
    names = (WEIGHT_CANDIDATES_EMA if use_ema else []) + WEIGHT_CANDIDATES_BEST
    for n in names:
        if (p / n).exists():
            return str(p / n), p
    files = sorted(p.glob("*.weights.h5")) + sorted(p.glob("*.h5")) + sorted(p.glob("*.npz"))
    if files:
        return str(files[0]), p
    raise FileNotFoundError(f"No weights files found in {p}")


def load_normalization_stats(checkpoint_dir: Path) -> Tuple[float, float]:
    """normalization_stats.json written by training (train_adipose_unet_v3.py:1194-1207); inference always applies
    z-score with these (full_evaluation_enhanced.py:1306), whatever normalization_method says."""
    f = Path(checkpoint_dir) / "normalization_stats.json"
    if not f.exists():
        print(f"⚠️  Warning: {f} not found, using default normalization")
        return 0.0, 1.0
    with open(f) as fh:
        s = json.load(fh)
    mean, std = float(s["mean"]), float(s["std"])
    print(f"✓ Loaded normalization stats: mean={mean:.4f}, std={std:.4f}")
    return mean, std


def detect_deep_supervision(checkpoint_dir: Path) -> bool:
    f = Path(checkpoint_dir) / "training_settings.log"
    if not f.exists():
        return False
    txt = f.read_text(errors="replace")
    return "use_deep_supervision: True" in txt or "deep_supervision: True" in txt


def read_gray(path) -> Optional[np.ndarray]:
    """cv2.imread(..., IMREAD_GRAYSCALE) as uint8/uint16 (the reference's tile reader, train_adipose_unet_v3.py:555)."""
    p = str(path)
    if p.lower().endswith((".tif", ".tiff")):
        a = cv2.imread(p, cv2.IMREAD_UNCHANGED)
        if a is None:
            return None
        if a.ndim == 3:
            a = cv2.cvtColor(a[..., :3], cv2.COLOR_BGR2GRAY)
        return a
    return cv2.imread(p, cv2.IMREAD_GRAYSCALE)


def read_mask(path) -> np.ndarray:
    """0/1 (or 0/255) TIFF mask as float32 in [0,1] (reconstruct_full_images.py:386-395)."""
    a = cv2.imread(str(path), cv2.IMREAD_UNCHANGED)
    if a is None:
        raise FileNotFoundError(path)
    a = a.astype(np.float32)
    if a.ndim == 3:
        a = a[..., 0]
    if a.max() > 1.0:
        a = a / 255.0
    return a


def write_tiff_u8(path, arr: np.ndarray):
    """uint8 TIFF, LZW: tifffile.imwrite(path, arr, compression='lzw') of the reference, by the library's parallel strip writer."""
    from ..api import write_tiff_lzw
    write_tiff_lzw(path, np.ascontiguousarray(arr.astype(np.uint8)))


def overlay(image_rgb_u8: np.ndarray, mask: np.ndarray, color: Tuple[int, int, int]) -> np.ndarray:
    """60 % image + 40 % colour mask (segmentation_inference.py:282-298, reconstruct_full_images.py:423-455)."""
    if image_rgb_u8.ndim == 2:
        image_rgb_u8 = cv2.cvtColor(image_rgb_u8, cv2.COLOR_GRAY2RGB)
    cm = image_rgb_u8.copy() if False else np.zeros_like(image_rgb_u8)
    cm[mask > 0] = color
    return cv2.addWeighted(image_rgb_u8, 0.6, cm, 0.4, 0)


def list_images(d: Path) -> List[Path]:
    return sorted(f for f in Path(d).iterdir() if f.suffix.lower() in IMAGE_EXTS and f.is_file())


def make_model(weights_file: str, precision: str, device: int = 0, max_forwards: int = 16):
    from .. import api
    m = api.AdiposeUNet(precision=precision, device=device, max_forwards=max_forwards)
    m.build_model()
    m.load_weights(weights_file)
    return m


def add_engine_args(parser):
    g = parser.add_argument_group("B200 engine (not in the reference)")
    g.add_argument("--precision", choices=["bf16", "bf16x3", "fp32"], default="bf16",
                   help="bf16 = tcgen05 tensor-core path (default, |dp| <= 1e-2); bf16x3 = hi/lo split on the tensor cores "
                        "(|dp| <= 1e-4); fp32 = exact CUDA-core path")
    g.add_argument("--device", type=int, default=0)
    g.add_argument("--batch-tiles", type=int, default=16, help="tiles per device batch")


def prefetch(gen, depth: int = 2):
    """tf.data's `dataset.prefetch(...)` of the reference's input pipeline (train_adipose_unet_v3.py:609-623): the generator
    runs in a background thread `depth` items ahead, so JPEG decode / augmentation of the next batch overlaps the device
    step of the current one (cv2 and the ctypes calls release the GIL).  Order is preserved; an exception raised by the
    generator is re-raised in the consumer; abandoning the iterator early stops the producer."""
    import queue
    import threading
    q: "queue.Queue" = queue.Queue(maxsize=max(1, depth))
    stop = threading.Event()
    _END, _ERR = object(), object()

    def put(item) -> bool:
        while not stop.is_set():
            try:
                q.put(item, timeout=0.1)
                return True
            except queue.Full:
                continue
        return False

    def producer():
        try:
            for item in gen:
                if not put(item):
                    return
            put(_END)
        except BaseException as ex:          # noqa: BLE001 - handed to the consumer
            put((_ERR, ex))

    th = threading.Thread(target=producer, daemon=True)
    th.start()
    try:
        while True:
            item = q.get()
            if item is _END:
                return
            if isinstance(item, tuple) and len(item) == 2 and item[0] is _ERR:
                raise item[1]
            yield item
    finally:
        stop.set()
