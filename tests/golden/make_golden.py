#!/usr/bin/env python
"""Generate tests/golden/*.npz by RUNNING THE REFERENCE'S OWN NumPy CODE.

Run in the build container only (needs /root/reference; the GPU box has no such
path and only reads the committed .npz files):

    python tests/golden/make_golden.py

The reference scripts import TensorFlow, tifffile, seaborn, skimage and
matplotlib at module level, none of which exist in this image; those modules
are replaced by inert stubs so that the pure-NumPy classes
(SlidingWindowInference, GaussianBlender, LinearBlender, TestTimeAugmentation,
binarize_prediction, calculate_pixel_metrics, parse_tile_filename,
infer_full_image_dimensions) can be imported and executed unmodified.  The
TF-dependent network itself cannot run here, so the U-Net forward is NOT
covered by these vectors (see oracle/unet.py header: parity unpinned).
"""
import os
import sys
import types
from unittest import mock

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def _stub(name):
    m = mock.MagicMock(name=name)
    m.__name__ = name
    m.__path__ = []
    m.__spec__ = None
    return m


def import_reference():
    for name in ["tensorflow", "tensorflow.keras", "tensorflow.keras.backend", "tensorflow.keras.layers",
                 "tensorflow.keras.models", "tensorflow.keras.optimizers", "tensorflow.keras.callbacks",
                 "tensorflow.keras.losses", "tifffile", "seaborn", "skimage", "skimage.morphology",
                 "skimage.measure", "matplotlib", "matplotlib.pyplot", "matplotlib.patches",
                 "matplotlib.gridspec", "matplotlib.colors"]:
        sys.modules.setdefault(name, _stub(name))
    sys.path.insert(0, os.path.join(REF, "Segmentation"))
    sys.path.insert(0, REF)
    import full_evaluation_enhanced as ev  # noqa
    import reconstruct_full_images as rc  # noqa
    return ev, rc


class FakeModel:
    """Deterministic stand-in for AdiposeUNet.predict_single so the reference's TTA /
    sliding-window plumbing can be exercised without TensorFlow: a fixed, spatially
    NON-symmetric local function of the normalised image."""

    def predict_single(self, image, mean, std):
        x = (image - mean) / (std + 1e-10)
        x = x.astype(np.float32)
        y = 0.6 * x + 0.3 * np.roll(x, 1, axis=0) - 0.2 * np.roll(x, 2, axis=1)
        return (1.0 / (1.0 + np.exp(-y))).astype(np.float32)


def main():
    ev, rc = import_reference()
    rng = np.random.default_rng(865)
    g = {}

    # ---- G1: tile positions
    cases = [(1024, 1024, 1024, 0.5), (1500, 2000, 1024, 0.5), (2560, 3072, 1024, 0.5),
             (2560, 3072, 1024, 0.75), (4096, 4096, 1024, 0.9), (1000, 1000, 1024, 0.5),
             (3000, 1024, 1024, 0.25), (32768, 32768, 1024, 0.5), (16384, 16384, 1024, 0.75),
             (300, 520, 128, 0.5), (257, 640, 128, 0.75), (128, 128, 128, 0.0)]
    g["pos_cases"] = np.array(cases, dtype=np.float64)
    for i, (h, w, t, ov) in enumerate(cases):
        sw = ev.SlidingWindowInference(tile_size=t, overlap=ov, blend_mode="none")
        pos = sw.extract_tile_positions((h, w))
        g[f"pos_{i}"] = np.array(pos, dtype=np.int64).reshape(-1, 2)
        g[f"pos_stride_{i}"] = np.int64(sw.stride)

    # ---- recon inverse geometry
    names = ["6 BEEF Shoulder -1_grid_5x5_r1_c2_r0_c1.jpg", "slide_name_r5_c3.jpg", "a_b_c_r12_c0.tif"]
    parsed = [rc.parse_tile_filename(n) for n in names]
    g["parse_names"] = np.array(names)
    g["parse_ids"] = np.array([p[0] for p in parsed])
    g["parse_rc"] = np.array([[p[1], p[2]] for p in parsed], dtype=np.int64)
    g["infer_dims"] = np.array(rc.infer_full_image_dimensions({(0, 0), (3, 5), (2, 1)}, 1024, 512), dtype=np.int64)

    # ---- G2: Gaussian windows (bit patterns)
    import hashlib
    for t, sf in [(1024, 0.25), (128, 0.25), (256, 0.5), (64, 0.25)]:
        gb = ev.GaussianBlender(tile_size=t, sigma_factor=sf)
        wm = gb.weight_map
        tag = f"gauss_{t}_{int(sf * 100)}"
        # full map only for the small ones; SHA-256 of the raw float32 bytes + probe rows for all
        g[tag + "_sha256"] = np.array(hashlib.sha256(np.ascontiguousarray(wm).tobytes()).hexdigest())
        g[tag + "_rows"] = wm[[0, 1, t // 2 - 1, t // 2, t // 2 + 1, t - 1], :]
        g[tag + "_diag"] = np.ascontiguousarray(np.diagonal(wm))
        if t <= 128:
            g[tag] = wm

    # ---- G3: blenders on a small slide (tile 128 so the fixture stays small)
    t = 64
    for tag, (h, w, ov) in {"a": (150, 260, 0.5), "b": (100, 200, 0.75)}.items():
        sw = ev.SlidingWindowInference(tile_size=t, overlap=ov, blend_mode="gaussian")
        pos = sw.extract_tile_positions((h, w))
        # fp16-representable values so the fixture stores them losslessly in half the bytes
        tiles = [rng.random((t, t), dtype=np.float32).astype(np.float16).astype(np.float32) for _ in pos]
        g[f"blend_{tag}_tiles"] = np.stack(tiles).astype(np.float16)
        g[f"blend_{tag}_pos"] = np.array(pos, dtype=np.int64)
        g[f"blend_{tag}_shape"] = np.array([h, w], dtype=np.int64)
        g[f"blend_{tag}_gauss"] = ev.GaussianBlender(tile_size=t).reconstruct(tiles, pos, (h, w))
        g[f"blend_{tag}_linear"] = ev.LinearBlender().reconstruct(tiles, pos, (h, w))

    # ---- M3: TTA transforms (aug, deaug) on a non-symmetric ramp + full TTA loop with FakeModel
    n = 8
    ramp = (np.arange(n * n, dtype=np.float32).reshape(n, n))
    for mode in ["minimal", "basic", "full"]:
        tta = ev.TestTimeAugmentation(mode=mode)
        g[f"tta_{mode}_aug"] = np.stack([np.ascontiguousarray(a(ramp)) for a, _ in tta.transforms])
        g[f"tta_{mode}_deaug"] = np.stack([np.ascontiguousarray(d(ramp)) for _, d in tta.transforms])
    img = (rng.random((64, 64)) * 255).astype(np.float32)
    g["tta_img"] = img
    for mode in ["minimal", "basic", "full"]:
        avg, _ = ev.TestTimeAugmentation(mode=mode).predict_with_tta(FakeModel(), img, 127.5, 50.0)
        g[f"tta_{mode}_avg"] = avg
    # sliding window end-to-end with the fake model (tile 64 on a 150x200 image)
    big = (rng.random((150, 200)) * 255).astype(np.float32)
    g["sw_img"] = big
    for blend in ["gaussian", "linear"]:
        sw = ev.SlidingWindowInference(tile_size=64, overlap=0.5, blend_mode=blend)
        if blend == "gaussian":
            sw.blender = ev.GaussianBlender(tile_size=64)
        g[f"sw_{blend}"] = sw.predict_with_sliding_window(big, FakeModel(), 127.5, 50.0, use_tta=True, tta_mode="full")

    # ---- G4: threshold + metrics
    pred = rng.random((96, 96), dtype=np.float32)
    gt = (rng.random((96, 96)) > 0.6).astype(np.uint8)
    g["met_pred"] = pred
    g["met_gt"] = gt
    g["met_bin"] = ev.binarize_prediction(pred, 0.5)
    keys = ["dice_score", "jaccard_index", "sensitivity", "specificity", "precision", "f1_score", "accuracy",
            "tp", "fp", "fn", "tn"]
    g["met_keys"] = np.array(keys)
    for tag, (p_, g_, thr) in {"rand": (pred, gt, 0.5), "thr7": (pred, gt, 0.7),
                               "empty": (np.zeros((16, 16), np.float32), np.zeros((16, 16), np.uint8), 0.5),
                               "nopred": (np.zeros((16, 16), np.float32), np.ones((16, 16), np.uint8), 0.5)}.items():
        m = ev.calculate_pixel_metrics(p_, g_, thr)
        g[f"met_{tag}"] = np.array([float(m[k]) for k in keys], dtype=np.float64)

    np.savez_compressed(os.path.join(OUT, "reference_numpy.npz"), **g)
    print("wrote", os.path.join(OUT, "reference_numpy.npz"), len(g), "arrays")


if __name__ == "__main__":
    main()
