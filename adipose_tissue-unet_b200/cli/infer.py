"""Folder inference — drop-in for Segmentation/segmentation_inference.py (argparse :324-350, outputs :455-470).

Outputs: masks/{stem}_mask.tif (uint8 {0,1}), probabilities/{stem}_prob.tif (uint8 trunc(p*255)),
overlays/{stem}_overlay.png.  Images that are not 1024x1024 are skipped with the reference's warning.
Tiles are sent to the GPU in batches; prediction (+TTA) and thresholding run on the device."""
from __future__ import annotations

import argparse
import sys
import time
from pathlib import Path

import cv2
import numpy as np

from . import common as C


def build_parser():
    p = argparse.ArgumentParser(description="Run segmentation inference on a folder of images")
    p.add_argument("--images-dir", type=str, required=True, help="Directory containing input images")
    p.add_argument("--output-dir", type=str, required=True, help="Directory to save predictions")
    p.add_argument("--weights", type=str, required=True, help="Path to model weights file or checkpoint directory")
    p.add_argument("--threshold", type=float, default=0.5, help="Binarization threshold (0-1, default: 0.5)")
    p.add_argument("--use-tta", action="store_true", default=False, help="Use Test Time Augmentation")
    p.add_argument("--tta-mode", type=str, default="basic", choices=["minimal", "basic", "full"])
    p.add_argument("--save-overlays", action="store_true", default=False)
    p.add_argument("--overlay-color", type=str, default="cyan", choices=list(C.OVERLAY_COLORS))
    p.add_argument("--save-probability", action="store_true", default=False)
    C.add_engine_args(p)
    return p


def main(argv=None) -> int:
    args = build_parser().parse_args(argv)
    images_dir, output_dir = Path(args.images_dir), Path(args.output_dir)
    if not images_dir.exists():
        print(f"❌ Error: Images directory not found: {images_dir}")
        return 1
    masks_dir = output_dir / "masks"; masks_dir.mkdir(parents=True, exist_ok=True)
    overlays_dir = output_dir / "overlays"; prob_dir = output_dir / "probabilities"
    if args.save_overlays:
        overlays_dir.mkdir(parents=True, exist_ok=True)
    if args.save_probability:
        prob_dir.mkdir(parents=True, exist_ok=True)
    print(f"\n{'=' * 80}\nADIPOSE TISSUE SEGMENTATION - INFERENCE\n{'=' * 80}")
    print(f"Images: {images_dir}\nOutput: {output_dir}\nThreshold: {args.threshold:.2f}")
    print(f"TTA: {'Enabled (' + args.tta_mode + ')' if args.use_tta else 'Disabled'}\n{'=' * 80}\n")
    weights_file, ckpt_dir = C.find_weights_file(args.weights)
    if weights_file.lower().endswith(".onnx"):
        print("❌ ONNX weights are served by the reference's OnnxUnetPredictor, not by this engine")
        return 1
    model = C.make_model(weights_file, args.precision, args.device, max(args.batch_tiles, 8))
    mean, std = C.load_normalization_stats(ckpt_dir)
    files = C.list_images(images_dir)
    if not files:
        print(f"❌ Error: No images found in {images_dir}\n   Looking for: {C.IMAGE_EXTS}")
        return 1
    print(f"\nFound {len(files)} images\nProcessing...\n")
    tta_mode = args.tta_mode if args.use_tta else None
    color = C.OVERLAY_COLORS[args.overlay_color]
    start = time.time()
    batch, names = [], []

    def flush():
        if not batch:
            return
        tiles = np.stack(batch)                                   # uint8 (n,1024,1024): gray tiles as read by OpenCV
        probs = model.predict_batch(tiles, mean, std, tta_mode)   # forward (+TTA mean) on the device
        for img, p, f in zip(batch, probs, names):
            if args.save_probability:
                C.write_tiff_u8(prob_dir / f"{f.stem}_prob.tif", (p * 255).astype(np.uint8))
            mask, _ = model.engine.threshold_metrics(p, None, args.threshold)
            C.write_tiff_u8(masks_dir / f"{f.stem}_mask.tif", mask)
            if args.save_overlays:
                ov = C.overlay(img, mask, color)
                cv2.imwrite(str(overlays_dir / f"{f.stem}_overlay.png"), cv2.cvtColor(ov, cv2.COLOR_RGB2BGR))
        batch.clear(); names.clear()

    for f in files:
        img = C.read_gray(f)
        if img is None:
            print(f"⚠️  Warning: Failed to load {f.name}, skipping")
            continue
        if img.shape != (1024, 1024):
            print(f"⚠️  Warning: {f.name} is {img.shape}, expected (1024, 1024), skipping")
            continue
        batch.append(img.astype(np.uint8) if img.dtype == np.uint8 else img.astype(np.float32))
        names.append(f)
        if len(batch) >= args.batch_tiles or batch[-1].dtype != batch[0].dtype:
            flush()
    flush()
    elapsed = time.time() - start
    print(f"\n{'=' * 80}\nINFERENCE COMPLETE\n{'=' * 80}")
    print(f"Processed: {len(files)} images\nTime: {elapsed:.1f}s ({elapsed / len(files):.2f}s per image)")
    print(f"\nOutput saved to:\n  Masks: {masks_dir}")
    if args.save_overlays:
        print(f"  Overlays: {overlays_dir}")
    if args.save_probability:
        print(f"  Probabilities: {prob_dir}")
    print(f"{'=' * 80}\n")
    return 0


if __name__ == "__main__":
    sys.exit(main())
