"""Whole-slide reconstruction from overlapping tiles - drop-in for Segmentation/reconstruct_full_images.py
(argparse :882-929, pipeline :586-866, per-slide reconstruction :334-417).

Per slide: tiles `{slide}_r{row}_c{col}.jpg` are predicted (+TTA) on the device and blended straight into the engine's
whole-slide accumulator (adp_wsi_*), the normalised probability map, mask and TP/FP/FN/TN come back from one finalize
call.  The ground-truth and RGB mosaics are blends of given tiles (no network) and use adp_blend_reconstruct.
Outputs keep the reference's names: {slide}/original_image.tif (BGR), prediction_mask.tif (probability*255),
ground_truth_mask.tif, gt_overlay.png, pred_overlay.png, metrics.txt, metrics/{slide}_metrics.json, metrics/summary.csv,
reconstruction_log.json.  With --boundary-refine every tile prediction passes through BoundaryRefiner.refine on the device
(adp_boundary_refine) before it is blended, as in reconstruct_full_images.py:378-380."""
from __future__ import annotations

import argparse
import csv
import json
import sys
import warnings
from collections import defaultdict
from datetime import datetime
from pathlib import Path

import cv2
import numpy as np

from . import common as C


def parse_tile_filename(filename: str):
    """`X_r{r}_c{c}.jpg` -> (slide, r, c) (reconstruct_full_images.py:121-147)."""
    parts = Path(filename).stem.split("_")
    if len(parts) >= 2 and parts[-2].startswith("r") and parts[-1].startswith("c"):
        try:
            return "_".join(parts[:-2]), int(parts[-2][1:]), int(parts[-1][1:])
        except (ValueError, IndexError):
            pass
    raise ValueError(f"Cannot parse tile position from filename: {filename}")


def group_tiles_by_slide(images_dir: Path, masks_dir: Path | None):
    slides = defaultdict(lambda: {"tiles": [], "positions": set()})
    mask_files = {p.stem: p for p in masks_dir.glob("*.tif")} if masks_dir and masks_dir.exists() else {}
    for img in sorted(images_dir.glob("*.jpg")):
        try:
            sid, r, c = parse_tile_filename(img.name)
        except ValueError as e:
            warnings.warn(f"Skipping file {img.name}: {e}")
            continue
        slides[sid]["tiles"].append((r, c, img, mask_files.get(img.stem)))
        slides[sid]["positions"].add((r, c))
    for info in slides.values():
        rows = [r for r, _ in info["positions"]]; cols = [c for _, c in info["positions"]]
        info["row_range"] = (min(rows), max(rows)); info["col_range"] = (min(cols), max(cols))
    return dict(slides)


def infer_full_image_dimensions(positions, tile_size: int, stride: int):
    """reconstruct_full_images.py:240-271 (fallback when the source image is not at hand)."""
    if not positions:
        return (0, 0)
    return (max(r for r, _ in positions) * stride + tile_size, max(c for _, c in positions) * stride + tile_size)


def build_parser():
    p = argparse.ArgumentParser(description="Reconstruct full images from overlapping tiles")
    p.add_argument("--weights", type=str, required=True)
    p.add_argument("--data-root", type=str, required=True)
    p.add_argument("--output-dir", type=str, required=True)
    p.add_argument("--tile-size", type=int, default=1024)
    p.add_argument("--stride", type=int, default=512)
    p.add_argument("--threshold", type=float, default=0.5)
    p.add_argument("--blend-mode", type=str, default="gaussian", choices=["gaussian", "linear", "hann"],
                   help="gaussian / linear as in the reference; hann is an extension (api.HannBlender)")
    p.add_argument("--use-tta", action="store_true", default=False)
    p.add_argument("--tta-mode", type=str, default="basic", choices=["minimal", "basic", "full"])
    p.add_argument("--boundary-refine", action="store_true", default=False)
    p.add_argument("--refine-kernel", type=int, default=5)
    p.add_argument("--save-masks", action="store_true", default=True)
    p.add_argument("--save-overlays", action="store_true", default=True)
    p.add_argument("--save-comparisons", action="store_true", default=True)
    p.add_argument("--save-metrics", action="store_true", default=True)
    p.add_argument("--min-coverage", type=float, default=0.90)
    p.add_argument("--max-tiles", type=int, default=None)
    C.add_engine_args(p)
    return p


def reconstruct_slide(model, tiles_info, full_shape, tile_size, stride, mean, std, blend_mode, tta_mode, threshold, batch_tiles,
                      refine_kernel=None):
    """-> (rgb float32 [0,1], probability, ground truth or None, mask, (tp,fp,fn,tn) or None)."""
    from .. import api, _lib
    eng = model.engine
    H, W = full_shape
    mode = _lib.BLEND_GAUSSIAN if blend_mode in ("gaussian", "hann") else _lib.BLEND_LINEAR
    window = api.blend_window(blend_mode, tile_size)
    ops = api.TTA_OPCODES[tta_mode] if tta_mode else None
    eng.wsi_begin(H, W, 0, tile_size, mode, window)
    positions, gt_tiles, rgb_tiles = [], [], []
    batch, ys, xs = [], [], []

    def flush():
        if batch:
            if refine_kernel:          # :372-380: predict (+TTA), BoundaryRefiner.refine per tile, then blend the refined tiles
                probs = eng.predict(np.stack(batch).astype(np.float32), mean, std, ops)
                eng.wsi_push_probs(eng.boundary_refine(probs, kernel_size=refine_kernel), ys, xs)
            else:
                eng.wsi_push_tiles(np.stack(batch).astype(np.float32), ys, xs, mean, std, ops)
            batch.clear(); ys.clear(); xs.clear()

    for row, col, img_path, mask_path in tiles_info:
        bgr = cv2.imread(str(img_path), cv2.IMREAD_COLOR)
        gray = cv2.imread(str(img_path), cv2.IMREAD_GRAYSCALE)
        rgb_tiles.append(cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB).astype(np.float32) / 255.0)
        y = min(row * stride, H - tile_size); x = min(col * stride, W - tile_size)      # edge clamp, :399-400
        positions.append((y, x))
        batch.append(gray); ys.append(y); xs.append(x)
        if len(batch) >= batch_tiles:
            flush()
        if mask_path is not None:
            gt_tiles.append(C.read_mask(mask_path))
    flush()
    full_gt = None
    if gt_tiles:
        full_gt = eng.blend(mode, gt_tiles, positions, (H, W), window)
    prob, mask, counts = eng.wsi_finalize(0, H, W, threshold, gt=full_gt, want_prob=True, want_mask=True)
    eng.wsi_end()
    rgb = np.zeros((H, W, 3), np.float32)
    for ch in range(3):
        rgb[:, :, ch] = eng.blend(mode, [t[:, :, ch] for t in rgb_tiles], positions, (H, W), window)
    return rgb, prob, full_gt, mask, (counts if full_gt is not None else None)


def main(argv=None) -> int:
    from .. import api
    args = build_parser().parse_args(argv)
    print(f"\n{'=' * 80}\nFULL IMAGE RECONSTRUCTION FROM OVERLAPPING TILES\n{'=' * 80}")
    data_root = Path(args.data_root)
    images_dir, masks_dir = data_root / "images", data_root / "masks"
    output_dir = Path(args.output_dir)
    if args.max_tiles:
        output_dir = output_dir.parent / f"{output_dir.name}_{args.max_tiles}x{args.max_tiles}"
    if not images_dir.exists():
        raise FileNotFoundError(f"Images directory not found: {images_dir}")
    for sub in ("masks", "overlays", "comparisons", "metrics"):
        (output_dir / sub).mkdir(parents=True, exist_ok=True)
    weights_file, ckpt_dir = C.find_weights_file(args.weights)
    mean, std = C.load_normalization_stats(ckpt_dir)
    model = C.make_model(weights_file, args.precision, args.device, max(args.batch_tiles, 8))
    print("✓ Model loaded successfully")
    if args.boundary_refine:
        print(f"✓ Boundary refinement enabled (kernel={args.refine_kernel})")
    slides = group_tiles_by_slide(images_dir, masks_dir)
    print(f"✓ Found {len(slides)} slide(s)")
    tta_mode = args.tta_mode if args.use_tta else None
    results = []
    for sid, info in slides.items():
        print(f"\n{'=' * 80}\nProcessing: {sid}\n{'=' * 80}")
        tiles, positions = info["tiles"], info["positions"]
        if args.max_tiles:
            tiles = [t for t in tiles if t[0] < args.max_tiles and t[1] < args.max_tiles]
            positions = {(r, c) for r, c in positions if r < args.max_tiles and c < args.max_tiles}
            row_range = col_range = (0, args.max_tiles - 1)
        else:
            row_range, col_range = info["row_range"], info["col_range"]
        expected = {(r, c) for r in range(row_range[0], row_range[1] + 1) for c in range(col_range[0], col_range[1] + 1)}
        missing = expected - positions
        coverage = len(positions) / len(expected)
        print(f"  Tiles found: {len(tiles)}\n  Coverage: {coverage:.1%}")
        if coverage < args.min_coverage:
            print(f"  ⚠️  Skipping (coverage {coverage:.1%} < {args.min_coverage:.1%})")
            continue
        if args.max_tiles:
            full_shape = ((args.max_tiles - 1) * args.stride + args.tile_size,) * 2
        else:
            full_shape = infer_full_image_dimensions(positions, args.tile_size, args.stride)
        print(f"  Reconstructing {full_shape[1]}x{full_shape[0]} with {args.blend_mode} blending...")
        rgb, prob, gt, mask, counts = reconstruct_slide(model, tiles, full_shape, args.tile_size, args.stride, mean, std,
                                                        args.blend_mode, tta_mode, args.threshold, args.batch_tiles,
                                                        args.refine_kernel if args.boundary_refine else None)
        sdir = output_dir / sid
        sdir.mkdir(parents=True, exist_ok=True)
        rgb8 = (rgb * 255).astype(np.uint8)
        cv2.imwrite(str(sdir / "original_image.tif"), cv2.cvtColor(rgb8, cv2.COLOR_RGB2BGR))
        C.write_tiff_u8(sdir / "prediction_mask.tif", (prob * 255).astype(np.uint8))
        metrics = {}
        if gt is not None:
            C.write_tiff_u8(sdir / "ground_truth_mask.tif", (gt * 255).astype(np.uint8))
            metrics = api.metrics_from_counts(*counts)
            print(f"  Dice: {metrics['dice_score']:.4f}\n  IoU: {metrics['jaccard_index']:.4f}")
            cv2.imwrite(str(sdir / "gt_overlay.png"), cv2.cvtColor(C.overlay(rgb8, gt > 0.5, (255, 255, 0)), cv2.COLOR_RGB2BGR))
            cv2.imwrite(str(sdir / "pred_overlay.png"), cv2.cvtColor(C.overlay(rgb8, mask, (255, 0, 255)), cv2.COLOR_RGB2BGR))
            with open(sdir / "metrics.txt", "w") as f:
                f.write(f"Full Image Reconstruction Metrics\n{'=' * 60}\n\nSlide: {sid}\n")
                f.write(f"Image Size: {full_shape[1]} x {full_shape[0]} pixels\nTiles Used: {len(tiles)}\nCoverage: {coverage:.1%}\n\n")
                f.write(f"Reconstruction Settings:\n  Blend Mode: {args.blend_mode}\n")
                f.write(f"  TTA: {'Yes (' + args.tta_mode + ')' if args.use_tta else 'No'}\n")
                f.write(f"  Boundary Refinement: {'Yes' if args.boundary_refine else 'No'}\n  Threshold: {args.threshold}\n\nPerformance Metrics:\n")
                for label, key in (("Dice Score:    ", "dice_score"), ("IoU (Jaccard): ", "jaccard_index"), ("Sensitivity:   ", "sensitivity"),
                                   ("Specificity:   ", "specificity"), ("Precision:     ", "precision"), ("F1-Score:      ", "f1_score")):
                    f.write(f"  {label} {metrics[key]:.4f}\n")
        if args.save_metrics:
            res = {"slide_id": sid,
                   "reconstruction": {"tiles_used": len(tiles), "tiles_missing": len(missing), "coverage_ratio": coverage,
                                      "blend_mode": args.blend_mode, "tta_enabled": args.use_tta,
                                      "tta_mode": args.tta_mode if args.use_tta else None, "boundary_refined": bool(args.boundary_refine)},
                   "dimensions": {"width": full_shape[1], "height": full_shape[0],
                                  "tiles_rows": info["row_range"][1] - info["row_range"][0] + 1,
                                  "tiles_cols": info["col_range"][1] - info["col_range"][0] + 1},
                   "metrics": metrics if gt is not None else None}
            results.append(res)
            with open(output_dir / "metrics" / f"{sid}_metrics.json", "w") as f:
                json.dump(res, f, indent=2)
    if results:
        rows = [{"slide_id": r["slide_id"], "dice_score": r["metrics"]["dice_score"], "jaccard_iou": r["metrics"]["jaccard_index"],
                 "sensitivity": r["metrics"]["sensitivity"], "specificity": r["metrics"]["specificity"],
                 "precision": r["metrics"]["precision"], "tiles_used": r["reconstruction"]["tiles_used"],
                 "tiles_missing": r["reconstruction"]["tiles_missing"], "coverage": r["reconstruction"]["coverage_ratio"]}
                for r in results if r["metrics"]]
        if rows:
            with open(output_dir / "metrics" / "summary.csv", "w", newline="") as f:
                w = csv.DictWriter(f, fieldnames=list(rows[0]))
                w.writeheader(); w.writerows(rows)
            print(f"\nMean Dice: {np.mean([r['dice_score'] for r in rows]):.4f}\nMean IoU: {np.mean([r['jaccard_iou'] for r in rows]):.4f}")
        with_m = [r for r in results if r["metrics"]]
        log = {"timestamp": datetime.now().isoformat(),
               "parameters": {"weights": str(args.weights), "data_root": str(args.data_root), "tile_size": args.tile_size, "stride": args.stride,
                              "threshold": args.threshold, "blend_mode": args.blend_mode, "use_tta": args.use_tta,
                              "tta_mode": args.tta_mode if args.use_tta else None, "boundary_refine": bool(args.boundary_refine), "refine_kernel": args.refine_kernel if args.boundary_refine else None},
               "slides_processed": len(results), "slide_results": results,
               "summary_statistics": {"mean_dice": float(np.mean([r["metrics"]["dice_score"] for r in with_m])) if with_m else None,
                                      "mean_coverage": float(np.mean([r["reconstruction"]["coverage_ratio"] for r in results])),
                                      "total_tiles_used": sum(r["reconstruction"]["tiles_used"] for r in results),
                                      "total_tiles_missing": sum(r["reconstruction"]["tiles_missing"] for r in results)}}
        with open(output_dir / "reconstruction_log.json", "w") as f:
            json.dump(log, f, indent=2)
    print(f"\n✅ Reconstruction complete!\n   Output directory: {output_dir}")
    return 0


if __name__ == "__main__":
    try:
        sys.exit(main())
    except Exception as e:          # exit code 1 with a traceback, reconstruct_full_images.py:949-953
        import traceback
        print(f"\n❌ Error: {e}")
        traceback.print_exc()
        sys.exit(1)
