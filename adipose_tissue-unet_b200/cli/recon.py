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
    p.add_argument("--decode", choices=["nvjpeg", "cv2"], default="nvjpeg",
                   help="B200 engine: JPEG tiles are decoded on the device by nvJPEG (default) or on the host by OpenCV "
                        "(bit-identical to the reference's cv2.imread; nvJPEG's IDCT differs by at most a grey level or two)")
    return p


def _load_tiles(eng, paths, tile_size, decode):
    """Decode a batch of tile files to device-resident uint8 gray + RGB (reconstruct_full_images.py:362-369 reads every tile
    twice: IMREAD_GRAYSCALE for the model, IMREAD_COLOR for the mosaic).  JPEG files go through nvJPEG on the device
    (decode == 'nvjpeg'); anything else, or decode == 'cv2', is decoded by OpenCV on the host (bit-identical to the reference's
    reads) and shipped as uint8.  Returns (gray, rgb): device addresses or uint8 arrays - never float32 tiles."""
    if decode == "nvjpeg" and all(str(p).lower().endswith((".jpg", ".jpeg")) for p in paths):
        d = eng.jpeg_decode([Path(p).read_bytes() for p in paths], tile_size, want_gray=True, want_rgb=True, to_host=False)
        return d["gray_dev"], d["rgb_dev"]
    gray = np.stack([cv2.imread(str(p), cv2.IMREAD_GRAYSCALE) for p in paths])
    rgb = np.stack([cv2.cvtColor(cv2.imread(str(p), cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB) for p in paths])
    return gray, rgb


def reconstruct_slide(model, tiles_info, full_shape, tile_size, stride, mean, std, blend_mode, tta_mode, threshold, batch_tiles,
                      refine_kernel=None, decode="nvjpeg"):
    """reconstruct_slide (reconstruct_full_images.py:334-417) with every accumulation on the device: the probability
    accumulator plus four auxiliary planes sharing its weights - R, G, B of the mosaic (:411-415) and the blended ground truth
    (:404-409).  Tiles are decoded straight into device memory (nvJPEG) and pushed as uint8; no float32 tile list exists.
    -> dict(bgr8, prob, prob8, gt8 or None, mask, counts or None)."""
    from .. import api, _lib
    eng = model.engine
    H, W = full_shape
    mode = _lib.BLEND_GAUSSIAN if blend_mode in ("gaussian", "hann") else _lib.BLEND_LINEAR
    window = api.blend_window(blend_mode, tile_size)
    ops = api.TTA_OPCODES[tta_mode] if tta_mode else None
    has_gt = bool(tiles_info) and all(t[3] is not None for t in tiles_info)
    eng.wsi_begin(H, W, 0, tile_size, mode, window)
    eng.wsi_aux_begin(4 if has_gt else 3)
    for i in range(0, len(tiles_info), batch_tiles):
        chunk = tiles_info[i:i + batch_tiles]
        # edge clamp, :399-400
        ys = [min(row * stride, H - tile_size) for row, _, _, _ in chunk]
        xs = [min(col * stride, W - tile_size) for _, col, _, _ in chunk]
        gray, rgb = _load_tiles(eng, [t[2] for t in chunk], tile_size, decode)
        if refine_kernel:          # :372-380: predict (+TTA), BoundaryRefiner.refine per tile, then blend the refined tiles
            probs = eng.predict_u8_dev(gray, len(chunk), tile_size, 1, mean, std, ops)
            eng.wsi_push_probs(eng.boundary_refine(probs, kernel_size=refine_kernel), ys, xs)
        else:
            eng.wsi_push_tiles_u8(gray, ys, xs, mean, std, ops, channels=1)
        eng.wsi_push_aux(0, rgb, ys, xs, n_planes=3, is_u8=True)
        if has_gt:
            eng.wsi_push_aux(3, np.stack([C.read_mask(t[3]) for t in chunk]), ys, xs, n_planes=1)
    out = dict(gt8=None, counts=None)
    if has_gt:      # calculate_pixel_metrics(full_pred, full_gt, threshold), :745-749, on the blended ground truth
        out["prob"], out["mask"], out["counts"] = eng.wsi_finalize_auxgt(3, 0, H, W, threshold)
        out["gt8"] = eng.wsi_export_u8(3, 1, 0, H, W)                      # (full_gt * 255).astype(uint8), :745
        out["gt_mask"] = eng.wsi_export_f32(3, 0, H, W) > 0.5              # the overlay's mask (:423-455)
    else:
        out["prob"], out["mask"], _ = eng.wsi_finalize(0, H, W, threshold, gt=None, want_prob=True, want_mask=True)
    out["prob8"] = eng.wsi_export_u8(-1, 1, 0, H, W)                       # (full_pred * 255).astype(uint8), :733
    out["bgr8"] = eng.wsi_export_u8(0, 3, 0, H, W, reverse=True)           # (rgb * 255).astype(uint8) -> COLOR_RGB2BGR, :726-727
    eng.wsi_end()
    return out


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
        r = reconstruct_slide(model, tiles, full_shape, args.tile_size, args.stride, mean, std, args.blend_mode, tta_mode,
                              args.threshold, args.batch_tiles, args.refine_kernel if args.boundary_refine else None, args.decode)
        prob, mask, counts, has_gt = r["prob"], r["mask"], r["counts"], r["gt8"] is not None
        sdir = output_dir / sid
        sdir.mkdir(parents=True, exist_ok=True)
        # tifffile.imwrite(path, original_bgr, compression='lzw') (:724-728): the BGR-ordered samples go into the file as they are
        api.write_tiff_lzw(sdir / "original_image.tif", r["bgr8"])
        api.write_tiff_lzw(sdir / "prediction_mask.tif", r["prob8"])
        rgb8 = r["bgr8"][:, :, ::-1]
        metrics = {}
        if has_gt:
            api.write_tiff_lzw(sdir / "ground_truth_mask.tif", r["gt8"])
            metrics = api.metrics_from_counts(*counts)
            print(f"  Dice: {metrics['dice_score']:.4f}\n  IoU: {metrics['jaccard_index']:.4f}")
            cv2.imwrite(str(sdir / "gt_overlay.png"), cv2.cvtColor(C.overlay(np.ascontiguousarray(rgb8), r["gt_mask"], (255, 255, 0)), cv2.COLOR_RGB2BGR))
            cv2.imwrite(str(sdir / "pred_overlay.png"), cv2.cvtColor(C.overlay(np.ascontiguousarray(rgb8), mask, (255, 0, 255)), cv2.COLOR_RGB2BGR))
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
                   "metrics": metrics if has_gt else None}
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
