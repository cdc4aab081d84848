"""Test-set evaluation - drop-in for Segmentation/full_evaluation_enhanced.py (argparse :1989-2036, pipeline :1446-1958).

Tiles are predicted in device batches (+TTA / sliding window), thresholding and TP/FP/FN/TN run on the device
(adp_threshold_metrics) for the fixed threshold and for every candidate of the slide-level F1 search (:891-940); slide
aggregation and the 10 000-sample bootstrap (:983-1018, RandomState(42)) are host statistics as in the reference.
ROC / PR AUC are the reference's scikit-learn host statistics (calculate_auc_metrics, :847-888; --skip-auc leaves them NaN).
Hausdorff95 / ASSD follow
the reference's statements as written (api.boundary_metrics_from_counts: the reference samples each mask's own distance
transform on its own surface, so the values are 0.0 / inf by case).  Not reproduced: the matplotlib panels (DESIGN.md
section 6).  --boundary-refine runs BoundaryRefiner.refine on the device (adp_boundary_refine)."""
from __future__ import annotations

import argparse
import csv
import sys
import time
from collections import defaultdict
from pathlib import Path

import numpy as np

from . import common as C

METRIC_KEYS = ["dice_score", "jaccard_index", "sensitivity", "specificity", "precision", "f1_score", "accuracy"]
TABLE_NAMES = ["Dice Score", "Jaccard Index (IoU)", "Sensitivity (Recall)", "Specificity", "Precision", "F1 Score", "Accuracy",
               "ROC AUC", "PR AUC", "Hausdorff95", "ASSD"]


def extract_slide_id(tile_path: str) -> str:
    """full_evaluation_enhanced.py:658-678."""
    stem = Path(tile_path).stem
    parts = stem.split("_")
    if len(parts) >= 2 and parts[-2].startswith("r") and parts[-1].startswith("c"):
        return "_".join(parts[:-2])
    if parts[-1].startswith(("r", "c")):
        return "_".join(parts[:-1])
    return stem


def load_validation_data(val_root: str):
    """Image/mask pairs by stem, '_mask' suffix tolerated (full_evaluation_enhanced.py:1386-1443)."""
    root = Path(val_root)
    images_dir, masks_dir = root / "images", root / "masks"
    if not images_dir.exists() or not masks_dir.exists():
        raise FileNotFoundError(f"Image/mask dirs not found:\n  {images_dir}\n  {masks_dir}")
    imgs = [p for p in images_dir.rglob("*") if p.suffix.lower() in {".jpg", ".jpeg", ".png", ".tif", ".tiff"}]
    masks = [p for p in masks_dir.rglob("*") if p.suffix.lower() in {".tif", ".tiff", ".png", ".jpg", ".jpeg"}]
    by_stem = {}
    for m in masks:
        by_stem.setdefault(m.stem, m)
        if m.stem.endswith("_mask"):
            by_stem.setdefault(m.stem[:-5], m)
    pairs = [(str(i), str(by_stem[i.stem])) for i in sorted(imgs) if i.stem in by_stem]
    if not pairs:
        raise FileNotFoundError("No paired image-mask files found. Ensure stems match (optionally with '_mask' on masks).")
    print(f"Found {len(pairs)} pairs (images: {len(imgs)}, masks: {len(masks)}, unpaired images: {len(imgs) - len(pairs)})")
    return pairs


def bootstrap_ci(data: np.ndarray, n_bootstrap: int = 10000, alpha: float = 0.05, seed: int = 42):
    """full_evaluation_enhanced.py:983-1018: mean, percentile CI over resampled slides; NaNs dropped first."""
    data = np.asarray(data, dtype=np.float64)
    data = data[np.isfinite(data)]
    if len(data) == 0:
        return float("nan"), (float("nan"), float("nan"))
    rng = np.random.RandomState(seed)
    n = len(data)
    stats = np.asarray([np.mean(rng.choice(data, size=n, replace=True)) for _ in range(n_bootstrap)])
    lo, hi = np.percentile(stats, [100 * alpha / 2, 100 * (1 - alpha / 2)])
    return float(np.mean(data)), (float(lo), float(hi))


def build_parser():
    p = argparse.ArgumentParser(description="Publication-Quality Full Evaluation for Adipose U-Net")
    p.add_argument("--weights", type=str, required=True)
    p.add_argument("--test-dataset", type=str, required=True)
    p.add_argument("--output", type=str, default=None)
    p.add_argument("--ema", action="store_true", default=False)
    p.add_argument("--optimize-threshold", action="store_true", default=False)
    p.add_argument("--no-visualizations", action="store_true", default=False)
    p.add_argument("--n-vis-samples", type=int, default=10)
    p.add_argument("--use-tta", action="store_true", default=False)
    p.add_argument("--tta-mode", type=str, default="basic", choices=["minimal", "basic", "full"])
    p.add_argument("--sliding-window", action="store_true", default=False)
    p.add_argument("--overlap", type=float, default=0.5)
    p.add_argument("--blend-mode", type=str, default="gaussian", choices=["gaussian", "linear", "none", "hann"],
                   help="gaussian / linear / none as in the reference; hann is an extension (api.HannBlender)")
    p.add_argument("--boundary-refine", action="store_true", default=False)
    p.add_argument("--refine-kernel", type=int, default=5)
    p.add_argument("--adaptive-threshold", action="store_true", default=False)
    p.add_argument("--save-overlays", action="store_true", default=False)
    p.add_argument("--n-positive", type=int, default=120)
    p.add_argument("--n-negative", type=int, default=30)
    C.add_engine_args(p)
    p.add_argument("--skip-auc", action="store_true", default=False,
                   help="(not in the reference) skip the scikit-learn ROC/PR AUC host statistics (~0.3 s per 1024^2 tile on the CPU)")
    return p


def output_folder_name(args, dataset_name: str, parent_name: str) -> str:
    """full_evaluation_enhanced.py:2055-2090."""
    src = "stain" if "stain" in parent_name.lower() else "original"
    suf = []
    if args.ema:
        suf.append("ema")
    if args.use_tta:
        suf.append(f"tta_{args.tta_mode}")
    if args.sliding_window:
        s = f"sw_{args.blend_mode}"
        if args.overlap != 0.5:
            s += f"_o{int(args.overlap * 100)}"
        suf.append(s)
    if args.boundary_refine:
        suf.append("refine" + (str(args.refine_kernel) if args.refine_kernel != 5 else ""))
    if args.adaptive_threshold:
        suf.append("adaptive")
    return f"{dataset_name}_{src}" + ("_" + "_".join(suf) if suf else "")


def predict_all(model, pairs, mean, std, args):
    """-> probabilities (list of float32 HxW), ground truths (uint8 0/1)."""
    from .. import api
    preds, gts = [], []
    tta_mode = args.tta_mode if args.use_tta else None
    sw = api.SlidingWindowInference(1024, args.overlap, args.blend_mode) if args.sliding_window else None
    batch, start = [], time.time()

    def flush():
        if batch:
            preds.extend(list(model.predict_batch(np.stack(batch), mean, std, tta_mode)))
            batch.clear()

    for i, (ip, mp) in enumerate(pairs):
        img = C.read_gray(ip).astype(np.float32)
        gts.append((C.read_mask(mp) > 0).astype(np.uint8))
        if sw is not None or img.shape != (1024, 1024):
            flush()
            if sw is not None:
                preds.append(sw.predict_with_sliding_window(img, model, mean, std, use_tta=args.use_tta, tta_mode=args.tta_mode).astype(np.float32))
            else:
                preds.append(model.predict(img, mean, std, use_tta=args.use_tta, tta_mode=args.tta_mode)[0].astype(np.float32))
        else:
            batch.append(img)
            if len(batch) >= args.batch_tiles:
                flush()
        if (i + 1) % 50 == 0:
            el = time.time() - start
            print(f"  Processed {i + 1}/{len(pairs)} samples | Rate: {(i + 1) / el:.1f}/s")
    flush()
    if args.boundary_refine:                      # BoundaryRefiner.refine on every prediction (:1574-1576), on the device
        refiner = api.BoundaryRefiner(kernel_size=args.refine_kernel, engine=model.engine)
        preds = [refiner.refine(p) for p in preds]
    print(f"✓ Inference completed in {(time.time() - start) / 60:.1f} minutes")
    return preds, gts


def auc_metrics(pred: np.ndarray, true: np.ndarray):
    """calculate_auc_metrics (:847-888): pixel-level ROC AUC and average precision with scikit-learn, exactly the reference's
    host statistics (NaN when only one class is present or when scikit-learn is not installed)."""
    true_flat = (np.asarray(true) > 0.5).astype(int).ravel()
    if len(np.unique(true_flat)) < 2:
        return {"roc_auc": float("nan"), "pr_auc": float("nan")}
    try:
        from sklearn.metrics import average_precision_score, roc_auc_score
        pred_flat = np.asarray(pred).ravel()
        return {"roc_auc": float(roc_auc_score(true_flat, pred_flat)), "pr_auc": float(average_precision_score(true_flat, pred_flat))}
    except Exception as e:                      # noqa: BLE001 - the reference warns and reports NaN (:882-888)
        import warnings
        warnings.warn(f"Error calculating AUC metrics: {e}")
        return {"roc_auc": float("nan"), "pr_auc": float("nan")}


def tile_metrics(engine, pred, gt, thr):
    """calculate_pixel_metrics (:721-785): threshold + TP/FP/FN/TN on the device, ratios (and the both-empty rule) on the host."""
    from .. import api
    return api.calculate_pixel_metrics(pred, gt, thr, engine=engine)


def optimize_threshold_slide_level(engine, preds, gts, paths, thresholds):
    """full_evaluation_enhanced.py:891-940 (slide-macro F1 per candidate, first maximum wins).  The confusion counts of ALL
    candidates of a tile come from one device pass (adp_threshold_sweep) instead of one pass per candidate."""
    from .. import api
    thresholds = np.asarray(thresholds, dtype=np.float64)
    per_tile = [engine.threshold_sweep(p, g, thresholds.astype(np.float32)) for p, g in zip(preds, gts)]
    slide_of = [extract_slide_id(path) for path in paths]
    best_t, best_f1 = 0.5, -1.0
    for j, t in enumerate(thresholds):
        per_slide = defaultdict(list)
        for counts, sid in zip(per_tile, slide_of):
            per_slide[sid].append(api.metrics_from_counts(*[int(c) for c in counts[j]])["f1_score"])
        f1 = float(np.mean([np.mean(v) for v in per_slide.values()]))
        print(f"  Threshold {t:.2f}: Slide-Macro F1 = {f1:.4f}")
        if f1 > best_f1:
            best_f1, best_t = f1, float(t)
    print(f"✓ Optimal threshold: {best_t:.2f} (Slide-Macro F1 = {best_f1:.4f})")
    return best_t


def main(argv=None) -> int:
    args = build_parser().parse_args(argv)
    weights_file, ckpt_dir = C.find_weights_file(args.weights, use_ema=args.ema)
    ds = Path(args.test_dataset)
    if not ds.is_dir():
        print(f"❌ Test dataset not found: {ds}")
        return 1
    if not (ds / "images").exists() or not (ds / "masks").exists():
        print(f"❌ Dataset structure invalid. Expected:\n   {ds}/images/\n   {ds}/masks/")
        return 1
    name = ds.name
    out = Path(args.output) if args.output else Path(ckpt_dir) / "evaluation" / output_folder_name(args, name, ds.parent.name)
    out.mkdir(parents=True, exist_ok=True)
    print(f"\n{'=' * 80}\nPUBLICATION-QUALITY EVALUATION: {name.upper()} DATASET\n{'=' * 80}")
    stats_file = Path(ckpt_dir) / "normalization_stats.json"
    if not stats_file.exists():
        print(f"❌ Training normalization statistics not found: {stats_file}")
        return 1
    mean, std = C.load_normalization_stats(ckpt_dir)
    pairs = load_validation_data(str(ds))
    if C.detect_deep_supervision(ckpt_dir):
        print("🔬 Deep-supervision checkpoint: auxiliary heads are training-only; main_out is evaluated")
    model = C.make_model(weights_file, args.precision, args.device, max(args.batch_tiles, 8))
    print("✓ Model loaded successfully")
    if args.boundary_refine:
        print(f"✓ Boundary refinement enabled (kernel={args.refine_kernel})")
    print(f"\nRunning inference on {len(pairs)} samples...")
    preds, gts = predict_all(model, pairs, mean, std, args)
    paths = [p for p, _ in pairs]
    eng = model.engine
    if args.optimize_threshold or args.adaptive_threshold:
        print(f"\nOptimizing threshold on {name} set...")
        if args.adaptive_threshold:
            coarse = optimize_threshold_slide_level(eng, preds, gts, paths, np.arange(0.1, 1.0, 0.1))
            lo, hi = max(0.1, coarse - 0.1), min(0.9, coarse + 0.1)
            thr = optimize_threshold_slide_level(eng, preds, gts, paths, np.arange(lo, hi + 0.01, 0.01))
        else:
            thr = optimize_threshold_slide_level(eng, preds, gts, paths, np.arange(0.1, 0.95, 0.05))
    else:
        thr = 0.5
        print(f"Using fixed threshold: {thr}")
    from .. import api
    slides = defaultdict(list)
    for p, g, path in zip(preds, gts, paths):
        m = tile_metrics(eng, p, g, thr)
        m.update(api.boundary_metrics_from_counts(m["tp"], m["fp"], m["fn"], m["tn"]))   # calculate_boundary_metrics, :788-844
        m.update(auc_metrics(p, g) if not args.skip_auc else {"roc_auc": float("nan"), "pr_auc": float("nan")})
        slides[extract_slide_id(path)].append(m)
    slide_vals = {k: np.array([np.mean([m[k] for m in tiles]) for tiles in slides.values()]) for k in METRIC_KEYS}

    def finite_mean(tiles, key):                  # slide value = mean of the finite tile values, NaN if none (:1701-1708)
        v = [m[key] for m in tiles if np.isfinite(m[key])]
        return float(np.mean(v)) if v else float("nan")
    for key in ("roc_auc", "pr_auc", "hausdorff95", "assd"):
        slide_vals[key] = np.array([finite_mean(tiles, key) for tiles in slides.values()])
    print(f"✓ Calculated slide-level metrics for {len(slides)} slides\n\nCalculating bootstrap confidence intervals (n=10000)...")
    summary = {k: bootstrap_ci(slide_vals[k]) for k in list(METRIC_KEYS) + ["roc_auc", "pr_auc", "hausdorff95", "assd"]}
    rows = [summary[k] for k in METRIC_KEYS] + [summary["roc_auc"], summary["pr_auc"], summary["hausdorff95"], summary["assd"]]
    table = out / f"{name}_comprehensive_results.csv"
    with open(table, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["Metric", "Mean", "CI_Lower", "CI_Upper", "N_Slides", "N_Tiles", "Mean_CI"])
        for nm, (m, (lo, hi)) in zip(TABLE_NAMES, rows):
            w.writerow([nm, m, lo, hi, len(slides), len(pairs), f"{m:.4f} [{lo:.4f}, {hi:.4f}]"])
    print(f"✓ Saved results table: {table}")
    print(f"\n{'=' * 80}\nEVALUATION COMPLETE: {name.upper()}\n{'=' * 80}")
    for label, k in (("Dice Score:     ", "dice_score"), ("Jaccard (IoU):  ", "jaccard_index"), ("Precision:      ", "precision"),
                     ("Sensitivity:    ", "sensitivity"), ("Specificity:    ", "specificity")):
        m, (lo, hi) = summary[k]
        print(f"  {label} {m:.4f} (95% CI: [{lo:.4f}, {hi:.4f}])")
    print(f"  Optimal Thresh:  {thr:.3f}\n  Slides:          {len(slides)}\n  Tiles:           {len(pairs)}\n\n📂 Results saved to: {out}\n{'=' * 80}\n")
    return 0


if __name__ == "__main__":
    try:
        sys.exit(main())
    except Exception as e:          # exit code 1 with a traceback, full_evaluation_enhanced.py:2201-2215
        import traceback
        print(f"\n{'=' * 80}\n❌ EVALUATION FAILED\n{'=' * 80}\nError: {e}")
        traceback.print_exc()
        sys.exit(1)
