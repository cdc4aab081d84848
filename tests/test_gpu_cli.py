"""End-to-end runs of the re-hosted CLIs on the GPU with a small synthetic dataset:
train (two short phases, checkpoints + side files) -> evaluate -> folder inference -> whole-slide reconstruction.
The reconstruction's probability map is checked against the oracle's blend of the engine's own per-tile predictions."""
import csv
import json

import cv2
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import adipose_unet_b200 as A
from adipose_unet_b200 import api
from adipose_unet_b200.cli import evaluate, infer, recon, train
from adipose_unet_b200.weights_io import load_weights_file, save_weights_file
from oracle import geometry as G

T = 1024


def _write_dataset(root, n_rows, n_cols, stride, seed=1):
    """A synthetic slide cut into overlapping 1024^2 tiles named {slide}_r{r}_c{c}.jpg + masks/*.tif."""
    (root / "images").mkdir(parents=True); (root / "masks").mkdir(parents=True)
    H, W = (n_rows - 1) * stride + T, (n_cols - 1) * stride + T
    big = cv2.resize(A.synth.ecm_tile(512, seed=seed), (W, H), interpolation=cv2.INTER_LINEAR)
    gt = (big > 140).astype(np.uint8)
    for r in range(n_rows):
        for c in range(n_cols):
            y, x = r * stride, c * stride
            tile = cv2.cvtColor(big[y:y + T, x:x + T], cv2.COLOR_GRAY2BGR)
            cv2.imwrite(str(root / "images" / f"slideA_r{r}_c{c}.jpg"), tile, [cv2.IMWRITE_JPEG_QUALITY, 95])
            cv2.imwrite(str(root / "masks" / f"slideA_r{r}_c{c}.tif"), gt[y:y + T, x:x + T])
    return H, W


@pytest.fixture(scope="module")
def workspace(tmp_path_factory):
    ws = tmp_path_factory.mktemp("cli")
    _write_dataset(ws / "build" / "dataset" / "train", 1, 3, 512, seed=1)
    _write_dataset(ws / "build" / "dataset" / "val", 1, 2, 512, seed=2)
    return ws


def test_train_then_eval_infer_recon(workspace, capsys):
    ws = workspace
    rc = train.main(["--data-root", str(ws / "build"), "--pretrained-weights", str(ws / "none.h5"), "--batch-size", "1",
                     "--epochs-phase1", "1", "--epochs-phase2", "1", "--normalization-method", "zscore",
                     "--checkpoint-root", str(ws / "ckpt"), "--max-steps-per-epoch", "2"])      # the reference's DEFAULT recipe:
    # deep supervision + hard-example mining + cosine schedule
    assert rc == 0
    ckpts = list((ws / "ckpt").iterdir())
    assert len(ckpts) == 1 and ckpts[0].name.endswith("_adipose_sybreosin_1024_finetune_v3")
    ck = ckpts[0]
    for f in ("phase1_best.weights.h5", "phase2_best.weights.h5", "weights_best_overall.weights.h5", "weights_ema.weights.h5",
              "normalization_stats.json",
              "training_settings.log", "phase1_training.log", "phase2_training.log"):
        assert (ck / f).exists(), f
    stats = json.loads((ck / "normalization_stats.json").read_text())
    assert set(stats) == {"mean", "std", "normalization_method", "dataset_path", "num_training_images", "build_timestamp", "version"}
    assert "use_deep_supervision: True" in (ck / "training_settings.log").read_text()
    wa = load_weights_file(str(ck / "weights_best_overall.weights.h5"), keep_aux=True)
    assert wa["aux_out1/kernel"].shape == (1, 1, 176, 1) and wa["aux_out2/kernel"].shape == (1, 1, 88, 1)
    w = load_weights_file(str(ck / "weights_best_overall.weights.h5"))
    assert "aux_out1/kernel" not in w                       # inference loads the 22 graph layers only
    assert w["down1_conv1/kernel"].shape == (3, 3, 1, 44) and w["output_softmax/kernel"].shape == (1, 1, 44, 2)
    rows = list(csv.DictReader(open(ck / "phase2_training.log")))
    assert len(rows) == 1 and np.isfinite(float(rows[0]["val_main_out_dice_coef"])) and 0.0 <= float(rows[0]["main_out_binary_accuracy"]) <= 1.0

    # evaluate on the validation split with the checkpoint DIRECTORY (weights discovery) and threshold search
    val = ws / "build" / "dataset" / "val"
    rc = evaluate.main(["--weights", str(ck), "--test-dataset", str(val), "--optimize-threshold", "--use-tta", "--tta-mode", "minimal",
                        "--no-visualizations"])
    assert rc == 0
    table = ck / "evaluation" / "val_original_tta_minimal" / "val_comprehensive_results.csv"
    rows = list(csv.DictReader(open(table)))
    assert [r["Metric"] for r in rows][:3] == ["Dice Score", "Jaccard Index (IoU)", "Sensitivity (Recall)"] and len(rows) == 11
    assert rows[0]["N_Tiles"] == "2" and rows[0]["N_Slides"] == "1" and 0.0 <= float(rows[0]["Mean"]) <= 1.0

    # folder inference
    out = ws / "infer_out"
    rc = infer.main(["--images-dir", str(val / "images"), "--output-dir", str(out), "--weights", str(ck), "--save-probability",
                     "--save-overlays", "--threshold", "0.5"])
    assert rc == 0
    m = cv2.imread(str(out / "masks" / "slideA_r0_c0_mask.tif"), cv2.IMREAD_UNCHANGED)
    p = cv2.imread(str(out / "probabilities" / "slideA_r0_c0_prob.tif"), cv2.IMREAD_UNCHANGED)
    assert m.shape == (T, T) and set(np.unique(m)) <= {0, 1} and p.dtype == np.uint8
    np.testing.assert_array_equal(m[p > 128], 1)              # p*255 truncated > 128 implies p > 0.5
    assert (out / "overlays" / "slideA_r0_c0_overlay.png").exists()

    # whole-slide reconstruction of the validation slide (2 tiles, 50 % overlap) against the oracle's blend
    rout = ws / "recon_out"
    rc = recon.main(["--weights", str(ck / "weights_best_overall.weights.h5"), "--data-root", str(val), "--output-dir", str(rout),
                     "--stride", "512", "--blend-mode", "gaussian", "--decode", "cv2"])      # host libjpeg: the reference's exact tile bytes
    assert rc == 0
    sdir = rout / "slideA"
    for f in ("original_image.tif", "prediction_mask.tif", "ground_truth_mask.tif", "gt_overlay.png", "pred_overlay.png", "metrics.txt"):
        assert (sdir / f).exists(), f
    res = json.loads((rout / "metrics" / "slideA_metrics.json").read_text())
    assert res["dimensions"] == {"width": 1536, "height": 1024, "tiles_rows": 1, "tiles_cols": 2}
    assert res["reconstruction"]["tiles_used"] == 2 and res["reconstruction"]["coverage_ratio"] == 1.0
    assert (rout / "metrics" / "summary.csv").exists() and (rout / "reconstruction_log.json").exists()
    model = api.AdiposeUNet(); model.build_model(); model.load_weights(str(ck / "weights_best_overall.weights.h5"))
    tiles = [cv2.imread(str(val / "images" / f"slideA_r0_c{c}.jpg"), cv2.IMREAD_GRAYSCALE).astype(np.float32) for c in range(2)]
    preds = [model.predict_single(t, stats["mean"], stats["std"]) for t in tiles]
    want = G.gaussian_reconstruct(preds, [(0, 0), (0, 512)], (1024, 1536), G.gaussian_window(T))
    got = cv2.imread(str(sdir / "prediction_mask.tif"), cv2.IMREAD_UNCHANGED)
    assert np.abs(got.astype(np.int32) - (want * 255).astype(np.uint8).astype(np.int32)).max() <= 1

    # the same slide with --boundary-refine (BoundaryRefiner.refine per tile before blending, reconstruct_full_images.py:378-380)
    # and the Hann blending extension: against the oracle's refine + blend statements
    from oracle import refine as R
    rout2 = ws / "recon_refined"
    rc = recon.main(["--weights", str(ck / "weights_best_overall.weights.h5"), "--data-root", str(val), "--output-dir", str(rout2),
                     "--stride", "512", "--blend-mode", "hann", "--boundary-refine", "--refine-kernel", "5", "--decode", "cv2"])
    assert rc == 0
    log = json.loads((rout2 / "reconstruction_log.json").read_text())
    assert log["parameters"]["boundary_refine"] is True and log["parameters"]["refine_kernel"] == 5
    want2 = G.hann_reconstruct([R.refine(p) for p in preds], [(0, 0), (0, 512)], (1024, 1536), G.hann_window(T))
    got2 = cv2.imread(str(rout2 / "slideA" / "prediction_mask.tif"), cv2.IMREAD_UNCHANGED)
    assert np.abs(got2.astype(np.int32) - (want2 * 255).astype(np.uint8).astype(np.int32)).max() <= 1


    # default decode path: nvJPEG on the device.  Its IDCT is not libjpeg-turbo's, so tile bytes may differ by a grey level
    # (tests/test_gpu_io.py measures it); the reconstructed probability map must stay within a few levels (of 255) of the host-decoded run.
    rout3 = ws / "recon_nvjpeg"
    rc = recon.main(["--weights", str(ck / "weights_best_overall.weights.h5"), "--data-root", str(val), "--output-dir", str(rout3),
                     "--stride", "512", "--blend-mode", "gaussian"])
    assert rc == 0
    got3 = cv2.imread(str(rout3 / "slideA" / "prediction_mask.tif"), cv2.IMREAD_UNCHANGED)
    d3 = np.abs(got3.astype(np.int32) - got.astype(np.int32))
    print("recon nvJPEG vs cv2 decode: prediction_mask.tif max diff", int(d3.max()), "levels, differing px", float((d3 > 0).mean()))
    assert d3.max() <= 5          # measured 2-3 of 255 levels on the bf16 path (a one-level input change moves bf16 roundings)
    o1 = cv2.imread(str(sdir / "original_image.tif"), cv2.IMREAD_COLOR).astype(np.int32)
    o3 = cv2.imread(str(rout3 / "slideA" / "original_image.tif"), cv2.IMREAD_COLOR).astype(np.int32)
    assert o1.shape == o3.shape == (1024, 1536, 3) and np.abs(o1 - o3).max() <= 3
    g1 = cv2.imread(str(sdir / "ground_truth_mask.tif"), cv2.IMREAD_UNCHANGED)
    g3 = cv2.imread(str(rout3 / "slideA" / "ground_truth_mask.tif"), cv2.IMREAD_UNCHANGED)
    np.testing.assert_array_equal(g1, g3)
