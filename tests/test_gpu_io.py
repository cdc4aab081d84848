"""GPU parity of the tile I/O front-end: nvJPEG decode against the reference's cv2.imread,
device-side RGB mosaic / blended ground truth against the oracle's blend statements, uint8 exports, fat-%."""
import cv2
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import adipose_unet_b200 as A
from adipose_unet_b200 import api, _lib
from oracle import geometry as G
from oracle import unet as U

MEAN, STD = A.synth.DEFAULT_MEAN, A.synth.DEFAULT_STD


@pytest.fixture(scope="module")
def eng():
    e = api.Engine(precision="fp32", max_forwards=8)
    e.set_weights(A.synth.init_weights())
    yield e
    e.close()


def _jpeg(img_bgr_or_gray, quality, sampling=None):
    params = [cv2.IMWRITE_JPEG_QUALITY, quality]
    if sampling is not None:
        params += [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, sampling]
    ok, buf = cv2.imencode(".jpg", img_bgr_or_gray, params)
    assert ok
    return buf.tobytes()


def test_nvjpeg_decode_vs_cv2_imread(eng):
    """The reference reads tiles with cv2.imread(path, IMREAD_GRAYSCALE) for the model and IMREAD_COLOR for the mosaic
    (reconstruct_full_images.py:362-369).  nvJPEG's inverse DCT and chroma upsampling are not libjpeg-turbo's, so the decode
    is NOT bit-identical: this test states by how much (asserted bounds = measured + margin) and what it does to the
    prediction - it is why the recon CLI keeps --decode cv2."""
    S = 512
    rgb = A.synth.rgb_tile(S, seed=5)
    gray3 = cv2.cvtColor(A.synth.ecm_tile(S, seed=6), cv2.COLOR_GRAY2BGR)
    blobs, names = [], []
    for q in (75, 95, 100):
        blobs.append(_jpeg(cv2.cvtColor(rgb, cv2.COLOR_RGB2BGR), q)); names.append(f"colour q{q} 4:2:0")
        blobs.append(_jpeg(gray3, q)); names.append(f"gray-as-BGR q{q}")
    blobs.append(_jpeg(cv2.cvtColor(rgb, cv2.COLOR_RGB2BGR), 95, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444)); names.append("colour q95 4:4:4")
    blobs.append(_jpeg(A.synth.ecm_tile(S, seed=7), 95)); names.append("1-component gray q95")
    d = eng.jpeg_decode(blobs, S, want_gray=True, want_rgb=True)
    worst_g, worst_c = 0, 0
    for i, (b, nm) in enumerate(zip(blobs, names)):
        arr = np.frombuffer(b, np.uint8)
        ref_g = cv2.imdecode(arr, cv2.IMREAD_GRAYSCALE)
        ref_c = cv2.cvtColor(cv2.imdecode(arr, cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB)
        dg = np.abs(d["gray"][i].astype(int) - ref_g.astype(int))
        dc = np.abs(d["rgb"][i].astype(int) - ref_c.astype(int))
        print(f"nvJPEG vs cv2 [{nm}]: gray max {dg.max()} levels, exact {float((dg == 0).mean()):.4f}; rgb max {dc.max()}, exact {float((dc == 0).mean()):.4f}")
        worst_g, worst_c = max(worst_g, int(dg.max())), max(worst_c, int(dc.max()))
        assert dg.max() <= 2 and (dg == 0).mean() >= 0.80, nm
        assert dc.max() <= 16 and (dc <= 2).mean() >= 0.95, nm
    # effect on the model output: same tile, both decodes, fp32 path
    arr = np.frombuffer(blobs[2], np.uint8)                                  # colour q95
    ref_g = cv2.imdecode(arr, cv2.IMREAD_GRAYSCALE)
    p_ref = eng.predict(ref_g[None], MEAN, STD)[0]
    p_dev = eng.predict_u8_dev(eng.jpeg_decode([blobs[2]], S, want_gray=True, to_host=False)["gray_dev"], 1, S, 1, MEAN, STD)[0]
    dp = float(np.abs(p_ref - p_dev).max())
    a, b = p_ref > 0.5, p_dev > 0.5
    dice = (2.0 * (a & b).sum() + 1e-10) / (a.sum() + b.sum() + 1e-10)
    print(f"prediction from nvJPEG-decoded vs cv2-decoded tile (fp32, random-init weights): max|dp| = {dp:.2e}, mask-dice = {dice:.5f}")
    assert dp <= 2e-2


def test_rgb_mosaic_and_blended_ground_truth_vs_oracle(eng):
    """Device-side auxiliary planes: three blender.reconstruct calls over the colour channels (float32 tile / 255,
    reconstruct_full_images.py:368, 411-415), the blended ground truth (:404-409), (x * 255).astype(uint8) exports
    (:726, 733, 745) and calculate_pixel_metrics on the blended truth (:749) - bit-exact against the NumPy statements."""
    S, H, W = 128, 300, 420
    rng = np.random.default_rng(9)
    for blend, overlap in (("gaussian", 0.5), ("linear", 0.75)):
        stride = G.stride_for(S, overlap)
        pos = G.tile_positions(H, W, S, stride)
        rgb_tiles = rng.integers(0, 256, size=(len(pos), S, S, 3), dtype=np.uint8)
        gt_tiles = (rng.random((len(pos), S, S)) < 0.4).astype(np.float32)
        gray = rng.integers(0, 256, size=(len(pos), S, S), dtype=np.uint8)
        win = G.gaussian_window(S) if blend == "gaussian" else None
        mode = _lib.BLEND_GAUSSIAN if blend == "gaussian" else _lib.BLEND_LINEAR
        ys, xs = [p[0] for p in pos], [p[1] for p in pos]
        eng.wsi_begin(H, W, 0, S, mode, win)
        eng.wsi_aux_begin(4)
        for i in range(0, len(pos), 7):
            sl = slice(i, i + 7)
            eng.wsi_push_tiles_u8(gray[sl], ys[sl], xs[sl], MEAN, STD, None, channels=1)
            eng.wsi_push_aux(0, rgb_tiles[sl], ys[sl], xs[sl], n_planes=3)
            eng.wsi_push_aux(3, gt_tiles[sl], ys[sl], xs[sl], n_planes=1)
        prob, mask, counts = eng.wsi_finalize_auxgt(3, 0, H, W, 0.5)
        bgr8 = eng.wsi_export_u8(0, 3, 0, H, W, reverse=True)
        rgb8 = eng.wsi_export_u8(0, 3, 0, H, W)
        gt8 = eng.wsi_export_u8(3, 1, 0, H, W)
        gtf = eng.wsi_export_f32(3, 0, H, W)
        prob8 = eng.wsi_export_u8(-1, 1, 0, H, W)
        # the same pushes through the float32 tile path give the same probabilities, bit for bit
        eng.wsi_end()
        eng.wsi_begin(H, W, 0, S, mode, win)
        eng.wsi_push_tiles(gray.astype(np.float32), ys, xs, MEAN, STD, None)
        prob_f, _, _ = eng.wsi_finalize(0, H, W, want_mask=False)
        eng.wsi_end()
        np.testing.assert_array_equal(prob, prob_f)
        recon = (lambda t: G.gaussian_reconstruct(t, pos, (H, W), win)) if blend == "gaussian" else (lambda t: G.linear_reconstruct(t, pos, (H, W)))
        tiles_f = [(t.astype(np.float32) / 255.0) for t in rgb_tiles]                     # cv2 tile .astype(float32) / 255.0
        full = np.zeros((H, W, 3), np.float32)
        for ch in range(3):
            full[:, :, ch] = recon([t[:, :, ch] for t in tiles_f])
        want_rgb8 = (full * 255).astype(np.uint8)
        np.testing.assert_array_equal(rgb8, want_rgb8)
        np.testing.assert_array_equal(bgr8, cv2.cvtColor(want_rgb8, cv2.COLOR_RGB2BGR))
        full_gt = recon(list(gt_tiles))
        np.testing.assert_array_equal(gtf, full_gt)
        np.testing.assert_array_equal(gt8, (full_gt * 255).astype(np.uint8))
        np.testing.assert_array_equal(prob8, (prob * 255).astype(np.uint8))
        m = G.pixel_metrics(prob, full_gt, 0.5)
        assert counts == (m["tp"], m["fp"], m["fn"], m["tn"])
        np.testing.assert_array_equal(mask, (prob > 0.5).astype(np.uint8))


def test_fat_percentage_and_classification(eng):
    """calculate_fat_percentage / classify_tile (tile_classification_evaluation.py:211-239)."""
    rng = np.random.default_rng(3)
    p = rng.random((5, 200, 300)).astype(np.float32)
    p[1] = 0.0; p[2] = 1.0; p[3, :100] = 0.5        # exactly at the threshold: not fat (strict >)
    got = eng.fat_percent(p, 0.5)
    want = np.array([((t > 0.5).astype(np.uint8).sum() / t.size) * 100.0 for t in p])
    np.testing.assert_array_equal(got, want)
    assert api.calculate_fat_percentage(p[0], 0.5, engine=eng) == want[0]
    assert api.classify_tile(want[0], want[0]) == "Has Fat" and api.classify_tile(0.0, 1e-9) == "No Fat"


def test_tiff_writer_threads_on_this_box(tmp_path):
    """The parallel strip writer on the GPU box's host cores (timing printed for the record)."""
    import time
    a = (cv2.GaussianBlur(np.random.default_rng(0).random((8192, 8192)).astype(np.float32), (0, 0), 8) > 0.5).astype(np.uint8) * 255
    for th in (1, 0):
        t0 = time.perf_counter()
        api.write_tiff_lzw(tmp_path / "m.tif", a, threads=th)
        dt = time.perf_counter() - t0
        print(f"TIFF-LZW 8192^2 mask, threads={'all' if th == 0 else th}: {dt * 1e3:.0f} ms = {a.size / dt / 1e6:.0f} MB/s")
    np.testing.assert_array_equal(cv2.imread(str(tmp_path / "m.tif"), cv2.IMREAD_UNCHANGED), a)
