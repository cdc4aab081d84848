"""Oracle (oracle/geometry.py) vs vectors produced by the reference's own NumPy code
(tests/golden/make_golden.py) — SURVEY.md section 8a rows G1-G4, M3."""
import hashlib

import numpy as np
import pytest

from oracle import geometry as G


def test_tile_positions_match_reference(golden):
    cases = golden["pos_cases"]
    for i, (h, w, t, ov) in enumerate(cases):
        h, w, t = int(h), int(w), int(t)
        stride = G.stride_for(t, float(ov))
        assert stride == int(golden[f"pos_stride_{i}"])
        pos = np.array(G.tile_positions(h, w, t, stride), dtype=np.int64).reshape(-1, 2)
        np.testing.assert_array_equal(pos, golden[f"pos_{i}"])


def test_tile_counts_of_baseline_configs():
    assert len(G.tile_positions(32768, 32768, 1024, G.stride_for(1024, 0.5))) == 3969
    assert len(G.tile_positions(16384, 16384, 1024, G.stride_for(1024, 0.75))) == 3721
    assert G.tile_positions(1000, 1000, 1024, 512) == []
    assert G.tile_positions(1024, 1024, 1024, 512) == [(0, 0)]


def test_recon_inverse_geometry(golden):
    for name, sid, rc in zip(golden["parse_names"], golden["parse_ids"], golden["parse_rc"]):
        s, r, c = G.parse_tile_filename(str(name))
        assert s == str(sid) and (r, c) == tuple(int(v) for v in rc)
    with pytest.raises(ValueError):
        G.parse_tile_filename("no_position.jpg")
    assert G.infer_full_dims({(0, 0), (3, 5), (2, 1)}, 1024, 512) == tuple(int(v) for v in golden["infer_dims"])
    assert G.recon_tile_origin(5, 3, 1024, 512, 3000, 2000) == (min(5 * 512, 3000 - 1024), min(3 * 512, 2000 - 1024))


@pytest.mark.parametrize("t,sf", [(1024, 0.25), (128, 0.25), (256, 0.5), (64, 0.25)])
def test_gaussian_window_bit_exact(golden, t, sf):
    w = G.gaussian_window(t, sf)
    tag = f"gauss_{t}_{int(sf * 100)}"
    assert w.dtype == np.float32
    assert hashlib.sha256(np.ascontiguousarray(w).tobytes()).hexdigest() == str(golden[tag + "_sha256"])
    np.testing.assert_array_equal(w[[0, 1, t // 2 - 1, t // 2, t // 2 + 1, t - 1], :], golden[tag + "_rows"])
    np.testing.assert_array_equal(np.diagonal(w), golden[tag + "_diag"])


def test_gaussian_window_probes():
    w = G.gaussian_window(1024, 0.25)
    assert w[512, 512] == np.float32(1.0)
    assert abs(float(w[0, 0]) - 0.0183156) < 1e-6
    assert abs(float(w[1023, 1023]) - 0.0186038) < 1e-6
    assert abs(float(w[0, 512]) - 0.135335) < 1e-6


@pytest.mark.parametrize("tag", ["a", "b"])
def test_blenders_bit_exact(golden, tag):
    tiles = [t.astype(np.float32) for t in golden[f"blend_{tag}_tiles"]]
    pos = [tuple(int(v) for v in p) for p in golden[f"blend_{tag}_pos"]]
    shape = tuple(int(v) for v in golden[f"blend_{tag}_shape"])
    win = G.gaussian_window(64, 0.25)
    np.testing.assert_array_equal(G.gaussian_reconstruct(tiles, pos, shape, win), golden[f"blend_{tag}_gauss"])
    np.testing.assert_array_equal(G.linear_reconstruct(tiles, pos, shape), golden[f"blend_{tag}_linear"])


@pytest.mark.parametrize("mode", ["minimal", "basic", "full"])
def test_tta_transforms(golden, mode):
    n = 8
    ramp = np.arange(n * n, dtype=np.float32).reshape(n, n)
    augs = golden[f"tta_{mode}_aug"]
    deaugs = golden[f"tta_{mode}_deaug"]
    assert len(G.TTA_MODES[mode]) == len(augs) == len(G.TTA_OPCODES[mode])
    for k, ((a, d), op) in enumerate(zip(G.TTA_MODES[mode], G.TTA_OPCODES[mode])):
        np.testing.assert_array_equal(a(ramp), augs[k])
        np.testing.assert_array_equal(d(ramp), deaugs[k])
        # device op codes: aug == d4_apply(op), deaug == d4_apply(inverse op)
        np.testing.assert_array_equal(G.d4_apply(op, ramp), augs[k])
        np.testing.assert_array_equal(G.d4_apply(G.D4_INVERSE[op], ramp), deaugs[k])
        np.testing.assert_array_equal(G.d4_apply(G.D4_INVERSE[op], G.d4_apply(op, ramp)), ramp)


@pytest.mark.parametrize("mode", ["minimal", "basic", "full"])
def test_tta_loop_with_fake_model(golden, fake_model, mode):
    img = golden["tta_img"]
    preds = []
    for aug, deaug in G.TTA_MODES[mode]:
        preds.append(deaug(fake_model.predict_single(aug(img), 127.5, 50.0)).astype(np.float32))
    np.testing.assert_array_equal(G.tta_mean(preds), golden[f"tta_{mode}_avg"])
    # sequential float32 sum then divide == np.mean over axis 0 (what the device kernel does)
    acc = preds[0].copy()
    for p in preds[1:]:
        acc = acc + p
    np.testing.assert_array_equal((acc / np.float32(len(preds))).astype(np.float32), golden[f"tta_{mode}_avg"])


@pytest.mark.parametrize("blend", ["gaussian", "linear"])
def test_sliding_window_with_fake_model(golden, fake_model, blend):
    img = golden["sw_img"]
    t, stride = 64, G.stride_for(64, 0.5)
    pos = G.tile_positions(img.shape[0], img.shape[1], t, stride)
    preds = []
    for (y, x) in pos:
        tile = img[y:y + t, x:x + t]
        ps = [deaug(fake_model.predict_single(aug(tile), 127.5, 50.0)).astype(np.float32) for aug, deaug in G.TTA_FULL]
        preds.append(G.tta_mean(ps))
    if blend == "gaussian":
        out = G.gaussian_reconstruct(preds, pos, img.shape, G.gaussian_window(64))
    else:
        out = G.linear_reconstruct(preds, pos, img.shape)
    np.testing.assert_array_equal(out, golden[f"sw_{blend}"])


def test_threshold_and_metrics(golden):
    pred, gt = golden["met_pred"], golden["met_gt"]
    np.testing.assert_array_equal(G.binarize(pred, 0.5), golden["met_bin"])
    keys = [str(k) for k in golden["met_keys"]]
    cases = {"rand": (pred, gt, 0.5), "thr7": (pred, gt, 0.7),
             "empty": (np.zeros((16, 16), np.float32), np.zeros((16, 16), np.uint8), 0.5),
             "nopred": (np.zeros((16, 16), np.float32), np.ones((16, 16), np.uint8), 0.5)}
    for tag, (p, g, thr) in cases.items():
        m = G.pixel_metrics(p, g, thr)
        np.testing.assert_array_equal(np.array([float(m[k]) for k in keys]), golden[f"met_{tag}"])
        m2 = G.metrics_from_counts(m["tp"], m["fp"], m["fn"], m["tn"])
        assert m2 == m
    assert G.prob_to_u8(np.array([0.0, 0.5, 0.999, 1.0], np.float32)).tolist() == [0, 127, 254, 255]


def test_hann_window_extension_properties():
    """Hann blending is an extension (the reference has Gaussian / linear only): formula h[i] = 0.5 - 0.5 cos(2 pi (i + 0.5) / T)."""
    for T in (64, 1024):
        w = G.hann_window(T)
        assert w.dtype == np.float32 and w.shape == (T, T)
        assert w.min() >= 0.99e-6                               # every weight above the blender's 1e-8 normalisation clamp
        np.testing.assert_array_equal(w, w[::-1, ::-1])         # symmetric (half-sample shift)
        np.testing.assert_array_equal(w, w.T)
        h = np.maximum(0.5 - 0.5 * np.cos(2 * np.pi * (np.arange(T) + 0.5) / T), G.HANN_FLOOR)
        np.testing.assert_allclose(h[: T // 2] + h[T // 2:], 1.0, atol=1.1e-3)  # partition of unity at 50 % overlap (floored ends aside)
        inner = slice(T // 8, T // 2 - T // 8)
        np.testing.assert_allclose(h[: T // 2][inner] + h[T // 2:][inner], 1.0, atol=1e-15)
    # a single 1024^2 tile: the corner weights (5.5e-12 without the floor) stay above the clamp, the blend returns the tile
    t = np.full((1024, 1024), 0.5, np.float32)
    np.testing.assert_allclose(G.hann_reconstruct([t], [(0, 0)], (1024, 1024)), 0.5, rtol=2e-6)
    # a constant field blends to the same constant everywhere, borders included; interior weight sum == 1
    T, H, W = 64, 192, 256
    pos = [(y, x) for y in range(0, H - T + 1, T // 2) for x in range(0, W - T + 1, T // 2)]
    tiles = [np.full((T, T), 0.625, np.float32) for _ in pos]
    res, acc, wsum = G.hann_reconstruct(tiles, pos, (H, W), return_parts=True)
    np.testing.assert_allclose(res, 0.625, rtol=2e-6)
    np.testing.assert_allclose(wsum[T // 2:H - T // 2, T // 2:W - T // 2], 1.0, rtol=2.5e-3)      # 1 + the floored ends' excess
    assert wsum.min() > 0


def test_boundary_metrics_case_analysis_matches_reference_statement():
    """calculate_boundary_metrics as written in the reference (own distance transform on own surface) reduces to a case
    analysis on the confusion counts: api.boundary_metrics_from_counts against the literal scipy restatement."""
    from adipose_unet_b200.api import boundary_metrics_from_counts
    rng = np.random.default_rng(5)
    h, w = 48, 64
    blobs = rng.random((h, w))
    cases = [(np.zeros((h, w)), np.zeros((h, w))), (np.ones((h, w)), np.zeros((h, w))), (np.zeros((h, w)), np.ones((h, w))),
             (np.ones((h, w)), np.ones((h, w))), (np.ones((h, w)), (blobs > 0.5) * 1.0), ((blobs > 0.3) * 1.0, np.ones((h, w))),
             (blobs, (rng.random((h, w)) > 0.6) * 1.0), ((blobs > 0.9) * 0.8, (blobs > 0.2) * 1.0)]
    edge = np.zeros((h, w)); edge[0, :] = 1.0; edge[:, -1] = 1.0            # touches the border only
    cases.append((edge, edge.T[:h, :w] if edge.T.shape == (h, w) else edge[::-1]))
    for pred, true in cases:
        pb, tb = pred > 0.5, true > 0.5
        counts = (int((pb & tb).sum()), int((pb & ~tb).sum()), int((~pb & tb).sum()), int((~pb & ~tb).sum()))
        assert boundary_metrics_from_counts(*counts) == G.boundary_metrics_reference(pred.astype(np.float32), true.astype(np.float32))


def test_host_windows_equal_oracle_statements():
    """The host-side window builders of the product (api.py) against the oracle statements, no GPU needed."""
    from adipose_unet_b200 import api
    for T in (64, 1024):
        np.testing.assert_array_equal(api.blend_window("gaussian", T), G.gaussian_window(T))
        np.testing.assert_array_equal(api.blend_window("hann", T), G.hann_window(T))
    assert api.blend_window("linear", 64) is None
    np.testing.assert_array_equal(api.HannBlender(128).weight_map, G.hann_window(128))
    np.testing.assert_array_equal(api.GaussianBlender(128, 0.25).weight_map, G.gaussian_window(128, 0.25))
