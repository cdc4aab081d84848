"""GPU parity of the post-processing kernels against the vectors the reference's own NumPy code
produced (tests/golden) — bit-exact — and against the oracle for the loss."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from adipose_unet_b200 import api
from oracle import geometry as G
from oracle import unet as U


@pytest.fixture(scope="module")
def eng():
    return api.default_engine()


@pytest.mark.parametrize("mode", ["minimal", "basic", "full"])
def test_tta_combine_bit_exact_vs_reference(golden, fake_model, mode):
    img = golden["tta_img"]
    avg, info = api.TestTimeAugmentation(mode).predict_with_tta(fake_model, img, 127.5, 50.0)
    np.testing.assert_array_equal(avg, golden[f"tta_{mode}_avg"])
    assert info["num_augmentations"] == len(G.TTA_OPCODES[mode])


@pytest.mark.parametrize("size", [40, 96, 1024])
def test_tta_combine_odd_sizes(eng, size):
    rng = np.random.default_rng(size)
    planes = rng.random((8, size, size), dtype=np.float32)
    out = eng.tta_combine(planes, G.TTA_OPCODES["full"])
    ref = G.tta_mean([G.d4_apply(G.D4_INVERSE[op], planes[k]) for k, op in enumerate(G.TTA_OPCODES["full"])])
    np.testing.assert_array_equal(out, ref)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_blenders_bit_exact_vs_reference(golden, tag):
    tiles = [t.astype(np.float32) for t in golden[f"blend_{tag}_tiles"]]
    pos = [tuple(int(v) for v in p) for p in golden[f"blend_{tag}_pos"]]
    shape = tuple(int(v) for v in golden[f"blend_{tag}_shape"])
    np.testing.assert_array_equal(api.GaussianBlender(tile_size=64).reconstruct(tiles, pos, shape), golden[f"blend_{tag}_gauss"])
    np.testing.assert_array_equal(api.LinearBlender().reconstruct(tiles, pos, shape), golden[f"blend_{tag}_linear"])


@pytest.mark.parametrize("blend", ["gaussian", "linear"])
def test_sliding_window_foreign_model_bit_exact(golden, fake_model, blend):
    sw = api.SlidingWindowInference(tile_size=64, overlap=0.5, blend_mode=blend, verbose=False)
    out = sw.predict_with_sliding_window(golden["sw_img"], fake_model, 127.5, 50.0, use_tta=True, tta_mode="full")
    np.testing.assert_array_equal(out, golden[f"sw_{blend}"])


def test_hann_blender_extension_vs_oracle(golden, fake_model):
    """Hann blending (extension): the device blend with the Hann window equals the oracle's statement bit for bit."""
    np.testing.assert_array_equal(api.HannBlender(64).weight_map, G.hann_window(64))
    np.testing.assert_array_equal(api.blend_window("hann", 1024), G.hann_window(1024))
    tiles = [t.astype(np.float32) for t in golden["blend_a_tiles"]]
    pos = [tuple(int(v) for v in p) for p in golden["blend_a_pos"]]
    shape = tuple(int(v) for v in golden["blend_a_shape"])
    np.testing.assert_array_equal(api.HannBlender(tile_size=64).reconstruct(tiles, pos, shape), G.hann_reconstruct(tiles, pos, shape))
    sw = api.SlidingWindowInference(tile_size=64, overlap=0.5, blend_mode="hann", verbose=False)
    out = sw.predict_with_sliding_window(golden["sw_img"], fake_model, 127.5, 50.0, use_tta=False)
    img = golden["sw_img"]
    positions = G.tile_positions(img.shape[0], img.shape[1], 64, G.stride_for(64, 0.5))
    preds = [fake_model.predict_single(img[y:y + 64, x:x + 64], 127.5, 50.0) for (y, x) in positions]
    np.testing.assert_array_equal(out, G.hann_reconstruct(preds, positions, img.shape[:2]))


def test_blend_edge_cases(eng):
    # uncovered pixels -> 0 (weight_sum clamp), single tile, empty list
    win = G.gaussian_window(64)
    t = np.full((64, 64), 0.75, np.float32)
    out = api.GaussianBlender(64).reconstruct([t], [(10, 20)], (100, 120))
    ref = G.gaussian_reconstruct([t], [(10, 20)], (100, 120), win)
    np.testing.assert_array_equal(out, ref)
    assert out[0, 0] == 0.0
    np.testing.assert_array_equal(api.LinearBlender().reconstruct([t], [(10, 20)], (100, 120)),
                                  G.linear_reconstruct([t], [(10, 20)], (100, 120)))
    assert api.LinearBlender().reconstruct([], [], (8, 8)).sum() == 0


def test_threshold_and_metrics_vs_reference(golden):
    pred, gt = golden["met_pred"], golden["met_gt"]
    np.testing.assert_array_equal(api.binarize_prediction(pred, 0.5), golden["met_bin"])
    keys = [str(k) for k in golden["met_keys"]]
    cases = {"rand": (pred, gt, 0.5), "thr7": (pred, gt, 0.7),
             "empty": (np.zeros((16, 16), np.float32), np.zeros((16, 16), np.uint8), 0.5),
             "nopred": (np.zeros((16, 16), np.float32), np.ones((16, 16), np.uint8), 0.5)}
    for tag, (p, g, thr) in cases.items():
        m = api.calculate_pixel_metrics(p, g, thr)
        np.testing.assert_array_equal(np.array([float(m[k]) for k in keys]), golden[f"met_{tag}"])


def test_metrics_large_counts_exact(eng):
    rng = np.random.default_rng(5)
    p = rng.random((4096, 4096), dtype=np.float32)
    g = (rng.random((4096, 4096)) > 0.3).astype(np.uint8)
    mask, counts = eng.threshold_metrics(p, g, 0.5)
    ref = G.pixel_metrics(p, g, 0.5)
    assert counts == (ref["tp"], ref["fp"], ref["fn"], ref["tn"])
    np.testing.assert_array_equal(mask, G.binarize(p, 0.5))
    assert sum(counts) == p.size


def test_loss_and_grad_vs_oracle(eng):
    rng = np.random.default_rng(9)
    y = (rng.random((2, 128, 128)) > 0.6).astype(np.float32)
    p = rng.random((2, 128, 128)).astype(np.float32)
    p[0, 0, :4] = [0.0, 1.0, 1e-8, 1 - 1e-8]          # exercise the clip
    res, g = eng.loss_metrics(p, y, want_grad=True)
    # float32 oracle: the clip bounds 1e-7 / 1-1e-7 are float32 quantities in TF as well
    pt = torch.from_numpy(p).requires_grad_(True)
    yt = torch.from_numpy(y)
    loss = U.combined_loss_standard(yt, pt)
    loss.backward()
    assert abs(res["loss"] - float(loss.detach())) <= 2e-6 * max(1.0, abs(float(loss.detach())))
    assert abs(res["dice_coef"] - float(U.dice_coef(yt, pt.detach()))) <= 2e-6
    gref = pt.grad.numpy()
    assert np.abs(g - gref).max() <= 2e-6 * max(1.0, np.abs(gref).max())
    # away from the clip the float64 oracle agrees to 1e-6 too
    p2 = np.clip(p, 0.01, 0.99)
    res2 = eng.loss_metrics(p2, y)
    l64 = float(U.combined_loss_standard(yt.double(), torch.from_numpy(p2).double()))
    assert abs(res2["loss"] - l64) <= 1e-6 * max(1.0, abs(l64))


# ---- loss recipes of compile_model (train_adipose_unet_v3.py:808-855): hard-example mining and label smoothing
@pytest.mark.parametrize("keep,eps", [(1.0, (0.03, 0.07)), (0.7, (0.0, 0.0)), (0.7, (0.03, 0.07)), (0.5, (0.0, 0.0))])
def test_loss_recipes_vs_oracle(keep, eps):
    import torch
    from oracle import unet as U
    rng = np.random.default_rng(7)
    B, S, SW = 3, 96, 64           # non-square: rows (H = 96) are what hard mining ranks, the row length (W = 64) what it averages
    y = (rng.random((B, S, SW)) < 0.3).astype(np.float32)
    # probabilities spanning saturated (clipped), confident and uncertain pixels
    z = rng.standard_normal((B, S, SW)).astype(np.float32) * 4.0
    p = (1.0 / (1.0 + np.exp(-z))).astype(np.float32)
    p[0, :4] = 0.0; p[1, :4] = 1.0                                   # exercise the clip and tie handling
    # float32 like TensorFlow: for saturated pixels clip(p) + 1e-7 and 1 - clip(p) + 1e-7 are float32 roundings
    # (1 - 1e-7 -> 0.99999988), which a float64 evaluation would not reproduce
    pt = torch.tensor(p, dtype=torch.float32, requires_grad=True)
    yt = torch.tensor(y, dtype=torch.float32)
    if keep < 1.0:
        loss = U.online_hard_example_mining_loss(yt, pt, keep, eps[0], eps[1])
    else:
        loss = U.combined_loss_with_label_smoothing(yt, pt, eps[0], eps[1])
    loss.backward()
    eng = api.default_engine()
    res, g = eng.loss_metrics(p, y, want_grad=True, ohem_keep_ratio=keep, eps_pos=eps[0], eps_neg=eps[1])
    assert abs(res["loss"] - float(loss)) <= 2e-5 * max(1.0, abs(float(loss)))          # float32 oracle sums vs float64 device sums
    assert abs(res["dice_coef"] - float(U.dice_coef(yt, pt.detach()))) <= 1e-5            # metric keeps the raw target
    gref = pt.grad.numpy()
    # same selected rows as the oracle (row means differ by many ulps here), gradient to float32 rounding
    l2 = np.linalg.norm((g - gref).ravel()) / np.linalg.norm(gref.ravel())
    differ = (np.abs(g - gref) > 1e-3 * np.abs(gref) + 1e-7 * np.abs(gref).max()).mean()      # a mis-selected row is off by ~100 %
    assert l2 <= 1e-5 and differ == 0.0, (l2, differ)
    if keep < 1.0:
        # the second (NumPy float64, hand-differentiated) restatement: loss, gradient and the number of selected rows
        from oracle import unet_numpy as N
        ln, gn, sel = N.ohem_loss_numpy(y, p, keep, eps[0], eps[1], dtype=np.float32)
        assert sel.sum() == B * int(np.float32(S) * np.float32(keep))
        assert abs(res["loss"] - ln) <= 2e-5 * max(1.0, abs(ln))
        assert np.linalg.norm((g - gn).ravel()) / np.linalg.norm(gn.ravel()) <= 1e-4


def test_threshold_sweep_equals_per_threshold_counts():
    """One-pass sweep (adp_threshold_sweep) == calculate_pixel_metrics per candidate threshold (the reference's search loop,
    full_evaluation_enhanced.py:891-980), including probabilities that sit exactly on a candidate (strict >)."""
    rng = np.random.default_rng(3)
    eng = api.default_engine()
    n = 300_007
    p = rng.random(n).astype(np.float32)
    thr = np.arange(0.1, 0.95, 0.05)                                   # the reference's default candidate range
    t32 = thr.astype(np.float32)
    p[:len(t32)] = t32                                                 # exactly on the candidates
    p[len(t32):2 * len(t32)] = np.nextafter(t32, np.float32(1))        # one ulp above
    gt = (rng.random(n) < 0.3).astype(np.uint8)
    sweep = eng.threshold_sweep(p, gt, t32)
    assert sweep.shape == (len(thr), 4)
    for j, t in enumerate(t32):
        _, counts = eng.threshold_metrics(p, gt, float(t), want_mask=False)
        assert tuple(sweep[j]) == counts, (j, t)
        ref = G.pixel_metrics(p, gt, float(t))
        assert (ref["tp"], ref["fp"], ref["fn"], ref["tn"]) == counts
    assert (sweep.sum(axis=1) == n).all()


@pytest.mark.parametrize("k", [3, 5, 9])
def test_boundary_refine_vs_oracle_and_cv2(eng, k):
    """adp_boundary_refine (BoundaryRefiner.refine, full_evaluation_enhanced.py:332-393): bit-exact against the restatement of
    OpenCV's algorithms (oracle/refine.py); against the cv2 of the image within one grey level (its 8-bit bilateral path is
    not bit-stable across builds) and with mask agreement."""
    from oracle import refine as R
    from refine_helpers import _blob_prob, reference_refine_cv2
    probs = np.stack([_blob_prob(200, 264, 1), (_blob_prob(200, 264, 2) > 0.55).astype(np.float32), np.zeros((200, 264), np.float32)])
    out = eng.boundary_refine(probs, kernel_size=k)
    for i in range(len(probs)):
        np.testing.assert_array_equal(out[i], R.refine(probs[i], kernel_size=k))
        ref = reference_refine_cv2(probs[i], k=k)
        assert np.abs(out[i] - ref).max() <= 1.0 / 255.0 + 1e-7
        assert ((out[i] > 0.5) != (ref > 0.5)).mean() <= 1e-4
    # device-resident input and output, reference-style object
    td = torch.from_numpy(probs).cuda()
    od = torch.empty_like(td)
    eng.boundary_refine(td, kernel_size=k, out=od)
    np.testing.assert_array_equal(od.cpu().numpy(), out)
    np.testing.assert_array_equal(api.BoundaryRefiner(kernel_size=k, engine=eng).refine(probs[0], image=None), out[0])
