"""Sanity of the PyTorch-CPU graph restatement (oracle/unet.py): structure constants
SURVEY.md fixes, and self-consistency fp32 vs float64."""
import numpy as np
import torch

import adipose_unet_b200 as A
from oracle import unet as U


def test_param_count_and_flops():
    assert A.layers.param_count() == 8_507_050
    assert abs(A.layers.forward_flops(1024) / 1e9 - 896.3) < 0.1
    w = A.synth.init_weights()
    assert sum(v.size for v in w.values()) == 8_507_050
    assert w["up3_conv2/kernel"].shape == (3, 3, 352, 176)
    assert w["output_softmax/kernel"].shape == (1, 1, 44, 2)


def test_forward_shapes_and_fp64_agreement():
    torch.manual_seed(0)
    w = A.synth.init_weights()
    img = A.synth.ecm_tile(64, seed=3).astype(np.float32)
    p32 = U.predict_single(img, 127.5, 50.0, U.to_torch_params(w, torch.float32))
    p64 = U.predict_single(img, 127.5, 50.0, U.to_torch_params(w, torch.float64), dtype=torch.float64)
    assert p32.shape == (64, 64) and p32.dtype == np.float32
    assert np.abs(p32 - p64).max() < 1e-5
    assert 0.0 < p32.min() and p32.max() < 1.0


def test_softmax_channel1_is_sigmoid_of_difference():
    z = torch.randn(2, 2, 5, 5)
    a = torch.softmax(z, dim=1)[:, 1]
    b = torch.sigmoid(z[:, 1] - z[:, 0])
    assert torch.allclose(a, b, atol=1e-6)


def test_loss_matches_manual():
    rng = np.random.default_rng(1)
    y = (rng.random((2, 8, 8)) > 0.5).astype(np.float32)
    p = rng.random((2, 8, 8)).astype(np.float32)
    yt, pt = torch.from_numpy(y), torch.from_numpy(p)
    loss = float(U.combined_loss_standard(yt, pt))
    pc = np.clip(p.astype(np.float64), 1e-7, 1 - 1e-7)
    bce = -(y * np.log(pc + 1e-7) + (1 - y) * np.log(1 - pc + 1e-7)).mean()
    dice = 1 - (2 * (y * pc).sum() + 1) / (y.sum() + pc.sum() + 1)
    assert abs(loss - (bce + dice)) < 1e-5


def test_keras_adam_differs_from_torch_adam_only_by_eps_placement():
    rng = np.random.default_rng(2)
    th = rng.standard_normal(100).astype(np.float32)
    g = rng.standard_normal(100).astype(np.float32)
    t1, m1, v1 = U.keras_adam_step(th, g, np.zeros_like(th), np.zeros_like(th), 1, 1e-3)
    # at t=1: m=(1-b1)g, v=(1-b2)g^2, alpha = lr*sqrt(1-b2)/(1-b1) -> step = lr*g/(|g| + eps/sqrt(1-b2)) approx
    expect = th - 1e-3 * np.sqrt(1 - 0.999) / (1 - 0.9) * (0.1 * g) / (np.sqrt(0.001 * g * g) + 1e-7)
    assert np.abs(t1 - expect).max() < 1e-6


def test_two_independent_restatements_agree():
    """oracle/unet.py (torch.nn.functional) against oracle/unet_numpy.py (plain NumPy float64 written from the Keras op
    semantics): forward incl. all six dilations, both pools/upsamplings/concats, the softmax head and the
    deep-supervision heads with the half-pixel bilinear resize."""
    from oracle import unet_numpy as N
    w = A.synth.init_weights(deep_supervision=True)
    rng = np.random.default_rng(5)
    x = rng.standard_normal((2, 64, 64)).astype(np.float32)
    with torch.no_grad():
        pt, a1t, a2t = U.forward(torch.from_numpy(x).to(torch.float64), U.to_torch_params(w, torch.float64), deep_supervision=True)
    pn, a1n, a2n = N.forward(x, w, deep_supervision=True)
    assert np.abs(pt.numpy() - pn).max() < 1e-10
    assert np.abs(a1t.numpy() - a1n).max() < 1e-10
    assert np.abs(a2t.numpy() - a2n).max() < 1e-10
    # the dilation-32 layer really sees its neighbours on a larger map (8x8 at 64^2 only ever reads the centre tap)
    k = rng.standard_normal((3, 3, 3, 2)); b = rng.standard_normal(2); t = rng.standard_normal((1, 70, 70, 3))
    ref = torch.nn.functional.conv2d(torch.from_numpy(t).permute(0, 3, 1, 2), torch.from_numpy(k).permute(3, 2, 0, 1),
                                     torch.from_numpy(b), padding=32, dilation=32).permute(0, 2, 3, 1).numpy()
    assert np.abs(N.conv2d_same(t, k, b, 32, relu=False) - ref).max() < 1e-10


def test_hard_mining_ranks_row_means_as_the_reference_writes_it():
    """online_hard_example_mining_loss (train_adipose_unet_v3.py:282-323): binary_crossentropy averages the LAST axis, so
    the top-k runs over the (B,H) per-row means with k = int(float32(H)*ratio) - 716 of 1024 rows at the CLI default.
    The torch restatement, the hand-differentiated NumPy restatement and a literal line-by-line evaluation agree."""
    from oracle import unet_numpy as N
    assert int(np.float32(1024) * np.float32(0.7)) == 716
    rng = np.random.default_rng(11)
    B, H, W = 3, 40, 24
    y = (rng.random((B, H, W)) < 0.3).astype(np.float32)
    p = (1.0 / (1.0 + np.exp(-4.0 * rng.standard_normal((B, H, W))))).astype(np.float32)
    p[0, :2] = 0.0; p[1, :2] = 1.0                                        # clipped rows
    for keep, ep, en in ((0.7, 0.0, 0.0), (0.5, 0.03, 0.07)):
        pt = torch.tensor(p, dtype=torch.float64, requires_grad=True)
        yt = torch.tensor(y, dtype=torch.float64)
        loss = U.online_hard_example_mining_loss(yt, pt, keep, ep, en)
        loss.backward()
        ln, gn, sel = N.ohem_loss_numpy(y, p, keep, ep, en)
        k = int(np.float32(H) * np.float32(keep))
        assert sel.sum(axis=1).tolist() == [k] * B
        assert abs(float(loss) - ln) < 1e-12
        assert np.abs(pt.grad.numpy() - gn).max() < 1e-12
        # literal evaluation of the reference's statements with NumPy stand-ins for the TF ops
        ys = y.astype(np.float64) * (1.0 - ep - en) + en if (ep or en) else y.astype(np.float64)
        pc = np.clip(p.astype(np.float64), 1e-7, 1 - 1e-7)
        per_pixel_bce = np.mean(-(ys * np.log(pc + 1e-7) + (1 - ys) * np.log(1 - pc + 1e-7)), axis=-1)   # :301 -> (B,H)
        flat_loss = per_pixel_bce.reshape(B, -1)                                                          # :305
        num_pixels = flat_loss.shape[1]                                                                   # :308
        assert num_pixels == H
        kk = int(np.float32(num_pixels) * np.float32(keep))                                               # :309
        top_k_loss = -np.sort(-flat_loss, axis=1)[:, :kk]                                                 # :312
        hard_bce = top_k_loss.mean()                                                                      # :313
        dice = 1.0 - (2.0 * (ys * pc).sum() + 1.0) / (ys.sum() + pc.sum() + 1.0)
        assert abs(hard_bce + dice - ln) < 1e-12
