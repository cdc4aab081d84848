"""Hand-derived known answers for the graph semantics (tests/known_answer.py): both CPU restatements of the reference graph
must reproduce a composition of shifts / pools / upsamplings / one sum that is written down independently of them."""
import numpy as np
import torch

import adipose_unet_b200 as A
from oracle import unet as U
from oracle import unet_numpy as UN

import known_answer as KA

MEAN, STD = A.synth.DEFAULT_MEAN, A.synth.DEFAULT_STD


def test_probe_network_matches_hand_derivation_torch_restatement():
    img = A.synth.ecm_tile(256, seed=21).astype(np.float32)
    want = KA.expected_probability(img, MEAN, STD)
    got = U.predict_single(img, MEAN, STD, U.to_torch_params(KA.probe_weights()))
    assert want.min() >= 0.5 and want.max() > 0.9 and np.unique(np.round(want, 3)).size > 100      # a non-trivial field
    assert np.abs(got - want).max() < 2e-6
    # every path matters: dropping the skip path or the deep path changes the answer by far more than the tolerance
    for name, idx in (("up1_conv2", (1, 1, 1, 0)), ("up1_conv2", (1, 1, 44, 1)), ("dilate5", (2, 1, 0, 0))):
        w = KA.probe_weights()
        w[name + "/kernel"][idx] = 0.0
        other = U.predict_single(img, MEAN, STD, U.to_torch_params(w))
        assert np.abs(other - want).max() > 1e-2, name


def test_probe_network_matches_hand_derivation_numpy_restatement():
    img = A.synth.ecm_tile(64, seed=22).astype(np.float32)          # 8 x 8 bottleneck: the d = 16 tap reads zero padding only
    want = KA.expected_probability(img, MEAN, STD)
    xn = ((img - MEAN) / (STD + 1e-10)).astype(np.float32)
    got = UN.forward(xn[None], KA.probe_weights())[0]
    assert np.abs(got - want).max() < 1e-12


def test_single_tap_conv_is_a_shift_in_the_cross_correlation_direction():
    """Conv2D 'same' with dilation d and the single tap (ky, kx): out[y, x] = in[y + d(ky-1), x + d(kx-1)], zero padded."""
    rng = np.random.default_rng(5)
    x = rng.random((12, 12))
    for d in (1, 2, 4):
        for ky in range(3):
            for kx in range(3):
                k = np.zeros((3, 3, 1, 1)); k[ky, kx, 0, 0] = 1.0
                want = KA._shift(x, d * (ky - 1), d * (kx - 1))
                got_np = UN.conv2d_same(x[None, :, :, None], k, np.zeros(1), d, relu=False)[0, :, :, 0]
                got_t = torch.nn.functional.conv2d(torch.from_numpy(x)[None, None], torch.from_numpy(k).permute(3, 2, 0, 1),
                                                   padding=d, dilation=d)[0, 0].numpy()
                assert np.array_equal(got_np, want) and np.array_equal(got_t, want), (d, ky, kx)


def test_bilinear_resize_half_pixel_known_values():
    """tf.image.resize(..., 'bilinear') of the deep-supervision heads (train_adipose_unet_v3.py:716-726) samples at half-pixel
    centres without antialiasing: doubling [0, 1] gives [0, 0.25, 0.75, 1] (edge clamp), doubling [0, 4, 8] gives
    [0, 1, 3, 5, 7, 8] - written down by hand, checked on both restatements."""
    a = np.array([[0.0, 1.0], [0.0, 1.0]])
    want = np.tile(np.array([0.0, 0.25, 0.75, 1.0]), (4, 1))
    got_np = UN.resize_bilinear_half_pixel(a[None], 4)[0]
    got_t = torch.nn.functional.interpolate(torch.from_numpy(a)[None, None], size=(4, 4), mode="bilinear", align_corners=False)[0, 0].numpy()
    assert np.allclose(got_np, want, atol=1e-12) and np.allclose(got_t, want, atol=1e-12)
    b = np.tile(np.array([0.0, 4.0, 8.0]), (3, 1))
    want_b = np.tile(np.array([0.0, 1.0, 3.0, 5.0, 7.0, 8.0]), (6, 1))
    assert np.allclose(UN.resize_bilinear_half_pixel(b[None], 6)[0], want_b, atol=1e-12)
    got_tb = torch.nn.functional.interpolate(torch.from_numpy(b)[None, None], size=(6, 6), mode="bilinear", align_corners=False)[0, 0].numpy()
    assert np.allclose(got_tb, want_b, atol=1e-12)
