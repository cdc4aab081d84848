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
