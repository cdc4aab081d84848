"""Hand-derived known-answer case for the U-Net graph, shared by the CPU oracle test and the GPU test.

The reference ships no golden tensor and TensorFlow is not installable here, so nothing TF-produced can pin the graph.  What CAN
be pinned without trusting either restatement is the WIRING: a sparse weight set turns the 22-layer network into a composition
of shifts, 2x2 max-pools, nearest upsamplings and one sum whose result is written down below in plain NumPy from the Keras op
semantics alone (Segmentation/train_adipose_unet_v3.py:664-752):

  * Conv2D(padding='same', dilation_rate=d) is a cross-correlation: a single tap (ky, kx) reads in[y + d(ky-1), x + d(kx-1)],
    zeros outside the image (exercised at d = 1, 2, 4, 16),
  * MaxPooling2D((2,2)) / UpSampling2D((2,2)) nearest (out[y,x] = in[y//2, x//2]),
  * Add([dilate1 .. dilate6]) is the plain sum of the six CHAINED conv outputs (:681-688),
  * Concatenate([down_s, up_s]) puts the skip tensor FIRST (:693, :700, :707): the up path enters conv2 at input channel
    C_skip, skip channel 1 at input channel 1,
  * the output is channel 1 of the 2-class softmax = sigmoid(z1 - z0) (:748-750),
  * the input is normalised as float32 (image - mean) / (std + 1e-10) before the graph (full_evaluation_enhanced.py:1306).

Every conv output here is non-negative after the first ReLU, so later ReLUs are identities and the expectation is exact up to
floating-point rounding.
"""
import numpy as np

from adipose_unet_b200.layers import conv_layers

HEAD = dict(z1_a=0.5, z0_a=-0.25, z1_b=0.25, z0_b=0.125)        # exactly representable in bf16


def probe_weights():
    w = {}
    for name, cin, cout, k, _ in conv_layers():
        w[name + "/kernel"] = np.zeros((k, k, cin, cout), np.float32)
        w[name + "/bias"] = np.zeros((cout,), np.float32)
    c1, c2, c4 = 44, 88, 176

    def tap(name, ky, kx, cin, cout, v=1.0):
        w[name + "/kernel"][ky, kx, cin, cout] = v

    tap("down1_conv1", 1, 1, 0, 0)            # channel 0: relu(x)
    tap("down1_conv1", 0, 2, 0, 1)            # channel 1: relu(x)[y-1, x+1]
    tap("down1_conv2", 1, 1, 0, 0)
    tap("down1_conv2", 1, 1, 1, 1)
    for name in ("down2_conv1", "down2_conv2", "down3_conv1", "down3_conv2", "dilate1", "dilate4", "dilate6",
                 "up3_conv1", "up3_conv3", "up2_conv1", "up2_conv3", "up1_conv1"):
        tap(name, 1, 1, 0, 0)
    tap("dilate2", 0, 1, 0, 0)                # d = 2:  in[y-2, x]
    tap("dilate3", 1, 2, 0, 0)                # d = 4:  in[y, x+4]
    tap("dilate5", 2, 1, 0, 0)                # d = 16: in[y+16, x]
    tap("up3_conv2", 1, 1, c4, 0)             # first channel of the UP half of Concatenate([down3, up3])
    tap("up2_conv2", 1, 1, c2, 0)
    tap("up1_conv2", 1, 1, 1, 0)              # skip channel 1 -> output channel 0   (path A)
    tap("up1_conv2", 1, 1, c1, 1)             # up channel 0   -> output channel 1   (path B)
    tap("up1_conv3", 1, 1, 0, 0)
    tap("up1_conv3", 1, 1, 1, 1)
    k = w["output_softmax/kernel"]
    k[0, 0, 0, 1], k[0, 0, 0, 0] = HEAD["z1_a"], HEAD["z0_a"]
    k[0, 0, 1, 1], k[0, 0, 1, 0] = HEAD["z1_b"], HEAD["z0_b"]
    return w


def _shift(a, dy, dx):
    """out[y, x] = a[y + dy, x + dx], zeros outside."""
    h, w_ = a.shape
    out = np.zeros_like(a)
    ys, ye = max(0, -dy), min(h, h - dy)
    xs, xe = max(0, -dx), min(w_, w_ - dx)
    if ye > ys and xe > xs:
        out[ys:ye, xs:xe] = a[ys + dy:ye + dy, xs + dx:xe + dx]
    return out


def _pool2(a):
    h, w_ = a.shape
    return a.reshape(h // 2, 2, w_ // 2, 2).max(axis=(1, 3))


def expected_probability(image, mean, std):
    """float64 evaluation of the composition the probe weights select; image: (S, S) float32 in 0..255, S a multiple of 8."""
    xn = ((image - mean) / (std + 1e-10)).astype(np.float32).astype(np.float64)
    r = np.maximum(xn, 0.0)
    a = _shift(r, -1, +1)
    t1 = _pool2(_pool2(_pool2(r)))
    t2 = _shift(t1, -2, 0)
    t3 = _shift(t2, 0, +4)
    t4 = t3
    t5 = _shift(t4, +16, 0)
    t6 = t5
    s = t1 + t2 + t3 + t4 + t5 + t6
    b = s.repeat(8, axis=0).repeat(8, axis=1)
    logit = (HEAD["z1_a"] - HEAD["z0_a"]) * a + (HEAD["z1_b"] - HEAD["z0_b"]) * b
    return 1.0 / (1.0 + np.exp(-logit))
