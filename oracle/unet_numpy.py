"""Second, independent CPU restatement of the reference graph in plain NumPy float64 (TEST INFRASTRUCTURE ONLY).

oracle/unet.py leans on torch.nn.functional for conv / pool / interpolate; this file re-derives every op from the Keras
semantics SURVEY.md Appendix C lists (no torch): if the two restatements agree, a convention slip (kernel flip, HWIO index
order, 'same' padding under dilation, concat order, softmax channel, half-pixel bilinear) in either one would show.
PARITY UNPINNED against TensorFlow itself (not installable here; the reference ships no golden tensors).

Graph: Segmentation/train_adipose_unet_v3.py:660-752 (deep-supervision heads :712-745)."""
from __future__ import annotations

import numpy as np


def conv2d_same(x: np.ndarray, k: np.ndarray, b: np.ndarray, dilation: int = 1, relu: bool = True) -> np.ndarray:
    """Conv2D(padding='same', strides=1, dilation_rate=d), channels_last: x (B,H,W,Cin), k (kh,kw,Cin,Cout) HWIO.
    out[b,y,x,o] = b[o] + sum_{i,j,c} x[b, y+(i-(kh-1)/2)d, x+(j-(kw-1)/2)d, c] * k[i,j,c,o]   (cross-correlation, zero pad)."""
    B, H, W, _ = x.shape
    kh, kw, _, co = k.shape
    ph, pw = dilation * (kh - 1) // 2, dilation * (kw - 1) // 2
    xp = np.zeros((B, H + 2 * ph, W + 2 * pw, x.shape[3]), np.float64)
    xp[:, ph:ph + H, pw:pw + W] = x
    out = np.zeros((B, H, W, co), np.float64) + b.astype(np.float64)
    for i in range(kh):
        for j in range(kw):
            out += xp[:, i * dilation:i * dilation + H, j * dilation:j * dilation + W] @ k[i, j].astype(np.float64)
    return np.maximum(out, 0.0) if relu else out


def maxpool2(x):          # MaxPooling2D((2,2), strides=(2,2)), even sizes
    B, H, W, C = x.shape
    return x.reshape(B, H // 2, 2, W // 2, 2, C).max(axis=(2, 4))


def upsample2(x):         # UpSampling2D((2,2)) nearest: out[y,x] = in[y//2, x//2]
    return np.repeat(np.repeat(x, 2, axis=1), 2, axis=2)


def resize_bilinear_half_pixel(a: np.ndarray, size: int) -> np.ndarray:
    """tf.image.resize(x, [size,size], 'bilinear') for (B,h,w): half-pixel centres, no antialias (TF2 default)."""
    B, h, w = a.shape

    def axis(n_in, n_out):
        src = (np.arange(n_out) + 0.5) * (n_in / n_out) - 0.5
        fl = np.floor(src)
        lo = np.maximum(fl, 0).astype(int); hi = np.minimum(np.ceil(src), n_in - 1).astype(int)
        return lo, hi, src - fl

    y0, y1, ty = axis(h, size); x0, x1, tx = axis(w, size)
    top = a[:, y0][:, :, x0] + (a[:, y0][:, :, x1] - a[:, y0][:, :, x0]) * tx
    bot = a[:, y1][:, :, x0] + (a[:, y1][:, :, x1] - a[:, y1][:, :, x0]) * tx
    return top + (bot - top) * ty[None, :, None]


def forward(x: np.ndarray, w: dict, deep_supervision: bool = False):
    """x: (B,H,W) normalised.  Returns probabilities (B,H,W) [and the two auxiliary outputs]."""
    def C(t, name, d=1, relu=True):
        return conv2d_same(t, w[name + "/kernel"], w[name + "/bias"], d, relu)

    t = x[..., None].astype(np.float64)                                   # Reshape((H,W,1))
    d1 = C(C(t, "down1_conv1"), "down1_conv2")
    d2 = C(C(maxpool2(d1), "down2_conv1"), "down2_conv2")
    d3 = C(C(maxpool2(d2), "down3_conv1"), "down3_conv2")
    ts = []
    cur = maxpool2(d3)
    for i, d in enumerate((1, 2, 4, 8, 16, 32)):
        cur = C(cur, f"dilate{i + 1}", d)
        ts.append(cur)
    s = sum(ts)                                                           # Add([dilate1..6])
    u3 = C(upsample2(s), "up3_conv1")
    u3 = C(C(np.concatenate([d3, u3], axis=-1), "up3_conv2"), "up3_conv3")   # Concatenate([down3, up3])
    u2 = C(upsample2(u3), "up2_conv1")
    u2 = C(C(np.concatenate([d2, u2], axis=-1), "up2_conv2"), "up2_conv3")
    u1 = C(upsample2(u2), "up1_conv1")
    u1 = C(C(np.concatenate([d1, u1], axis=-1), "up1_conv2"), "up1_conv3")
    z = C(u1, "output_softmax", relu=False)                               # (B,H,W,2)
    e = np.exp(z - z.max(axis=-1, keepdims=True))
    prob = (e / e.sum(axis=-1, keepdims=True))[..., 1]                    # softmax, channel 1, squeezed
    if not deep_supervision:
        return prob
    a1 = 1.0 / (1.0 + np.exp(-C(u3, "aux_out1", relu=False)[..., 0]))
    a2 = 1.0 / (1.0 + np.exp(-C(u2, "aux_out2", relu=False)[..., 0]))
    H = x.shape[1]
    return prob, resize_bilinear_half_pixel(a1, H), resize_bilinear_half_pixel(a2, H)


def ohem_loss_numpy(y: np.ndarray, p: np.ndarray, keep_ratio: float = 0.7, epsilon_pos: float = 0.0, epsilon_neg: float = 0.0,
                    dtype=np.float64):
    """Second restatement (float64, no torch) of online_hard_example_mining_loss[_with_smoothing]
    (Segmentation/train_adipose_unet_v3.py:282-363) for y, p of shape (B,H,W), with its gradient dL/dp written out by hand.

    Line by line: binary_crossentropy (:301) = mean over the LAST axis of -[y log(pc+eps) + (1-y) log(1-pc+eps)] with
    pc = clip(p, eps, 1-eps) -> (B,H); reshape to (B,-1) (:305) -> (B,H); num_pixels = H (:308);
    k = int(float32(H)*keep_ratio) (:309); top_k per image (:312), ties to the lower index; reduce_mean over (B,k) (:313);
    plus dice_loss over every pixel (:316, :217-225).  Returns (loss, dL/dp, selected-row mask (B,H)).
    dtype = np.float32 evaluates the element-wise statements in float32 like TensorFlow (for saturated pixels
    clip(p) + 1e-7 and 1 - clip(p) + 1e-7 are float32 roundings that float64 does not reproduce); reductions stay float64."""
    eps = dtype(1e-7)
    y = y.astype(dtype); p = p.astype(dtype)
    one = dtype(1.0)
    ys = y * dtype(1.0 - epsilon_pos - epsilon_neg) + dtype(epsilon_neg) if (epsilon_pos or epsilon_neg) else y
    B, H, W = p.shape
    pc = np.clip(p, eps, one - eps)
    inside = (p >= eps) & (p <= one - eps)                       # gradient of clip
    bce = -(ys * np.log(pc + eps) + (one - ys) * np.log(one - pc + eps))
    rows = bce.astype(np.float64).mean(axis=-1)                  # (B,H)
    ys = ys.astype(np.float64); pc = pc.astype(np.float64); eps = 1e-7
    k = int(np.float32(H) * np.float32(keep_ratio))
    sel = np.zeros((B, H), bool)
    for b in range(B):
        order = np.argsort(-rows[b], kind="stable")              # descending, ties keep the lower index first
        sel[b, order[:k]] = True
    hard = rows[sel].sum() / (B * k)
    inter = (ys * pc).sum(); den = ys.sum() + pc.sum() + 1.0
    loss = hard + 1.0 - (2.0 * inter + 1.0) / den
    dbce = -(ys / (pc + eps) - (1.0 - ys) / (1.0 - pc + eps))
    ddice = -(2.0 * ys * den - (2.0 * inter + 1.0)) / den ** 2
    g = (sel[:, :, None] * dbce / (W * B * k) + ddice) * inside
    return float(loss), g, sel
