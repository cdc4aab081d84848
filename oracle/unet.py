"""ORACLE (test infrastructure, never shipped, never on the product path).

PyTorch-CPU restatement of the reference U-Net graph, its inference seam, the
BCE+Dice loss and the Keras-2.13 Adam step.  PARITY UNPINNED for these rows
(SURVEY.md section 8a M1, M2, T1-T3): the arithmetic of the reference lives in
tensorflow==2.13.1 / keras==2.13.1 (requirements.txt:5-6), which are not
installable in this image (no wheel, Python 3.12, no network), and the
reference ships no checkpoint, golden tensor or test.  The restatement follows
the Keras op semantics written out in SURVEY.md Appendix C; float64 is the
arbiter when fp32 implementations disagree.

Paths cited are relative to /root/reference.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from . import geometry as G

# (name, cin, cout, k, dilation) — Segmentation/train_adipose_unet_v3.py:668-750
def _layers(init_nb=44):
    c1, c2, c4, c8 = init_nb, 2 * init_nb, 4 * init_nb, 8 * init_nb
    return [("down1_conv1", 1, c1, 3, 1), ("down1_conv2", c1, c1, 3, 1),
            ("down2_conv1", c1, c2, 3, 1), ("down2_conv2", c2, c2, 3, 1),
            ("down3_conv1", c2, c4, 3, 1), ("down3_conv2", c4, c4, 3, 1),
            ("dilate1", c4, c8, 3, 1), ("dilate2", c8, c8, 3, 2), ("dilate3", c8, c8, 3, 4),
            ("dilate4", c8, c8, 3, 8), ("dilate5", c8, c8, 3, 16), ("dilate6", c8, c8, 3, 32),
            ("up3_conv1", c8, c4, 3, 1), ("up3_conv2", c8, c4, 3, 1), ("up3_conv3", c4, c4, 3, 1),
            ("up2_conv1", c4, c2, 3, 1), ("up2_conv2", c4, c2, 3, 1), ("up2_conv3", c2, c2, 3, 1),
            ("up1_conv1", c2, c1, 3, 1), ("up1_conv2", c2, c1, 3, 1), ("up1_conv3", c1, c1, 3, 1),
            ("output_softmax", c1, 2, 1, 1)]


DILATION = {n: d for n, _, _, _, d in _layers()}
DILATION.update({"aux_out1": 1, "aux_out2": 1})     # deep-supervision heads, train_adipose_unet_v3.py:715, 722


def to_torch_params(weights: Dict[str, np.ndarray], dtype=torch.float32, requires_grad=False):
    """HWIO kernels -> OIHW tensors; name -> (w, b)."""
    p = {}
    names = [n for n, *_ in _layers()] + [n for n in ("aux_out1", "aux_out2") if n + "/kernel" in weights]
    for name in names:
        k = torch.from_numpy(np.ascontiguousarray(weights[name + "/kernel"])).to(dtype)
        b = torch.from_numpy(np.ascontiguousarray(weights[name + "/bias"])).to(dtype)
        w = k.permute(3, 2, 0, 1).contiguous()
        if requires_grad:
            w.requires_grad_(True); b.requires_grad_(True)
        p[name] = (w, b)
    return p


def _conv(x, p, name, act=True):
    """Conv2D(padding='same', dilation_rate=d) + bias (+ReLU): zero-pad d per side,
    cross-correlation.  train_adipose_unet_v3.py:668-709."""
    w, b = p[name]
    d = DILATION[name]
    pad = d * (w.shape[-1] - 1) // 2
    y = F.conv2d(x, w, b, padding=pad, dilation=d)
    return F.relu(y) if act else y


def forward(x: torch.Tensor, p, dropout_masks: Optional[Dict[str, torch.Tensor]] = None,
            taps: Optional[Dict[str, torch.Tensor]] = None, deep_supervision: bool = False):
    """Graph of train_adipose_unet_v3.py:664-752.  x: (B,H,W) already normalised.
    Returns (B,H,W) probabilities = softmax(z)[...,1] = sigmoid(z1-z0).
    dropout_masks (training only): name -> 0/1 mask, applied with the inverted
    scaling 1/(1-0.3) at the four Dropout sites; None = inference (identity).
    taps: optional dict filled with every layer's post-activation output (NCHW)."""
    def rec(name, t):
        if taps is not None:
            taps[name] = t
        return t

    def drop(name, t):
        if dropout_masks is None or name not in dropout_masks:
            return t
        return t * dropout_masks[name] * (1.0 / 0.7)

    x = x.unsqueeze(1)  # Reshape((1024,1024,1)) — channel dim
    d1 = rec("down1_conv1", _conv(x, p, "down1_conv1"))
    d1 = rec("down1_conv2", _conv(d1, p, "down1_conv2"))
    d1p = F.max_pool2d(d1, 2, 2)
    d2 = rec("down2_conv1", _conv(d1p, p, "down2_conv1"))
    d2 = rec("down2_conv2", _conv(d2, p, "down2_conv2"))
    d2p = F.max_pool2d(d2, 2, 2)
    d3 = rec("down3_conv1", _conv(d2p, p, "down3_conv1"))
    d3 = rec("down3_conv2", _conv(d3, p, "down3_conv2"))
    d3p = F.max_pool2d(d3, 2, 2)

    t1 = rec("dilate1", _conv(d3p, p, "dilate1"))
    t1 = drop("dropout_dilate1", t1)          # the Add sees the post-dropout tensor (:681-688)
    t2 = rec("dilate2", _conv(t1, p, "dilate2"))
    t3 = rec("dilate3", _conv(t2, p, "dilate3"))
    t4 = rec("dilate4", _conv(t3, p, "dilate4"))
    t5 = rec("dilate5", _conv(t4, p, "dilate5"))
    t6 = rec("dilate6", _conv(t5, p, "dilate6"))
    s = rec("dilate_add", t1 + t2 + t3 + t4 + t5 + t6)

    def up(t):  # UpSampling2D((2,2)) nearest: out[y,x] = in[y//2,x//2]
        return t.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)

    u3 = rec("up3_conv1", _conv(up(s), p, "up3_conv1"))
    u3 = torch.cat([d3, u3], dim=1)           # Concatenate([down3, up3]) — skip first (:693)
    u3 = rec("up3_conv2", _conv(u3, p, "up3_conv2"))
    u3 = rec("up3_conv3", _conv(u3, p, "up3_conv3"))
    u3 = drop("dropout_up3", u3)
    u2 = rec("up2_conv1", _conv(up(u3), p, "up2_conv1"))
    u2 = torch.cat([d2, u2], dim=1)
    u2 = rec("up2_conv2", _conv(u2, p, "up2_conv2"))
    u2 = rec("up2_conv3", _conv(u2, p, "up2_conv3"))
    u2 = drop("dropout_up2", u2)
    u1 = rec("up1_conv1", _conv(up(u2), p, "up1_conv1"))
    u1 = torch.cat([d1, u1], dim=1)
    u1 = rec("up1_conv2", _conv(u1, p, "up1_conv2"))
    u1 = rec("up1_conv3", _conv(u1, p, "up1_conv3"))
    u1 = drop("dropout_up1", u1)

    z = _conv(u1, p, "output_softmax", act=False)   # (B,2,H,W)
    prob = torch.softmax(z, dim=1)[:, 1]            # channel 1, squeezed (:748-750)
    if not deep_supervision:
        return rec("prob", prob)
    # deep supervision (:712-745): sigmoid 1x1 heads on the post-dropout up3 / up2 tensors, tf.image.resize(...,
    # 'bilinear') to the input size = half-pixel centres without antialiasing = F.interpolate(align_corners=False)
    size = x.shape[-2:]
    a1 = torch.sigmoid(_conv(u3, p, "aux_out1", act=False))
    a2 = torch.sigmoid(_conv(u2, p, "aux_out2", act=False))
    a1 = F.interpolate(a1, size=size, mode="bilinear", align_corners=False)[:, 0]
    a2 = F.interpolate(a2, size=size, mode="bilinear", align_corners=False)[:, 0]
    return rec("prob", prob), rec("aux_out1", a1), rec("aux_out2", a2)


def predict_single(image: np.ndarray, mean: float, std: float, p, dtype=torch.float32) -> np.ndarray:
    """AdiposeUNet.predict_single — full_evaluation_enhanced.py:1303-1321:
    NumPy float32 normalisation with Python-float scalars, batch of one."""
    norm = (image - mean) / (std + 1e-10)
    with torch.no_grad():
        out = forward(torch.from_numpy(np.ascontiguousarray(norm)).to(dtype).unsqueeze(0), p)
    return out[0].to(torch.float32).numpy()


def predict_with_tta(image: np.ndarray, mean: float, std: float, p, mode: str = "full",
                     dtype=torch.float32) -> np.ndarray:
    """TestTimeAugmentation.predict_with_tta — full_evaluation_enhanced.py:577-600."""
    preds = []
    for aug, deaug in G.TTA_MODES[mode]:
        pr = predict_single(np.ascontiguousarray(aug(image)), mean, std, p, dtype)
        preds.append(deaug(pr).astype(np.float32))
    return G.tta_mean(preds)


# ------------------------------------------------------------------ T1
EPS = 1e-7  # K.epsilon()


def dice_loss(y, p):
    """train_adipose_unet_v3.py:217-225 (flattened over the WHOLE batch)."""
    pc = torch.clamp(p, EPS, 1.0 - EPS)
    inter = torch.sum(y * pc)
    return 1.0 - (2.0 * inter + 1.0) / (torch.sum(y) + torch.sum(pc) + 1.0)


def bce_mean(y, p):
    """keras.losses.binary_crossentropy on probabilities (Keras 2.13 backend):
    clip to [eps,1-eps], -[y log(p+eps) + (1-y) log(1-p+eps)], mean over the last
    axis, then Keras' loss reduction averages the rest => mean over all pixels."""
    pc = torch.clamp(p, EPS, 1.0 - EPS)
    bce = -(y * torch.log(pc + EPS) + (1.0 - y) * torch.log(1.0 - pc + EPS))
    return bce.mean()


def combined_loss_standard(y, p):
    """train_adipose_unet_v3.py:228-241: (B,1024)+scalar, reduced by mean."""
    return bce_mean(y, p) + dice_loss(y, p)


def smooth_labels(y, epsilon_pos=0.03, epsilon_neg=0.07):
    """Asymmetric label smoothing, train_adipose_unet_v3.py:271-274: 1 -> 1-eps_pos-eps_neg, 0 -> eps_neg."""
    return y * (1.0 - epsilon_pos - epsilon_neg) + epsilon_neg


def combined_loss_with_label_smoothing(y, p, epsilon_pos=0.03, epsilon_neg=0.07):
    """train_adipose_unet_v3.py:244-279."""
    ys = smooth_labels(y, epsilon_pos, epsilon_neg)
    return bce_mean(ys, p) + dice_loss(ys, p)


def bce_last_axis_mean(y, p):
    """tf.keras.losses.binary_crossentropy(y_true, y_pred) as a TENSOR (before Keras' loss reduction): the backend's
    element-wise BCE on clipped probabilities, averaged over the LAST axis only.  For the reference's (B,H,W) output and
    target (train_adipose_unet_v3.py:748-750, 611-613) the result is (B,H): one mean per image ROW."""
    pc = torch.clamp(p, EPS, 1.0 - EPS)
    bce = -(y * torch.log(pc + EPS) + (1.0 - y) * torch.log(1.0 - pc + EPS))
    return bce.mean(dim=-1)


def online_hard_example_mining_loss(y, p, keep_ratio=0.7, epsilon_pos=0.0, epsilon_neg=0.0):
    """train_adipose_unet_v3.py:282-323 (and :326-363 with smoothing), AS WRITTEN: `per_pixel_bce =
    tf.keras.losses.binary_crossentropy(y_true, y_pred)` (:301) averages the trailing axis, so for y, p of shape (B,H,W) it
    is the (B,H) tensor of per-row means; `flat_loss = reshape(.., [B,-1])` (:305) stays (B,H); `num_pixels` (:308) is H and
    k = int(float32(H) * keep_ratio) (:309) = 716 for H = 1024, ratio 0.7; `top_k` (:312) keeps the k largest ROW means of
    every image and `hard_bce` (:313) is their mean over the whole (B,k) tensor; dice_loss runs over ALL pixels (:316)."""
    ys = smooth_labels(y, epsilon_pos, epsilon_neg) if (epsilon_pos or epsilon_neg) else y
    flat = bce_last_axis_mean(ys, p).reshape(p.shape[0], -1)
    k = int(np.float32(flat.shape[1]) * np.float32(keep_ratio))
    top, _ = torch.topk(flat, k, dim=1, sorted=False)
    return top.mean() + dice_loss(ys, p)


def dice_coef(y, p):
    """src/utils/model.py:93-98 (no clip, smooth=1)."""
    return (2.0 * torch.sum(y * p) + 1.0) / (torch.sum(y) + torch.sum(p) + 1.0)


# ------------------------------------------------------------------ T2
def loss_and_grads(x: np.ndarray, y: np.ndarray, weights, dtype=torch.float32,
                   dropout_masks=None, loss_fn=None, ds_weights=None, loss_fn_aux=None):
    """One forward+backward of combined_loss_standard through the graph.
    x: (B,H,W) normalised float32, y: (B,H,W) {0,1}.  Returns loss, dice_coef,
    prob, dL/dprob and name -> (dW HWIO, db)."""
    p = to_torch_params(weights, dtype, requires_grad=True)
    xt = torch.from_numpy(np.ascontiguousarray(x)).to(dtype)
    yt = torch.from_numpy(np.ascontiguousarray(y)).to(dtype)
    dm = None
    if dropout_masks is not None:
        dm = {k: torch.from_numpy(v).to(dtype) for k, v in dropout_masks.items()}
    if ds_weights is not None:
        # compile(loss={main, aux1, aux2}, loss_weights=...) (:858-872): total = sum of weighted output losses
        prob, a1, a2 = forward(xt, p, dropout_masks=dm, deep_supervision=True)
        prob.retain_grad()
        fa = loss_fn_aux or combined_loss_standard
        loss = ds_weights[0] * (loss_fn or combined_loss_standard)(yt, prob) + ds_weights[1] * fa(yt, a1) + ds_weights[2] * fa(yt, a2)
    else:
        prob = forward(xt, p, dropout_masks=dm)
        prob.retain_grad()
        loss = (loss_fn or combined_loss_standard)(yt, prob)
    loss.backward()
    grads = {}
    for name, (w, b) in p.items():
        grads[name + "/kernel"] = w.grad.permute(2, 3, 1, 0).contiguous().to(torch.float32).numpy()
        grads[name + "/bias"] = b.grad.to(torch.float32).numpy()
    return (float(loss), float(dice_coef(yt, prob.detach())), prob.detach().to(torch.float32).numpy(),
            prob.grad.to(torch.float32).numpy(), grads)


# ------------------------------------------------------------------ T3
def keras_adam_step(theta, g, m, v, t: int, lr: float, beta1=0.9, beta2=0.999, eps=1e-7,
                    weight_decay: float = 0.0):
    """Keras 2.13 Adam.update_step (epsilon OUTSIDE the bias correction):
    alpha = lr*sqrt(1-b2^t)/(1-b1^t); m += (g-m)(1-b1); v += (g^2-v)(1-b2);
    theta -= alpha*m/(sqrt(v)+eps).  AdamW first does theta -= theta*wd*lr.
    Reference call sites: train_adipose_unet_v3.py:801-806.  float32 in, float32 out;
    t is 1-based (iterations+1)."""
    theta = theta.astype(np.float32).copy(); m = m.astype(np.float32).copy(); v = v.astype(np.float32).copy()
    g = g.astype(np.float32)
    if weight_decay:
        theta = theta - theta * np.float32(weight_decay) * np.float32(lr)
    b1p = np.float32(beta1) ** np.float32(t)
    b2p = np.float32(beta2) ** np.float32(t)
    alpha = np.float32(lr) * np.sqrt(np.float32(1) - b2p) / (np.float32(1) - b1p)
    m = m + (g - m) * np.float32(1 - beta1)
    v = v + (g * g - v) * np.float32(1 - beta2)
    theta = theta - (m * alpha) / (np.sqrt(v) + np.float32(eps))
    return theta.astype(np.float32), m.astype(np.float32), v.astype(np.float32)
