"""Layer table of the adipose U-Net (graph order == Keras creation order).

Follows /root/reference Segmentation/train_adipose_unet_v3.py:660-752
(identical inference clones: full_evaluation_enhanced.py:1163-1264,
segmentation_inference.py:88-146).  Names are the Keras layer names and
therefore the keys of the legacy ``.weights.h5`` layout.

Each entry: (name, cin, cout, ksize, dilation).  Kernel layout is Keras HWIO
``(k, k, cin, cout)`` float32, bias ``(cout,)`` float32.
"""
from __future__ import annotations

INIT_NB = 44
TILE = 1024


def conv_layers(init_nb: int = INIT_NB):
    c1, c2, c4, c8 = init_nb, init_nb * 2, init_nb * 4, init_nb * 8
    return [
        ("down1_conv1", 1, c1, 3, 1),
        ("down1_conv2", c1, c1, 3, 1),
        ("down2_conv1", c1, c2, 3, 1),
        ("down2_conv2", c2, c2, 3, 1),
        ("down3_conv1", c2, c4, 3, 1),
        ("down3_conv2", c4, c4, 3, 1),
        ("dilate1", c4, c8, 3, 1),
        ("dilate2", c8, c8, 3, 2),
        ("dilate3", c8, c8, 3, 4),
        ("dilate4", c8, c8, 3, 8),
        ("dilate5", c8, c8, 3, 16),
        ("dilate6", c8, c8, 3, 32),
        ("up3_conv1", c8, c4, 3, 1),
        ("up3_conv2", c8, c4, 3, 1),   # cin = [down3 | up3_conv1]
        ("up3_conv3", c4, c4, 3, 1),
        ("up2_conv1", c4, c2, 3, 1),
        ("up2_conv2", c4, c2, 3, 1),   # cin = [down2 | up2_conv1]
        ("up2_conv3", c2, c2, 3, 1),
        ("up1_conv1", c2, c1, 3, 1),
        ("up1_conv2", c2, c1, 3, 1),   # cin = [down1 | up1_conv1]
        ("up1_conv3", c1, c1, 3, 1),
        ("output_softmax", c1, 2, 1, 1),
    ]


LAYER_NAMES = [l[0] for l in conv_layers()]


AUX_NAMES = ("aux_out1", "aux_out2")      # deep-supervision heads (train_adipose_unet_v3.py:712-725)


def aux_weight_shapes(init_nb: int = INIT_NB):
    return {"aux_out1": ((1, 1, 4 * init_nb, 1), (1,)), "aux_out2": ((1, 1, 2 * init_nb, 1), (1,))}


def weight_shapes(init_nb: int = INIT_NB):
    """name -> (kernel_shape HWIO, bias_shape)."""
    return {n: ((k, k, ci, co), (co,)) for n, ci, co, k, _ in conv_layers(init_nb)}


def param_count(init_nb: int = INIT_NB) -> int:
    return sum(k * k * ci * co + co for _, ci, co, k, _ in conv_layers(init_nb))


def forward_flops(size: int = TILE, init_nb: int = INIT_NB) -> float:
    """Algorithmic FLOPs of one forward (2*H*W*k*k*Cin*Cout, unpadded channels),
    the numerator SURVEY.md section 8(d) fixes: 896.3 GFLOP at 1024^2."""
    res = {"down1": 1, "down2": 2, "down3": 4, "dilate": 8, "up3": 4, "up2": 2, "up1": 1, "output": 1}
    total = 0.0
    for n, ci, co, k, _ in conv_layers(init_nb):
        for pref, div in res.items():
            if n.startswith(pref):
                hw = (size // div) ** 2
                break
        total += 2.0 * hw * k * k * ci * co
    return total
