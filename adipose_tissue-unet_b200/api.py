"""Host-side mirror of the reference's inference seam.

Same class names, method names, argument meaning and return types as
Segmentation/full_evaluation_enhanced.py (AdiposeUNet :1156-1353,
TestTimeAugmentation :522-600, GaussianBlender :115-183, LinearBlender :186-204,
SlidingWindowInference :207-329, binarize_prediction :716-718,
calculate_pixel_metrics :721-785) so that reconstruct_full_images.py /
segmentation_inference.py style callers work unchanged — but every array
operation runs in libadipose_b200.so on the GPU.  NumPy here is plumbing
(buffers, slicing views, float64 ratios of four integers).
"""
from __future__ import annotations

import ctypes as C
import math
import time
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .layers import LAYER_NAMES, weight_shapes, AUX_NAMES, aux_weight_shapes

TTA_OPCODES = {"minimal": [0, 4], "basic": [0, 4, 5, 1], "full": [0, 1, 2, 3, 4, 5, 6, 7]}


def _f32c(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


class Engine:
    """Owner of one adp_engine handle (one GPU, one stream)."""

    def __init__(self, precision: str = "bf16", device: int = 0, init_nb: int = 44, max_forwards: int = 16):
        self.lib = _lib.load()
        self.precision = precision
        h = C.c_void_p()
        _lib.check(self.lib.adp_create(device, _lib.PRECISIONS[precision], init_nb, max_forwards, C.byref(h)))
        self.h = h
        self.device = device
        self.init_nb = init_nb

    def close(self):
        if getattr(self, "h", None):
            self.lib.adp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- weights
    def set_weights(self, weights: Dict[str, np.ndarray]):
        """weights: '<layer>/kernel' (HWIO float32) and '<layer>/bias' per Keras layer name."""
        names = list(LAYER_NAMES) + [n for n in AUX_NAMES if n + "/kernel" in weights]     # aux heads: deep-supervision training only
        for name in names:
            k = _f32c(weights[name + "/kernel"])
            b = _f32c(weights[name + "/bias"])
            shp = (C.c_int64 * 4)(*k.shape)
            _lib.check(self.lib.adp_set_weight(self.h, name.encode(), _lib.ptr(k), shp, _lib.ptr(b), b.size))
            if name in AUX_NAMES:
                self._has_aux = True

    def get_weights(self) -> Dict[str, np.ndarray]:
        out = {}
        shapes = dict(weight_shapes(self.init_nb))
        if getattr(self, "_has_aux", False):
            shapes.update(aux_weight_shapes(self.init_nb))
        for name, (ks, bs) in shapes.items():
            k = np.empty(ks, np.float32)
            b = np.empty(bs, np.float32)
            _lib.check(self.lib.adp_get_weight(self.h, name.encode(), _lib.ptr(k), k.size, _lib.ptr(b), b.size))
            out[name + "/kernel"] = k
            out[name + "/bias"] = b
        return out

    # ---- inference
    def predict(self, tiles, mean: float, std: float, ops: Optional[Sequence[int]] = None, out=None):
        """tiles: (n,S,S) float32 / uint8 (n,S,S) or (n,S,S,3); host ndarray or CUDA torch tensor.
        Returns (n,S,S) float32 probabilities (ndarray, or `out` if given)."""
        is_np = isinstance(tiles, np.ndarray)
        shape = tuple(tiles.shape)
        n, S = shape[0], shape[1]
        assert shape[2] == S
        ops = list(ops) if ops else []
        ops_arr = _lib.int_array(ops) if ops else None
        if out is None:
            out = np.empty((n, S, S), np.float32)
        dt = str(tiles.dtype)
        if "uint8" in dt:
            ch = shape[3] if len(shape) == 4 else 1
            if is_np:
                tiles = np.ascontiguousarray(tiles)
            _lib.check(self.lib.adp_predict_u8(self.h, _lib.ptr(tiles), n, S, ch, mean, std, ops_arr, len(ops), _lib.ptr(out)))
        else:
            if is_np:
                tiles = _f32c(tiles)
            _lib.check(self.lib.adp_predict(self.h, _lib.ptr(tiles), n, S, mean, std, ops_arr, len(ops), _lib.ptr(out)))
        return out

    def debug_layer(self, name: str, idx: int = 0) -> np.ndarray:
        shp = (C.c_int64 * 3)()
        _lib.check(self.lib.adp_debug_layer(self.h, name.encode(), idx, None, 0, shp))
        out = np.empty(tuple(shp), np.float32)
        _lib.check(self.lib.adp_debug_layer(self.h, name.encode(), idx, _lib.ptr(out), out.size, shp))
        return out

    def tta_combine(self, planes: np.ndarray, ops: Sequence[int]) -> np.ndarray:
        planes = _f32c(planes)
        k, S, _ = planes.shape
        out = np.empty((S, S), np.float32)
        _lib.check(self.lib.adp_tta_combine(self.h, _lib.ptr(planes), S, _lib.int_array(list(ops)), k, _lib.ptr(out)))
        return out

    def threshold_metrics(self, prob, gt=None, threshold: float = 0.5, want_mask: bool = True):
        n = int(np.prod(prob.shape))
        if isinstance(prob, np.ndarray):
            prob = _f32c(prob)
        mask = np.empty(prob.shape, np.uint8) if want_mask else None
        counts = (C.c_int64 * 4)()
        g = None
        if gt is not None:
            g = np.ascontiguousarray((np.asarray(gt) > 0.5).astype(np.uint8)) if isinstance(gt, np.ndarray) else gt
        _lib.check(self.lib.adp_threshold_metrics(self.h, _lib.ptr(prob), _lib.ptr(g), n, threshold, _lib.ptr(mask), counts))
        return mask, tuple(int(c) for c in counts)

    def threshold_sweep(self, prob, gt, thresholds) -> np.ndarray:
        """(n_thr, 4) int64 {tp, fp, fn, tn} for ascending candidate thresholds in one device pass."""
        prob = _f32c(prob)
        g = np.ascontiguousarray((np.asarray(gt) > 0.5).astype(np.uint8))
        thr = np.ascontiguousarray(thresholds, dtype=np.float32)
        counts = np.zeros((len(thr), 4), np.int64)
        _lib.check(self.lib.adp_threshold_sweep(self.h, _lib.ptr(prob), _lib.ptr(g), prob.size, _lib.ptr(thr), len(thr), _lib.ptr(counts)))
        return counts

    def boundary_refine(self, mask, kernel_size: int = 5, bilateral_d: int = 5, sigma_color: float = 50.0,
                        sigma_space: float = 50.0, out=None):
        """BoundaryRefiner.refine on the device.  mask: (H,W) or (n,H,W) float32, host ndarray or CUDA torch tensor."""
        is_np = isinstance(mask, np.ndarray)
        if is_np:
            mask = _f32c(mask)
        shape = tuple(mask.shape)
        n, (h, w) = (1, shape) if len(shape) == 2 else (shape[0], shape[1:])
        if out is None:
            out = np.empty(shape, np.float32)
        _lib.check(self.lib.adp_boundary_refine(self.h, _lib.ptr(mask), n, h, w, int(kernel_size), int(bilateral_d),
                                                float(sigma_color), float(sigma_space), _lib.ptr(out)))
        return out

    def blend(self, mode: int, tiles: Sequence[np.ndarray], positions, output_shape, window: Optional[np.ndarray]):
        h, w = int(output_shape[0]), int(output_shape[1])
        out = np.empty((h, w), np.float32)
        n = len(tiles)
        if n == 0:
            out[:] = 0
            return out
        th, tw = tiles[0].shape[:2]
        if any(t.shape[:2] != (th, tw) for t in tiles):
            raise ValueError("all tiles must have the same shape")
        arr = _f32c(np.stack(tiles)) if not (isinstance(tiles, np.ndarray) and tiles.ndim == 3) else _f32c(tiles)
        ys = np.ascontiguousarray([p[0] for p in positions], dtype=np.int32)
        xs = np.ascontiguousarray([p[1] for p in positions], dtype=np.int32)
        win = None
        if window is not None:
            win = _f32c(window[:th, :tw])
        _lib.check(self.lib.adp_blend_reconstruct(self.h, mode, _lib.ptr(arr), n, th, tw, _lib.ptr(ys), _lib.ptr(xs),
                                                  _lib.ptr(win), h, w, _lib.ptr(out)))
        return out

    def loss_metrics(self, p, y, want_grad: bool = False, ohem_keep_ratio: float = 1.0, eps_pos: float = 0.0,
                     eps_neg: float = 0.0):
        """combined_loss_standard by default; ohem_keep_ratio < 1 / eps_* > 0 select the reference's hard-mining and
        label-smoothing losses.  p, y: (B, H, W) (or a single image); the last axis is the one Keras' binary_crossentropy
        averages, i.e. hard mining ranks the (B, H) row means (train_adipose_unet_v3.py:301-313)."""
        p = _f32c(p); y = _f32c(y)
        out = (C.c_double * 4)()
        g = np.empty(p.shape, np.float32) if want_grad else None
        batch = p.shape[0] if p.ndim == 3 else 1
        row_len = p.shape[-1] if p.ndim >= 2 else 0
        _lib.check(self.lib.adp_loss_metrics_ex(self.h, _lib.ptr(p), _lib.ptr(y), batch, p.size // batch, row_len, ohem_keep_ratio,
                                                eps_pos, eps_neg, _lib.ptr(g), out))
        res = dict(loss=out[0], bce=out[1], dice_loss=out[2], dice_coef=out[3])
        return (res, g) if want_grad else res

    # ---- training step (Keras train_step: train_adipose_unet_v3.py:1316-1324)
    DROPOUT_SITES = ("dropout_dilate1", "dropout_up3", "dropout_up2", "dropout_up1")

    def train_begin(self, batch: int, size: int, dropout_rate: float = 0.3, seed: int = 865):
        _lib.check(self.lib.adp_train_begin(self.h, batch, size, dropout_rate, seed))
        self._train_shape = (batch, size, size)

    def train_forward(self, x, y, dropout_masks: Optional[Dict[str, np.ndarray]] = None, want_sums: bool = True):
        """x, y: (B,S,S) float32 (normalised image, target).  dropout_masks: site name -> (B,h,w,C) 0/1.
        Returns the eight loss sums (float64) of this batch (include/adipose_b200.h); want_sums=False leaves them in the
        device buffer (train_sums_buffer) and returns None without synchronising."""
        if isinstance(x, np.ndarray):
            x = _f32c(x)
        if isinstance(y, np.ndarray):
            y = _f32c(y)
        nout = int(self.lib.adp_train_outputs(self.h))
        sums = (C.c_double * (8 * nout))()
        mptr = None
        keep = []
        if dropout_masks is not None:
            arr = (C.c_void_p * 4)()
            for i, site in enumerate(self.DROPOUT_SITES):
                m = np.ascontiguousarray(dropout_masks[site], dtype=np.uint8)
                keep.append(m)
                arr[i] = m.ctypes.data
            mptr = arr
        _lib.check(self.lib.adp_train_forward(self.h, _lib.ptr(x), _lib.ptr(y), int(x.shape[0]), mptr, sums if want_sums else None))
        return np.array(list(sums), dtype=np.float64) if want_sums else None

    def train_sums_buffer(self) -> Tuple[int, int]:
        """(device address, count) of the float64 loss sums of the last forward (8 per output)."""
        p = C.c_void_p(); n = C.c_int()
        _lib.check(self.lib.adp_train_sums_buffer(self.h, C.byref(p), C.byref(n)))
        return int(p.value), int(n.value)

    def train_sums_read(self) -> np.ndarray:
        """Host copy of the device loss sums (synchronises the engine's stream)."""
        n = 8 * int(self.lib.adp_train_outputs(self.h))
        out = np.empty(n, np.float64)
        _lib.check(self.lib.adp_train_sums_read(self.h, _lib.ptr(out), n))
        return out

    def train_set_loss(self, ohem_keep_ratio: float = 1.0, eps_pos: float = 0.0, eps_neg: float = 0.0):
        """Loss recipe of the following steps (train_adipose_unet_v3.py:808-855)."""
        _lib.check(self.lib.adp_train_set_loss(self.h, ohem_keep_ratio, eps_pos, eps_neg))

    def train_set_deep_supervision(self, on: bool = True, w_main: float = 1.0, w_aux1: float = 0.4, w_aux2: float = 0.3):
        """aux_out1 / aux_out2 heads with loss weights (train_adipose_unet_v3.py:712-745, 858-872); needs their weights set."""
        _lib.check(self.lib.adp_train_set_deep_supervision(self.h, 1 if on else 0, w_main, w_aux1, w_aux2))
        self._ds_weights = (w_main, w_aux1, w_aux2) if on else None

    def train_loss(self, sums) -> Dict[str, float]:
        """{loss, bce, dice_loss, dice_coef}: with deep supervision `loss` is the weighted total of the three outputs and
        bce / dice_loss / dice_coef are those of main_out (what Keras logs as main_out_*)."""
        sums = np.asarray(sums, dtype=np.float64)
        outs = []
        for o in range(len(sums) // 8):
            s = (C.c_double * 8)(*[float(v) for v in sums[8 * o:8 * o + 8]])
            out = (C.c_double * 4)()
            _lib.check(self.lib.adp_train_loss(s, out))
            outs.append(dict(loss=out[0], bce=out[1], dice_loss=out[2], dice_coef=out[3]))
        res = dict(outs[0])
        dsw = getattr(self, "_ds_weights", None)
        if len(outs) == 3 and dsw is not None:
            res["loss"] = dsw[0] * outs[0]["loss"] + dsw[1] * outs[1]["loss"] + dsw[2] * outs[2]["loss"]
            res["main_out_loss"], res["aux_out1_loss"], res["aux_out2_loss"] = outs[0]["loss"], outs[1]["loss"], outs[2]["loss"]
        return res

    def train_backward(self, sums=None, freeze_encoder: bool = False):
        """sums: host values the loss is defined over, or None = the device sums buffer as it stands."""
        s = (C.c_double * len(sums))(*[float(v) for v in sums]) if sums is not None else None
        _lib.check(self.lib.adp_train_backward(self.h, s, 1 if freeze_encoder else 0))

    def train_grad_buffer(self) -> Tuple[int, int]:
        """(device address, element count) of the flat fp32 gradient — what a data-parallel wrapper all-reduces."""
        p = C.c_void_p(); n = C.c_int64()
        _lib.check(self.lib.adp_train_grad_buffer(self.h, C.byref(p), C.byref(n)))
        return int(p.value), int(n.value)

    def train_accuracy_read(self) -> Tuple[float, float]:
        """(matching pixels, pixels) of Keras' binary_accuracy for the last backward's batch (needs set_option('train_accuracy', 1))."""
        out = (C.c_double * 2)()
        _lib.check(self.lib.adp_train_accuracy_read(self.h, out))
        return float(out[0]), float(out[1])

    def train_grad_buckets(self) -> List[Tuple[int, int]]:
        """[lo, hi) element ranges of the flat gradient in the order the backward pass completes them."""
        lo = (C.c_int64 * 8)(); hi = (C.c_int64 * 8)()
        n = _lib.check(self.lib.adp_train_grad_buckets(self.h, lo, hi, 8))
        return [(int(lo[i]), int(hi[i])) for i in range(n)]

    def train_bucket_wait(self, bucket: int, stream_ptr: int):
        _lib.check(self.lib.adp_train_bucket_wait(self.h, bucket, C.c_void_p(stream_ptr)))

    def train_join(self, stream_ptr: int):
        _lib.check(self.lib.adp_train_join(self.h, C.c_void_p(stream_ptr)))

    def grad_to_host(self) -> np.ndarray:
        _, n = self.train_grad_buffer()
        g = np.empty(n, np.float32)
        _lib.check(self.lib.adp_train_grad_read(self.h, _lib.ptr(g), n))
        return g

    def grad_from_host(self, g: np.ndarray):
        g = _f32c(g).ravel()
        _lib.check(self.lib.adp_train_grad_write(self.h, _lib.ptr(g), g.size))

    def train_grads(self) -> Dict[str, np.ndarray]:
        out = {}
        shapes = dict(weight_shapes(self.init_nb))
        if int(self.lib.adp_train_outputs(self.h)) == 3:
            shapes.update(aux_weight_shapes(self.init_nb))
        for name, (ks, bs) in shapes.items():
            k = np.empty(ks, np.float32); b = np.empty(bs, np.float32)
            _lib.check(self.lib.adp_train_get_grad(self.h, name.encode(), _lib.ptr(k), k.size, _lib.ptr(b), b.size))
            out[name + "/kernel"] = k; out[name + "/bias"] = b
        return out

    def train_probs(self) -> np.ndarray:
        out = np.empty(self._train_shape, np.float32)
        _lib.check(self.lib.adp_train_probs(self.h, _lib.ptr(out), out.size))
        return out

    def train_apply(self, lr: float, optimizer: str = "adam", grad_scale: float = 1.0, beta1: float = 0.9,
                    beta2: float = 0.999, eps: float = 1e-7, weight_decay: float = 0.01, freeze_encoder: bool = False):
        _lib.check(self.lib.adp_train_apply(self.h, _lib.OPTIMIZERS[optimizer], lr, grad_scale, beta1, beta2, eps,
                                            weight_decay, 1 if freeze_encoder else 0))

    def train_step(self, x, y, lr: float, optimizer: str = "adam", weight_decay: float = 0.01,
                   freeze_encoder: bool = False) -> Dict[str, float]:
        if isinstance(x, np.ndarray):
            x = _f32c(x)
        if isinstance(y, np.ndarray):
            y = _f32c(y)
        out = (C.c_double * 4)()
        _lib.check(self.lib.adp_train_step(self.h, _lib.ptr(x), _lib.ptr(y), int(x.shape[0]), _lib.OPTIMIZERS[optimizer], lr,
                                           weight_decay, 1 if freeze_encoder else 0, out))
        return dict(loss=out[0], bce=out[1], dice_loss=out[2], dice_coef=out[3])

    def train_iterations(self) -> int:
        return int(self.lib.adp_train_iterations(self.h))

    def train_end(self):
        _lib.check(self.lib.adp_train_end(self.h))

    def adam_update(self, theta, grad, m, v, t: int, lr: float, optimizer: str = "adam", beta1: float = 0.9,
                    beta2: float = 0.999, eps: float = 1e-7, weight_decay: float = 0.0):
        """One Keras-2.13 Adam/AdamW update on caller arrays (returns new theta, m, v)."""
        theta = _f32c(theta).copy(); m = _f32c(m).copy(); v = _f32c(v).copy(); grad = _f32c(grad)
        _lib.check(self.lib.adp_adam_update(self.h, _lib.ptr(theta), _lib.ptr(grad), _lib.ptr(m), _lib.ptr(v), theta.size, t,
                                            _lib.OPTIMIZERS[optimizer], lr, beta1, beta2, eps, weight_decay))
        return theta, m, v

    # ---- whole-slide accumulator
    def wsi_begin(self, rows, W, y0, tile, mode, window):
        win = _f32c(window) if window is not None else None
        _lib.check(self.lib.adp_wsi_begin(self.h, rows, W, y0, tile, mode, _lib.ptr(win)))

    def wsi_push_tiles(self, tiles, ys, xs, mean, std, ops):
        if isinstance(tiles, np.ndarray):
            tiles = _f32c(tiles)
        ys = np.ascontiguousarray(ys, dtype=np.int32); xs = np.ascontiguousarray(xs, dtype=np.int32)
        ops = list(ops) if ops else []
        _lib.check(self.lib.adp_wsi_push_tiles(self.h, _lib.ptr(tiles), len(ys), _lib.ptr(ys), _lib.ptr(xs), mean, std,
                                               _lib.int_array(ops) if ops else None, len(ops)))

    def wsi_push_from_slide(self, slide_dev, region_y0, region_rows, ys, xs, mean, std, ops, channels: int = 1,
                            defer_below_row: int = 0):
        """slide_dev: device uint8 region [region_rows, W] (gray) or [region_rows, W, 3] (RGB, channels=3).
        defer_below_row: see adp_wsi_push_from_slide (exact multi-GPU boundary zone)."""
        ys = np.ascontiguousarray(ys, dtype=np.int32); xs = np.ascontiguousarray(xs, dtype=np.int32)
        ops = list(ops) if ops else []
        _lib.check(self.lib.adp_wsi_push_from_slide(self.h, _lib.ptr(slide_dev), int(channels), region_y0, region_rows, len(ys),
                                                    _lib.ptr(ys), _lib.ptr(xs), mean, std,
                                                    _lib.int_array(ops) if ops else None, len(ops), int(defer_below_row)))

    def wsi_replay_deferred(self):
        _lib.check(self.lib.adp_wsi_replay_deferred(self.h))

    def wsi_push_probs(self, probs, ys, xs):
        probs = _f32c(probs)
        ys = np.ascontiguousarray(ys, dtype=np.int32); xs = np.ascontiguousarray(xs, dtype=np.int32)
        _lib.check(self.lib.adp_wsi_push_probs(self.h, _lib.ptr(probs), len(ys), _lib.ptr(ys), _lib.ptr(xs)))

    def wsi_export(self, y, rows, W):
        acc = np.empty((rows, W), np.float32); wt = np.empty((rows, W), np.float32)
        _lib.check(self.lib.adp_wsi_export(self.h, y, rows, _lib.ptr(acc), _lib.ptr(wt)))
        return acc, wt

    def wsi_export_into(self, y, rows, acc, wt):
        """Raw accumulator rows into caller buffers (NumPy arrays or device tensors with data_ptr())."""
        _lib.check(self.lib.adp_wsi_export(self.h, y, rows, _lib.ptr(acc), _lib.ptr(wt)))

    def wsi_import_add(self, y, acc, wt):
        if isinstance(acc, np.ndarray):
            acc = _f32c(acc); wt = _f32c(wt)
        _lib.check(self.lib.adp_wsi_import_add(self.h, y, int(acc.shape[0]), _lib.ptr(acc), _lib.ptr(wt)))

    def wsi_finalize(self, y, rows, W, threshold=0.5, gt=None, want_prob=True, want_mask=True):
        prob = np.empty((rows, W), np.float32) if want_prob else None
        mask = np.empty((rows, W), np.uint8) if want_mask else None
        counts = (C.c_int64 * 4)()
        g = None
        if gt is not None:          # the kernel tests gt != 0, which equals the reference's gt > 0.5 for uint8 masks: no host pass
            g = np.ascontiguousarray(gt) if getattr(gt, "dtype", None) == np.uint8 else \
                np.ascontiguousarray((np.asarray(gt) > 0.5).astype(np.uint8))
        _lib.check(self.lib.adp_wsi_finalize(self.h, y, rows, threshold, _lib.ptr(prob), _lib.ptr(mask), _lib.ptr(g), counts))
        return prob, mask, tuple(int(c) for c in counts)

    def wsi_end(self):
        _lib.check(self.lib.adp_wsi_end(self.h))

    # ---- tile I/O front-end
    def jpeg_decode(self, blobs: Sequence[bytes], size: int, want_gray: bool = True, want_rgb: bool = False, to_host: bool = True):
        """nvJPEG decode of JPEG byte strings (size x size tiles).  Returns dict with 'gray' (n,S,S) / 'rgb' (n,S,S,3) uint8 host
        arrays (to_host) and 'gray_dev' / 'rgb_dev' device addresses (valid until the next decode) for predict / wsi_push_*."""
        n = len(blobs)
        bufs = [np.frombuffer(b, dtype=np.uint8) for b in blobs]
        ptrs = (C.c_void_p * n)(*[b.ctypes.data for b in bufs])
        lens = (C.c_size_t * n)(*[b.size for b in bufs])
        gray = np.empty((n, size, size), np.uint8) if (want_gray and to_host) else None
        rgb = np.empty((n, size, size, 3), np.uint8) if (want_rgb and to_host) else None
        gd, rd = C.c_void_p(), C.c_void_p()
        _lib.check(self.lib.adp_jpeg_decode(self.h, ptrs, lens, n, size, _lib.ptr(gray), _lib.ptr(rgb),
                                            C.byref(gd) if want_gray else None, C.byref(rd) if want_rgb else None))
        return dict(gray=gray, rgb=rgb, gray_dev=gd.value, rgb_dev=rd.value, n=n, size=size)

    def predict_u8_dev(self, dev_ptr: int, n: int, size: int, channels: int, mean: float, std: float,
                       ops: Optional[Sequence[int]] = None) -> np.ndarray:
        """adp_predict_u8 on device-resident uint8 tiles (a device address, e.g. the output of jpeg_decode) or a host uint8 array."""
        ops = list(ops) if ops else []
        out = np.empty((n, size, size), np.float32)
        if isinstance(dev_ptr, np.ndarray):
            dev_ptr = np.ascontiguousarray(dev_ptr, dtype=np.uint8)
        _lib.check(self.lib.adp_predict_u8(self.h, _lib.ptr(dev_ptr), n, size, channels, mean, std,
                                           _lib.int_array(ops) if ops else None, len(ops), _lib.ptr(out)))
        return out

    def wsi_push_tiles_u8(self, tiles, ys, xs, mean, std, ops, channels: int = 1):
        """tiles: uint8 ndarray (n,S,S[,3]) or a device address (int)."""
        if isinstance(tiles, np.ndarray):
            tiles = np.ascontiguousarray(tiles, dtype=np.uint8)
        ys = np.ascontiguousarray(ys, dtype=np.int32); xs = np.ascontiguousarray(xs, dtype=np.int32)
        ops = list(ops) if ops else []
        _lib.check(self.lib.adp_wsi_push_tiles_u8(self.h, _lib.ptr(tiles), int(channels), len(ys), _lib.ptr(ys), _lib.ptr(xs), mean, std,
                                                  _lib.int_array(ops) if ops else None, len(ops)))

    def wsi_aux_begin(self, n_planes: int):
        _lib.check(self.lib.adp_wsi_aux_begin(self.h, n_planes))

    def wsi_push_aux(self, plane0: int, tiles, ys, xs, n_planes: int = 1, is_u8: Optional[bool] = None):
        """Blend tiles into auxiliary planes plane0..plane0+n_planes-1: uint8 (scaled by 1/255) or float32; ndarray
        (n,S,S[,n_planes]) or a device address (then is_u8 must be given)."""
        if isinstance(tiles, np.ndarray):
            is_u8 = tiles.dtype == np.uint8
            tiles = np.ascontiguousarray(tiles) if is_u8 else _f32c(tiles)
        ys = np.ascontiguousarray(ys, dtype=np.int32); xs = np.ascontiguousarray(xs, dtype=np.int32)
        _lib.check(self.lib.adp_wsi_push_aux(self.h, plane0, n_planes, _lib.ptr(tiles), 1 if is_u8 else 0, len(ys), _lib.ptr(ys), _lib.ptr(xs)))

    def wsi_export_u8(self, plane0: int, n_planes: int, y: int, rows: int, W: int, reverse: bool = False) -> np.ndarray:
        """(normalised plane * 255).astype(uint8); plane0 = -1: the probability plane.  (rows, W) or (rows, W, n_planes)."""
        out = np.empty((rows, W) if n_planes == 1 else (rows, W, n_planes), np.uint8)
        _lib.check(self.lib.adp_wsi_export_u8(self.h, plane0, n_planes, 1 if reverse else 0, y, rows, _lib.ptr(out)))
        return out

    def wsi_export_f32(self, plane: int, y: int, rows: int, W: int) -> np.ndarray:
        out = np.empty((rows, W), np.float32)
        _lib.check(self.lib.adp_wsi_export_f32(self.h, plane, y, rows, _lib.ptr(out)))
        return out

    def wsi_finalize_auxgt(self, gt_plane: int, y: int, rows: int, W: int, threshold: float = 0.5, want_prob=True, want_mask=True):
        prob = np.empty((rows, W), np.float32) if want_prob else None
        mask = np.empty((rows, W), np.uint8) if want_mask else None
        counts = (C.c_int64 * 4)()
        _lib.check(self.lib.adp_wsi_finalize_auxgt(self.h, gt_plane, y, rows, threshold, _lib.ptr(prob), _lib.ptr(mask), counts))
        return prob, mask, tuple(int(c) for c in counts)

    def fat_percent(self, probs, threshold: float = 0.5) -> np.ndarray:
        """calculate_fat_percentage (tile_classification_evaluation.py:211-225) for a batch (n,H,W) or one tile (H,W)."""
        p = _f32c(probs)
        n = p.shape[0] if p.ndim == 3 else 1
        out = np.empty(n, np.float64)
        _lib.check(self.lib.adp_tile_fat_percent(self.h, _lib.ptr(p), n, p.size // n, threshold, _lib.ptr(out)))
        return out

    # ---- profiling
    def profile(self, on: bool):
        self.lib.adp_profile_enable(self.h, 1 if on else 0)
        if on:
            self.lib.adp_profile_reset(self.h)

    def profile_rows(self) -> List[dict]:
        rows = (_lib.ProfRow * 512)()
        n = self.lib.adp_profile_read(self.h, rows, 512)
        return [dict(name=rows[i].name.decode(), launches=rows[i].launches, ms=rows[i].ms, flops=rows[i].flops,
                     bytes=rows[i].bytes) for i in range(n)]

    def launch_count(self) -> int:
        return int(self.lib.adp_launch_count(self.h))

    def synchronize(self):
        _lib.check(self.lib.adp_synchronize(self.h))

    def set_option(self, key: str, value: int):
        _lib.check(self.lib.adp_set_option(self.h, key.encode(), int(value)))

    def stream_ptr(self) -> int:
        return int(self.lib.adp_stream(self.h) or 0)


_default_engine: Optional[Engine] = None


def default_engine() -> Engine:
    """Engine used by the stateless reference-style objects (blenders, metrics)."""
    global _default_engine
    if _default_engine is None:
        _default_engine = Engine(precision="fp32", max_forwards=8)
    return _default_engine


# =====================================================================================
# Reference-shaped objects
# =====================================================================================
class AdiposeUNet:
    """Drop-in for full_evaluation_enhanced.AdiposeUNet / segmentation_inference.AdiposeUNet."""

    def __init__(self, precision: str = "bf16", device: int = 0, max_forwards: int = 16):
        self.net = None
        self.use_deep_supervision = False
        self._precision, self._device, self._max_forwards = precision, device, max_forwards
        self.engine: Optional[Engine] = None

    def build_model(self, init_nb: int = 44, dropout_rate: float = 0.3, use_deep_supervision: bool = False):
        # Dropout is identity at inference; the deep-supervision heads do not feed main_out
        # (full_evaluation_enhanced.py:1313-1319 only ever returns 'main_out').
        self.use_deep_supervision = use_deep_supervision
        self.engine = Engine(self._precision, self._device, init_nb, self._max_forwards)
        self.net = self.engine
        return self.net

    def set_weights(self, weights: Dict[str, np.ndarray]):
        self.engine.set_weights(weights)

    def load_weights(self, weights_path: str):
        from .weights_io import load_weights_file
        self.engine.set_weights(load_weights_file(weights_path, self.engine.init_nb))
        print(f"✓ Loaded weights from {weights_path}")

    def predict_single(self, image: np.ndarray, mean: float, std: float) -> np.ndarray:
        return self.engine.predict(np.asarray(image)[None], float(mean), float(std))[0]

    def predict_batch(self, tiles, mean: float, std: float, tta_mode: Optional[str] = None):
        ops = TTA_OPCODES[tta_mode] if tta_mode else None
        return self.engine.predict(tiles, float(mean), float(std), ops)

    def predict(self, image: np.ndarray, mean: float, std: float, use_tta: bool = False,
                tta_mode: str = "basic") -> Tuple[np.ndarray, dict]:
        start = time.time()
        if not use_tta:
            pred = self.predict_single(image, mean, std)
            return pred, {"num_augmentations": 1, "total_time": time.time() - start, "tta_enabled": False}
        pred, timing = TestTimeAugmentation(mode=tta_mode).predict_with_tta(self, image, mean, std)
        timing["tta_enabled"] = True
        timing["tta_mode"] = tta_mode
        return pred, timing


class TestTimeAugmentation:
    """full_evaluation_enhanced.TestTimeAugmentation (:522-600).  For the native model the
    eight forwards, the inverse dihedral ops and the mean run as one device batch; for any other
    object with predict_single (the reference's ONNX seam) the per-augmentation predictions are
    de-augmented and averaged by the device kernel."""
    __test__ = False

    def __init__(self, mode: str = "basic"):
        mode = (mode or "basic").lower()
        if mode not in TTA_OPCODES:
            mode = "basic"
        self.mode = mode
        self.ops = TTA_OPCODES[mode]
        self.transforms = [(_aug_fn(op), _aug_fn(_INV[op])) for op in self.ops]

    def predict_with_tta(self, model_wrapper: Any, image: np.ndarray, mean: float, std: float):
        start = time.time()
        if isinstance(model_wrapper, AdiposeUNet):
            avg = model_wrapper.engine.predict(np.asarray(image)[None], float(mean), float(std), self.ops)[0]
        else:
            planes = [np.asarray(model_wrapper.predict_single(aug(image), mean, std), dtype=np.float32)
                      for aug, _ in self.transforms]
            avg = default_engine().tta_combine(np.stack(planes), self.ops)
        return avg, {"num_augmentations": len(self.ops), "total_time": time.time() - start}


_INV = [0, 3, 2, 1, 4, 5, 6, 7]


def _aug_fn(op: int):
    """NumPy *view* of the dihedral op (host plumbing for foreign models only)."""
    fns = {0: lambda x: x, 1: lambda x: np.rot90(x, 1), 2: lambda x: np.rot90(x, 2), 3: lambda x: np.rot90(x, 3),
           4: lambda x: np.flip(x, axis=1), 5: lambda x: np.flip(x, axis=0),
           6: lambda x: np.flip(np.rot90(x, 1), axis=1), 7: lambda x: np.flip(np.rot90(x, 1), axis=0)}
    return fns[op]


class GaussianBlender:
    """full_evaluation_enhanced.GaussianBlender (:115-183)."""

    def __init__(self, tile_size: int = 1024, sigma_factor: float = 0.25, engine: Optional[Engine] = None):
        self.tile_size = tile_size
        self.sigma = tile_size * sigma_factor
        self.weight_map = self._create_gaussian_weight_map()
        self._engine = engine

    def _create_gaussian_weight_map(self) -> np.ndarray:
        # float64 -> /max -> float32, centre = tile/2 (:133-147); 4 MB computed once on the host
        center = self.tile_size / 2
        y, x = np.ogrid[0:self.tile_size, 0:self.tile_size]
        w = np.exp(-((x - center) ** 2 + (y - center) ** 2) / (2 * self.sigma ** 2))
        return (w / w.max()).astype(np.float32)

    def reconstruct(self, tiles, positions, output_shape) -> np.ndarray:
        eng = self._engine or default_engine()
        return eng.blend(_lib.BLEND_GAUSSIAN, tiles, positions, output_shape, self.weight_map)


class BoundaryRefiner:
    """full_evaluation_enhanced.BoundaryRefiner (:332-393): morphological boundary refinement of a [0,1] mask /
    probability map, on the device (adp_boundary_refine).  `image` is accepted and unused, as in the reference."""

    def __init__(self, kernel_size: int = 5, bilateral_d: int = 5, bilateral_sigma_color: float = 50,
                 bilateral_sigma_space: float = 50, engine: Optional[Engine] = None):
        self.kernel_size = kernel_size
        self.bilateral_d = bilateral_d
        self.sigma_color = bilateral_sigma_color
        self.sigma_space = bilateral_sigma_space
        self._engine = engine

    def refine(self, mask: np.ndarray, image: Optional[np.ndarray] = None) -> np.ndarray:
        eng = self._engine or default_engine()
        return eng.boundary_refine(np.asarray(mask, dtype=np.float32), self.kernel_size, self.bilateral_d, self.sigma_color,
                                   self.sigma_space)


class HannBlender:
    """EXTENSION (not in the reference, which has only Gaussian and linear blenders): Hann-window blending as named by
    BASELINE.json.  Weights h[y]*h[x], h[i] = max(0.5 - 0.5*cos(2*pi*(i + 0.5)/tile), 1e-3) in float64 -> float32: the
    half-sample shift keeps the window symmetric and positive, the floor keeps every weight above the 1e-8 clamp of the
    normalisation (slide corners are covered by a single tile), and h[i] + h[i + tile/2] == 1 at 50 % overlap away from
    the floored ends.  Same accumulate / normalise statements as GaussianBlender.reconstruct: same device kernels."""

    def __init__(self, tile_size: int = 1024, engine: Optional[Engine] = None):
        self.tile_size = tile_size
        self.weight_map = self._create_hann_weight_map()
        self._engine = engine

    def _create_hann_weight_map(self) -> np.ndarray:
        i = np.arange(self.tile_size, dtype=np.float64)
        h = np.maximum(0.5 - 0.5 * np.cos(2.0 * np.pi * (i + 0.5) / self.tile_size), 1e-3)
        return np.outer(h, h).astype(np.float32)

    def reconstruct(self, tiles, positions, output_shape) -> np.ndarray:
        eng = self._engine or default_engine()
        return eng.blend(_lib.BLEND_GAUSSIAN, tiles, positions, output_shape, self.weight_map)


def blend_window(blend_mode: str, tile_size: int = 1024) -> Optional[np.ndarray]:
    """Host window of a weighted blend mode ('gaussian', or the 'hann' extension); None for 'linear'."""
    if blend_mode == "gaussian":
        return GaussianBlender(tile_size).weight_map
    if blend_mode == "hann":
        return HannBlender(tile_size).weight_map
    return None


class LinearBlender:
    """full_evaluation_enhanced.LinearBlender (:186-204)."""

    def __init__(self, engine: Optional[Engine] = None):
        self._engine = engine

    def reconstruct(self, tiles, positions, output_shape) -> np.ndarray:
        eng = self._engine or default_engine()
        return eng.blend(_lib.BLEND_LINEAR, tiles, positions, output_shape, None)


class SlidingWindowInference:
    """full_evaluation_enhanced.SlidingWindowInference (:207-329)."""

    def __init__(self, tile_size: int = 1024, overlap: float = 0.5, blend_mode: str = "gaussian", verbose: bool = True):
        self.tile_size = tile_size
        self.overlap = max(0.0, min(overlap, 0.75))
        self.stride = int(tile_size * (1 - self.overlap))
        self.blend_mode = blend_mode
        if blend_mode == "gaussian":
            self.blender = GaussianBlender(tile_size)
        elif blend_mode == "linear":
            self.blender = LinearBlender()
        elif blend_mode == "hann":                      # extension, see HannBlender
            self.blender = HannBlender(tile_size)
        else:
            self.blender = None
        if verbose:
            print(f"[SlidingWindow] Initialized: tile={tile_size}, stride={self.stride}, "
                  f"overlap={overlap:.1%}, blend={blend_mode}")

    def extract_tile_positions(self, image_shape) -> List[Tuple[int, int]]:
        h, w = image_shape[:2]
        t, s = self.tile_size, self.stride
        positions = []
        y_steps = max(1, math.ceil((h - t) / s) + 1)
        x_steps = max(1, math.ceil((w - t) / s) + 1)
        for yi in range(y_steps):
            for xi in range(x_steps):
                y = min(yi * s, h - t)
                x = min(xi * s, w - t)
                if y >= 0 and x >= 0 and y + t <= h and x + t <= w:
                    positions.append((y, x))
        return positions

    def extract_tiles(self, image: np.ndarray):
        positions = self.extract_tile_positions(image.shape)
        t = self.tile_size
        return [image[y:y + t, x:x + t] for y, x in positions], positions

    def predict_with_sliding_window(self, image: np.ndarray, model, mean: float, std: float, use_tta: bool = False,
                                    tta_mode: str = "basic") -> np.ndarray:
        tiles, positions = self.extract_tiles(image)
        h, w = image.shape[:2]
        # window-weighted blenders (Gaussian, and the Hann extension) carry a weight_map; linear / none average uniformly
        win = getattr(self.blender, "weight_map", None)
        if isinstance(model, AdiposeUNet):
            eng = model.engine
            ops = TTA_OPCODES[tta_mode if tta_mode in TTA_OPCODES else "basic"] if use_tta else None
            eng.wsi_begin(h, w, 0, self.tile_size, _lib.BLEND_GAUSSIAN if win is not None else _lib.BLEND_LINEAR, win)
            try:
                step = 16
                for i in range(0, len(tiles), step):
                    chunk = np.stack([np.asarray(t, dtype=np.float32) for t in tiles[i:i + step]])
                    pos = positions[i:i + step]
                    eng.wsi_push_tiles(chunk, [p[0] for p in pos], [p[1] for p in pos], float(mean), float(std), ops)
                prob, _, _ = eng.wsi_finalize(0, h, w, want_mask=False)
            finally:
                eng.wsi_end()
            return prob
        preds = []
        for tile in tiles:
            if use_tta:
                pred, _ = TestTimeAugmentation(mode=tta_mode).predict_with_tta(model, tile, mean, std)
            else:
                pred = model.predict_single(tile, mean, std)
            preds.append(np.asarray(pred, dtype=np.float32))
        blender = self.blender if self.blender is not None else LinearBlender()
        return blender.reconstruct(preds, positions, (h, w))


def binarize_prediction(pred: np.ndarray, threshold: float = 0.5, engine: Optional[Engine] = None) -> np.ndarray:
    """full_evaluation_enhanced.binarize_prediction (:716-718)."""
    mask, _ = (engine or default_engine()).threshold_metrics(pred, None, threshold)
    return mask


def metrics_from_counts(tp: int, fp: int, fn: int, tn: int) -> Dict[str, float]:
    """Ratios of full_evaluation_enhanced.calculate_pixel_metrics (:735-785) from the four counts."""
    if tp == 0 and fp == 0 and fn == 0:
        return {"dice_score": 1.0, "jaccard_index": 1.0, "sensitivity": 1.0, "specificity": 1.0, "precision": 1.0,
                "f1_score": 1.0, "accuracy": 1.0, "tp": 0, "fp": 0, "fn": 0, "tn": int(tn)}
    precision = tp / (tp + fp + 1e-10)
    sensitivity = tp / (tp + fn + 1e-10)
    specificity = tn / (tn + fp + 1e-10)
    accuracy = (tp + tn) / (tp + fp + fn + tn + 1e-10)
    f1 = 2 * tp / (2 * tp + fp + fn + 1e-10)
    jaccard = tp / (tp + fp + fn + 1e-10)
    return {"dice_score": float(f1), "jaccard_index": float(jaccard), "sensitivity": float(sensitivity),
            "specificity": float(specificity), "precision": float(precision), "f1_score": float(f1),
            "accuracy": float(accuracy), "tp": int(tp), "fp": int(fp), "fn": int(fn), "tn": int(tn)}


def boundary_metrics_from_counts(tp: int, fp: int, fn: int, tn: int) -> Dict[str, float]:
    """full_evaluation_enhanced.calculate_boundary_metrics (:788-844) from the four confusion counts.

    The reference samples each mask's OWN distance transform on its OWN surface (`pred_dt[pred_surface]`,
    `true_dt[true_surface]`, :829-830): `pred_dt = distance_transform_edt(~pred_bin)` is 0 on every pixel of `pred_bin`,
    and a surface is a subset of its mask, so all sampled distances are 0 and both metrics are 0.0 whenever both
    surfaces are non-empty.  What remains of the function is its case analysis, which the counts decide:
    both masks empty -> 0.0; exactly one empty -> inf; a surface empty -> inf.  With skimage's binary_erosion
    (3x3 cross, outside pixels count as foreground) a non-empty mask has an empty surface only when it covers the
    whole image.  A drop-in follows the code as written, not the metric's name."""
    pred_any, true_any = (tp + fp) > 0, (tp + fn) > 0
    if not pred_any and not true_any:
        return {"hausdorff95": 0.0, "assd": 0.0}
    if not pred_any or not true_any:
        return {"hausdorff95": float("inf"), "assd": float("inf")}
    pred_all, true_all = (fn + tn) == 0, (fp + tn) == 0
    if pred_all or true_all:
        return {"hausdorff95": float("inf"), "assd": float("inf")}
    return {"hausdorff95": 0.0, "assd": 0.0}


def calculate_boundary_metrics(pred: np.ndarray, true: np.ndarray, threshold: float = 0.5, spacing=(1.0, 1.0),
                               engine: Optional[Engine] = None) -> Dict[str, float]:
    """full_evaluation_enhanced.calculate_boundary_metrics (:788-844): counts on the device, case analysis here
    (see boundary_metrics_from_counts for why no distance transform is needed to reproduce the reference)."""
    _, counts = (engine or default_engine()).threshold_metrics(pred, true, threshold, want_mask=False)
    return boundary_metrics_from_counts(*counts)


def calculate_pixel_metrics(pred: np.ndarray, true: np.ndarray, threshold: float = 0.5,
                            engine: Optional[Engine] = None) -> Dict[str, float]:
    """full_evaluation_enhanced.calculate_pixel_metrics (:721-785): counts on the device, ratios here."""
    _, (tp, fp, fn, tn) = (engine or default_engine()).threshold_metrics(pred, true, threshold, want_mask=False)
    return metrics_from_counts(tp, fp, fn, tn)


def calculate_fat_percentage(mask: np.ndarray, threshold: float = 0.5, engine: Optional[Engine] = None) -> float:
    """tile_classification_evaluation.calculate_fat_percentage (:211-225): percentage of pixels above the threshold."""
    return float((engine or default_engine()).fat_percent(np.asarray(mask, dtype=np.float32), threshold)[0])


def classify_tile(fat_percentage: float, classification_threshold: float) -> str:
    """tile_classification_evaluation.classify_tile (:228-239)."""
    return "Has Fat" if fat_percentage >= classification_threshold else "No Fat"


def write_tiff_lzw(path, array: np.ndarray, threads: int = 0):
    """tifffile.imwrite(path, array, compression='lzw') for uint8 (H,W) / (H,W,3) arrays: baseline TIFF, strips compressed in
    parallel by libadipose_b200's host-side writer (no GPU needed)."""
    a = np.ascontiguousarray(array)
    if a.dtype != np.uint8 or a.ndim not in (2, 3) or (a.ndim == 3 and a.shape[2] != 3):
        raise ValueError("write_tiff_lzw: uint8 (H,W) or (H,W,3) arrays only")
    _lib.check(_lib.load().adp_tiff_write_lzw(str(path).encode(), _lib.ptr(a), a.shape[0], a.shape[1], 1 if a.ndim == 2 else 3, 0, threads))
