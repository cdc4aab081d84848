"""Weight files of the reference (`*.weights.h5`, SURVEY.md section 8b "Weights file").

load_weights_file(path) -> {'<layer>/kernel': HWIO float32, '<layer>/bias': float32}
save_weights_file(path, weights): legacy Keras HDF5 layout (by-name loadable by the reference's
own fallback, full_evaluation_enhanced.py:1285-1301).  `.npz` is accepted for tests.
"""
from __future__ import annotations

import numpy as np

from .layers import LAYER_NAMES, weight_shapes


def _validate(w: dict, init_nb: int) -> dict:
    shapes = weight_shapes(init_nb)
    out = {}
    for name in LAYER_NAMES:
        ks, bs = shapes[name]
        k = np.asarray(w[name + "/kernel"], dtype=np.float32)
        b = np.asarray(w[name + "/bias"], dtype=np.float32)
        if k.shape != ks or b.shape != bs:
            raise ValueError(f"{name}: expected kernel {ks} bias {bs}, file has {k.shape} {b.shape}")
        out[name + "/kernel"], out[name + "/bias"] = k, b
    return out


def load_weights_file(path: str, init_nb: int = 44) -> dict:
    if path.endswith(".npz"):
        with np.load(path) as f:
            return _validate({k: f[k] for k in f.files}, init_nb)
    from .hdf5_min import read_keras_weights
    return _validate(read_keras_weights(path, init_nb), init_nb)


def save_weights_file(path: str, weights: dict, init_nb: int = 44) -> None:
    w = _validate(weights, init_nb)
    if path.endswith(".npz"):
        np.savez(path, **w)
        return
    from .hdf5_min import write_keras_legacy_weights
    write_keras_legacy_weights(path, w)
