"""Weight files of the reference (`*.weights.h5`, SURVEY.md section 8b "Weights file").

load_weights_file(path) -> {'<layer>/kernel': HWIO float32, '<layer>/bias': float32}
save_weights_file(path, weights): `*.weights.h5` names get the hybrid file of hdf5_min.write_keras_weights_hybrid (Keras-2.13
saving_lib groups hard-linked to legacy by-name datasets: loadable by net.load_weights AND by the reference's legacy
fallback, full_evaluation_enhanced.py:1266-1301); other `.h5` names the plain legacy layout.  `.npz` is accepted for tests.
"""
from __future__ import annotations

import numpy as np

from .layers import AUX_NAMES, LAYER_NAMES, aux_weight_shapes, weight_shapes


def _validate(w: dict, init_nb: int, keep_aux: bool = False) -> dict:
    """The 22 graph layers (required) and, with keep_aux, the deep-supervision heads when the file has them."""
    shapes = dict(weight_shapes(init_nb))
    names = list(LAYER_NAMES)
    if keep_aux and all(n + "/kernel" in w for n in AUX_NAMES):
        shapes.update(aux_weight_shapes(init_nb))
        names += list(AUX_NAMES)
    out = {}
    for name in names:
        ks, bs = shapes[name]
        k = np.asarray(w[name + "/kernel"], dtype=np.float32)
        b = np.asarray(w[name + "/bias"], dtype=np.float32)
        if k.shape != ks or b.shape != bs:
            raise ValueError(f"{name}: expected kernel {ks} bias {bs}, file has {k.shape} {b.shape}")
        out[name + "/kernel"], out[name + "/bias"] = k, b
    return out


def load_weights_file(path: str, init_nb: int = 44, keep_aux: bool = False) -> dict:
    """keep_aux: also return aux_out1 / aux_out2 of a deep-supervision checkpoint (training resumes with them;
    inference never needs them, full_evaluation_enhanced.py:1313-1319)."""
    if path.endswith(".npz"):
        with np.load(path) as f:
            return _validate({k: f[k] for k in f.files}, init_nb, keep_aux)
    from .hdf5_min import read_keras_weights
    return _validate(read_keras_weights(path, init_nb), init_nb, keep_aux)


def save_weights_file(path: str, weights: dict, init_nb: int = 44) -> None:
    w = _validate(weights, init_nb, keep_aux=True)
    if path.endswith(".npz"):
        np.savez(path, **w)
        return
    from .hdf5_min import write_keras_legacy_weights, write_keras_weights_hybrid
    if path.endswith(".weights.h5"):
        write_keras_weights_hybrid(path, w)
    else:
        write_keras_legacy_weights(path, w)
