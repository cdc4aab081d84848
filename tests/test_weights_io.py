"""Weight files (SURVEY.md section 8b "Weights file"): legacy Keras HDF5 layout round trip with the in-tree
minimal HDF5 writer/reader, tolerant read of the Keras-2.13 'vars' layout, chunked/deflate datasets."""
import os

import numpy as np
import pytest

import adipose_unet_b200 as A
from adipose_unet_b200 import hdf5_min as H
from adipose_unet_b200 import weights_io as W
from adipose_unet_b200.layers import LAYER_NAMES


@pytest.fixture(scope="module")
def weights():
    return A.synth.init_weights(seed=5)


def test_legacy_round_trip_is_exact(tmp_path, weights):
    p = str(tmp_path / "weights_best_overall.weights.h5")
    W.save_weights_file(p, weights)
    back = W.load_weights_file(p)
    assert set(back) == set(weights)
    for k in weights:
        assert back[k].dtype == np.float32 and back[k].shape == weights[k].shape
        assert back[k].tobytes() == weights[k].tobytes()          # exact bytes of the tensors
    r = H.Hdf5Reader(p)
    root = r.attrs["/"]
    assert [n.decode() for n in root["layer_names"]] == LAYER_NAMES
    assert root["backend"] == b"tensorflow" and root["keras_version"] == b"2.13.1"
    assert [n.decode() for n in r.attrs["/dilate3"]["weight_names"]] == ["dilate3/kernel:0", "dilate3/bias:0"]
    assert "/up2_conv2/up2_conv2/kernel:0" in r.datasets
    with open(p, "rb") as f:
        assert f.read(8) == b"\x89HDF\r\n\x1a\n"


def test_legacy_under_model_weights_group(tmp_path, weights):
    """A full-model .h5 keeps the same tree under /model_weights (hdf5_format.load_weights_from_hdf5_group)."""
    w = H.Hdf5Writer()
    for n in LAYER_NAMES:
        w.create_dataset(f"/model_weights/{n}/{n}/kernel:0", weights[n + "/kernel"])
        w.create_dataset(f"/model_weights/{n}/{n}/bias:0", weights[n + "/bias"])
    p = str(tmp_path / "full_model.h5")
    w.save(p)
    back = W.load_weights_file(p)
    for k in weights:
        assert np.array_equal(back[k], weights[k])


@pytest.mark.parametrize("container", ["layers", "_layer_checkpoint_dependencies"])
def test_v3_vars_layout_is_matched_by_creation_order(tmp_path, weights, container):
    p = str(tmp_path / "phase2_best.weights.h5")
    H.write_keras_v3_like_weights(p, weights, LAYER_NAMES, container=container)
    back = W.load_weights_file(p)
    for k in weights:
        assert np.array_equal(back[k], weights[k]), k


def test_v3_layout_with_swapped_identical_shapes_is_detected(tmp_path, weights):
    order = list(LAYER_NAMES)
    i, j = order.index("down3_conv1"), order.index("dilate1")      # different shapes: order violation must be caught
    order[i], order[j] = order[j], order[i]
    p = str(tmp_path / "bad.weights.h5")
    H.write_keras_v3_like_weights(p, weights, order)
    with pytest.raises(H.Hdf5Error):
        W.load_weights_file(p)


def test_shape_validation(tmp_path, weights):
    bad = dict(weights)
    bad["dilate2/kernel"] = bad["dilate2/kernel"][:, :, :-1]
    p = str(tmp_path / "bad_shape.weights.h5")
    H.write_keras_legacy_weights(p, bad, LAYER_NAMES)
    with pytest.raises(ValueError):
        W.load_weights_file(p)


@pytest.mark.parametrize("deflate", [False, True])
def test_chunked_dataset_read(tmp_path, deflate):
    rng = np.random.default_rng(0)
    a = rng.standard_normal((3, 3, 20, 44)).astype(np.float32)
    b = rng.integers(-5, 5, size=(10, 7)).astype(np.int32)
    w = H.Hdf5Writer()
    w.create_dataset("/g/a", a, chunks=(2, 3, 8, 16), deflate=deflate)
    w.create_dataset("/g/b", b, chunks=(4, 4), deflate=deflate)
    w.create_dataset("/scalarish", np.array([1.5], np.float64))
    p = str(tmp_path / "chunked.h5")
    w.save(p)
    r = H.Hdf5Reader(p)
    assert np.array_equal(r.datasets["/g/a"], a)
    assert np.array_equal(r.datasets["/g/b"], b)
    assert r.datasets["/scalarish"][0] == 1.5


def test_many_children_spill_over_several_symbol_nodes(tmp_path):
    w = H.Hdf5Writer()
    names = [f"layer_{i:03d}" for i in range(100)]
    for i, n in enumerate(names):
        w.create_dataset(f"/{n}/v", np.full((2,), i, np.float32))
    p = str(tmp_path / "many.h5")
    w.save(p)
    r = H.Hdf5Reader(p)
    assert sorted(r.datasets) == [f"/{n}/v" for n in names]
    assert all(r.datasets[f"/{n}/v"][0] == i for i, n in enumerate(names))


def test_npz_and_missing_file(tmp_path, weights):
    p = str(tmp_path / "w.npz")
    W.save_weights_file(p, weights)
    assert np.array_equal(W.load_weights_file(p)["dilate6/bias"], weights["dilate6/bias"])
    with pytest.raises(FileNotFoundError):
        W.load_weights_file(str(tmp_path / "absent.weights.h5"))
    junk = tmp_path / "junk.weights.h5"
    junk.write_bytes(b"not hdf5" * 100)
    with pytest.raises(H.Hdf5Error):
        W.load_weights_file(str(junk))


def test_hybrid_weights_file_serves_every_loader_of_the_reference(tmp_path):
    """`*.weights.h5` written by this repo: legacy by-name datasets (load_weights_from_hdf5_group[_by_name],
    full_evaluation_enhanced.py:1285-1301) AND the Keras-2.13 saving_lib groups (`layers/` and the 2.12 container
    `_layer_checkpoint_dependencies/`, `<class>[_k]/vars/{0,1}` in model.layers order) as hard links to the same tensors."""
    import os
    import adipose_unet_b200 as A
    from adipose_unet_b200 import hdf5_min as H
    from adipose_unet_b200.weights_io import load_weights_file, save_weights_file
    for ds in (False, True):
        w = A.synth.init_weights(deep_supervision=ds)
        path = str(tmp_path / f"ds{int(ds)}.weights.h5")
        save_weights_file(path, w)
        assert os.path.getsize(path) < 1.02 * sum(v.nbytes for v in w.values()) + 200_000       # linked, not copied
        r = H.Hdf5Reader(path)
        assert [n.decode() for n in r.attrs["/"]["layer_names"]][:2] == ["down1_conv1", "down1_conv2"]
        assert "/vars" in r.groups and "/layers/reshape/vars" in r.groups and "/layers/lambda_1/vars" in r.groups
        for container in ("/layers", "/_layer_checkpoint_dependencies"):
            np.testing.assert_array_equal(r.datasets[f"{container}/conv2d/vars/0"], w["down1_conv1/kernel"])
            np.testing.assert_array_equal(r.datasets[f"{container}/conv2d_6/vars/0"], w["dilate1/kernel"])
            np.testing.assert_array_equal(r.datasets[f"{container}/conv2d_12/vars/1"], w["up3_conv1/bias"])
            np.testing.assert_array_equal(r.datasets[f"{container}/conv2d_21/vars/0"], w["output_softmax/kernel"])
            if ds:
                np.testing.assert_array_equal(r.datasets[f"{container}/conv2d_22/vars/0"], w["aux_out1/kernel"])
                np.testing.assert_array_equal(r.datasets[f"{container}/conv2d_23/vars/1"], w["aux_out2/bias"])
        for layout in ("auto", "vars"):
            got = H.read_keras_weights(path, layout=layout)
            assert set(got) == set(w) and all(np.array_equal(got[k], w[k]) for k in w)
        assert set(load_weights_file(path)) == {k for k in w if not k.startswith("aux_out")}
        assert set(load_weights_file(path, keep_aux=True)) == set(w)
    # plain `.h5` names keep the legacy layout only
    save_weights_file(str(tmp_path / "legacy.h5"), A.synth.init_weights())
    assert not any(g.startswith("/layers") for g in H.Hdf5Reader(str(tmp_path / "legacy.h5")).groups)
