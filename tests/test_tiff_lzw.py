"""Host-side TIFF-LZW writer of the library (adp_tiff_write_lzw, csrc/tiff_lzw.h): replaces tifffile.imwrite(...,
compression='lzw') of reconstruct_full_images.py:724-734 / segmentation_inference.py:455-464.  Decoded by libtiff (OpenCV,
PIL) the files must give back the exact bytes; no GPU involved."""
import os

import cv2
import numpy as np
import pytest
from PIL import Image

from adipose_unet_b200 import api


def _cases():
    rng = np.random.default_rng(0)
    smooth = cv2.GaussianBlur(rng.random((700, 900)).astype(np.float32), (0, 0), 6)
    smooth = ((smooth - smooth.min()) / (smooth.max() - smooth.min()) * 255).astype(np.uint8)
    yield "smooth", smooth
    yield "mask01", (smooth > 128).astype(np.uint8)                       # masks/{stem}_mask.tif: {0,1}
    yield "noise", rng.integers(0, 256, size=(513, 257), dtype=np.uint8)  # incompressible: many table clears
    yield "constant", np.full((2048, 2048), 7, np.uint8)                  # long runs: 12-bit codes, width schedule
    yield "rgb", np.stack([smooth, 255 - smooth, (smooth // 2)], axis=-1)
    yield "one_pixel", np.array([[5]], np.uint8)
    yield "odd", rng.integers(0, 4, size=(17, 5), dtype=np.uint8)


@pytest.mark.parametrize("name,arr", list(_cases()))
def test_round_trip_through_libtiff(tmp_path, name, arr):
    path = tmp_path / f"{name}.tif"
    api.write_tiff_lzw(path, arr)
    got = cv2.imread(str(path), cv2.IMREAD_UNCHANGED)
    if arr.ndim == 3:
        got = got[..., ::-1]            # OpenCV hands back BGR; the file holds the samples in the caller's order
    np.testing.assert_array_equal(got.reshape(arr.shape), arr)
    np.testing.assert_array_equal(np.array(Image.open(path)).reshape(arr.shape), arr)
    with Image.open(path) as im:
        assert im.info.get("compression") == "tiff_lzw"
    if name in ("constant", "mask01"):
        assert os.path.getsize(path) < arr.size // 10


def test_single_strip_and_thread_counts_give_identical_pixels(tmp_path):
    from adipose_unet_b200 import _lib
    rng = np.random.default_rng(1)
    a = (rng.random((1500, 1100)) * 40).astype(np.uint8)
    lib = _lib.load()
    for rps, th in ((1500, 1), (64, 1), (64, 4), (1, 3), (0, 0)):
        p = tmp_path / f"r{rps}_t{th}.tif"
        assert lib.adp_tiff_write_lzw(str(p).encode(), _lib.ptr(a), a.shape[0], a.shape[1], 1, rps, th) == 0
        np.testing.assert_array_equal(cv2.imread(str(p), cv2.IMREAD_UNCHANGED), a)


def test_bad_arguments_are_errors(tmp_path):
    with pytest.raises(ValueError):
        api.write_tiff_lzw(tmp_path / "x.tif", np.zeros((4, 4), np.float32))
    with pytest.raises(Exception):
        api.write_tiff_lzw(tmp_path / "no_such_dir" / "x.tif", np.zeros((4, 4), np.uint8))
