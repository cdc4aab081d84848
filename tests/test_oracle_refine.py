"""oracle/refine.py (BoundaryRefiner restatement) against the cv2 of this image: structuring element and morphology
bit-exact; 8-bit bilateral within one grey level (not bit-stable across OpenCV builds, see oracle/refine.py); the
reference's refine() statement, re-typed here with cv2 calls, against the restatement."""
import cv2
import numpy as np
import pytest

from oracle import refine as R


from refine_helpers import _blob_prob, reference_refine_cv2      # tests/ is on sys.path (conftest.py lives there)


@pytest.mark.parametrize("k", [1, 3, 5, 7, 9, 11, 15])
def test_ellipse_kernel_matches_cv2(k):
    np.testing.assert_array_equal(R.ellipse_kernel(k), cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (k, k)))


@pytest.mark.parametrize("k", [3, 5, 9])
def test_morphology_bit_exact_vs_cv2(k):
    rng = np.random.default_rng(k)
    kernel = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (k, k))
    for img in (rng.integers(0, 256, (67, 93), dtype=np.uint8), ((_blob_prob(128, 160, k) > 0.5) * 255).astype(np.uint8)):
        np.testing.assert_array_equal(R.erode(img, kernel), cv2.erode(img, kernel))
        np.testing.assert_array_equal(R.dilate(img, kernel), cv2.dilate(img, kernel))
        np.testing.assert_array_equal(R.dilate(R.erode(img, kernel), kernel), cv2.morphologyEx(img, cv2.MORPH_OPEN, kernel))
        np.testing.assert_array_equal(R.erode(R.dilate(img, kernel), kernel), cv2.morphologyEx(img, cv2.MORPH_CLOSE, kernel))


def test_bilateral_within_one_grey_level_of_cv2():
    rng = np.random.default_rng(3)
    for img in (rng.integers(0, 256, (65, 77), dtype=np.uint8), (_blob_prob(128, 128, 1) * 255).astype(np.uint8)):
        d = R.bilateral_u8(img).astype(np.int32) - cv2.bilateralFilter(img, 5, 50, 50).astype(np.int32)
        assert np.abs(d).max() <= 1


@pytest.mark.parametrize("binary", [False, True])
def test_refine_vs_reference_statement(binary):
    p = _blob_prob(256, 320, 7)
    if binary:                                  # evaluation passes thresholded masks too
        p = (p > 0.5).astype(np.float32)
    ours, ref = R.refine(p), reference_refine_cv2(p)
    assert np.abs(ours - ref).max() <= 1.0 / 255.0 + 1e-7                 # one grey level (bilateral rounding of the cv2 build)
    a, b = ours > 0.5, ref > 0.5
    assert (a != b).mean() <= 1e-4
