"""Shared by tests/test_oracle_refine.py and tests/test_gpu_post.py: synthetic probability maps and the reference's
BoundaryRefiner.refine statement typed with cv2 itself."""
import cv2
import numpy as np


def _blob_prob(h, w, seed):
    rng = np.random.default_rng(seed)
    b = cv2.GaussianBlur(rng.random((h, w)).astype(np.float32), (0, 0), 6)
    return ((b - b.min()) / (b.max() - b.min())).astype(np.float32)


def reference_refine_cv2(mask, k=5, d=5, sc=50.0, ss=50.0):
    """full_evaluation_enhanced.py:357-393 statement by statement, with cv2 itself."""
    kernel = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (k, k))
    mask_u8 = (mask * 255).astype(np.uint8)
    eroded = cv2.erode(mask_u8, kernel, iterations=1)
    dilated = cv2.dilate(mask_u8, kernel, iterations=1)
    boundary = np.logical_xor(dilated > 0, eroded > 0).astype(np.uint8)
    filtered = cv2.bilateralFilter(mask_u8, d, sc, ss)
    refined = np.where(boundary > 0, filtered, mask_u8)
    refined = cv2.morphologyEx(refined, cv2.MORPH_OPEN, kernel, iterations=1)
    refined = cv2.morphologyEx(refined, cv2.MORPH_CLOSE, kernel, iterations=1)
    return (refined / 255.0).astype(np.float32)
