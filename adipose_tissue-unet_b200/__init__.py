"""adipose_unet_b200 — B200-native U-Net segmentation hot path of
MAGIC-SCAN/adipose_tissue-unet behind the reference's own Python seam.

The compute lives in ``csrc/`` (hand-written sm_100a CUDA behind the C ABI
declared in ``include/adipose_b200.h``); this package is the thin ctypes host
layer that mirrors the reference's objects (AdiposeUNet, TestTimeAugmentation,
GaussianBlender, LinearBlender, SlidingWindowInference).  There is no CPU
fallback: importing :mod:`adipose_unet_b200.api` without the built library raises.
"""
from . import layers, synth  # noqa: F401

__all__ = ["layers", "synth"]
