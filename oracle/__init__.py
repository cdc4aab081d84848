"""ORACLE — test infrastructure only.

CPU restatement of the reference's algorithm for the U-Net segmentation hot path.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package, and only as the checker; the product path
(adipose_unet_b200) never imports it and fails loudly without its CUDA library.

  geometry.py  NumPy: tile positions, windows, blenders, D4 TTA, threshold, metrics
               — PINNED against the reference's own NumPy code (tests/golden/).
  unet.py      PyTorch-CPU: graph, predict_single/TTA, BCE+Dice, backward, Keras Adam
               — PARITY UNPINNED (TensorFlow 2.13 not installable here; no vectors upstream).
"""
