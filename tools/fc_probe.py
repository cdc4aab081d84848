#!/usr/bin/env python
"""Smallest run of the fused first conv (one 128^2 tile, bf16): for compute-sanitizer when the fusion test fails."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adipose_unet_b200 as A
from adipose_unet_b200 import api
S = int(sys.argv[1]) if len(sys.argv) > 1 else 128
eng = api.Engine(precision="bf16", max_forwards=2)
eng.set_weights(A.synth.init_weights())
t = A.synth.ecm_tiles(1, S).astype(np.float32)
eng.set_option("fuse_first", 0); a = eng.predict(t, 127.5, 50.0)
eng.set_option("fuse_first", 1); b = eng.predict(t, 127.5, 50.0)
print("max|fused - unfused| =", float(np.abs(a - b).max()))
