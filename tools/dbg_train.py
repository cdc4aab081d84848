import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import adipose_unet_b200 as A
from adipose_unet_b200 import api
from oracle import unet as U
from test_gpu_train import batch, dropout_masks, rel_err, l2_err, run_engine
w = A.synth.init_weights()
S = int(sys.argv[1]) if len(sys.argv) > 1 else 128
x, y = batch(2, S)
for seed in (3, 4, 5):
    masks = dropout_masks(2, S, seed=seed)
    om = {k: np.ascontiguousarray(v.transpose(0, 3, 1, 2)).astype(np.float32) for k, v in masks.items()}
    r64 = U.loss_and_grads(x, y, w, dtype=torch.float64, dropout_masks=om)
    r32 = U.loss_and_grads(x, y, w, dropout_masks=om)
    eng, loss, prob, g = run_engine("fp32", w, x, y, masks)
    e_mine = {k: rel_err(g[k], r64[4][k]) for k in g}
    e_or = {k: rel_err(r32[4][k], r64[4][k]) for k in g}
    top = sorted(e_mine.items(), key=lambda kv: -kv[1])[:6]
    print("seed", seed, "prob err", np.abs(prob - r64[2]).max(), [(k, f"{v:.2e}", f"oracle32 {e_or[k]:.2e}") for k, v in top])
    eng.train_end(); eng.close()
