"""Data-parallel training step on the GPU:
(1) two in-process 'ranks' (two engines on cuda:0, exchanges done on the host) against ONE engine stepping on the
    concatenated batch - the exchange arithmetic with the real kernels, runnable on a one-GPU box;
(2) the same over real NCCL with one process per GPU when the box has >= 2 GPUs (tools/dp_train_check.py)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import adipose_unet_b200 as A
from adipose_unet_b200 import api, train as T


def batch(n, S, seed=5):
    tiles = np.stack([A.synth.ecm_tile(S, seed=seed + 31 * i) for i in range(n)])
    x = ((tiles.astype(np.float32) - A.synth.DEFAULT_MEAN) / (A.synth.DEFAULT_STD + 1e-10)).astype(np.float32)
    y = np.stack([A.synth.mask_from_tile(t, 140) for t in tiles]).astype(np.float32)
    return x, y

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _moved(w, w0, keys):
    return np.concatenate([(w[k] - w0[k]).ravel() for k in keys]).astype(np.float64)


@pytest.mark.parametrize("prec,mode", [("fp32", "global"), ("bf16", "global"), ("fp32", "replica")])
def test_two_emulated_ranks_equal_one_engine_on_concatenated_batch(prec, mode):
    weights = A.synth.init_weights()
    b, S, lr, steps = 1, 128, 1e-4, 3
    x, y = batch(2 * b, S, seed=17)
    engs = [api.Engine(precision=prec, max_forwards=4) for _ in range(2)]
    for e in engs:
        e.set_weights(weights)
    trs = [T.DataParallelTrainer(e, b, S, rank=r, world=2, dropout_rate=0.0, dice_mode=mode) for r, e in enumerate(engs)]
    losses = []
    for _ in range(steps):
        outs = T.emulated_step(trs, [x[:b], x[b:]], [y[:b], y[b:]], lr)
        losses.append(outs[0]["loss"])
    w_dp = [e.get_weights() for e in engs]
    for t in trs:
        t.close()
    keys = sorted(weights)
    for k in keys:                                                   # replicas stay bit-identical
        assert np.array_equal(w_dp[0][k], w_dp[1][k]), k
    if mode == "replica":
        # per-replica Dice: a different (documented) loss than the whole-batch one; only sanity here
        assert np.isfinite(losses).all() and np.abs(_moved(w_dp[0], weights, keys)).max() > 0.5 * lr
        for e in engs:
            e.close()
        return
    one = api.Engine(precision=prec, max_forwards=4)
    one.set_weights(weights)
    one.train_begin(2 * b, S, dropout_rate=0.0)
    losses1 = [one.train_step(x, y, lr)["loss"] for _ in range(steps)]
    w1 = one.get_weights()
    one.train_end(); one.close()
    for e in engs:
        e.close()
    # Appendix B bound (1e-5) is for ONE step from identical parameters; later steps start from parameters that differ
    # by Adam's amplification of summation-order noise (measured 1.1e-5 at step 3), asserted 2e-4
    np.testing.assert_allclose(losses[0], losses1[0], rtol=1e-5 if prec == "fp32" else 2e-2)
    np.testing.assert_allclose(losses, losses1, rtol=2e-4 if prec == "fp32" else 2e-2)
    d = _moved(w_dp[0], w1, keys); mv = _moved(w1, weights, keys)
    rms_ratio = float(np.sqrt((d ** 2).mean()) / np.sqrt((mv ** 2).mean()))
    print(prec, "DP(2, emulated) vs single: loss", losses, losses1, "theta rms ratio", rms_ratio)
    # summation order of the weight gradient differs (per-rank partial sums): Adam turns noise-floor gradients into
    # +-lr steps, so the bound is statistical as in test_three_steps_fp32_vs_oracle
    assert rms_ratio <= (0.05 if prec == "fp32" else 0.5)


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_nccl_data_parallel_equals_single_gpu(prec):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (covered on one GPU by the emulated-rank test and on CPU by the gloo test)")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tools", "dp_train_check.py"), "--precision", prec]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    out = json.loads(lines[-1])
    print(out)
    assert out["ok"] and out["replicas_identical"]
