"""Host-side logic of the data-parallel training step on CPU over a world_size-2
gloo group.  The device engine is replaced by a stand-in with the same phase methods whose 'network' is a
per-pixel logistic model p = sigmoid(a*x + b) differentiated through the oracle's loss statements
(oracle/unet.py: combined_loss_standard) - what is under test is the trainer's exchange arithmetic:
global-Dice mode must reproduce the single-rank step on the concatenated batch, replica mode must equal the mean
of the per-replica gradients."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from adipose_unet_b200 import train as T
from oracle import unet as U

S, B = 16, 2


class FakeTrainEngine:
    """adp_train_* semantics for theta = (a, b); float64 autograd through the oracle's loss."""

    device = 0

    def __init__(self):
        self.theta = np.array([0.7, -0.2], np.float32)
        self.m = np.zeros(2, np.float32); self.v = np.zeros(2, np.float32); self.t = 0
        self.grad = np.zeros(2, np.float32)

    def train_begin(self, batch, size, dropout_rate, seed):
        self.seed = seed

    def train_grad_buffer(self):
        return 0, 2

    def stream_ptr(self):
        return 0

    def train_forward(self, x, y, masks=None):
        self.x = torch.tensor(x, dtype=torch.float64); self.y = torch.tensor(y, dtype=torch.float64)
        p = torch.sigmoid(float(self.theta[0]) * self.x + float(self.theta[1]))
        pc = torch.clamp(p, U.EPS, 1 - U.EPS)
        bce = -(self.y * torch.log(pc + U.EPS) + (1 - self.y) * torch.log(1 - pc + U.EPS))
        return np.array([bce.sum(), (self.y * pc).sum(), self.y.sum(), pc.sum(), (self.y * p).sum(), p.sum(), self.y.sum(),
                         float(self.y.numel())], np.float64)

    @staticmethod
    def train_loss(s):
        n = s[7]
        dl = 1.0 - (2 * s[1] + 1) / (s[2] + s[3] + 1)
        return dict(loss=s[0] / n + dl, bce=s[0] / n, dice_loss=dl, dice_coef=(2 * s[4] + 1) / (s[6] + s[5] + 1))

    def train_backward(self, gs, freeze):
        n_global = gs[7]
        # dL/dp for LOCAL pixels of the loss defined by the (global) sums: bce/n + 1 - (2I+1)/(Y+P+1)
        th = torch.tensor(self.theta.astype(np.float64), requires_grad=True)
        p = torch.sigmoid(th[0] * self.x + th[1])
        pc = torch.clamp(p, U.EPS, 1 - U.EPS)
        bce = -(self.y * torch.log(pc + U.EPS) + (1 - self.y) * torch.log(1 - pc + U.EPS))
        I, D = gs[1], gs[2] + gs[3] + 1.0
        dldpc = -2.0 * self.y / D + (2 * I + 1) / (D * D)          # d(dice_loss)/d(pc), global sums held fixed
        surrogate = bce.sum() / n_global + (dldpc.detach() * pc).sum()
        surrogate.backward()
        self.grad = th.grad.numpy().astype(np.float32)

    def grad_to_host(self):
        return self.grad.copy()

    def grad_from_host(self, g):
        self.grad = np.asarray(g, np.float32).copy()

    def train_apply(self, lr, optimizer, grad_scale=1.0, weight_decay=0.01, freeze_encoder=False, **kw):
        self.t += 1
        self.theta, self.m, self.v = U.keras_adam_step(self.theta, self.grad * np.float32(grad_scale), self.m, self.v,
                                                       self.t, lr, weight_decay=weight_decay if optimizer == "adamw" else 0.0)

    def train_end(self):
        pass


def _data(world):
    rng = np.random.default_rng(11)
    x = rng.standard_normal((world * B, S, S)).astype(np.float32)
    y = (rng.random((world * B, S, S)) < 0.3).astype(np.float32)
    return x, y


def _single_rank_reference(world, steps=3, lr=1e-2):
    """One 'GPU' holding the concatenated batch: the oracle's loss differentiated directly."""
    x, y = _data(world)
    theta = np.array([0.7, -0.2], np.float32); m = np.zeros(2, np.float32); v = np.zeros(2, np.float32)
    losses = []
    for t in range(1, steps + 1):
        th = torch.tensor(theta.astype(np.float64), requires_grad=True)
        p = torch.sigmoid(th[0] * torch.tensor(x, dtype=torch.float64) + th[1])
        loss = U.combined_loss_standard(torch.tensor(y, dtype=torch.float64), p)
        loss.backward()
        losses.append(float(loss))
        theta, m, v = U.keras_adam_step(theta, th.grad.numpy().astype(np.float32), m, v, t, lr)
    return theta, losses


def _worker(rank, world, port, mode, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x, y = _data(world)
    tr = T.DataParallelTrainer(FakeTrainEngine(), B, S, dist=dist, rank=rank, world=world, dropout_rate=0.0,
                               dice_mode=mode)
    losses = [tr.step(x[rank * B:(rank + 1) * B], y[rank * B:(rank + 1) * B], 1e-2)["loss"] for _ in range(3)]
    q.put((rank, tr.engine.theta.copy(), losses))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _spawn(world, mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def test_global_dice_dp_equals_single_rank_on_concatenated_batch():
    world = 2
    theta_ref, losses_ref = _single_rank_reference(world)
    res = _spawn(world, "global")
    for rank, theta, losses in res:
        np.testing.assert_allclose(theta, theta_ref, rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(losses, losses_ref, rtol=1e-9)
    np.testing.assert_array_equal(res[0][1], res[1][1])            # replicas stay bit-identical


def test_replica_mode_is_mean_of_replica_gradients():
    world = 2
    res = _spawn(world, "replica")
    np.testing.assert_array_equal(res[0][1], res[1][1])
    # expected: per-replica loss gradients, averaged
    x, y = _data(world)
    theta = np.array([0.7, -0.2], np.float32); m = np.zeros(2, np.float32); v = np.zeros(2, np.float32)
    for t in range(1, 4):
        gs = []
        for r in range(world):
            th = torch.tensor(theta.astype(np.float64), requires_grad=True)
            p = torch.sigmoid(th[0] * torch.tensor(x[r * B:(r + 1) * B], dtype=torch.float64) + th[1])
            U.combined_loss_standard(torch.tensor(y[r * B:(r + 1) * B], dtype=torch.float64), p).backward()
            gs.append(th.grad.numpy().astype(np.float32))
        theta, m, v = U.keras_adam_step(theta, (gs[0] + gs[1]) * np.float32(0.5), m, v, t, 1e-2)
    np.testing.assert_allclose(res[0][1], theta, rtol=1e-5, atol=1e-7)


def test_emulated_step_matches_gloo_step():
    world = 2
    x, y = _data(world)
    trs = [T.DataParallelTrainer(FakeTrainEngine(), B, S, rank=r, world=world, dropout_rate=0.0) for r in range(world)]
    for _ in range(3):
        T.emulated_step(trs, [x[:B], x[B:]], [y[:B], y[B:]], 1e-2)
    theta_ref, _ = _single_rank_reference(world)
    np.testing.assert_allclose(trs[0].engine.theta, theta_ref, rtol=1e-5, atol=1e-7)


def test_cosine_warmup_schedule_matches_reference_expression():
    # CosineAnnealingWithWarmup, phase-1 constants (train_adipose_unet_v3.py:1294-1300)
    mx, mn, wu, tot = 1e-4, 1e-7, 5, 75
    for e in range(tot):
        want = (mx / wu) * (e + 1) if e < wu else mn + 0.5 * (mx - mn) * (1 + np.cos(np.pi * ((e - wu) / (tot - wu))))
        assert T.cosine_warmup_lr(e, mx, mn, wu, tot) == want
    assert T.cosine_warmup_lr(0, mx, mn, wu, tot) == pytest.approx(2e-5)
    assert T.cosine_warmup_lr(wu, mx, mn, wu, tot) == pytest.approx(mx)
