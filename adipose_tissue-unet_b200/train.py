"""Data-parallel training step and the reference's learning-rate schedule.

One process per GPU.  A step is the engine's forward -> backward -> apply (include/adipose_b200.h, adp_train_*)
with two exchanges between ranks:

* eight float64 loss sums before backward (``dice_mode="global"``): the reference's Dice term is defined over the
  WHOLE batch (train_adipose_unet_v3.py:217-225), so the exact data-parallel equivalent of a single-GPU step on the
  concatenated batch needs sum(y*p), sum(y), sum(p) over all ranks before dL/dp is formed.  ``dice_mode="replica"``
  skips this exchange and averages per-replica gradients instead (what wrapping the reference in a DP strategy does).
  With NCCL the sums never leave the device: the all-reduce runs in place on the engine's sums buffer and stream;
* the all-reduce (sum) of the flat fp32 gradient (8.5 M parameters, 34 MB) before the optimizer, done by
  ``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU tests) directly on the engine's device
  buffer - PyTorch is only the collective plumbing here.  With NCCL it is split into the engine's four completion
  buckets (decoder, dilate6..4, dilate3..1, encoder) and each bucket's all-reduce is enqueued on a side stream behind
  the event the backward pass records when that bucket's last weight gradient is written, so only the encoder's
  2.3 MB are exchanged after the last backward kernel.  A step has no host synchronisation; the loss of a step is
  read from a pinned mirror of the sums that lands at the start of its backward pass.

Nothing in this file computes on the CPU: without the CUDA library ``api.Engine`` raises.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np


class _CudaView:
    """__cuda_array_interface__ over a raw device pointer so torch can alias the engine's gradient buffer."""

    def __init__(self, ptr: int, count: int, typestr: str = "<f4"):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False), "version": 3,
                                         "strides": None}


class DataParallelTrainer:
    """Drives one engine per rank.  ``dist`` is ``torch.distributed`` (initialised) or None for a single rank.

    step(x, y, lr) == Keras ``train_step`` on the global batch (train_adipose_unet_v3.py:1316-1324, 1413-1421):
    returns {loss, bce, dice_loss, dice_coef} of the GLOBAL batch in dice_mode="global".  The three phases
    (forward_sums / backward / apply) are public so a test can interleave several in-process 'ranks'."""

    def __init__(self, engine, batch: int, size: int, *, dist=None, rank: int = 0, world: int = 1,
                 dropout_rate: float = 0.3, seed: int = 865, optimizer: str = "adam", weight_decay: float = 0.01,
                 dice_mode: str = "global", freeze_encoder: bool = False):
        assert dice_mode in ("global", "replica")
        self.engine, self.dist, self.rank, self.world = engine, dist, rank, world
        self.batch, self.size = batch, size
        self.optimizer, self.weight_decay = optimizer, weight_decay
        self.dice_mode, self.freeze_encoder = dice_mode, freeze_encoder
        # per-rank dropout stream: the same seed would drop the same units on every replica
        engine.train_begin(batch, size, dropout_rate, seed + 7919 * rank)
        self._grad_t = None
        self._sums_t = None
        self._torch_stream = None
        self._side_stream = None
        self._buckets = []
        self.allreduce_bytes = 4 * engine.train_grad_buffer()[1]
        self.collective_path = "none"
        if dist is not None and world > 1:
            self._bind_collective()

    # -- plumbing ---------------------------------------------------------------------------------------------
    def _bind_collective(self):
        import torch
        if "nccl" in str(self.dist.get_backend()):      # "nccl" or a mixed "cpu:gloo,cuda:nccl" group
            ptr, n = self.engine.train_grad_buffer()
            dev = getattr(self.engine, "device", 0)
            self._grad_t = torch.as_tensor(_CudaView(ptr, n), device=f"cuda:{dev}")
            sp = self.engine.stream_ptr()
            self._torch_stream = torch.cuda.ExternalStream(sp, device=dev) if sp else None
            self.collective_path = "nccl on the engine's device buffer"
            if self._torch_stream is not None and hasattr(self.engine, "train_sums_buffer"):
                sptr, sn = self.engine.train_sums_buffer()
                self._sums_t = torch.as_tensor(_CudaView(sptr, sn, "<f8"), device=f"cuda:{dev}")
                self._side_stream = torch.cuda.Stream(device=dev)
                self._buckets = [b for b in self.engine.train_grad_buckets() if b[1] > b[0]]
                self.collective_path = ("nccl on the engine's device buffers: loss sums in place on the engine stream, gradient in "
                                        "%d buckets on a side stream behind the backward pass's completion events" % len(self._buckets))
        else:
            self.collective_path = "host-staged (%s)" % self.dist.get_backend()

    def _allreduce_sums(self, sums: np.ndarray) -> np.ndarray:
        if self.dist is None or self.world == 1:
            return sums
        import torch
        t = torch.tensor(np.asarray(sums, dtype=np.float64), dtype=torch.float64,
                         device=self._grad_t.device if self._grad_t is not None else "cpu")
        self.dist.all_reduce(t)
        return t.cpu().numpy()

    def _allreduce_grad(self):
        if self.dist is None or self.world == 1:
            return
        import torch
        if self._grad_t is not None:
            # enqueued behind the backward kernels on the engine's own stream, the optimizer kernel behind it:
            # no host synchronisation inside the step
            if self._torch_stream is not None:
                with torch.cuda.stream(self._torch_stream):
                    self.dist.all_reduce(self._grad_t)
            else:                    # engine runs on the legacy default stream
                self.engine.synchronize()
                self.dist.all_reduce(self._grad_t)
                torch.cuda.synchronize(self._grad_t.device)
            return
        g = torch.from_numpy(self.engine.grad_to_host())     # host-staged collective (gloo)
        self.dist.all_reduce(g)
        self.engine.grad_from_host(g.numpy())

    # -- phases -----------------------------------------------------------------------------------------------
    @property
    def n_local(self) -> int:
        return self.batch * self.size * self.size

    def forward_sums(self, x, y, dropout_masks=None) -> np.ndarray:
        return self.engine.train_forward(x, y, dropout_masks)

    def backward(self, sums: np.ndarray, global_sums: Optional[np.ndarray] = None) -> Dict[str, float]:
        """Backward for the loss the mode defines; returns that loss.  The sums carry their own term count
        (sums[7]), so hard-example mining and label smoothing need no special casing here."""
        gs = sums if (self.dice_mode != "global" or global_sums is None) else global_sums
        self.engine.train_backward(gs, self.freeze_encoder)
        return self.engine.train_loss(gs)

    def apply(self, lr: float):
        # global: d(global loss)/d(theta) is the SUM of the ranks' contributions; replica: mean of replica gradients
        scale = 1.0 if self.dice_mode == "global" else 1.0 / self.world
        self.engine.train_apply(lr, self.optimizer, grad_scale=scale, weight_decay=self.weight_decay,
                                freeze_encoder=self.freeze_encoder)

    def _device_resident(self) -> bool:
        """True when a step can stay on the device: one rank, or an NCCL group bound to the engine's buffers."""
        if not hasattr(self.engine, "train_sums_read"):
            return False             # stand-in engines of the CPU tests
        return self.dist is None or self.world == 1 or self._sums_t is not None

    def step(self, x, y, lr: float, dropout_masks=None, want_loss: bool = True) -> Optional[Dict[str, float]]:
        if not self._device_resident():      # host-staged exchanges (gloo)
            sums = self.forward_sums(x, y, dropout_masks)
            gsums = self._allreduce_sums(sums) if self.dice_mode == "global" else None
            out = self.backward(sums, gsums)
            self._allreduce_grad()
            self.apply(lr)
            return out
        import torch
        eng = self.engine
        eng.train_forward(x, y, dropout_masks, want_sums=False)
        multi = self.dist is not None and self.world > 1
        if multi and self.dice_mode == "global":
            with torch.cuda.stream(self._torch_stream):
                self.dist.all_reduce(self._sums_t)
        eng.train_backward(None, self.freeze_encoder)          # enqueued; records the bucket events
        if multi:
            side = self._side_stream
            for b, (lo, hi) in enumerate(eng.train_grad_buckets()):
                if hi <= lo:
                    continue
                eng.train_bucket_wait(b, side.cuda_stream)
                with torch.cuda.stream(side):
                    self.dist.all_reduce(self._grad_t[lo:hi])
            eng.train_join(side.cuda_stream)
        self.apply(lr)
        if not want_loss:
            return None
        return eng.train_loss(eng.train_sums_read())

    def close(self):
        self._grad_t = None
        self.engine.train_end()


def emulated_step(trainers, xs, ys, lr: float, dropout_masks=None):
    """The same data-parallel step with every 'rank' living in this process (several engines on one device):
    the two exchanges are done on the host.  Test vehicle for the N>1 arithmetic on a single-GPU box."""
    sums = [t.forward_sums(x, y, None if dropout_masks is None else dropout_masks[i])
            for i, (t, x, y) in enumerate(zip(trainers, xs, ys))]
    gs = np.sum(np.stack(sums), axis=0)
    outs = [t.backward(s, gs) for t, s in zip(trainers, sums)]
    total = np.sum(np.stack([t.engine.grad_to_host().astype(np.float64) for t in trainers]), axis=0).astype(np.float32)
    for t in trainers:
        t.engine.grad_from_host(total)
        t.apply(lr)
    return outs


# ---- learning-rate schedule -------------------------------------------------------------------------------------
def cosine_warmup_lr(epoch: int, max_lr: float, min_lr: float, warmup_epochs: int, total_epochs: int) -> float:
    """CosineAnnealingWithWarmup.on_epoch_begin (train_adipose_unet_v3.py:393-404), same float64 expression order:
    linear warm-up (max_lr/warmup)*(epoch+1), then min_lr + 0.5*(max_lr-min_lr)*(1+cos(pi*progress))."""
    if epoch < warmup_epochs:
        return (max_lr / warmup_epochs) * (epoch + 1)
    progress = (epoch - warmup_epochs) / (total_epochs - warmup_epochs)
    return float(min_lr + 0.5 * (max_lr - min_lr) * (1 + np.cos(np.pi * progress)))
