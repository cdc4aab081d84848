"""Two-phase fine-tuning - drop-in for Segmentation/train_adipose_unet_v3.py (argparse :1455-1630, driver :1080-1450).

What runs on the device: the whole training step (forward with Dropout(0.3), combined_loss_standard = BCE + Dice,
backward, Keras Adam/AdamW; include/adipose_b200.h adp_train_*), optionally data-parallel over the GPUs of one box
(`torchrun --nproc-per-node N -m adipose_unet_b200.cli.train ...`: NCCL all-reduce of the gradient, train.py).
What stays on the host, as in the reference: tile files, shuffling, augmentation, the epoch loop, the cosine/warm-up
schedule (:393-404), best-checkpoint bookkeeping, EMA of the weights (:407-505) and the log files.

Recipe coverage: deep supervision (--use-deep-supervision, --ds-weight-*: aux_out1 / aux_out2
heads, adp_train_set_deep_supervision), hard-example mining (--use-hard-mining, --hard-example-ratio) and asymmetric
label smoothing (--use-label-smoothing, --label-smooth-epsilon-*; adp_train_set_loss) all run on the device, i.e. the
reference's default recipe.  --no-cosine-schedule runs Keras' ReduceLROnPlateau (factor 0.5, patience 5, :1303-1313) on
the monitored validation Dice.  phase{1,2}_training.log has CSVLogger's columns for the reference's compile() (epoch + the
sorted log keys: loss, dice_coef, binary_accuracy and their val_ twins; main_out_/aux_out*_ names with deep supervision).
Still host-side / not reproduced: the intensity and elastic parts of the augmentation levels (geometric D4 part only)."""
from __future__ import annotations

import argparse
import csv
import json
import math
import os
import sys
import time
from datetime import datetime
from pathlib import Path

import cv2
import numpy as np

from . import common as C

TILE = 1024


def build_parser():
    p = argparse.ArgumentParser(description="Train Adipose U-Net (v3 recipe) on the B200 engine")
    p.add_argument("--data-root", type=str, default=os.path.expanduser("~/Data_for_ML/Meat_Luci_Tulane/_build"))
    p.add_argument("--pretrained-weights", type=str, default="checkpoints/unet_1024_dilation/weights_loss_val.weights.h5")
    p.add_argument("--batch-size", type=int, default=2)
    p.add_argument("--epochs-phase1", type=int, default=75)
    p.add_argument("--epochs-phase2", type=int, default=150)
    p.add_argument("--normalization-method", choices=["zscore", "percentile"], default="percentile")
    p.add_argument("--percentile-low", type=float, default=1.0)
    p.add_argument("--percentile-high", type=float, default=99.0)
    p.add_argument("--augmentation-level", choices=["none", "light", "moderate", "heavy", "tta-style"], default="moderate")
    p.add_argument("--checkpoint-suffix", type=str, default="")
    for name, default in (("deep-supervision", True), ("hard-mining", True), ("label-smoothing", False), ("cosine-schedule", True)):
        dest = "use_" + name.replace("-", "_")
        p.add_argument(f"--use-{name}", dest=dest, action="store_true", default=default)
        p.add_argument(f"--no-{name}", dest=dest, action="store_false", default=default)
    p.add_argument("--hard-example-ratio", type=float, default=0.7)
    p.add_argument("--ema-decay", type=float, default=0.995)
    p.add_argument("--optimizer", choices=["adam", "adamw"], default="adam")
    p.add_argument("--label-smooth-epsilon-pos", type=float, default=0.03)
    p.add_argument("--label-smooth-epsilon-neg", type=float, default=0.07)
    p.add_argument("--warmup-epochs-phase1", type=int, default=5)
    p.add_argument("--warmup-epochs-phase2", type=int, default=3)
    p.add_argument("--ds-weight-main", type=float, default=1.0)
    p.add_argument("--ds-weight-aux1", type=float, default=0.4)
    p.add_argument("--ds-weight-aux2", type=float, default=0.3)
    g = p.add_argument_group("B200 engine (not in the reference)")
    g.add_argument("--precision", choices=["bf16", "fp32"], default="bf16")
    g.add_argument("--checkpoint-root", type=str, default="checkpoints/segmentation")
    g.add_argument("--max-steps-per-epoch", type=int, default=0, help="cap for smoke runs (0 = whole epoch)")
    g.add_argument("--seed", type=int, default=865)
    return p


class TileDataset:
    """Paired `images/*.jpg` + `masks/*.tif` (train_adipose_unet_v3.py:510-600): gray float32 tile, mask as stored."""

    def __init__(self, images_dir: Path, masks_dir: Path, method: str, mean: float, std: float, p_low: float, p_high: float):
        masks = {p.stem: p for p in Path(masks_dir).glob("*.tif")}
        self.pairs = [(p, masks[p.stem]) for p in sorted(Path(images_dir).glob("*.jpg")) if p.stem in masks]
        self.method, self.mean, self.std, self.p_low, self.p_high = method, mean, std, p_low, p_high
        print(f"Found {len(self.pairs)} paired tiles in {Path(images_dir).name}")

    def __len__(self):
        return len(self.pairs)

    def load(self, idx: int):
        ip, mp = self.pairs[idx]
        img = cv2.imread(str(ip), cv2.IMREAD_GRAYSCALE).astype(np.float32)
        mask = cv2.imread(str(mp), cv2.IMREAD_UNCHANGED).astype(np.float32)
        if mask.ndim == 3:
            mask = mask[..., 0]
        return img, mask

    def normalise(self, img: np.ndarray) -> np.ndarray:
        if self.method == "zscore":
            return ((img - self.mean) / (self.std + 1e-10)).astype(np.float32)
        lo, hi = np.percentile(img, (self.p_low, self.p_high))              # src/utils/data.py:413-416
        return np.clip((img - lo) / max(hi - lo, 1e-3), 0, 1).astype(np.float32)


def augment_d4(img, mask, rng):
    """Dihedral flips / 90-degree rotations (the geometric part of every augmentation level of the reference;
    its intensity / elastic jitter is host-side image processing outside this engine)."""
    k = int(rng.randint(0, 4))
    if k:
        img, mask = np.rot90(img, k), np.rot90(mask, k)
    if rng.rand() < 0.5:
        img, mask = img[:, ::-1], mask[:, ::-1]
    if rng.rand() < 0.5:
        img, mask = img[::-1], mask[::-1]
    return np.ascontiguousarray(img), np.ascontiguousarray(mask)


def batches(ds: TileDataset, batch: int, rng, augment: bool, rank: int, world: int, shuffle: bool, shuffle_rng=None):
    """rng: this rank's stream (augmentation); shuffle_rng: a stream that is IDENTICAL on every rank (same seed, same
    number of draws), so that all ranks slice their shards from one permutation and an epoch covers every tile once."""
    idx = np.arange(len(ds))
    if shuffle:
        (shuffle_rng if shuffle_rng is not None else rng).shuffle(idx)
    per_step = batch * world
    for i in range(0, len(idx), per_step):
        chunk = list(idx[i:i + per_step])
        while len(chunk) < per_step:              # pad the last global batch by repetition (:583-585)
            chunk.append(chunk[-1])
        mine = chunk[rank * batch:(rank + 1) * batch]
        xs, ys = [], []
        for j in mine:
            img, mask = ds.load(int(j))
            if augment:
                img, mask = augment_d4(img, mask, rng)
            xs.append(ds.normalise(img)); ys.append(mask.astype(np.float32))
        yield np.stack(xs), np.stack(ys)


def compute_mean_std(paths):
    vals = np.concatenate([cv2.imread(str(p), cv2.IMREAD_GRAYSCALE).astype(np.float32).reshape(-1) for p in paths])
    return float(vals.mean()), float(vals.std() + 1e-10)


def validate(engine, ds: TileDataset, batch: int, deep_supervision: bool = False):
    """The validation pass of net.fit (train_adipose_unet_v3.py:1316-1324): the training graph with Dropout inactive
    (engine option train_eval_mode), every output's loss, dice_coef and binary_accuracy of main_out.  Keras semantics:
    losses and dice_coef are means over the batches, binary_accuracy is matching pixels / pixels over the whole pass."""
    logs = {}
    acc = [0.0, 0.0]
    engine.set_option("train_eval_mode", 1)
    try:
        rng = np.random.RandomState(0)
        for x, y in batches(ds, batch, rng, False, 0, 1, False):
            m = engine.train_loss(engine.train_forward(x, y))
            a = engine.train_accuracy_read()
            acc[0] += a[0]; acc[1] += a[1]
            for k, v in m.items():
                logs.setdefault(k, []).append(v)
    finally:
        engine.set_option("train_eval_mode", 0)
    out = {k: float(np.mean(v)) for k, v in logs.items()}
    out["binary_accuracy"] = acc[0] / max(acc[1], 1.0)
    return out


def keras_logs(train: dict, val: dict, deep_supervision: bool) -> dict:
    """The `logs` dict Keras hands CSVLogger at epoch end for the reference's compile() (train_adipose_unet_v3.py:858-879):
    single output: loss, dice_coef, binary_accuracy (+ val_*); deep supervision: loss, {main_out,aux_out1,aux_out2}_loss,
    main_out_dice_coef, main_out_binary_accuracy (+ val_*).  No learning-rate entry: CSVLogger precedes the schedule callback
    in the callback list (:1270-1314), so 'lr' is not yet in `logs` when it fixes its columns."""
    def one(d):
        if not deep_supervision:
            return {"loss": d["loss"], "dice_coef": d["dice_coef"], "binary_accuracy": d["binary_accuracy"]}
        return {"loss": d["loss"], "main_out_loss": d["main_out_loss"], "aux_out1_loss": d["aux_out1_loss"], "aux_out2_loss": d["aux_out2_loss"],
                "main_out_dice_coef": d["dice_coef"], "main_out_binary_accuracy": d["binary_accuracy"]}
    logs = one(train)
    if val:
        logs.update({"val_" + k: v for k, v in one(val).items()})
    return logs


class ReduceLROnPlateau:
    """keras.callbacks.ReduceLROnPlateau(monitor, mode='max', factor=0.5, patience=5, min_lr=1e-7) as the reference's
    --no-cosine-schedule branch configures it (train_adipose_unet_v3.py:1303-1313, 1398-1408); Keras defaults
    min_delta=1e-4, cooldown=0."""

    def __init__(self, lr: float, factor: float = 0.5, patience: int = 5, min_lr: float = 1e-7, min_delta: float = 1e-4, cooldown: int = 0):
        self.lr, self.factor, self.patience, self.min_lr, self.min_delta, self.cooldown = lr, factor, patience, min_lr, min_delta, cooldown
        self.best, self.wait, self.cooldown_counter = -np.inf, 0, 0

    def on_epoch_end(self, current: float) -> float:
        if self.cooldown_counter > 0:
            self.cooldown_counter -= 1
            self.wait = 0
        if current > self.best + self.min_delta:
            self.best, self.wait = current, 0
        elif self.cooldown_counter <= 0:
            self.wait += 1
            if self.wait >= self.patience:
                if self.lr > np.float32(self.min_lr):
                    self.lr = max(self.lr * self.factor, self.min_lr)
                    self.cooldown_counter = self.cooldown
                    self.wait = 0
        return self.lr


def main(argv=None) -> int:
    from .. import api, train as T
    from ..weights_io import load_weights_file, save_weights_file
    from .. import synth
    args = build_parser().parse_args(argv)
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_
        torch.cuda.set_device(local)
        dist_.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_
    log = print if rank == 0 else (lambda *a, **k: None)
    log("=" * 80 + "\nTRAIN ADIPOSE U-NET VERSION 3.1 (B200 engine)\n" + "=" * 80)
    log("✓ Deep supervision: ", f"ENABLED (weights main={args.ds_weight_main}, aux1={args.ds_weight_aux1}, aux2={args.ds_weight_aux2})"
        if args.use_deep_supervision else "DISABLED")
    log("✓ Hard example mining:", f"ENABLED (keep {args.hard_example_ratio})" if args.use_hard_mining else "DISABLED")
    log("✓ Label smoothing:", f"ENABLED (eps_pos={args.label_smooth_epsilon_pos}, eps_neg={args.label_smooth_epsilon_neg})"
        if args.use_label_smoothing else "DISABLED")
    data_root = Path(args.data_root)
    tr_img, tr_msk = data_root / "dataset" / "train" / "images", data_root / "dataset" / "train" / "masks"
    va_img, va_msk = data_root / "dataset" / "val" / "images", data_root / "dataset" / "val" / "masks"
    train_paths = sorted(tr_img.glob("*.jpg"))
    if not train_paths:
        print(f"❌ No training tiles under {tr_img}")
        return 1
    mean, std = compute_mean_std(train_paths)
    log(f"Global normalization stats: mean={mean:.2f}, std={std:.2f}")
    train_ds = TileDataset(tr_img, tr_msk, args.normalization_method, mean, std, args.percentile_low, args.percentile_high)
    val_ds = TileDataset(va_img, va_msk, args.normalization_method, mean, std, args.percentile_low, args.percentile_high)
    steps_per_epoch = max(1, math.ceil(len(train_ds) / (args.batch_size * world)))
    if args.max_steps_per_epoch:
        steps_per_epoch = min(steps_per_epoch, args.max_steps_per_epoch)
    stamp = datetime.now().strftime("%Y%m%d_%H%M%S")
    if dist is not None:          # one checkpoint directory for the whole job: rank 0's clock names it
        box = [stamp]
        dist.broadcast_object_list(box, src=0)
        stamp = box[0]
    name = "adipose_sybreosin" + (f"_{args.checkpoint_suffix}" if args.checkpoint_suffix else "")
    ckpt = Path(args.checkpoint_root) / f"{stamp}_{name}_1024_finetune_v3"          # :650-652
    if rank == 0:
        ckpt.mkdir(parents=True, exist_ok=True)
        with open(ckpt / "normalization_stats.json", "w") as f:                      # :1194-1207
            json.dump({"mean": mean, "std": std, "normalization_method": args.normalization_method, "dataset_path": str(data_root),
                       "num_training_images": len(train_paths), "build_timestamp": stamp, "version": "3.0"}, f, indent=2)
        with open(ckpt / "training_settings.log", "w") as f:                         # :984-1053 (sniffed at eval:513-516)
            f.write("=" * 80 + "\nTRAINING SETTINGS LOG - VERSION 3\n" + "=" * 80 + f"\n\nGenerated: {stamp}\n")
            f.write("Script: adipose_unet_b200.cli.train\n" + f"Checkpoint Directory: {ckpt}\n\n" + "-" * 60 + "\nCOMMAND LINE ARGUMENTS\n" + "-" * 60 + "\n")
            settings = dict(vars(args))
            for k, v in settings.items():
                f.write(f"  {k}: {v}\n")
            f.write("\n" + "-" * 60 + "\nMACHINE READABLE FORMAT (JSON)\n" + "-" * 60 + "\n" + json.dumps(settings, indent=2, default=str) + "\n")
    engine = api.Engine(precision=args.precision, device=local, max_forwards=max(8, args.batch_size))
    if args.pretrained_weights and Path(args.pretrained_weights).exists():
        w0 = load_weights_file(args.pretrained_weights, keep_aux=True)
        log(f"✓ Loaded pretrained weights from {args.pretrained_weights}")
    else:
        log("WARNING: No pretrained weights found, training from scratch")
        w0 = synth.init_weights(seed=args.seed)
    if args.use_deep_supervision and "aux_out1/kernel" not in w0:
        # fresh heads as Keras creates them: glorot_uniform kernel, zero bias (Conv2D defaults, :715, 722)
        r0 = np.random.RandomState(args.seed)
        for nm, cin in (("aux_out1", 176), ("aux_out2", 88)):
            lim = np.sqrt(6.0 / (cin + 1))
            w0[nm + "/kernel"] = r0.uniform(-lim, lim, size=(1, 1, cin, 1)).astype(np.float32)
            w0[nm + "/bias"] = np.zeros(1, np.float32)
    if not args.use_deep_supervision:
        w0 = {k: v for k, v in w0.items() if not k.startswith("aux_out")}
    engine.set_weights(w0)
    engine.set_option("train_accuracy", 1)            # compile(metrics=[dice_coef, 'binary_accuracy']), :850-853, 877-878
    if args.use_deep_supervision:
        engine.train_set_deep_supervision(True, args.ds_weight_main, args.ds_weight_aux1, args.ds_weight_aux2)
    engine.train_set_loss(args.hard_example_ratio if args.use_hard_mining else 1.0,
                          args.label_smooth_epsilon_pos if args.use_label_smoothing else 0.0,
                          args.label_smooth_epsilon_neg if args.use_label_smoothing else 0.0)
    rng = np.random.RandomState(args.seed + rank)          # per-rank: augmentation draws
    shuffle_rng = np.random.RandomState(args.seed)         # shared: the epoch permutation (identical on every rank)
    best_overall = -1.0
    for phase, epochs, max_lr, min_lr, warm, freeze, decay in (
            (1, args.epochs_phase1, 1e-4, 1e-7, args.warmup_epochs_phase1, True, 0.999),
            (2, args.epochs_phase2, 1e-5, 1e-8, args.warmup_epochs_phase2, False, args.ema_decay)):
        if epochs <= 0:
            continue
        log(f"\n{'=' * 60}\nPHASE {phase}: {'frozen encoder' if freeze else 'fine-tuning all layers'} ({epochs} epochs)\n{'=' * 60}")
        if dist is not None:
            dist.barrier()           # rank 0 has written phase1_best: every replica restarts phase 2 from the same file
        if phase == 2 and (ckpt / "phase1_best.weights.h5").exists():
            engine.set_weights(load_weights_file(str(ckpt / "phase1_best.weights.h5"), keep_aux=True))       # :1336-1339
        trainer = T.DataParallelTrainer(engine, args.batch_size, TILE, dist=dist, rank=rank, world=world, dropout_rate=0.3,
                                        seed=args.seed + 1000 * phase, optimizer=args.optimizer, freeze_encoder=freeze)
        best_phase, since_best = -1.0, 0
        # EMACallback (:410-505): one instance per phase (phase 1: decay 0.999, never saved; phase 2: --ema-decay, best snapshot
        # by the monitored validation Dice), updated at EPOCH end from the current weights, initialised at the first epoch end
        ema, ema_best, ema_saved = None, -np.inf, False
        logf, w, keys = None, None, None
        if rank == 0:
            logf = open(ckpt / f"phase{phase}_training.log", "w", newline="")
            w = csv.writer(logf)
        plateau = None if args.use_cosine_schedule else ReduceLROnPlateau(max_lr, min_lr=min_lr)      # legacy mode, :1303-1313
        ds_on = bool(args.use_deep_supervision)
        for epoch in range(epochs):
            lr = T.cosine_warmup_lr(epoch, max_lr, min_lr, warm, epochs) if args.use_cosine_schedule else plateau.lr
            t0, steps_logs, acc = time.time(), {}, [0.0, 0.0]
            # the next batch is decoded / augmented while the device runs this one (the reference's dataset.prefetch, :620)
            for step, (x, y) in enumerate(C.prefetch(batches(train_ds, args.batch_size, rng, args.augmentation_level != "none", rank, world, True, shuffle_rng))):
                if step >= steps_per_epoch:
                    break
                out = trainer.step(x, y, lr)
                for k, v in out.items():
                    steps_logs.setdefault(k, []).append(v)
                a = engine.train_accuracy_read()
                acc[0] += a[0]; acc[1] += a[1]
            tlog = {k: float(np.mean(v)) for k, v in steps_logs.items()}
            tlog["binary_accuracy"] = acc[0] / max(acc[1], 1.0)
            vlog = validate(engine, val_ds, args.batch_size, ds_on) if len(val_ds) else {}
            logs = keras_logs(tlog, vlog, ds_on)
            vloss, vdice = vlog.get("loss", float("nan")), vlog.get("dice_coef", float("nan"))
            log(f"Epoch {epoch + 1}/{epochs} - {time.time() - t0:.0f}s - " + " - ".join(f"{k}: {v:.4f}" for k, v in logs.items()) + f" - lr: {lr:.2e}")
            if plateau is not None:
                # monitor = val_main_out_dice_coef / val_dice_coef (:1268), the training value when there is no validation split
                plateau.on_epoch_end(vdice if np.isfinite(vdice) else tlog["dice_coef"])
            if rank == 0:
                if keys is None:                      # CSVLogger: 'epoch' + the sorted keys of the first epoch's logs
                    keys = sorted(logs)
                    w.writerow(["epoch"] + keys)
                w.writerow([epoch] + [logs.get(k, "NA") for k in keys]); logf.flush()
                score = vdice if np.isfinite(vdice) else tlog["dice_coef"]
                if score > best_phase:
                    best_phase, since_best = score, 0
                    save_weights_file(str(ckpt / f"phase{phase}_best.weights.h5"), engine.get_weights())
                else:
                    since_best += 1
                if score > best_overall:
                    best_overall = score
                    save_weights_file(str(ckpt / "weights_best_overall.weights.h5"), engine.get_weights())
                cur = engine.get_weights()
                ema = {k: v.copy() for k, v in cur.items()} if ema is None else \
                    {k: (decay * ema[k] + (1 - decay) * cur[k]).astype(np.float32) for k in cur}
                if phase == 2 and np.isfinite(vdice) and vdice > ema_best:          # save_best_only on the monitor
                    ema_best, ema_saved = vdice, True
                    save_weights_file(str(ckpt / "weights_ema.weights.h5"), ema)
            stop = since_best >= 15                                  # EarlyStopping(patience=15), :1280-1284, 1370-1374
            if dist is not None:
                import torch
                flag = torch.tensor([1 if (rank == 0 and stop) else 0], device=f"cuda:{local}")
                dist.broadcast(flag, 0)
                stop = bool(flag.item())
            if stop:
                log(f"Early stopping at epoch {epoch + 1}")
                break
        if rank == 0:
            logf.close()
            save_weights_file(str(ckpt / f"weights_phase{phase}_final.weights.h5"), engine.get_weights())
            if phase == 2 and ema is not None and not ema_saved:                      # on_train_end fallback (:484-491)
                save_weights_file(str(ckpt / "weights_ema.weights.h5"), ema)
        trainer.close()
    if rank == 0 and (ckpt / "phase2_best.weights.h5").exists():       # best_overall := best phase-2 weights (:1425-1428)
        save_weights_file(str(ckpt / "weights_best_overall.weights.h5"),
                          load_weights_file(str(ckpt / "phase2_best.weights.h5"), keep_aux=True))
    log(f"\n✓ Training complete. Checkpoints in {ckpt}\n  - phase1_best.weights.h5\n  - phase2_best.weights.h5\n"
        f"  - weights_best_overall.weights.h5\n  - weights_ema.weights.h5")
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
