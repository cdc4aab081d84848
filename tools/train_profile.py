"""Per-kernel timing of one training step (forward + backward + Adam) at a given size/batch."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import adipose_unet_b200 as A
from adipose_unet_b200 import api

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 8
prec = sys.argv[3] if len(sys.argv) > 3 else "bf16"
eng = api.Engine(precision=prec, max_forwards=8)
eng.set_weights(A.synth.init_weights())
if os.environ.get("ADP_FUSE_DROPOUT") is not None:      # experiment: dropout in the conv epilogue on / off
    eng.set_option("fuse_dropout", int(os.environ["ADP_FUSE_DROPOUT"]))
eng.train_begin(nb, S, dropout_rate=0.3)
rng = np.random.default_rng(0)
x = rng.standard_normal((nb, S, S)).astype(np.float32)
y = (rng.random((nb, S, S)) < 0.3).astype(np.float32)
for _ in range(2):
    eng.train_step(x, y, 1e-4)
t0 = time.time()
REPS = 10
for _ in range(REPS):
    out = eng.train_step(x, y, 1e-4)
eng.synchronize()
dt = (time.time() - t0) / REPS
fl = 3 * A.layers.forward_flops(S) * nb
print(f"train step S={S} nb={nb} {prec}: {dt*1e3:.1f} ms wall (host buffers), {nb/dt:.2f} tiles/s, ~{fl/dt/1e12:.0f} TFLOP/s (3x fwd flops), loss={out['loss']:.4f}")
eng.profile(True)
eng.train_step(x, y, 1e-4)
rows = eng.profile_rows()
tot = sum(r["ms"] for r in rows)
agg = {}
for r in rows:
    k = r["name"].split("/")[0]
    a = agg.setdefault(k, dict(ms=0, flops=0, bytes=0, n=0))
    a["ms"] += r["ms"]; a["flops"] += r["flops"]; a["bytes"] += r["bytes"]; a["n"] += r["launches"]
print(f"profiled step: {tot:.1f} ms in {sum(r['launches'] for r in rows)} launches")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"  {k:28s} {a['ms']:8.2f} ms {100*a['ms']/tot:5.1f}%  n={a['n']:3d}  {a['flops']/max(a['ms'],1e-9)/1e9:8.1f} TFLOP/s {a['bytes']/max(a['ms'],1e-9)/1e6:8.1f} GB/s")
print("per layer (conv kernels):")
for r in sorted(rows, key=lambda r: -r["ms"]):
    if "/" in r["name"]:
        print(f"  {r['name']:40s} {r['ms']:7.3f} ms  {r['flops']/max(r['ms'],1e-9)/1e9:8.1f} TFLOP/s")
