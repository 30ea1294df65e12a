"""Per-kernel time breakdown of one training step (CUDA events around every pre-bound launch, eager mode).
Usage: python scripts/profile_step.py [--batch 16] [--hw 256] [--dtype bf16]  ->  gpurun_out/step_breakdown.txt"""
import argparse, os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16); ap.add_argument("--hw", type=int, default=256)
ap.add_argument("--dtype", default="bf16"); ap.add_argument("--classes", type=int, default=6)
ap.add_argument("--detail", action="store_true")
a = ap.parse_args()
import __graft_entry__ as ge; ge.build()
from oracle import resuneta_oracle as O
from resuneta_b200 import Adam, Tanimoto_dual_loss
from resuneta_b200.builder import build_model
heads = ("seg", "bound", "dist", "color")
m = build_model((a.hw, a.hw, 3), a.classes, True, "v2", dtype=a.dtype)
m.use_cuda_graph = False
m.compile(optimizer=Adam(lr=1e-3), loss={h: Tanimoto_dual_loss() for h in heads})
x, y = O.synth_batch(a.batch, a.hw, 3, a.classes, seed=1)
for _ in range(3): m.train_on_batch(x, y)
pl = m.net.plan(a.batch, True, m.loss_spec)
st = torch.cuda.current_stream().cuda_stream
seq = [("pack", m.net.pack_launch)] if m.net.pack_launch else []
seq += [("fwd", op) for op in pl.fwd] + [("fwd", pl.bn_update)] + [("bwd", op) for op in pl.bwd] + [("opt", m._opt_launch)]
m._push_lr(); pl.scratch.zero_(); m.net.params.grad.zero_()
evs = []
torch.cuda.synchronize()
t0 = torch.cuda.Event(enable_timing=True); t0.record()
for phase, op in seq:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); op(st); e1.record(); evs.append((phase, op, e0, e1))
t1 = torch.cuda.Event(enable_timing=True); t1.record()
torch.cuda.synchronize()
agg = collections.OrderedDict()
rows = []
for phase, op, e0, e1 in evs:
    name = getattr(op, "kernel", "?") + ("/" + op.tag if hasattr(op, "tag") else "")
    ms = e0.elapsed_time(e1)
    k = (phase, name)
    c, t = agg.get(k, (0, 0.0)); agg[k] = (c + 1, t + ms)
    rows.append((phase, name, ms, getattr(op, "flops", 0.0), getattr(getattr(op, "cell", [op])[0] if getattr(op, "cell", None) else op, "ints", ())))
tot = sum(t for _, t in agg.values())
lines = [f"step (eager, event-timed): {t0.elapsed_time(t1):.2f} ms wall, {tot:.2f} ms summed kernels; plan bytes {pl.bytes/2**30:.2f} GiB"]
for (phase, name), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"{phase:4s} {name:44s} n={c:4d} {t:9.3f} ms {100*t/tot:5.1f}%")
if a.detail:
    lines.append("---- conv launches (ms, TFLOP/s)")
    for phase, name, ms, fl, ints in rows:
        if fl: lines.append(f"{phase} {name:40s} {ms:8.3f} ms {fl/ms/1e9:8.1f} TF")
    lines.append("---- non-conv launches over 40 us")
    for phase, name, ms, fl, ints in sorted(rows, key=lambda r: -r[2]):
        if not fl and ms > 0.04: lines.append(f"{phase} {name:28s} {ms:8.3f} ms  args={ints}")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "step_breakdown.txt"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:40]))
