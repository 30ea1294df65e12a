"""Per-launch durations INSIDE the graph-replayed training step (CUPTI via torch.profiler), attributed to the plan's
launches.  Eager CUDA-event timing inflates short kernels by 3-5 us each; this is the timeline the step really runs.
An eager pass with a marker fill between launches gives the number of kernels each launch issues; the replayed step's
kernels (same order) are then split by those counts.
Usage: python scripts/trace_launches.py [--steps 3]  ->  gpurun_out/trace_launches.txt"""
import argparse, os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
ap = argparse.ArgumentParser(); ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
import __graft_entry__ as ge; ge.build()
from oracle import resuneta_oracle as O
from resuneta_b200 import Adam, Tanimoto_dual_loss
from resuneta_b200.builder import build_model
from torch.profiler import profile, ProfilerActivity
heads = ("seg", "bound", "dist", "color")
m = build_model((256, 256, 3), 6, True, "v2", dtype="bf16")
m.compile(optimizer=Adam(lr=1e-3), loss={h: Tanimoto_dual_loss() for h in heads})
x, y = O.synth_batch(16, 256, 3, 6, seed=1)
for _ in range(4): m.train_on_batch(x, y)
pl = m.net.plan(16, True, m.loss_spec)
st = torch.cuda.current_stream().cuda_stream
seq = [("pack", m.net.pack_launch)] if m.net.pack_launch else []
seq += [("fwd", op) for op in pl.fwd] + [("fwd", pl.bn_update)] + [("bwd", op) for op in pl.bwd] + [("opt", m._opt_launch)]


def kernels(prof):
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "Memcpy" not in e.name and "Memset" not in e.name]
    return sorted(evs, key=lambda e: e.time_range.start)


# ---- eager pass with markers -> kernels per launch
marker = torch.zeros(1, device="cuda")
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    marker.fill_(1.0)
    for _, op in seq:
        op(st); marker.fill_(1.0)
    torch.cuda.synchronize()
counts, cur, started = [], 0, False
for e in kernels(prof):
    is_marker = "FillFunctor" in e.name
    if not started:
        started = is_marker
        continue
    if is_marker:
        counts.append(cur); cur = 0
    else:
        cur += 1
assert len(counts) == len(seq), (len(counts), len(seq))

# ---- replayed steps
m._push_lr(); m._execute(pl, True); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(a.steps):
        m._execute(pl, True)
    torch.cuda.synchronize()
evs = [e for e in kernels(prof) if "FillFunctor" not in e.name]
per = sum(counts)
assert len(evs) == per * a.steps, (len(evs), per, a.steps)
dur = [0.0] * len(seq)
for s in range(a.steps):
    i = s * per
    for j, c in enumerate(counts):
        for e in evs[i:i + c]:
            dur[j] += (e.time_range.end - e.time_range.start) / 1e3 / a.steps      # ms
        i += c


def describe(op):
    name = getattr(op, "kernel", "?") + ("/" + op.tag if hasattr(op, "tag") else "")
    cell = getattr(op, "cell", None)
    ints = getattr(cell[0], "ints", ()) if cell else getattr(op, "ints", ())
    return name, ints


agg = collections.OrderedDict()
rows = []
for (phase, op), ms in zip(seq, dur):
    name, ints = describe(op)
    c, t = agg.get((phase, name), (0, 0.0)); agg[(phase, name)] = (c + 1, t + ms)
    rows.append((phase, name, ms, getattr(op, "flops", 0.0), ints))
tot = sum(dur)
lines = [f"graph-replayed step, {a.steps} steps averaged: {tot:.3f} ms summed kernel time, {len(seq)} launches, {per} kernels"]
for (phase, name), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"{phase:4s} {name:44s} n={c:4d} {t:9.3f} ms {100 * t / tot:5.1f}%")
lines.append("---- every launch in order (ms; TFLOP/s where the launch declares flops; integer args)")
for i, (phase, name, ms, fl, ints) in enumerate(rows):
    tf = f"{fl / ms / 1e9:7.1f} TF" if fl and ms > 0 else " " * 10
    lines.append(f"{i:4d} {phase:4s} {name:40s} {ms * 1e3:8.1f} us {tf}  {tuple(ints)[:16]}")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "trace_launches.txt"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:45]))
