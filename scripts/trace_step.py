"""Kernel timeline of graph-replayed training steps (CUPTI via torch.profiler): per-kernel-name totals per step.
Usage: python scripts/trace_step.py [--steps 3]  -> gpurun_out/trace_step_<tag>.txt"""
import argparse, os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
ap = argparse.ArgumentParser(); ap.add_argument("--steps", type=int, default=3); ap.add_argument("--tag", default="a")
a = ap.parse_args()
import __graft_entry__ as ge; ge.build()
from oracle import resuneta_oracle as O
from resuneta_b200 import Adam, Tanimoto_dual_loss
from resuneta_b200.builder import build_model
heads = ("seg", "bound", "dist", "color")
m = build_model((256, 256, 3), 6, True, "v2", dtype="bf16")
m.compile(optimizer=Adam(lr=1e-3), loss={h: Tanimoto_dual_loss() for h in heads})
x, y = O.synth_batch(16, 256, 3, 6, seed=1)
for _ in range(4): m.train_on_batch(x, y)
pl = m.net.plan(16, True, m.loss_spec)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(a.steps):
        m._push_lr(); m._execute(pl, True)
    torch.cuda.synchronize()
agg = collections.OrderedDict(); tot = 0.0
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
t0 = min(e.time_range.start for e in evs); t1 = max(e.time_range.end for e in evs)
for e in evs:
    k = e.name.split("(")[0].replace("void ", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    c, t = agg.get(k, (0, 0.0)); agg[k] = (c + 1, t + e.cuda_time if hasattr(e, "cuda_time") else t + (e.time_range.end - e.time_range.start)); 
    tot += (e.time_range.end - e.time_range.start)
lines = [f"{a.steps} graph-replayed steps: span {(t1 - t0) / a.steps / 1e3:.3f} ms/step, summed kernel time {tot / a.steps / 1e3:.3f} ms/step"]
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"{k[:70]:70s} n/step={c / a.steps:6.1f} {t / a.steps / 1e3:8.3f} ms/step")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", f"trace_step_{a.tag}.txt"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:45]))
