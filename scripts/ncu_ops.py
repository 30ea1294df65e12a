"""Target for `ncu --profile-from-start off`: after three warm-up steps of BASELINE config 2 (batch 16, bf16) it brackets
with cudaProfilerStart/Stop ONLY the launches of one family, in plan order, and writes the matching list
(gpurun_out/r2_ops_<what>.json: kernel, tag, integer arguments, algorithmic FLOPs / bytes), so that row i of ncu's CSV is
entry i of the list.  what = conv (the tagged 3x3 convolution launches the roofline fraction is computed on) | hbm (the
bandwidth-bound launches: BatchNorm, pooling, heads, losses, optimizer).  Numbers printed under ncu are never bench values.

    ncu --profile-from-start off --clock-control none --metrics <...> --csv --log-file gpurun_out/r2_ncu_conv.csv \
        python scripts/ncu_ops.py conv
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import __graft_entry__ as ge  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "conv"
ge.build()
from bench import HEADS, synth_train  # noqa: E402
from resuneta_b200 import Adam, Tanimoto_dual_loss  # noqa: E402
from resuneta_b200.builder import build_model  # noqa: E402

torch.cuda.set_device(0)
B = int(os.environ.get("B", "16"))
model = build_model((256, 256, 3), 6, True, "v2", dtype="bf16", seed=1234)
model.use_cuda_graph = False
model.compile(optimizer=Adam(lr=1e-3), loss={h: Tanimoto_dual_loss() for h in HEADS}, loss_weights={h: 1.0 for h in HEADS})
x, y = synth_train(2, B, 1234)
for _ in range(3):
    model.train_on_batch(x, y)
pl = model.net.plan(B, True, model.loss_spec)
seq = [model.net.pack_launch] + list(pl.fwd) + [pl.bn_update] + list(pl.bwd) + [model._opt_launch]
if what == "conv":
    ops = [op for op in seq if getattr(op, "tag", None)]
else:
    ops = [op for op in seq if getattr(op, "hbm_bytes", None)]
rows = []
for op in ops:
    inner = op.cell[0] if getattr(op, "cell", None) else op
    rows.append(dict(kernel=getattr(op, "kernel", getattr(inner, "kernel", "?")), tag=getattr(op, "tag", None),
                     ints=list(getattr(inner, "ints", ())), flops=getattr(op, "flops", None),
                     bytes=getattr(op, "conv_bytes", None) or getattr(op, "hbm_bytes", None)))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", f"r2_ops_{what}.json"), "w"))
stream = torch.cuda.current_stream().cuda_stream
torch.cuda.synchronize()
torch.cuda.profiler.start()
for op in ops:
    op(stream)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(f"{len(ops)} {what} launches profiled")
