"""Debug: per-parameter update difference between emission order and the hostile order (tests/sched_util.py)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import __graft_entry__ as ge; ge.build()
from oracle import resuneta_oracle as O
from resuneta_b200 import SGD, Tanimoto_dual_loss
from resuneta_b200.builder import build_model
from sched_util import adversarial_run
from test_model_gpu import rand_params, LW
hw, n, B = int(os.environ.get("HW", "64")), 5, int(os.environ.get("B", "4"))
p = rand_params("v2", hw, 3, n)
x, y = O.synth_batch(B, hw, 3, n, seed=21, block=16)
modes = sys.argv[1:] or ["serial", "serial", "hostile"]
upd = []
for mode in modes:
    m = build_model((hw, hw, 3), n, True, "v2", dtype="bf16")
    m.use_cuda_graph = False
    m.net.set_weights(p)
    m.compile(optimizer=SGD(lr=1.0), loss={k: Tanimoto_dual_loss() for k in LW}, loss_weights=LW)
    if mode == "hostile":
        type(m)._run_ops_saved = type(m)._run_ops
        type(m)._run_ops = lambda self, ops, stream: adversarial_run(list(ops), stream)
    elif mode == "nolane":   # hostile only towards side launches
        type(m)._run_ops_saved = type(m)._run_ops
        type(m)._run_ops = lambda self, ops, stream: adversarial_run(list(ops), stream, nl=1)
    elif mode == "one":
        os.environ["RSA_WGRAD_STREAM"] = "0"; os.environ["RSA_LANES"] = "0"
    before = {k: v.clone() for k, v in m.net.get_weights().items()}
    m.train_on_batch(x, y)
    after = m.net.get_weights()
    if mode in ("hostile", "nolane"):
        type(m)._run_ops = type(m)._run_ops_saved
    os.environ.pop("RSA_WGRAD_STREAM", None); os.environ.pop("RSA_LANES", None)
    upd.append({k: (after[k] - before[k]).double() for k in before if "/moving_" not in k})
for i in range(1, len(modes)):
    rs = []
    for k in upd[0]:
        a, b = upd[i][k], upd[0][k]
        if float(b.norm()) > 1e-6 * max(1.0, float(b.numel()) ** 0.5):
            rs.append((float((a - b).norm() / b.norm()), k))
    rs.sort(reverse=True)
    tot = float(torch.cat([(upd[i][k] - upd[0][k]).flatten() for k in upd[0]]).norm() / torch.cat([upd[0][k].flatten() for k in upd[0]]).norm())
    print(f"== {modes[i]} vs {modes[0]}: whole-update rel {tot:.3e}; median {rs[len(rs)//2][0]:.2e}; worst {[(f'{r:.2e}', k) for r, k in rs[:4]]}")
