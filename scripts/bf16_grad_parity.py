"""Measure: whole-gradient rel-L2 of the bf16 tensor-core step against the fp64 / fp32 CPU oracle (and fp32 GPU mode)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import __graft_entry__ as ge; ge.build()
from oracle import resuneta_oracle as O
from resuneta_b200 import SGD, Tanimoto_dual_loss
from resuneta_b200.builder import build_model
from test_model_gpu import rand_params, LW
for variant, hw, n, B in [(a.split(":")[0], int(a.split(":")[1]), 5, int(a.split(":")[2])) for a in (sys.argv[1:] or ["v2:64:4", "v2:128:2", "v1:64:4"])]:
    p = rand_params(variant, hw, 3, n)
    x, y = O.synth_batch(B, hw, 3, n, seed=21, block=16)
    yt = {k: torch.from_numpy(v) for k, v in y.items()}
    p64 = {k: v.double() for k, v in p.items()}
    tot, per, _, grads, _ = O.loss_and_grads(p64, torch.from_numpy(x).double(), {k: v.double() for k, v in yt.items()},
                                             {k: O.tanimoto_dual_loss for k in LW}, LW, n, variant=variant)
    for dtype in os.environ.get("DTYPES", "fp32,bf16").split(","):
        m = build_model((hw, hw, 3), n, True, variant, dtype=dtype)
        m.net.set_weights(p)
        m.compile(optimizer=SGD(lr=1.0), loss={k: Tanimoto_dual_loss() for k in LW}, loss_weights=LW)
        before = {k: v.clone() for k, v in m.net.get_weights().items()}
        res = m.train_on_batch(x, y)
        after = m.net.get_weights()
        keys = [k for k in grads if k in before and "/moving_" not in k]
        g_mine = torch.cat([(before[k] - after[k]).double().flatten() for k in keys])
        g_ref = torch.cat([grads[k].double().flatten() for k in keys])
        rel = float((g_mine - g_ref).norm() / g_ref.norm())
        cos = float((g_mine @ g_ref) / (g_mine.norm() * g_ref.norm()))
        rows = []
        for k in keys:
            a, b = (before[k] - after[k]).double().flatten(), grads[k].double().flatten()
            if float(b.norm()) >= 3e-3 * float(g_ref.norm()):
                rows.append((float(a @ b / (a.norm() * b.norm() + 1e-300)), float(a.norm() / b.norm()), float(b.norm() / g_ref.norm()), k))
        rows.sort()
        print("   lowest per-parameter cosines (cos, |mine|/|ref|, share of total norm):", [(f"{c:.3f}", f"{r:.2f}", f"{sh:.3f}", k) for c, r, sh, k in rows[:14]])
        print(f"{variant} hw={hw} B={B} {dtype}: loss {res[0]:.6f} oracle {tot.item():.6f}; whole-gradient rel-L2 {rel:.3e}, cosine {cos:.6f}", flush=True)
