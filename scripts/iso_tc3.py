"""Which pipeline bounds conv_tc3?  Times the config-2 thin-layer launches (no tracing) with parts of the kernel switched
off through RSA_TC3_DEBUG (1 no MMAs, 2 no epilogue data movement, 4 no TMA operand loads, 8 no TMA stores; results are
wrong, only the time matters; 16 / 32 / 64 switch off the epilogue's tcgen05.ld / proxy fence / staging stores).  CUDA events over rotating buffer sets larger than L2."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge; ge.build()
from resuneta_b200 import _capi
lib = _capi.Lib()
N, dt = int(os.environ.get("N", "16")), torch.bfloat16
st = torch.cuda.current_stream().cuda_stream
NB = 4
names = {0: "full", 8: "no store", 2: "no epilogue", 1: "no MMA", 9: "noMMA,no store", 17: "noMMA,no tmem ld", 33: "noMMA,no fence",
         65: "noMMA,no STS", 121: "noMMA,epi shell", 3: "no MMA, no epi", 7: "barriers only"}
for C, H in ((32, 256), (64, 128)):
    xs = [torch.randn(N, H, H, C, device="cuda").to(dt) for _ in range(NB)]
    outs = [torch.zeros(N, H, H, C, device="cuda", dtype=dt) for _ in range(NB)]
    res = [torch.randn(N, H, H, C, device="cuda").to(dt) for _ in range(NB)]
    w = (torch.randn(9, C, C, device="cuda") / (3 * C ** 0.5)).to(dt).view(-1); b = torch.randn(C, device="cuda")
    stats = torch.zeros(2 * C, dtype=torch.float64, device="cuda")
    fst = torch.rand(2 * C, dtype=torch.float64, device="cuda") * (N * H * H)
    fst[C:] = fst[:C] ** 2 / (N * H * H) + (N * H * H)
    for d in (1, 15):
        for label in ("plain", "stats", "accum", "bnr"):
            row = []
            for dbg in (0, 8, 2, 1, 9, 17, 33, 65, 121, 3, 7):
                os.environ["RSA_TC3_DEBUG"] = str(dbg)
                ops = []
                for i in range(NB):
                    kw = dict(plain={}, stats=dict(stats=stats), accum=dict(accumulate=True),
                              bnr=dict(stats=stats, bnr=(res[i], fst, float(N * H * H), 1e-3, b, b, 1)))[label]
                    ops.append(lib.conv_tc3_fwd([xs[i]], [w], [b], [d if label != "bnr" else -d], outs[i], N, H, H, C, **kw))
                for op in ops: op(st)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    for op in ops: op(st)
                e1.record(); torch.cuda.synchronize()
                row.append((names[dbg], e0.elapsed_time(e1) * 1e3 / (5 * NB)))
            os.environ["RSA_TC3_DEBUG"] = "0"
            print(f"C={C} d={d:2d} {label:6s} | " + "  ".join(f"{n}: {t:5.1f}" for n, t in row), flush=True)
