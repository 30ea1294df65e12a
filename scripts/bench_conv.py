"""Micro-benchmark of the 3x3 convolution kernels at the config-2 layer shapes (CUDA events, L2 flushed by rotating
over distinct buffers larger than L2).  Usage: python scripts/bench_conv.py [--C 32]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge
ge.build()
from resuneta_b200 import _capi
lib = _capi.Lib()
ap = argparse.ArgumentParser(); ap.add_argument("--C", type=int, default=32); ap.add_argument("--N", type=int, default=16)
a = ap.parse_args()
C, N = a.C, a.N
H = W = 256 * 32 // C
dt = torch.bfloat16
st = torch.cuda.current_stream().cuda_stream
NB = 4   # rotating buffer sets: 4 x (x, out) x 67 MB > L2
xs = [torch.randn(N, H, W, C, device="cuda").to(dt) for _ in range(NB)]
outs = [torch.zeros(N, H, W, C, device="cuda", dtype=dt) for _ in range(NB)]
wt = [(torch.randn(9, C, C, device="cuda") / (3 * C ** 0.5)).to(dt) for _ in range(4)]
bias = [torch.randn(C, device="cuda") for _ in range(4)]
stats = torch.zeros(2 * C, dtype=torch.float64, device="cuda")
flops = 2.0 * N * H * W * 9 * C * C

def timeit(make, reps=20):
    ops = [make(i) for i in range(NB)]
    for op in ops: op(st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(reps):
        ops[r % NB](st)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

for d in (1, 3, 15, 31):
    t2 = timeit(lambda i: lib.conv_tc2_fwd(xs[i], None, wt[0].view(-1), C, bias[0], outs[i], N, H, W, C, taps=9, dil=d, stats=stats))
    t2p = timeit(lambda i: lib.conv_tc2_fwd(xs[i], None, wt[0].view(-1), C, bias[0], outs[i], N, H, W, C, taps=9, dil=d))
    line = f"C={C} H={H} d={d:2d}  tc2 stats {t2:7.1f} us ({flops/t2/1e6:6.0f} TF)  tc2 plain {t2p:7.1f} us"
    if lib.conv_tc3_supported(N, H, W, C):
        t3 = timeit(lambda i: lib.conv_tc3_fwd([xs[i]], [wt[0].view(-1)], [bias[0]], [d], outs[i], N, H, W, C, stats=stats))
        t3p = timeit(lambda i: lib.conv_tc3_fwd([xs[i]], [wt[0].view(-1)], [bias[0]], [d], outs[i], N, H, W, C))
        t3r = timeit(lambda i: lib.conv_tc3_fwd([xs[i]], [wt[0].view(-1)], [bias[0]], [d], outs[i], N, H, W, C, accumulate=True))
        fst = torch.randn(2 * C, dtype=torch.float64, device="cuda").abs_() * (N * H * W) + 1.0
        fst[C:] = fst[:C] ** 2 / (N * H * W) + (N * H * W)          # variance 1
        bnr = (xs[(0 + 2) % NB], fst, float(N * H * W), 1e-3, bias[1], bias[2], 1)
        t3b = timeit(lambda i: lib.conv_tc3_fwd([xs[i]], [wt[0].view(-1)], [None], [-d], outs[i], N, H, W, C, stats=stats,
                                                bnr=(xs[(i + 2) % NB],) + bnr[1:]))
        t3m = timeit(lambda i: lib.conv_tc3_fwd([xs[i]], [wt[0].view(-1)], [None], [-d], outs[i], N, H, W, C, mask=xs[(i + 2) % NB]))
        line += f" | tc3 stats {t3:7.1f} us ({flops/t3/1e6:6.0f} TF)  plain {t3p:7.1f} us  accum {t3r:7.1f} us  dgrad+bnr {t3b:7.1f} us  dgrad+mask {t3m:7.1f} us"
    print(line, flush=True)
if lib.conv_tc3_supported(N, H, W, C) and C == 32:
    dils = [1, 3, 15, 31]
    t4 = timeit(lambda i: lib.conv_tc3_fwd([xs[(i + k) % NB] for k in range(4)], [w.view(-1) for w in wt], bias, dils, outs[i], N, H, W, C,
                                           residual=xs[i], relu=True))
    print(f"C={C} fused 4-branch (1,3,15,31) + identity: {t4:7.1f} us ({4*flops/t4/1e6:6.0f} TF)")
if lib.conv_tc3_supported(N, H, W, C):
    xw = torch.randn(N, H, W, C, device="cuda").to(dt)
    dw = torch.zeros(9 * C * C, device="cuda")
    for d in (1, 3, 15, 31):
        tw = timeit(lambda i: lib.conv_tc_wgrad(xs[i], outs[(i + 1) % NB], dw, N, H, W, C, C, d))
        if not lib.conv_tc3_wgrad_supported(N, H, W, C, d):
            print(f"C={C} wgrad d={d}: {tw:7.1f} us ({flops/tw/1e6:6.0f} TF)"); continue
        tw3 = timeit(lambda i: lib.conv_tc3_wgrad(xs[i], outs[(i + 1) % NB], dw, N, H, W, C, d))
        print(f"C={C} wgrad d={d}: {tw:7.1f} us ({flops/tw/1e6:6.0f} TF) | tc3 {tw3:7.1f} us ({flops/tw3/1e6:6.0f} TF)")
