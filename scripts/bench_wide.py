"""Micro-benchmark of the C >= 128 3x3 convolution launches of config 2 (conv_tc2 forward / data gradient, conv_tc weight
gradient): CUDA events, rotating buffers.  RSA_TC2_MT=1 gives the one-sub-tile kernel.  Usage: python scripts/bench_wide.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge
ge.build()
from resuneta_b200 import _capi
lib = _capi.Lib()
N, dt = 16, torch.bfloat16
st = torch.cuda.current_stream().cuda_stream
NB = 6


def timeit(make, reps=24):
    ops = [make(i) for i in range(NB)]
    for op in ops: op(st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(reps):
        ops[r % NB](st)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for C, H in ((128, 64), (256, 32), (512, 16), (1024, 8)):
    xs = [torch.randn(N, H, H, C, device="cuda").to(dt) for _ in range(NB)]
    outs = [torch.zeros(N, H, H, C, device="cuda", dtype=dt) for _ in range(NB)]
    wt = (torch.randn(9, C, C, device="cuda") / (3 * C ** 0.5)).to(dt).view(-1)
    bias = torch.randn(C, device="cuda")
    stats = torch.zeros(2 * C, dtype=torch.float64, device="cuda")
    dw = torch.zeros(9 * C * C, device="cuda")
    flops = 2.0 * N * H * H * 9 * C * C
    for d in ((1, 3, 15) if C <= 256 else (1,)):
        tf = timeit(lambda i: lib.conv_tc2_fwd(xs[i], None, wt, C, bias, outs[i], N, H, H, C, taps=9, dil=d, stats=stats))
        ta = timeit(lambda i: lib.conv_tc2_fwd(xs[i], None, wt, C, bias, outs[i], N, H, H, C, taps=9, dil=d, accumulate=True))
        tg = timeit(lambda i: lib.conv_tc2_fwd(xs[i], None, wt, C, None, outs[i], N, H, H, C, taps=9, dil=-d, mask=xs[(i + 1) % NB]))
        tw = timeit(lambda i: lib.conv_tc_wgrad(xs[i], outs[(i + 1) % NB], dw, N, H, H, C, C, d))
        print(f"C={C:4d} H={H:2d} d={d:2d}  fwd+stats {tf:6.1f} us ({flops / tf / 1e6:5.0f} TF)  fwd accumulate {ta:6.1f} us  "
              f"dgrad+mask {tg:6.1f} us ({flops / tg / 1e6:5.0f} TF)  wgrad {tw:6.1f} us ({flops / tw / 1e6:5.0f} TF)", flush=True)
