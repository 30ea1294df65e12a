"""Micro-benchmark of the thin 1x1 convolution launches of config 2 (rsa_conv_tc2_fwd, which routes them to the streaming
kernel of pw_stream.cu unless RSA_PW_STREAM=0) and of rsa_pw_wgrad_tc: CUDA events, rotating over buffer sets larger than
L2, algorithmic bytes / time against the measured HBM peak.  Usage: python scripts/bench_pw.py [--N 16]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge
ge.build()
from resuneta_b200 import _capi
lib = _capi.Lib()
ap = argparse.ArgumentParser(); ap.add_argument("--N", type=int, default=16); ap.add_argument("--wgrad", type=int, default=1)
a = ap.parse_args()
N, dt = a.N, torch.bfloat16
st = torch.cuda.current_stream().cuda_stream
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    PEAK = float(PEAK.get("hbm_gbs") or 0) or None
except Exception:
    PEAK = None
NB = 4


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


def T(H, C, s=1.0):
    return [(torch.randn(N, H, H, C, device="cuda") * s).to(dt) for _ in range(NB)]


# (label, C0, C1, Cout, H, in_stride, out_stride, ups, residual, mask, accumulate, relu, stats, k_base, k_total)
CASES = [
    ("combine 32+32->32 relu stats", 32, 32, 32, 256, 1, 1, (), 0, 0, 0, 1, 1, 0, 0),
    ("dec1 skip+up2 ->32 stats", 32, 0, 32, 256, 1, 1, (1,), 0, 0, 0, 0, 1, 16, 48),
    ("PSP out 32->32 +3 ups +res", 32, 0, 32, 256, 1, 1, (1, 2, 3), 1, 0, 0, 0, 1, 32, 64),
    ("32->8", 32, 0, 8, 256, 1, 1, (), 0, 0, 0, 0, 1, 0, 0),
    ("8->32", 8, 0, 32, 256, 1, 1, (), 0, 0, 0, 0, 1, 0, 64),
    ("dgrad 32->32 plain", 32, 0, 32, 256, 1, 1, (), 0, 0, 0, 0, 0, 0, 0),
    ("dgrad 32->32 acc", 32, 0, 32, 256, 1, 1, (), 0, 0, 1, 0, 0, 0, 0),
    ("dgrad 32->32 mask+acc", 32, 0, 32, 256, 1, 1, (), 0, 1, 1, 0, 0, 0, 0),
    ("dgrad 8->32 mask+acc", 8, 0, 32, 256, 1, 1, (), 0, 1, 1, 0, 0, 0, 0),
    ("dgrad 32->8", 32, 0, 8, 256, 1, 1, (), 0, 0, 0, 0, 0, 0, 0),
    ("down 32->64 s2", 32, 0, 64, 128, 2, 1, (), 0, 0, 0, 0, 1, 0, 0),
    ("dgrad 64->32 out_stride 2 acc", 64, 0, 32, 128, 1, 2, (), 0, 0, 1, 0, 0, 0, 0),
    ("64->16", 64, 0, 16, 128, 1, 1, (), 0, 0, 0, 0, 1, 0, 0),
    ("16->64 dgrad", 16, 0, 64, 128, 1, 1, (), 0, 0, 0, 0, 0, 0, 0),
    ("64->64 skip+up2", 64, 0, 64, 128, 1, 1, (1,), 0, 0, 0, 0, 1, 32, 96),
]
for lab, C0, C1, Co, H, istr, ostr, ups, res, msk, acc, relu, stats, kb, kt in CASES:
    K = C0 + C1
    Kt = kt if kt else K
    Hs, Ho = H * istr, H * ostr
    x0 = T(Hs, C0); x1 = T(Hs, C1) if C1 else None
    wt = (torch.randn(1, Co, Kt, device="cuda") / K ** 0.5).to(dt)
    b = torch.randn(Co, device="cuda")
    qs = [[(torch.randn(N, H >> s, H >> s, Co, device="cuda").to(dt), s) for s in ups] for _ in range(NB)]
    r = T(Ho, Co) if res else None
    m = T(Ho, Co) if msk else None
    out = T(Ho, Co)
    sts = torch.zeros(2 * Co, dtype=torch.float64, device="cuda")
    t = timeit(lambda i: lib.conv_tc2_fwd(x0[i], x1[i] if C1 else None, wt, Co, b, out[i], N, H, H, Co, taps=1, in_stride=istr,
                                          ups=qs[i], residual=r[i] if res else None, mask=m[i] if msk else None,
                                          stats=sts if stats else None, accumulate=bool(acc), relu=bool(relu), k_base=kb, k_total=kt,
                                          out_stride=ostr))
    M = N * H * H
    byt = 2 * M * (K + Co * (1 + res + msk + acc) + sum(Co / 4 ** s for s in ups))
    line = f"{lab:34s} K={K:3d} N={Co:3d} H={H:3d}  {t:7.1f} us  {byt / 1e6:6.1f} MB  {byt / t / 1e3:7.0f} GB/s"
    if PEAK:
        line += f"  {byt / t / 1e3 / PEAK:5.2f} of peak"
    print(line, flush=True)
if a.wgrad:
    for Cin, Co, H in ((32, 32, 256), (32, 8, 256), (8, 32, 256), (64, 64, 128), (64, 16, 128), (128, 128, 64)):
        x, dz = T(H, Cin), T(H, Co)
        dw = torch.zeros(Cin * Co, device="cuda")
        t = timeit(lambda i: lib.pw_wgrad_tc(x[i], dz[i], dw, Co, N, H, H, Cin, Co, 1))
        byt = 2 * N * H * H * (Cin + Co)
        print(f"pw_wgrad {Cin:3d}x{Co:3d} H={H:3d}  {t:7.1f} us  {byt / 1e6:6.1f} MB  {byt / t / 1e3:7.0f} GB/s", flush=True)
