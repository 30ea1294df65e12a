"""Where do the roles of conv_tc3 wait?  Per-CTA barrier-wait cycles (rsa_conv_tc3_set_trace) for the config-2 thin layers."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge; ge.build()
from resuneta_b200 import _capi
lib = _capi.Lib()
N, dt = 16, torch.bfloat16
st = torch.cuda.current_stream().cuda_stream
names = ["prod:A-ring", "(unused)", "mma:tempty", "mma:full", "-", "store:sready", "epi:tfull", "epi:sfree", "epi:ifull", "epi:total", "subtiles", "epi:tmem_ld", "epi:fence"]
for C, H in ((32, 256), (64, 128)):
    x = torch.randn(N, H, H, C, device="cuda").to(dt); out = torch.zeros_like(x); res = torch.randn_like(x)
    w = (torch.randn(9, C, C, device="cuda") / (3 * C ** 0.5)).to(dt).view(-1); b = torch.randn(C, device="cuda")
    stats = torch.zeros(2 * C, dtype=torch.float64, device="cuda")
    for d in (1,):
        fst = torch.rand(2 * C, dtype=torch.float64, device="cuda") * (N * H * H)
        fst[C:] = fst[:C] ** 2 / (N * H * H) + (N * H * H)
        bnr = (res, fst, float(N * H * H), 1e-3, b, b, 1)
        for label, kw in (("plain", {}), ("stats", dict(stats=stats)), ("accum", dict(accumulate=True)), ("mask", dict(mask=res)),
                          ("bnr", dict(stats=stats, bnr=bnr))):
            op = lib.conv_tc3_fwd([x], [w], [b], [d], out, N, H, H, C, **kw)
            for _ in range(3): op(st)
            torch.cuda.synchronize()
            tr = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
            lib.dll.rsa_conv_tc3_set_trace(tr.data_ptr())
            op2 = lib.conv_tc3_fwd([x], [w], [b], [d], out, N, H, H, C, **kw)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); op2(st); e1.record(); torch.cuda.synchronize()
            lib.dll.rsa_conv_tc3_set_trace(None)
            t = tr.view(148, 16).double().mean(0)
            us = e0.elapsed_time(e1) * 1e3
            print(f"C={C} d={d:2d} {label:6s} {us:6.1f} us | " + "  ".join(f"{n}={t[i].item():8.0f}" for i, n in enumerate(names) if n != "-"), flush=True)
