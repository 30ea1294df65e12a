"""GPU debug: isolate maxpool_pyr_bwd mismatch per level / accumulate flag."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from emul_lib import EmulLib
from resuneta_b200 import _capi
lib = _capi.Lib(); emu = EmulLib()
st = torch.cuda.current_stream().cuda_stream
N, H, W, C = 2, 16, 16, 24
g = torch.Generator().manual_seed(1)
x = torch.randn((N, H, W, C), generator=g)
for levels in [(2,), (4,), (8,), (2, 4), (2, 4, 8)]:
    for acc in (False, True):
        dp = {k: (torch.randn((N, H // k, W // k, C), generator=g) if k in levels else None) for k in (2, 4, 8)}
        dx = torch.randn((N, H, W, C), generator=g)
        dd = {k: (v.cuda() if v is not None else None) for k, v in dp.items()}
        dxd = dx.cuda()
        emu.maxpool_pyr_bwd(x, N, H, W, C, dp[2], dp[4], dp[8], dx, acc)(0)
        lib.maxpool_pyr_bwd(x.cuda(), N, H, W, C, dd[2], dd[4], dd[8], dxd, acc)(st)
        torch.cuda.synchronize()
        err = (dxd.cpu() - dx).abs()
        bad = (err > 1e-5).nonzero()
        print(levels, acc, "maxerr", err.max().item(), "nbad", len(bad), bad[:6].tolist())
