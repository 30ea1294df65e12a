"""Per-kernel parity through the C-ABI on the B200: every CUDA kernel against the torch-CPU emulation
of its documented semantics (tests/emul_lib.py, fp64 math) on the same seeded inputs.
Tolerances: fp32 kernels 1e-5 relative (accumulation order), bf16 storage 1 ulp of bf16 (2^-8)."""
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from emul_lib import EmulLib  # noqa: E402
from resuneta_b200._capi import Seg  # noqa: E402

EMU = EmulLib()


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from resuneta_b200 import _capi
    return _capi.Lib()


def to_dev(a):
    if isinstance(a, torch.Tensor):
        return a.cuda()
    if isinstance(a, Seg):
        s = copy.copy(a)
        s.src = a.src.cuda()
        return s
    if isinstance(a, (list, tuple)):
        return type(a)(to_dev(v) for v in a)
    return a


def run_pair(lib, method, args, kwargs, outs, rtol, atol=0.0):
    """outs: indices (into args) / keys (into kwargs) of tensors the kernel writes."""
    dargs, dkw = to_dev(args), {k: to_dev(v) for k, v in kwargs.items()}
    getattr(EMU, method)(*args, **kwargs)(0)
    getattr(lib, method)(*dargs, **dkw)(torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    for o in outs:
        ref = kwargs[o] if isinstance(o, str) else args[o]
        got = dkw[o] if isinstance(o, str) else dargs[o]
        if isinstance(ref, (list, tuple)):
            pairs = list(zip(ref, got))
        else:
            pairs = [(ref, got)]
        for r, g in pairs:
            r64, g64 = r.double(), g.cpu().double()
            err = (r64 - g64).abs().max().item()
            scale = r64.abs().max().item()
            assert err <= rtol * scale + atol, (method, o, err, scale)


def rnd(shape, dtype, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dtype)


DT = [torch.float32, torch.bfloat16]
TOL = {torch.float32: 2e-5, torch.bfloat16: 1.0 / 128}


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("C,Co,H,d", [(32, 32, 24, 1), (32, 32, 40, 15), (64, 64, 16, 3), (32, 32, 32, 31), (8, 12, 9, 1)])
def test_igemm_conv3x3_fwd_epilogues(lib, dt, C, Co, H, d):
    N, W = 2, H + 3
    x = rnd((N, H, W, C), dt, 1)
    w = rnd((9 * C * Co,), torch.float32, 2, 0.1)
    b = rnd((Co,), torch.float32, 3)
    segs = [Seg(x, C, H, W, off_h=(ky - 1) * d, off_w=(kx - 1) * d, w_off=(ky * 3 + kx) * C * Co)
            for ky in range(3) for kx in range(3)]
    out = rnd((N, H, W, Co), dt, 4)
    res = rnd((N, H, W, Co), dt, 5)
    stats = torch.zeros(2 * Co, dtype=torch.float64)
    run_pair(lib, "igemm_fwd", [segs, w, Co, False, b, out, N, H, W, Co],
             dict(residual=res, stats=stats, accumulate=True, relu=False), [5, "stats"], TOL[dt], 1e-4)
    mask = rnd((N, H, W, Co), dt, 6)
    run_pair(lib, "igemm_fwd", [segs, w, Co, False, b, out, N, H, W, Co], dict(relu=True, mask=mask), [5], TOL[dt], 1e-4)


@pytest.mark.parametrize("dt", DT)
def test_igemm_dgrad_transposed_weights(lib, dt):
    N, H, W, C, Co, d = 2, 20, 20, 32, 64, 3
    dy = rnd((N, H, W, Co), dt, 1)
    w = rnd((9 * C * Co,), torch.float32, 2, 0.1)
    sg = [Seg(dy, Co, H, W, off_h=-(ky - 1) * d, off_w=-(kx - 1) * d, w_off=(ky * 3 + kx) * C * Co)
          for ky in range(3) for kx in range(3)]
    dx = rnd((N, H, W, C), dt, 3)
    run_pair(lib, "igemm_fwd", [sg, w, Co, True, None, dx, N, H, W, C], dict(accumulate=True), [5], TOL[dt], 1e-4)


@pytest.mark.parametrize("dt", DT)
def test_igemm_gather_modes_concat_up_stride(lib, dt):
    N = 2
    a = rnd((N, 8, 8, 16), dt, 1)      # up-sampled x2, ReLU'd
    b = rnd((N, 16, 16, 24), dt, 2)    # plain
    c = rnd((N, 2, 2, 8), dt, 3)       # up-sampled x8
    Co = 20
    w = rnd(((16 + 24 + 8) * Co,), torch.float32, 4, 0.2)
    bias = rnd((Co,), torch.float32, 5)
    segs = [Seg(a, 16, 8, 8, shift=1, relu_in=True, w_off=0), Seg(b, 24, 16, 16, w_off=16 * Co),
            Seg(c, 8, 2, 2, shift=3, w_off=40 * Co)]
    out = torch.zeros((N, 16, 16, Co), dtype=torch.float32)
    stats = torch.zeros(2 * Co, dtype=torch.float64)
    run_pair(lib, "igemm_fwd", [segs, w, Co, False, bias, out, N, 16, 16, Co], dict(stats=stats), [5, "stats"], 2e-5, 1e-5)
    # stride-2 sampling (down conv) and its transposed gather
    x = rnd((N, 16, 16, 32), dt, 6)
    w2 = rnd((32 * 64,), torch.float32, 7, 0.2)
    o2 = torch.zeros((N, 8, 8, 64), dtype=dt)
    run_pair(lib, "igemm_fwd", [[Seg(x, 32, 16, 16, mult=2)], w2, 64, False, None, o2, N, 8, 8, 64], {}, [5], TOL[dt], 1e-4)
    dy = rnd((N, 8, 8, 64), dt, 8)
    dx = rnd((N, 16, 16, 32), dt, 9)
    run_pair(lib, "igemm_fwd", [[Seg(dy, 64, 8, 8, shift=1, aligned=True)], w2, 64, True, None, dx, N, 16, 16, 32],
             dict(accumulate=True), [5], TOL[dt], 1e-4)
    # adjoint of x2 up-sampling as four child segments
    dyf = rnd((N, 16, 16, Co), dt, 10)
    da = torch.zeros((N, 8, 8, 16), dtype=dt)
    sg = [Seg(dyf, Co, 16, 16, mult=2, off_h=i, off_w=j, w_off=0) for i in (0, 1) for j in (0, 1)]
    run_pair(lib, "igemm_fwd", [sg, w, Co, True, None, da, N, 8, 8, 16], dict(mask=a), [5], TOL[dt], 1e-4)


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("C,Co", [(32, 32), (3, 32), (96, 6)])
def test_igemm_wgrad(lib, dt, C, Co):
    N, H, W, d = 2, 24, 24, 3
    x = rnd((N, H, W, C), dt, 1)
    dy = rnd((N, H, W, Co), dt, 2)
    segs = [Seg(x, C, H, W, off_h=(ky - 1) * d, off_w=(kx - 1) * d, relu_in=(ky == 1), w_off=(ky * 3 + kx) * C * Co)
            for ky in range(3) for kx in range(3)]
    dw = torch.zeros(9 * C * Co, dtype=torch.float32)
    db = torch.zeros(Co, dtype=torch.float32)
    run_pair(lib, "igemm_wgrad", [segs, dy, dw, Co, db, N, H, W, Co], {}, [2, 4], 1e-4, 1e-4)
    dyf = dy.float()
    run_pair(lib, "igemm_wgrad", [segs, dyf, dw, Co, None, N, H, W, Co], {}, [2], 1e-4, 1e-4)


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("C", [8, 32, 256, 1024])
def test_bn_forward_backward(lib, dt, C):
    M = 1500
    x = rnd((M, C), dt, 1) * 2 + 0.5
    stats = torch.zeros(2 * C, dtype=torch.float64)
    run_pair(lib, "bn_stats", [x, M, C, stats], {}, [3], 1e-5, 1e-6)
    EMU.bn_stats(x, M, C, stats)(0)
    gam = [rnd((C,), torch.float32, 10 + k) for k in range(3)]
    bet = [rnd((C,), torch.float32, 20 + k) for k in range(3)]
    outs = [torch.zeros((M, C), dtype=dt) for _ in range(3)]
    run_pair(lib, "bn_apply", [x, M, C, outs, gam, bet, stats, float(M), None, None, 1e-3, True], {}, [3], TOL[dt], 1e-5)
    mm = [rnd((C,), torch.float32, 30)]
    mv = [rnd((C,), torch.float32, 31).abs() + 0.5]
    run_pair(lib, "bn_apply", [x, M, C, outs[:1], gam[:1], bet[:1], None, 1.0, mm, mv, 1e-3, False], {}, [3], TOL[dt], 1e-5)
    dy = rnd((M, C), dt, 40)
    act = rnd((M, C), dt, 41)
    red = torch.zeros(2 * C, dtype=torch.float64)
    run_pair(lib, "bn_bwd_reduce", [dy, x, act, M, C, stats, float(M), 1e-3, red], {}, [8], 1e-4, 1e-4)
    EMU.bn_bwd_reduce(dy, x, act, M, C, stats, float(M), 1e-3, red)(0)
    dx = rnd((M, C), dt, 42)
    dg = torch.zeros(C)
    db = torch.zeros(C)
    run_pair(lib, "bn_bwd_apply", [dy, x, act, M, C, stats, float(M), 1e-3, gam[0], red, dx, True, dg, db], {},
             [10, 12, 13], TOL[dt], 1e-4)
    dst = torch.zeros(2 * C, dtype=torch.float64)
    run_pair(lib, "bn_derive_stats", [stats, float(M), gam[0], bet[0], 1e-3, dst, 4.0 * M, C], {}, [5], 1e-6, 1e-9)


def test_bn_update_moving(lib):
    Cs = [32, 8, 256]
    stats = torch.rand(2 * sum(Cs), dtype=torch.float64) * 100 + 50
    params = torch.rand(2 * sum(Cs) + 64)
    tab, cnt, so, po = [], [], 0, 0
    for C in Cs:
        stats[so + C:so + 2 * C] += stats[so:so + C] ** 2 / 100   # keep variance positive
        tab += [so, C, po, po + C]
        cnt += [100.0, 400.0]
        so += 2 * C
        po += 2 * C
    table = torch.tensor(tab, dtype=torch.int64)
    counts = torch.tensor(cnt, dtype=torch.float64)
    run_pair(lib, "bn_update_moving", [stats, params, table, counts, len(Cs), 0.99], {}, [1], 1e-6, 1e-7)


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("levels,H,C", [((2, 4, 8), 16, 24), ((2, 4), 8, 24), ((2,), 6, 24),
                                        # 32 channels, 8-pixel windows: the warp-per-window backward kernel (bf16)
                                        ((2, 4, 8), 16, 32), ((2, 4, 8), 40, 32), ((4, 8), 16, 32)])
def test_pool_pyramids(lib, dt, levels, H, C):
    N, W = 2, H
    x = rnd((N, H, W, C), dt, 1)
    p = {k: (torch.zeros((N, H // k, W // k, C), dtype=dt) if k in levels else None) for k in (2, 4, 8)}
    run_pair(lib, "maxpool_pyr_fwd", [x, N, H, W, C, p[2], p[4], p[8]], {}, [i for i, k in ((5, 2), (6, 4), (7, 8)) if k in levels], 0.0)
    dp = {k: (rnd((N, H // k, W // k, C), dt, 10 + k) if k in levels else None) for k in (2, 4, 8)}
    dx = rnd((N, H, W, C), dt, 3)
    run_pair(lib, "maxpool_pyr_bwd", [x, N, H, W, C, dp[2], dp[4], dp[8], dx, True], {}, [8], TOL[dt], 1e-6)
    # ties: quantised input, gradient must go to the FIRST maximum of each window
    xq = (x.float() * 2).round().to(dt)
    dx0 = torch.zeros((N, H, W, C), dtype=dt)
    run_pair(lib, "maxpool_pyr_bwd", [xq, N, H, W, C, dp[2], dp[4], dp[8], dx0, False], {}, [8], TOL[dt], 1e-6)
    s = {k: (torch.zeros((N, H // k, W // k, C), dtype=dt) if k in levels else None) for k in (2, 4, 8)}
    run_pair(lib, "sumpool_pyr", [x, N, H, W, C, s[2], s[4], s[8]], {}, [i for i, k in ((5, 2), (6, 4), (7, 8)) if k in levels], TOL[dt], 1e-5)


@pytest.mark.parametrize("C", [3, 6, 12])
def test_head_activations(lib, C):
    M = 5000
    z = rnd((M, C), torch.float32, 1, 3.0)
    p = torch.zeros((M, C))
    run_pair(lib, "softmax_fwd", [z, p, M, C], {}, [1], 1e-6, 1e-7)
    EMU.softmax_fwd(z, p, M, C)(0)
    dp = rnd((M, C), torch.float32, 2)
    dz = torch.zeros((M, C))
    run_pair(lib, "softmax_bwd", [p, dp, dz, M, C], {}, [2], 1e-5, 1e-7)
    run_pair(lib, "sigmoid_fwd", [z, p, M * C], {}, [1], 1e-6, 1e-7)
    run_pair(lib, "sigmoid_bwd", [p, dp, dz, M * C], {}, [2], 1e-5, 1e-7)


@pytest.mark.parametrize("C,single_class", [(6, False), (3, False), (6, True), (12, False)])
def test_tanimoto_forward_backward(lib, C, single_class):
    B, H = 3, 40
    g = torch.Generator().manual_seed(C)
    pred = torch.softmax(torch.randn((B, H, H, C), generator=g) * 2, -1)
    lab = torch.nn.functional.one_hot(torch.randint(0, C, (B, H, H), generator=g), C).float()
    if single_class:     # classes absent from the labels: inf -> max-weight path (multitasking_utils.py:52-53)
        lab = torch.zeros_like(lab)
        lab[..., 0] = 1.0
    sums = torch.zeros(B * C * 5, dtype=torch.float64)
    run_pair(lib, "tanimoto_sums", [pred, lab, B, H * H, C, sums], {}, [5], 1e-6, 1e-6)
    EMU.tanimoto_sums(pred, lab, B, H * H, C, sums)(0)
    lb, lm, coef = torch.zeros(B), torch.zeros(1), torch.zeros(B * C * 3)
    run_pair(lib, "tanimoto_finalize", [sums, B, H * H, C, 0.7, lb, lm, coef], {}, [5, 6, 7], 1e-5, 1e-9)
    EMU.tanimoto_finalize(sums, B, H * H, C, 0.7, lb, lm, coef)(0)
    dpred = torch.zeros_like(pred)
    run_pair(lib, "tanimoto_bwd", [pred, lab, coef, B, H * H, C, dpred], {}, [6], 1e-5, 1e-12)


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_pixel_losses(lib, kind):
    M, C = 7000, 5
    g = torch.Generator().manual_seed(kind)
    pred = torch.softmax(torch.randn((M, C), generator=g) * 4, -1)
    pred[:10] = torch.tensor([1.0, 0, 0, 0, 0])           # exercises the 1e-7 clip
    lab = torch.nn.functional.one_hot(torch.randint(0, C, (M,), generator=g), C).float()
    w = torch.tensor([1.1, 2.0, 0.5, 3.0, 0.0]) if kind == 0 else None
    ls = torch.zeros(1, dtype=torch.float64)
    run_pair(lib, "pixel_loss_fwd", [kind, pred, lab, w, M, C, ls], {}, [6], 1e-4)   # fp32 logf per pixel
    dp = torch.zeros_like(pred)
    run_pair(lib, "pixel_loss_bwd", [kind, pred, lab, w, M, C, 0.37, dp], {}, [7], 1e-5, 1e-7)


def test_metrics_and_confusion_are_bit_exact(lib):
    M, C = 100000, 6
    g = torch.Generator().manual_seed(0)
    prob = torch.softmax(torch.randn((M, C), generator=g) * 3, -1)
    prob[:100] = 1.0 / C                                   # ties -> first maximum like numpy argmax
    lab = torch.randint(0, C, (M,), generator=g)
    onehot = torch.nn.functional.one_hot(lab, C).float()
    out = torch.zeros(5, dtype=torch.int64)
    run_pair(lib, "seg_metrics", [prob, onehot, M, C, out], {}, [4], 0.0)
    pl, cm = torch.zeros(M, dtype=torch.int32), torch.zeros(C * C, dtype=torch.int64)
    run_pair(lib, "argmax_confusion", [prob, M, C, pl, lab.to(torch.int32), C, cm], {}, [3, 6], 0.0)
    from sklearn.metrics import confusion_matrix
    d_pl, d_cm = torch.zeros(M, dtype=torch.int32).cuda(), torch.zeros(C * C, dtype=torch.int64).cuda()
    lib.argmax_confusion(prob.cuda(), M, C, d_pl, lab.to(torch.int32).cuda(), C, d_cm)(torch.cuda.current_stream().cuda_stream)
    np.testing.assert_array_equal(d_pl.cpu().numpy(), prob.numpy().argmax(-1))
    np.testing.assert_array_equal(d_cm.cpu().numpy().reshape(C, C), confusion_matrix(lab.numpy(), prob.numpy().argmax(-1)))


def test_optimizers_axpy_cast(lib):
    n = 100003
    p, g = rnd((n,), torch.float32, 1), rnd((n,), torch.float32, 2, 1e-3)
    m, v = rnd((n,), torch.float32, 3, 1e-4), rnd((n,), torch.float32, 4, 1e-4).abs()
    lr = torch.tensor([3e-4])
    run_pair(lib, "adam_step", [p, g, m, v, n, lr, 0.9, 0.999, 1e-7, 0.5], {}, [0, 2, 3], 1e-6, 1e-9)
    run_pair(lib, "sgd_step", [p, g, m, n, lr, 0.8, 1.0], {}, [0, 2], 1e-6, 1e-9)
    for dt in DT:
        a, b = rnd((n,), dt, 5), rnd((n,), dt, 6)
        run_pair(lib, "axpy", [a, b, n, True], {}, [0], TOL[dt])
        run_pair(lib, "axpy", [a, b, n, False], {}, [0], 0.0)
    src = rnd((n,), torch.float32, 7)
    dst = torch.zeros(n, dtype=torch.bfloat16)
    run_pair(lib, "cast", [src, dst, n], {}, [1], 0.0)


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("C,k,relu", [(32, 4, True), (8, 1, False), (256, 3, True), (1024, 1, True), (64, 2, False)])
def test_bn_backward_multi_branch(lib, dt, C, k, relu):
    M = 1300
    x = rnd((M, C), dt, 1) * 2 + 0.5
    stats = torch.zeros(2 * C, dtype=torch.float64)
    EMU.bn_stats(x, M, C, stats)(0)
    dys = [rnd((M, C), dt, 10 + b) for b in range(k)]
    gam = [rnd((C,), torch.float32, 20 + b) for b in range(k)]
    bet = [rnd((C,), torch.float32, 30 + b) for b in range(k)]
    reds = [torch.zeros(2 * C, dtype=torch.float64) for _ in range(k)]
    run_pair(lib, "bn_bwd_reduce_multi", [dys, x, M, C, stats, float(M), 1e-3, gam, bet, relu, reds], {}, [10], 2e-3 if dt == torch.bfloat16 else 1e-4, 3.0)   # atol: one borderline ReLU-mask flip moves a sum by |dy|
    for r in reds:
        r.zero_()
    EMU.bn_bwd_reduce_multi(dys, x, M, C, stats, float(M), 1e-3, gam, bet, relu, reds)(0)
    dx = rnd((M, C), dt, 40)
    dg = [torch.zeros(C) for _ in range(k)]
    db = [torch.zeros(C) for _ in range(k)]
    run_pair(lib, "bn_bwd_apply_multi", [dys, x, M, C, stats, float(M), 1e-3, gam, bet, relu, reds, dx, True, dg, db], {},
             [11, 13, 14], TOL[dt] * 4, 1e-2)


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("n", [3, 14])
def test_stem_thin_kernels(lib, dt, n):
    M = 40000
    x = rnd((M, n), dt, 1)
    w = rnd((n * 32,), torch.float32, 2, 0.3)
    b = rnd((32,), torch.float32, 3)
    out = torch.zeros((M, 32), dtype=dt)
    stats = torch.zeros(64, dtype=torch.float64)
    run_pair(lib, "stem_fwd", [x, w, b, out, M, n, stats], {}, [3, 6], TOL[dt], 1e-2)
    dy = rnd((M, 32), dt, 4)
    dw, db = torch.zeros(n * 32), torch.zeros(32)
    run_pair(lib, "stem_wgrad", [x, dy, M, n, dw, db], {}, [4, 5], 1e-4, 1e-3)


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("n,acc,mask", [(6, False, True), (3, True, False), (6, True, True)])
def test_head_bwd_thin_kernel(lib, dt, n, acc, mask):
    M = 30000
    h = rnd((M, 32), dt, 1)
    dz = rnd((M, n), torch.float32, 2)
    w = rnd((32 * n,), torch.float32, 3, 0.3)
    dh = rnd((M, 32), dt, 4)
    dw, db = torch.zeros(32 * n), torch.zeros(n)
    run_pair(lib, "head_bwd", [h, dz, w, M, n, dh, acc, mask, dw, db], {}, [5, 8, 9], max(TOL[dt], 1e-4), 1e-3)


@pytest.mark.parametrize("n", [6, 3, 8])
def test_head_fwd_thin_kernel(lib, n):
    """bf16 features -> fp32 logits (model2.py:159,168,180,186) against the fp64 product of the same operands."""
    M = 30011
    h = rnd((M, 32), torch.bfloat16, 1)
    w = rnd((32 * n,), torch.float32, 2, 0.3)
    b = rnd((n,), torch.float32, 3)
    z = torch.zeros((M, n), dtype=torch.float32).cuda()
    lib.head_fwd(h.cuda(), w.cuda(), b.cuda(), z, M, n)(torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ref = h.double() @ w.view(32, n).double() + b.double()
    assert (z.cpu().double() - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()
