"""tcgen05/TMA convolution kernels (conv_tc.cu, conv_tc2.cu, conv_tc3.cu) through the C-ABI against the fp64 emulation of the same
implicit GEMM on bf16-rounded operands.  Kept in its own file so that it runs in its own process on the
GPU box.  Tolerance: 1e-2 of the output range (bf16 output rounding + fp32 accumulation order)."""
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


def rnd(shape, dtype, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dtype)


# ---- tensor-core (tcgen05 + TMA) convolution vs the fp64 emulation of the same implicit GEMM --------------
def _pack(w_hwio_flat, taps, cin, cout):
    w = w_hwio_flat.view(taps, cin, cout)
    return w.permute(0, 2, 1).contiguous().to(torch.bfloat16), w.contiguous().to(torch.bfloat16)


def test_pack_weights_tc(lib):
    import struct
    layers = [(9, 32, 32), (9, 64, 128)]
    total = sum(t * a * b for t, a, b in layers)
    params = rnd((total + 64,), torch.float32, 1)
    table, soff, doff = b"", 0, 0
    for t, a, b in layers:
        n = t * a * b
        table += struct.pack("<qqqiiii", soff, doff, doff + n, t, a, b, 0)
        soff += n
        doff += 2 * n
    shadow = torch.zeros(doff, dtype=torch.bfloat16).cuda()
    tab = torch.frombuffer(bytearray(table), dtype=torch.uint8).cuda()
    lib.pack_weights_tc(params.cuda(), shadow, tab, len(layers), max(t * a * b for t, a, b in layers))(
        torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    soff = doff = 0
    for t, a, b in layers:
        n = t * a * b
        wf, wb = _pack(params[soff:soff + n], t, a, b)
        assert torch.equal(shadow[doff:doff + n].cpu(), wf.reshape(-1))
        assert torch.equal(shadow[doff + n:doff + 2 * n].cpu(), wb.reshape(-1))
        soff += n
        doff += 2 * n


@pytest.mark.parametrize("N,H,C,d", [(2, 32, 32, 1), (2, 64, 32, 31), (2, 32, 64, 3), (3, 16, 128, 1), (2, 16, 256, 15),
                                     (4, 8, 512, 1), (16, 4, 1024, 1), (16, 64, 32, 15)])
def test_conv_tc_wgrad_and_bias_grad(lib, N, H, C, d):
    W, dt = H, torch.bfloat16
    x = rnd((N, H, W, C), dt, 1)
    dy = rnd((N, H, W, C), dt, 2)
    segs = [Seg(x, C, H, W, off_h=(ky - 1) * d, off_w=(kx - 1) * d, w_off=(ky * 3 + kx) * C * C)
            for ky in range(3) for kx in range(3)]
    dw = torch.zeros(9 * C * C, dtype=torch.float32)
    db = torch.zeros(C, dtype=torch.float32)
    EMU.igemm_wgrad(segs, dy, dw, C, db, N, H, W, C)(0)
    st = torch.cuda.current_stream().cuda_stream
    d_dw = torch.zeros(9 * C * C, dtype=torch.float32).cuda()
    lib.conv_tc_wgrad(x.cuda(), dy.cuda(), d_dw, N, H, W, C, C, d)(st)
    d_db = [torch.zeros(C, dtype=torch.float32).cuda() for _ in range(3)]
    lib.bias_grad(dy.cuda(), N * H * W, C, d_db)(st)
    torch.cuda.synchronize()
    scale = dw.abs().max().item()
    err = (d_dw.cpu() - dw).abs().max().item()
    assert err <= 2e-3 * scale, (err, scale)
    for t_ in d_db:
        np.testing.assert_allclose(t_.cpu().numpy(), db.numpy(), rtol=1e-3, atol=1e-3 * db.abs().max().item())


# ---- persistent generalised kernel (conv_tc2.cu) ------------------------------------------------------------
def _up(t, s):
    k = 1 << s
    return t.repeat_interleave(k, 1).repeat_interleave(k, 2)


@pytest.mark.parametrize("N,H,C,Co,d", [(2, 32, 32, 32, 1), (2, 64, 32, 32, 31), (2, 32, 64, 64, 3), (3, 16, 128, 128, 1),
                                        (2, 16, 256, 256, 15), (4, 8, 512, 512, 1), (16, 4, 1024, 1024, 1),
                                        (16, 64, 32, 32, 15), (5, 32, 64, 128, 3),
                                        # >= 120 tiles of 256 pixels: two sub-tiles per CTA share the weight slices (MT = 2)
                                        (8, 64, 128, 128, 1), (8, 64, 128, 128, 15), (16, 32, 256, 256, 3)])
def test_conv_tc2_3x3_matches_emulation(lib, N, H, C, Co, d):
    W, dt = H, torch.bfloat16
    x = rnd((N, H, W, C), dt, 1)
    w = rnd((9 * C * Co,), torch.float32, 2, 1.0 / (3 * C ** 0.5)).to(dt).float()
    b = rnd((Co,), torch.float32, 3)
    wf, _ = _pack(w, 9, C, Co)
    segs = [Seg(x, C, H, W, off_h=(ky - 1) * d, off_w=(kx - 1) * d, w_off=(ky * 3 + kx) * C * Co)
            for ky in range(3) for kx in range(3)]
    out, res = rnd((N, H, W, Co), dt, 4), rnd((N, H, W, Co), dt, 5)
    stats = torch.zeros(2 * Co, dtype=torch.float64)
    EMU.igemm_fwd(segs, w, Co, False, b, out, N, H, W, Co, residual=res, stats=stats, accumulate=True)(0)
    st = torch.cuda.current_stream().cuda_stream
    d_out, d_stats = rnd((N, H, W, Co), dt, 4).cuda(), torch.zeros(2 * Co, dtype=torch.float64).cuda()
    lib.conv_tc2_fwd(x.cuda(), None, wf.cuda(), Co, b.cuda(), d_out, N, H, W, Co, taps=9, dil=d, residual=res.cuda(),
                     stats=d_stats, accumulate=True)(st)
    torch.cuda.synchronize()
    scale = out.float().abs().max().item()
    assert (d_out.cpu().float() - out.float()).abs().max().item() <= scale / 100
    np.testing.assert_allclose(d_stats.cpu().numpy(), stats.numpy(), rtol=2e-2, atol=2e-2 * N * H * W ** 0.5)
    mask = rnd((N, H, W, Co), dt, 6)
    EMU.igemm_fwd(segs, w, Co, False, b, out, N, H, W, Co, relu=True, mask=mask)(0)
    lib.conv_tc2_fwd(x.cuda(), None, wf.cuda(), Co, b.cuda(), d_out, N, H, W, Co, taps=9, dil=d, relu=True,
                     mask=mask.cuda())(st)
    torch.cuda.synchronize()
    assert (d_out.cpu().float() - out.float()).abs().max().item() <= scale / 100


@pytest.mark.parametrize("N,H,C0,C1,Co,stride,ups,f32", [
    (2, 64, 32, 32, 32, 1, (), False),            # final combine: two plain sources
    (2, 64, 32, 0, 64, 2, (), False),             # down conv: stride 2
    (2, 32, 64, 0, 64, 1, (1,), False),           # decoder combine: skip + up2(q)
    (2, 64, 32, 0, 32, 1, (1, 2, 3), False),      # PSP final conv: x + up2/4/8 addends
    (2, 32, 32, 0, 6, 1, (), True),               # head: 6 classes, fp32 logits, padded N
    (2, 32, 64, 0, 16, 1, (), False),             # dec1 up conv: 64 -> 16
    (2, 32, 16, 0, 32, 1, (), False),             # K = 16 source (SWIZZLE_32B)
    (3, 8, 1024, 0, 256, 1, (), False),           # PSP mid branch
    (2, 16, 256, 512, 512, 1, (), False),         # wide two-source concat
    (2, 32, 32, 0, 8, 1, (), False),              # PSP out branch: 8 output channels
    (2, 32, 8, 0, 32, 1, (), False),              # 8-channel source: 16-wide TMA box, upper half zero-filled
    (2, 64, 8, 0, 32, 1, (), False)])
def test_conv_tc2_pointwise_modes(lib, N, H, C0, C1, Co, stride, ups, f32):
    dt = torch.bfloat16
    Hs = H * stride
    x0 = rnd((N, Hs, Hs, C0), dt, 1)
    x1 = rnd((N, Hs, Hs, C1), dt, 2) if C1 else None
    K = C0 + C1
    w = rnd((K * Co,), torch.float32, 3, 1.0 / K ** 0.5).to(dt).float()       # [K][Co]
    b = rnd((Co,), torch.float32, 4)
    BNt = 128 if Co >= 128 else (64 if Co >= 64 else (32 if Co >= 32 else 16))
    CoP = (Co + BNt - 1) // BNt * BNt
    wt = torch.zeros((1, CoP, K), dtype=dt)
    wt[0, :Co] = w.view(K, Co).t().to(dt)
    segs = [Seg(x0, C0, Hs, Hs, mult=stride, w_off=0)]
    if C1:
        segs.append(Seg(x1, C1, Hs, Hs, mult=stride, w_off=C0 * Co))
    qs = [rnd((N, H >> s, H >> s, Co), dt, 10 + s) for s in ups]
    res = sum((_up(q.float(), s) for q, s in zip(qs, ups)), torch.zeros((N, H, H, Co))) if ups else None
    out = torch.zeros((N, H, H, Co), dtype=torch.float32 if f32 else dt)
    stats = torch.zeros(2 * Co, dtype=torch.float64)
    EMU.igemm_fwd(segs, w, Co, False, b, out, N, H, H, Co, residual=res, stats=stats)(0)
    st = torch.cuda.current_stream().cuda_stream
    d_out = torch.zeros_like(out).cuda()
    d_stats = torch.zeros(2 * Co, dtype=torch.float64).cuda()
    lib.conv_tc2_fwd(x0.cuda(), x1.cuda() if C1 else None, wt.cuda(), CoP, b.cuda(), d_out, N, H, H, Co, taps=1,
                     in_stride=stride, ups=[(q.cuda(), s) for q, s in zip(qs, ups)], stats=d_stats)(st)
    torch.cuda.synchronize()
    scale = out.float().abs().max().item()
    assert (d_out.cpu().float() - out.float()).abs().max().item() <= scale / 100
    np.testing.assert_allclose(d_stats.cpu().numpy(), stats.numpy(), rtol=2e-2, atol=2e-2 * N * H * H ** 0.5)


# every channel combination the streaming 1x1 kernel (pw_stream.cu, entered through rsa_conv_tc2_fwd) is compiled for, with the
# epilogue options the plan uses on them: (C0, C1, Co, in_stride, out_stride, ups, residual, mask, accumulate, relu, stats, k_base, k_total)
_PWS = [
    (32, 0, 32, 1, 1, (), False, True, True, False, False, 0, 0),        # data gradient into a ReLU-masked, shared tensor
    (32, 0, 32, 1, 1, (), False, False, False, False, True, 0, 0),       # plain + statistics
    (32, 32, 32, 1, 1, (), False, False, False, True, True, 0, 0),       # combine: concat, ReLU, statistics
    (32, 0, 32, 1, 1, (1, 2, 3), True, False, False, False, True, 32, 64),   # PSP output: three up-sampled addends + residual
    (32, 0, 32, 1, 1, (1,), False, False, False, False, True, 16, 48),   # decoder combine: skip + up2(q)
    (32, 0, 8, 1, 1, (), False, False, False, False, False, 0, 0),
    (8, 0, 32, 1, 1, (), False, True, True, False, False, 0, 0),
    (8, 0, 32, 1, 1, (), False, False, False, False, False, 8, 64),
    (32, 0, 64, 2, 1, (), False, False, False, False, True, 0, 0),       # down convolution: stride 2
    (64, 0, 32, 1, 2, (), False, True, True, False, False, 0, 0),        # its data gradient: strided accumulate
    (64, 0, 16, 1, 1, (), False, False, False, True, False, 0, 0),
    (16, 0, 64, 1, 1, (), False, True, False, False, False, 0, 0),
    (16, 0, 32, 1, 1, (), False, False, False, False, True, 0, 48),
    (16, 0, 16, 1, 1, (), True, False, False, False, False, 0, 0),
    (8, 0, 8, 1, 1, (), False, False, True, False, True, 0, 0),
    (8, 0, 16, 1, 1, (), False, False, False, False, False, 0, 0),
    (8, 0, 64, 1, 1, (), False, False, False, False, True, 0, 0),
    (16, 0, 8, 1, 1, (), False, False, False, False, False, 0, 0),
    (32, 0, 16, 1, 1, (2,), False, False, False, False, False, 0, 0),
    (64, 0, 8, 1, 1, (), False, False, False, False, True, 0, 0),
    (64, 0, 32, 1, 1, (), False, False, False, False, True, 0, 0),
    (32, 32, 16, 1, 1, (), False, False, False, False, False, 0, 0),
    (32, 32, 8, 1, 1, (), False, True, False, False, False, 0, 0),
    (64, 0, 64, 1, 1, (1,), False, False, False, False, True, 32, 96),   # 64 -> 64 (B fragments in shared memory): skip + up2(q)
    (64, 0, 64, 1, 1, (), False, True, True, False, False, 0, 0),
    (32, 32, 64, 1, 1, (), True, False, False, True, True, 0, 0),
]


@pytest.mark.parametrize("C0,C1,Co,istr,ostr,ups,res,msk,acc,relu,stats,kb,kt", _PWS)
@pytest.mark.parametrize("N,H", [(3, 32), (1, 4)])
def test_pw_stream_modes(lib, N, H, C0, C1, Co, istr, ostr, ups, res, msk, acc, relu, stats, kb, kt):
    from emul_lib import EmulLibTC
    if ups and (H >> max(ups)) < 1:
        pytest.skip("addend resolution below one pixel")
    emu, dt = EmulLibTC(), torch.bfloat16
    K, Hs, Ho = C0 + C1, H * istr, H * ostr
    Kt = kt if kt else K
    x0 = rnd((N, Hs, Hs, C0), dt, 1)
    x1 = rnd((N, Hs, Hs, C1), dt, 2) if C1 else None
    wt = rnd((1, Co, Kt), torch.float32, 3, 1.0 / K ** 0.5).to(dt)
    b = rnd((Co,), torch.float32, 4)
    qs = [(rnd((N, H >> s, H >> s, Co), dt, 10 + s), s) for s in ups]
    r = rnd((N, Ho, Ho, Co), dt, 5) if res else None
    m = rnd((N, Ho, Ho, Co), dt, 6) if msk else None
    out0 = rnd((N, Ho, Ho, Co), dt, 7)
    ref, st_ref = out0.clone(), torch.zeros(2 * Co, dtype=torch.float64)
    kw = dict(taps=1, in_stride=istr, residual=r, mask=m, accumulate=acc, relu=relu, k_base=kb, k_total=kt, out_stride=ostr)
    emu.conv_tc2_fwd(x0, x1, wt, Co, b, ref, N, H, H, Co, ups=qs, stats=st_ref if stats else None, **kw)(0)
    d_out, d_st = out0.clone().cuda(), torch.zeros(2 * Co, dtype=torch.float64).cuda()
    cu = lambda v: None if v is None else v.cuda()
    kw.update(residual=cu(r), mask=cu(m))
    lib.conv_tc2_fwd(x0.cuda(), cu(x1), wt.cuda(), Co, b.cuda(), d_out, N, H, H, Co, ups=[(q.cuda(), s) for q, s in qs],
                     stats=d_st if stats else None, **kw)(torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    scale = ref.float().abs().max().item()
    assert (d_out.cpu().float() - ref.float()).abs().max().item() <= scale / 100
    if ostr == 2:      # pixels between the strided stores stay untouched
        keep = torch.ones(Ho, Ho, dtype=torch.bool); keep[::2, ::2] = False
        assert torch.equal(d_out.cpu()[:, keep], out0[:, keep])
    if stats:
        np.testing.assert_allclose(d_st.cpu().numpy(), st_ref.numpy(), rtol=2e-2, atol=2e-2 * N * H * H ** 0.5)


@pytest.mark.parametrize("N,H,Cin,Cout,stride", [(2, 64, 32, 32, 1), (2, 32, 32, 64, 2), (2, 32, 16, 32, 1), (2, 32, 64, 16, 1),
                                                  (3, 16, 256, 512, 1), (2, 8, 1024, 256, 1), (16, 4, 512, 1024, 2),
                                                  (4, 64, 64, 128, 1), (2, 64, 128, 32, 1), (2, 64, 32, 8, 1), (2, 64, 8, 32, 1),
                                                  (2, 32, 8, 8, 1)])
def test_pw_wgrad_tc(lib, N, H, Cin, Cout, stride):
    dt = torch.bfloat16
    Hs = H * stride
    x = rnd((N, Hs, Hs, Cin), dt, 1)
    dz = rnd((N, H, H, Cout), dt, 2)
    ldw = Cout + 16                     # a row-slice of a wider [K_total, ldw] gradient matrix
    dw = torch.zeros(Cin * ldw, dtype=torch.float32)
    EMU.igemm_wgrad([Seg(x, Cin, Hs, Hs, mult=stride, w_off=0)], dz, dw, ldw, None, N, H, H, Cout)(0)
    d_dw = torch.zeros(Cin * ldw, dtype=torch.float32).cuda()
    lib.pw_wgrad_tc(x.cuda(), dz.cuda(), d_dw, ldw, N, H, H, Cin, Cout, stride)(torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    scale = dw.abs().max().item()
    assert (d_dw.cpu() - dw).abs().max().item() <= 2e-3 * scale


def test_conv_tc2_k_base_and_strided_store(lib):
    """A K sub-range of a wider weight matrix (q-convs of the decoder) and the transposed stride-2 store."""
    dt = torch.bfloat16
    N, H, C, Co, Kt, kb = 2, 32, 32, 64, 96, 64
    x = rnd((N, H, H, C), dt, 1)
    wfull = rnd((Kt * Co,), torch.float32, 2, 0.1).to(dt).float()          # [Kt][Co]
    wt = wfull.view(Kt, Co).t().contiguous().to(dt).view(1, Co, Kt)
    out = torch.zeros((N, H, H, Co), dtype=dt)
    EMU.igemm_fwd([Seg(x, C, H, H, w_off=kb * Co)], wfull, Co, False, None, out, N, H, H, Co)(0)
    st = torch.cuda.current_stream().cuda_stream
    d_out = torch.zeros_like(out).cuda()
    lib.conv_tc2_fwd(x.cuda(), None, wt.cuda(), Co, None, d_out, N, H, H, Co, k_base=kb, k_total=Kt)(st)
    torch.cuda.synchronize()
    assert (d_out.cpu().float() - out.float()).abs().max().item() <= out.float().abs().max().item() / 100
    # data gradient of a stride-2 conv: dx[2h,2w] += W^T dz[h,w], other pixels untouched
    dz = rnd((N, H, H, Co), dt, 3)
    wb = wfull.view(Kt, Co)[:C].contiguous().to(dt).view(1, C, Co)        # [N=C][K=Co]
    dx = rnd((N, 2 * H, 2 * H, C), dt, 4)
    ref = dx.clone()
    EMU.igemm_fwd([Seg(dz, Co, H, H, shift=1, aligned=True, w_off=0)], wfull, Co, True, None, ref, N, 2 * H, 2 * H, C,
                  accumulate=True)(0)
    d_dx = dx.clone().cuda()
    lib.conv_tc2_fwd(dz.cuda(), None, wb.cuda(), C, None, d_dx, N, H, H, C, out_stride=2, accumulate=True)(st)
    torch.cuda.synchronize()
    assert (d_dx.cpu().float() - ref.float()).abs().max().item() <= ref.float().abs().max().item() / 100


# ---- thin-layer kernel (conv_tc3.cu): halo tiles, resident weights, fused branches, tensor-core statistics ----
def _conv3_ref(x, w, C, d, N, H, W):
    """fp64 'same' dilated 3x3 conv of bf16 operands; w flat HWIO."""
    segs = [Seg(x, C, H, W, off_h=(ky - 1) * d, off_w=(kx - 1) * d, w_off=(ky * 3 + kx) * C * C)
            for ky in range(3) for kx in range(3)]
    o = torch.zeros((N, H, W, C), dtype=torch.float64)
    EMU.igemm_fwd(segs, w, C, False, None, o, N, H, W, C)(0)
    return o


@pytest.mark.parametrize("C", [32, 64])
@pytest.mark.parametrize("N,H,W,d", [(2, 32, 32, 1), (1, 16, 64, 3), (2, 32, 32, 15), (2, 64, 64, 31), (3, 48, 96, 3),
                                     (1, 32, 32, 31),
                                     # W a multiple of 128 with a large dilation: band items (rows of 128 pixels, one box per tap row)
                                     (1, 16, 128, 15), (2, 32, 256, 31), (3, 48, 128, 31), (1, 64, 128, 15)])
def test_conv_tc3_single_branch(lib, N, H, W, d, C):
    dt = torch.bfloat16
    assert lib.conv_tc3_supported(N, H, W, C)
    x = rnd((N, H, W, C), dt, 1)
    w = rnd((9 * C * C,), torch.float32, 2, 1.0 / (3 * C ** 0.5)).to(dt).float()
    b = rnd((C,), torch.float32, 3)
    wf, wb = _pack(w, 9, C, C)
    res, prev, mask = rnd((N, H, W, C), dt, 5), rnd((N, H, W, C), dt, 4), rnd((N, H, W, C), dt, 6)
    st = torch.cuda.current_stream().cuda_stream
    conv = _conv3_ref(x, w, C, d, N, H, W)
    # bias + identity input (first branch of a ResBlock-a) + statistics
    ref = conv + b.double() + res.double()
    d_out, d_stats = torch.zeros((N, H, W, C), dtype=dt).cuda(), torch.zeros(2 * C, dtype=torch.float64).cuda()
    lib.conv_tc3_fwd([x.cuda()], [wf.cuda()], [b.cuda()], [d], d_out, N, H, W, C, residual=res.cuda(), stats=d_stats)(st)
    torch.cuda.synchronize()
    scale = ref.abs().max().item()
    assert (d_out.cpu().double() - ref).abs().max().item() <= scale / 100, "forward + residual"
    got = d_out.cpu().double().reshape(-1, C)
    # the statistics are those of the stored (bf16) tensor
    np.testing.assert_allclose(d_stats[:C].cpu().numpy(), got.sum(0).numpy(), rtol=1e-4, atol=1e-2)
    np.testing.assert_allclose(d_stats[C:].cpu().numpy(), (got * got).sum(0).numpy(), rtol=1e-4, atol=1e-2)
    # bias + running branch sum (accumulate into out) + ReLU
    ref1 = (conv + b.double() + prev.double()).clamp_min(0)
    d_out1 = prev.clone().cuda()
    lib.conv_tc3_fwd([x.cuda()], [wf.cuda()], [b.cuda()], [d], d_out1, N, H, W, C, accumulate=True, relu=True)(st)
    torch.cuda.synchronize()
    assert (d_out1.cpu().double() - ref1).abs().max().item() <= scale / 100, "forward + accumulate"
    # relu + mask, plain store
    ref2 = (conv + b.double()).clamp_min(0) * (mask.double() > 0)
    d_out2 = torch.zeros((N, H, W, C), dtype=dt).cuda()
    lib.conv_tc3_fwd([x.cuda()], [wf.cuda()], [b.cuda()], [d], d_out2, N, H, W, C, mask=mask.cuda(), relu=True)(st)
    torch.cuda.synchronize()
    assert (d_out2.cpu().double() - ref2).abs().max().item() <= scale / 100, "relu/mask"
    # data gradient: negated dilation with the [tap][ci][co] copy
    dy = rnd((N, H, W, C), dt, 7)
    sg = [Seg(dy, C, H, W, off_h=-(ky - 1) * d, off_w=-(kx - 1) * d, w_off=(ky * 3 + kx) * C * C)
          for ky in range(3) for kx in range(3)]
    dx = torch.zeros((N, H, W, C), dtype=torch.float64)
    EMU.igemm_fwd(sg, w, C, True, None, dx, N, H, W, C)(0)
    d_dx = torch.zeros((N, H, W, C), dtype=dt).cuda()
    lib.conv_tc3_fwd([dy.cuda()], [wb.cuda()], None, [-d], d_dx, N, H, W, C)(st)
    torch.cuda.synchronize()
    assert (d_dx.cpu().double() - dx).abs().max().item() <= dx.abs().max().item() / 100, "dgrad"
    # masked and accumulated (second consumer of a ReLU'd tensor)
    ref3 = prev.double() + dx * (mask.double() > 0)        # mask applies to the sum in the kernel's order: (+prev), then mask
    ref3 = (prev.double() + dx) * (mask.double() > 0)
    d_dx3 = prev.clone().cuda()
    lib.conv_tc3_fwd([dy.cuda()], [wb.cuda()], None, [-d], d_dx3, N, H, W, C, mask=mask.cuda(), accumulate=True)(st)
    torch.cuda.synchronize()
    assert (d_dx3.cpu().double() - ref3).abs().max().item() <= ref3.abs().max().item() / 100, "dgrad mask+acc"


@pytest.mark.parametrize("N,H,W,C,d", [(16, 256, 256, 32, 1), (16, 256, 256, 32, 3), (16, 256, 256, 32, 15),
                                       (16, 256, 256, 32, 31), (16, 128, 128, 64, 1), (16, 128, 128, 64, 3),
                                       (16, 128, 128, 64, 15), (16, 128, 128, 64, 31)])
def test_conv_tc3_at_the_benchmarked_shapes(lib, N, H, W, C, d):
    """The thin-layer launches of BASELINE config 2 (batch 16: enc1 / dec1 / heads at 256^2 x 32, enc2 / dec2 at
    128^2 x 64; halo mode for d <= 3, one box per tap for d = 15 / 31): forward with identity + statistics, data gradient,
    weight gradient, each against the fp64 implicit GEMM on the same bf16 operands (VERDICT r1 missing-3)."""
    dt = torch.bfloat16
    assert lib.conv_tc3_supported(N, H, W, C)
    x, res, dy = rnd((N, H, W, C), dt, 1), rnd((N, H, W, C), dt, 5), rnd((N, H, W, C), dt, 7)
    w = rnd((9 * C * C,), torch.float32, 2, 1.0 / (3 * C ** 0.5)).to(dt).float()
    b = rnd((C,), torch.float32, 3)
    wf, wb = _pack(w, 9, C, C)
    st = torch.cuda.current_stream().cuda_stream
    ref = _conv3_ref(x, w, C, d, N, H, W) + b.double() + res.double()
    d_out, d_stats = torch.zeros((N, H, W, C), dtype=dt).cuda(), torch.zeros(2 * C, dtype=torch.float64).cuda()
    lib.conv_tc3_fwd([x.cuda()], [wf.cuda()], [b.cuda()], [d], d_out, N, H, W, C, residual=res.cuda(), stats=d_stats)(st)
    torch.cuda.synchronize()
    got = d_out.cpu().double()
    assert (got - ref).abs().max().item() <= ref.abs().max().item() / 100, "forward"
    assert float((got - ref).norm() / ref.norm()) <= 4e-3          # bf16 output rounding: ~2^-9 / sqrt(3)
    g2 = got.reshape(-1, C)
    np.testing.assert_allclose(d_stats[:C].cpu().numpy(), g2.sum(0).numpy(), rtol=1e-4, atol=1.0)
    np.testing.assert_allclose(d_stats[C:].cpu().numpy(), (g2 * g2).sum(0).numpy(), rtol=1e-4, atol=1.0)
    sg = [Seg(dy, C, H, W, off_h=-(ky - 1) * d, off_w=-(kx - 1) * d, w_off=(ky * 3 + kx) * C * C)
          for ky in range(3) for kx in range(3)]
    dx = torch.zeros((N, H, W, C), dtype=torch.float64)
    EMU.igemm_fwd(sg, w, C, True, None, dx, N, H, W, C)(0)
    d_dx = torch.zeros((N, H, W, C), dtype=dt).cuda()
    lib.conv_tc3_fwd([dy.cuda()], [wb.cuda()], None, [-d], d_dx, N, H, W, C)(st)
    torch.cuda.synchronize()
    assert float((d_dx.cpu().double() - dx).norm() / dx.norm()) <= 4e-3, "dgrad"
    segs = [Seg(x, C, H, W, off_h=(ky - 1) * d, off_w=(kx - 1) * d, w_off=(ky * 3 + kx) * C * C)
            for ky in range(3) for kx in range(3)]
    dw = torch.zeros(9 * C * C, dtype=torch.float32)
    EMU.igemm_wgrad(segs, dy, dw, C, None, N, H, W, C)(0)
    d_dw = torch.zeros(9 * C * C, dtype=torch.float32).cuda()
    if lib.conv_tc3_wgrad_supported(N, H, W, C, d):
        lib.conv_tc3_wgrad(x.cuda(), dy.cuda(), d_dw, N, H, W, C, d)(st)
    else:
        lib.conv_tc_wgrad(x.cuda(), dy.cuda(), d_dw, N, H, W, C, C, d)(st)
    torch.cuda.synchronize()
    assert float((d_dw.cpu().double() - dw.double()).norm() / dw.double().norm()) <= 1e-3, "wgrad"


@pytest.mark.parametrize("N,H,W,dils", [(2, 64, 64, (1, 3, 15, 31)), (1, 32, 64, (1, 3, 15)), (2, 32, 32, (3, 31)),
                                        (2, 32, 128, (15, 31)), (1, 64, 256, (15, 31)), (1, 32, 128, (1, 3, 15, 31))])
def test_conv_tc3_fused_branches(lib, N, H, W, dils):
    """ResBlock-a branch sum + identity in one launch (model2.py:23-31)."""
    C, dt = 32, torch.bfloat16
    xs = [rnd((N, H, W, C), dt, 10 + i) for i in range(len(dils))]
    ws = [rnd((9 * C * C,), torch.float32, 20 + i, 1.0 / (3 * C ** 0.5)).to(dt).float() for i in range(len(dils))]
    bs = [rnd((C,), torch.float32, 30 + i) for i in range(len(dils))]
    res = rnd((N, H, W, C), dt, 5)
    ref = res.double()
    for x, w, b, d in zip(xs, ws, bs, dils):
        ref = ref + _conv3_ref(x, w, C, d, N, H, W) + b.double()
    ref = ref.clamp_min(0)
    st = torch.cuda.current_stream().cuda_stream
    d_out = torch.zeros((N, H, W, C), dtype=dt).cuda()
    d_stats = torch.zeros(2 * C, dtype=torch.float64).cuda()
    lib.conv_tc3_fwd([x.cuda() for x in xs], [_pack(w, 9, C, C)[0].cuda() for w in ws], [b.cuda() for b in bs], list(dils),
                     d_out, N, H, W, C, residual=res.cuda(), relu=True, stats=d_stats)(st)
    torch.cuda.synchronize()
    assert (d_out.cpu().double() - ref).abs().max().item() <= ref.abs().max().item() / 100
    got = d_out.cpu().double().reshape(-1, C)
    np.testing.assert_allclose(d_stats[:C].cpu().numpy(), got.sum(0).numpy(), rtol=1e-4, atol=1e-2)
    np.testing.assert_allclose(d_stats[C:].cpu().numpy(), (got * got).sum(0).numpy(), rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("N,H,W,d", [(2, 32, 32, 1), (1, 16, 64, 3), (2, 32, 32, 15), (2, 64, 64, 31), (16, 64, 64, 1),
                                     (1, 32, 32, 31), (3, 48, 96, 3),
                                     # W a multiple of 128: one band box per tap row (incl. tap rows that never touch the image)
                                     (1, 16, 128, 15), (2, 32, 256, 31), (3, 64, 128, 31), (1, 48, 128, 15)])
@pytest.mark.parametrize("C", [32, 64])
def test_conv_tc3_wgrad(lib, N, H, W, d, C):
    dt = torch.bfloat16
    if not lib.conv_tc3_wgrad_supported(N, H, W, C, d):
        pytest.skip("large dilations at 64 channels stay on conv_tc_wgrad")
    x = rnd((N, H, W, C), dt, 1)
    dy = rnd((N, H, W, C), dt, 2)
    segs = [Seg(x, C, H, W, off_h=(ky - 1) * d, off_w=(kx - 1) * d, w_off=(ky * 3 + kx) * C * C)
            for ky in range(3) for kx in range(3)]
    dw = torch.zeros(9 * C * C, dtype=torch.float32)
    EMU.igemm_wgrad(segs, dy, dw, C, None, N, H, W, C)(0)
    st = torch.cuda.current_stream().cuda_stream
    d_dw = torch.zeros(9 * C * C, dtype=torch.float32).cuda()
    lib.conv_tc3_wgrad(x.cuda(), dy.cuda(), d_dw, N, H, W, C, d)(st)
    torch.cuda.synchronize()
    scale = dw.abs().max().item()
    assert (d_dw.cpu() - dw).abs().max().item() <= scale * 2e-3


@pytest.mark.parametrize("C", [32, 64])
@pytest.mark.parametrize("N,H,W,d,relu", [(2, 32, 32, 1, True), (2, 32, 64, 15, True), (1, 32, 32, 3, False)])
def test_conv_tc3_fused_bn_backward_reductions(lib, N, H, W, d, relu, C):
    """Data gradient into a = [relu](BN(x)) with FusedBatchNormGrad's reductions {sum g, sum g*xhat} in the epilogue."""
    dt, eps = torch.bfloat16, 1e-3
    dy = rnd((N, H, W, C), dt, 7)
    w = rnd((9 * C * C,), torch.float32, 2, 1.0 / (3 * C ** 0.5)).to(dt).float()
    _, wb = _pack(w, 9, C, C)
    x = (rnd((N, H, W, C), torch.float32, 8) * 1.5 + 0.3).to(dt)
    gamma, beta = rnd((C,), torch.float32, 9) * 0.5 + 1.0, rnd((C,), torch.float32, 10) * 0.3
    xd = x.double().reshape(-1, C)
    cnt = float(xd.shape[0])
    fstats = torch.cat([xd.sum(0), (xd * xd).sum(0)])
    mean = fstats[:C] / cnt
    var = (fstats[C:] / cnt - mean * mean).clamp_min(0)
    xhat = (xd - mean) / torch.sqrt(var + eps)
    sg = [Seg(dy, C, H, W, off_h=-(ky - 1) * d, off_w=-(kx - 1) * d, w_off=(ky * 3 + kx) * C * C)
          for ky in range(3) for kx in range(3)]
    da = torch.zeros((N, H, W, C), dtype=torch.float64)
    EMU.igemm_fwd(sg, w, C, True, None, da, N, H, W, C)(0)
    g = da.reshape(-1, C)
    if relu:
        g = g * ((gamma.double() * xhat + beta.double()) > 0)
    st = torch.cuda.current_stream().cuda_stream
    d_out = torch.zeros((N, H, W, C), dtype=dt).cuda()
    d_red = torch.zeros(2 * C, dtype=torch.float64).cuda()
    lib.conv_tc3_fwd([dy.cuda()], [wb.cuda()], None, [-d], d_out, N, H, W, C, stats=d_red,
                     bnr=(x.cuda(), fstats.cuda(), cnt, eps, gamma.cuda(), beta.cuda(), relu))(st)
    torch.cuda.synchronize()
    scale = g.abs().max().item()
    # elements whose pre-activation sits within rounding of zero may flip their mask between fp32 and fp64
    near = ((gamma.double() * xhat + beta.double()).abs() < 1e-3).reshape(N, H, W, C)
    diff = (d_out.cpu().double() - g.reshape(N, H, W, C)).abs()
    assert diff[~near].max().item() <= scale / 100
    ref_s, ref_q = g.sum(0), (g * xhat).sum(0)
    tol = 2e-2 * (g.abs().sum(0).max().item() / 10 + 1)
    np.testing.assert_allclose(d_red[:C].cpu().numpy(), ref_s.numpy(), atol=tol)
    np.testing.assert_allclose(d_red[C:].cpu().numpy(), ref_q.numpy(), atol=tol)


@pytest.mark.parametrize("N,H,C,d", [(2, 32, 128, 1), (2, 16, 256, 3), (4, 8, 512, 1), (16, 8, 1024, 1), (2, 64, 128, 15)])
def test_conv_tc2_fused_bn_backward_reductions(lib, N, H, C, d):
    """Data gradient of a C >= 128 layer into a = relu(BN(x)): the epilogue masks with the activated tensor, stores g and
    accumulates FusedBatchNormGrad's {sum g, sum g*xhat} from x and the {mean, invstd} table (graph.Plan._wide_dgrad)."""
    dt, eps, W = torch.bfloat16, 1e-3, H
    dy = rnd((N, H, W, C), dt, 7)
    w = rnd((9 * C * C,), torch.float32, 2, 1.0 / (3 * C ** 0.5)).to(dt).float()
    _, wb = _pack(w, 9, C, C)
    x = (rnd((N, H, W, C), torch.float32, 8) * 1.5 + 0.3).to(dt)
    gamma, beta = rnd((C,), torch.float32, 9) * 0.5 + 1.0, rnd((C,), torch.float32, 10) * 0.3
    xd = x.double().reshape(-1, C)
    cnt = float(xd.shape[0])
    fstats = torch.cat([xd.sum(0), (xd * xd).sum(0)])
    mean = fstats[:C] / cnt
    var = (fstats[C:] / cnt - mean * mean).clamp_min(0)
    inv = 1.0 / torch.sqrt(var + eps)
    xhat = (xd - mean) * inv
    a = (gamma.double() * xhat + beta.double()).clamp_min(0).reshape(N, H, W, C).to(dt)       # the stored activation
    # the {mean, invstd} table as the forward BatchNorm launch writes it
    outs = [torch.zeros((N, H, W, C), dtype=dt).cuda()]
    coef = torch.zeros(2 * C, dtype=torch.float32).cuda()
    st = torch.cuda.current_stream().cuda_stream
    lib.bn_apply(x.cuda(), N * H * W, C, outs, [gamma.cuda()], [beta.cuda()], fstats.cuda(), cnt, None, None, eps, True,
                 coef)(st)
    torch.cuda.synchronize()
    np.testing.assert_allclose(coef[:C].cpu().numpy(), mean.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(coef[C:].cpu().numpy(), inv.numpy(), rtol=1e-5)
    assert (outs[0].cpu().float() - a.float()).abs().max().item() <= 0.02 * a.float().abs().max().item()
    sg = [Seg(dy, C, H, W, off_h=-(ky - 1) * d, off_w=-(kx - 1) * d, w_off=(ky * 3 + kx) * C * C)
          for ky in range(3) for kx in range(3)]
    da = torch.zeros((N, H, W, C), dtype=torch.float64)
    EMU.igemm_fwd(sg, w, C, True, None, da, N, H, W, C)(0)
    g = (da * (a.double() > 0)).reshape(-1, C)
    d_out = torch.zeros((N, H, W, C), dtype=dt).cuda()
    d_red = torch.zeros(2 * C, dtype=torch.float64).cuda()
    lib.conv_tc2_fwd(dy.cuda(), None, wb.cuda(), C, None, d_out, N, H, W, C, taps=9, dil=-d, mask=a.cuda(), stats=d_red,
                     bnr_x=x.cuda(), bnr_coef=coef)(st)
    torch.cuda.synchronize()
    scale = g.abs().max().item()
    assert (d_out.cpu().double().reshape(-1, C) - g).abs().max().item() <= scale / 100
    tol = 2e-2 * (g.abs().sum(0).max().item() / 10 + 1)
    np.testing.assert_allclose(d_red[:C].cpu().numpy(), g.sum(0).numpy(), atol=tol)
    np.testing.assert_allclose(d_red[C:].cpu().numpy(), (g * xhat).sum(0).numpy(), atol=tol)
