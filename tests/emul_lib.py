"""CPU emulation of the kernel launcher (resuneta_b200._capi.Lib) — TEST INFRASTRUCTURE ONLY.

Implements the documented semantics of every C-ABI entry point (include/resuneta.h) with plain
torch-CPU ops so that the *host-side* logic (graph construction, backward tape, buffer routing,
Keras surface) can be tested against the oracle without a GPU.  The `-m gpu` tests then check each
CUDA kernel against these same semantics/oracle.  Never importable from the product package.
"""
import math

import torch

F64 = torch.float64


def _gather(seg, N, Ho, Wo):
    src = seg.src.reshape(N, seg.Hs, seg.Ws, seg.C).to(F64)
    h = torch.arange(Ho)
    w = torch.arange(Wo)
    hm, wm = h * seg.mult, w * seg.mult
    vh = torch.ones(Ho, dtype=torch.bool)
    vw = torch.ones(Wo, dtype=torch.bool)
    if seg.aligned:
        msk = (1 << seg.shift) - 1
        vh &= (hm & msk) == 0
        vw &= (wm & msk) == 0
    hs = (hm >> seg.shift) + seg.off_h
    ws = (wm >> seg.shift) + seg.off_w
    vh &= (hs >= 0) & (hs < seg.Hs)
    vw &= (ws >= 0) & (ws < seg.Ws)
    g = src[:, hs.clamp(0, seg.Hs - 1)][:, :, ws.clamp(0, seg.Ws - 1)]
    g = g * (vh[:, None] & vw[None, :])[None, :, :, None]
    if seg.relu_in:
        g = g.clamp_min(0)
    return g.reshape(-1, seg.C)


def _store(dst, val):
    dst.reshape(-1).copy_(val.reshape(-1).to(dst.dtype))


def _mean_inv(stats, count, C, eps, mm=None, mv=None):
    if stats is not None:
        mu = stats[:C] / count
        var = (stats[C:2 * C] / count - mu * mu).clamp_min(0)
        return mu.to(torch.float32).to(F64), (1.0 / torch.sqrt(var + eps)).to(torch.float32).to(F64)
    return mm.to(F64), torch.rsqrt(mv + eps).to(F64)


class EmulLib:
    is_emulation = True

    def __init__(self):
        self.launches = 0

    def _wrap(self, fn, name):
        def launch(stream=0):
            self.launches += 1
            fn()
        launch.kernel = name
        return launch

    # -- igemm ------------------------------------------------------------------------------------------
    def igemm_fwd(self, segs, w, ldw, transB, bias, out, N, Ho, Wo, Co, residual=None, mask=None, stats=None,
                  accumulate=False, relu=False):
        def run():
            acc = torch.zeros(N * Ho * Wo, Co, dtype=F64)
            for s in segs:
                A = _gather(s, N, Ho, Wo)
                if transB:
                    B = torch.as_strided(w, (s.C, Co), (1, ldw), w.storage_offset() + s.w_off).to(F64)
                else:
                    B = torch.as_strided(w, (s.C, Co), (ldw, 1), w.storage_offset() + s.w_off).to(F64)
                acc += A @ B
            if bias is not None:
                acc += bias.to(F64)
            if residual is not None:
                acc += residual.reshape(-1, Co).to(F64)
            if accumulate:
                acc += out.reshape(-1, Co).to(F64)
            if relu:
                acc = acc.clamp_min(0)
            if mask is not None:
                acc = acc * (mask.reshape(-1, Co).to(F64) > 0)
            _store(out, acc)
            if stats is not None:
                v = acc.to(torch.float32).to(F64)
                stats[:Co] += v.sum(0)
                stats[Co:2 * Co] += (v * v).sum(0)
        return self._wrap(run, "rsa_igemm_fwd")

    def igemm_wgrad(self, segs, dy, dw, ldw, dbias, N, Ho, Wo, Co):
        def run():
            g = dy.reshape(-1, Co).to(F64)
            for s in segs:
                A = _gather(s, N, Ho, Wo)
                d = (A.T @ g).to(torch.float32)
                view = torch.as_strided(dw, (s.C, Co), (ldw, 1), dw.storage_offset() + s.w_off)
                view += d
            if dbias is not None:
                dbias.add_(g.sum(0).to(torch.float32))
        return self._wrap(run, "rsa_igemm_wgrad")

    # -- bn ---------------------------------------------------------------------------------------------
    def bn_stats(self, x, M, C, stats):
        def run():
            v = x.reshape(M, C).to(F64)
            stats[:C] += v.sum(0)
            stats[C:2 * C] += (v * v).sum(0)
        return self._wrap(run, "rsa_bn_stats")

    def bn_apply(self, x, M, C, outs, gammas, betas, stats, count, mmeans, mvars, eps, relu, meaninv=None):
        def run():
            v = x.reshape(M, C).to(F64)
            if meaninv is not None:
                mu, inv = _mean_inv(stats, count, C, eps, None if mmeans is None else mmeans[0],
                                    None if mvars is None else mvars[0])
                meaninv[:C] = mu.to(meaninv.dtype)
                meaninv[C:2 * C] = inv.to(meaninv.dtype)
            for k, o in enumerate(outs):
                mu, inv = _mean_inv(stats, count, C, eps, None if mmeans is None else mmeans[k],
                                    None if mvars is None else mvars[k])
                sc = gammas[k].to(F64) * inv
                sh = betas[k].to(F64) - mu * sc
                y = v * sc + sh
                if relu:
                    y = y.clamp_min(0)
                _store(o, y)
        return self._wrap(run, "rsa_bn_apply")

    def bn_bwd_reduce(self, dy, x, act, M, C, stats, count, eps, red):
        def run():
            g = dy.reshape(M, C).to(F64)
            if act is not None:
                g = g * (act.reshape(M, C).to(F64) > 0)
            mu, inv = _mean_inv(stats, count, C, eps)
            xh = (x.reshape(M, C).to(F64) - mu) * inv
            red[:C] += g.sum(0)
            red[C:2 * C] += (g * xh).sum(0)
        return self._wrap(run, "rsa_bn_bwd_reduce")

    def bn_bwd_apply(self, dy, x, act, M, C, stats, count, eps, gamma, red, dx, accumulate, dgamma, dbeta):
        def run():
            g = dy.reshape(M, C).to(F64)
            if act is not None:
                g = g * (act.reshape(M, C).to(F64) > 0)
            mu, inv = _mean_inv(stats, count, C, eps)
            xh = (x.reshape(M, C).to(F64) - mu) * inv
            d = gamma.to(F64) * inv * (g - red[:C] / count - xh * red[C:2 * C] / count)
            if accumulate:
                d = d + dx.reshape(M, C).to(F64)
            _store(dx, d)
            if dgamma is not None:
                dgamma.copy_(red[C:2 * C].to(torch.float32))
            if dbeta is not None:
                dbeta.copy_(red[:C].to(torch.float32))
        return self._wrap(run, "rsa_bn_bwd_apply")

    def bn_bwd_reduce_multi(self, dys, x, M, C, stats, count, eps, gammas, betas, relu, reds):
        def run():
            mu, inv = _mean_inv(stats, count, C, eps)
            xv = x.reshape(M, C).to(F64)
            xh = (xv - mu) * inv
            for dy, ga, be, red in zip(dys, gammas, betas, reds):
                g = dy.reshape(M, C).to(F64)
                if relu:
                    sc = (ga.to(torch.float32) * inv.to(torch.float32)).to(F64)
                    sh = (be.to(torch.float32) - mu.to(torch.float32) * sc.to(torch.float32)).to(F64)
                    g = g * ((xv * sc + sh) > 0)
                red[:C] += g.sum(0)
                red[C:2 * C] += (g * xh).sum(0)
        return self._wrap(run, "rsa_bn_bwd_reduce_multi")

    def bn_bwd_apply_multi(self, dys, x, M, C, stats, count, eps, gammas, betas, relu, reds, dx, accumulate, dgammas,
                           dbetas):
        def run():
            mu, inv = _mean_inv(stats, count, C, eps)
            xv = x.reshape(M, C).to(F64)
            xh = (xv - mu) * inv
            tot = dx.reshape(M, C).to(F64).clone() if accumulate else torch.zeros(M, C, dtype=F64)
            for b, (dy, ga, be, red) in enumerate(zip(dys, gammas, betas, reds)):
                g = dy.reshape(M, C).to(F64)
                if relu:
                    sc = (ga.to(torch.float32) * inv.to(torch.float32)).to(F64)
                    sh = (be.to(torch.float32) - mu.to(torch.float32) * sc.to(torch.float32)).to(F64)
                    g = g * ((xv * sc + sh) > 0)
                tot += ga.to(F64) * inv * (g - red[:C] / count - xh * red[C:2 * C] / count)
                if dgammas is not None and dgammas[b] is not None:
                    dgammas[b].copy_(red[C:2 * C].to(torch.float32))
                if dbetas is not None and dbetas[b] is not None:
                    dbetas[b].copy_(red[:C].to(torch.float32))
            _store(dx, tot)
        return self._wrap(run, "rsa_bn_bwd_apply_multi")

    def bn_derive_stats(self, src_stats, count, gamma, beta, eps, dst_stats, dst_count, C):
        def run():
            mu = src_stats[:C] / count
            var = (src_stats[C:2 * C] / count - mu * mu).clamp_min(0)
            g, b = gamma.to(F64), beta.to(F64)
            vy = g * g * var / (var + eps)
            dst_stats[:C] = dst_count * b
            dst_stats[C:2 * C] = dst_count * (vy + b * b)
        return self._wrap(run, "rsa_bn_derive_stats")

    def bn_update_moving(self, stats_base, param_base, table, counts, nlayers, momentum):
        def run():
            tab = table.reshape(nlayers, 4).tolist()
            cnt = counts.reshape(nlayers, 2).tolist()
            for (soff, C, mo, vo), (n, nfull) in zip(tab, cnt):
                st = stats_base[soff:soff + 2 * C]
                mu = st[:C] / n
                var = (st[C:] / n - mu * mu).clamp_min(0)
                var_u = var * (nfull / (nfull - 1.0)) if nfull > 1 else var
                mm = param_base[mo:mo + C]
                mv = param_base[vo:vo + C]
                mm.copy_((mm.to(F64) * momentum + mu * (1 - momentum)).to(torch.float32))
                mv.copy_((mv.to(F64) * momentum + var_u * (1 - momentum)).to(torch.float32))
        return self._wrap(run, "rsa_bn_update_moving")

    # -- pooling ----------------------------------------------------------------------------------------
    def maxpool_pyr_fwd(self, x, N, H, W, C, p2, p4, p8):
        def run():
            v = x.reshape(N, H, W, C).permute(0, 3, 1, 2).float()
            for k, p in ((2, p2), (4, p4), (8, p8)):
                if p is not None:
                    _store(p, torch.nn.functional.max_pool2d(v, k, k).permute(0, 2, 3, 1))
        return self._wrap(run, "rsa_maxpool_pyr_fwd")

    def maxpool_pyr_bwd(self, x, N, H, W, C, dp2, dp4, dp8, dx, accumulate):
        def run():
            v = x.reshape(N, H, W, C).permute(0, 3, 1, 2).to(F64)
            tot = dx.reshape(N, H, W, C).to(F64).clone() if accumulate else torch.zeros(N, H, W, C, dtype=F64)
            for k, dp in ((2, dp2), (4, dp4), (8, dp8)):
                if dp is None:
                    continue
                _, idx = torch.nn.functional.max_pool2d(v, k, k, return_indices=True)
                g = dp.reshape(N, H // k, W // k, C).permute(0, 3, 1, 2).to(F64)
                z = torch.zeros(N, C, H * W, dtype=F64)
                z.scatter_add_(2, idx.reshape(N, C, -1), g.reshape(N, C, -1))
                tot += z.reshape(N, C, H, W).permute(0, 2, 3, 1)
            _store(dx, tot)
        return self._wrap(run, "rsa_maxpool_pyr_bwd")

    def sumpool_pyr(self, x, N, H, W, C, s2, s4, s8):
        def run():
            v = x.reshape(N, H, W, C).permute(0, 3, 1, 2).to(F64)
            for k, s in ((2, s2), (4, s4), (8, s8)):
                if s is not None:
                    _store(s, (torch.nn.functional.avg_pool2d(v, k, k) * (k * k)).permute(0, 2, 3, 1))
        return self._wrap(run, "rsa_sumpool_pyr")

    # -- heads / losses ---------------------------------------------------------------------------------
    def softmax_fwd(self, z, p, M, C):
        return self._wrap(lambda: _store(p, torch.softmax(z.reshape(M, C).to(F64), -1)), "rsa_softmax_fwd")

    def softmax_bwd(self, p, dp, dz, M, C):
        def run():
            pv, g = p.reshape(M, C).to(F64), dp.reshape(M, C).to(F64)
            _store(dz, pv * (g - (pv * g).sum(-1, keepdim=True)))
        return self._wrap(run, "rsa_softmax_bwd")

    def sigmoid_fwd(self, z, p, n):
        return self._wrap(lambda: _store(p, torch.sigmoid(z.reshape(-1).to(F64))), "rsa_sigmoid_fwd")

    def sigmoid_bwd(self, p, dp, dz, n):
        def run():
            pv = p.reshape(-1).to(F64)
            _store(dz, dp.reshape(-1).to(F64) * pv * (1 - pv))
        return self._wrap(run, "rsa_sigmoid_bwd")

    def tanimoto_sums(self, pred, label, B, HW, C, sums):
        def run():
            p = pred.reshape(B, HW, C).to(F64)
            l = label.reshape(B, HW, C).to(F64)
            s = torch.stack([p.sum(1), (p * p).sum(1), l.sum(1), (l * l).sum(1), (p * l).sum(1)], dim=-1)
            sums[:B * C * 5] += s.reshape(-1)
        return self._wrap(run, "rsa_tanimoto_sums")

    def tanimoto_finalize(self, sums, B, HW, C, scale, loss_b, loss_mean, coef):
        def run():
            S = sums[:B * C * 5].reshape(B, C, 5)
            sp, sp2, sl, sl2, spl = (S[..., i] for i in range(5))
            sm = 1e-5
            V1 = sp.mean(0)
            V2 = (HW - sl).mean(0)
            w1, w2 = 1 / V1 ** 2, 1 / V2 ** 2
            for w in (w1, w2):
                inf = torch.isinf(w)
                fin = torch.where(inf, torch.zeros_like(w), w)
                w[inf] = fin.max()
            V1 = torch.where(torch.isinf(1 / sp.mean(0) ** 2), torch.zeros_like(V1), V1)
            n1 = (w1 * spl).sum(1)
            d1 = (w1 * (sp2 + sl2 - spl)).sum(1)
            prod = HW - sl - sp + spl
            sq = (HW - 2 * sp + sp2) + (HW - 2 * sl + sl2)
            n2 = (w2 * prod).sum(1)
            d2 = (w2 * (sq - prod)).sum(1)
            lb = 1 - 0.5 * ((n1 + sm) / (d1 + sm) + (n2 + sm) / (d2 + sm))
            if loss_b is not None:
                loss_b.copy_(lb.to(torch.float32))
            if loss_mean is not None:
                loss_mean.copy_(lb.mean().to(torch.float32).reshape(1))
            if coef is not None:
                A1, R1 = 1 / (d1 + sm), (n1 + sm) / (d1 + sm) ** 2
                A2, R2 = 1 / (d2 + sm), (n2 + sm) / (d2 + sm) ** 2
                acc = (A1[:, None] * spl - R1[:, None] * (sp2 + sl2 - spl)).sum(0)
                G = torch.where(V1 > 0, (-2.0 / (B * V1.clamp_min(1e-300) ** 3)) * acc, torch.zeros_like(acc))
                k = -scale / (2.0 * B)
                c0 = k * (G[None, :] + w2[None, :] * (R2 - A2)[:, None])
                c1 = k * (-2 * R1[:, None] * w1[None, :] - 2 * R2[:, None] * w2[None, :])
                c2 = k * (w1[None, :] * (A1 + R1)[:, None] + w2[None, :] * (A2 + R2)[:, None])
                coef.copy_(torch.stack([c0, c1, c2], -1).reshape(-1).to(torch.float32))
        return self._wrap(run, "rsa_tanimoto_finalize")

    def tanimoto_bwd(self, pred, label, coef, B, HW, C, dpred):
        def run():
            c = coef.reshape(B, 1, C, 3).to(F64)
            p = pred.reshape(B, HW, C).to(F64)
            l = label.reshape(B, HW, C).to(F64)
            _store(dpred, c[..., 0] + c[..., 1] * p + c[..., 2] * l)
        return self._wrap(run, "rsa_tanimoto_bwd")

    @staticmethod
    def _pixel_loss(kind, p, y, w):
        eps = 1e-7
        if kind == 0:
            q = (p / p.sum(-1, keepdim=True)).clamp(eps, 1 - eps)
            ww = w.to(F64) if w is not None else 1.0
            return -(y * torch.log(q) * ww).sum(-1)
        if kind == 1:
            q = p.clamp(eps, 1 - eps)
            return -(y * torch.log(q) + (1 - y) * torch.log(1 - q)).mean(-1)
        return ((p - y) ** 2).mean(-1)

    def pixel_loss_fwd(self, kind, pred, label, weights, M, C, loss_sum):
        def run():
            l = self._pixel_loss(kind, pred.reshape(M, C).to(F64), label.reshape(M, C).to(F64), weights)
            loss_sum.add_(l.sum())
        return self._wrap(run, "rsa_pixel_loss_fwd")

    def pixel_loss_elem(self, kind, pred, label, weights, M, C, out):
        def run():
            _store(out, self._pixel_loss(kind, pred.reshape(M, C).to(F64), label.reshape(M, C).to(F64), weights))
        return self._wrap(run, "rsa_pixel_loss_elem")

    def pixel_loss_bwd(self, kind, pred, label, weights, M, C, scale, dpred):
        def run():
            p = pred.reshape(M, C).to(F64).clone().requires_grad_(True)
            l = self._pixel_loss(kind, p, label.reshape(M, C).to(F64), weights).sum() * scale
            g, = torch.autograd.grad(l, p)
            _store(dpred, g)
        return self._wrap(run, "rsa_pixel_loss_bwd")

    def seg_metrics(self, pred, label, M, C, out):
        def run():
            p, y = pred.reshape(M, C), label.reshape(M, C)
            t, q = y > 0.5, p > 0.5
            out[0] += (p.argmax(-1) == y.argmax(-1)).sum()
            out[1] += (t & q).sum()
            out[2] += (~t & q).sum()
            out[3] += (~t & ~q).sum()
            out[4] += (t & ~q).sum()
        return self._wrap(run, "rsa_seg_metrics")

    def argmax_confusion(self, prob, M, C, pred_label, true_label, K, cm):
        def run():
            a = prob.reshape(M, C).argmax(-1).to(torch.int32)
            if pred_label is not None:
                pred_label.copy_(a)
            if cm is not None and true_label is not None:
                idx = true_label.to(torch.int64) * K + a.to(torch.int64)
                cm.add_(torch.bincount(idx, minlength=K * K))
        return self._wrap(run, "rsa_argmax_confusion")

    # -- thin 1x1 convolutions --------------------------------------------------------------------------
    def stem_fwd(self, x, w, b, out, M, n, stats):
        def run():
            v = x.reshape(M, n).to(F64) @ w[:n * 32].reshape(n, 32).to(F64)
            if b is not None:
                v = v + b.to(F64)
            _store(out, v)
            if stats is not None:
                vv = v.to(torch.float32).to(F64)
                stats[:32] += vv.sum(0)
                stats[32:64] += (vv * vv).sum(0)
        return self._wrap(run, "rsa_stem_fwd")

    def stem_wgrad(self, x, dy, M, n, dw, db):
        def run():
            g = dy.reshape(M, 32).to(F64)
            dw[:n * 32] += (x.reshape(M, n).to(F64).T @ g).reshape(-1).to(torch.float32)
            if db is not None:
                db.add_(g.sum(0).to(torch.float32))
        return self._wrap(run, "rsa_stem_wgrad")

    def head_bwd(self, h, dz, w, M, n, dh, accumulate, relu_mask, dw, db):
        def run():
            hv = h.reshape(M, 32).to(F64)
            z = dz.reshape(M, n).to(F64)
            if dh is not None:
                d = z @ w[:32 * n].reshape(32, n).to(F64).T
                if relu_mask:
                    d = d * (hv > 0)
                if accumulate:
                    d = d + dh.reshape(M, 32).to(F64)
                _store(dh, d)
            dw[:32 * n] += (hv.T @ z).reshape(-1).to(torch.float32)
            if db is not None:
                db.add_(z.sum(0).to(torch.float32))
        return self._wrap(run, "rsa_head_bwd")

    # -- optim / misc -----------------------------------------------------------------------------------
    def adam_step(self, param, grad, m, v, n, lr_dev, b1, b2, eps, grad_scale):
        def run():
            g = grad[:n] * grad_scale
            m[:n] = b1 * m[:n] + (1 - b1) * g
            v[:n] = b2 * v[:n] + (1 - b2) * g * g
            param[:n] -= lr_dev[0] * m[:n] / (v[:n].sqrt() + eps)
        return self._wrap(run, "rsa_adam_step")

    def sgd_step(self, param, grad, vel, n, lr_dev, momentum, grad_scale):
        def run():
            vel[:n] = momentum * vel[:n] - lr_dev[0] * grad[:n] * grad_scale
            param[:n] += vel[:n]
        return self._wrap(run, "rsa_sgd_step")

    def axpy(self, dst, src, n, accumulate):
        def run():
            v = src.reshape(-1)[:n].to(F64)
            if accumulate:
                v = v + dst.reshape(-1)[:n].to(F64)
            dst.reshape(-1)[:n] = v.to(dst.dtype)
        return self._wrap(run, "rsa_axpy")

    def cast(self, src, dst, n):
        return self._wrap(lambda: _store(dst, src), "rsa_cast")


# ------------------------------------------------------------------------------------------------------
# bf16 tensor-core path (conv_tc*.cu, thin.cu head forward) restated on the CPU: the same launch list the GPU runs in
# bf16 mode - K-concatenated 1x1 convolutions with up-sampled addends, thin-layer kernels with the fused BatchNorm-backward
# sums, packed weight copies, tensor-core weight gradients - so that the host logic of that mode (fusions, side / join /
# lane / chain hints) is testable without a GPU.  Math in float64, tensors stored in bf16 exactly like the kernels' outputs.
# ------------------------------------------------------------------------------------------------------
def _taps_conv(x, w, taps, dil, s, H, W):
    """x f64 [N,Hs,Ws,C], w f64 [taps][Co][C] -> [N,H,W,Co]; sample (h*s+dy*dil, w*s+dx*dil), zero outside."""
    N, Hs, Ws, C = x.shape
    Co = w.shape[1]
    out = torch.zeros(N, H, W, Co, dtype=F64)
    pad = abs(dil) if taps == 9 else 0
    xp = torch.nn.functional.pad(x, (0, 0, pad, pad, pad, pad))
    for t in range(taps):
        dy, dx = (t // 3 - 1, t % 3 - 1) if taps == 9 else (0, 0)
        h0, w0 = pad + dy * dil, pad + dx * dil
        win = xp[:, h0:h0 + (H - 1) * s + 1:s, w0:w0 + (W - 1) * s + 1:s, :]
        out += win @ w[t].T
    return out


def _taps_wgrad(x, dy, dil):
    """dw[tap][ci][co] = sum_pix x[pix + off(tap)*dil, ci] * dy[pix, co]; x, dy f64 NHWC."""
    N, H, W, C = x.shape
    Co = dy.shape[-1]
    pad = abs(dil)
    xp = torch.nn.functional.pad(x, (0, 0, pad, pad, pad, pad))
    g = dy.reshape(-1, Co)
    dw = torch.zeros(9, C, Co, dtype=F64)
    for t in range(9):
        ddy, ddx = t // 3 - 1, t % 3 - 1
        win = xp[:, pad + ddy * dil:pad + ddy * dil + H, pad + ddx * dil:pad + ddx * dil + W, :]
        dw[t] = win.reshape(-1, C).T @ g
    return dw


class EmulLibTC(EmulLib):
    """EmulLib + the tensor-core entry points: Net takes the bf16 tcgen05 path on the CPU (graph.Net gating)."""
    emulates_tensor_core = True

    @staticmethod
    def _pow2(v):
        return v >= 1 and (v & (v - 1)) == 0

    # shapes: conv_tc.cu:300, conv_tc2.cu:355, conv_tc3.cu (rsa_conv_tc3_supported / _wgrad_supported)
    def conv_tc_supported(self, N, H, W, Cin, Cout):
        okc = lambda c: c == 32 or (c >= 64 and c % 64 == 0)
        return okc(Cin) and okc(Cout) and H == W and self._pow2(W) and W >= 4

    def conv_tc2_supported(self, N, H, W, C0, C1, Cout):
        if not (self._pow2(H) and self._pow2(W)) or N < 1:
            return False
        if C0 == 8 and C1 == 0:
            return Cout >= 1
        return not (C0 < 16 or C0 % 16 or (C1 and C1 % 16) or Cout < 1)

    def conv_tc3_supported(self, N, H, W, C):
        return C in (32, 64) and N >= 1 and H >= 16 and H % 16 == 0 and W >= 32 and W % 32 == 0

    def conv_tc3_wgrad_supported(self, N, H, W, C, dil):
        return self.conv_tc3_supported(N, H, W, C) and dil > 0 and (C == 32 or dil <= 3 or (W % 128 == 0 and 128 + 2 * dil <= 256))

    def pack_weights_tc(self, params, shadow, table, nlayers, max_elems):
        import struct
        raw = bytes(table.cpu().numpy().tobytes())

        def run():
            for l in range(nlayers):
                src, fwd, bwd, taps, cin, cout, pad = struct.unpack_from("<qqqiiii", raw, l * 40)
                coutp = pad if pad > 0 else cout
                w = params[src:src + taps * cin * cout].reshape(taps, cin, cout)
                shadow[bwd:bwd + taps * cin * cout] = w.reshape(-1).to(torch.bfloat16)
                f = shadow[fwd:fwd + taps * coutp * cin].view(taps, coutp, cin)
                f[:, :cout, :] = w.permute(0, 2, 1).to(torch.bfloat16)
        return self._wrap(run, "rsa_pack_weights_tc")

    @staticmethod
    def _epilogue(acc, out, Cout, residual, mask, accumulate, relu, stats, os_=1):
        """acc f64 [N,H,W,Cout] -> store (bf16 or fp32) at stride os_; order: residual, accumulate, relu, mask."""
        N, H, W, _ = acc.shape
        ov = out.reshape(N, H * os_, W * os_, Cout)[:, ::os_, ::os_, :]
        if residual is not None:
            acc = acc + residual.reshape(N, H * os_, W * os_, Cout)[:, ::os_, ::os_, :].to(F64)
        if accumulate:
            acc = acc + ov.to(F64)
        if relu:
            acc = acc.clamp_min(0)
        if mask is not None:
            acc = acc * (mask.reshape(N, H * os_, W * os_, Cout)[:, ::os_, ::os_, :].to(F64) > 0)
        ov.copy_(acc.to(out.dtype))
        if stats is not None:
            v = ov.to(F64).reshape(-1, Cout)       # statistics of the stored (rounded) values
            stats[:Cout] += v.sum(0)
            stats[Cout:2 * Cout] += (v * v).sum(0)

    def conv_tc2_fwd(self, x0, x1, wt, CoutP, bias, out, N, H, W, Cout, taps=1, dil=1, in_stride=1, ups=(),
                     residual=None, mask=None, stats=None, accumulate=False, relu=False, k_base=0, k_total=0,
                     out_stride=1, bnr_x=None, bnr_coef=None):
        C0 = x0.shape[-1]
        C1 = x1.shape[-1] if x1 is not None else 0
        K = C0 + C1
        kt = k_total if k_total else K

        def run():
            s = in_stride
            xs = x0.reshape(N, H * s, W * s, C0).to(F64)
            if x1 is not None:
                xs = torch.cat([xs, x1.reshape(N, H * s, W * s, C1).to(F64)], -1)
            # rows [0, Cout) of every tap's [CoutP][kt] matrix, columns [k_base, k_base + K)
            w = torch.stack([torch.as_strided(wt, (Cout, K), (kt, 1), wt.storage_offset() + t * CoutP * kt + k_base)
                             for t in range(taps)]).to(F64)
            acc = _taps_conv(xs, w, taps, dil, s, H, W)
            if bias is not None:
                acc = acc + bias[:Cout].to(F64)
            for q, sh in ups:
                qv = q.reshape(N, H >> sh, W >> sh, Cout).to(F64)
                acc = acc + qv.repeat_interleave(1 << sh, 1).repeat_interleave(1 << sh, 2)
            if bnr_x is None:
                self._epilogue(acc, out, Cout, residual, mask, accumulate, relu, stats, out_stride)
                return
            # fused BatchNorm-backward reductions: the stored value g is masked by the activated tensor `mask`; stats +=
            # {sum g, sum g * xhat} with xhat from bnr_x and the {mean, invstd} table
            assert stats is not None and mask is not None and out_stride == 1 and not accumulate and residual is None
            g = acc * (mask.reshape(N, H, W, Cout).to(F64) > 0)
            mean, inv = bnr_coef[:Cout].to(F64), bnr_coef[Cout:2 * Cout].to(F64)
            xh = (bnr_x.reshape(N, H, W, Cout).to(F64) - mean) * inv
            stats[:Cout] += g.reshape(-1, Cout).sum(0)
            stats[Cout:2 * Cout] += (g * xh).reshape(-1, Cout).sum(0)
            out.reshape(N, H, W, Cout).copy_(g.to(out.dtype))
        return self._wrap(run, "rsa_conv_tc2_fwd")

    def conv_tc3_fwd(self, xs, wts, biases, dils, out, N, H, W, C, residual=None, mask=None, stats=None,
                     accumulate=False, relu=False, bnr=None):
        assert not (residual is not None and accumulate), "one addend"

        def run():
            acc = torch.zeros(N, H, W, C, dtype=F64)
            for b, (x, w, d) in enumerate(zip(xs, wts, dils)):
                acc += _taps_conv(x.reshape(N, H, W, C).to(F64), w.reshape(9, C, C).to(F64), 9, int(d), 1, H, W)
                if biases is not None and biases[b] is not None:
                    acc += biases[b].to(F64)
            if bnr is None:
                self._epilogue(acc, out, C, residual, mask, accumulate, relu, stats)
                return
            bx, bst, cnt, eps, gam, bet, brelu = bnr
            assert mask is None and stats is not None
            if residual is not None:
                acc = acc + residual.reshape(N, H, W, C).to(F64)
            if accumulate:
                acc = acc + out.reshape(N, H, W, C).to(F64)
            mean, inv = _mean_inv(bst, cnt, C, eps)
            xh = (bx.reshape(N, H, W, C).to(F64) - mean) * inv
            if brelu:
                acc = acc * ((gam.to(F64) * xh + bet.to(F64)) > 0)
            stats[:C] += acc.reshape(-1, C).sum(0)                   # sums of the fp32 values, before the bf16 store
            stats[C:2 * C] += (acc * xh).reshape(-1, C).sum(0)
            out.reshape(N, H, W, C).copy_(acc.to(out.dtype))
        return self._wrap(run, "rsa_conv_tc3_fwd")

    def conv_tc_wgrad(self, x, dy, dw, N, H, W, Cin, Cout, dil):
        def run():
            d = _taps_wgrad(x.reshape(N, H, W, Cin).to(F64), dy.reshape(N, H, W, Cout).to(F64), dil)
            dw[:9 * Cin * Cout] += d.reshape(-1).to(torch.float32)
        return self._wrap(run, "rsa_conv_tc_wgrad")

    def conv_tc3_wgrad(self, x, dy, dw, N, H, W, C, dil):
        op = self.conv_tc_wgrad(x, dy, dw, N, H, W, C, C, dil)
        op.kernel = "rsa_conv_tc3_wgrad"
        return op

    def pw_wgrad_tc(self, x, dz, dw, ldw, N, H, W, Cin, Cout, in_stride=1):
        def run():
            s = in_stride
            xv = x.reshape(N, H * s, W * s, Cin)[:, ::s, ::s, :].to(F64).reshape(-1, Cin)
            d = (xv.T @ dz.reshape(-1, Cout).to(F64)).to(torch.float32)
            view = torch.as_strided(dw, (Cin, Cout), (ldw, 1), dw.storage_offset())
            view += d
        return self._wrap(run, "rsa_pw_wgrad_tc")

    def bias_grad(self, dy, M, C, dbs):
        def run():
            g = dy.reshape(M, C).to(F64).sum(0).to(torch.float32)
            for db in dbs:
                if db is not None:
                    db.add_(g)
        return self._wrap(run, "rsa_bias_grad")

    def head_fwd(self, h, w, b, z, M, n):
        def run():
            v = h.reshape(M, 32).to(F64) @ w[:32 * n].reshape(32, n).to(F64)
            if b is not None:
                v = v + b.to(F64)
            _store(z, v)
        return self._wrap(run, "rsa_head_fwd")

