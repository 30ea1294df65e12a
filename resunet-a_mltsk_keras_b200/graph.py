"""Execution plans for the ResUnet-a hot path: forward ops, backward tape, buffers.

A :class:`Net` owns the parameters (one flat fp32 buffer, Keras names and HWIO kernels, see
SURVEY.md §C) and builds :class:`Plan` objects — flat lists of pre-bound kernel launches for a
given (batch, mode).  The graph that the reference expresses as ~294 Keras layers
(ResUnet_a/model2.py:14-193, ResUnet_a/model.py:14-171) is emitted here as ~10 kernel kinds:

  * every convolution is an implicit GEMM over K-segments (``_capi.Seg``): dilated 3x3 taps,
    channel concat, nearest up-sampling and stride-2 sampling are gather maps, so no padded,
    concatenated or up-sampled tensor is ever materialised;
  * training-mode BatchNormalization is split into {statistics in the producer's epilogue,
    normalise(+ReLU) in one light pass that writes all branch variants at once};
  * 1x1 convolutions that keras applies to an up-sampled tensor (PSPPooling model2.py:55-68,
    UpSampling model2.py:89-94) run at the pooled resolution — Conv1x1(Up(x)) == Up(Conv1x1(x))
    and the biased batch statistics of Up(t) equal those of t.

Backward is a tape of closures emitted in reverse; gradient buffers are written by their first
consumer and accumulated by the others (``Plan.gacc``).
"""
from __future__ import annotations

import contextlib
import math
import os
import struct
from collections import OrderedDict

import torch

from ._capi import Seg

BN_EPS = 1e-3
BN_MOMENTUM = 0.99
FILTERS = [32, 64, 128, 256, 512, 1024]
DILATIONS = [[1, 3, 15, 31], [1, 3, 15, 31], [1, 3, 15], [1, 3, 15], [1], [1]]
ALIGN = 64   # every parameter starts on a 256-byte boundary inside the flat buffer


def psp_levels(img_width):
    """PSPPooling levels are gated on the *input patch* width (model2.py:49-52)."""
    lv = [1, 2]
    if img_width >= 128:
        lv.append(4)
    if img_width >= 256:
        lv.append(8)
    return lv


class T:
    """Activation tensor handle (NHWC, dense)."""
    __slots__ = ("name", "N", "H", "W", "C", "dtype", "data", "grad", "grad_written", "needs_grad",
                 "relu_masked", "stats", "count", "bn_src", "grad_writers")

    def __init__(self, name, N, H, W, C, dtype):
        self.name, self.N, self.H, self.W, self.C, self.dtype = name, N, H, W, C, dtype
        self.data = None
        self.grad = None
        self.grad_written = False
        self.needs_grad = True
        self.relu_masked = False   # value went through a fused ReLU: writers of .grad mask with data>0
        self.stats = None          # double[2C] {sum, sumsq} view, valid after the producer ran
        self.count = 0.0           # elements per channel behind .stats
        self.bn_src = None         # this tensor is [relu](BatchNorm(x)): what a fused backward reduction needs
        self.grad_writers = 0      # launches that write / accumulate into .grad (Plan.gacc, identity alias)

    @property
    def M(self):
        return self.N * self.H * self.W

    @property
    def shape(self):
        return (self.N, self.H, self.W, self.C)


class ParamStore:
    """Flat fp32 parameter / gradient buffers; trainable tensors first."""

    def __init__(self):
        self.spec = OrderedDict()   # name -> (shape, trainable, init)
        self.off = {}
        self.n_train = 0
        self.n_total = 0
        self.data = None
        self.grad = None

    def add(self, name, shape, trainable, init):
        if name in self.spec:
            raise ValueError(f"duplicate parameter {name}")
        self.spec[name] = (tuple(shape), trainable, init)

    def finalize(self, device, seed):
        off = 0
        for trainable in (True, False):
            for name, (shape, tr, _) in self.spec.items():
                if tr != trainable:
                    continue
                self.off[name] = off
                n = math.prod(shape)
                off += (n + ALIGN - 1) // ALIGN * ALIGN
            if trainable:
                self.n_train = off
        self.n_total = off
        host = torch.zeros(self.n_total, dtype=torch.float32)
        gen = torch.Generator().manual_seed(seed)
        for name, (shape, _, init) in self.spec.items():
            n = math.prod(shape)
            v = host[self.off[name]:self.off[name] + n].view(shape)
            if init == "glorot":
                kh, kw, cin, cout = shape
                limit = math.sqrt(6.0 / (kh * kw * cin + kh * kw * cout))
                v.copy_(((torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * limit).float())
            elif init == "ones":
                v.fill_(1.0)
        self.data = host.to(device)
        self.grad = torch.zeros(self.n_train, dtype=torch.float32, device=device)

    def view(self, name):
        shape = self.spec[name][0]
        o = self.off[name]
        return self.data[o:o + math.prod(shape)].view(shape)

    def flat(self, name):
        shape = self.spec[name][0]
        o = self.off[name]
        return self.data[o:o + math.prod(shape)]

    def gflat(self, name):
        shape, tr, _ = self.spec[name]
        assert tr
        o = self.off[name]
        return self.grad[o:o + math.prod(shape)]

    def gview(self, name):
        return self.gflat(name).view(self.spec[name][0])


class _Ops(list):
    """Launch list of a plan (forward or backward) carrying the hints the executor (keras_api._run_ops) schedules by:
      side   weight/bias-gradient launches (Plan._side): they only feed the optimizer and may run on a second stream;
      join   the first launch that overwrites a buffer a pending side launch still reads (Plan.gacc notices the write);
      lane   launches of one ResBlock-a branch (Plan.lane): branches are independent chains, so alternating them over
             two streams lets one branch's bandwidth-bound BatchNorm pass overlap the other's convolutions;
      chain  launches that read-modify-write the same tensor across lanes (branch sum) and must keep their order."""

    def __init__(self, plan):
        super().__init__()
        self.plan = plan

    def append(self, op):
        pl = self.plan
        if pl._lane is not None:
            op.lane = pl._lane
        if not getattr(op, "side", False) and pl._chain_next is not None:
            if not hasattr(op, "chain"):
                op.chain = pl._chain_next
            pl._chain_next = None
        if not getattr(op, "side", False) and pl._join_next:
            op.join = True
            pl._join_next = False
            pl._side_reads.clear()
        super().append(op)


class _Tape(list):
    """Backward closures; one appended inside a Plan.lane() region emits its launches into that lane again."""

    def __init__(self, plan):
        super().__init__()
        self.plan = plan

    def append(self, fn):
        pl, lane = self.plan, self.plan._lane
        if lane is None:
            super().append(fn)
            return

        def in_lane():
            prev, pl._lane = pl._lane, lane
            try:
                fn()
            finally:
                pl._lane = prev
        super().append(in_lane)


class Plan:
    """One walk of the network for a fixed batch size and mode.

    spec_only=True registers parameters only (no allocation, no launches)."""

    def __init__(self, net, N, training, spec_only=False, with_loss=False):
        self.net = net
        self.lib = net.lib
        self.N = N
        self.training = training
        self.spec_only = spec_only
        self.with_loss = with_loss
        self.device = net.device
        self.adt = net.act_dtype
        self._nconv = 0
        self._nbn = 0
        self._side_reads = set()   # buffers read by side launches emitted since the last join
        self._join_next = False
        self._chain_next = None
        self._lane = None
        self.fwd = _Ops(self)      # launch(stream) callables
        self.bwd = _Ops(self)
        self.tape = _Tape(self)    # closures that emit backward launches (run reversed)
        self.bn_table = []   # (stats view, C, moving_mean name, moving_var name, count, count_full)
        self._scratch_chunks = []
        self.scratch = None
        self.tensors = []
        self.outputs = OrderedDict()
        self.labels = OrderedDict()
        self.loss_out = OrderedDict()
        self.bytes = 0
        self.metrics = None
        self.bn_update = None
        self.grad_ready = {}  # param name -> index of the backward op that completes its gradient
        self.res_f32 = None   # [8] fp32: tanimoto loss means per head
        self.res_z = None     # 16 zeroed 8-byte words: [0..7] pixel-loss sums (double), [8..12] seg metrics (int64)

    # -- names (keras creation order) -----------------------------------------------------------
    def name_conv(self):
        i = self._nconv
        self._nconv += 1
        return "conv2d" if i == 0 else f"conv2d_{i}"

    def name_bn(self):
        i = self._nbn
        self._nbn += 1
        return "batch_normalization" if i == 0 else f"batch_normalization_{i}"

    # -- memory -----------------------------------------------------------------------------------
    def alloc(self, shape, dtype):
        if self.spec_only:
            return None
        t = torch.empty(shape, dtype=dtype, device=self.device)
        self.bytes += t.numel() * t.element_size()
        return t

    def tensor(self, name, H, W, C, dtype=None, N=None):
        t = T(name, self.N if N is None else N, H, W, C, dtype or self.adt)
        t.data = self.alloc(t.shape, t.dtype)
        self.tensors.append(t)
        return t

    def zeroed(self, n):
        """n 8-byte words of scratch, zeroed at the start of every step (double / int64 views)."""
        if self.spec_only:
            return None
        holder = [None]
        self._scratch_chunks.append((n, holder))
        return holder

    def _finalize_scratch(self):
        total = sum((n + 7) // 8 * 8 for n, _ in self._scratch_chunks)
        self.scratch = torch.zeros(max(total, 8), dtype=torch.float64, device=self.device)
        off = 0
        for n, holder in self._scratch_chunks:
            holder[0] = self.scratch[off:off + n]
            off += (n + 7) // 8 * 8

    def gacc(self, t):
        """Gradient buffer of t and whether the caller must accumulate into it."""
        if t.grad is None:
            t.grad = self.alloc(t.shape, t.dtype)
        acc = t.grad_written
        t.grad_written = True
        t.grad_writers += 1
        if self._side_reads and t.grad.data_ptr() in self._side_reads:
            self._join_next = True     # the next main-stream launch overwrites what a side launch still reads
        if self._lane is not None:
            self._chain_next = t.grad.data_ptr()   # write / accumulate order on a gradient shared across lanes
        return t.grad, acc

    @contextlib.contextmanager
    def lane(self, k):
        """Launches (and backward closures) emitted inside belong to branch k: an independent chain between the
        surrounding launches, which act as barriers.  Anything shared across lanes must be read-only or `chain`ed."""
        prev, self._lane = self._lane, k
        try:
            yield
        finally:
            self._lane = prev

    def _side(self, op, *bufs):
        """Tag a weight/bias-gradient launch as runnable beside the main chain; bufs = the tensors it reads."""
        op.side = True
        self._side_reads.update(b.data_ptr() for b in bufs if b is not None)
        return op

    @staticmethod
    def _tag(op, tag, flops, nbytes=None):
        """3x3 convolution launch: algorithmic FLOPs (dense 9-tap count) and minimal HBM bytes (one operand tensor read, one
        result tensor written / second operand read; side inputs and weights not counted)."""
        op.tag = tag
        op.flops = flops
        op.conv_bytes = nbytes
        return op

    @staticmethod
    def _hb(op, nbytes):
        """Algorithmic HBM bytes of a bandwidth-bound launch (what it must read + write once): bench.py and
        scripts/hbm_kernels.py divide them by the measured duration (roofline of the PSP / BN / head / loss kernels)."""
        op.hbm_bytes = float(nbytes)
        return op

    def _ready(self, *names):
        for n in names:
            self.grad_ready[n] = len(self.bwd) - 1

    def P(self, name):
        return self.net.params.flat(name)

    def G(self, name):
        return self.net.params.gflat(name)

    # -- parameter registration (spec pass) ---------------------------------------------------------
    def _reg_conv(self, name, k, cin, cout):
        if self.spec_only and name + "/kernel" not in self.net.params.spec:
            self.net.params.add(name + "/kernel", (k, k, cin, cout), True, "glorot")
            self.net.params.add(name + "/bias", (cout,), True, "zeros")

    def _reg_bn(self, name, c):
        if self.spec_only and name + "/gamma" not in self.net.params.spec:
            self.net.params.add(name + "/gamma", (c,), True, "ones")
            self.net.params.add(name + "/beta", (c,), True, "zeros")
            self.net.params.add(name + "/moving_mean", (c,), False, "zeros")
            self.net.params.add(name + "/moving_variance", (c,), False, "ones")

    def _new_stats(self, t, count):
        if self.training and not self.spec_only:
            t.stats = self.zeroed(2 * t.C)
            t.count = float(count)

    # ---------------------------------------------------------------------------------------------
    # generic 1x1 convolution over gathered / concatenated inputs
    # ---------------------------------------------------------------------------------------------
    def conv1x1(self, inputs, cout, name, stats=False, out_dtype=None, bias_grad=True, relu=False):
        """inputs: list of (T, mode, relu_in); mode = 'plain' | 's2' | ('up', s).
        keras: Conv2D(cout,(1,1)) on Concatenate/UpSampling2D/strides=2 inputs
        (model2.py:37,84,92,101-111,159-187).  relu=True fuses the consumer's Activation('relu') into the
        epilogue (ReLU commutes with nearest up-sampling)."""
        k_total = sum(t.C for t, _, _ in inputs)
        self._reg_conv(name, 1, k_total, cout)
        t0, m0, _ = inputs[0]
        if m0 == "plain":
            Ho, Wo = t0.H, t0.W
        elif m0 == "s2":
            Ho, Wo = (t0.H + 1) // 2, (t0.W + 1) // 2
        else:
            Ho, Wo = t0.H << m0[1], t0.W << m0[1]
        out = self.tensor(name, Ho, Wo, cout, out_dtype)
        if self.spec_only:
            return out
        if stats:
            self._new_stats(out, out.M)
        out.relu_masked = relu
        lib, N = self.lib, self.N
        W_ = self.P(name + "/kernel")
        b_ = self.P(name + "/bias")
        if self._thin_stem(inputs, cout, name, out, stats, bias_grad, relu):
            return out
        if self._thin_head(inputs, cout, name, out, stats, bias_grad, relu):
            return out
        pw = self._pw_classify(inputs, cout, name, out)
        if pw is not None:
            self._conv1x1_tc(inputs, pw, cout, name, out, stats, bias_grad, relu)
            return out
        segs, koff = [], 0
        for t, mode, relu_in in inputs:
            if mode == "plain":
                assert (t.H, t.W) == (Ho, Wo), (name, t.shape, Ho, Wo)
                s = Seg(t.data, t.C, t.H, t.W, relu_in=relu_in, w_off=koff * cout)
            elif mode == "s2":
                s = Seg(t.data, t.C, t.H, t.W, mult=2, relu_in=relu_in, w_off=koff * cout)
            else:
                assert (t.H << mode[1], t.W << mode[1]) == (Ho, Wo), (name, t.shape, Ho, Wo)
                s = Seg(t.data, t.C, t.H, t.W, shift=mode[1], relu_in=relu_in, w_off=koff * cout)
            segs.append(s)
            koff += t.C
        st = out.stats
        self.fwd.append(self._late(lambda: lib.igemm_fwd(segs, W_, cout, False, b_, out.data, N, Ho, Wo, cout,
                                                          stats=st[0] if st else None, relu=relu)))
        if self.training:
            self.tape.append(lambda: self._conv1x1_bwd(inputs, segs, out, name, cout, bias_grad))
        return out

    def _thin_stem(self, inputs, cout, name, out, stats, bias_grad, relu):
        """Stem Conv2D(32,(1,1)) on the raw (<= 16 band) input, model2.py:101: streaming thin kernel."""
        if len(inputs) != 1 or relu or cout != 32:
            return False
        t, mode, relu_in = inputs[0]
        if mode != "plain" or relu_in or t.C > 16 or t.dtype != out.dtype or not hasattr(self.lib, "stem_fwd"):
            return False
        lib, M, n = self.lib, out.M, t.C
        W_, b_ = self.P(name + "/kernel"), self.P(name + "/bias")
        st = out.stats
        esz = out.data.element_size()
        self.fwd.append(self._hb(self._late(lambda: lib.stem_fwd(t.data, W_, b_, out.data, M, n, st[0] if st else None)),
                                 M * (n + 32) * esz))
        if self.training:
            def bwd():
                if out.grad is None:
                    return
                assert not t.needs_grad, "the stem input carries no gradient"
                self.bwd.append(self._hb(lib.stem_wgrad(t.data, out.grad, M, n, self.G(name + "/kernel"),
                                                        self.G(name + "/bias") if bias_grad else None), M * (n + 32) * esz))
                self._ready(name + "/kernel", name + "/bias")
            self.tape.append(bwd)
        return True

    def _thin_head(self, inputs, cout, name, out, stats, bias_grad, relu):
        """Final Conv2D 1x1 of a head (32 -> n classes, fp32 logits, model2.py:159,168,180,186): streaming kernel."""
        if len(inputs) != 1 or relu or stats or cout > 8 or out.dtype != torch.float32 or not hasattr(self.lib, "head_fwd"):
            return False
        t, mode, relu_in = inputs[0]
        if mode != "plain" or relu_in or t.C != 32 or t.dtype != torch.bfloat16:
            return False
        W_, b_ = self.P(name + "/kernel"), self.P(name + "/bias")
        self.fwd.append(self._hb(self.lib.head_fwd(t.data, W_, b_, out.data, out.M, cout), out.M * (64 + 4 * cout)))
        if self.training:
            def bwd():
                if out.grad is None:
                    return
                ok = self._head_bwd(inputs, out, name, cout, bias_grad)
                assert ok, "head backward kernel unavailable"
            self.tape.append(bwd)
        return True

    def _head_bwd(self, inputs, out, name, cout, bias_grad):
        """Backward of a head's last 1x1 conv (32 -> n classes, fp32 logits): one fused streaming kernel."""
        if len(inputs) != 1 or out.dtype != torch.float32 or cout > 16 or not hasattr(self.lib, "head_bwd"):
            return False
        t, mode, relu_in = inputs[0]
        if mode != "plain" or relu_in or t.C != 32:
            return False
        g, acc = self.gacc(t) if t.needs_grad else (None, False)
        self.bwd.append(self._hb(self.lib.head_bwd(t.data, out.grad, self.P(name + "/kernel"), out.M, cout, g, acc,
                                                   t.relu_masked, self.G(name + "/kernel"),
                                                   self.G(name + "/bias") if bias_grad else None),
                                 out.M * (64 + 4 * cout + (64 * (2 if acc else 1) if g is not None else 0))))
        self._ready(name + "/kernel", name + "/bias")
        return True

    # ---------------------------------------------------------------------------------------------
    # tensor-core path of the 1x1 convolutions (bf16 mode)
    # ---------------------------------------------------------------------------------------------
    def _pw_classify(self, inputs, cout, name, out):
        """Split the sources of a 1x1 conv into `mains` (K-concatenated through TMA by conv_tc2) and `sides`
        (evaluated as their own small conv at the source resolution and added in the epilogue, using
        Conv1x1(Up(v)) == Up(Conv1x1(v))).  None -> CUDA-core implicit GEMM."""
        ent = self.net.tc.get(name)
        if ent is None or ent["taps"] != 1 or self.adt != torch.bfloat16:
            return None
        if out.H != out.W or out.W < 1 or (out.W & (out.W - 1)) or cout > 1024:
            return None
        mains, sides, koff = [], [], 0
        for t, mode, relu_in in inputs:
            if relu_in or t.dtype != torch.bfloat16:
                return None
            if mode in ("plain", "s2") and t.C % 16 == 0 and len(mains) < 2 and \
                    (not mains or (mains[-1][1] == mode == "plain" and mains[-1][2] + mains[-1][0].C == koff)):
                mains.append((t, mode, koff))
            elif mode == "s2":
                return None
            else:
                sides.append((t, 0 if mode == "plain" else mode[1], koff))
            koff += t.C
        if not mains or (mains[0][1] == "s2" and (sides or len(mains) > 1)):
            return None
        if sum(1 for _, sh, _ in sides if sh == 0) > 1 or sum(1 for _, sh, _ in sides if sh > 0) > 4:
            return None
        if out.dtype == torch.float32 and any(sh == 0 for _, sh, _ in sides):
            return None
        return mains, sides, koff

    @staticmethod
    def _tc_spatial_ok(H, W, min_side=1):
        return H == W and W >= min_side and (W & (W - 1)) == 0

    def _hb_stream(self, op, c0, c1, cout, M, out_dtype, sides=0.0):
        """A 1x1 launch that rsa_conv_tc2_fwd hands to the streaming kernel (pw_stream.cu: at most 64 input channels in total,
        8 / 16 / 32 / 64 output channels, bf16 out) is bandwidth-bound: tag its algorithmic bytes - operand rows read, result
        written, `sides` further result-shaped tensors read (mask, running sum, residual, up-sampled addends as fractions)."""
        if (out_dtype == torch.bfloat16 and c0 in (8, 16, 32, 64) and (c1 == 0 or (c0, c1) == (32, 32)) and cout in (8, 16, 32, 64)
                and M % 16 == 0 and os.environ.get("RSA_PW_STREAM", "1") != "0"):
            self._hb(op, 2.0 * M * (c0 + c1 + cout * (1.0 + sides)))
        return op

    def _pw_fwd_one(self, src, koff, K, cout, name, out, Hq, Wq):
        """q = W[koff:koff+C] . src at the source's own resolution (no bias)."""
        lib, N = self.lib, self.N
        ent = self.net.tc[name]
        if (src.C % 16 == 0 or src.C == 8) and self._tc_spatial_ok(Hq, Wq):
            wt = self.net.shadow[ent["fwd"]:ent["fwd"] + ent["coutp"] * K]
            return self._hb_stream(lib.conv_tc2_fwd(src.data, None, wt, ent["coutp"], None, out, N, Hq, Wq, cout, k_base=koff,
                                                    k_total=K), src.C, 0, cout, N * Hq * Wq, out.dtype)
        W_ = self.P(name + "/kernel")
        return lib.igemm_fwd([Seg(src.data, src.C, Hq, Wq, w_off=koff * cout)], W_, cout, False, None, out, N, Hq, Wq, cout)

    def _conv1x1_tc(self, inputs, pw, cout, name, out, stats, bias_grad, relu):
        lib, N = self.lib, self.N
        mains, sides, K = pw
        ent = self.net.tc[name]
        b_ = self.P(name + "/bias")
        wt = self.net.shadow[ent["fwd"]:ent["fwd"] + ent["coutp"] * K]
        qs = []
        for t, sh, koff in sides:
            q = self.alloc((N, t.H, t.W, cout), torch.bfloat16)
            self.fwd.append(self._pw_fwd_one(t, koff, K, cout, name, q, t.H, t.W))
            qs.append((t, sh, koff, q))
        ups = [(q, sh) for _, sh, _, q in qs if sh > 0]
        res = next((q for _, sh, _, q in qs if sh == 0), None)
        x0 = mains[0][0]
        x1 = mains[1][0] if len(mains) > 1 else None
        stride = 2 if mains[0][1] == "s2" else 1
        st = out.stats
        self.fwd.append(self._hb_stream(self._late(lambda: lib.conv_tc2_fwd(
            x0.data, x1.data if x1 is not None else None, wt, ent["coutp"], b_, out.data, N, out.H, out.W, cout,
            in_stride=stride, ups=ups, residual=res, stats=st[0] if st else None, relu=relu, k_base=mains[0][2],
            k_total=K)), x0.C, x1.C if x1 is not None else 0, cout, out.M, out.dtype,
            sides=(res is not None) + sum(0.25 ** sh for _, sh in ups)))
        if not self.training:
            return

        def bwd():
            if out.grad is None:
                return
            if self._head_bwd(inputs, out, name, cout, bias_grad):
                return
            dz = out.grad
            dW = self.G(name + "/kernel")
            W_ = self.P(name + "/kernel")
            wb = self.net.shadow[ent["bwd"]:ent["bwd"] + K * cout]
            bf = dz.dtype == torch.bfloat16
            p2 = lambda v: v >= 8 and (v & (v - 1)) == 0
            db_simt = [None]
            if bias_grad:
                if cout % 8 == 0 and bf:
                    self.bwd.append(self._side(self._hb(lib.bias_grad(dz, out.M, cout, [self.G(name + "/bias")]),
                                                        out.M * cout * dz.element_size()), dz))
                else:   # fp32 head logits / odd channel counts: folded into the CUDA-core wgrad of the first main
                    db_simt[0] = self.G(name + "/bias")
            sp = {}

            def one_source(t, koff, dq, Hq, Wq, in_stride):
                """weight + data gradient of one source given the gradient dq at the conv's own resolution."""
                sp_ok = self._tc_spatial_ok(Hq, Wq)
                if bf and p2(t.C) and p2(cout) and self._tc_spatial_ok(Hq, Wq, 4):
                    self.bwd.append(self._side(lib.pw_wgrad_tc(t.data, dq, dW[koff * cout:], cout, N, Hq, Wq, t.C, cout,
                                                               in_stride), t.data, dq))
                else:
                    seg = Seg(t.data, t.C, t.H, t.W, mult=in_stride, w_off=koff * cout)
                    db, db_simt[0] = (db_simt[0], None) if dq is dz else (None, db_simt[0])
                    self.bwd.append(lib.igemm_wgrad([seg], dq, dW, cout, db, N, Hq, Wq, cout))
                if not t.needs_grad:
                    return
                g, acc = self.gacc(t)
                mask = t.data if t.relu_masked else None
                ok8 = lambda v: v == 8 or v % 16 == 0
                if bf and ok8(cout) and ok8(t.C) and sp_ok:
                    wsl = wb[koff * cout:(koff + t.C) * cout]
                    if in_stride == 2 and not acc:
                        self.bwd.append(lambda s_, g=g: g.zero_())      # odd pixels receive no gradient
                    self.bwd.append(self._hb_stream(
                        lib.conv_tc2_fwd(dq, None, wsl, max(t.C, 16), None, g, N, Hq, Wq, t.C, mask=mask,
                                         accumulate=acc or in_stride == 2, out_stride=in_stride),
                        cout, 0, t.C, N * Hq * Wq, g.dtype, sides=(mask is not None) + bool(acc or in_stride == 2)))
                else:
                    if in_stride == 2:
                        sg = [Seg(dq, cout, Hq, Wq, shift=1, aligned=True, w_off=koff * cout)]
                    else:
                        sg = [Seg(dq, cout, Hq, Wq, w_off=koff * cout)]
                    self.bwd.append(lib.igemm_fwd(sg, W_, cout, True, None, g, N, t.H, t.W, t.C, mask=mask, accumulate=acc))

            for i, (t, mode, koff) in enumerate(mains):
                one_source(t, koff, dz, out.H, out.W, 2 if mode == "s2" else 1)
            need = sorted({sh for _, sh, _, _ in qs if sh > 0})
            if need:   # adjoint of nearest up-sampling: window sums of dz, all levels in one pass
                sp = {sh: self.alloc((N, out.H >> sh, out.W >> sh, cout), dz.dtype) for sh in need}
                self.bwd.append(self._hb(lib.sumpool_pyr(dz, N, out.H, out.W, cout, sp.get(1), sp.get(2), sp.get(3)),
                                         out.M * cout * dz.element_size() * (1 + sum(0.25 ** sh for sh in need))))
            for t, sh, koff, q in qs:
                one_source(t, koff, dz if sh == 0 else sp[sh], t.H, t.W, 1)
            assert db_simt[0] is None, "bias gradient was not emitted"
            self._ready(name + "/kernel", name + "/bias")
        self.tape.append(bwd)

    def _thin_dgrad(self, src, dy, wt, g, acc, mask, dil, N, H, W, C):
        """conv_tc3 data gradient into `src`; with RSA_BNR=1, when src = [relu](BatchNorm(x)) and this launch is the only
        writer of d(src), the BatchNormalization backward reductions ride in its epilogue (no separate reduction pass).
        Off by default since round 2: on the B200 the step with the fused sums takes 13.52 ms, with separate
        rsa_bn_bwd_reduce launches 13.42 ms (20 more launches, but they are bandwidth kernels that overlap with the other
        lane's convolutions, while the fused epilogue adds 15-30 us to a tensor-core kernel on the critical path)."""
        lib = self.lib
        b = src.bn_src
        if (b is not None and not acc and mask is None and b["x"].dtype == torch.bfloat16 and b["x"].shape == src.shape
                and os.environ.get("RSA_BNR", "0") == "1"):
            # provisional: the BatchNorm's own backward closure (emitted after every writer of d(src)) withdraws the
            # fusion when another launch accumulates into d(src) as well (two heads read the PSP output) - sums taken in
            # this epilogue would miss that contribution.  The launch is bound late, so it follows the final decision.
            b["fused"] = True
            return self._late(lambda: lib.conv_tc3_fwd(
                [dy], [wt], None, [-dil], g, N, H, W, C, stats=b["red"][0],
                bnr=(b["x"].data, b["xs"][0], b["cnt"], BN_EPS, b["gamma"], b["beta"], b["relu"]))
                if b["fused"] else lib.conv_tc3_fwd([dy], [wt], None, [-dil], g, N, H, W, C))
        return lib.conv_tc3_fwd([dy], [wt], None, [-dil], g, N, H, W, C, mask=mask, accumulate=acc)

    def _wide_dgrad(self, src, dy, wt, g, acc, mask, dil, N, H, W, C, Cdy):
        """conv_tc2 data gradient (C >= 128 layers) into `src`, with the same fusion as _thin_dgrad: for
        src = relu(BatchNorm(x)) written by this launch alone, the epilogue masks with the activated tensor (src > 0 is the
        ReLU mask), stores g and accumulates FusedBatchNormGrad's {sum g, sum g * xhat} from the BatchNorm input tile and
        the {mean, invstd} table of the forward statistics (written by the forward rsa_bn_apply launch).
        Off unless RSA_BNR_WIDE=1: measured on the B200 (profiles/r2_wide_bnr.txt) the 22 saved reduction launches do not pay
        for the longer register epilogue of conv_tc2 - the separate reductions overlap with convolutions on the other lane,
        the fused ones sit on the critical path (13.88 -> 14.12 ms per step)."""
        lib = self.lib
        b = src.bn_src
        plain = lambda: lib.conv_tc2_fwd(dy, None, wt, C, None, g, N, H, W, C, taps=9, dil=-dil, mask=mask, accumulate=acc)
        if (b is not None and not acc and mask is None and b["relu"] and b.get("coef") is not None and C % 32 == 0
                and b["x"].dtype == torch.bfloat16 and b["x"].shape == src.shape
                and os.environ.get("RSA_BNR_WIDE", "0") == "1"):
            b["fused"] = True
            return self._late(lambda: lib.conv_tc2_fwd(dy, None, wt, C, None, g, N, H, W, C, taps=9, dil=-dil, mask=src.data,
                                                       stats=b["red"][0], bnr_x=b["x"].data, bnr_coef=b["coef"])
                              if b["fused"] else plain())
        return plain()

    def _late(self, make):
        """Bind a launch lazily: scratch views (statistics) only exist after _finalize_scratch()."""
        cell = []

        def launch(stream):
            if not cell:
                cell.append(make())
            cell[0](stream)
        launch.make = make
        launch.cell = cell
        return launch

    def _conv1x1_bwd(self, inputs, segs, out, name, cout, bias_grad):
        lib, N = self.lib, self.N
        if out.grad is None:
            return   # nothing consumed this output
        if self._head_bwd(inputs, out, name, cout, bias_grad):
            return
        dy = out.grad
        W_ = self.P(name + "/kernel")
        dW = self.G(name + "/kernel")
        db = self.G(name + "/bias") if bias_grad else None
        self.bwd.append(lib.igemm_wgrad(segs, dy, dW, cout, db, N, out.H, out.W, cout))
        pyr = {}
        koff = 0
        for t, mode, relu_in in inputs:
            if t.needs_grad:
                g, acc = self.gacc(t)
                mask = t.data if (relu_in or t.relu_masked) else None
                woff = koff * cout
                if mode == "plain":
                    sg = [Seg(dy, cout, out.H, out.W, w_off=woff)]
                elif mode == "s2":
                    sg = [Seg(dy, cout, out.H, out.W, shift=1, aligned=True, w_off=woff)]
                elif mode[1] == 1:
                    sg = [Seg(dy, cout, out.H, out.W, mult=2, off_h=i, off_w=j, w_off=woff)
                          for i in (0, 1) for j in (0, 1)]
                else:
                    if not pyr:   # adjoint of nearest up-sampling = window sums of dy, one pass for all levels
                        need = sorted({m[1] for _, m, _ in inputs if isinstance(m, tuple) and m[1] > 1})
                        lv = {s: self.alloc((N, out.H >> s, out.W >> s, cout), dy.dtype) for s in need}
                        self.bwd.append(self._hb(lib.sumpool_pyr(dy, N, out.H, out.W, cout, lv.get(1), lv.get(2), lv.get(3)),
                                                 out.M * cout * dy.element_size() * (1 + sum(0.25 ** sh for sh in need))))
                        pyr.update(lv)
                    sg = [Seg(pyr[mode[1]], cout, t.H, t.W, w_off=woff)]
                self.bwd.append(lib.igemm_fwd(sg, W_, cout, True, None, g, N, t.H, t.W, t.C, mask=mask,
                                              accumulate=acc))
            koff += t.C
        self._ready(name + "/kernel", name + "/bias")         # after the data gradients, the last readers of the weights

    # ---------------------------------------------------------------------------------------------
    # 3x3 dilated 'same' convolution (ResBlock-a branches model2.py:19-24, heads :153-178)
    # ---------------------------------------------------------------------------------------------
    def conv3x3(self, x, cout, dil, name, relu=False, residual=None, out=None, stats=False, bias_grad=True):
        self._reg_conv(name, 3, x.C, cout)
        accumulate = out is not None
        if out is None:
            out = self.tensor(name, x.H, x.W, cout)
        if self.spec_only:
            return out
        if stats:
            self._new_stats(out, out.M)
        out.relu_masked = relu
        lib, N, H, W, C = self.lib, self.N, x.H, x.W, x.C
        W_ = self.P(name + "/kernel")
        b_ = self.P(name + "/bias")
        segs = [Seg(x.data, C, H, W, off_h=(ky - 1) * dil, off_w=(kx - 1) * dil, w_off=(ky * 3 + kx) * C * cout)
                for ky in range(3) for kx in range(3)]
        st = out.stats
        res = residual.data if residual is not None else None
        flops = 2.0 * N * H * W * 9 * C * cout
        nb = 2.0 * N * H * W * (C + cout)        # bf16 operand in + result out (fwd / dgrad) or two operands (wgrad)
        tcw = self.net.tc_weights(name, N, H, W)
        thin = tcw is not None and self.net.thin_ok(N, H, W, C, cout)
        if thin:
            self.fwd.append(self._tag(self._late(lambda: lib.conv_tc3_fwd(
                [x.data], [tcw[0]], [b_], [dil], out.data, N, H, W, C, residual=res,
                stats=st[0] if st else None, accumulate=accumulate, relu=relu)), "conv3x3_fwd", flops, nb))
        elif tcw is not None:
            self.fwd.append(self._tag(self._late(lambda: lib.conv_tc2_fwd(
                x.data, None, tcw[0], cout, b_, out.data, N, H, W, cout, taps=9, dil=dil, residual=res,
                stats=st[0] if st else None, accumulate=accumulate, relu=relu)), "conv3x3_fwd", flops, nb))
        else:
            self.fwd.append(self._tag(self._late(lambda: lib.igemm_fwd(
                segs, W_, cout, False, b_, out.data, N, H, W, cout, residual=res, stats=st[0] if st else None,
                accumulate=accumulate, relu=relu)), "conv3x3_fwd", flops, nb))
        if self.training:
            def bwd():
                if out.grad is None:
                    return
                dy = out.grad
                dW = self.G(name + "/kernel")
                db = self.G(name + "/bias") if bias_grad else None
                if tcw is not None and C == cout:
                    if thin and lib.conv_tc3_wgrad_supported(N, H, W, C, dil):
                        self.bwd.append(self._side(self._tag(lib.conv_tc3_wgrad(x.data, dy, dW, N, H, W, C, dil),
                                                             "conv3x3_wgrad", flops, nb), x.data, dy))
                    else:
                        self.bwd.append(self._side(self._tag(lib.conv_tc_wgrad(x.data, dy, dW, N, H, W, C, cout, dil),
                                                             "conv3x3_wgrad", flops, nb), x.data, dy))
                    if db is not None:
                        self.bwd.append(self._side(self._hb(lib.bias_grad(dy, N * H * W, cout, [db]),
                                                            N * H * W * cout * dy.element_size()), dy))
                else:
                    self.bwd.append(self._tag(lib.igemm_wgrad(segs, dy, dW, cout, db, N, H, W, cout),
                                              "conv3x3_wgrad", flops, nb))
                if x.needs_grad:
                    g, acc = self.gacc(x)
                    sg = [Seg(dy, cout, H, W, off_h=-(ky - 1) * dil, off_w=-(kx - 1) * dil,
                              w_off=(ky * 3 + kx) * C * cout) for ky in range(3) for kx in range(3)]
                    mask = x.data if x.relu_masked else None
                    if thin:
                        self.bwd.append(self._tag(self._thin_dgrad(x, dy, tcw[1], g, acc, mask, dil, N, H, W, C),
                                                  "conv3x3_dgrad", flops, nb))
                    elif tcw is not None:
                        self.bwd.append(self._tag(self._wide_dgrad(x, dy, tcw[1], g, acc, mask, dil, N, H, W, C, cout),
                                                  "conv3x3_dgrad", flops, nb))
                    else:
                        self.bwd.append(self._tag(lib.igemm_fwd(sg, W_, cout, True, None, g, N, H, W, C, mask=mask,
                                                                accumulate=acc), "conv3x3_dgrad", flops, nb))
                # after the data gradient: "ready" also means that no later launch READS this layer's weights (the step
                # updates them and refreshes their bf16 copies while backward still runs, keras_api._early_opt_plan)
                self._ready(name + "/kernel", name + "/bias")
            self.tape.append(bwd)
        return out

    # ---------------------------------------------------------------------------------------------
    # BatchNormalization (+ReLU), one input -> len(names) outputs (model2.py:17,21,38,86,93)
    # ---------------------------------------------------------------------------------------------
    def bn(self, x, names, relu, full_mult=1, derive=False):
        for n in names:
            self._reg_bn(n, x.C)
        outs = [self.tensor(n, x.H, x.W, x.C) for n in names]
        if self.spec_only:
            return outs
        lib, M, C = self.lib, x.M, x.C
        gam = [self.P(n + "/gamma") for n in names]
        bet = [self.P(n + "/beta") for n in names]
        mm = [self.P(n + "/moving_mean") for n in names]
        mv = [self.P(n + "/moving_variance") for n in names]
        if self.training:
            if x.stats is None:   # producer could not fuse the statistics: one extra pass
                self._new_stats(x, x.M)
                xs = x.stats
                self.fwd.append(self._late(lambda: lib.bn_stats(x.data, M, C, xs[0])))
            xs, cnt = x.stats, x.count
            for n in names:
                self.bn_table.append((xs, C, n + "/moving_mean", n + "/moving_variance", cnt, cnt * full_mult))
            esz = x.data.element_size()
            # C >= 128: the consumers' data gradients run on conv_tc2, whose fused BatchNorm-backward epilogue wants the
            # {mean, invstd} table (see _wide_dgrad); the first block of this launch writes it
            coef = None
            if relu and C >= 128 and x.dtype == torch.bfloat16 and self.net.conv_engine != "igemm_simt":
                coef = self.alloc((2 * C,), torch.float32)
            self.fwd.append(self._hb(self._late(lambda: lib.bn_apply(x.data, M, C, [o.data for o in outs], gam, bet, xs[0],
                                                                      cnt, None, None, BN_EPS, relu, coef)),
                                     (1 + len(outs)) * M * C * esz))
            if derive:   # statistics of y = gamma*xhat+beta are known in closed form
                o = outs[0]
                self._new_stats(o, o.M)
                os_ = o.stats
                self.fwd.append(self._late(lambda: lib.bn_derive_stats(xs[0], cnt, gam[0], bet[0], BN_EPS, os_[0],
                                                                        float(o.M), C)))
            reds = [self.zeroed(2 * C) for _ in names]
            for k, o in enumerate(outs):
                o.bn_src = dict(x=x, xs=xs, cnt=cnt, gamma=gam[k], beta=bet[k], relu=relu, red=reds[k], fused=False, coef=coef)

            def bwd():
                live = [k for k, o in enumerate(outs) if o.grad is not None]
                if not live:
                    return
                dys = [outs[k].grad for k in live]
                gl, bl, rl = [gam[k] for k in live], [bet[k] for k in live], [reds[k] for k in live]
                for k in live:       # several writers of d(out): the fused sums would be partial (see _thin_dgrad)
                    if outs[k].bn_src["fused"] and outs[k].grad_writers > 1:
                        outs[k].bn_src["fused"] = False
                # branches whose data-gradient kernel already produced {sum g, sum g*xhat} (conv_tc3 epilogue) are skipped
                unf = [i for i, k in enumerate(live) if not outs[k].bn_src["fused"]]
                if unf:
                    self.bwd.append(self._hb(self._late(lambda: lib.bn_bwd_reduce_multi(
                        [dys[i] for i in unf], x.data, M, C, xs[0], cnt, BN_EPS, [gl[i] for i in unf], [bl[i] for i in unf],
                        relu, [rl[i][0] for i in unf])), (len(unf) + 1) * M * C * esz))
                if x.needs_grad:
                    g, acc = self.gacc(x)
                    self.bwd.append(self._hb(self._late(lambda: lib.bn_bwd_apply_multi(
                        dys, x.data, M, C, xs[0], cnt, BN_EPS, gl, bl, relu, [r[0] for r in rl], g, acc,
                        [self.G(names[k] + "/gamma") for k in live], [self.G(names[k] + "/beta") for k in live])),
                        (len(live) + 2 + (1 if acc else 0)) * M * C * esz))
                    self._ready(*[names[k] + sfx for k in live for sfx in ("/gamma", "/beta")])
            self.tape.append(bwd)
        else:
            self.fwd.append(self._hb(lib.bn_apply(x.data, M, C, [o.data for o in outs], gam, bet, None, 1.0, mm, mv, BN_EPS,
                                                  relu), (1 + len(outs)) * M * C * x.data.element_size()))
        return outs

    # ---------------------------------------------------------------------------------------------
    def identity_add(self, x, out):
        """Backward of the identity term of Add([x, b_1..b_k]) (model2.py:27-31): x.grad += out.grad."""
        if self.spec_only or not self.training:
            return

        def bwd():
            if out.grad is None or not x.needs_grad:
                return
            if x.grad is None:
                x.grad = out.grad          # alias: branch gradients accumulate on top of d(out)
                x.grad_written = True
                x.grad_writers += 1
            else:
                g, acc = self.gacc(x)
                self.bwd.append(self._hb(self.lib.axpy(g, out.grad, x.M * x.C, acc),
                                         (3 if acc else 2) * x.M * x.C * g.element_size()))
        self.tape.append(bwd)

    def block_bias_grad(self, out, conv_names, f):
        """Tensor-core path: the second-conv biases of all branches of a ResBlock-a share d(out); one
        column-sum pass feeds them all (the SIMT wgrad kernel folds its own bias gradient instead)."""
        if self.spec_only or not self.training:
            return
        if self.net.tc_weights(conv_names[0], self.N, out.H, out.W) is None or f != out.C:
            return

        def bwd():
            if out.grad is None:
                return
            dbs = [self.G(n + "/bias") for n in conv_names]
            self.bwd.append(self._side(self._hb(self.lib.bias_grad(out.grad, out.M, out.C, dbs),
                                                out.M * out.C * out.grad.element_size()), out.grad))
            self._ready(*[n + "/bias" for n in conv_names])
        self.tape.append(bwd)

    def maxpool(self, x, levels):
        """MaxPooling2D(k, strides=k) for k in levels (subset of 2,4,8) in one pass (model2.py:47-52)."""
        outs = {k: self.tensor(f"pool{k}", x.H // k, x.W // k, x.C) for k in levels}
        if self.spec_only:
            return outs
        lib, N = self.lib, self.N
        d = lambda k: outs[k].data if k in outs else None
        esz = 2 if x.dtype == torch.bfloat16 else 4
        frac = sum(1.0 / (k * k) for k in levels)
        self.fwd.append(self._hb(lib.maxpool_pyr_fwd(x.data, N, x.H, x.W, x.C, d(2), d(4), d(8)), x.M * x.C * esz * (1 + frac)))
        if self.training:
            def bwd():
                gr = lambda k: outs[k].grad if k in outs else None
                if all(gr(k) is None for k in (2, 4, 8)) or not x.needs_grad:
                    return
                g, acc = self.gacc(x)
                self.bwd.append(self._hb(lib.maxpool_pyr_bwd(x.data, N, x.H, x.W, x.C, gr(2), gr(4), gr(8), g, acc),
                                         x.M * x.C * esz * (2 + frac + (1 if acc else 0))))
            self.tape.append(bwd)
        return outs

    def activation(self, z, kind):
        """softmax / sigmoid head activation in place on fp32 logits (model2.py:162,171,182,186)."""
        p = T(z.name + "/" + kind, z.N, z.H, z.W, z.C, z.dtype)
        p.data = z.data
        if self.spec_only:
            return p
        lib = self.lib
        if kind == "softmax":
            self.fwd.append(self._hb(lib.softmax_fwd(z.data, p.data, z.M, z.C), 8 * z.M * z.C))
        else:
            self.fwd.append(self._hb(lib.sigmoid_fwd(z.data, p.data, z.M * z.C), 8 * z.M * z.C))
        if self.training:
            def bwd():
                if p.grad is None:
                    return
                z.grad = p.grad
                z.grad_written = True
                if kind == "softmax":
                    self.bwd.append(self._hb(lib.softmax_bwd(p.data, p.grad, z.grad, z.M, z.C), 12 * z.M * z.C))
                else:
                    self.bwd.append(self._hb(lib.sigmoid_bwd(p.data, p.grad, z.grad, z.M * z.C), 12 * z.M * z.C))
            self.tape.append(bwd)
        return p

    # ---------------------------------------------------------------------------------------------
    # losses and metrics (train_ISPRS.py:411-452)
    # ---------------------------------------------------------------------------------------------
    def attach_loss(self, head, p, kind, weight, class_weights=None):
        """kind: 'tanimoto' | 'cce' | 'bce' | 'mse'.  Forward value lands in loss_out[head]."""
        lib = self.lib
        y = self.tensor("label/" + head, p.H, p.W, p.C, torch.float32)
        y.needs_grad = False
        self.labels[head] = y
        B, HW, C = p.N, p.H * p.W, p.C
        if self.res_f32 is None:
            self.res_f32 = torch.zeros(8, dtype=torch.float32, device=self.device)
            self.res_z = self.zeroed(16)
        slot = len(self.loss_out)
        rz = self.res_z
        if kind == "tanimoto":
            sums = self.zeroed(B * C * 5)
            res = self.res_f32[slot:slot + 1]
            coef = self.alloc((B * C * 3,), torch.float32)
            self.fwd.append(self._hb(self._late(lambda: lib.tanimoto_sums(p.data, y.data, B, HW, C, sums[0])), 8 * B * HW * C))
            self.fwd.append(self._late(lambda: lib.tanimoto_finalize(sums[0], B, HW, C, weight, None, res, coef)))
            self.loss_out[head] = ("mean", slot)
            if self.training:
                def bwd():
                    g, acc = self.gacc(p)
                    assert not acc
                    self.bwd.append(self._hb(lib.tanimoto_bwd(p.data, y.data, coef, B, HW, C, g), 12 * B * HW * C))
                self.tape.append(bwd)
        else:
            code = {"cce": 0, "bce": 1, "mse": 2}[kind]
            cw = None
            if class_weights is not None:
                cw = torch.tensor(list(class_weights), dtype=torch.float32).to(self.device)
                assert cw.numel() == C
            self.fwd.append(self._hb(self._late(lambda: lib.pixel_loss_fwd(code, p.data, y.data, cw, B * HW, C,
                                                                           rz[0][slot:slot + 1])), 8 * B * HW * C))
            self.loss_out[head] = ("sum", slot, float(B * HW))
            if self.training:
                def bwd():
                    g, acc = self.gacc(p)
                    assert not acc
                    self.bwd.append(self._hb(lib.pixel_loss_bwd(code, p.data, y.data, cw, B * HW, C, weight / (B * HW), g),
                                             12 * B * HW * C))
                self.tape.append(bwd)

    def attach_seg_metrics(self, p):
        y = self.labels["seg"]
        rz = self.res_z
        self.metrics = True
        self.fwd.append(self._hb(self._late(lambda: self.lib.seg_metrics(p.data, y.data, p.M, p.C,
                                                                          rz[0][8:13].view(torch.int64))), 8 * p.M * p.C))

    # ---------------------------------------------------------------------------------------------
    def finish(self):
        """Bind late launches, emit the backward list and the moving-statistics update."""
        if self.spec_only:
            return
        lib = self.lib
        if self.training:
            for rec in reversed(self.tape):
                rec()
        self._finalize_scratch()
        if self.training and self.bn_table:
            base = self.scratch
            pbase = self.net.params.data
            esz = base.element_size()
            tab, cnt = [], []
            for (st, C, mmn, mvn, n, nfull) in self.bn_table:
                soff = (st[0].data_ptr() - base.data_ptr()) // esz
                tab += [soff, C, self.net.params.off[mmn], self.net.params.off[mvn]]
                cnt += [n, nfull]
            self._bn_tab = torch.tensor(tab, dtype=torch.int64).to(self.device)
            self._bn_cnt = torch.tensor(cnt, dtype=torch.float64).to(self.device)
            self.bn_update = lib.bn_update_moving(base, pbase, self._bn_tab, self._bn_cnt, len(self.bn_table),
                                                  BN_MOMENTUM)
        # resolve late bindings now so that the first step is not special (and graph capture is clean)
        for op in self.fwd + self.bwd:
            if hasattr(op, "make") and not op.cell:
                op.cell.append(op.make())
                op.kernel = getattr(op.cell[0], "kernel", "?")

    def zero_scratch(self, stream):
        self.scratch.zero_()


# ------------------------------------------------------------------------------------------------------
# network definition (shared by the spec pass and every plan)
# ------------------------------------------------------------------------------------------------------
def _resblock(pl, x, f, dils, identity, relu_out=False):
    """ResBlock-a: x + Σ_d Conv3x3_d(ReLU(BN(Conv3x3_d(ReLU(BN_d(x)))))) (model2.py:15-34);
    model.py:15-33 has no identity term."""
    names = [(pl.name_bn(), pl.name_conv(), pl.name_bn(), pl.name_conv()) for _ in dils]
    for nb1, nc1, nb2, nc2 in names:     # register in keras creation order (branch by branch)
        pl._reg_bn(nb1, x.C)
        pl._reg_conv(nc1, 3, x.C, f)
        pl._reg_bn(nb2, f)
        pl._reg_conv(nc2, 3, f, f)
    a1 = pl.bn(x, [n[0] for n in names], relu=True)     # every branch normalises the same input
    out = pl.tensor("resblock_out", x.H, x.W, f)
    pl.block_bias_grad(out, [n[3] for n in names], f)
    if identity:
        pl.identity_add(x, out)
    # C = 32 blocks on the tensor-core path: the second convolutions of the branches are issued as fused launches whose
    # branches accumulate in the same TMEM tiles (north_star (1), model2.py:23-31) - the halo-mode dilations (1, 3) with the
    # identity input in one launch, the box-mode ones (15, 31) in a second that adds onto it: `out` is written twice instead
    # of four times and rounded to bf16 half as often.  (All four in one launch measured no faster than four single ones:
    # 72 KB of resident weights leave too shallow an operand ring, profiles/r2_bench_conv.txt.)
    fuse = (not pl.spec_only and len(dils) > 1 and f == x.C == 32 and os.environ.get("RSA_FUSE_BRANCHES", "1") != "0"
            and pl.net.tc_weights(names[0][3], pl.N, x.H, x.W) is not None and pl.net.thin_ok(pl.N, x.H, x.W, x.C, f))
    pending = []
    for i, d in enumerate(dils):
        with pl.lane(i if len(dils) > 1 else None):
            y1 = pl.conv3x3(a1[i], f, d, names[i][1], stats=True, bias_grad=False)   # bias before BN: zero gradient
            a2 = pl.bn(y1, [names[i][2]], relu=True)[0]
            piece = _conv_into(pl, a2, f, d, names[i][3], out, first=(i == 0), residual=x if identity else None,
                               relu=relu_out and i == len(dils) - 1, emit_fwd=not fuse)
            pending.append((d, piece))
    if fuse:
        groups = [[p_ for p_ in pending if abs(p_[0]) <= 3], [p_ for p_ in pending if abs(p_[0]) > 3]]
        groups = [g for g in groups if g]
        lib, N, H, W, C = pl.lib, pl.N, x.H, x.W, x.C
        for gi, grp in enumerate(groups):
            xs, wts, bs, ds = [q[1][0] for q in grp], [q[1][1] for q in grp], [q[1][2] for q in grp], [q[0] for q in grp]
            first, last = gi == 0, gi == len(groups) - 1
            flops = 2.0 * N * H * W * 9 * C * f * len(grp)
            nb = 2.0 * N * H * W * (C * len(grp) + f)
            op = lib.conv_tc3_fwd(xs, wts, bs, ds, out.data, N, H, W, C, residual=x.data if (first and identity) else None,
                                  accumulate=not first, relu=relu_out and last)
            op.convs = len(grp)
            pl.fwd.append(pl._tag(op, "conv3x3_fwd", flops, nb))
            pl.fwd[-1].chain = out.data.data_ptr()
    if relu_out:
        # the block's only consumer applies Activation('relu') (combine, model2.py:82): fused into the last
        # branch's epilogue; the stored tensor is relu(out) and gradients written to it are masked with it
        out.relu_masked = True
    return out


def _conv_into(pl, a, f, d, name, out, first, residual, relu=False, emit_fwd=True):
    """Second conv of a branch: writes (first) or accumulates into the block output; the first one
    also adds the identity input so that the branch sum + identity never exist as separate tensors.
    emit_fwd=False: the caller issues the forward as part of a fused multi-branch launch and gets
    (input, packed forward weights, bias) back; the backward launches are emitted here either way."""
    pl._reg_conv(name, 3, a.C, f)
    if pl.spec_only:
        return None
    lib, N, H, W, C = pl.lib, pl.N, a.H, a.W, a.C
    W_ = pl.P(name + "/kernel")
    b_ = pl.P(name + "/bias")
    segs = [Seg(a.data, C, H, W, off_h=(ky - 1) * d, off_w=(kx - 1) * d, w_off=(ky * 3 + kx) * C * f)
            for ky in range(3) for kx in range(3)]
    res = residual.data if (first and residual is not None) else None
    flops = 2.0 * N * H * W * 9 * C * f
    nb = 2.0 * N * H * W * (C + f)
    tcw = pl.net.tc_weights(name, N, H, W)
    thin = tcw is not None and pl.net.thin_ok(N, H, W, C, f)
    if not emit_fwd:
        assert thin
    elif thin:
        pl.fwd.append(pl._tag(lib.conv_tc3_fwd([a.data], [tcw[0]], [b_], [d], out.data, N, H, W, C, residual=res,
                                               accumulate=not first, relu=relu), "conv3x3_fwd", flops, nb))
    elif tcw is not None:
        pl.fwd.append(pl._tag(lib.conv_tc2_fwd(a.data, None, tcw[0], f, b_, out.data, N, H, W, f, taps=9, dil=d,
                                               residual=res, accumulate=not first, relu=relu), "conv3x3_fwd", flops, nb))
    else:
        pl.fwd.append(pl._tag(lib.igemm_fwd(segs, W_, f, False, b_, out.data, N, H, W, f, residual=res,
                                            accumulate=not first, relu=relu), "conv3x3_fwd", flops, nb))
    if emit_fwd:
        pl.fwd[-1].chain = out.data.data_ptr()     # the branch sum is a read-modify-write sequence on `out`
    if pl.training:
        def bwd():
            if out.grad is None:
                return
            dy = out.grad
            if tcw is not None and C == f:
                # bias gradient: one column-sum of d(out) per block (block_bias_grad)
                if thin and lib.conv_tc3_wgrad_supported(N, H, W, C, d):
                    pl.bwd.append(pl._side(pl._tag(lib.conv_tc3_wgrad(a.data, dy, pl.G(name + "/kernel"), N, H, W, C, d),
                                                   "conv3x3_wgrad", flops, nb), a.data, dy))
                else:
                    pl.bwd.append(pl._side(pl._tag(lib.conv_tc_wgrad(a.data, dy, pl.G(name + "/kernel"), N, H, W, C, f, d),
                                                   "conv3x3_wgrad", flops, nb), a.data, dy))
            else:
                pl.bwd.append(pl._tag(lib.igemm_wgrad(segs, dy, pl.G(name + "/kernel"), f, pl.G(name + "/bias"), N,
                                                      H, W, f), "conv3x3_wgrad", flops, nb))
            g, acc = pl.gacc(a)
            sg = [Seg(dy, f, H, W, off_h=-(ky - 1) * d, off_w=-(kx - 1) * d, w_off=(ky * 3 + kx) * C * f)
                  for ky in range(3) for kx in range(3)]
            if thin:
                pl.bwd.append(pl._tag(pl._thin_dgrad(a, dy, tcw[1], g, acc, None, d, N, H, W, C), "conv3x3_dgrad", flops, nb))
            elif tcw is not None:
                pl.bwd.append(pl._tag(pl._wide_dgrad(a, dy, tcw[1], g, acc, None, d, N, H, W, C, f), "conv3x3_dgrad", flops, nb))
            else:
                pl.bwd.append(pl._tag(lib.igemm_fwd(sg, W_, f, True, None, g, N, H, W, C, accumulate=acc),
                                      "conv3x3_dgrad", flops, nb))
            pl._ready(name + "/kernel", name + "/bias")      # after the data gradient, the last reader of the weights
        pl.tape.append(bwd)
    return (a.data, tcw[0], b_) if thin else None


def _psp(pl, x, f, v2):
    """PSPPooling (model2.py:41-79 / model.py:35-64) at pooled resolution, no concat buffer."""
    levels = psp_levels(pl.net.img_width)
    pooled = {1: x}
    if len(levels) > 1:
        pooled.update(pl.maxpool(x, [k for k in levels if k > 1]))
    br = []
    for k in levels:
        cn = pl.name_conv()
        c = pl.conv1x1([(pooled[k], "plain", False)], f // 4, cn, stats=v2, bias_grad=not v2)
        if v2:
            c = pl.bn(c, [pl.name_bn()], relu=False, full_mult=k * k)[0]
        s = int(math.log2(k))
        br.append((c, "plain" if s == 0 else ("up", s), False))
    cn = pl.name_conv()
    y = pl.conv1x1(br + [(x, "plain", False)], f, cn, stats=v2, bias_grad=not v2)
    if v2:
        y = pl.bn(y, [pl.name_bn()], relu=True)[0]   # Conv2DN then the call-site ReLU (model2.py:116,142)
    return y


def define_network(pl):
    net = pl.net
    v2 = net.variant == "v2"
    n = net.num_classes
    x = pl.tensor("input", net.img_height, net.img_width, net.img_channel)
    x.needs_grad = False
    pl.input = x
    t = pl.conv1x1([(x, "plain", False)], 32, pl.name_conv(), stats=True)           # stem, model2.py:101
    skips = [t]
    for lvl, (f, dils) in enumerate(zip(FILTERS, DILATIONS)):
        if lvl > 0:
            t = pl.conv1x1([(t, "s2", False)], f, pl.name_conv(), stats=True)       # model2.py:103-111
        t = _resblock(pl, t, f, dils, identity=v2)
        if lvl < 5:
            skips.append(t)
    t = _psp(pl, t, 1024, v2)
    for lvl in range(4, -1, -1):
        f, dils = FILTERS[lvl], DILATIONS[lvl]
        skip = skips[lvl + 1]
        if v2:
            # UpSampling: up2 -> conv(f/2) -> BN evaluated before the up-sampling (model2.py:89-94),
            # then combine: ReLU, concat skip, conv(f), BN (model2.py:81-87)
            u = pl.conv1x1([(t, "plain", False)], f // 2, pl.name_conv(), stats=True, bias_grad=False)
            v = pl.bn(u, [pl.name_bn()], relu=True, full_mult=4)[0]
            z = pl.conv1x1([(v, ("up", 1), False), (skip, "plain", False)], f, pl.name_conv(), stats=True,
                           bias_grad=False)
            t = pl.bn(z, [pl.name_bn()], relu=False, derive=True)[0]
        else:
            # model.py:93-94 Conv1x1(f) -> UpSampling2D; combine's ReLU (model.py:67) fused into this conv
            u = pl.conv1x1([(t, "plain", False)], f, pl.name_conv(), relu=True)
            t = pl.conv1x1([(u, ("up", 1), False), (skip, "plain", False)], f, pl.name_conv(), stats=True)
        t = _resblock(pl, t, f, dils, identity=v2, relu_out=(lvl == 0))
    xc = pl.conv1x1([(t, "plain", False), (skips[0], "plain", False)], 32, pl.name_conv(), stats=v2,
                    bias_grad=not v2)                                               # model2.py:140 (ReLU fused above)
    if v2:
        x_comb = pl.bn(xc, [pl.name_bn()], relu=False)[0]
    else:
        x_comb = xc
    x_psp = _psp(pl, x_comb, 32, v2)
    f32 = torch.float32
    if not net.multitask:
        z = pl.conv1x1([(x_psp, "plain", False)], n, pl.name_conv(), out_dtype=f32)
        pl.outputs["seg"] = pl.activation(z, "softmax")
        return
    # the four heads are independent chains (shared inputs are read-only, the shared input gradients are `chain`ed):
    # one lane each, like the ResBlock-a branches (HEAD_LANE is also used for their losses)
    with pl.lane(HEAD_LANE["seg"]):
        h = pl.conv3x3(x_psp, 32, 1, "seg1", relu=True)
        h = pl.conv3x3(h, 32, 1, "seg2", relu=True)
        z = pl.conv1x1([(h, "plain", False)], n, "seg3", out_dtype=f32)
        pl.outputs["seg"] = pl.activation(z, "softmax")
    with pl.lane(HEAD_LANE["bound"]):
        h = pl.conv3x3(x_psp, 32, 1, pl.name_conv(), relu=True)
        z = pl.conv1x1([(h, "plain", False)], n, pl.name_conv(), out_dtype=f32)
        pl.outputs["bound"] = pl.activation(z, "sigmoid")
    with pl.lane(HEAD_LANE["dist"]):
        h = pl.conv3x3(x_comb, 32, 1, pl.name_conv(), relu=True)
        h = pl.conv3x3(h, 32, 1, pl.name_conv(), relu=True)
        z = pl.conv1x1([(h, "plain", False)], n, pl.name_conv(), out_dtype=f32)
        pl.outputs["dist"] = pl.activation(z, "softmax")                                 # softmax, model2.py:182
    with pl.lane(HEAD_LANE["color"]):
        z = pl.conv1x1([(x_comb, "plain", False)], 3, "color", out_dtype=f32)
        pl.outputs["color"] = pl.activation(z, "sigmoid")


HEAD_LANE = {"seg": 0, "bound": 1, "dist": 2, "color": 3}


class Net:
    """Parameters + plan cache for one ResUnet-a instance."""

    def __init__(self, input_shape, num_classes, multitask, variant="v2", dtype="bf16", seed=1234, lib=None,
                 device=None):
        from . import _capi
        self.lib = lib or _capi.get_lib()
        self.img_height, self.img_width, self.img_channel = input_shape
        self.num_classes = num_classes
        self.multitask = bool(multitask)
        self.variant = variant
        self.act_dtype = {"bf16": torch.bfloat16, "fp32": torch.float32}[dtype]
        self.dtype_name = dtype
        if device is None:
            device = torch.device("cpu") if getattr(self.lib, "is_emulation", False) else torch.device(
                "cuda", torch.cuda.current_device())
        self.device = device
        if self.img_height % 32 or self.img_width % 32 or self.img_height != self.img_width:
            raise ValueError("ResUnet-a d6 needs square inputs with a side that is a multiple of 32")
        if (self.img_width // 32) % max(psp_levels(self.img_width)):
            raise ValueError(f"input width {self.img_width}: the 1/32-resolution map is not divisible by the "
                             f"PSPPooling levels {psp_levels(self.img_width)} (keras fails the same way)")
        self.params = ParamStore()
        spec = Plan(self, 1, True, spec_only=True)
        define_network(spec)
        self.output_names = list(spec.outputs)
        self.params.finalize(self.device, seed)
        self.plans = {}
        self._init_tensor_core_path()

    # -- bf16 tensor-core path: bf16 weight copies [tap][Cout][Cin] (fwd) / [tap][Cin][Cout] (dgrad) -------
    def _init_tensor_core_path(self):
        self.tc = {}
        self.shadow = None
        self.pack_launch = None
        self._pack_split = {}
        self.shadow_dirty = True
        engine = os.environ.get("RSA_CONV_ENGINE", "tc")
        emulated = getattr(self.lib, "is_emulation", False) and not getattr(self.lib, "emulates_tensor_core", False)
        if self.act_dtype != torch.bfloat16 or emulated or engine != "tc" or not hasattr(self.lib, "conv_tc2_fwd"):
            self.conv_engine = "igemm_simt"
            return
        okc = lambda c: c == 32 or (c >= 64 and c % 64 == 0)
        off, table, max_elems = 0, b"", 0
        for name, (shape, _, _) in self.params.spec.items():
            if not name.endswith("/kernel"):
                continue
            k, _, cin, cout = shape
            if k == 3:
                if not (okc(cin) and okc(cout)):
                    continue
                coutp = cout
            else:
                if cin < 16:
                    continue       # stem (3 or 14 input channels): CUDA-core path
                bn = 128 if cout >= 128 else (64 if cout >= 64 else (32 if cout >= 32 else 16))
                coutp = (cout + bn - 1) // bn * bn
            taps = k * k
            nf, nb = taps * coutp * cin, taps * cin * cout
            off = (off + 127) // 128 * 128
            self.tc[name[:-len("/kernel")]] = dict(fwd=off, bwd=off + nf, taps=taps, cin=cin, cout=cout, coutp=coutp)
            table += struct.pack("<qqqiiii", self.params.off[name], off, off + nf, taps, cin, cout, coutp)
            off += nf + nb
            off = (off + 127) // 128 * 128
            max_elems = max(max_elems, taps * cin * cout)
        if not self.tc:
            self.conv_engine = "igemm_simt"
            return
        self.shadow = torch.zeros(off, dtype=torch.bfloat16, device=self.device)
        self._pack_table = torch.frombuffer(bytearray(table), dtype=torch.uint8).to(self.device)
        self.pack_launch = self.lib.pack_weights_tc(self.params.data, self.shadow, self._pack_table, len(self.tc),
                                                    max_elems)
        self.pack_launch.hbm_bytes = 8.0 * sum(e["taps"] * e["cin"] * e["cout"] for e in self.tc.values())
        self.conv_engine = "tcgen05 (3x3 fwd/dgrad/wgrad persistent TMA kernels; 1x1 on tensor cores where K,N % 16 == 0)"

    def pack_launches_split(self, off):
        """The bf16 weight refresh as two launches: the layers whose kernels start at a flat parameter offset >= off and
        the rest - the step refreshes the first set as soon as the optimizer has updated those parameters, while the
        backward of the shallow levels still runs (keras_api._train_step_overlapped).  (hi, lo); either may be None."""
        key = ("split", int(off))
        if key not in self._pack_split:
            hi, lo = [], []
            for name, e in self.tc.items():
                ent = struct.pack("<qqqiiii", self.params.off[name + "/kernel"], e["fwd"], e["bwd"], e["taps"], e["cin"],
                                  e["cout"], e["coutp"])
                (hi if self.params.off[name + "/kernel"] >= off else lo).append((ent, e["taps"] * e["cin"] * e["cout"]))
            launches = []
            for part in (hi, lo):
                if not part:
                    launches.append(None)
                    continue
                tab = torch.frombuffer(bytearray(b"".join(t for t, _ in part)), dtype=torch.uint8).to(self.device)
                launches.append(self.lib.pack_weights_tc(self.params.data, self.shadow, tab, len(part), max(n for _, n in part)))
            self._pack_split[key] = tuple(launches)
        return self._pack_split[key]

    def tc_weights(self, name, N, H, W):
        """(fwd copy, dgrad copy) bf16 views if the tensor-core kernel handles this layer at this size."""
        ent = self.tc.get(name)
        if ent is None or ent["taps"] != 9 or not self.lib.conv_tc_supported(N, H, W, ent["cin"], ent["cout"]):
            return None
        n = 9 * ent["cin"] * ent["cout"]
        return self.shadow[ent["fwd"]:ent["fwd"] + n], self.shadow[ent["bwd"]:ent["bwd"] + n]

    def thin_ok(self, N, H, W, cin, cout):
        """conv_tc3 (thin-layer kernel: halo tiles, resident weights) handles this 3x3 convolution."""
        return (cin == cout and hasattr(self.lib, "conv_tc3_supported") and os.environ.get("RSA_TC3", "1") != "0"
                and self.lib.conv_tc3_supported(N, H, W, cin))

    def ensure_shadow(self, stream):
        if self.pack_launch is not None and self.shadow_dirty:
            self.pack_launch(stream)
            self.shadow_dirty = False

    def plan(self, N, training, loss_spec=None):
        """loss_spec: None or tuple of (head, kind, weight, class_weights) — part of the cache key."""
        key = (N, training, loss_spec)
        if key not in self.plans:
            pl = Plan(self, N, training, with_loss=loss_spec is not None)
            define_network(pl)
            pl.n_fwd_net = len(pl.fwd)      # forward launches before this index need only x (labels may still be in flight)
            if loss_spec is not None:
                for head, kind, weight, cw in loss_spec:
                    with pl.lane(HEAD_LANE.get(head) if self.multitask else None):
                        pl.attach_loss(head, pl.outputs[head], kind, weight, cw)
                pl.attach_seg_metrics(pl.outputs["seg"])
            pl.finish()
            self.plans[key] = pl
        return self.plans[key]

    # -- weights interchange (keras names, HWIO) ----------------------------------------------------------
    def get_weights(self):
        return OrderedDict((k, self.params.view(k).detach().cpu().clone()) for k in self.params.spec)

    def set_weights(self, weights):
        for k, v in weights.items():
            if k not in self.params.spec:
                raise KeyError(f"unknown parameter {k}")
            dst = self.params.view(k)
            v = torch.as_tensor(v, dtype=torch.float32)
            if tuple(v.shape) != tuple(dst.shape):
                raise ValueError(f"shape mismatch for {k}: {tuple(v.shape)} vs {tuple(dst.shape)}")
            dst.copy_(v.to(self.device))
        self.shadow_dirty = True
