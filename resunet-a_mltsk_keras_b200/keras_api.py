"""Keras-compatible surface of the B200 ResUnet-a: Model, optimizers, losses, load_model.

Mirrors what the reference scripts call on ``Resunet_a(...).model``:
``summary / compile / train_on_batch / test_on_batch / predict / fit / save / output_names /
optimizer.lr`` (train_ISPRS.py:148,186,292,445-452,478-480; test_ISPRS.py:26-36,278;
amazon_py/main_tcc.py:218).  Inputs and outputs are NHWC float32 numpy arrays with one-hot labels,
exactly like the Keras model; underneath every call runs the pre-bound CUDA launches of a
``graph.Plan``, optionally replayed as one CUDA graph.
"""
from __future__ import annotations

import json
import math
import os
from collections import OrderedDict

import numpy as np
import torch

from . import graph

_COPY_POOL = None
_HOST_REG = {}      # data pointer -> [nbytes, sightings, registered, weakref to the owning array]


def _registered_view(a):
    """Zero-copy path for host batches that are REUSED across steps (the reference refills the same x_train_b / y_*_b
    buffers every step, train_ISPRS.py:71-92,121-141): the second time the same float32 buffer is seen it is page-locked
    with cudaHostRegister, from then on the H2D copy reads the caller's memory directly and the staging memcpy disappears.
    Returns a torch view that is_pinned(), or None (fresh / small / foreign memory: staging path).  The registration is
    dropped when the owning array is garbage collected."""
    import weakref
    if os.environ.get("RSA_HOST_REGISTER", "1") == "0" or a.nbytes < (1 << 20) or not a.flags.owndata and a.base is None:
        return None
    owner = a
    while isinstance(owner.base, np.ndarray):
        owner = owner.base
    if owner.base is not None or not owner.flags.owndata:
        return None                               # memmap / foreign buffer: do not pin what we do not understand
    ptr, nb = a.ctypes.data, a.nbytes
    ent = _HOST_REG.get(ptr)
    if ent is None or ent[0] != nb or ent[3]() is not owner:
        if ent is not None and ent[2]:
            torch.cuda.cudart().cudaHostUnregister(ptr)
        _HOST_REG[ptr] = ent = [nb, 0, False, weakref.ref(owner)]
    ent[1] += 1
    if not ent[2]:
        if ent[1] < 2:
            return None
        try:
            rc = torch.cuda.cudart().cudaHostRegister(ptr, nb, 0)
        except Exception:
            rc = 1
        if int(rc) != 0:
            ent[1] = -(1 << 30)                   # never try this buffer again
            return None
        ent[2] = True

        def _drop(_ptr=ptr):
            e = _HOST_REG.pop(_ptr, None)
            if e is not None and e[2]:
                try:
                    torch.cuda.cudart().cudaHostUnregister(_ptr)
                except Exception:
                    pass
        weakref.finalize(owner, _drop)
    t = torch.from_numpy(a)
    return t if t.is_pinned() else None


def _host_copy(dst_np, src_np, min_chunk=4 << 20):
    """Parallel memcpy of a host batch into its pinned staging buffer.  Own thread pool (numpy releases the GIL while it
    copies): torch's intra-op threads are not usable for this under torchrun, which sets OMP_NUM_THREADS=1 per rank and
    turned the 100 MB/step staging copy into a 20+ ms single-thread memcpy at N > 1."""
    global _COPY_POOL
    n = dst_np.size
    nthreads = int(os.environ.get("RSA_COPY_THREADS", "8"))
    if nthreads <= 1 or n * dst_np.itemsize < 2 * min_chunk:
        np.copyto(dst_np, src_np)
        return
    if _COPY_POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _COPY_POOL = ThreadPoolExecutor(nthreads)
    d, s_ = dst_np.reshape(-1), src_np.reshape(-1)
    step = max(min_chunk // dst_np.itemsize, (n + nthreads - 1) // nthreads)
    futs = [_COPY_POOL.submit(np.copyto, d[i:i + step], s_[i:i + step]) for i in range(0, n, step)]
    for f in futs:
        f.result()


# ------------------------------------------------------------------------------------------------------
# optimizers (train_ISPRS.py:404-407)
# ------------------------------------------------------------------------------------------------------
class Optimizer:
    def __init__(self, lr):
        self.lr = float(lr)
        self.iterations = 0

    @property
    def learning_rate(self):
        return self.lr

    @learning_rate.setter
    def learning_rate(self, v):
        self.lr = float(v)


class Adam(Optimizer):
    """keras Adam(lr, beta_1=.9, beta_2=.999, epsilon=1e-7, amsgrad=False)."""

    def __init__(self, lr=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7, learning_rate=None):
        super().__init__(lr if learning_rate is None else learning_rate)
        self.beta_1, self.beta_2, self.epsilon = beta_1, beta_2, epsilon

    def config(self):
        return dict(kind="adam", lr=self.lr, beta_1=self.beta_1, beta_2=self.beta_2, epsilon=self.epsilon)


class SGD(Optimizer):
    """keras SGD(lr, momentum): v <- momentum*v - lr*g ; p <- p + v."""

    def __init__(self, lr=1e-2, momentum=0.0, learning_rate=None):
        super().__init__(lr if learning_rate is None else learning_rate)
        self.momentum = momentum

    def config(self):
        return dict(kind="sgd", lr=self.lr, momentum=self.momentum)


# ------------------------------------------------------------------------------------------------------
# losses: factories return callables tagged with the kernel that implements them
# ------------------------------------------------------------------------------------------------------
def _run_loss_standalone(kind, y_true, y_pred, class_weights=None):
    """Evaluate a loss on the device outside a model (what calling the keras loss fn directly does)."""
    from . import _capi
    lib = _capi.get_lib()
    dev = torch.device("cpu") if getattr(lib, "is_emulation", False) else torch.device("cuda")
    p = torch.as_tensor(np.asarray(y_pred), dtype=torch.float32).to(dev).contiguous()
    y = torch.as_tensor(np.asarray(y_true), dtype=torch.float32).to(dev).contiguous()
    B, H, W, C = p.shape
    stream = 0 if dev.type == "cpu" else torch.cuda.current_stream().cuda_stream
    if kind == "tanimoto":
        sums = torch.zeros(B * C * 5, dtype=torch.float64, device=dev)
        lb = torch.empty(B, dtype=torch.float32, device=dev)
        lib.tanimoto_sums(p, y, B, H * W, C, sums)(stream)
        lib.tanimoto_finalize(sums, B, H * W, C, 1.0, lb, None, None)(stream)
        return lb.cpu().numpy()
    raise NotImplementedError("element-wise losses are evaluated inside the model (loss value returned by "
                              "train_on_batch/test_on_batch)")


def Tanimoto_dual_loss():
    """multitasking_utils.py:71-85 — returns ``loss(label, pred) -> [B]``."""
    def loss(label, pred):
        return _run_loss_standalone("tanimoto", label, pred)
    loss.rsa_kind = "tanimoto"
    loss.rsa_class_weights = None
    return loss


def weighted_categorical_crossentropy(weights):
    """utils.py:466-491 — returns ``loss(y_true, y_pred) -> [B,H,W]``."""
    w = [float(v) for v in np.asarray(weights).ravel()]

    def loss(y_true, y_pred):
        p = np.asarray(y_pred, dtype=np.float64)
        p = p / p.sum(-1, keepdims=True)
        p = np.clip(p, 1e-7, 1 - 1e-7)
        return -(np.asarray(y_true) * np.log(p) * np.asarray(w)).sum(-1)
    loss.rsa_kind = "cce"
    loss.rsa_class_weights = tuple(w)
    return loss


class _KerasLoss:
    rsa_class_weights = None


class CategoricalCrossentropy(_KerasLoss):
    rsa_kind = "cce"


class BinaryCrossentropy(_KerasLoss):
    rsa_kind = "bce"


class MeanSquaredError(_KerasLoss):
    rsa_kind = "mse"


_LOSS_STRINGS = {"categorical_crossentropy": "cce", "binary_crossentropy": "bce", "mse": "mse",
                 "mean_squared_error": "mse", "tanimoto": "tanimoto"}


def _resolve_loss(l):
    if isinstance(l, str):
        if l not in _LOSS_STRINGS:
            raise ValueError(f"unsupported loss '{l}'")
        return _LOSS_STRINGS[l], None
    kind = getattr(l, "rsa_kind", None)
    if kind is None:
        raise ValueError("loss must be one of Tanimoto_dual_loss(), weighted_categorical_crossentropy(w), "
                         "CategoricalCrossentropy(), BinaryCrossentropy(), MeanSquaredError() or a keras loss "
                         "name: arbitrary Python callables cannot be lowered to the fused CUDA loss kernels")
    return kind, getattr(l, "rsa_class_weights", None)


# ------------------------------------------------------------------------------------------------------
# callbacks used by the reference's fit() call (amazon_py/main_tcc.py:212-218)
# ------------------------------------------------------------------------------------------------------
class EarlyStopping:
    def __init__(self, monitor="val_loss", min_delta=0.0, patience=0, verbose=0, mode="min"):
        self.monitor, self.min_delta, self.patience = monitor, min_delta, patience
        self.best, self.wait = math.inf, 0

    def on_epoch_end(self, model, epoch, logs):
        cur = logs.get(self.monitor)
        if cur is None:
            return False
        if cur < self.best - self.min_delta:
            self.best, self.wait = cur, 0
            return False
        self.wait += 1
        return self.wait >= self.patience


class ModelCheckpoint:
    def __init__(self, filepath, monitor="val_loss", verbose=0, save_best_only=False, mode="min"):
        self.filepath, self.monitor, self.save_best_only = filepath, monitor, save_best_only
        self.best = math.inf

    def on_epoch_end(self, model, epoch, logs):
        cur = logs.get(self.monitor)
        if not self.save_best_only or (cur is not None and cur < self.best):
            if cur is not None:
                self.best = min(self.best, cur)
            model.save(self.filepath)
        return False


class History:
    def __init__(self):
        self.history = {}

    def add(self, logs):
        for k, v in logs.items():
            self.history.setdefault(k, []).append(v)


# ------------------------------------------------------------------------------------------------------
class Model:
    """The object the reference obtains as ``Resunet_a(...).model``."""

    def __init__(self, net: graph.Net, config: dict):
        self.net = net
        self.config = config
        self.output_names = list(net.output_names)
        self.optimizer = None
        self.loss_spec = None
        self.loss_weights = {}
        self.metrics_names = []
        self._opt_state = None
        self._graphs = {}
        self.use_cuda_graph = os.environ.get("RSA_CUDA_GRAPH", "1") != "0"
        self._staging = {}
        self.dp = None          # set by distribute.MirroredStrategy.scope()
        self._lr_host = None
        self._lr_dev = None

    # -- introspection ---------------------------------------------------------------------------------
    def count_params(self):
        return sum(math.prod(s) for s, _, _ in self.net.params.spec.values())

    def summary(self, print_fn=print):
        ps = self.net.params.spec
        print_fn(f'Model: "ResUnet-a d6 ({self.net.variant}, {"multitask" if self.net.multitask else "single-task"})"')
        print_fn(f"{'Layer (keras name)':40s}{'Param shape':28s}{'# Params':>10s}")
        for name, (shape, tr, _) in ps.items():
            print_fn(f"{name:40s}{str(shape):28s}{math.prod(shape):>10d}")
        tot = self.count_params()
        tr = sum(math.prod(s) for s, t, _ in ps.values() if t)
        print_fn(f"Total params: {tot:,}\nTrainable params: {tr:,}\nNon-trainable params: {tot - tr:,}")

    # -- compile ---------------------------------------------------------------------------------------------
    def compile(self, optimizer=None, loss=None, loss_weights=None, metrics=None):
        if isinstance(optimizer, str):
            optimizer = {"adam": Adam, "sgd": SGD}[optimizer.lower()]()
        self.optimizer = optimizer
        heads = self.output_names
        if isinstance(loss, dict):
            missing = [h for h in heads if h not in loss]
            if missing:
                raise ValueError(f"no loss given for outputs {missing}")
            per = {h: loss[h] for h in heads}
        else:
            per = {h: loss for h in heads}
        lw = dict(loss_weights or {})
        spec = []
        for h in heads:
            kind, cw = _resolve_loss(per[h])
            spec.append((h, kind, float(lw.get(h, 1.0)), cw))
        self.loss_spec = tuple(spec)
        self.loss_weights = {h: float(lw.get(h, 1.0)) for h in heads}
        if self.net.multitask:
            self.metrics_names = ["loss"] + [f"{h}_loss" for h in heads] + [
                "seg_accuracy", "seg_true_positives", "seg_false_positives", "seg_true_negatives",
                "seg_false_negatives"]
        else:
            self.metrics_names = ["loss", "accuracy", "true_positives", "false_positives", "true_negatives",
                                  "false_negatives"]
        self._graphs.clear()
        from . import distribute
        strat = distribute.current_strategy()
        if strat is not None and strat.dp.world_size > 1:
            self.dp = strat.dp
            self.dp.broadcast_parameters(self.net.params)
            self._opt_state = None

    # -- data movement -------------------------------------------------------------------------------------
    def _stream(self):
        return 0 if self.net.device.type == "cpu" else torch.cuda.current_stream().cuda_stream

    def _stage(self, key, arr, dst):
        """host numpy (any float) -> pinned fp32 staging -> device tensor `dst` (fp32) asynchronously.  A float32 torch
        tensor that already lives in pinned memory (data.PatchBatchLoader) is copied from directly."""
        if isinstance(arr, torch.Tensor):
            if (arr.dtype == torch.float32 and arr.is_contiguous() and tuple(arr.shape) == tuple(dst.shape)
                    and arr.device.type == "cpu" and self.net.device.type != "cpu" and arr.is_pinned()):
                dst.copy_(arr, non_blocking=True)
                return arr.numel() * 4
            arr = arr.detach().cpu().numpy()
        a = np.ascontiguousarray(arr, dtype=np.float32)
        if tuple(a.shape) != tuple(dst.shape):
            raise ValueError(f"{key}: expected shape {tuple(dst.shape)}, got {tuple(a.shape)}")
        if self.net.device.type == "cpu":
            dst.copy_(torch.from_numpy(a))
            return a.nbytes
        if a is arr or (isinstance(arr, np.ndarray) and np.shares_memory(a, arr)):
            rv = _registered_view(a)
            if rv is not None:
                dst.copy_(rv, non_blocking=True)
                return a.nbytes
        st = self._staging.get((key, a.shape))
        if st is None:
            st = torch.empty(a.shape, dtype=torch.float32, pin_memory=True)
            self._staging[(key, a.shape)] = st
        _host_copy(st.numpy(), a)              # multi-threaded host copy into the pinned staging buffer
        dst.copy_(st, non_blocking=True)
        return a.nbytes

    def _load_inputs(self, pl, x, y):
        return self._load_x(pl, x) + self._load_labels(pl, y)

    def _load_x(self, pl, x):
        nb = 0
        if pl.input.dtype == torch.float32:
            nb += self._stage("x", x, pl.input.data)
        else:
            xf = getattr(pl, "_x_f32", None)
            if xf is None:
                xf = pl._x_f32 = torch.empty(pl.input.shape, dtype=torch.float32, device=self.net.device)
                pl._x_cast = self.net.lib.cast(xf, pl.input.data, xf.numel())
            nb += self._stage("x", x, xf)
            pl._x_cast(self._stream())
        return nb

    def _load_labels(self, pl, y):
        nb = 0
        if y is not None:
            if not isinstance(y, dict):
                y = {self.output_names[0]: y}
            for h, t in pl.labels.items():
                if h not in y:
                    raise ValueError(f"missing labels for output '{h}'")
                nb += self._stage("y/" + h, y[h], t.data)
        return nb

    # -- optimizer step ---------------------------------------------------------------------------------------
    def _ensure_opt(self):
        if self.optimizer is None:
            raise RuntimeError("compile() the model with an optimizer before training")
        ps = self.net.params
        if self._opt_state is None:
            dev = self.net.device
            z = lambda: torch.zeros(ps.n_train, dtype=torch.float32, device=dev)
            self._opt_state = dict(m=z(), v=z()) if isinstance(self.optimizer, Adam) else dict(vel=z())
            saved = getattr(self, "_saved_opt_state", None)
            if saved:
                for k, v in saved.items():
                    if k in self._opt_state:
                        self._opt_state[k].copy_(torch.from_numpy(np.asarray(v)).to(dev))
            self._lr_dev = torch.zeros(1, dtype=torch.float32, device=dev)
            self._lr_host = torch.zeros(64, dtype=torch.float32, pin_memory=dev.type != "cpu")
            lib, opt, n = self.net.lib, self.optimizer, ps.n_train
            gs = 1.0 / (self.dp.world_size if self.dp else 1)
            if isinstance(opt, Adam):
                self._opt_launch = lib.adam_step(ps.data, ps.grad, self._opt_state["m"], self._opt_state["v"], n,
                                                 self._lr_dev, opt.beta_1, opt.beta_2, opt.epsilon, gs)
            else:
                self._opt_launch = lib.sgd_step(ps.data, ps.grad, self._opt_state["vel"], n, self._lr_dev,
                                                opt.momentum, gs)

    def _push_lr(self):
        opt = self.optimizer
        opt.iterations += 1
        if isinstance(opt, Adam):
            t = opt.iterations
            lr_t = opt.lr * math.sqrt(1.0 - opt.beta_2 ** t) / (1.0 - opt.beta_1 ** t)
        else:
            lr_t = opt.lr
        slot = opt.iterations % 64      # ring of pinned slots: up to 64 steps may be in flight
        self._lr_host[slot] = lr_t
        self._lr_dev.copy_(self._lr_host[slot:slot + 1], non_blocking=True)

    # -- the step itself ----------------------------------------------------------------------------------------
    def _run_fwd_bwd(self, pl, stream):
        pl.scratch.zero_()
        self.net.params.grad.zero_()
        if self.net.pack_launch is not None:      # refresh the bf16 weight copies of the tensor-core path
            self.net.pack_launch(stream)
        self._run_ops(pl.fwd, stream)
        if pl.bn_update is not None:
            pl.bn_update(stream)
        self._run_ops(pl.bwd, stream)

    def _run_ops(self, ops, stream):
        """Issue a range of plan launches with the concurrency the plan allows (hints: graph._Ops).
        * `side` launches (weight / bias gradients) only feed the optimizer: they go to a second stream behind an event of
          the stream they were emitted on and run beside the data-gradient chain, whose bandwidth-bound BatchNorm and
          pooling kernels leave the tensor cores idle.  A launch tagged `join` overwrites a buffer a pending side launch
          reads: its stream waits for the side stream first.
        * `lane` launches belong to one ResBlock-a branch; branches alternate over RSA_LANES (default 2) streams, launches
          without a lane are barriers for all of them, `chain` launches keep their order across lanes (branch sum).
        Everything is joined into the calling stream at the end of the range, so ranges compose (graph capture, the
        data-parallel split).  RSA_WGRAD_STREAM=0 / RSA_LANES=0 switch the two mechanisms off."""
        use_side = os.environ.get("RSA_WGRAD_STREAM", "1") != "0"
        nl = int(os.environ.get("RSA_LANES", "2"))
        if self.net.device.type != "cuda" or not (use_side or nl > 0):
            for op in ops:
                op(stream)
            return
        main = torch.cuda.current_stream()
        if getattr(self, "_wg_stream", None) is None:
            self._wg_stream = torch.cuda.Stream()
            self._lane_streams = []
        while len(self._lane_streams) < nl:
            self._lane_streams.append(torch.cuda.Stream())
        side = self._wg_stream
        forked = set()       # lane streams carrying work since the last barrier
        pending = False      # side launches not yet waited for
        chain_ev = {}        # chain key -> (event, stream) of its last launch
        for op in ops:
            ln = getattr(op, "lane", None)
            in_lane = ln is not None and nl > 0
            s = self._lane_streams[ln % nl] if in_lane else main
            if getattr(op, "side", False) and use_side:
                if in_lane and (ln % nl) not in forked:
                    s = main                      # nothing of this lane is in flight: the data is final on main
                side.wait_stream(s)
                op(side.cuda_stream)
                pending = True
                continue
            if in_lane:
                if (ln % nl) not in forked:
                    s.wait_stream(main)
                    forked.add(ln % nl)
            else:
                for k in forked:
                    main.wait_stream(self._lane_streams[k])
                forked.clear()
            if pending and getattr(op, "join", False):
                s.wait_stream(side)
                pending = in_lane                 # only a wait on the main stream orders every later launch
            ck = getattr(op, "chain", None)
            if ck is not None and ck in chain_ev and chain_ev[ck][1] is not s:
                s.wait_event(chain_ev[ck][0])
            op(stream if s is main else s.cuda_stream)
            if ck is not None and nl > 0:
                ev = torch.cuda.Event()
                ev.record(s)
                chain_ev[ck] = (ev, s)
        for k in forked:
            main.wait_stream(self._lane_streams[k])
        if pending:
            main.wait_stream(side)

    def _run_train_ops(self, pl, stream):
        self._run_fwd_bwd(pl, stream)
        self._opt_launch(stream)

    def _run_eval_ops(self, pl, stream):
        pl.scratch.zero_()
        self._run_ops(pl.fwd, stream)

    def _graph(self, key, fn):
        """Capture `fn` (allocation-free pre-bound launches) once and replay it afterwards."""
        g = self._graphs.get(key)
        if g is None:
            g = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            with torch.cuda.graph(g):
                fn(torch.cuda.current_stream().cuda_stream)
            self._graphs[key] = g
        g.replay()

    def _execute(self, pl, train):
        if train:
            self.net.shadow_dirty = True      # parameters change: bf16 copies are refreshed lazily for eval
        stream = self._stream()
        on_gpu = self.net.device.type == "cuda"
        dp = train and self.dp is not None and self.dp.world_size > 1
        if not train:
            self.net.ensure_shadow(stream)
        if not (self.use_cuda_graph and on_gpu):
            if dp and self.dp.overlap:
                # eager launches: bucket all-reduces are issued as their gradients complete (distribute.py)
                pl.scratch.zero_()
                self.net.params.grad.zero_()
                if self.net.pack_launch is not None:
                    self.net.pack_launch(stream)
                self._run_ops(pl.fwd, stream)
                if pl.bn_update is not None:
                    pl.bn_update(stream)
                self.dp.run_backward(pl, stream)
                self._opt_launch(stream)
            elif dp:
                self._run_fwd_bwd(pl, stream)
                self.dp.all_reduce_sum_(self.net.params.grad)
                self._opt_launch(stream)
            else:
                (self._run_train_ops if train else self._run_eval_ops)(pl, stream)
            return
        if not train:
            self._graph((id(pl), "eval"), lambda s: self._run_eval_ops(pl, s))
        elif dp:
            # data parallel: replayed graphs for fwd+bwd with the gradient all-reduce overlapped, replayed optimizer
            def head(st):
                pl.scratch.zero_()
                self.net.params.grad.zero_()
                if self.net.pack_launch is not None:
                    self.net.pack_launch(st)
            self._dp_graph_backward(pl, 0, head, "X")
            self._graph((id(pl), "opt"), lambda s: self._opt_launch(s))
        else:
            self._graph((id(pl), "train"), lambda s: self._run_train_ops(pl, s))

    def _train_step_overlapped(self, pl, x, y):
        """Graph-replayed training step whose label upload hides behind the forward pass: the network part of the
        forward only needs x, so the (much larger) one-hot label tensors are staged and copied on a second stream
        while the GPU already runs, and the loss/backward/optimizer graph waits for that copy's event."""
        self.net.shadow_dirty = True
        main = torch.cuda.current_stream()
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream()
        nb = self._load_x(pl, x)
        self._push_lr()
        k = pl.n_fwd_net

        def part_a(st):
            pl.scratch.zero_()
            self.net.params.grad.zero_()
            if self.net.pack_launch is not None:
                self.net.pack_launch(st)
            self._run_ops(pl.fwd[:k], st)

        def part_b(st):
            self._run_ops(pl.fwd[k:], st)
            if pl.bn_update is not None:
                pl.bn_update(st)
            self._run_ops(pl.bwd, st)

        self._graph((id(pl), "A"), part_a)
        with torch.cuda.stream(self._copy_stream):
            nb += self._load_labels(pl, y)
            ev = self._copy_stream.record_event()
        main.wait_event(ev)
        if self.dp is not None and self.dp.world_size > 1:
            self._dp_graph_backward(pl, k, None, "B")
            self._graph((id(pl), "opt"), lambda st: self._opt_launch(st))
        else:
            self._graph((id(pl), "B+opt"), lambda st: (part_b(st), self._opt_launch(st)))
        return nb

    def _dp_graph_backward(self, pl, fwd_from, head, tag):
        """Forward launches [fwd_from:], BN moving statistics, backward and the NCCL gradient sum of a data-parallel step,
        graph-replayed.  The all-reduce is overlapped with backward: the parameter-heavy deep levels are final after
        backward launch ks (distribute.two_phase_split) and travel over NVLink on NCCL's stream while the shallow levels'
        backward (a second graph) still runs; the small remainder follows.  RSA_DP_GRAPH_OVERLAP=0: one all-reduce after
        the whole backward."""
        split = self.dp.two_phase_split(pl, self.net.params) if os.environ.get("RSA_DP_GRAPH_OVERLAP", "1") != "0" else None
        grad = self.net.params.grad
        ks = split[0] if split else len(pl.bwd) - 1

        def first(st):
            if head is not None:
                head(st)
            self._run_ops(pl.fwd[fwd_from:], st)
            if pl.bn_update is not None:
                pl.bn_update(st)
            self._run_ops(pl.bwd[:ks + 1], st)

        self._graph((id(pl), tag + "1"), first)
        if split is None:
            self.dp.all_reduce_sum_(grad)
            return
        off = split[1]
        h = self.dp.all_reduce_async(grad[off:])
        self._graph((id(pl), tag + "2"), lambda st: self._run_ops(pl.bwd[ks + 1:], st))
        self.dp.all_reduce_sum_(grad[:off])
        h.wait()

    def _collect(self, pl):
        """Device -> host read of the step results; returns the keras metrics list."""
        if self.net.device.type == "cuda":
            torch.cuda.current_stream().synchronize()
        rf = pl.res_f32.cpu().numpy()
        rz = pl.res_z[0].cpu()
        sums = rz[:8].numpy()
        met = rz[8:13].view(torch.int64).numpy()
        per = []
        for h, rec in pl.loss_out.items():
            per.append(float(rf[rec[1]]) if rec[0] == "mean" else float(sums[rec[1]] / rec[2]))
        total = sum(self.loss_weights[h] * v for h, v in zip(pl.loss_out, per))
        seg = pl.outputs["seg"]
        acc = float(met[0]) / float(seg.M)
        tail = [acc, float(met[1]), float(met[2]), float(met[3]), float(met[4])]
        if self.net.multitask:
            return [total] + per + tail
        return [total] + tail

    def train_on_batch(self, x, y=None, sample_weight=None, class_weight=None, reset_metrics=True,
                       return_dict=False):
        """One fwd + bwd + optimizer update on a batch (train_ISPRS.py:131,148)."""
        if self.loss_spec is None:
            raise RuntimeError("compile() the model before train_on_batch")
        self._ensure_opt()
        N = int(np.shape(x)[0])
        pl = self.net.plan(N, True, self.loss_spec)
        if self.use_cuda_graph and self.net.device.type == "cuda" and not (self.dp is not None and self.dp.overlap):
            self.last_h2d_bytes = self._train_step_overlapped(pl, x, y)
        else:
            self.last_h2d_bytes = self._load_inputs(pl, x, y)
            self._push_lr()
            self._execute(pl, True)
        res = self._collect(pl)
        self.last_d2h_bytes = 8 * 4 + 16 * 8
        if return_dict:
            return dict(zip(self.metrics_names, res))
        return res

    def test_on_batch(self, x, y=None, sample_weight=None, reset_metrics=True, return_dict=False):
        """Loss + metrics in inference mode (moving BN statistics), train_ISPRS.py:167,186."""
        if self.loss_spec is None:
            raise RuntimeError("compile() the model before test_on_batch")
        N = int(np.shape(x)[0])
        pl = self.net.plan(N, False, self.loss_spec)
        self._load_inputs(pl, x, y)
        self._execute(pl, False)
        res = self._collect(pl)
        if return_dict:
            return dict(zip(self.metrics_names, res))
        return res

    def predict(self, x, batch_size=32, verbose=0):
        """Forward in inference mode; dict of arrays keyed by head for the multitask model, one array
        otherwise (test_ISPRS.py:26-36,295)."""
        x = np.asarray(x)
        n = x.shape[0]
        outs = None
        for i in range(0, n, batch_size):
            xb = x[i:i + batch_size]
            pl = self.net.plan(xb.shape[0], False, None)
            self._load_inputs(pl, xb, None)
            self._execute(pl, False)
            if self.net.device.type == "cuda":
                torch.cuda.current_stream().synchronize()
            if outs is None:
                outs = OrderedDict((h, np.empty((n,) + t.shape[1:], dtype=np.float32)) for h, t in pl.outputs.items())
            for h, t in pl.outputs.items():
                outs[h][i:i + xb.shape[0]] = t.data.float().cpu().numpy()
        if self.net.multitask:
            return dict(outs)
        return outs["seg"]

    def __call__(self, x, training=False):
        return self.predict(x, batch_size=int(np.shape(x)[0]))

    def evaluate(self, x, y, batch_size=32, verbose=0):
        n = np.shape(x)[0]
        acc = None
        nb = 0
        for i in range(0, n - batch_size + 1, batch_size):
            yb = {k: v[i:i + batch_size] for k, v in y.items()} if isinstance(y, dict) else y[i:i + batch_size]
            r = np.array(self.test_on_batch(x[i:i + batch_size], yb))
            acc = r if acc is None else acc + r
            nb += 1
        return (acc / max(nb, 1)).tolist()

    def fit(self, x, y, batch_size=32, epochs=1, verbose=1, callbacks=None, validation_data=None, shuffle=True):
        """Epoch loop with the semantics the reference relies on (amazon_py/main_tcc.py:218):
        per-epoch mean of the batch metrics, validation via test_on_batch, EarlyStopping /
        ModelCheckpoint callbacks."""
        hist = History()
        n = np.shape(x)[0]
        rng = np.random.RandomState(0)
        nb = max(n // batch_size, 1)
        take = lambda a, idx: ({k: v[idx] for k, v in a.items()} if isinstance(a, dict) else a[idx])
        for ep in range(epochs):
            order = rng.permutation(n) if shuffle else np.arange(n)
            tot = None
            for b in range(nb):
                idx = order[b * batch_size:(b + 1) * batch_size]
                r = np.array(self.train_on_batch(x[idx], take(y, idx)))
                tot = r if tot is None else tot + r
            logs = dict(zip(self.metrics_names, (tot / nb).tolist()))
            if validation_data is not None:
                xv, yv = validation_data[0], validation_data[1]
                bs = min(batch_size, np.shape(xv)[0])
                logs.update({"val_" + k: v for k, v in zip(self.metrics_names, self.evaluate(xv, yv, bs))})
            hist.add(logs)
            if verbose:
                print(f"Epoch {ep + 1}/{epochs} - " + " - ".join(f"{k}: {v:.4f}" for k, v in logs.items()))
            stop = False
            for cb in callbacks or []:
                stop = bool(cb.on_epoch_end(self, ep, logs)) or stop
            if stop:
                break
        return hist

    # -- persistence (native .npz keyed by keras names; h5py is not available) ---------------------
    def get_weights_dict(self):
        return self.net.get_weights()

    def set_weights_dict(self, w):
        self.net.set_weights(w)

    def save(self, path):
        if self.dp is not None:
            self.dp.sync_moving_statistics(self.net.params)
        w = {k: v.numpy() for k, v in self.net.get_weights().items()}
        cfg = dict(self.config)
        if self.optimizer is not None:
            cfg["optimizer"] = self.optimizer.config()
            cfg["iterations"] = self.optimizer.iterations
        w["__config__"] = np.frombuffer(json.dumps(cfg).encode(), dtype=np.uint8)
        if self._opt_state is not None:
            for k, v in self._opt_state.items():
                w["__opt__/" + k] = v.cpu().numpy()
        with open(path, "wb") as f:   # keep the caller's file name (the reference saves 'best_model.h5')
            np.savez(f, **w)


def load_model(path, compile=True, custom_objects=None):
    """Counterpart of keras ``load_model`` for files written by :meth:`Model.save`
    (test_ISPRS.py:278, train_ISPRS.py:471-480)."""
    from .builder import build_model
    z = np.load(path, allow_pickle=False)
    cfg = json.loads(bytes(z["__config__"]).decode())
    model = build_model(tuple(cfg["input_shape"]), cfg["num_classes"], cfg["multitask"], cfg["variant"],
                        dtype=cfg.get("dtype", "bf16"))
    model.net.set_weights({k: z[k] for k in z.files if not k.startswith("__")})
    if compile and "optimizer" in cfg:
        oc = dict(cfg["optimizer"])
        kind = oc.pop("kind")
        model.optimizer = Adam(**oc) if kind == "adam" else SGD(**oc)
        model.optimizer.iterations = cfg.get("iterations", 0)
        model._saved_opt_state = {k[len("__opt__/"):]: z[k] for k in z.files if k.startswith("__opt__/")}
    return model
