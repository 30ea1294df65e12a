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


def _copy_pool():
    global _COPY_POOL
    if _COPY_POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _COPY_POOL = ThreadPoolExecutor(max(1, int(os.environ.get("RSA_COPY_THREADS", "8"))))
    return _COPY_POOL


def _host_copy(dst_np, src_np, min_chunk=4 << 20):
    """Parallel memcpy of a host batch into its pinned staging buffer.  Own thread pool (numpy releases the GIL while it
    copies): torch's intra-op threads are not usable for this under torchrun, which sets OMP_NUM_THREADS=1 per rank and
    turned the 100 MB/step staging copy into a 20+ ms single-thread memcpy at N > 1."""
    n = dst_np.size
    nthreads = int(os.environ.get("RSA_COPY_THREADS", "8"))
    if nthreads <= 1 or n * dst_np.itemsize < 2 * min_chunk:
        np.copyto(dst_np, src_np)
        return
    pool = _copy_pool()
    d, s_ = dst_np.reshape(-1), src_np.reshape(-1)
    step = max(min_chunk // dst_np.itemsize, (n + nthreads - 1) // nthreads)
    futs = [pool.submit(np.copyto, d[i:i + step], s_[i:i + step]) for i in range(0, n, step)]
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
    """Evaluate a loss on the device outside a model (what calling the keras loss fn directly does).
    tanimoto -> [B] (multitasking_utils.py:71-85); cce -> [B,H,W] (utils.py:466-491); bce / mse -> [B,H,W] (mean over the
    last axis, keras BinaryCrossentropy / MeanSquaredError).  Same kernels as inside a model: rsa_tanimoto_* and
    rsa_pixel_loss_elem."""
    from . import _capi
    lib = _capi.get_lib()
    dev = torch.device("cpu") if getattr(lib, "is_emulation", False) else torch.device("cuda")
    p = torch.as_tensor(np.asarray(y_pred), dtype=torch.float32).to(dev).contiguous()
    y = torch.as_tensor(np.asarray(y_true), dtype=torch.float32).to(dev).contiguous()
    if p.shape != y.shape or p.dim() != 4:
        raise ValueError(f"loss: y_true {tuple(y.shape)} and y_pred {tuple(p.shape)} must be equal [B,H,W,C] tensors")
    B, H, W, C = p.shape
    stream = 0 if dev.type == "cpu" else torch.cuda.current_stream().cuda_stream
    if kind == "tanimoto":
        sums = torch.zeros(B * C * 5, dtype=torch.float64, device=dev)
        lb = torch.empty(B, dtype=torch.float32, device=dev)
        lib.tanimoto_sums(p, y, B, H * W, C, sums)(stream)
        lib.tanimoto_finalize(sums, B, H * W, C, 1.0, lb, None, None)(stream)
        return lb.cpu().numpy()
    code = {"cce": 0, "bce": 1, "mse": 2}[kind]
    cw = None
    if class_weights is not None:
        cw = torch.tensor(list(class_weights), dtype=torch.float32).to(dev)
        if cw.numel() != C:
            raise ValueError(f"loss: {cw.numel()} class weights for {C} classes")
    out = torch.empty(B * H * W, dtype=torch.float32, device=dev)
    lib.pixel_loss_elem(code, p, y, cw, B * H * W, C, out)(stream)
    return out.view(B, H, W).cpu().numpy()


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
        return _run_loss_standalone("cce", y_true, y_pred, w)
    loss.rsa_kind = "cce"
    loss.rsa_class_weights = tuple(w)
    return loss


class _KerasLoss:
    rsa_class_weights = None

    def __call__(self, y_true, y_pred):
        """keras loss objects reduce with SUM_OVER_BATCH_SIZE: the mean over every element the function returns."""
        return float(_run_loss_standalone(self.rsa_kind, y_true, y_pred).mean())


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
# metrics accepted by compile(metrics=...) (train_ISPRS.py:446-452): all are computed by rsa_seg_metrics
# ------------------------------------------------------------------------------------------------------
class _Metric:
    def __init__(self, name=None):
        self.name = name or self.rsa_metric


class TruePositives(_Metric):
    rsa_metric = "true_positives"


class FalsePositives(_Metric):
    rsa_metric = "false_positives"


class TrueNegatives(_Metric):
    rsa_metric = "true_negatives"


class FalseNegatives(_Metric):
    rsa_metric = "false_negatives"


_METRIC_SLOTS = {"accuracy": 0, "acc": 0, "categorical_accuracy": 0, "true_positives": 1, "false_positives": 2,
                 "true_negatives": 3, "false_negatives": 4}


def _resolve_metric(m):
    """-> (reported name, slot in the rsa_seg_metrics result)."""
    if isinstance(m, str):
        if m not in _METRIC_SLOTS:
            raise ValueError(f"unsupported metric '{m}' (have: accuracy, TruePositives, FalsePositives, TrueNegatives, "
                             "FalseNegatives)")
        return ("accuracy" if _METRIC_SLOTS[m] == 0 else m), _METRIC_SLOTS[m]
    kind = getattr(m, "rsa_metric", None)
    if kind is None:
        raise ValueError(f"unsupported metric object {m!r}")
    return m.name, _METRIC_SLOTS[kind]


# ------------------------------------------------------------------------------------------------------
# callbacks used by the reference's fit() call (amazon_py/main_tcc.py:212-218)
# ------------------------------------------------------------------------------------------------------
def _monitor_sign(monitor, mode):
    """+1 when smaller is better, -1 when larger is better (keras 'auto': accuracy-like monitors maximise)."""
    if mode not in ("min", "max", "auto"):
        raise ValueError(f"mode must be 'min', 'max' or 'auto', got {mode!r}")
    if mode == "auto":
        mode = "max" if ("acc" in monitor or monitor.startswith("fmeasure")) else "min"
    return 1.0 if mode == "min" else -1.0


class EarlyStopping:
    def __init__(self, monitor="val_loss", min_delta=0.0, patience=0, verbose=0, mode="auto"):
        self.monitor, self.min_delta, self.patience = monitor, abs(min_delta), patience
        self.sign = _monitor_sign(monitor, mode)
        self.best, self.wait = math.inf, 0

    def on_epoch_end(self, model, epoch, logs):
        cur = logs.get(self.monitor)
        if cur is None:
            return False
        cur = self.sign * cur
        if cur < self.best - self.min_delta:
            self.best, self.wait = cur, 0
            return False
        self.wait += 1
        return self.wait >= self.patience


class ModelCheckpoint:
    def __init__(self, filepath, monitor="val_loss", verbose=0, save_best_only=False, mode="auto"):
        self.filepath, self.monitor, self.save_best_only = filepath, monitor, save_best_only
        self.sign = _monitor_sign(monitor, mode)
        self.best = math.inf

    def on_epoch_end(self, model, epoch, logs):
        cur = logs.get(self.monitor)
        cur = None if cur is None else self.sign * cur
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
        self._metric_sel = []
        self._metrics_cfg = None
        self._opt_state = None
        self._opt_launch = None
        self._saved_opt_state = None
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
        """keras Model.compile (train_ISPRS.py:446-461).  optimizer=None keeps the optimizer (and its restored slots) of a
        model that came from load_model(); a different optimizer object drops the previous optimizer's state.  metrics:
        list (single output) or dict keyed by output name of 'accuracy' / TruePositives() / FalsePositives() /
        TrueNegatives() / FalseNegatives() on the segmentation output, reported in the order given; None reports all
        five (the list the reference trains with)."""
        if isinstance(optimizer, str):
            optimizer = {"adam": Adam, "sgd": SGD}[optimizer.lower()]()
        if optimizer is not None and optimizer is not self.optimizer:
            self.optimizer = optimizer
            self._opt_state = None
            self._opt_launch = None
            self._saved_opt_state = None
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
        self._set_loss_spec(spec, metrics)
        self._graphs.clear()
        from . import distribute
        strat = distribute.current_strategy()
        if strat is not None and strat.dp.world_size > 1:
            self.dp = strat.dp
            self.dp.broadcast_parameters(self.net.params)
            self._opt_state = None

    def _set_loss_spec(self, spec, metrics=None):
        heads = self.output_names
        self.loss_spec = tuple((h, k, float(w), None if cw is None else tuple(cw)) for h, k, w, cw in spec)
        self.loss_weights = {h: w for h, _, w, _ in self.loss_spec}
        if metrics is None:
            sel = [(n, i) for i, n in enumerate(("accuracy", "true_positives", "false_positives", "true_negatives",
                                                 "false_negatives"))]
        else:
            if isinstance(metrics, dict):
                extra = [h for h in metrics if h != "seg"]
                if extra:
                    raise ValueError(f"metrics are computed on the 'seg' output only, got {extra}")
                metrics = metrics.get("seg", [])
            sel = [_resolve_metric(m) for m in (metrics if isinstance(metrics, (list, tuple)) else [metrics])]
        self._metric_sel = sel
        self._metrics_cfg = [n for n, _ in sel] if metrics is not None else None
        pre = "seg_" if self.net.multitask else ""
        self.metrics_names = ["loss"] + ([f"{h}_loss" for h in heads] if self.net.multitask else []) + [
            pre + n for n, _ in sel]

    # -- data movement -------------------------------------------------------------------------------------
    def _stream(self):
        return 0 if self.net.device.type == "cpu" else torch.cuda.current_stream().cuda_stream

    def _stage(self, key, arr, dst):
        """host numpy (any float) -> pinned fp32 staging -> device tensor `dst` (fp32) asynchronously.  A float32 torch
        tensor that already lives in pinned memory (data.PatchBatchLoader) is copied from directly."""
        if isinstance(arr, torch.Tensor):
            if arr.device == dst.device and tuple(arr.shape) == tuple(dst.shape) and arr.device.type != "cpu":
                dst.copy_(arr)                    # already resident on this device (inference.SceneOnDevice)
                return 0
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
        ent = self._staging.get((key, a.shape))
        if ent is None:
            ent = self._staging[(key, a.shape)] = [torch.empty(a.shape, dtype=torch.float32, pin_memory=True), None]
        st, ev = ent
        if ev is not None:
            ev.synchronize()                   # the previous asynchronous copy out of this buffer may still be queued
        _host_copy(st.numpy(), a)              # multi-threaded host copy into the pinned staging buffer
        dst.copy_(st, non_blocking=True)
        ent[1] = torch.cuda.Event()
        ent[1].record(torch.cuda.current_stream())
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
            opt, n = self.optimizer, ps.n_train
            self._opt_launch = self._opt_range(0, n)
            self._early_opt = {}
            # algorithmic HBM bytes per parameter: Adam reads p, g, m, v and writes p, m, v; SGD reads p, g, vel, writes p, vel
            self._opt_launch.hbm_bytes = float(n) * (28 if isinstance(opt, Adam) else 20)

    def _opt_range(self, lo, hi):
        """Optimizer launch over the flat parameter range [lo, hi) (the update is element-wise)."""
        lib, opt, ps, st = self.net.lib, self.optimizer, self.net.params, self._opt_state
        gs = 1.0 / (self.dp.world_size if self.dp else 1)
        if isinstance(opt, Adam):
            return lib.adam_step(ps.data[lo:hi], ps.grad[lo:hi], st["m"][lo:hi], st["v"][lo:hi], hi - lo, self._lr_dev,
                                 opt.beta_1, opt.beta_2, opt.epsilon, gs)
        return lib.sgd_step(ps.data[lo:hi], ps.grad[lo:hi], st["vel"][lo:hi], hi - lo, self._lr_dev, opt.momentum, gs)

    def _early_opt_plan(self, pl):
        """(k, [opt_hi, pack_hi], [opt_lo, pack_lo]) or None.  The flat gradient buffer becomes final from its END (the deep,
        parameter-heavy levels: backward reaches them first), so after backward launch k the optimizer can already update
        every parameter at offset >= o and the bf16 copies of those layers can be refreshed - ~98 % of the 1.5 GB the
        optimizer and the refresh move - on a side stream while the shallow levels (a third of the backward time,
        tensor-core bound) still run.  Same split as the data-parallel all-reduce ranges (distribute.phase_splits)."""
        if os.environ.get("RSA_EARLY_OPT", "1") == "0" or self.net.pack_launch is None:
            return None
        if id(pl) not in self._early_opt:
            from .distribute import DataParallel
            helper = self.dp if self.dp is not None else DataParallel()
            splits = helper.phase_splits(pl, self.net.params)
            plan = None
            if splits:
                k, off = splits[-1]
                n = self.net.params.n_train
                pack_hi, pack_lo = self.net.pack_launches_split(off)
                hi = [self._opt_range(off, n)] + ([pack_hi] if pack_hi is not None else [])
                lo = [self._opt_range(0, off)] + ([pack_lo] if pack_lo is not None else [])
                plan = (k, hi, lo)
            self._early_opt[id(pl)] = plan
        return self._early_opt[id(pl)]

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
        main = torch.cuda.current_stream()
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream()
            self._opt_stream = torch.cuda.Stream()
        dp = self.dp is not None and self.dp.world_size > 1
        early = None if dp else self._early_opt_plan(pl)
        if early is not None:
            self.net.ensure_shadow(main.cuda_stream)      # first step / new weights; afterwards every step leaves them fresh
        self.net.shadow_dirty = True
        nb = self._load_x(pl, x)
        self._push_lr()
        k = pl.n_fwd_net

        def part_a(st):
            pl.scratch.zero_()
            self.net.params.grad.zero_()
            if self.net.pack_launch is not None and early is None:
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
        if dp:
            self._dp_graph_backward(pl, k, None, "B")
            self._graph((id(pl), "opt"), lambda st: self._opt_launch(st))
        elif early is None:
            self._graph((id(pl), "B+opt"), lambda st: (part_b(st), self._opt_launch(st)))
        else:
            ks, hi_ops, lo_ops = early

            def part_b_early(st):
                cur = torch.cuda.current_stream()
                self._run_ops(pl.fwd[k:], st)
                if pl.bn_update is not None:
                    pl.bn_update(st)
                self._run_ops(pl.bwd[:ks + 1], st)               # ends with every side launch joined: grad[off:] is final
                self._opt_stream.wait_stream(cur)
                for op in hi_ops:
                    op(self._opt_stream.cuda_stream)
                self._run_ops(pl.bwd[ks + 1:], st)
                for op in lo_ops:
                    op(st)
                cur.wait_stream(self._opt_stream)

            self._graph((id(pl), "B+early-opt"), part_b_early)
            self.net.shadow_dirty = False                         # the step ended with the bf16 copies refreshed
        return nb

    def _dp_graph_backward(self, pl, fwd_from, head, tag):
        """Forward launches [fwd_from:], BN moving statistics, backward and the NCCL gradient sum of a data-parallel step,
        graph-replayed.  The all-reduce is overlapped with backward: the flat gradient buffer becomes final from its END
        (deep levels: backward reaches them first) towards its start, so after the launches distribute.phase_splits names a
        suffix goes out on NCCL's stream while the next backward graph runs; only the last few per cent (shallow levels)
        are reduced after the final launch.  RSA_DP_GRAPH_OVERLAP=0: one all-reduce after the whole backward;
        RSA_DP_RANGES=1: the two-range schedule of round 1."""
        splits = []
        if os.environ.get("RSA_DP_GRAPH_OVERLAP", "1") != "0":
            splits = self.dp.phase_splits(pl, self.net.params)
            if os.environ.get("RSA_DP_RANGES", "2") == "1":
                splits = splits[:1]
        grad = self.net.params.grad

        def first(st):
            if head is not None:
                head(st)
            self._run_ops(pl.fwd[fwd_from:], st)
            if pl.bn_update is not None:
                pl.bn_update(st)
            self._run_ops(pl.bwd[:(splits[0][0] + 1) if splits else len(pl.bwd)], st)

        self._graph((id(pl), tag + "1"), first)
        if not splits:
            self.dp.all_reduce_sum_(grad)
            return
        handles, hi = [], grad.numel()
        for i, (k, off) in enumerate(splits):
            handles.append(self.dp.all_reduce_async(grad[off:hi]))
            hi = off
            k_next = splits[i + 1][0] if i + 1 < len(splits) else len(pl.bwd) - 1
            self._graph((id(pl), tag + str(i + 2)), lambda st, a=k + 1, b=k_next + 1: self._run_ops(pl.bwd[a:b], st))
        self.dp.all_reduce_sum_(grad[:hi])
        for h in handles:
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
        vals = [float(met[0]) / float(seg.M), float(met[1]), float(met[2]), float(met[3]), float(met[4])]
        tail = [vals[i] for _, i in self._metric_sel]
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

    def _count_mask(self):
        """True for the entries of a metrics vector that keras accumulates as totals (TP / FP / TN / FN counters); losses
        and accuracy are sample-weighted means."""
        nloss = len(self.metrics_names) - len(self._metric_sel)
        return np.array([False] * nloss + [slot != 0 for _, slot in self._metric_sel])

    def _reduce_batches(self, rows, sizes):
        rows, sizes = np.asarray(rows, dtype=np.float64), np.asarray(sizes, dtype=np.float64)
        mean = (rows * sizes[:, None]).sum(0) / sizes.sum()
        return np.where(self._count_mask(), rows.sum(0), mean)

    def evaluate(self, x, y, batch_size=32, verbose=0):
        """keras evaluate: losses and accuracy are sample-weighted means over all batches (the short last batch
        included), the confusion counters are totals."""
        n = int(np.shape(x)[0])
        batch_size = max(1, min(int(batch_size), n))
        rows, sizes = [], []
        for i in range(0, n, batch_size):
            yb = {k: v[i:i + batch_size] for k, v in y.items()} if isinstance(y, dict) else y[i:i + batch_size]
            rows.append(self.test_on_batch(x[i:i + batch_size], yb))
            sizes.append(min(batch_size, n - i))
        return self._reduce_batches(rows, sizes).tolist()

    def fit(self, x, y, batch_size=32, epochs=1, verbose=1, callbacks=None, validation_data=None, shuffle=True):
        """Epoch loop with the semantics the reference relies on (amazon_py/main_tcc.py:218):
        per-epoch mean of the batch metrics, validation via test_on_batch, EarlyStopping /
        ModelCheckpoint callbacks."""
        hist = History()
        n = np.shape(x)[0]
        rng = np.random.RandomState(0)
        batch_size = max(1, min(int(batch_size), n))
        take = lambda a, idx: ({k: v[idx] for k, v in a.items()} if isinstance(a, dict) else a[idx])
        for ep in range(epochs):
            order = rng.permutation(n) if shuffle else np.arange(n)
            rows, sizes = [], []
            for b0 in range(0, n, batch_size):     # keras trains on the short last batch too
                idx = order[b0:b0 + batch_size]
                rows.append(self.train_on_batch(x[idx], take(y, idx)))
                sizes.append(len(idx))
            logs = dict(zip(self.metrics_names, self._reduce_batches(rows, sizes).tolist()))
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
        if self.loss_spec is not None:
            cfg["loss_spec"] = [[h, k, lw_, None if cw is None else list(cw)] for h, k, lw_, cw in self.loss_spec]
            cfg["metrics"] = self._metrics_cfg
        if self.optimizer is not None:
            cfg["optimizer"] = self.optimizer.config()
            cfg["iterations"] = self.optimizer.iterations
        w["__config__"] = np.frombuffer(json.dumps(cfg).encode(), dtype=np.uint8)
        if self._opt_state is not None:
            for k, v in self._opt_state.items():
                w["__opt__/" + k] = v.cpu().numpy()
        if self.dp is not None and self.dp.world_size > 1 and self.dp.rank != 0:
            return                    # replicas are identical after the statistics sync: rank 0 writes the file
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
        if cfg.get("loss_spec"):
            # a compiled checkpoint resumes training without a new compile() (train_ISPRS.py:471-480)
            model._set_loss_spec([(h, k, w, cw) for h, k, w, cw in cfg["loss_spec"]], cfg.get("metrics"))
            from . import distribute
            strat = distribute.current_strategy()
            if strat is not None and strat.dp.world_size > 1:
                model.dp = strat.dp
                model.dp.broadcast_parameters(model.net.params)
    return model
