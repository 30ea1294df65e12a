"""Batch-sharded data-parallel training: one process per GPU, NCCL over NVLink/NVSwitch.

Replaces ``tf.distribute.MirroredStrategy`` (train_ISPRS.py:347,432; test_ISPRS.py:276-277) — the
reference's single-process in-graph replication — with the B200 idiom: ``torchrun`` starts one rank
per GPU, every rank owns a full replica, BatchNormalization statistics stay per-replica (plain
``BatchNormalization`` under MirroredStrategy is not synchronised) and the only
exchange is one sum-all-reduce of the flat fp32 gradient buffer per step.  The buffer is cut into
buckets in *backward completion order* and each bucket's all-reduce is issued on NCCL's stream the
moment its last producer kernel has been enqueued, so the transfer overlaps the rest of backward.
The mean (1/world) is folded into the optimizer kernel's grad_scale.
"""
from __future__ import annotations

import contextlib
import os

import torch
import torch.distributed as dist

_CURRENT = None


def current_strategy():
    return _CURRENT


class DataParallel:
    def __init__(self, group=None, n_buckets=8, overlap=None):
        self.group = group
        # overlap=True: eager launches with bucketed all-reduces overlapped with backward;
        # overlap=False (default): CUDA-graph replay of fwd+bwd followed by one all-reduce of the flat buffer
        # (171 MB over NVLink is ~0.5 ms of a >20 ms step; graph replay saves more than the overlap hides)
        self.overlap = (os.environ.get("RSA_DP_OVERLAP", "0") == "1") if overlap is None else overlap
        self.world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_buckets = n_buckets
        self._sched = {}

    # ---------------------------------------------------------------------------------------------------
    def broadcast_parameters(self, params):
        """Identical initial replicas (MirroredStrategy mirrors variables at creation)."""
        if self.world_size > 1:
            dist.broadcast(params.data, src=0, group=self.group)

    def _schedule(self, pl, params):
        """bucket list [(ready_op_index, lo, hi)] sorted by readiness."""
        key = id(pl)
        if key in self._sched:
            return self._sched[key]
        n = params.n_train
        ready = torch.full((n,), -1, dtype=torch.int32)
        for name, idx in pl.grad_ready.items():
            o = params.off[name]
            sz = 1
            for d in params.spec[name][0]:
                sz *= d
            ready[o:o + sz] = idx
        # padding / never-written (exactly-zero) gradients are ready from the start
        nb = max(1, min(self.n_buckets, n // 1024))
        edges = [n * i // nb for i in range(nb + 1)]
        buckets = []
        for lo, hi in zip(edges[:-1], edges[1:]):
            buckets.append((int(ready[lo:hi].max().item()), lo, hi))
        buckets.sort()
        self._sched[key] = buckets
        return buckets

    def two_phase_split(self, pl, params, min_frac=0.6, max_pos=0.8):
        """(k, o) for the graph-replayed overlapped step: after backward launch k every gradient element at flat offset
        >= o is final, so `grad[o:]` (the deep, parameter-heavy levels: backward reaches them first and they sit at the
        END of the creation-ordered buffer) can be all-reduced on NCCL's stream while backward launches k+1.. still run;
        `grad[:o]` follows after the last launch.  None when no useful split exists."""
        key = ("2p", id(pl))
        if key in self._sched:
            return self._sched[key]
        n, nb = params.n_train, len(pl.bwd)
        ready = torch.full((n,), -1, dtype=torch.int32)
        for name, idx in pl.grad_ready.items():
            o = params.off[name]
            sz = 1
            for d in params.spec[name][0]:
                sz *= d
            ready[o:o + sz] = idx
        # suffix maximum: smax[o] = last launch that writes anything at or beyond offset o
        smax = torch.flip(torch.cummax(torch.flip(ready, [0]), 0).values, [0])
        best = None
        for k in sorted(set(int(v) for v in torch.unique(ready).tolist())):
            if k < 0 or k > max_pos * nb:
                continue
            ok = (smax <= k).nonzero()
            if ok.numel() == 0:
                continue
            o = int(ok[0].item())
            o = (o + 63) // 64 * 64                      # keep the two ranges 256-byte aligned
            if n - o >= min_frac * n:
                best = (k, o)
                break
        self._sched[key] = best
        return best

    def phase_splits(self, pl, params, fracs=(0.6, 0.97), min_gap=8):
        """[(k_1, o_1), (k_2, o_2), ...] with k increasing and o decreasing: after backward launch k_i every gradient element
        at flat offset >= o_i is final.  The graph-replayed step all-reduces grad[o_1:] while launches k_1+1.. run, then
        grad[o_2:o_1] while launches k_2+1.. run, and only grad[:o_last] after the last launch.  The first range is the one
        two_phase_split finds (the deep, parameter-heavy levels); the second closes once >= 97 % of the buffer is final - what
        is left are the shallow levels (enc1-4, stem: ~3 % of the parameters but a third of the backward time), so the
        exposed all-reduce shrinks from a quarter of the buffer to a few per cent."""
        key = ("np", id(pl), tuple(fracs))
        if key in self._sched:
            return self._sched[key]
        first = self.two_phase_split(pl, params, min_frac=fracs[0])
        out = []
        if first is not None:
            out.append(first)
            n, nb = params.n_train, len(pl.bwd)
            ready = torch.full((n,), -1, dtype=torch.int32)
            for name, idx in pl.grad_ready.items():
                o = params.off[name]
                sz = 1
                for d in params.spec[name][0]:
                    sz *= d
                ready[o:o + sz] = idx
            smax = torch.flip(torch.cummax(torch.flip(ready, [0]), 0).values, [0])
            for frac in fracs[1:]:
                k_prev, o_prev = out[-1]
                for k in sorted(set(int(v) for v in torch.unique(ready).tolist())):
                    if k < k_prev + min_gap or k > nb - 1 - min_gap:
                        continue
                    ok = (smax <= k).nonzero()
                    if ok.numel() == 0:
                        continue
                    o = (int(ok[0].item()) + 63) // 64 * 64
                    if o < o_prev and n - o >= frac * n:
                        out.append((k, o))
                        break
        self._sched[key] = out
        return out

    def run_backward(self, pl, stream, params=None):
        """Run the backward launches, issuing bucket all-reduces as their gradients complete."""
        params = params or pl.net.params
        buckets = list(self._schedule(pl, params))
        handles = []
        bi = 0
        for i, op in enumerate(pl.bwd):
            op(stream)
            while bi < len(buckets) and buckets[bi][0] <= i:
                _, lo, hi = buckets[bi]
                handles.append(dist.all_reduce(params.grad[lo:hi], op=dist.ReduceOp.SUM, group=self.group,
                                               async_op=True))
                bi += 1
        while bi < len(buckets):
            _, lo, hi = buckets[bi]
            handles.append(dist.all_reduce(params.grad[lo:hi], op=dist.ReduceOp.SUM, group=self.group,
                                           async_op=True))
            bi += 1
        for h in handles:
            h.wait()          # the compute stream waits for NCCL before the optimizer kernel

    def sync_moving_statistics(self, params):
        """BN moving mean/variance are per-replica during training and mean-aggregated when read
        (MirroredStrategy's ON_READ / MEAN aggregation)."""
        if self.world_size > 1:
            tail = params.data[params.n_train:]
            dist.all_reduce(tail, op=dist.ReduceOp.SUM, group=self.group)
            tail.div_(self.world_size)

    def mean_host(self, values):
        """Mean over ranks of a small host vector (epoch metrics): every rank gets the same numbers back."""
        import numpy as np
        v = np.asarray(values, dtype=np.float64)
        if self.world_size <= 1:
            return v
        dev = "cuda" if dist.get_backend(self.group) == "nccl" else "cpu"
        t = torch.from_numpy(v.copy()).to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return (t / self.world_size).cpu().numpy()

    def all_reduce_async(self, t):
        """Sum all-reduce on NCCL's stream (ordered after the work already enqueued on the current stream); the returned
        handle's wait() makes the current stream wait for it."""
        return dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def all_reduce_sum_(self, t):
        if self.world_size > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t


class MirroredStrategy:
    """Drop-in for ``tf.distribute.MirroredStrategy()``: under ``torchrun`` (WORLD_SIZE > 1) models
    compiled inside ``scope()`` train data-parallel; in a single process it is a no-op."""

    def __init__(self, backend=None, n_buckets=8, overlap=None):
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if world > 1 and not dist.is_initialized():
            use_cuda = torch.cuda.is_available()
            if use_cuda:
                torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group(backend=backend or ("nccl" if use_cuda else "gloo"))
        self.dp = DataParallel(n_buckets=n_buckets, overlap=overlap)
        self.num_replicas_in_sync = self.dp.world_size

    @contextlib.contextmanager
    def scope(self):
        global _CURRENT
        prev, _CURRENT = _CURRENT, self
        try:
            yield self
        finally:
            _CURRENT = prev
