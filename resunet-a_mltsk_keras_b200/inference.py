"""Sliding-window inference side of the hot path (test_ISPRS.py:26-36,48-87,102-152,295-314).

chop -> predict -> argmax -> confusion matrix / metrics -> reconstruction.  The patch extraction and
the paste-back are pure index arithmetic on host arrays (as in the reference); argmax and the K x K
confusion histogram run on the device right behind the forward pass (rsa_argmax_confusion, int64
counts, bit-exact), so the 35 M-pixel label arrays of a 6000^2 scene never round-trip as fp32
probabilities.
"""
from __future__ import annotations

import numpy as np
import torch


def extract_patches(img, patch_size):
    """Non-overlapping stride=patch_size tiles, row-major (h outer, w inner), remainder dropped
    (extract_patches_train / extract_patches_test, test_ISPRS.py:102-152).  Works for (H,W) and (H,W,C)."""
    img = np.asarray(img)
    h, w = img.shape[:2]
    nh, nw = h // patch_size, w // patch_size
    v = img[:nh * patch_size, :nw * patch_size]
    v = v.reshape((nh, patch_size, nw, patch_size) + img.shape[2:])
    v = np.swapaxes(v, 1, 2)
    return np.ascontiguousarray(v.reshape((nh * nw, patch_size, patch_size) + img.shape[2:]))


def pred_recostruction(patch_size, pred_labels, binary_img_test_ref, img_type=1):
    """Paste patch predictions back row-major into a zero-initialised float64 image of the reference
    shape; the right/bottom remainder stays 0 (test_ISPRS.py:48-87; spelling kept from the reference)."""
    h, w = np.asarray(binary_img_test_ref).shape[:2]
    nh, nw = h // patch_size, w // patch_size
    pred_labels = np.asarray(pred_labels)
    tail = pred_labels.shape[3:] if img_type == 2 else ()
    out = np.zeros((h, w) + tail)
    p = pred_labels[:nh * nw].reshape((nh, nw, patch_size, patch_size) + tail)
    p = np.swapaxes(p, 1, 2).reshape((nh * patch_size, nw * patch_size) + tail)
    out[:nh * patch_size, :nw * patch_size] = p
    return out


def confusion_matrix(y_true, y_pred, labels=None):
    """sklearn.metrics.confusion_matrix semantics on host arrays (test_ISPRS.py:314): int64, rows = true,
    labels = sorted union of the values present."""
    y_true = np.asarray(y_true).ravel()
    y_pred = np.asarray(y_pred).ravel()
    if labels is None:
        labels = np.union1d(np.unique(y_true), np.unique(y_pred))
    labels = np.asarray(labels)
    k = len(labels)
    ti = np.searchsorted(labels, y_true)
    pi = np.searchsorted(labels, y_pred)
    ok = (ti < k) & (pi < k)
    ok &= (labels[np.minimum(ti, k - 1)] == y_true) & (labels[np.minimum(pi, k - 1)] == y_pred)
    cm = np.bincount(ti[ok] * k + pi[ok], minlength=k * k).astype(np.int64)
    return cm.reshape(k, k)


def compact_confusion(cm_full):
    """Drop classes that appear neither as truth nor as prediction — sklearn builds the matrix over the
    union of labels present, so it can be smaller than num_classes^2."""
    cm_full = np.asarray(cm_full)
    present = (cm_full.sum(0) + cm_full.sum(1)) > 0
    return cm_full[np.ix_(present, present)], np.nonzero(present)[0]


def metrics_from_confusion(cm):
    """accuracy, f1, recall, precision (x100, per class) — utils.py:52-57 / test_ISPRS.py:39-45."""
    cm = np.asarray(cm, dtype=np.float64)
    diag = np.diag(cm)
    acc = 100.0 * diag.sum() / cm.sum()
    with np.errstate(divide="ignore", invalid="ignore"):
        rec = np.where(cm.sum(1) > 0, diag / cm.sum(1), 0.0)
        prec = np.where(cm.sum(0) > 0, diag / cm.sum(0), 0.0)
        f1 = np.where(rec + prec > 0, 2 * rec * prec / (rec + prec), 0.0)
    return acc, 100 * f1, 100 * rec, 100 * prec


def compute_metrics(true_labels, predicted_labels):
    """Same return tuple as utils.compute_metrics (utils.py:52-57)."""
    return metrics_from_confusion(confusion_matrix(true_labels, predicted_labels))


def _world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist, dist.get_world_size(), dist.get_rank()
    return None, 1, 0


def _patch_rows(image, patch_size):
    """(nh, nw, [nh, ps, nw, ps, ...] strided view of the covered part of the scene): patch i = view[i // nw, :, i % nw]."""
    img = np.asarray(image)
    h, w = img.shape[:2]
    nh, nw = h // patch_size, w // patch_size
    v = img[:nh * patch_size, :nw * patch_size].reshape((nh, patch_size, nw, patch_size) + img.shape[2:])
    return nh, nw, v


def _gather_patches(view, nw, lo, hi, dst):
    """dst[k] = patch lo + k of the scene, written straight into (pinned) memory: one host copy, spread over the copy pool
    (numpy releases the GIL while it copies)."""
    from . import keras_api as KA
    pool = KA._copy_pool()
    futs = [pool.submit(np.copyto, dst[k], view[(lo + k) // nw, :, (lo + k) % nw]) for k in range(hi - lo)]
    for f in futs:
        f.result()


class SceneOnDevice:
    """The device part of scene inference with this rank's patches and reference labels resident in HBM: per batch one
    forward (graph-replayed), argmax + int64 confusion counts right behind it.  bench.py times `run()` for the
    device-resident figure of BASELINE config 5; predict_scene is the end-to-end path."""

    def __init__(self, model, image, reference, patch_size, batch_size, num_classes):
        net = model.net
        self.model, self.net, self.K, self.bs = model, net, int(num_classes or net.num_classes), int(batch_size)
        dist, world, rank = _world()
        nh, nw, view = _patch_rows(image, patch_size)
        P = nh * nw
        share = -(-P // world)
        lo, hi = min(P, rank * share), min(P, (rank + 1) * share)
        self.n = hi - lo
        dev = net.device
        host = np.empty((max(self.n, 1), patch_size, patch_size) + view.shape[4:], dtype=np.float32)
        if self.n:
            _gather_patches(view, nw, lo, hi, host)
        self.x = torch.from_numpy(host[:self.n]).to(dev)
        self.ref = None
        if reference is not None:
            rp = extract_patches(np.asarray(reference), patch_size)[lo:hi].astype(np.int32)
            self.ref = torch.from_numpy(rp.reshape(self.n, -1)).to(dev)
        self.labels = torch.zeros((max(self.n, 1), patch_size * patch_size), dtype=torch.int32, device=dev)
        self.cm = torch.zeros(self.K * self.K, dtype=torch.int64, device=dev)
        self.launches_per_run = 0
        for i in range(0, self.n, self.bs):
            pl = net.plan(min(self.bs, self.n - i), False, None)
            self.launches_per_run += len(pl.fwd) + 2      # + input cast + argmax/confusion

    def run(self):
        model, net, lib = self.model, self.net, self.net.lib
        self.cm.zero_()
        for i in range(0, self.n, self.bs):
            n = min(self.bs, self.n - i)
            pl = net.plan(n, False, None)
            model._load_x(pl, self.x[i:i + n])
            model._execute(pl, False)
            prob = pl.outputs["seg"]
            tl = self.ref[i:i + n].reshape(-1) if self.ref is not None else None
            lib.argmax_confusion(prob.data, prob.M, prob.C, self.labels[i:i + n].view(-1), tl, self.K,
                                 self.cm if tl is not None else None)(model._stream())
        return self.labels, self.cm


def _paste_rows(out, seg_pred, nw, patch_size, r0, r1):
    """out[rows of tile-rows r0..r1) = tiles, converting to out's dtype on the way (one pass, no intermediate array)."""
    ps = patch_size
    for r in range(r0, r1):
        t = seg_pred[r * nw:(r + 1) * nw]                               # [nw, ps, ps]
        out[r * ps:(r + 1) * ps, :nw * ps].reshape(ps, nw, ps)[...] = np.swapaxes(t, 0, 1)


def reconstruct_threaded(patch_size, seg_pred, shape):
    """pred_recostruction for label tiles, tile-rows spread over the copy pool: a 6000 x 6000 float64 map is 288 MB of
    freshly faulted pages, which one thread fills in ~150 ms."""
    from . import keras_api as KA
    h, w = shape
    nh, nw = h // patch_size, w // patch_size
    out = np.empty((h, w), dtype=np.float64)
    out[nh * patch_size:, :] = 0
    out[:nh * patch_size, nw * patch_size:] = 0
    if nh * nw < 64:
        _paste_rows(out, seg_pred, nw, patch_size, 0, nh)
        return out
    pool = KA._copy_pool()
    step = max(1, -(-nh // 8))
    futs = [pool.submit(_paste_rows, out, seg_pred, nw, patch_size, r, min(nh, r + step)) for r in range(0, nh, step)]
    for f in futs:
        f.result()
    return out


_FINISHER = None


def _finisher():
    """Two host threads that assemble finished batches (kept apart from the copy pool, whose workers the batch gather
    saturates)."""
    global _FINISHER
    if _FINISHER is None:
        from concurrent.futures import ThreadPoolExecutor
        _FINISHER = ThreadPoolExecutor(2)
    return _FINISHER


def predict_scene(model, image, reference=None, patch_size=256, batch_size=64, num_classes=None):
    """Whole-scene inference (test_ISPRS.py:268-333): returns dict with
    ``seg_pred`` [P,ps,ps] int32, ``reconstructed`` (H,W) float64, and — when ``reference`` (H,W)
    integer labels is given — ``confusion`` (sklearn-shaped int64), ``labels`` and ``metrics``.

    Host side: every batch (patches and, if given, their reference labels) is gathered from the scene straight into one
    of two pinned buffers (one multi-threaded copy, no intermediate patch array) while the previous batch is still on the
    GPU; the label tiles come back in one copy into pinned memory at the end and are pasted into the map by the same
    thread pool.  Under ``torch.distributed`` (one process per GPU) the patches are independent: every rank
    predicts a contiguous share of them, the int64 confusion matrices are summed with one all-reduce and the label tiles
    are all-gathered, so every rank returns the complete result (as ``MirroredStrategy.predict`` does,
    test_ISPRS.py:276-277)."""
    net = model.net
    lib = net.lib
    K = int(num_classes or net.num_classes)
    nh, nw, view = _patch_rows(image, patch_size)
    rview = _patch_rows(reference, patch_size)[2] if reference is not None else None
    P = nh * nw
    dev = net.device
    on_gpu = dev.type == "cuda"
    dist, world, rank = _world()
    share = -(-P // world)                                   # patches per rank (the last ranks may get fewer / none)
    lo, hi = min(P, rank * share), min(P, (rank + 1) * share)
    cm = torch.zeros(K * K, dtype=torch.int64, device=dev)
    shape = (batch_size, patch_size, patch_size) + view.shape[4:]
    key = ("scene", shape, share)
    cache = getattr(model, "_scene_ring", None)
    if cache is None:
        cache = model._scene_ring = {}
    ent = cache.get(key)
    if ent is None:
        pin = dict(pin_memory=True) if on_gpu else {}
        ent = cache[key] = dict(
            x=[[torch.empty(shape, dtype=torch.float32, **pin), None] for _ in range(2)],
            y=[torch.empty((batch_size, patch_size * patch_size), dtype=torch.int32, **pin) for _ in range(2)],
            ydev=[torch.empty((batch_size, patch_size * patch_size), dtype=torch.int32, device=dev) for _ in range(2)],
            local=torch.zeros((share, patch_size, patch_size), dtype=torch.int32, device=dev),
            host=torch.empty((share, patch_size, patch_size), dtype=torch.int32, **pin))
    local = ent["local"]
    h, w = np.asarray(image).shape[:2]
    single = world == 1
    if single:
        # results are assembled on the host WHILE the GPU works: each batch's label tiles are copied to pinned memory
        # behind its argmax launch and a worker thread moves them into the result arrays as soon as that copy has landed
        seg_pred = np.empty((P, patch_size, patch_size), dtype=np.int32)
        recon = np.empty((h, w), dtype=np.float64)
        recon[nh * patch_size:, :] = 0
        recon[:nh * patch_size, nw * patch_size:] = 0
        finisher = _finisher()
        pending = []

        def finish(ev, i, n):
            if ev is not None:
                ev.synchronize()
            tiles = ent["host"][i:i + n].numpy()
            seg_pred[i:i + n] = tiles
            for k in range(n):
                r, c = divmod(i + k, nw)
                recon[r * patch_size:(r + 1) * patch_size, c * patch_size:(c + 1) * patch_size] = tiles[k]
    for bi, i in enumerate(range(lo, hi, batch_size)):
        n = min(batch_size, hi - i)
        buf, ybuf, ydev = ent["x"][bi & 1], ent["y"][bi & 1], ent["ydev"][bi & 1]
        if buf[1] is not None:
            buf[1].synchronize()                              # the H2D copies that last read these buffers have finished
        _gather_patches(view, nw, i, i + n, buf[0].numpy())     # np.copyto casts other dtypes to float32 on the way
        if rview is not None:
            _gather_patches(rview, nw, i, i + n, ybuf.numpy().reshape(batch_size, patch_size, patch_size))
        pl = net.plan(n, False, None)
        model._load_x(pl, buf[0][:n])                         # pinned tensor: asynchronous copy, no staging
        tl = None
        if rview is not None:
            ydev[:n].copy_(ybuf[:n], non_blocking=True)
            tl = ydev[:n].view(-1)
        if on_gpu:
            buf[1] = torch.cuda.Event()
            buf[1].record(torch.cuda.current_stream())
        model._execute(pl, False)
        prob = pl.outputs["seg"]
        lab = local[i - lo:i - lo + n].view(-1)
        lib.argmax_confusion(prob.data, prob.M, prob.C, lab, tl, K, cm if tl is not None else None)(model._stream())
        if single:
            ent["host"][i:i + n].copy_(local[i:i + n], non_blocking=True)
            ev = None
            if on_gpu:
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream())
            pending.append(finisher.submit(finish, ev, i, n))
    if single:
        for f in pending:
            f.result()
        out = dict(seg_pred=seg_pred, reconstructed=recon)
    else:
        if on_gpu:
            torch.cuda.current_stream().synchronize()
        if reference is not None:
            dist.all_reduce(cm)
        parts = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(parts, local)
        seg_pred = torch.cat(parts)[:P].cpu().numpy()
        out = dict(seg_pred=seg_pred, reconstructed=reconstruct_threaded(patch_size, seg_pred, (h, w)))
    if reference is not None:
        full = cm.cpu().numpy().reshape(K, K)
        out["confusion_full"] = full
        out["confusion"], out["labels"] = compact_confusion(full)
        out["metrics"] = metrics_from_confusion(out["confusion"])
    return out
