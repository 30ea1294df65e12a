"""Patch dataset reader and training-loop shell of the ISPRS workflow.

On-disk format (written by the reference's preprocess_save_patches_ISPRS.py:178-228):

    <root>/train/patch_<i>.npy                 float32 [H, W, C]   image patch (already /255)
    <root>/labels/seg/patch_<i>.npy            float32 [H, W, n]   one-hot classes
    <root>/labels/{bound,dist,color}/...       float32 [H, W, n] / [H, W, 3]  multitask targets

The reference reads 5*B files per step synchronously with np.load inside the training loop
(train_ISPRS.py:115-141, 159-186), so on real data the step is I/O-bound.  Here a thread pool reads the .npy payloads
straight into a ring of pinned host batches a few steps ahead; `Model.train_on_batch` recognises pinned tensors and
skips its own staging copy, so the only host work left on the step path is the asynchronous H2D copy.
"""
from __future__ import annotations

import os
import queue
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

HEADS = ("seg", "bound", "dist", "color")


# ---------------------------------------------------------------------------------------------------------------
# dataset layout
# ---------------------------------------------------------------------------------------------------------------
def save_patch_dataset(root, x, y):
    """Write patches in the reference layout (preprocess_save_patches_ISPRS.py:178-228).  x: [N,H,W,C]; y: dict head ->
    [N,H,W,*] or a single array (segmentation only)."""
    if not isinstance(y, dict):
        y = {"seg": y}
    os.makedirs(os.path.join(root, "train"), exist_ok=True)
    for h in y:
        os.makedirs(os.path.join(root, "labels", h), exist_ok=True)
    for i in range(len(x)):
        np.save(os.path.join(root, "train", f"patch_{i}.npy"), np.asarray(x[i], dtype=np.float32))
        for h, v in y.items():
            np.save(os.path.join(root, "labels", h, f"patch_{i}.npy"), np.asarray(v[i], dtype=np.float32))


def list_patch_dataset(root, multitasking=True):
    """(x_paths, {head: paths}) paired BY FILE NAME (the reference pairs independent os.listdir() orders,
    train_ISPRS.py:354-379, which only works when every directory lists identically)."""
    names = sorted(os.listdir(os.path.join(root, "train")), key=_natural)
    heads = HEADS if multitasking else HEADS[:1]
    x_paths = [os.path.join(root, "train", n) for n in names]
    y_paths = {}
    for h in heads:
        d = os.path.join(root, "labels", h)
        missing = [n for n in names if not os.path.exists(os.path.join(d, n))]
        if missing:
            raise FileNotFoundError(f"{len(missing)} patches have no '{h}' label under {d} (first: {missing[0]})")
        y_paths[h] = [os.path.join(d, n) for n in names]
    return x_paths, y_paths


def _natural(name):
    stem = os.path.splitext(name)[0]
    tail = stem.rsplit("_", 1)[-1]
    return (0, int(tail)) if tail.isdigit() else (1, stem)


def train_val_split(x_paths, y_paths, test_size=0.2, random_state=42):
    """The reference's sklearn train_test_split(test_size=0.2, random_state=42) (train_ISPRS.py:381-384) applied to one
    index vector, so that every head stays paired with its image."""
    from sklearn.model_selection import train_test_split
    idx = np.arange(len(x_paths))
    tr, va = train_test_split(idx, test_size=test_size, random_state=random_state)
    pick = lambda lst, ids: [lst[i] for i in ids]
    return (pick(x_paths, tr), {h: pick(p, tr) for h, p in y_paths.items()},
            pick(x_paths, va), {h: pick(p, va) for h, p in y_paths.items()})


def read_npy_into(path, dst):
    """Read a .npy payload directly into `dst` (a C-contiguous float32 numpy view, e.g. of pinned memory) without the
    intermediate array np.load would allocate; other dtypes / layouts go through np.load + cast like the reference's
    `.astype(np.float32)` (train_ISPRS.py:123-139)."""
    with open(path, "rb") as f:
        major, minor = np.lib.format.read_magic(f)
        shape, fortran, dtype = (np.lib.format.read_array_header_1_0(f) if (major, minor) == (1, 0)
                                 else np.lib.format.read_array_header_2_0(f))
        if tuple(shape) != tuple(dst.shape):
            raise ValueError(f"{path}: shape {tuple(shape)} does not match the batch slot {tuple(dst.shape)}")
        if dtype == np.float32 and not fortran and dst.flags.c_contiguous:
            got = f.readinto(memoryview(dst).cast("B"))
            if got != dst.nbytes:
                raise IOError(f"{path}: truncated payload ({got} of {dst.nbytes} bytes)")
            return
    dst[...] = np.load(path).astype(np.float32)


# ---------------------------------------------------------------------------------------------------------------
# prefetching batch loader
# ---------------------------------------------------------------------------------------------------------------
class PatchBatchLoader:
    """Iterates (x, y) batches over lists of patch files, `prefetch` batches ahead of the consumer.

    * same batching as the reference loop: n // batch_size full batches, remainder dropped (train_ISPRS.py:102,116);
    * shuffle=True draws a new permutation per epoch from `seed` (the reference calls sklearn.utils.shuffle with the
      global RNG, train_ISPRS.py:105-112);
    * batches are torch tensors in pinned host memory when CUDA is available (plain memory otherwise), float32, NHWC;
      a slot is only refilled after the consumer has moved two batches on, by which time the asynchronous H2D copy
      issued by `train_on_batch` for it has completed (the step that follows collects its results on the same stream);
    * world/rank shard the batch list for data-parallel training (one process per GPU): rank r takes batches r, r+world, ...
      of the first (n // batch_size) // world * world batches, so every rank runs the same number of steps (each step is
      one gradient all-reduce; a rank with an extra batch would wait for peers that have already left the epoch).
    """

    def __init__(self, x_paths, y_paths, batch_size, shuffle=False, seed=0, workers=8, prefetch=3, rank=0, world=1,
                 pin=None):
        if not isinstance(y_paths, dict):
            y_paths = {"seg": list(y_paths)}
        self.x_paths, self.y_paths = list(x_paths), {h: list(p) for h, p in y_paths.items()}
        for h, p in self.y_paths.items():
            if len(p) != len(self.x_paths):
                raise ValueError(f"{len(p)} '{h}' labels for {len(self.x_paths)} patches")
        self.batch_size, self.shuffle, self.seed = int(batch_size), bool(shuffle), int(seed)
        self.workers, self.prefetch = max(1, int(workers)), max(3, int(prefetch))
        self.rank, self.world = int(rank), int(world)
        self.pin = torch.cuda.is_available() if pin is None else bool(pin)
        self.epoch = 0
        if len(self.x_paths) < self.batch_size * self.world:
            raise ValueError("fewer patches than one batch per rank")
        self._shapes = {"x": self._probe(self.x_paths[0])}
        self._shapes.update({h: self._probe(p[0]) for h, p in self.y_paths.items()})
        self._slots = None

    @staticmethod
    def _probe(path):
        a = np.load(path, mmap_mode="r")
        return tuple(a.shape)

    def _n_batches(self):
        """Batches of the whole job per epoch: a multiple of the world size."""
        return len(self.x_paths) // self.batch_size // self.world * self.world

    def __len__(self):
        return self._n_batches() // self.world

    def _alloc(self):
        if self._slots is None:
            mk = lambda shp: torch.empty((self.batch_size,) + shp, dtype=torch.float32, pin_memory=self.pin)
            self._slots = [{k: mk(s) for k, s in self._shapes.items()} for _ in range(self.prefetch)]
        return self._slots

    def order(self, epoch):
        n = len(self.x_paths)
        return np.random.RandomState(self.seed + epoch).permutation(n) if self.shuffle else np.arange(n)

    def __iter__(self):
        slots = self._alloc()
        order = self.order(self.epoch)
        self.epoch += 1
        mine = list(range(self.rank, self._n_batches(), self.world))
        free, ready = queue.Queue(), queue.Queue()
        for s in range(len(slots)):
            free.put(s)
        stop = threading.Event()

        def fill(s, ids):
            slot = slots[s]
            jobs = []
            for b, i in enumerate(ids):
                jobs.append((self.x_paths[i], slot["x"][b].numpy()))
                for h, p in self.y_paths.items():
                    jobs.append((p[i], slot[h][b].numpy()))
            return list(pool.map(lambda j: read_npy_into(*j), jobs))

        def producer():
            try:
                for bi in mine:
                    while True:
                        if stop.is_set():
                            return
                        try:
                            s = free.get(timeout=0.05)
                            break
                        except queue.Empty:
                            continue
                    fill(s, order[bi * self.batch_size:(bi + 1) * self.batch_size])
                    ready.put((s, None))
                ready.put((None, None))
            except BaseException as e:      # surface I/O errors in the consumer thread
                ready.put((None, e))

        pool = ThreadPoolExecutor(self.workers)
        th = threading.Thread(target=producer, daemon=True)
        th.start()
        held = []                           # slots handed to the consumer, oldest first
        try:
            while True:
                s, err = ready.get()
                if err is not None:
                    raise err
                if s is None:
                    break
                held.append(s)
                if len(held) > 2:           # the batch two steps back is no longer read by any copy in flight
                    free.put(held.pop(0))
                slot = slots[s]
                y = {h: slot[h] for h in self.y_paths}
                yield slot["x"], (y if len(y) > 1 or "seg" not in y else y["seg"])
        finally:
            stop.set()
            th.join(timeout=5)
            pool.shutdown(wait=True)


# ---------------------------------------------------------------------------------------------------------------
# training-loop shell (train_ISPRS.py:55-292)
# ---------------------------------------------------------------------------------------------------------------
def compute_mcc(tp, tn, fp, fn):
    """Matthews correlation coefficient from the seg head's confusion counts (utils.py compute_mcc)."""
    den = np.sqrt(float(tp + fp) * float(tp + fn) * float(tn + fp) * float(tn + fn))
    return float((tp * tn - fp * fn) / den) if den > 0 else 0.0


def _scalar_writers(log_dir, log):
    """(train, val) TensorBoard writers like train_ISPRS.py:57-63 (`<log_dir>/train`, `<log_dir>/val`), or (None, None)."""
    if not log_dir:
        return None, None
    try:
        from torch.utils.tensorboard import SummaryWriter
    except Exception as e:          # tensorboard is an optional dependency
        log(f"TensorBoard scalars disabled: {e}")
        return None, None
    return SummaryWriter(os.path.join(log_dir, "train")), SummaryWriter(os.path.join(log_dir, "val"))


def add_tensorboard_scalars(train_writer, val_writer, epoch, metric_name, train_loss, val_loss, train_acc=None, val_acc=None,
                            val_mcc=None):
    """Same tags as the reference's add_tensorboard_scalars (train_ISPRS.py:35-53): <Task>/Loss, /Accuracy, /MCC."""
    if train_writer is None:
        return
    train_writer.add_scalar(metric_name + "/Loss", train_loss, epoch)
    if train_acc is not None:
        train_writer.add_scalar(metric_name + "/Accuracy", train_acc, epoch)
    val_writer.add_scalar(metric_name + "/Loss", val_loss, epoch)
    if val_acc is not None:
        val_writer.add_scalar(metric_name + "/Accuracy", val_acc, epoch)
    if val_mcc is not None:
        val_writer.add_scalar(metric_name + "/MCC", val_mcc, epoch)


def train_model(net, train_loader, val_loader, epochs, results_path, patience=10, delta=0.001, metrics_names=None,
                log=print, save_name="best_model.npz", tensorboard_dir=None):
    """Epoch loop of the reference trainer: mean of the per-batch train_on_batch / test_on_batch vectors
    (train_ISPRS.py:97-189), per-task table, MCC of the segmentation head, TensorBoard scalars per task when
    `tensorboard_dir` is given (train_ISPRS.py:35-63,226-268; rank 0 only), early stopping on the validation loss with
    `delta` / `patience` and a checkpoint of the best model (train_ISPRS.py:276-292).  Returns (net, history)."""
    names = list(metrics_names or net.metrics_names)
    dp = getattr(net, "dp", None)
    multi = dp is not None and dp.world_size > 1
    if not multi or dp.rank == 0:
        os.makedirs(results_path, exist_ok=True)
    min_loss, cont, history = float("inf"), 0, []
    tw, vw = _scalar_writers(tensorboard_dir if (not multi or dp.rank == 0) else None, log)
    tb_names = dict(Seg="Segmentation", Bound="Boundary", Dist="Distance", Color="Color", Total="Total")
    for epoch in range(epochs):
        tr = np.zeros(len(names))
        nb = 0
        for x, y in train_loader:
            tr += np.asarray(net.train_on_batch(x, y, return_dict=False), dtype=np.float64)
            nb += 1
        tr /= max(nb, 1)
        va = np.zeros(len(names))
        nv = 0
        for x, y in val_loader:
            va += np.asarray(net.test_on_batch(x, y), dtype=np.float64)
            nv += 1
        va /= max(nv, 1)
        if multi:
            # every rank must take the same early-stop / checkpoint decision (Model.save is a collective): average the
            # per-rank means; BN moving statistics are per-replica, so the validation losses differ across ranks
            tr, va = dp.mean_host(tr), dp.mean_host(va)
        trm, vam = dict(zip(names, tr.tolist())), dict(zip(names, va.tolist()))
        pre = "seg_" if "seg_true_positives" in vam else ""
        mcc = None
        if pre + "true_positives" in vam:
            mcc = compute_mcc(vam[pre + "true_positives"], vam[pre + "true_negatives"], vam[pre + "false_positives"],
                              vam[pre + "false_negatives"])
        rows = [(t, trm.get(f"{t.lower()}_loss"), vam.get(f"{t.lower()}_loss"), trm.get(f"{t.lower()}_accuracy"),
                 vam.get(f"{t.lower()}_accuracy")) for t in ("Seg", "Bound", "Dist", "Color") if f"{t.lower()}_loss" in trm]
        rows.append(("Total", trm["loss"], vam["loss"], trm.get("accuracy"), vam.get("accuracy")))
        log(f"Epoch: {epoch}")
        log(f"{'Task':8s} {'Loss':>10s} {'Val Loss':>10s} {'Acc %':>9s} {'Val Acc %':>10s}")
        for t, l, vl, a, vacc in rows:
            log(f"{t:8s} {l:10.5f} {vl:10.5f} {100 * (a or 0):9.5f} {100 * (vacc or 0):10.5f}")
            add_tensorboard_scalars(tw, vw, epoch, tb_names[t], l, vl, a, vacc, val_mcc=mcc if t == "Seg" else None)
        if mcc is not None:
            log(f"Validation MCC: {mcc:.5f}")
        val_loss = vam["loss"]
        history.append(dict(epoch=epoch, train=trm, val=vam, mcc=mcc))
        if val_loss >= min_loss + delta:
            cont += 1
            log(f"EarlyStopping counter: {cont} out of {patience}")
            if cont >= patience:
                log("Early Stopping! \t Training Stopped")
                break
        else:
            cont, min_loss = 0, val_loss
            log("Saving best model...")
            net.save(os.path.join(results_path, save_name))
    for w_ in (tw, vw):
        if w_ is not None:
            w_.close()
    return net, history
