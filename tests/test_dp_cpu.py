"""Data-parallel path with world_size 2 over gloo on CPU (host logic of distribute.py): bucket schedule,
overlapped all-reduce, 1/world gradient scaling, parameter broadcast, per-replica BatchNorm."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_CLS, HW = 4, 64
LW = dict(seg=1.0, bound=0.5, dist=1.0, color=1.0)


def _worker(rank, world, port, q, overlap):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    torch.set_num_threads(2)
    import resuneta_b200  # noqa: F401
    from emul_lib import EmulLib
    from oracle import resuneta_oracle as O
    from resuneta_b200 import SGD, Tanimoto_dual_loss, _capi
    from resuneta_b200.builder import build_model
    from resuneta_b200.distribute import MirroredStrategy
    _capi.set_lib(EmulLib())
    strat = MirroredStrategy(backend="gloo", n_buckets=5, overlap=overlap)
    assert strat.num_replicas_in_sync == world
    with strat.scope():
        # different seeds per rank: compile() must broadcast rank 0's parameters
        m = build_model((HW, HW, 3), N_CLS, True, "v2", dtype="fp32", seed=100 + rank)
        m.compile(optimizer=SGD(lr=1e-2, momentum=0.0), loss={k: Tanimoto_dual_loss() for k in LW}, loss_weights=LW)
    w0 = {k: v.clone() for k, v in m.net.get_weights().items()}
    x, y = O.synth_batch(2, HW, 3, N_CLS, seed=50 + rank, block=16)
    res = m.train_on_batch(x, y)
    sched = m.dp._schedule(m.net.plan(2, True, m.loss_spec), m.net.params)
    q.put((rank, {k: v.numpy() for k, v in w0.items()}, {k: v.numpy() for k, v in m.net.get_weights().items()}, res,
           [b[0] for b in sched]))
    import torch.distributed as dist
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.timeout(600)
@pytest.mark.parametrize("overlap", [True, False])
def test_two_rank_data_parallel_step_matches_oracle_average(overlap):
    from oracle import resuneta_oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, overlap)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=500) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, w0a, w1a, resa, scha), (_, w0b, w1b, resb, schb) = out
    # broadcast: both ranks started from rank 0's parameters
    for k in w0a:
        np.testing.assert_array_equal(w0a[k], w0b[k])
    # replicas stay identical after the step (trainable parameters)
    for k in w1a:
        if O.is_trainable(k):
            np.testing.assert_array_equal(w1a[k], w1b[k])
    # buckets are issued in backward completion order
    assert scha == sorted(scha) and len(scha) == 5
    # per-replica losses differ (different shards), and the update is p - lr * mean_r(grad_r)
    assert abs(resa[0] - resb[0]) > 1e-6
    p64 = {k: torch.from_numpy(v).double() for k, v in w0a.items()}
    gsum = None
    for r in range(2):
        x, y = O.synth_batch(2, HW, 3, N_CLS, seed=50 + r, block=16)
        y64 = {k: torch.from_numpy(v).double() for k, v in y.items()}
        _, _, _, g, _ = O.loss_and_grads(p64, torch.from_numpy(x).double(), y64,
                                         {k: O.tanimoto_dual_loss for k in LW}, LW, N_CLS)
        gsum = g if gsum is None else {k: gsum[k] + g[k] for k in g}
    gmax = max(v.norm().item() for v in gsum.values()) / 2
    for k, g in gsum.items():
        want = p64[k] - 1e-2 * g / 2
        err = (torch.from_numpy(w1a[k]).double() - want).norm().item()
        assert err <= 1e-2 * (2e-3 * g.norm().item() / 2 + 1e-5 * gmax) + 1e-6 * want.norm().item(), (k, err)


def _scene_worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    torch.set_num_threads(2)
    import torch.distributed as dist
    import resuneta_b200  # noqa: F401
    from emul_lib import EmulLib
    from resuneta_b200 import _capi, inference
    from resuneta_b200.builder import build_model
    _capi.set_lib(EmulLib())
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = build_model((HW, HW, 3), N_CLS, True, "v2", dtype="fp32", seed=3)      # same seed: identical replicas
    scene, ref = _scene()
    r = inference.predict_scene(m, scene, ref, patch_size=HW, batch_size=2, num_classes=N_CLS)
    q.put((rank, r["seg_pred"], r["confusion_full"], r["reconstructed"]))
    dist.barrier()
    dist.destroy_process_group()


def _scene():
    rs = np.random.RandomState(4)
    return rs.rand(3 * HW + 5, 2 * HW + 9, 3).astype(np.float32), rs.randint(0, N_CLS, (3 * HW + 5, 2 * HW + 9))


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world", [2, 4])
def test_scene_inference_shards_patches_and_sums_confusion(world):
    """SURVEY §8e, inference: the 6 patches of the scene are split over the ranks (4 ranks: 2+2+2+0), the int64 confusion
    matrices are all-reduced and the label tiles all-gathered - every rank must return the single-process result."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import resuneta_b200  # noqa: F401
    from emul_lib import EmulLib
    from resuneta_b200 import _capi, inference
    from resuneta_b200.builder import build_model
    old = _capi._LIB
    _capi.set_lib(EmulLib())
    try:
        m = build_model((HW, HW, 3), N_CLS, True, "v2", dtype="fp32", seed=3)
        scene, ref = _scene()
        want = inference.predict_scene(m, scene, ref, patch_size=HW, batch_size=4, num_classes=N_CLS)
    finally:
        _capi.set_lib(old)
    assert want["seg_pred"].shape == (6, HW, HW) and want["confusion_full"].sum() == 6 * HW * HW
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_scene_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=500) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, seg, cm, rec in out:
        np.testing.assert_array_equal(seg, want["seg_pred"])
        np.testing.assert_array_equal(cm, want["confusion_full"])
        np.testing.assert_array_equal(rec, want["reconstructed"])
