"""Two-rank check of the data-parallel step (run with torchrun on 2 GPUs): the graph-replayed step with the all-reduce
overlapped with backward must give the same losses and parameters as the non-overlapped one, and both ranks must hold
identical parameters afterwards."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
import __graft_entry__ as ge
if rank == 0: ge.build()
from oracle import resuneta_oracle as O
from resuneta_b200 import SGD, Tanimoto_dual_loss
from resuneta_b200.builder import build_model
from resuneta_b200.distribute import MirroredStrategy
strat = MirroredStrategy()
dist.barrier()
heads = ("seg", "bound", "dist", "color")
hw, n, B = 64, 4, 4
x, y = O.synth_batch(B, hw, 3, n, seed=100 + rank, block=16)
res = {}
for mode in ("1", "0", "0b", "0s"):   # "0b": the plain mode again = run-to-run noise floor (atomics reorder fp32 sums)
    os.environ["RSA_DP_GRAPH_OVERLAP"] = mode[0]
    # "0s": everything on one stream (no weight-gradient side stream, no branch lanes) = the serial reference
    os.environ["RSA_WGRAD_STREAM"], os.environ["RSA_LANES"] = ("0", "0") if mode == "0s" else ("1", "2")
    with strat.scope():
        m = build_model((hw, hw, 3), n, True, "v2", dtype="bf16", seed=7)
        m.compile(optimizer=SGD(lr=1e-2, momentum=0.5), loss={h: Tanimoto_dual_loss() for h in heads})
    pl = m.net.plan(B, True, m.loss_spec)
    split = m.dp.phase_splits(pl, m.net.params) if mode == "1" else None
    out = [m.train_on_batch(x, y) for _ in range(4)]
    # the device-resident path of bench.py (_execute) as well
    m._push_lr(); m._execute(pl, True); torch.cuda.synchronize()
    p = m.net.params.data[:m.net.params.n_train].double()
    chk = torch.stack([p.sum(), (p * p).sum()])
    allchk = [torch.zeros_like(chk) for _ in range(2)]
    dist.all_gather(allchk, chk)
    res[mode] = (np.array(out), chk.cpu().numpy(), [c.cpu().numpy() for c in allchk], split)
if rank == 0:
    a, b = res["1"], res["0"]
    print("split (launch index, offset) =", a[3], "of", len(pl.bwd), "backward launches /", m.net.params.n_train, "gradient elements")
    same_ranks = all(np.allclose(a[2][0], a[2][1], rtol=1e-12) for _ in [0]) and np.allclose(b[2][0], b[2][1], rtol=1e-12)
    c = res["0b"]
    noise_l = np.abs(b[0] - c[0]).max(); noise_p = np.abs(b[1] - c[1]).max()
    dl = np.abs(a[0] - b[0]).max(); dp_ = np.abs(a[1] - b[1]).max()
    print(f"max |loss diff| overlap-vs-plain {dl:.3e} (plain-vs-plain noise {noise_l:.3e}); checksum diff {dp_:.3e} (noise {noise_p:.3e})")
    loss_ok = dl <= 5 * noise_l + 1e-6
    par_ok = dp_ <= 5 * noise_p + 1e-6 * np.abs(b[1]).max()
    sr = res["0s"]
    ds_l = np.abs(sr[0] - b[0]).max(); ds_p = np.abs(sr[1] - b[1]).max()
    ser_ok = ds_l <= 5 * noise_l + 1e-6 and ds_p <= 5 * noise_p + 1e-6 * np.abs(b[1]).max()
    print(f"multi-stream vs single-stream launches: max |loss diff| {ds_l:.3e}, checksum diff {ds_p:.3e}: {ser_ok}")
    print("ranks hold identical parameters:", same_ranks, "| losses overlap vs plain:", loss_ok, "| parameter checksums:", par_ok)
    print("losses (overlap):", a[0][:, 0], "(plain):", b[0][:, 0])
    print("DP CHECK", "PASS" if (same_ranks and loss_ok and par_ok and ser_ok and a[3]) else "FAIL")
dist.barrier(); dist.destroy_process_group()
