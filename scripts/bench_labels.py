"""Throughput of the §8f rows next to their CPU counterparts (run on the GPU box):
  * multitask label generation: GPU kernels (labels.cu) vs OpenCV on the host cores (the reference's path,
    multitasking_utils.py:6-34) when cv2 is importable, else the numpy oracle on a bounded sample;
  * patch loader: batches/s of PatchBatchLoader vs the reference's synchronous np.load loop (train_ISPRS.py:115-141).
Usage: python scripts/bench_labels.py  -> gpurun_out/bench_labels.txt"""
import os, sys, time, tempfile, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import __graft_entry__ as ge; ge.build()
from resuneta_b200 import labels as L, data as D
from oracle import labels_oracle as LO
out = []
N, hw, n = 16, 256, 6
r = np.random.RandomState(0)
cls = r.randint(0, n, (N, hw // 16, hw // 16)).repeat(16, 1).repeat(16, 2)
onehot = np.eye(n, dtype=np.float32)[cls]
rgb = r.randint(0, 256, (N, hw, hw, 3)).astype(np.uint8)
gen = L.LabelGenerator(N, hw, hw, n)
oh_d, rgb_d = torch.from_numpy(onehot).cuda(), torch.from_numpy(rgb).cuda()
for _ in range(3): gen.multitask_targets(oh_d, rgb_d)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): y = gen.multitask_targets(oh_d, rgb_d)
e1.record(); torch.cuda.synchronize()
gpu_ms = e0.elapsed_time(e1) / 10
out.append(f"label generation, batch {N} x {hw}x{hw}x{n} one-hot (+RGB): GPU {gpu_ms:.3f} ms/batch = {N / gpu_ms * 1e3:.0f} patches/s")
try:
    import cv2
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden_labels as G
    t0 = time.perf_counter()
    for i in range(N):
        G.ref_boundary(onehot[i]); G.ref_distance(onehot[i]); G.ref_color(rgb[i])
    cpu_ms = (time.perf_counter() - t0) * 1e3
    out.append(f"  OpenCV {cv2.__version__} on 1 host thread (the reference's per-patch loop): {cpu_ms:.1f} ms/batch = {N / cpu_ms * 1e3:.0f} patches/s  -> x{cpu_ms / gpu_ms:.0f}")
    b = y["bound"].cpu().numpy(); d = y["dist"].cpu().numpy(); c = y["color"].cpu().numpy()
    ok_b = all(np.array_equal(b[i], G.ref_boundary(onehot[i])) for i in range(N))
    ok_c = all(np.array_equal(c[i], G.ref_color(rgb[i])) for i in range(N))
    err_d = max(np.abs(d[i] - G.ref_distance(onehot[i])).max() for i in range(N))
    out.append(f"  parity on this batch vs OpenCV: boundary bit-exact={ok_b}, colour bit-exact={ok_c}, distance max abs err={err_d:.2e}")
except ImportError:
    t0 = time.perf_counter(); LO.get_boundary_label(onehot[0]); LO.get_color_label(rgb[0]); cpu_ms = (time.perf_counter() - t0) * 1e3
    out.append(f"  cv2 not importable; numpy oracle boundary+colour of ONE patch: {cpu_ms:.1f} ms")
# ---- loader
root = tempfile.mkdtemp(prefix="rsa_patches_")
try:
    M = 64
    x = r.rand(M, hw, hw, 3).astype(np.float32)
    yy = {"seg": np.eye(n, dtype=np.float32)[r.randint(0, n, (M, hw, hw))], "bound": r.rand(M, hw, hw, n).astype(np.float32),
          "dist": r.rand(M, hw, hw, n).astype(np.float32), "color": r.rand(M, hw, hw, 3).astype(np.float32)}
    D.save_patch_dataset(root, x, yy)
    xp, yp = D.list_patch_dataset(root)
    B = 16
    t0 = time.perf_counter()
    for b in range(M // B):
        xb = np.stack([np.load(f) for f in xp[b * B:(b + 1) * B]])
        yb = {h: np.stack([np.load(f).astype(np.float32) for f in yp[h][b * B:(b + 1) * B]]) for h in yp}
    sync_ms = (time.perf_counter() - t0) * 1e3 / (M // B)
    ld = D.PatchBatchLoader(xp, yp, B, workers=16, prefetch=4)
    for _ in ld: pass                                   # warm page cache / allocate pinned slots
    t0 = time.perf_counter()
    for _ in range(3):
        for xb, yb in ld: pass
    ld_ms = (time.perf_counter() - t0) * 1e3 / (3 * (M // B))
    mb = (x[0].nbytes + sum(v[0].nbytes for v in yy.values())) * B / 1e6
    out.append(f"patch loader, batch {B} ({mb:.0f} MB of .npy per batch, page cache warm): reference-style np.load loop {sync_ms:.1f} ms/batch; "
               f"PatchBatchLoader(16 threads, pinned ring) {ld_ms:.1f} ms/batch = {B / ld_ms * 1e3:.0f} patches/s ({mb / ld_ms:.1f} GB/s)")
finally:
    shutil.rmtree(root, ignore_errors=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "bench_labels.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
