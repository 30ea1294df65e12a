"""Patch dataset reader / training shell against the reference's synchronous loop
(train_ISPRS.py:97-189): same batches in the same order, remainder dropped, early stopping and best-model save."""
import os

import numpy as np
import pytest

from resuneta_b200 import data as D


def _make(root, n=11, hw=8, c=3, k=4, seed=0):
    r = np.random.RandomState(seed)
    x = r.rand(n, hw, hw, c).astype(np.float32)
    y = {"seg": np.eye(k, dtype=np.float32)[r.randint(0, k, (n, hw, hw))],
         "bound": (r.rand(n, hw, hw, k) < 0.1).astype(np.float32),
         "dist": r.rand(n, hw, hw, k).astype(np.float32),
         "color": r.rand(n, hw, hw, 3).astype(np.float32)}
    D.save_patch_dataset(root, x, y)
    return x, y


def test_layout_and_pairing(tmp_path):
    x, y = _make(str(tmp_path))
    xp, yp = D.list_patch_dataset(str(tmp_path), multitasking=True)
    assert [os.path.basename(p) for p in xp] == [f"patch_{i}.npy" for i in range(11)]       # natural order, paired by name
    assert set(yp) == {"seg", "bound", "dist", "color"}
    assert all(os.path.basename(a) == os.path.basename(b) for a, b in zip(xp, yp["dist"]))
    xs, ys = D.list_patch_dataset(str(tmp_path), multitasking=False)
    assert list(ys) == ["seg"]
    os.remove(yp["color"][3])
    with pytest.raises(FileNotFoundError):
        D.list_patch_dataset(str(tmp_path), multitasking=True)


def test_split_matches_sklearn_on_paired_lists(tmp_path):
    from sklearn.model_selection import train_test_split
    _make(str(tmp_path), n=10)
    xp, yp = D.list_patch_dataset(str(tmp_path))
    xtr, ytr, xva, yva = D.train_val_split(xp, yp)
    ref = train_test_split(xp, yp["seg"], yp["bound"], yp["dist"], yp["color"], test_size=0.2, random_state=42)
    assert xtr == ref[0] and xva == ref[1] and ytr["seg"] == ref[2] and yva["seg"] == ref[3]
    assert ytr["color"] == ref[8] and yva["color"] == ref[9]


@pytest.mark.parametrize("workers,prefetch", [(1, 3), (4, 3), (8, 5)])
def test_loader_equals_the_synchronous_reference_loop(tmp_path, workers, prefetch):
    x, y = _make(str(tmp_path), n=11)
    xp, yp = D.list_patch_dataset(str(tmp_path))
    B = 3
    ld = D.PatchBatchLoader(xp, yp, B, shuffle=False, workers=workers, prefetch=prefetch)
    assert len(ld) == 11 // B
    got = [(xb.numpy().copy(), {h: v.numpy().copy() for h, v in yb.items()}) for xb, yb in ld]
    assert len(got) == 3                                                    # remainder (2 patches) dropped
    for b, (xb, yb) in enumerate(got):
        ref_x = np.stack([np.load(p) for p in xp[b * B:(b + 1) * B]])       # train_ISPRS.py:121-123
        np.testing.assert_array_equal(xb, ref_x)
        for h in yp:
            np.testing.assert_array_equal(yb[h], np.stack([np.load(p).astype(np.float32) for p in yp[h][b * B:(b + 1) * B]]))


def test_shuffle_is_seeded_and_changes_per_epoch(tmp_path):
    _make(str(tmp_path), n=12)
    xp, yp = D.list_patch_dataset(str(tmp_path))
    ld = D.PatchBatchLoader(xp, yp, 4, shuffle=True, seed=7)
    e0 = np.concatenate([xb.numpy().copy() for xb, _ in ld])
    e1 = np.concatenate([xb.numpy().copy() for xb, _ in ld])
    ld2 = D.PatchBatchLoader(xp, yp, 4, shuffle=True, seed=7)
    f0 = np.concatenate([xb.numpy().copy() for xb, _ in ld2])
    np.testing.assert_array_equal(e0, f0)
    assert not np.array_equal(e0, e1)
    assert sorted(e0.reshape(12, -1).sum(1).tolist()) == pytest.approx(sorted(e1.reshape(12, -1).sum(1).tolist()))


def test_rank_sharding_partitions_the_batches(tmp_path):
    _make(str(tmp_path), n=16)
    xp, yp = D.list_patch_dataset(str(tmp_path))
    full = [xb.numpy().copy() for xb, _ in D.PatchBatchLoader(xp, yp, 2)]
    r0 = [xb.numpy().copy() for xb, _ in D.PatchBatchLoader(xp, yp, 2, rank=0, world=2)]
    r1 = [xb.numpy().copy() for xb, _ in D.PatchBatchLoader(xp, yp, 2, rank=1, world=2)]
    assert len(r0) == len(r1) == 4
    for i in range(4):
        np.testing.assert_array_equal(r0[i], full[2 * i])
        np.testing.assert_array_equal(r1[i], full[2 * i + 1])


def test_non_float32_and_bad_shapes(tmp_path):
    root = str(tmp_path)
    _make(root, n=4)
    xp, yp = D.list_patch_dataset(root)
    np.save(yp["seg"][1], np.load(yp["seg"][1]).astype(np.uint8))          # the reference casts with .astype(float32)
    ref = np.load(yp["seg"][1]).astype(np.float32)
    got = [yb["seg"].numpy().copy() for _, yb in D.PatchBatchLoader(xp, yp, 4)]
    np.testing.assert_array_equal(got[0][1], ref)
    np.save(xp[2], np.zeros((5, 5, 3), np.float32))
    with pytest.raises(ValueError):
        list(D.PatchBatchLoader(xp, yp, 4))


def test_mcc_and_training_shell_early_stopping(tmp_path):
    assert D.compute_mcc(50, 40, 5, 5) == pytest.approx((50 * 40 - 25) / np.sqrt(55 * 55 * 45 * 45))
    assert D.compute_mcc(0, 0, 0, 0) == 0.0

    class FakeNet:
        metrics_names = ["loss", "seg_loss", "seg_accuracy", "seg_true_positives", "seg_true_negatives",
                         "seg_false_positives", "seg_false_negatives"]

        def __init__(self, val_losses):
            self.val_losses, self.e, self.saved = val_losses, 0, []

        def train_on_batch(self, x, y, return_dict=False):
            return [1.0, 1.0, 0.5, 1, 1, 1, 1]

        def test_on_batch(self, x, y):
            return [self.val_losses[self.e], 0.3, 0.6, 8, 6, 1, 1]

        def save(self, path):
            self.saved.append((self.e, path))

    class Train:                       # starting a training epoch advances the fake's epoch counter
        def __iter__(self):
            net.e += 1
            return iter([(0, 0)] * 3)

    net = FakeNet([1.0, 0.8, 0.8005 + 0.05, 0.81, 0.82, 0.5])
    net.e = -1
    lines = []
    _, hist = D.train_model(net, Train(), [(0, 0)] * 2, epochs=6, results_path=str(tmp_path / "res"), patience=3,
                            delta=0.001, log=lines.append, tensorboard_dir=str(tmp_path / "tb"))
    # epochs 0,1 improve (saved); 2,3,4 are within/above min+delta -> counter 3 -> stop before epoch 5
    assert [e for e, _ in net.saved] == [0, 1]
    assert len(hist) == 5 and any("Early Stopping" in l for l in lines)
    assert hist[0]["mcc"] == pytest.approx(D.compute_mcc(8, 6, 1, 1))
    # TensorBoard scalars with the reference's tags (train_ISPRS.py:35-53), one event file per writer
    from tensorboard.backend.event_processing.event_accumulator import EventAccumulator
    for split, want in (("train", {"Segmentation/Loss", "Segmentation/Accuracy", "Total/Loss"}),
                        ("val", {"Segmentation/Loss", "Segmentation/Accuracy", "Segmentation/MCC", "Total/Loss"})):
        acc = EventAccumulator(str(tmp_path / "tb" / split))
        acc.Reload()
        assert want <= set(acc.Tags()["scalars"]), (split, acc.Tags()["scalars"])
        assert len(acc.Scalars("Total/Loss")) == 5
