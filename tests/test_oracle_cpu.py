"""CPU tests that pin the oracle: reference known-answer vectors (KAT-M), the numpy direct-loop
cross-check of conventions, analytic known answers and the
committed golden regression vectors."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import numpy_loops as NL
from oracle import resuneta_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


# ---- KAT-M1..4: metrics recomputed from the reference's own printed confusion matrices -----
def test_kat_metrics_from_reference_results_file():
    d = json.load(open(os.path.join(GOLD, "kat_metrics.json")))
    assert len(d["blocks"]) == 5
    for b in d["blocks"]:
        cm = np.array(b["cm"], dtype=np.int64)
        assert cm.sum() == 6488064          # 99 patches of 256^2
        acc, f1, rec, prec = O.metrics_from_confusion(cm)
        assert abs(acc - b["accuracy"]) < 1e-9
        np.testing.assert_allclose(f1, b["f1"], rtol=0, atol=5e-7)
        np.testing.assert_allclose(rec, b["recall"], rtol=0, atol=5e-7)
        np.testing.assert_allclose(prec, b["precision"], rtol=0, atol=5e-7)


def test_confusion_matrix_matches_sklearn():
    from sklearn.metrics import confusion_matrix as skcm
    rng = np.random.RandomState(0)
    t = rng.randint(0, 6, 10000)
    p = rng.randint(0, 6, 10000)
    np.testing.assert_array_equal(O.confusion_matrix(t, p), skcm(t, p))
    # labels = union of values present: a missing class shrinks the matrix
    t2 = np.where(t == 3, 2, t)
    p2 = np.where(p == 3, 2, p)
    np.testing.assert_array_equal(O.confusion_matrix(t2, p2), skcm(t2, p2))
    assert O.confusion_matrix(t2, p2).shape == (5, 5)


def test_metrics_match_sklearn_compute_metrics():
    from sklearn.metrics import accuracy_score, f1_score, precision_score, recall_score
    rng = np.random.RandomState(1)
    t = rng.randint(0, 5, 5000)
    p = np.where(rng.rand(5000) < 0.7, t, rng.randint(0, 5, 5000))
    acc, f1, rec, prec = O.metrics_from_confusion(O.confusion_matrix(t, p))
    assert abs(acc - 100 * accuracy_score(t, p)) < 1e-9
    np.testing.assert_allclose(f1, 100 * f1_score(t, p, average=None), atol=1e-9)
    np.testing.assert_allclose(rec, 100 * recall_score(t, p, average=None), atol=1e-9)
    np.testing.assert_allclose(prec, 100 * precision_score(t, p, average=None), atol=1e-9)


# ---- conventions: torch oracle vs pure-numpy loops -----------------------------------------
@pytest.mark.parametrize("d", [1, 3, 5])
def test_conv_same_dilated_matches_numpy_loops(d):
    rng = np.random.RandomState(d)
    x = rng.randn(2, 9, 11, 3)
    w = rng.randn(3, 3, 3, 4)
    b = rng.randn(4)
    cx = O._Ctx({"c/kernel": torch.from_numpy(w), "c/bias": torch.from_numpy(b)}, False, False,
                torch.float64, None, None)
    y = cx.conv(torch.from_numpy(x).permute(0, 3, 1, 2), 4, 3, 1, d, "same", name="c")
    np.testing.assert_allclose(y.permute(0, 2, 3, 1).numpy(), NL.conv2d(x, w, b, 1, d, True), atol=1e-12)


def test_conv1x1_stride2_samples_even_pixels():
    rng = np.random.RandomState(3)
    x = rng.randn(1, 7, 8, 2)
    w = rng.randn(1, 1, 2, 3)
    b = rng.randn(3)
    cx = O._Ctx({"c/kernel": torch.from_numpy(w), "c/bias": torch.from_numpy(b)}, False, False,
                torch.float64, None, None)
    y = cx.conv(torch.from_numpy(x).permute(0, 3, 1, 2), 3, 1, 2, name="c").permute(0, 2, 3, 1).numpy()
    np.testing.assert_allclose(y, NL.conv2d(x, w, b, 2, 1, False), atol=1e-12)
    np.testing.assert_allclose(y, x[:, 0::2, 0::2, :] @ w[0, 0] + b, atol=1e-12)


def test_pool_upsample_bn_match_numpy_loops():
    rng = np.random.RandomState(4)
    x = rng.randn(2, 8, 8, 3)
    xt = torch.from_numpy(x).permute(0, 3, 1, 2)
    for k in (2, 4, 8):
        p = O._pool(xt, k)
        np.testing.assert_array_equal(p.permute(0, 2, 3, 1).numpy(), NL.maxpool(x, k))
        u = O._up(p, k)
        np.testing.assert_array_equal(u.permute(0, 2, 3, 1).numpy(), NL.upsample(NL.maxpool(x, k), k))
        assert u.shape == xt.shape                                  # KAT-P1 (notebook cell 2)
    g, be = rng.rand(3) + 0.5, rng.randn(3)
    P = {"batch_normalization/gamma": torch.from_numpy(g), "batch_normalization/beta": torch.from_numpy(be),
         "batch_normalization/moving_mean": torch.zeros(3, dtype=torch.float64),
         "batch_normalization/moving_variance": torch.ones(3, dtype=torch.float64)}
    st = {}
    cx = O._Ctx(P, True, False, torch.float64, None, st)
    y = cx.bn(xt).permute(0, 2, 3, 1).numpy()
    yr, mean, var = NL.bn_train(x, g, be)
    np.testing.assert_allclose(y, yr, atol=1e-12)
    n = 2 * 8 * 8
    np.testing.assert_allclose(st["batch_normalization/moving_mean"].numpy(), 0.01 * mean, atol=1e-12)
    np.testing.assert_allclose(st["batch_normalization/moving_variance"].numpy(),
                               0.99 + 0.01 * var * n / (n - 1), atol=1e-12)
    # G4: train-mode output has mean beta and variance gamma^2 var/(var+eps)
    np.testing.assert_allclose(y.reshape(-1, 3).mean(0), be, atol=1e-12)
    np.testing.assert_allclose(y.reshape(-1, 3).var(0), g ** 2 * var / (var + 1e-3), atol=1e-12)


def test_tanimoto_matches_numpy_loops():
    rng = np.random.RandomState(5)
    lab = np.eye(4)[rng.randint(0, 4, (3, 6, 6))]
    pred = rng.rand(3, 6, 6, 4)
    pred /= pred.sum(-1, keepdims=True)
    a = O.tanimoto_dual_loss(torch.from_numpy(lab), torch.from_numpy(pred)).numpy()
    np.testing.assert_allclose(a, NL.tanimoto_dual(lab, pred), rtol=1e-12)


# ---- analytic known answers ------------------------------------------------------------------
def _onehot(rng, b, h, c):
    return np.eye(c)[rng.randint(0, c, (b, h, h))]


def test_T1_dual_of_identical_onehot_is_zero():
    y = torch.from_numpy(_onehot(np.random.RandomState(0), 2, 8, 3))
    assert O.tanimoto_dual_loss(y, y).abs().max().item() < 1e-12


def test_T2_rolled_onehot():
    y = _onehot(np.random.RandomState(0), 2, 8, 3)
    p = np.roll(y, 1, axis=-1)
    v = O.tanimoto_dual_loss(torch.from_numpy(y), torch.from_numpy(p))
    assert ((v > 0.7) & (v < 0.95)).all()      # ~0.83: complement term keeps it below 1


def test_T3_single_class_label_stays_finite():
    y = np.zeros((2, 8, 8, 3))
    y[..., 0] = 1.0                              # classes 1,2 absent -> V=0 -> inf -> max-weight path
    p = np.random.RandomState(1).rand(2, 8, 8, 3)
    p /= p.sum(-1, keepdims=True)
    v = O.tanimoto_dual_loss(torch.from_numpy(y), torch.from_numpy(p))
    assert torch.isfinite(v).all()
    # all-inf weights (label all ones -> 1-label all zeros): weights collapse to 0 -> T = 1
    ones = torch.ones(1, 4, 4, 2, dtype=torch.float64)
    t = O.tanimoto_loss(1 - ones, torch.rand(1, 4, 4, 2, dtype=torch.float64))
    assert abs(t.item() - 1.0) < 1e-12


def test_T4_gradient_flows_through_prediction_derived_weights():
    rng = np.random.RandomState(2)
    y = torch.from_numpy(_onehot(rng, 2, 8, 4))
    z = torch.from_numpy(rng.randn(2, 8, 8, 4)).requires_grad_(True)
    p = torch.softmax(z, -1)
    g_full, = torch.autograd.grad(O.tanimoto_dual_loss(y, p).mean(), z, retain_graph=True)

    def detached(label, pred):
        vli = pred.detach().sum(dim=(1, 2)).mean(0)
        w = 1 / vli ** 2
        sp = (pred * label).sum((1, 2))
        sq = (pred ** 2 + label ** 2).sum((1, 2))
        l1 = ((w * sp).sum(-1) + 1e-5) / ((w * (sq - sp)).sum(-1) + 1e-5)
        return 1 - 0.5 * (l1 + O.tanimoto_loss(1 - label, 1 - pred))
    g_det, = torch.autograd.grad(detached(y, p).mean(), z)
    rel = ((g_full - g_det).norm() / g_full.norm()).item()
    assert rel > 0.02          # materially different: the kernel must differentiate through w


@pytest.mark.parametrize("variant", ["v2", "v1"])
def test_G1_resblock_with_zero_kernels(variant):
    rng = np.random.RandomState(6)
    f, dils = 4, [1, 3]
    P, x = {}, torch.from_numpy(rng.randn(2, f, 8, 8))
    cx = O._Ctx(P, True, True, torch.float64, torch.Generator().manual_seed(0), None)
    O._resblock(cx, x, f, dils, identity=variant == "v2")
    b2 = []
    for k in list(P):
        if k.endswith("/kernel"):
            P[k] = torch.zeros_like(P[k])
        if k.endswith("/bias"):
            P[k] = torch.from_numpy(rng.randn(f))
        if k.endswith("/gamma") or k.endswith("/beta"):
            P[k] = torch.from_numpy(rng.randn(f))
    b2 = [P["conv2d_1/bias"], P["conv2d_3/bias"]]
    cx = O._Ctx(P, True, False, torch.float64, None, None)
    out = O._resblock(cx, x, f, dils, identity=variant == "v2")
    want = (b2[0] + b2[1]).view(1, f, 1, 1) + (x if variant == "v2" else 0)
    np.testing.assert_allclose(out.numpy(), want.expand_as(out).numpy(), atol=1e-12)


def test_G2_conv1x1_commutes_with_nearest_upsample():
    rng = np.random.RandomState(7)
    x = torch.from_numpy(rng.randn(1, 3, 4, 4))
    w = torch.from_numpy(rng.randn(5, 3, 1, 1))
    a = torch.nn.functional.conv2d(O._up(x, 2), w)
    b = O._up(torch.nn.functional.conv2d(x, w), 2)
    assert torch.equal(a, b)


def test_G5_dilation_31_on_32x32_only_centre_and_strip():
    x = np.zeros((1, 32, 32, 1))
    x[0, 5, 7, 0] = 1.0
    w = np.ones((3, 3, 1, 1))
    y = NL.conv2d(x, w, np.zeros(1), 1, 31, True)[0, :, :, 0]
    # an impulse reaches only the centre tap plus positions exactly +-31 away (a 1-px strip)
    assert y[5, 7] == 1.0 and y.sum() <= 4.0
    cx = O._Ctx({"c/kernel": torch.from_numpy(w), "c/bias": torch.zeros(1, dtype=torch.float64)}, False,
                False, torch.float64, None, None)
    yt = cx.conv(torch.from_numpy(x).permute(0, 3, 1, 2), 1, 3, 1, 31, "same", name="c")[0, 0].numpy()
    np.testing.assert_array_equal(y, yt)


# ---- structure + golden regression ----------------------------------------------------------
def test_param_counts_match_survey():
    for variant, total, bn in (("v2", 42736773, 83), ("v1", 43246517, 62)):
        p = O.init_params((256, 256, 3), 6, True, variant)
        assert sum(v.numel() for v in p.values()) == total
        assert len([k for k in p if k.endswith("/gamma")]) == bn
        assert len([k for k in p if k.endswith("/kernel")]) == 98
    p = O.init_params((256, 256, 3), 6, True, "v2")
    assert p["conv2d_83/kernel"].shape == (1, 1, 64, 32)      # final combine
    assert p["conv2d_9/kernel"].shape == (1, 1, 32, 64)       # down1
    assert p["conv2d_43/kernel"].shape == (1, 1, 1024, 256)   # dec5 UpSampling conv


def test_golden_regression_small():
    g = np.load(os.path.join(GOLD, "oracle_small.npz"))
    for variant in ("v2", "v1"):
        p = O.init_params((64, 64, 3), 5, True, variant, seed=7)
        x, y = O.synth_batch(2, 64, 3, 5, seed=11, block=8)
        o = O.forward(p, torch.from_numpy(x), True, 5, True, variant)
        np.testing.assert_allclose(o["seg"].numpy()[:, ::2, ::2], g[f"{variant}_train_seg"], rtol=2e-4, atol=1e-6)
        np.testing.assert_allclose(o["color"].numpy()[:, ::2, ::2], g[f"{variant}_train_color"], rtol=2e-4, atol=1e-6)
        oi = O.forward(p, torch.from_numpy(x), False, 5, True, variant)
        np.testing.assert_allclose(oi["seg"].numpy()[:, ::2, ::2], g[f"{variant}_infer_seg"], rtol=2e-4, atol=1e-6)
        for k in oi:
            np.testing.assert_allclose(O.tanimoto_dual_loss(torch.from_numpy(y[k]), oi[k]).numpy(),
                                       g[f"{variant}_tanimoto_{k}"], rtol=1e-4)


def test_train_step_reduces_loss_and_updates_moving_stats():
    p = O.init_params((64, 64, 3), 4, True, "v2", seed=3)
    x, y = O.synth_batch(2, 64, 3, 4, seed=5, block=16)
    xt = torch.from_numpy(x)
    yt = {k: torch.from_numpy(v) for k, v in y.items()}
    losses = {k: O.tanimoto_dual_loss for k in yt}
    opt = O.Adam(lr=1e-3)
    first = O.train_on_batch(p, opt, xt, yt, losses, {}, 4)
    assert len(first) == 10
    for _ in range(3):
        last = O.train_on_batch(p, opt, xt, yt, losses, {}, 4)
    assert last[0] < first[0]
    assert not torch.allclose(p["batch_normalization/moving_mean"], torch.zeros(32))
    ev = O.test_on_batch(p, xt, yt, losses, {}, 4)
    assert len(ev) == 10 and np.isfinite(ev).all()
