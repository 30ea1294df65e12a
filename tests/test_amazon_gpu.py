"""GPU Amazon post-processing (csrc/postproc.cu through amazon.py) against the oracle: integer / index work, bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import amazon_oracle as AO  # noqa: E402


@pytest.fixture(scope="module")
def A():
    import __graft_entry__ as g
    g.build()
    from resuneta_b200 import amazon
    return amazon


def test_chop_paste_match_reference_loops(A):
    r = np.random.RandomState(0)
    img, img3 = r.rand(70, 53), r.rand(70, 53, 4)
    np.testing.assert_array_equal(A.extrac_patch2(img, 16, 1), AO.extrac_patch2(img, 16, 1))
    np.testing.assert_array_equal(A.extrac_patch2(img3, 16, 2), AO.extrac_patch2(img3, 16, 2))
    p = AO.extrac_patch2(img, 16, 1)
    np.testing.assert_array_equal(A.pred_recostruction(16, p, img), AO.pred_recostruction(16, p, img))


@pytest.mark.parametrize("H,W,density,thr", [(48, 64, 0.45, 4), (200, 333, 0.55, 69), (257, 129, 0.62, 25), (64, 64, 1.0, 10),
                                             (64, 64, 0.0, 1)])
def test_area_opening_bit_exact(A, H, W, density, thr):
    r = np.random.RandomState(H + W)
    img = (r.rand(H, W) < density).astype(np.float64)
    np.testing.assert_array_equal(A.area_opening(img, thr, connectivity=1), AO.area_opening_binary(img, thr))


def test_large_map_properties(A):
    """Full-size property checks (the flood-fill oracle is too slow at scene size): idempotence, monotonicity in the
    threshold, every surviving component is large enough (recount with scipy)."""
    from scipy import ndimage
    r = np.random.RandomState(3)
    img = (r.rand(2048, 2048) < 0.57).astype(np.uint8)
    a = A.area_opening(img, 69)
    np.testing.assert_array_equal(A.area_opening(a, 69), a)
    b = A.area_opening(img, 200)
    assert (b <= a).all() and (a <= img).all()
    lab, n = ndimage.label(a)
    sizes = np.bincount(lab.ravel())[1:]
    assert sizes.min() >= 69
    lab0, n0 = ndimage.label(img)
    s0 = np.bincount(lab0.ravel())
    keep = s0 >= 69
    keep[0] = False
    np.testing.assert_array_equal(a, keep[lab0].astype(np.uint8))


def test_consider_and_threshold_sweep(A):
    r = np.random.RandomState(5)
    H, W = 96, 128
    prob = r.rand(H, W)
    prob[20:60, 30:90] += 0.6
    ref = np.zeros((H, W)); ref[25:55, 35:85] = 1; ref[70:80, :] = 2
    tiles = np.ones((H, W)); tiles[:, 100:] = 0
    rec = (prob >= 0.5).astype(np.float64)
    rf, pf, cm = A.consider(rec, ref, tiles, 20)
    orf, opf, _ = AO.consider(rec, ref, tiles, 20)
    np.testing.assert_array_equal(rf, orf)
    np.testing.assert_array_equal(pf, opf)
    assert cm.sum() == len(orf) and cm[1, 1] == np.sum((orf == 1) & (opf == 1)) and cm[0, 1] == np.sum((orf == 0) & (opf == 1))
    got = A.matrics_AA_recall([0.3, 0.5, 0.9], prob, ref, tiles, 20)
    want = np.array([AO.metrics_aa_recall_one(t, prob, ref, tiles, 20) for t in (0.3, 0.5, 0.9)])
    np.testing.assert_allclose(got, want, rtol=1e-12, equal_nan=True)
