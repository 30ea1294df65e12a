"""Oracle of the Amazon post-processing (oracle/amazon_oracle.py): chop/paste conventions and the area opening against
scipy.ndimage.label, an independent connected-component implementation (scikit-image itself is not installed)."""
import numpy as np
import pytest
from scipy import ndimage

from oracle import amazon_oracle as AO


def test_column_major_chop_and_paste_roundtrip():
    img = np.arange(7 * 10).reshape(7, 10).astype(np.float64)
    p = AO.extrac_patch2(img, 3, 1)
    assert p.shape == (2 * 3, 3, 3)
    np.testing.assert_array_equal(p[0], img[0:3, 0:3])
    np.testing.assert_array_equal(p[1], img[3:6, 0:3])            # row index runs fastest: column-major tile order
    np.testing.assert_array_equal(p[2], img[0:3, 3:6])
    rec = AO.pred_recostruction(3, p, img)
    np.testing.assert_array_equal(rec, img[:6, :9])


@pytest.mark.parametrize("seed,thr", [(0, 1), (1, 4), (2, 9), (3, 69)])
def test_area_opening_matches_scipy_components(seed, thr):
    r = np.random.RandomState(seed)
    img = (r.rand(48, 64) < 0.45).astype(np.float64)
    lab, n = ndimage.label(img)                                   # default structure = 4-connectivity
    sizes = np.bincount(lab.ravel(), minlength=n + 1)
    keep = sizes >= thr
    keep[0] = False
    np.testing.assert_array_equal(AO.area_opening_binary(img, thr), keep[lab].astype(np.float64))


def test_consider_pipeline_small_case():
    rec = np.zeros((6, 8)); rec[0, 0] = 1; rec[2:5, 2:6] = 1       # a 1-pixel blob (dropped at area 4) and a 12-pixel blob
    ref = np.zeros((6, 8)); ref[2:4, 2:6] = 1; ref[5, :] = 2        # past deforestation on the last row
    clip = np.ones((6, 8)); clip[:, 7] = 0
    ref_f, pre_f, m = AO.consider(rec, ref, clip, 4)
    assert m[0, 0] == 0 and m[5, 0] == 0 and m[3, 3] == 1
    assert len(ref_f) == 6 * 8 - 1 - 8 - 5                         # minus the small blob, the past-deforestation row, column 7 (5 rows left)
    assert pre_f.sum() == 12 and ref_f.sum() == 8


def test_alarm_area_counts_masked_pixels_in_the_denominator():
    """utils2.py:335-336,352: ref_final = ref_consider[mask_amazon_ts == 1] keeps the pixels the area filter / the border
    mask zeroed, so the alarm area divides by the number of mask_amazon_ts == 1 pixels."""
    prob = np.zeros((6, 8)); prob[0, 0] = 0.9; prob[2:5, 2:6] = 0.8
    ref = np.zeros((6, 8)); ref[2:4, 2:6] = 1; ref[5, :] = 2
    mask = np.ones((6, 8)); mask[:, 7] = 0
    rec, prec, aa = AO.metrics_aa_recall_one(0.5, prob, ref, mask, 4)
    assert rec == 1.0 and prec == pytest.approx(8 / 12)
    assert aa == pytest.approx(12 / (6 * 7))          # 42 selected pixels, not the 34 that survive the masks
