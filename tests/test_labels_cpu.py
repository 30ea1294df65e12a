"""Label-generation oracle (oracle/labels_oracle.py) against OpenCV: the committed cv2-generated known answers
(tests/golden/labels_kat.npz, tests/golden/make_golden_labels.py) and, where cv2 is importable, cv2 itself on fresh
random inputs.  Boundary and colour targets are integer computations: bit-exact.  The distance target is float32:
OpenCV's optimised distanceTransform (IPP) does not round sqrt correctly everywhere (1-ulp differences, e.g. sqrt(37)), the
oracle uses the correctly rounded value; tolerance 1e-6 absolute on values in [0, 1] / 2e-6 relative on raw distances."""
import os

import numpy as np
import pytest

from oracle import labels_oracle as LO

KAT = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "labels_kat.npz"))
NCASE = len([k for k in KAT.files if k.startswith("label_")])


@pytest.mark.parametrize("i", range(NCASE))
def test_oracle_matches_opencv_known_answers(i):
    lab, img = KAT[f"label_{i}"].astype(np.float32), KAT[f"img_{i}"]
    np.testing.assert_array_equal(LO.get_boundary_label(lab), KAT[f"bound_{i}"])
    np.testing.assert_allclose(LO.get_distance_label(lab), KAT[f"dist_{i}"], rtol=0, atol=1e-6)
    np.testing.assert_array_equal(LO.get_color_label(img), KAT[f"color_{i}"])


def test_edge_cases():
    lab = np.zeros((16, 16, 2), np.float32)
    lab[..., 0] = 1.0                                    # class 0 covers the patch, class 1 is absent
    b, d = LO.get_boundary_label(lab), LO.get_distance_label(lab)
    assert b.max() == 0 and d.max() == 0
    lab[4:9, 5:11, 0] = 0
    lab[4:9, 5:11, 1] = 1
    d = LO.get_distance_label(lab)
    assert d[..., 1].max() == 1.0 and d[0, 0, 1] == 0.0 and d[6, 7, 1] == 1.0       # centre of the 5x6 box is the farthest


def test_against_cv2_when_available():
    cv2 = pytest.importorskip("cv2")
    r = np.random.RandomState(5)
    for t in range(40):
        k = int(r.choice([1, 2, 4, 8]))
        n = int(r.randint(2, 6))
        lab = r.randint(0, n, (64 // k, 64 // k)).repeat(k, 0).repeat(k, 1)
        m = (lab == 1).astype(np.uint8)
        np.testing.assert_array_equal(LO.canny_0_1(m), cv2.Canny(m, 0, 1))
        d = LO.edt_exact(m)
        if d is not None:
            dc = cv2.distanceTransform(m, cv2.DIST_L2, 0)
            np.testing.assert_allclose(d, dc, rtol=2e-6, atol=0)
            np.testing.assert_array_equal(LO.minmax_01(dc), cv2.normalize(dc.copy(), None, 0, 1.0, cv2.NORM_MINMAX))
        img = r.randint(0, 256, (32, 32, 3)).astype(np.uint8)
        np.testing.assert_array_equal(LO.rgb_to_hsv_u8(img), cv2.cvtColor(img, cv2.COLOR_RGB2HSV))
