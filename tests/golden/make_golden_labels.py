"""Generates tests/golden/labels_kat.npz with OpenCV (the library behind the reference's label generation,
multitasking_utils.py:6-34, preprocess_save_patches_ISPRS.py:224-228).  Run where cv2 is importable:
    python tests/golden/make_golden_labels.py"""
import os
import sys

import cv2
import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))


def ref_boundary(label):       # multitasking_utils.py:6-22, verbatim call sequence
    out = np.empty_like(label, dtype=np.float32)
    for c in range(label.shape[2]):
        t = label.astype(np.uint8)
        e = cv2.Canny(t[:, :, c], 0, 1)
        e = cv2.dilate(e, cv2.getStructuringElement(cv2.MORPH_CROSS, (3, 3)), iterations=1)
        out[:, :, c] = e.astype(np.float32) / 255.
    return out


def ref_distance(label):       # multitasking_utils.py:25-34
    out = np.empty_like(label, dtype=np.float32)
    for c in range(label.shape[2]):
        d = cv2.distanceTransform(label[:, :, c].astype(np.uint8), cv2.DIST_L2, 0)
        out[:, :, c] = cv2.normalize(d, d, 0, 1.0, cv2.NORM_MINMAX)
    return out


def ref_color(img):            # preprocess_save_patches_ISPRS.py:224-228 with normalize_hsv norm_type 1 (:89-94)
    hsv = cv2.cvtColor(img, cv2.COLOR_RGB2HSV).astype(np.float32)
    hsv[:, :, 0] /= 179.
    hsv[:, :, 1] /= 255.
    hsv[:, :, 2] /= 255.
    return hsv


def cases():
    r = np.random.RandomState(20261017)
    out = []
    for hw, k, ncls in [(32, 4, 4), (32, 1, 3), (48, 8, 5), (64, 16, 6), (64, 2, 2)]:
        lab = r.randint(0, ncls, (hw // k, hw // k)).repeat(k, 0).repeat(k, 1)
        onehot = np.eye(ncls + 1, dtype=np.float32)[lab]        # last class absent: an all-zero plane
        img = r.randint(0, 256, (hw, hw, 3)).astype(np.uint8)
        img[: hw // 4] = (r.randint(0, 4, (hw // 4, hw, 3)) * 85).astype(np.uint8)      # ties between channels
        out.append((onehot, img))
    return out


if __name__ == "__main__":
    d = {}
    for i, (onehot, img) in enumerate(cases()):
        d[f"label_{i}"] = onehot.astype(np.uint8)
        d[f"img_{i}"] = img
        d[f"bound_{i}"] = ref_boundary(onehot)
        d[f"dist_{i}"] = ref_distance(onehot)
        d[f"color_{i}"] = ref_color(img)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "labels_kat.npz")
    np.savez_compressed(path, **d)
    print("wrote", path, os.path.getsize(path), "bytes; OpenCV", cv2.__version__)
