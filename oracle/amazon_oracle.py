"""CPU restatement of the Amazon evaluation post-processing.  TEST INFRASTRUCTURE ONLY.

Reference: utils.py:402-436 (extrac_patch2: column-major chop), :449-464 (pred_recostruction), :505-548 (prediction),
utils2.py:312-356 (matrics_AA_recall).  The only third-party arithmetic is skimage.morphology.area_opening(img,
area_threshold, connectivity=1) — scikit-image is not installed here (unpinned in the reference); for the BINARY maps the
reference feeds it, its published definition is: keep the ones whose 4-connected component has at least area_threshold
pixels.  tests/test_amazon_cpu.py cross-checks this restatement against scipy.ndimage.label (an independent connected-
component implementation); with neither skimage nor reference fixtures available the area-opening parity stays
"unpinned at the skimage boundary".
"""
import numpy as np


def extrac_patch2(img, stride, img_type):
    """utils.py:402-436: tiles enumerated with the COLUMN index outermost, remainder dropped."""
    h, w = img.shape[:2]
    nh, nw = int(h / stride), int(w / stride)
    out = []
    for i in range(nw):
        for j in range(nh):
            out.append(img[stride * j:stride * (j + 1), stride * i:stride * (i + 1)])
    return np.asarray(out)


def pred_recostruction(patch_size, pred_labels, image_ref):
    """utils.py:449-464: inverse of extrac_patch2 into a zero float64 image of the covered size."""
    h, w = image_ref.shape
    nh, nw = int(h / patch_size), int(w / patch_size)
    out = np.zeros((nh * patch_size, nw * patch_size))
    count = 0
    for i in range(nw):
        for j in range(nh):
            out[patch_size * j:patch_size * (j + 1), patch_size * i:patch_size * (i + 1)] = pred_labels[count]
            count += 1
    return out


def area_opening_binary(img, area_threshold):
    """4-connected components of ones with fewer than area_threshold pixels are removed (flood fill)."""
    img = np.asarray(img)
    H, W = img.shape
    fg = img > 0
    seen = np.zeros((H, W), bool)
    out = np.zeros((H, W), img.dtype)
    for y0 in range(H):
        for x0 in range(W):
            if not fg[y0, x0] or seen[y0, x0]:
                continue
            comp, stack = [], [(y0, x0)]
            seen[y0, x0] = True
            while stack:
                y, x = stack.pop()
                comp.append((y, x))
                for yy, xx in ((y - 1, x), (y + 1, x), (y, x - 1), (y, x + 1)):
                    if 0 <= yy < H and 0 <= xx < W and fg[yy, xx] and not seen[yy, xx]:
                        seen[yy, xx] = True
                        stack.append((yy, xx))
            if len(comp) >= area_threshold:
                for y, x in comp:
                    out[y, x] = 1
    return out


def consider(img_reconstructed, ref_clip, clipping_mask_, area):
    """utils.py:527-545 given the reconstructed maps: returns (ref_final, pre_final, mask_no_consider)."""
    mask_areas_pred = np.ones_like(img_reconstructed)
    opened = area_opening_binary(img_reconstructed, area)
    area_no_consider = img_reconstructed - opened
    mask_areas_pred[area_no_consider == 1] = 0
    mask_borders = np.ones_like(img_reconstructed)
    mask_borders[ref_clip == 2] = 0
    mask_no_consider = mask_areas_pred * mask_borders
    ref_consider = mask_no_consider * ref_clip
    pred_consider = mask_no_consider * img_reconstructed
    sel = clipping_mask_ * mask_no_consider == 1
    return ref_consider[sel], pred_consider[sel], mask_no_consider


def metrics_aa_recall_one(thr, prob_map, reference, mask_amazon_ts, area):
    """One threshold of utils2.py:312-356: returns (recall, precision, alarm area)."""
    rec = (prob_map >= thr).astype(np.float64)
    _, _, mask_no_consider = consider(rec, reference, (mask_amazon_ts == 1).astype(np.float64), area)
    # utils2.py:332-336: the selection uses mask_amazon_ts alone; masked-out pixels remain as (0, 0) pairs
    ref_final = (mask_no_consider * reference)[mask_amazon_ts == 1]
    pre_final = (mask_no_consider * rec)[mask_amazon_ts == 1]
    tp = np.sum((ref_final == 1) & (pre_final == 1))
    fp = np.sum((ref_final == 0) & (pre_final == 1))
    fn = np.sum((ref_final == 1) & (pre_final == 0))
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.float64(tp) / (tp + fn), np.float64(tp) / (tp + fp), (tp + fp) / len(ref_final)
