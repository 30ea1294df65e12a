"""Amazon deforestation evaluation post-processing: drop-ins for utils.extrac_patch2 (:402-436),
utils.pred_recostruction (:449-464), the evaluation part of utils.prediction (:505-548) and utils2.matrics_AA_recall
(:312-356).  Chopping / pasting are index arithmetic on host arrays like in the reference; the connected-component area
filter, the mask pipeline and the confusion counts run on the GPU (csrc/postproc.cu).  No CPU path."""
from __future__ import annotations

import time

import numpy as np
import torch

from . import _capi
from .inference import metrics_from_confusion


def extrac_patch2(img, stride, img_type=None):
    """Column-major non-overlapping tiles (column index outermost), remainder dropped — utils.py:402-436."""
    img = np.asarray(img)
    h, w = img.shape[:2]
    nh, nw = int(h / stride), int(w / stride)
    v = img[:nh * stride, :nw * stride].reshape((nh, stride, nw, stride) + img.shape[2:])
    v = np.moveaxis(v, 2, 0)                     # (nw, nh, stride, stride, ...)
    return np.ascontiguousarray(v.reshape((nw * nh, stride, stride) + img.shape[2:]))


def pred_recostruction(patch_size, pred_labels, image_ref):
    """Inverse of extrac_patch2 into a float64 image of the covered size — utils.py:449-464 (spelling kept)."""
    h, w = np.asarray(image_ref).shape[:2]
    nh, nw = int(h / patch_size), int(w / patch_size)
    p = np.asarray(pred_labels)[:nh * nw].reshape(nw, nh, patch_size, patch_size)
    return np.ascontiguousarray(np.transpose(p, (1, 2, 0, 3)).reshape(nh * patch_size, nw * patch_size)).astype(np.float64)


class _Post:
    def __init__(self):
        self.lib = _capi.get_lib()
        self.dev = torch.device("cuda", torch.cuda.current_device())

    def u8(self, a):
        return torch.from_numpy(np.ascontiguousarray(a).astype(np.uint8)).to(self.dev)

    def stream(self):
        return torch.cuda.current_stream().cuda_stream


def area_opening(img, area_threshold, connectivity=1):
    """skimage.morphology.area_opening for the binary maps of the reference (utils.py:531): numpy in / numpy out."""
    if connectivity != 1:
        raise ValueError("only 4-connectivity (connectivity=1), the reference's setting, is implemented")
    a = np.asarray(img)
    P = _Post()
    H, W = a.shape
    src, out = P.u8(a > 0), torch.empty((H, W), dtype=torch.uint8, device=P.dev)
    ws = torch.empty(2 * H * W, dtype=torch.int32, device=P.dev)
    P.lib.area_opening_binary(src, out, H, W, area_threshold, ws)(P.stream())
    return out.cpu().numpy().astype(a.dtype)


def consider(img_reconstructed, ref_clip, clipping_mask_, area):
    """utils.py:527-545 on reconstructed maps -> (ref_final, pre_final, confusion[3,3] int64)."""
    P = _Post()
    rec = np.asarray(img_reconstructed)
    H, W = rec.shape
    n = H * W
    pred, ref, clip = P.u8(rec), P.u8(ref_clip), P.u8(np.asarray(clipping_mask_) == 1)
    opened = torch.empty((H, W), dtype=torch.uint8, device=P.dev)
    ws = torch.empty(2 * n, dtype=torch.int32, device=P.dev)
    P.lib.area_opening_binary(pred, opened, H, W, area, ws)(P.stream())
    rc, pc, sel = (torch.empty(n, dtype=torch.uint8, device=P.dev) for _ in range(3))
    cm = torch.zeros(9, dtype=torch.int64, device=P.dev)
    P.lib.amazon_consider(pred, opened, ref, clip, rc, pc, sel, n, cm)(P.stream())
    selh = sel.cpu().numpy().astype(bool)
    dt = rec.dtype if rec.dtype.kind == "f" else np.float64
    return rc.cpu().numpy()[selh].astype(dt), pc.cpu().numpy()[selh].astype(dt), cm.cpu().numpy().reshape(3, 3)


def prediction(model, image_array, image_ref, final_mask, mask_amazon_ts_, patch_size, area):
    """utils.prediction (:505-548): same 7-tuple.  `model.predict` returns [P, ps, ps, n] probabilities."""
    patch_ts = extrac_patch2(image_array, patch_size, 2)
    patches_lb = extrac_patch2(image_ref, patch_size, 1)
    clipping_ref = extrac_patch2(final_mask, patch_size, 1)
    t0 = time.time()
    predictions = model.predict(patch_ts)
    if isinstance(predictions, dict):
        predictions = predictions["seg"]
    probs = predictions[:, :, :, 1]
    p_labels = predictions.argmax(axis=3)
    end_test = time.time() - t0
    ref_reconstructed = pred_recostruction(patch_size, patches_lb, image_ref)
    img_reconstructed = pred_recostruction(patch_size, p_labels, image_ref)
    prob_recontructed = pred_recostruction(patch_size, probs, image_ref)
    ref_clip = pred_recostruction(patch_size, clipping_ref, image_ref)
    clipping_mask_ = pred_recostruction(patch_size, extrac_patch2(mask_amazon_ts_, patch_size, 1), image_ref)
    ref_final, pre_final, _ = consider(img_reconstructed, ref_clip, clipping_mask_, area)
    return ref_final, pre_final, prob_recontructed, ref_reconstructed, ref_clip, clipping_mask_, end_test


def matrics_AA_recall(thresholds, prob_map, reference, mask_amazon_ts, area, verbose=False):
    """utils2.matrics_AA_recall (:312-356): rows of (recall, precision, alarm area) per threshold."""
    out = []
    # utils2.py:335-336 selects with mask_amazon_ts == 1 ONLY: pixels removed by the area filter or the reference borders
    # stay in ref_final / pre_final as (0, 0) pairs, so they count in the alarm-area denominator (TP, FP, FN are the same)
    n_sel = int(np.count_nonzero(np.asarray(mask_amazon_ts) == 1))
    for thr in thresholds:
        rec = (np.asarray(prob_map) >= thr).astype(np.float64)
        _, _, cm3 = consider(rec, reference, np.asarray(mask_amazon_ts) == 1, area)
        cm = cm3[:2, :2]
        if verbose:
            print(thr, "\n", cm, metrics_from_confusion(cm))
        tp, fp, fn = cm[1, 1], cm[0, 1], cm[1, 0]
        with np.errstate(divide="ignore", invalid="ignore"):
            out.append(np.hstack((np.float64(tp) / (tp + fn), np.float64(tp) / (tp + fp), (tp + fp) / max(n_sel, 1))))
    return np.asarray(out)
