"""Multitask label generation on the GPU: drop-ins for the reference's
multitasking_utils.get_boundary_label / get_distance_label (multitasking_utils.py:6-34) and the HSV colour target of
preprocess_save_patches_ISPRS.py:89-94,224-228, plus a batch entry point that derives all three targets from the one-hot
segmentation batch already in HBM (the reference stores them as four extra .npy files per patch).

No CPU path: the kernels live in libresuneta.so (csrc/labels.cu)."""
from __future__ import annotations

import numpy as np
import torch

from . import _capi


class LabelGenerator:
    """Pre-allocates the workspace for a fixed [N, H, W, C] batch shape; methods take / return device tensors."""

    def __init__(self, N, H, W, C, lib=None, device=None):
        self.lib = lib or _capi.get_lib()
        self.shape = (int(N), int(H), int(W), int(C))
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.ws = torch.empty(self.lib.label_workspace_bytes(*self.shape), dtype=torch.uint8, device=self.device)

    def _stream(self):
        return torch.cuda.current_stream().cuda_stream

    def boundary(self, onehot, out=None):
        N, H, W, C = self.shape
        assert tuple(onehot.shape) == self.shape and onehot.dtype == torch.float32 and onehot.is_contiguous()
        out = torch.empty_like(onehot) if out is None else out
        self.lib.label_boundary(onehot, out, self.ws, N, H, W, C)(self._stream())
        return out

    def distance(self, onehot, out=None):
        N, H, W, C = self.shape
        assert tuple(onehot.shape) == self.shape and onehot.dtype == torch.float32 and onehot.is_contiguous()
        out = torch.empty_like(onehot) if out is None else out
        self.lib.label_distance(onehot, out, self.ws, N, H, W, C)(self._stream())
        return out

    def color(self, rgb_u8, out=None):
        assert rgb_u8.dtype == torch.uint8 and rgb_u8.shape[-1] == 3 and rgb_u8.is_contiguous()
        out = torch.empty(rgb_u8.shape, dtype=torch.float32, device=rgb_u8.device) if out is None else out
        self.lib.label_hsv(rgb_u8, out, rgb_u8.numel() // 3)(self._stream())
        return out

    def multitask_targets(self, onehot, rgb_u8=None):
        """{'seg', 'bound', 'dist'[, 'color']} for Model.train_on_batch (train_ISPRS.py:141-146)."""
        y = {"seg": onehot, "bound": self.boundary(onehot), "dist": self.distance(onehot)}
        if rgb_u8 is not None:
            y["color"] = self.color(rgb_u8)
        return y


_GEN = {}


def _gen(shape):
    g = _GEN.get(shape)
    if g is None:
        g = _GEN[shape] = LabelGenerator(*shape)
    return g


def get_boundary_label(label, kernel_size=(3, 3)):
    """Drop-in for multitasking_utils.get_boundary_label (:6-22): numpy [H, W, C] one-hot in, float32 [H, W, C] out."""
    if tuple(kernel_size) != (3, 3):
        raise ValueError("only the reference's 3x3 cross is implemented")
    a = np.ascontiguousarray(label, dtype=np.float32)
    g = _gen((1,) + a.shape)
    return g.boundary(torch.from_numpy(a).to(g.device).unsqueeze(0).contiguous())[0].cpu().numpy()


def get_distance_label(label):
    """Drop-in for multitasking_utils.get_distance_label (:25-34)."""
    a = np.ascontiguousarray(label, dtype=np.float32)
    g = _gen((1,) + a.shape)
    return g.distance(torch.from_numpy(a).to(g.device).unsqueeze(0).contiguous())[0].cpu().numpy()


def get_color_label(img_u8):
    """cv2.cvtColor(img, COLOR_RGB2HSV).astype(float32) / [179, 255, 255] (preprocess_save_patches_ISPRS.py:224-228)."""
    a = np.ascontiguousarray(img_u8, dtype=np.uint8)
    g = _gen((1, a.shape[0], a.shape[1], 1))
    return g.color(torch.from_numpy(a).to(g.device))[...].cpu().numpy()
