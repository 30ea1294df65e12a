"""ctypes binding of libresuneta.so (C-ABI declared in include/resuneta.h).

Every method of :class:`Lib` validates/convert its arguments ONCE and returns a callable
``launch(stream)`` that issues the kernel on the given ``cudaStream_t`` — the execution plans
(graph.py) pre-bind all launches at plan-build time so that a training step is a flat loop of
C calls (and can be captured in a CUDA graph).  Tensors are torch CUDA tensors used purely as
device-memory handles (``data_ptr()``); they can equally come from DLPack
(``torch.from_dlpack``) — the boundary itself only sees raw pointers.

There is NO CPU fallback: if the shared library is missing or the device is not sm_100 the
constructor raises.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libresuneta.so")

RSA_F32, RSA_BF16 = 0, 1
MAX_SEG = 16


class SegC(C.Structure):
    _fields_ = [("src", C.c_void_p), ("C", C.c_int32), ("Hs", C.c_int32), ("Ws", C.c_int32),
                ("mult", C.c_int32), ("shift", C.c_int32), ("off_h", C.c_int32), ("off_w", C.c_int32),
                ("relu_in", C.c_int32), ("aligned", C.c_int32), ("w_off", C.c_int64)]


@dataclass
class Seg:
    """One K-segment of an implicit GEMM (see rsa_seg_t in include/resuneta.h)."""
    src: torch.Tensor
    C: int
    Hs: int
    Ws: int
    mult: int = 1
    shift: int = 0
    off_h: int = 0
    off_w: int = 0
    relu_in: bool = False
    aligned: bool = False
    w_off: int = 0


def dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return RSA_F32
    if t.dtype == torch.bfloat16:
        return RSA_BF16
    raise TypeError(f"unsupported dtype {t.dtype}")


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


_SIGS = {
    "rsa_igemm_fwd": [C.POINTER(SegC), C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                      C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                      C.c_int, C.c_void_p],
    "rsa_igemm_wgrad": [C.POINTER(SegC), C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                        C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p],
    "rsa_bn_stats": [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_void_p],
    "rsa_bn_apply": [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_void_p),
                     C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p, C.c_double,
                     C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_float, C.c_int, C.c_void_p, C.c_void_p],
    "rsa_bn_bwd_reduce": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_double,
                          C.c_float, C.c_void_p, C.c_void_p],
    "rsa_bn_bwd_apply": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_double,
                         C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                         C.c_void_p],
    "rsa_bn_bwd_reduce_multi": [C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_void_p,
                                C.c_double, C.c_float, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int,
                                C.POINTER(C.c_void_p), C.c_void_p],
    "rsa_bn_bwd_apply_multi": [C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_void_p,
                               C.c_double, C.c_float, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int,
                               C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.POINTER(C.c_void_p),
                               C.POINTER(C.c_void_p), C.c_void_p],
    "rsa_bn_meaninv": [C.c_void_p, C.c_double, C.c_float, C.c_void_p, C.c_int, C.c_void_p],
    "rsa_bn_derive_stats": [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_double,
                            C.c_int, C.c_void_p],
    "rsa_bn_update_moving": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p],
    "rsa_maxpool_pyr_fwd": [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                            C.c_void_p, C.c_void_p],
    "rsa_maxpool_pyr_bwd": [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                            C.c_void_p, C.c_void_p, C.c_int, C.c_void_p],
    "rsa_sumpool_pyr": [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                        C.c_void_p, C.c_void_p],
    "rsa_softmax_fwd": [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p],
    "rsa_softmax_bwd": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p],
    "rsa_sigmoid_fwd": [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p],
    "rsa_sigmoid_bwd": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p],
    "rsa_tanimoto_sums": [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_void_p],
    "rsa_tanimoto_finalize": [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_float, C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_void_p],
    "rsa_tanimoto_bwd": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_void_p],
    "rsa_pixel_loss_fwd": [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p,
                           C.c_void_p],
    "rsa_pixel_loss_bwd": [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_float,
                           C.c_void_p, C.c_void_p],
    "rsa_pixel_loss_elem": [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p,
                            C.c_void_p],
    "rsa_seg_metrics": [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p],
    "rsa_adam_step": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_void_p, C.c_float,
                      C.c_float, C.c_float, C.c_float, C.c_void_p],
    "rsa_sgd_step": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_void_p, C.c_float, C.c_float,
                     C.c_void_p],
    "rsa_argmax_confusion": [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                             C.c_void_p],
    "rsa_stem_fwd": [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p,
                     C.c_void_p],
    "rsa_stem_wgrad": [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p],
    "rsa_head_bwd": [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_int,
                     C.c_void_p, C.c_void_p, C.c_void_p],
    "rsa_area_opening_binary": [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p],
    "rsa_amazon_consider": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                            C.c_void_p, C.c_void_p],
    "rsa_label_boundary": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p],
    "rsa_label_distance": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p],
    "rsa_label_hsv": [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p],
    "rsa_head_fwd": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p],
    "rsa_axpy": [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p],
    "rsa_conv_tc2_fwd": [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                         C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, C.c_void_p,
                         C.c_void_p, C.c_int, C.c_int, C.c_void_p],
    "rsa_conv_tc_wgrad": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                          C.c_void_p],
    "rsa_conv_tc3_wgrad": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p],
    "rsa_conv_tc3_fwd": [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_int,
                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                         C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p],
    "rsa_pw_wgrad_tc": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                        C.c_int, C.c_void_p],
    "rsa_bias_grad": [C.c_void_p, C.c_int, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                      C.c_void_p],
    "rsa_pack_weights_tc": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_void_p],
    "rsa_cast": [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int64, C.c_void_p],
}

EXPORTS = sorted(list(_SIGS) + ["rsa_version", "rsa_last_error", "rsa_device_check", "rsa_conv_tc_supported",
                                "rsa_conv_tc2_supported", "rsa_conv_tc3_supported", "rsa_conv_tc3_wgrad_supported",
                                "rsa_label_workspace_bytes"])


def load_cdll(path=LIB_PATH):
    """dlopen the library and declare prototypes.  No device needed (used by the CPU symbol test)."""
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
    dll = C.CDLL(path)
    for name, sig in _SIGS.items():
        fn = getattr(dll, name)
        fn.argtypes = sig
        fn.restype = C.c_int
    dll.rsa_version.restype = C.c_char_p
    dll.rsa_last_error.restype = C.c_char_p
    dll.rsa_device_check.restype = C.c_int
    dll.rsa_conv_tc_supported.argtypes = [C.c_int] * 5
    dll.rsa_conv_tc_supported.restype = C.c_int
    dll.rsa_conv_tc2_supported.argtypes = [C.c_int] * 6
    dll.rsa_conv_tc2_supported.restype = C.c_int
    dll.rsa_conv_tc3_supported.argtypes = [C.c_int] * 4
    dll.rsa_conv_tc3_supported.restype = C.c_int
    dll.rsa_conv_tc3_wgrad_supported.argtypes = [C.c_int] * 5
    dll.rsa_conv_tc3_wgrad_supported.restype = C.c_int
    dll.rsa_label_workspace_bytes.argtypes = [C.c_int] * 4
    dll.rsa_label_workspace_bytes.restype = C.c_int64
    return dll


class Lib:
    """Kernel launcher over the C-ABI.  Methods return ``launch(stream:int)`` callables."""

    is_emulation = False

    def __init__(self, path=LIB_PATH, check_device=True):
        self.dll = load_cdll(path)
        if check_device:
            if not torch.cuda.is_available():
                raise RuntimeError("libresuneta needs a CUDA device (B200); there is no CPU path")
            rc = self.dll.rsa_device_check()
            if rc:
                raise RuntimeError("rsa_device_check: " + self.dll.rsa_last_error().decode())
        self.launches = 0   # kernels launched through this object (bench.py's gpu_launches claim)

    # ------------------------------------------------------------------------------------------
    def _bind(self, name, *args, keep=()):
        fn = getattr(self.dll, name)
        dll = self.dll
        lib = self

        def launch(stream, _fn=fn, _args=args, _keep=keep):
            rc = _fn(*_args, stream)
            lib.launches += 1
            if rc:
                raise RuntimeError(f"{name} failed ({rc}): {dll.rsa_last_error().decode()}")
        launch.kernel = name
        launch.ints = tuple(a for a in args if isinstance(a, int) and not isinstance(a, bool))
        return launch

    @staticmethod
    def _segs(segs):
        assert 1 <= len(segs) <= MAX_SEG, len(segs)
        arr = (SegC * len(segs))()
        dt = None
        for i, s in enumerate(segs):
            d = dtype_code(s.src)
            assert dt is None or dt == d, "all segments must share a dtype"
            dt = d
            arr[i] = SegC(s.src.data_ptr(), s.C, s.Hs, s.Ws, s.mult, s.shift, s.off_h, s.off_w,
                          int(s.relu_in), int(s.aligned), s.w_off)
        return arr, dt

    # -- implicit GEMM -------------------------------------------------------------------------
    def igemm_fwd(self, segs, w, ldw, transB, bias, out, N, Ho, Wo, Co, residual=None, mask=None, stats=None,
                  accumulate=False, relu=False):
        arr, dt = self._segs(segs)
        assert w.dtype == torch.float32 and (bias is None or bias.dtype == torch.float32)
        return self._bind("rsa_igemm_fwd", arr, len(segs), dt, _p(w), ldw, int(transB), _p(bias), _p(out),
                          dtype_code(out), _p(residual), _p(mask), _p(stats), N, Ho, Wo, Co, int(accumulate),
                          int(relu), keep=(segs, w, bias, out, residual, mask, stats))

    def igemm_wgrad(self, segs, dy, dw, ldw, dbias, N, Ho, Wo, Co):
        arr, dt = self._segs(segs)
        assert dw.dtype == torch.float32
        return self._bind("rsa_igemm_wgrad", arr, len(segs), dt, _p(dy), dtype_code(dy), _p(dw), ldw, _p(dbias),
                          N, Ho, Wo, Co, keep=(segs, dy, dw, dbias))

    # -- batch norm -----------------------------------------------------------------------------
    def bn_stats(self, x, M, C_, stats):
        return self._bind("rsa_bn_stats", _p(x), dtype_code(x), M, C_, _p(stats), keep=(x, stats))

    @staticmethod
    def _ptr_array(ts):
        if ts is None:
            return None
        arr = (C.c_void_p * len(ts))()
        for i, t in enumerate(ts):
            arr[i] = t.data_ptr()
        return arr

    def bn_apply(self, x, M, C_, outs, gammas, betas, stats, count, mmeans, mvars, eps, relu, meaninv=None):
        return self._bind("rsa_bn_apply", _p(x), dtype_code(x), M, C_, len(outs), self._ptr_array(outs),
                          self._ptr_array(gammas), self._ptr_array(betas), _p(stats), float(count),
                          self._ptr_array(mmeans), self._ptr_array(mvars), float(eps), int(relu), _p(meaninv),
                          keep=(x, outs, gammas, betas, stats, mmeans, mvars, meaninv))

    def bn_bwd_reduce(self, dy, x, act, M, C_, stats, count, eps, red):
        return self._bind("rsa_bn_bwd_reduce", _p(dy), _p(x), _p(act), dtype_code(x), M, C_, _p(stats),
                          float(count), float(eps), _p(red), keep=(dy, x, act, stats, red))

    def bn_bwd_apply(self, dy, x, act, M, C_, stats, count, eps, gamma, red, dx, accumulate, dgamma, dbeta):
        return self._bind("rsa_bn_bwd_apply", _p(dy), _p(x), _p(act), dtype_code(x), M, C_, _p(stats),
                          float(count), float(eps), _p(gamma), _p(red), _p(dx), int(accumulate), _p(dgamma),
                          _p(dbeta), keep=(dy, x, act, stats, gamma, red, dx, dgamma, dbeta))

    def bn_bwd_reduce_multi(self, dys, x, M, C_, stats, count, eps, gammas, betas, relu, reds):
        return self._bind("rsa_bn_bwd_reduce_multi", self._ptr_array(dys), _p(x), dtype_code(x), M, C_, len(dys),
                          _p(stats), float(count), float(eps), self._ptr_array(gammas), self._ptr_array(betas),
                          int(relu), self._ptr_array(reds), keep=(dys, x, stats, gammas, betas, reds))

    def bn_bwd_apply_multi(self, dys, x, M, C_, stats, count, eps, gammas, betas, relu, reds, dx, accumulate, dgammas,
                           dbetas):
        return self._bind("rsa_bn_bwd_apply_multi", self._ptr_array(dys), _p(x), dtype_code(x), M, C_, len(dys),
                          _p(stats), float(count), float(eps), self._ptr_array(gammas), self._ptr_array(betas),
                          int(relu), self._ptr_array(reds), _p(dx), int(accumulate), self._ptr_array(dgammas),
                          self._ptr_array(dbetas), keep=(dys, x, stats, gammas, betas, reds, dx, dgammas, dbetas))

    def bn_meaninv(self, stats, count, eps, out, C_):
        return self._bind("rsa_bn_meaninv", _p(stats), float(count), float(eps), _p(out), C_, keep=(stats, out))

    def bn_derive_stats(self, src_stats, count, gamma, beta, eps, dst_stats, dst_count, C_):
        return self._bind("rsa_bn_derive_stats", _p(src_stats), float(count), _p(gamma), _p(beta), float(eps),
                          _p(dst_stats), float(dst_count), C_, keep=(src_stats, gamma, beta, dst_stats))

    def bn_update_moving(self, stats_base, param_base, table, counts, nlayers, momentum):
        return self._bind("rsa_bn_update_moving", _p(stats_base), _p(param_base), _p(table), _p(counts), nlayers,
                          float(momentum), keep=(stats_base, param_base, table, counts))

    # -- pooling pyramid ------------------------------------------------------------------------
    def maxpool_pyr_fwd(self, x, N, H, W, C_, p2, p4, p8):
        return self._bind("rsa_maxpool_pyr_fwd", _p(x), dtype_code(x), N, H, W, C_, _p(p2), _p(p4), _p(p8),
                          keep=(x, p2, p4, p8))

    def maxpool_pyr_bwd(self, x, N, H, W, C_, dp2, dp4, dp8, dx, accumulate):
        return self._bind("rsa_maxpool_pyr_bwd", _p(x), dtype_code(x), N, H, W, C_, _p(dp2), _p(dp4), _p(dp8),
                          _p(dx), int(accumulate), keep=(x, dp2, dp4, dp8, dx))

    def sumpool_pyr(self, x, N, H, W, C_, s2, s4, s8):
        return self._bind("rsa_sumpool_pyr", _p(x), dtype_code(x), N, H, W, C_, _p(s2), _p(s4), _p(s8),
                          keep=(x, s2, s4, s8))

    # -- heads / losses -------------------------------------------------------------------------
    def softmax_fwd(self, z, p, M, C_):
        return self._bind("rsa_softmax_fwd", _p(z), _p(p), M, C_, keep=(z, p))

    def softmax_bwd(self, p, dp, dz, M, C_):
        return self._bind("rsa_softmax_bwd", _p(p), _p(dp), _p(dz), M, C_, keep=(p, dp, dz))

    def sigmoid_fwd(self, z, p, n):
        return self._bind("rsa_sigmoid_fwd", _p(z), _p(p), n, keep=(z, p))

    def sigmoid_bwd(self, p, dp, dz, n):
        return self._bind("rsa_sigmoid_bwd", _p(p), _p(dp), _p(dz), n, keep=(p, dp, dz))

    def tanimoto_sums(self, pred, label, B, HW, C_, sums):
        return self._bind("rsa_tanimoto_sums", _p(pred), _p(label), B, HW, C_, _p(sums), keep=(pred, label, sums))

    def tanimoto_finalize(self, sums, B, HW, C_, scale, loss_b, loss_mean, coef):
        return self._bind("rsa_tanimoto_finalize", _p(sums), B, HW, C_, float(scale), _p(loss_b), _p(loss_mean),
                          _p(coef), keep=(sums, loss_b, loss_mean, coef))

    def tanimoto_bwd(self, pred, label, coef, B, HW, C_, dpred):
        return self._bind("rsa_tanimoto_bwd", _p(pred), _p(label), _p(coef), B, HW, C_, _p(dpred),
                          keep=(pred, label, coef, dpred))

    def pixel_loss_fwd(self, kind, pred, label, weights, M, C_, loss_sum):
        return self._bind("rsa_pixel_loss_fwd", kind, _p(pred), _p(label), _p(weights), M, C_, _p(loss_sum),
                          keep=(pred, label, weights, loss_sum))

    def pixel_loss_bwd(self, kind, pred, label, weights, M, C_, scale, dpred):
        return self._bind("rsa_pixel_loss_bwd", kind, _p(pred), _p(label), _p(weights), M, C_, float(scale),
                          _p(dpred), keep=(pred, label, weights, dpred))

    def pixel_loss_elem(self, kind, pred, label, weights, M, C_, out):
        return self._bind("rsa_pixel_loss_elem", kind, _p(pred), _p(label), _p(weights), M, C_, _p(out),
                          keep=(pred, label, weights, out))

    def seg_metrics(self, pred, label, M, C_, out):
        return self._bind("rsa_seg_metrics", _p(pred), _p(label), M, C_, _p(out), keep=(pred, label, out))

    def argmax_confusion(self, prob, M, C_, pred_label, true_label, K, cm):
        return self._bind("rsa_argmax_confusion", _p(prob), M, C_, _p(pred_label), _p(true_label), K, _p(cm),
                          keep=(prob, pred_label, true_label, cm))

    # -- tensor-core convolution (bf16) ------------------------------------------------------------
    def conv_tc_supported(self, N, H, W, Cin, Cout):
        return bool(self.dll.rsa_conv_tc_supported(N, H, W, Cin, Cout))

    def conv_tc2_supported(self, N, H, W, C0, C1, Cout):
        return bool(self.dll.rsa_conv_tc2_supported(N, H, W, C0, C1, Cout))

    def conv_tc2_fwd(self, x0, x1, wt, CoutP, bias, out, N, H, W, Cout, taps=1, dil=1, in_stride=1, ups=(),
                     residual=None, mask=None, stats=None, accumulate=False, relu=False, k_base=0, k_total=0,
                     out_stride=1, bnr_x=None, bnr_coef=None):
        """ups: sequence of (q tensor, shift).  x1 may be None.  out bf16 or fp32."""
        assert x0.dtype == torch.bfloat16 and wt.dtype == torch.bfloat16
        C0 = x0.shape[-1]
        C1 = x1.shape[-1] if x1 is not None else 0
        nup = len(ups)
        up_ptrs = (C.c_void_p * max(nup, 1))()
        up_sh = (C.c_int * max(nup, 1))()
        for i, (q, sft) in enumerate(ups):
            assert q.dtype == torch.bfloat16
            up_ptrs[i] = q.data_ptr()
            up_sh[i] = sft
        return self._bind("rsa_conv_tc2_fwd", _p(x0), C0, _p(x1), C1, _p(wt), CoutP, _p(bias), _p(out),
                          int(out.dtype == torch.float32), _p(residual), _p(mask), _p(stats), N, H, W, Cout, taps, dil,
                          in_stride, nup, up_ptrs, up_sh, k_base, k_total, out_stride, _p(bnr_x), _p(bnr_coef),
                          int(accumulate), int(relu),
                          keep=(x0, x1, wt, bias, out, residual, mask, stats, ups, up_ptrs, up_sh, bnr_x, bnr_coef))

    def conv_tc3_supported(self, N, H, W, C_):
        return bool(self.dll.rsa_conv_tc3_supported(N, H, W, C_))

    def conv_tc3_fwd(self, xs, wts, biases, dils, out, N, H, W, C_, residual=None, mask=None, stats=None,
                     accumulate=False, relu=False, bnr=None):
        """Thin-layer 3x3 convolution; len(xs) branches accumulate into one tile (conv_tc3.cu).
        bnr = (x, fwd_stats, count, eps, gamma, beta, relu): fuse that BatchNorm's backward reductions (stats = output)."""
        nbr = len(xs)
        assert nbr == len(wts) == len(dils) and 1 <= nbr <= 4
        assert all(x.dtype == torch.bfloat16 for x in xs) and out.dtype == torch.bfloat16
        xa, wa = self._ptr_array(xs), self._ptr_array(wts)
        ba = (C.c_void_p * nbr)()
        for i in range(nbr):
            ba[i] = biases[i].data_ptr() if biases is not None and biases[i] is not None else None
        da = (C.c_int * nbr)(*[int(d) for d in dils])
        bx, bst, bcnt, beps, bg, bb, brelu = bnr if bnr is not None else (None, None, 0.0, 0.0, None, None, 0)
        return self._bind("rsa_conv_tc3_fwd", xa, wa, ba, da, nbr, _p(out), _p(residual), _p(mask), _p(stats), N, H, W,
                          C_, int(accumulate), int(relu), _p(bx), _p(bst), float(bcnt), float(beps), _p(bg), _p(bb),
                          int(brelu), keep=(xs, wts, biases, out, residual, mask, stats, xa, wa, ba, da, bnr))

    def conv_tc3_wgrad_supported(self, N, H, W, C_, dil):
        return bool(self.dll.rsa_conv_tc3_wgrad_supported(N, H, W, C_, dil))

    def conv_tc3_wgrad(self, x, dy, dw, N, H, W, C_, dil):
        assert x.dtype == torch.bfloat16 and dy.dtype == torch.bfloat16 and dw.dtype == torch.float32
        return self._bind("rsa_conv_tc3_wgrad", _p(x), _p(dy), _p(dw), N, H, W, C_, dil, keep=(x, dy, dw))

    def conv_tc_wgrad(self, x, dy, dw, N, H, W, Cin, Cout, dil):
        assert x.dtype == torch.bfloat16 and dy.dtype == torch.bfloat16 and dw.dtype == torch.float32
        return self._bind("rsa_conv_tc_wgrad", _p(x), _p(dy), _p(dw), N, H, W, Cin, Cout, dil, keep=(x, dy, dw))

    def pw_wgrad_tc(self, x, dz, dw, ldw, N, H, W, Cin, Cout, in_stride=1):
        assert x.dtype == torch.bfloat16 and dz.dtype == torch.bfloat16 and dw.dtype == torch.float32
        return self._bind("rsa_pw_wgrad_tc", _p(x), _p(dz), _p(dw), ldw, N, H, W, Cin, Cout, in_stride,
                          keep=(x, dz, dw))

    def bias_grad(self, dy, M, C_, dbs):
        d = list(dbs) + [None] * (4 - len(dbs))
        return self._bind("rsa_bias_grad", _p(dy), dtype_code(dy), M, C_, _p(d[0]), _p(d[1]), _p(d[2]), _p(d[3]),
                          keep=(dy, dbs))

    def pack_weights_tc(self, params, shadow, table, nlayers, max_elems):
        return self._bind("rsa_pack_weights_tc", _p(params), _p(shadow), _p(table), nlayers, max_elems,
                          keep=(params, shadow, table))

    # -- thin 1x1 convolutions ---------------------------------------------------------------------
    def stem_fwd(self, x, w, b, out, M, n, stats):
        return self._bind("rsa_stem_fwd", _p(x), dtype_code(x), _p(w), _p(b), _p(out), dtype_code(out), M, n, _p(stats),
                          keep=(x, w, b, out, stats))

    def stem_wgrad(self, x, dy, M, n, dw, db):
        return self._bind("rsa_stem_wgrad", _p(x), _p(dy), dtype_code(x), M, n, _p(dw), _p(db), keep=(x, dy, dw, db))

    # -- Amazon evaluation post-processing (postproc.cu) ----------------------------------------------
    def area_opening_binary(self, img, out, H, W, area_threshold, ws):
        assert img.dtype == torch.uint8 and out.dtype == torch.uint8 and ws.dtype == torch.int32 and ws.numel() >= 2 * H * W
        return self._bind("rsa_area_opening_binary", _p(img), _p(out), H, W, int(area_threshold), _p(ws), keep=(img, out, ws))

    def amazon_consider(self, pred, opened, ref_clip, clip_mask, ref_consider, pred_consider, selected, n, cm):
        assert cm.dtype == torch.int64 and cm.numel() >= 9
        return self._bind("rsa_amazon_consider", _p(pred), _p(opened), _p(ref_clip), _p(clip_mask), _p(ref_consider),
                          _p(pred_consider), _p(selected), n, _p(cm),
                          keep=(pred, opened, ref_clip, clip_mask, ref_consider, pred_consider, selected, cm))

    # -- multitask label generation (labels.cu) ------------------------------------------------------
    def label_workspace_bytes(self, N, H, W, C_):
        return int(self.dll.rsa_label_workspace_bytes(N, H, W, C_))

    def label_boundary(self, label, out, ws, N, H, W, C_):
        assert label.dtype == torch.float32 and out.dtype == torch.float32 and ws.dtype == torch.uint8
        return self._bind("rsa_label_boundary", _p(label), _p(out), _p(ws), N, H, W, C_, keep=(label, out, ws))

    def label_distance(self, label, out, ws, N, H, W, C_):
        assert label.dtype == torch.float32 and out.dtype == torch.float32 and ws.dtype == torch.uint8
        return self._bind("rsa_label_distance", _p(label), _p(out), _p(ws), N, H, W, C_, keep=(label, out, ws))

    def label_hsv(self, rgb, out, npix):
        assert rgb.dtype == torch.uint8 and out.dtype == torch.float32
        return self._bind("rsa_label_hsv", _p(rgb), _p(out), npix, keep=(rgb, out))

    def head_fwd(self, h, w, b, z, M, n):
        assert h.dtype == torch.bfloat16 and z.dtype == torch.float32
        return self._bind("rsa_head_fwd", _p(h), _p(w), _p(b), _p(z), M, n, keep=(h, w, b, z))

    def head_bwd(self, h, dz, w, M, n, dh, accumulate, relu_mask, dw, db):
        assert dz.dtype == torch.float32
        return self._bind("rsa_head_bwd", _p(h), dtype_code(h), _p(dz), _p(w), M, n, _p(dh), int(accumulate),
                          int(relu_mask), _p(dw), _p(db), keep=(h, dz, w, dh, dw, db))

    # -- optimizers / misc ----------------------------------------------------------------------
    def adam_step(self, param, grad, m, v, n, lr_dev, b1, b2, eps, grad_scale):
        """lr_dev: 1-element fp32 device tensor holding keras' bias-corrected lr_t for this step."""
        return self._bind("rsa_adam_step", _p(param), _p(grad), _p(m), _p(v), n, 0.0, _p(lr_dev), float(b1),
                          float(b2), float(eps), float(grad_scale), keep=(param, grad, m, v, lr_dev))

    def sgd_step(self, param, grad, vel, n, lr_dev, momentum, grad_scale):
        return self._bind("rsa_sgd_step", _p(param), _p(grad), _p(vel), n, 0.0, _p(lr_dev), float(momentum),
                          float(grad_scale), keep=(param, grad, vel, lr_dev))

    def axpy(self, dst, src, n, accumulate):
        return self._bind("rsa_axpy", _p(dst), _p(src), dtype_code(dst), n, int(accumulate), keep=(dst, src))

    def cast(self, src, dst, n):
        return self._bind("rsa_cast", _p(src), dtype_code(src), _p(dst), dtype_code(dst), n, keep=(src, dst))


_LIB = None


def get_lib():
    """Process-wide launcher; raises if the CUDA extension or device is missing."""
    global _LIB
    if _LIB is None:
        _LIB = Lib()
    return _LIB


def set_lib(lib):
    """Test hook: install another launcher object (the CPU test-suite installs an emulation that
    exercises the host-side graph logic; product code never does this)."""
    global _LIB
    _LIB = lib
