"""ctypes wrapper of oracle/_ref/libcorr_ref.so — the reference's OWN correlation CUDA kernels
(avod/core/ops/correlation/*.cu.cc compiled unmodified, see oracle/build_oracle.py) behind the
host logic of oracle/ref_corr_driver.cu. TEST INFRASTRUCTURE ONLY; needs a GPU to run.

`available()` is False when the library was never built (no reference checkout at build time)."""
import ctypes
import os

import numpy as np

from . import build_oracle

_lib = None


def available():
    return os.path.exists(build_oracle.CORR_REF_LIB)


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build_oracle.CORR_REF_LIB)
        f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
        i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
        lib.ref_correlation_out_shape.restype = ctypes.c_int
        lib.ref_correlation_out_shape.argtypes = [ctypes.c_int] * 7 + [i32p]
        lib.ref_correlation.restype = ctypes.c_int
        lib.ref_correlation.argtypes = [f32p, f32p] + [ctypes.c_int] * 9 + [f32p, f32p, ctypes.c_int]
        lib.ref_correlation_grad.restype = ctypes.c_int
        lib.ref_correlation_grad.argtypes = [f32p, f32p, f32p] + [ctypes.c_int] * 9 + [f32p, f32p, f32p, ctypes.c_int]
        _lib = lib
    return _lib


def out_shape(H, W, kernel_size, max_displacement, stride_1, stride_2, padding):
    hwc = np.zeros(3, dtype=np.int32)
    rc = _load().ref_correlation_out_shape(H, W, kernel_size, max_displacement, stride_1, stride_2, padding, hwc)
    if rc:
        raise ValueError("reference op rejects these attributes (code %d)" % rc)
    return tuple(int(v) for v in hwc)


def correlation(input_a, input_b, kernel_size=1, max_displacement=20, stride_1=1, stride_2=2, padding=20,
                reps=1, return_ms=False):
    """PadData x2 + CorrelateData on the current CUDA device. ms = (pads, correlate) per run."""
    a = np.ascontiguousarray(input_a, dtype=np.float32)
    b = np.ascontiguousarray(input_b, dtype=np.float32)
    N, H, W, C = a.shape
    oh, ow, oc = out_shape(H, W, kernel_size, max_displacement, stride_1, stride_2, padding)
    out = np.empty((N, oh, ow, oc), dtype=np.float32)
    ms = np.zeros(2, dtype=np.float32)
    rc = _load().ref_correlation(a, b, N, H, W, C, kernel_size, max_displacement, stride_1, stride_2, padding,
                                 out, ms, reps)
    if rc:
        raise RuntimeError("ref_correlation failed (code %d)" % rc)
    return (out, ms) if return_ms else out


def correlation_grad(gradients, input_a, input_b, kernel_size=1, max_displacement=20, stride_1=1, stride_2=2,
                     padding=20, reps=1, return_ms=False):
    """PadData x2 + CorrelateDataBackward0 / Backward1. ms = (pads, grad A, grad B) per run."""
    a = np.ascontiguousarray(input_a, dtype=np.float32)
    b = np.ascontiguousarray(input_b, dtype=np.float32)
    g = np.ascontiguousarray(gradients, dtype=np.float32)
    N, H, W, C = a.shape
    assert g.shape == (N,) + out_shape(H, W, kernel_size, max_displacement, stride_1, stride_2, padding)
    ga, gb = np.empty_like(a), np.empty_like(b)
    ms = np.zeros(3, dtype=np.float32)
    rc = _load().ref_correlation_grad(g, a, b, N, H, W, C, kernel_size, max_displacement, stride_1, stride_2,
                                      padding, ga, gb, ms, reps)
    if rc:
        raise RuntimeError("ref_correlation_grad failed (code %d)" % rc)
    return (ga, gb, ms) if return_ms else (ga, gb)
