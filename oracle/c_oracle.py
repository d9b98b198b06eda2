"""ctypes wrapper of the C oracle (oracle/c_oracle.c) — TEST INFRASTRUCTURE ONLY. Same call
signatures as the NumPy oracle (oracle/np_oracle.py) for S3, S4, S5."""
import ctypes

import numpy as np

from . import build_oracle

_lib = None


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build_oracle.build())
        f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
        i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
        lib.oracle_correlation.restype = ctypes.c_int
        lib.oracle_correlation.argtypes = [f32p, f32p] + [ctypes.c_int] * 9 + [f32p]
        lib.oracle_correlation_grad.restype = ctypes.c_int
        lib.oracle_correlation_grad.argtypes = [f32p, f32p, f32p] + [ctypes.c_int] * 9 + [f32p, f32p]
        lib.oracle_crop_and_resize.restype = ctypes.c_int
        lib.oracle_crop_and_resize.argtypes = [f32p] + [ctypes.c_int] * 4 + [f32p, i32p, ctypes.c_int,
                                                                          ctypes.c_int, ctypes.c_int,
                                                                          ctypes.c_float, f32p]
        lib.oracle_nms.restype = ctypes.c_int
        lib.oracle_nms.argtypes = [f32p, f32p, ctypes.c_int, ctypes.c_int, ctypes.c_float, i32p]
        _lib = lib
    return _lib


def correlation(input_a, input_b, kernel_size=1, max_displacement=20, stride_1=1, stride_2=2,
                padding=20):
    from .np_oracle import correlation_out_shape
    a = np.ascontiguousarray(input_a, dtype=np.float32)
    b = np.ascontiguousarray(input_b, dtype=np.float32)
    N, H, W, C = a.shape
    oh, ow, oc = correlation_out_shape(H, W, kernel_size, max_displacement, stride_1, stride_2, padding)
    out = np.empty((N, oh, ow, oc), dtype=np.float32)
    rc = _load().oracle_correlation(a, b, N, H, W, C, kernel_size, max_displacement, stride_1,
                                    stride_2, padding, out)
    assert rc == 0, rc
    return out


def correlation_grad(gradients, input_a, input_b, kernel_size=1, max_displacement=20, stride_1=1,
                     stride_2=2, padding=20):
    """(backprops_a, backprops_b) of the reference's CorrelationGrad op
    (avod/core/corr_layers/correlation.py:30-48)."""
    from .np_oracle import correlation_out_shape
    a = np.ascontiguousarray(input_a, dtype=np.float32)
    b = np.ascontiguousarray(input_b, dtype=np.float32)
    g = np.ascontiguousarray(gradients, dtype=np.float32)
    N, H, W, C = a.shape
    assert g.shape == (N,) + correlation_out_shape(H, W, kernel_size, max_displacement, stride_1,
                                                    stride_2, padding)
    ga, gb = np.empty_like(a), np.empty_like(b)
    rc = _load().oracle_correlation_grad(g, a, b, N, H, W, C, kernel_size, max_displacement,
                                         stride_1, stride_2, padding, ga, gb)
    assert rc == 0, rc
    return ga, gb


def crop_and_resize(image, boxes, box_ind, crop_size, extrapolation_value=0.0):
    image = np.ascontiguousarray(image, dtype=np.float32)
    boxes = np.ascontiguousarray(boxes, dtype=np.float32).reshape(-1, 4)
    box_ind = np.ascontiguousarray(box_ind, dtype=np.int32)
    B, H, W, C = image.shape
    n = boxes.shape[0]
    out = np.zeros((n, int(crop_size[0]), int(crop_size[1]), C), dtype=np.float32)
    if n:
        _load().oracle_crop_and_resize(image, B, H, W, C, boxes, box_ind, n, int(crop_size[0]),
                                       int(crop_size[1]), float(extrapolation_value), out)
    return out


def non_max_suppression(boxes, scores, max_output_size, iou_threshold=0.5):
    boxes = np.ascontiguousarray(boxes, dtype=np.float32).reshape(-1, 4)
    scores = np.ascontiguousarray(scores, dtype=np.float32).reshape(-1)
    n = boxes.shape[0]
    sel = np.zeros(max(min(int(max_output_size), n), 1), dtype=np.int32)
    k = _load().oracle_nms(boxes, scores, n, int(max_output_size), float(iou_threshold), sel)
    assert k >= 0
    return sel[:k].copy()
