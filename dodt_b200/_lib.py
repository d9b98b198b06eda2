"""ctypes binding of libdodt_fe.so (include/dodt_fe.h). There is no fallback: if the library is
missing the import of any compute entry point raises, and on a host without a CUDA device the
entry points return DODT_ECUDA which is raised as RuntimeError."""
import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_int64,
                    c_size_t, c_void_p)

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdodt_fe.so")

DODT_OK, DODT_EINVAL, DODT_ESHAPE, DODT_ECAPACITY, DODT_ECUDA, DODT_EALIGN = 0, -1, -2, -3, -4, -5
DODT_F32, DODT_F64 = 0, 1
MAX_SLICES = 15
MAX_DENSITY_LUT = 64
NMS_WINDOW = 1536
NMS_CHUNK_WINDOWS = 2
CORR_STREAM_MAX_PAIRS = 8
BEV_STATS_LEN = 24
STAT_DENSITY, STAT_OCC, STAT_TOUCHED, STAT_OVERFLOW, STAT_OOB = 16, 17, 18, 19, 20


class BevParams(Structure):
    """struct dodt_bev_params (include/dodt_fe.h)."""
    _fields_ = [
        ("plane", c_double * 4),
        ("extents", c_double * 6),
        ("voxel_size", c_double),
        ("height_lo", c_double),
        ("height_hi", c_double),
        ("num_slices", c_int32),
        ("filter_mode", c_int32),
        ("occ_lo", c_double),
        ("occ_hi", c_double),
        ("density_lut_len", c_int32),
        ("reserved", c_int32),
        ("density_lut", c_double * MAX_DENSITY_LUT),
    ]


class AnchorGrid(Structure):
    """struct dodt_anchor_grid."""
    _fields_ = [("extents", c_double * 6), ("stride", c_double * 2), ("plane", c_double * 4),
                ("sizes", c_double * 48), ("n_sizes", c_int32), ("reserved", c_int32)]


class GatherSpec(Structure):
    """struct dodt_gather_spec."""
    _fields_ = [("src", c_void_p), ("dst", c_void_p), ("width", c_int32), ("reserved", c_int32)]


class CropSpec(Structure):
    """struct dodt_crop_spec."""
    _fields_ = [("image", c_void_p), ("boxes", c_void_p), ("crops", c_void_p), ("height", c_int32),
                ("width", c_int32), ("channels", c_int32), ("reserved", c_int32)]


# name -> (restype, argtypes); every symbol include/dodt_fe.h declares
SIGNATURES = {
    "dodt_strerror": (c_char_p, [c_int]),
    "dodt_last_cuda_error": (c_char_p, []),
    "dodt_version": (c_int, []),
    "dodt_launch_count": (c_int64, []),
    "dodt_bev_grid": (c_int, [POINTER(c_double), c_double, POINTER(c_int32)]),
    "dodt_bev_workspace_bytes": (c_size_t, [c_int64, c_int32, c_int32, c_int32]),
    "dodt_lidar_workspace_bytes": (c_size_t, [c_int64]),
    "dodt_lidar_to_camera": (c_int, [c_void_p, c_int64, POINTER(c_double), POINTER(c_double), c_int32,
                                     c_int32, c_void_p, c_int32, c_int64, c_void_p, c_void_p, c_size_t,
                                     c_void_p]),
    "dodt_lidar_to_camera_aligned": (c_int, [c_void_p, c_int64, POINTER(c_double), POINTER(c_double), c_void_p,
                                             POINTER(c_double), POINTER(c_double), c_int32, c_int32, c_void_p,
                                             c_int32, c_int64, c_void_p, c_void_p, c_size_t, c_void_p]),
    "dodt_bev_slices": (c_int, [c_void_p, c_int32, c_int64, c_void_p, c_int64, POINTER(BevParams), c_void_p,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                c_void_p]),
    "dodt_integral_workspace_bytes": (c_size_t, [c_int32, c_int32]),
    "dodt_integral_image_2d": (c_int, [c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_size_t,
                                       c_void_p]),
    "dodt_map_to_index": (c_int, [c_void_p, c_int32, c_int64, c_double, c_int32, c_int32, c_int32,
                                  c_int32, c_void_p, c_void_p]),
    "dodt_anchor_filter_2d": (c_int, [c_void_p, c_int32, c_int64, c_void_p, c_int32, c_int32,
                                      c_int32, c_int32, c_double, c_double, c_void_p, c_void_p,
                                      c_void_p]),
    "dodt_compact_workspace_bytes": (c_size_t, [c_int64]),
    "dodt_compact_mask": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_size_t,
                                  c_void_p]),
    "dodt_gather_rows": (c_int, [c_void_p, c_int32, c_void_p, c_void_p, c_int64, c_void_p,
                                 c_void_p]),
    "dodt_gather_rows_multi": (c_int, [POINTER(GatherSpec), c_int32, c_void_p, c_void_p, c_int64,
                                       c_void_p]),
    "dodt_grid_anchor_shape": (c_int, [POINTER(c_double), POINTER(c_double), c_int32, POINTER(c_int32)]),
    "dodt_grid_anchors": (c_int, [POINTER(c_double), POINTER(c_double), c_int32, POINTER(c_double),
                                  POINTER(c_double), c_void_p, c_void_p]),
    "dodt_project_to_bev": (c_int, [c_void_p, c_int32, c_int64, POINTER(c_double), c_int32, c_void_p,
                                    c_void_p, c_void_p]),
    "dodt_project_to_image_space": (c_int, [c_void_p, c_int32, c_int64, POINTER(c_double), c_int32,
                                            c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "dodt_offset_to_anchor": (c_int, [c_void_p, c_int32, c_void_p, c_int32, c_int64, c_void_p, c_void_p]),
    "dodt_rpn_decode": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, POINTER(c_double),
                                POINTER(c_double), c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "dodt_emit_detections": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_int32,
                                     c_void_p]),
    "dodt_crop_and_resize_multi": (c_int, [POINTER(CropSpec), c_int32, c_int32, c_void_p, c_int64,
                                           c_void_p, c_int32, c_int32, c_float, c_void_p]),
    "dodt_crop_and_resize": (c_int, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p,
                                     c_void_p, c_int64, c_void_p, c_int32, c_int32, c_float,
                                     c_void_p, c_void_p]),
    "dodt_correlation_out_shape": (c_int, [c_int32, c_int32, c_int32, c_int32, c_int32, c_int32,
                                           c_int32, POINTER(c_int32)]),
    "dodt_correlation": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32,
                                 c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    "dodt_correlation_shared": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32,
                                        c_int32, c_int32, c_int32, c_int32, c_void_p, c_int32,
                                        c_void_p]),
    "dodt_correlation_stream": (c_int, [POINTER(c_void_p), c_int32, POINTER(c_void_p), c_int32, c_int32,
                                        c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32,
                                        c_void_p]),
    "dodt_correlation_grad_workspace_bytes": (c_size_t, [c_int32] * 9),
    "dodt_correlation_grad": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                      c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p,
                                      c_void_p, c_size_t, c_void_p]),
    "dodt_integral_banded_workspace_bytes": (c_size_t, [c_int32, c_int32]),
    "dodt_integral_band_rows": (c_int32, []),
    "dodt_integral_image_2d_banded": (c_int, [c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_size_t,
                                              POINTER(c_void_p), c_void_p]),
    "dodt_anchor_filter_fused_workspace_bytes": (c_size_t, [c_int64]),
    "dodt_anchor_filter_fused": (c_int, [c_void_p, POINTER(AnchorGrid), c_int64, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                                         c_int32, c_int32, c_double, c_double, c_void_p, c_void_p, c_void_p,
                                         c_void_p, POINTER(c_double), c_int32, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "dodt_three_d_iou_matrix": (c_int, [c_void_p, c_int32, c_void_p, c_int32, c_void_p, c_void_p]),
    "dodt_nms_workspace_bytes": (c_size_t, [c_int64]),
    "dodt_nms_state_offset": (c_size_t, [c_int64]),
    "dodt_nms": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int32, c_float, c_int32, c_int32,
                         c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
}

DODT_FE_VERSION = 201   # include/dodt_fe.h

_lib = None
_diag = False


def use_diag_library():
    """tools/ only: load libdodt_fe_diag.so (python -m dodt_b200._build --diag), the variant whose
    measurement knobs read the environment. Must be called before the first load(); the package
    itself never calls it, and no environment variable selects it."""
    global _diag
    if _lib is not None:
        raise RuntimeError("use_diag_library() must come before the library is first used")
    _diag = True


def load():
    """Load libdodt_fe.so (built by dodt_b200._build.build / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    _path = os.path.join(os.path.dirname(LIB_PATH), "libdodt_fe_diag.so") if _diag else LIB_PATH
    if not os.path.exists(_path):
        raise RuntimeError(
            "%s is not built. Run `python -m dodt_b200._build` (needs nvcc); "
            "this package has no CPU fallback." % _path)
    lib = ctypes.CDLL(_path)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the library lacks a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    got = lib.dodt_version()
    if got != DODT_FE_VERSION:
        raise RuntimeError("%s is version %d, this package expects %d: rebuild it (python -m dodt_b200._build)"
                           % (_path, got, DODT_FE_VERSION))
    _lib = lib
    return lib


class DodtError(RuntimeError):
    pass


def check(code, what):
    """Map a DODT_E* return code to the exception type the reference raises for that failure."""
    if code == DODT_OK:
        return
    lib = load()
    msg = "%s: %s" % (what, lib.dodt_strerror(code).decode())
    if code in (DODT_EINVAL, DODT_ESHAPE):
        raise ValueError(msg)
    if code == DODT_ECAPACITY:
        raise MemoryError(msg)
    if code == DODT_ECUDA:
        raise DodtError(msg + " — " + lib.dodt_last_cuda_error().decode())
    raise DodtError(msg)
