"""dodt_b200 — B200 (sm_100a) implementation of the DODT/AVOD per-frame proposal front end:
BEV slice maps, empty-anchor filter, crop_and_resize, inter-frame correlation and NMS, each behind
the reference's own Python call signature. All computation happens in hand-written CUDA kernels
reached through the C ABI of include/dodt_fe.h (dodt_b200/libdodt_fe.so); there is no CPU fallback.
"""
from ._lib import LIB_PATH, DodtError, load  # noqa: F401
from .anchor_filter import get_empty_anchor_filter_2d  # noqa: F401
from .bev_slices import BevSlices  # noqa: F401
from .correlation import correlation, correlation_grad, correlation_stream  # noqa: F401
from .evaluation import iou_3d, three_d_iou, three_d_iou_matrix, track_iou  # noqa: F401
from .tf_image import crop_and_resize, non_max_suppression  # noqa: F401
from .voxel_grid_2d import VoxelGrid2D  # noqa: F401

__all__ = ["BevSlices", "VoxelGrid2D", "get_empty_anchor_filter_2d", "crop_and_resize",
           "non_max_suppression", "correlation", "correlation_grad", "correlation_stream", "three_d_iou", "three_d_iou_matrix", "iou_3d", "track_iou", "load", "LIB_PATH", "DodtError"]
