"""Tensor-level wrappers of the C ABI (include/dodt_fe.h): torch CUDA tensors in, torch CUDA tensors
out, work enqueued on torch's current stream. PyTorch is used for device memory and streams only;
every computation happens in libdodt_fe.so. No CPU path exists: inputs must live on a CUDA device.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import (BEV_STATS_LEN, DODT_F32, DODT_F64, MAX_DENSITY_LUT, MAX_SLICES, AnchorGrid, BevParams,
                   check, load)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _need_cuda(*tensors):
    """Every tensor of a call must live on the CURRENT CUDA device: kernels are enqueued on
    torch.cuda.current_stream(), which belongs to it (wrap the call in torch.cuda.device(...))."""
    cur = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("dodt_b200 runs on CUDA tensors only (there is no CPU fallback)")
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            raise RuntimeError("tensor on cuda:%d but the current device is cuda:%d: call inside "
                               "torch.cuda.device(tensor.device)" % (t.device.index, cur))


def _dtype_code(t):
    if t.dtype == torch.float32:
        return DODT_F32
    if t.dtype == torch.float64:
        return DODT_F64
    raise TypeError("expected a float32 or float64 tensor, got %s" % t.dtype)


def launch_count():
    return int(load().dodt_launch_count())


# ------------------------------------------------------------------------------------------ S1


def bev_grid(area_extents, voxel_size):
    """(nx, ny, nz, min_x, min_y, min_z) of wavedata voxel_grid_2d.py:125-130,142-143."""
    ext = (ctypes.c_double * 6)(*[float(v) for v in np.asarray(area_extents, dtype=np.float64).reshape(6)])
    grid = (ctypes.c_int32 * 6)()
    check(load().dodt_bev_grid(ext, float(voxel_size), grid), "dodt_bev_grid")
    return tuple(int(v) for v in grid)


def density_lut(norm_value):
    """min(1, log(n + 1) / norm) for n = 0.. until it saturates, computed with NumPy's log so that
    the density map is bit-identical to avod/core/bev_generators/bev_generator.py:34-35."""
    n = np.arange(MAX_DENSITY_LUT + 1)
    vals = np.minimum(1.0, np.log(n + 1) / norm_value)
    sat = np.flatnonzero(vals >= 1.0)
    if len(sat) == 0 or sat[0] > MAX_DENSITY_LUT:
        raise ValueError("density norm_value %r needs more than %d LUT entries" %
                         (norm_value, MAX_DENSITY_LUT))
    return vals[:sat[0]]


def make_bev_params(ground_plane, area_extents, voxel_size, height_lo, height_hi, num_slices,
                    filter_mode=True, occ_lo=0.2, occ_hi=2.0, norm_value=np.log(16)):
    if not 0 <= int(num_slices) <= MAX_SLICES:
        raise ValueError("num_slices must be in [0, %d]" % MAX_SLICES)
    p = BevParams()
    plane = [0.0, 0.0, 0.0, 0.0] if ground_plane is None else [float(v) for v in ground_plane]
    if len(plane) != 4:
        raise ValueError("ground_plane must have 4 coefficients")
    ext = np.asarray(area_extents, dtype=np.float64).reshape(-1)
    if ext.shape != (6,):
        raise ValueError("Extents are the wrong shape {}".format(np.asarray(area_extents).shape))
    p.plane[:] = plane
    p.extents[:] = [float(v) for v in ext]
    p.voxel_size = float(voxel_size)
    p.height_lo = float(height_lo)
    p.height_hi = float(height_hi)
    p.num_slices = int(num_slices)
    p.filter_mode = 1 if filter_mode else 0
    p.occ_lo = float(occ_lo)
    p.occ_hi = float(occ_hi)
    lut = density_lut(norm_value)
    p.density_lut_len = len(lut)
    for i, v in enumerate(lut):
        p.density_lut[i] = float(v)
    return p


def bev_workspace_bytes(n_points, num_slices, nx, nz):
    return int(load().dodt_bev_workspace_bytes(int(n_points), int(num_slices), int(nx), int(nz)))


def bev_slices(points, params, maps, occ, stats, workspace, winner_idx=None, counts=None,
               n_dev=None):
    """points (3, N) CUDA f32/f64 with unit inner stride. Fills maps [(S+1), nz, nx] f32,
    occ [nx, nz] u8 (or None), stats [24] i32. n_dev: optional device int32 point count."""
    _need_cuda(points, maps, occ, stats, workspace, winner_idx, counts, n_dev)
    if points.dim() != 2 or points.shape[0] != 3:
        raise ValueError("Points have the wrong shape: {}".format(tuple(points.shape)))
    n = points.shape[1]
    if n > 0 and points.stride(1) != 1:
        raise ValueError("points rows must be contiguous")
    row_stride = points.stride(0) if n > 0 else 0
    rc = load().dodt_bev_slices(_ptr(points), _dtype_code(points), n, _ptr(n_dev), max(row_stride, n),
                                ctypes.byref(params), _ptr(maps), _ptr(occ), _ptr(stats),
                                _ptr(winner_idx), _ptr(counts), _ptr(workspace),
                                workspace.numel() * workspace.element_size(), _stream())
    check(rc, "dodt_bev_slices")


# ------------------------------------------------------------------------------------------ S2


def integral_workspace_bytes(nx, nz):
    return int(load().dodt_integral_workspace_bytes(int(nx), int(nz)))


def integral_image_2d(occ, ii=None, workspace=None):
    """occ [nx, nz] u8 -> ii [(nx+1), (nz+1)] i32 (wavedata integral_image_2d.py:17-37)."""
    _need_cuda(occ, ii, workspace)
    if occ.dim() != 2:
        raise ValueError("Not a 2D image for integral image: input dim {}".format(occ.dim()))
    if occ.dtype != torch.uint8 or not occ.is_contiguous():
        raise TypeError("occupancy grid must be a contiguous uint8 tensor")
    nx, nz = occ.shape
    if ii is None:
        ii = torch.empty((nx + 1, nz + 1), dtype=torch.int32, device=occ.device)
    if workspace is None:
        workspace = torch.empty(integral_workspace_bytes(nx, nz), dtype=torch.uint8, device=occ.device)
    check(load().dodt_integral_image_2d(_ptr(occ), nx, nz, _ptr(ii), _ptr(workspace),
                                        workspace.numel(), _stream()), "dodt_integral_image_2d")
    return ii


def map_to_index(coords, voxel_size, min_x, min_z, nx, nz):
    """coords (n, 2) f32/f64 CUDA -> (n, 2) i32 (wavedata voxel_grid_2d.py:162-186)."""
    _need_cuda(coords)
    coords = coords.contiguous()
    n = coords.shape[0]
    out = torch.empty((n, 2), dtype=torch.int32, device=coords.device)
    check(load().dodt_map_to_index(_ptr(coords), _dtype_code(coords), n, float(voxel_size), min_x,
                                   min_z, nx, nz, _ptr(out), _stream()), "dodt_map_to_index")
    return out


def anchor_filter_2d(anchors, ii, nx, nz, min_x, min_z, voxel_size, density_threshold=1,
                     keep=None, scores=None):
    """anchors (n, 6) f32/f64 CUDA, ii from integral_image_2d -> keep (n,) u8."""
    _need_cuda(anchors, ii, keep, scores)
    if anchors.dim() != 2 or anchors.shape[1] != 6:
        raise TypeError("Given input does not have valid number of attributes. "
                        "Should be N x 6 for anchor.")
    anchors = anchors.contiguous()
    n = anchors.shape[0]
    if keep is None:
        keep = torch.empty((n,), dtype=torch.uint8, device=anchors.device)
    check(load().dodt_anchor_filter_2d(_ptr(anchors), _dtype_code(anchors), n, _ptr(ii), nx, nz,
                                       min_x, min_z, float(voxel_size), float(density_threshold),
                                       _ptr(keep), _ptr(scores), _stream()),
          "dodt_anchor_filter_2d")
    return keep


def integral_banded_workspace_bytes(nx, nz):
    return int(load().dodt_integral_banded_workspace_bytes(int(nx), int(nz)))


def integral_image_2d_banded(occ, ii_local, workspace):
    """occ [nx, nz] u8 -> band-local integral image ii_local [(nx+1), (nz+1)] i32 and, inside
    `workspace` (zero before its first use), the exclusive band offsets: ONE launch. Returns the
    band offsets as a tensor view [bands, nz] of the workspace and the rows per band; the full image
    is ii_local[X, Z] + bandoff[(X-1) // band_rows, Z-1] (X, Z >= 1)."""
    _need_cuda(occ, ii_local, workspace)
    if occ.dim() != 2 or occ.dtype != torch.uint8 or not occ.is_contiguous():
        raise TypeError("occupancy grid must be a contiguous 2-D uint8 tensor")
    nx, nz = occ.shape
    out = ctypes.c_void_p(0)
    check(load().dodt_integral_image_2d_banded(_ptr(occ), nx, nz, _ptr(ii_local), _ptr(workspace),
                                               workspace.numel(), ctypes.byref(out), _stream()),
          "dodt_integral_image_2d_banded")
    band_rows = int(load().dodt_integral_band_rows())
    bands = (nx + band_rows - 1) // band_rows
    off = out.value - workspace.data_ptr()
    bandoff = workspace[off:off + bands * nz * 4].view(torch.int32).view(bands, nz)
    return bandoff, band_rows


def anchor_filter_fused_workspace(n, device):
    """Zeroed workspace of anchor_filter_fused (every call leaves it zero again)."""
    return torch.zeros(max(int(load().dodt_anchor_filter_fused_workspace_bytes(int(n))), 256),
                       dtype=torch.uint8, device=device)


def anchor_filter_fused(anchors, ii, nx, nz, min_x, min_z, voxel_size, density_threshold, keep, kept_idx,
                        n_kept, workspace, bandoff=None, band_rows=0, anchor_bev_boxes=None, k_bev_boxes=None,
                        anchor_img_boxes=None, k_img_boxes=None, rpn_scores=None, k_scores=None,
                        rpn_offsets=None, bev_extents=None, k_rpn_boxes=None, tf_float32=False, grid=None):
    """S2 of the frame stream in one launch: keep mask of the float64 anchors, ordered compaction
    (kept_idx, n_kept on the device) and, at the compacted positions, the anchors' crop boxes, RPN
    scores and decoded BEV boxes. ii is the full integral image, or the band-local one when
    bandoff / band_rows (integral_image_2d_banded) are given. tf_float32: decode like rpn_decode.
    grid = (area_extents, anchor_3d_sizes, anchor_stride, ground_plane) with anchors=None: the anchors
    are that grid (grid_anchors) and are evaluated from the index inside the kernel, not read."""
    _need_cuda(anchors, ii, bandoff, keep, kept_idx, n_kept, workspace, anchor_bev_boxes, k_bev_boxes,
               anchor_img_boxes, k_img_boxes, rpn_scores, k_scores, rpn_offsets, k_rpn_boxes)
    grid_arg = None
    if anchors is None:
        if grid is None:
            raise ValueError("anchor_filter_fused needs anchors or grid")
        ext, sizes, stride, plane = grid
        sizes = np.asarray(sizes, dtype=np.float64).reshape(-1, 3)
        ga = AnchorGrid()
        ga.extents[:] = [float(v) for v in np.asarray(ext, dtype=np.float64).reshape(-1)]
        ga.stride[:] = [float(v) for v in stride]
        ga.plane[:] = [float(v) for v in plane]
        for i, v in enumerate(sizes.reshape(-1)):
            ga.sizes[i] = float(v)
        ga.n_sizes = len(sizes)
        grid_arg = ctypes.byref(ga)
        n = keep.shape[0]
    else:
        if anchors.dtype != torch.float64 or anchors.dim() != 2 or anchors.shape[1] != 6 or not anchors.is_contiguous():
            raise TypeError("anchor_filter_fused expects contiguous float64 anchors (n, 6)")
        n = anchors.shape[0]
    check(load().dodt_anchor_filter_fused(
        _ptr(anchors), grid_arg, n, _ptr(ii), _ptr(bandoff), int(band_rows), nx, nz, min_x, min_z, float(voxel_size),
        float(density_threshold), _ptr(anchor_bev_boxes), _ptr(anchor_img_boxes), _ptr(rpn_scores),
        _ptr(rpn_offsets), _dbl(bev_extents, 4) if bev_extents is not None else None, int(bool(tf_float32)),
        _ptr(keep),
        _ptr(kept_idx), _ptr(n_kept), _ptr(k_bev_boxes), _ptr(k_img_boxes), _ptr(k_scores), _ptr(k_rpn_boxes),
        _ptr(workspace), workspace.numel(), _stream()), "dodt_anchor_filter_fused")


# ------------------------------------------------------------------------------------------ S3


def compact_mask(keep, idx=None, count=None, workspace=None):
    """keep [n] u8 -> (idx [n] i32 ascending positions of non-zero entries, count [1] i32), device."""
    _need_cuda(keep, idx, count, workspace)
    if keep.dtype == torch.bool:
        keep = keep.view(torch.uint8)
    keep = keep.contiguous()
    n = keep.numel()
    dev = keep.device
    if idx is None:
        idx = torch.empty((n,), dtype=torch.int32, device=dev)
    if count is None:
        count = torch.empty((1,), dtype=torch.int32, device=dev)
    if workspace is None:
        workspace = torch.empty(max(int(load().dodt_compact_workspace_bytes(n)), 256),
                                dtype=torch.uint8, device=dev)
    check(load().dodt_compact_mask(_ptr(keep), n, _ptr(idx), _ptr(count), _ptr(workspace),
                                   workspace.numel(), _stream()), "dodt_compact_mask")
    return idx, count


def gather_rows(src, idx, count, out=None):
    """out[i] = src[idx[i]] for i < count[0]; src [m, w] (or [m]) f32, idx [n] i32, count [1] i32."""
    _need_cuda(src, idx, count, out)
    src = src.contiguous()
    width = 1 if src.dim() == 1 else int(src.shape[1])
    n_max = idx.numel()
    if out is None:
        out = torch.empty((n_max,) if src.dim() == 1 else (n_max, width), dtype=torch.float32,
                          device=src.device)
    check(load().dodt_gather_rows(_ptr(src), width, _ptr(idx), _ptr(count), n_max, _ptr(out),
                                  _stream()), "dodt_gather_rows")
    return out


def gather_rows_multi(pairs, idx, count):
    """One launch for several (src [m, w], dst [n_max, w]) pairs sharing idx / count."""
    specs = (_lib.GatherSpec * len(pairs))()
    for k, (src, dst) in enumerate(pairs):
        _need_cuda(src, dst)
        if not src.is_contiguous() or not dst.is_contiguous():
            raise ValueError("gather_rows_multi needs contiguous tensors")
        specs[k].src = src.data_ptr()
        specs[k].dst = dst.data_ptr()
        specs[k].width = 1 if src.dim() == 1 else int(src.shape[1])
    check(load().dodt_gather_rows_multi(specs, len(pairs), _ptr(idx), _ptr(count), idx.numel(),
                                        _stream()), "dodt_gather_rows_multi")


# ------------------------------------------------------------------------------------------ lidar


def lidar_to_camera(velo, rectified, p2=None, im_size=None, dtype=torch.float64, out=None,
                    count=None, workspace=None, ego=None, aligned=None):
    """velo [n, 4] CUDA float32 (KITTI .bin rows) -> (points (3, n) of dtype with the first
    count[0] columns valid, count [1] int32), both on the device, no synchronisation.
    rectified: rows 0..2 of R0_rect(4x4).Tr_velo_to_cam(4x4); im_size = [w, h] enables the
    image-frustum filter of wavedata tracking_utils.get_lidar_point_cloud.
    ego = (trans[3], matrix[3, 3]): first move the scan into the previous frame's LiDAR frame,
    float32((xyz + trans) @ matrix) (kitti_tracking_dataset.py:317-328); aligned: optional [n, 4]
    float32 out tensor that receives the moved scan."""
    _need_cuda(velo, out, count, workspace, aligned)
    if velo.dim() != 2 or velo.shape[1] != 4 or velo.dtype != torch.float32:
        raise ValueError("velo must be an [n, 4] float32 tensor (x, y, z, intensity)")
    velo = velo.contiguous()
    n = velo.shape[0]
    dev = velo.device
    if out is None:
        out = torch.empty((3, max(n, 1)), dtype=dtype, device=dev)
    if count is None:
        count = torch.empty((1,), dtype=torch.int32, device=dev)
    if workspace is None:
        workspace = torch.empty(int(load().dodt_lidar_workspace_bytes(n)), dtype=torch.uint8, device=dev)
    if aligned is not None and (ego is None or aligned.shape != velo.shape or aligned.dtype != torch.float32
                                or not aligned.is_contiguous()):
        raise ValueError("aligned must be a contiguous [n, 4] float32 tensor and needs ego")
    w, h = (int(im_size[0]), int(im_size[1])) if im_size is not None else (0, 0)
    trans, matrix = (None, None) if ego is None else (_dbl(ego[0], 3), _dbl(ego[1], 9))
    check(load().dodt_lidar_to_camera_aligned(_ptr(velo), n, trans, matrix,
                                              _ptr(aligned) if aligned is not None else None,
                                              _dbl(np.asarray(rectified)[:3], 12),
                                              _dbl(p2, 12) if p2 is not None else None, w, h, _ptr(out),
                                              _dtype_code(out), out.stride(0), _ptr(count), _ptr(workspace),
                                              workspace.numel(), _stream()), "dodt_lidar_to_camera_aligned")
    return out, count


def lidar_align(velo, trans, matrix, aligned=None):
    """velo [n, 4] CUDA float32 -> the scan moved into the previous frame's LiDAR frame,
    float32((xyz + trans) @ matrix), intensity unchanged (kitti_tracking_dataset.py:317-328)."""
    _need_cuda(velo, aligned)
    if velo.dim() != 2 or velo.shape[1] != 4 or velo.dtype != torch.float32:
        raise ValueError("velo must be an [n, 4] float32 tensor (x, y, z, intensity)")
    velo = velo.contiguous()
    if aligned is None:
        aligned = torch.empty_like(velo)
    check(load().dodt_lidar_to_camera_aligned(_ptr(velo), velo.shape[0], _dbl(trans, 3), _dbl(matrix, 9),
                                              _ptr(aligned), None, None, 0, 0, None, DODT_F32, 0, None, None, 0,
                                              _stream()), "dodt_lidar_to_camera_aligned")
    return aligned


# ------------------------------------------------------------------------------------------ anchors


def _dbl(values, n):
    arr = np.asarray(values, dtype=np.float64).reshape(-1)
    if arr.shape != (n,):
        raise ValueError("expected %d values, got shape %r" % (n, np.shape(values)))
    return (ctypes.c_double * n)(*[float(v) for v in arr])


def grid_anchor_shape(area_extents, anchor_stride, n_sizes):
    """(nz, nx, n_sizes, 2) of avod grid_anchor_3d_generator.tile_anchors_3d."""
    shape = (ctypes.c_int32 * 4)()
    check(load().dodt_grid_anchor_shape(_dbl(area_extents, 6), _dbl(anchor_stride, 2), int(n_sizes), shape),
          "dodt_grid_anchor_shape")
    return tuple(int(v) for v in shape)


def grid_anchors(area_extents, anchor_3d_sizes, anchor_stride, ground_plane, device=None):
    """box_3d_to_anchor(tile_anchors_3d(...)) as a CUDA float64 [N, 6] tensor."""
    sizes = np.asarray(anchor_3d_sizes, dtype=np.float64).reshape(-1, 3)
    nz, nx, ns, nr = grid_anchor_shape(area_extents, anchor_stride, len(sizes))
    out = torch.empty((nz * nx * ns * nr, 6), dtype=torch.float64, device=device or "cuda")
    check(load().dodt_grid_anchors(_dbl(area_extents, 6), _dbl(sizes, 3 * len(sizes)), len(sizes),
                                   _dbl(anchor_stride, 2), _dbl(ground_plane, 4), _ptr(out), _stream()),
          "dodt_grid_anchors")
    return out


def _anchors_arg(anchors):
    _need_cuda(anchors)
    if anchors.dim() != 2 or anchors.shape[1] != 6:
        raise TypeError("Given input does not have valid number of attributes. "
                        "Should be N x 6 for anchor.")
    return anchors.contiguous()


def project_to_bev(anchors, bev_extents, tf_order=False, want_metres=False):
    """anchors [n, 6] f32/f64 CUDA -> normalised corners float32 [n, 4] (and corners in metres)."""
    a = _anchors_arg(anchors)
    n = a.shape[0]
    norm = torch.empty((n, 4), dtype=torch.float32, device=a.device)
    metres = torch.empty((n, 4), dtype=torch.float32, device=a.device) if want_metres else None
    check(load().dodt_project_to_bev(_ptr(a), _dtype_code(a), n, _dbl(bev_extents, 4), int(bool(tf_order)),
                                     _ptr(norm), _ptr(metres), _stream()), "dodt_project_to_bev")
    return (metres, norm) if want_metres else norm


def project_to_image_space(anchors, stereo_calib_p2, image_shape, tf_order=False, want_pixels=False):
    a = _anchors_arg(anchors)
    n = a.shape[0]
    norm = torch.empty((n, 4), dtype=torch.float32, device=a.device)
    pixels = torch.empty((n, 4), dtype=torch.float32, device=a.device) if want_pixels else None
    check(load().dodt_project_to_image_space(_ptr(a), _dtype_code(a), n, _dbl(stereo_calib_p2, 12),
                                             int(image_shape[0]), int(image_shape[1]),
                                             int(bool(tf_order)), _ptr(norm), _ptr(pixels), _stream()),
          "dodt_project_to_image_space")
    return (pixels, norm) if want_pixels else norm


def offset_to_anchor(anchors, offsets):
    a = _anchors_arg(anchors)
    o = _anchors_arg(offsets)
    if a.shape != o.shape:
        raise ValueError("anchors and offsets must have the same shape")
    out = torch.empty((a.shape[0], 6), dtype=torch.float64, device=a.device)
    check(load().dodt_offset_to_anchor(_ptr(a), _dtype_code(a), _ptr(o), _dtype_code(o), a.shape[0],
                                       _ptr(out), _stream()), "dodt_offset_to_anchor")
    return out


def rpn_decode(anchors, offsets, idx, count, bev_extents, stereo_calib_p2, image_shape, bev_boxes,
               img_boxes, idx2=None, tf_float32=False):
    """Regressed + projected boxes of the anchors idx[:count[0]] (device-side count), or of
    idx[idx2[:count[0]]]: bev_boxes [n_max, 4] [z1,x1,z2,x2], img_boxes [n_max, 4] [y1,x1,y2,x2],
    float32, normalised; either may be None. tf_float32: the float32 tf.Tensor branches the
    reference's inference graph runs (dt_rpn_model.py:568-591) instead of the float64 NumPy ones."""
    _need_cuda(anchors, offsets, idx, count, bev_boxes, img_boxes, idx2)
    if anchors.dtype != torch.float64 or offsets.dtype != torch.float32:
        raise TypeError("rpn_decode expects float64 anchors and float32 offsets")
    if not (anchors.is_contiguous() and offsets.is_contiguous()):
        raise ValueError("rpn_decode needs contiguous tensors")
    n_max = (bev_boxes if bev_boxes is not None else img_boxes).shape[0]
    check(load().dodt_rpn_decode(_ptr(anchors), _ptr(offsets), _ptr(idx), _ptr(idx2), _ptr(count), n_max,
                                 _dbl(bev_extents, 4), _dbl(stereo_calib_p2, 12), int(image_shape[0]),
                                 int(image_shape[1]), int(bool(tf_float32)), _ptr(bev_boxes), _ptr(img_boxes),
                                 _stream()),
          "dodt_rpn_decode")


def emit_detections(boxes, scores, keep, n_keep, block, frame_id=None, row_io=None, rewrite=False):
    """Append keep[:n_keep[0]] of one frame to a dodt_b200.shard.DetectionBlock living on the device
    (rows = box, score, index). frame_id: optional device int32 [2] (sequence, frame). row_io:
    optional device int32 [1] that receives the block row the frame was given; with rewrite=True
    the list replaces row row_io[0] instead of taking a new one."""
    _need_cuda(boxes, scores, keep, n_keep, block.rows, block.counts, block.frame_ids, block.cursor,
               frame_id, row_io)
    if rewrite and row_io is None:
        raise ValueError("rewrite needs row_io")
    max_frames, max_det = int(block.rows.shape[0]), int(block.rows.shape[1])
    if keep.numel() < max_det:
        raise ValueError("keep holds fewer than max_det entries")
    check(load().dodt_emit_detections(_ptr(boxes), _ptr(scores), _ptr(keep), _ptr(n_keep), max_det,
                                      _ptr(frame_id), _ptr(block.rows), _ptr(block.counts),
                                      _ptr(block.frame_ids), _ptr(block.cursor), max_frames, _ptr(row_io),
                                      1 if rewrite else 0, _stream()),
          "dodt_emit_detections")


def crop_and_resize_multi(triples, crop_size, extrapolation_value=0.0, n_dev=None, box_ind=None):
    """One launch for several (image [B,H,W,C], boxes [n,4], out [n,ch,cw,C]) triples that share
    the box count, box_ind (None = zeros) and crop size."""
    specs = (_lib.CropSpec * len(triples))()
    n = batch = None
    for k, (image, boxes, out) in enumerate(triples):
        _need_cuda(image, boxes, out)
        B, H, W, C = image.shape
        if n is None:
            n, batch = boxes.shape[0], B
        if boxes.shape[0] != n or B != batch or not (image.is_contiguous() and boxes.is_contiguous()
                                                      and out.is_contiguous()):
            raise ValueError("crop_and_resize_multi: inconsistent or non-contiguous inputs")
        specs[k].image, specs[k].boxes, specs[k].crops = image.data_ptr(), boxes.data_ptr(), out.data_ptr()
        specs[k].height, specs[k].width, specs[k].channels = H, W, C
    check(load().dodt_crop_and_resize_multi(specs, len(triples), batch, _ptr(box_ind), n,
                                            _ptr(n_dev), int(crop_size[0]), int(crop_size[1]),
                                            float(extrapolation_value), _stream()),
          "dodt_crop_and_resize_multi")


def crop_and_resize(image, boxes, box_ind, crop_size, extrapolation_value=0.0, out=None,
                    n_dev=None):
    """image [B,H,W,C] f32 NHWC, boxes [n,4] f32, box_ind [n] i32 (None = all zero) ->
    [n, ch, cw, C] f32. n_dev: optional device int32 box count (rows past it are left untouched)."""
    if box_ind is None:
        return _crop_no_ind(image, boxes, crop_size, extrapolation_value, out, n_dev)
    _need_cuda(image, boxes, box_ind, out, n_dev)
    if image.dim() != 4:
        raise ValueError("image must be 4-D [batch, height, width, channels]")
    if boxes.dim() != 2 or boxes.shape[1] != 4:
        raise ValueError("boxes must be 2-D [num_boxes, 4]")
    if box_ind.dim() != 1 or box_ind.shape[0] != boxes.shape[0]:
        raise ValueError("box_ind must be 1-D with one entry per box")
    ch, cw = int(crop_size[0]), int(crop_size[1])
    if ch <= 0 or cw <= 0:
        raise ValueError("crop dimensions must be positive")
    image = image.contiguous()
    boxes = boxes.contiguous()
    box_ind = box_ind.contiguous()
    if image.dtype != torch.float32 or boxes.dtype != torch.float32 or box_ind.dtype != torch.int32:
        raise TypeError("crop_and_resize expects float32 image/boxes and int32 box_ind")
    B, H, W, C = image.shape
    n = boxes.shape[0]
    if out is None:
        # rows with an out-of-range box_ind are left untouched by TF; give them a defined value
        out = torch.zeros((n, ch, cw, C), dtype=torch.float32, device=image.device) \
            if n and bool(((box_ind < 0) | (box_ind >= B)).any()) \
            else torch.empty((n, ch, cw, C), dtype=torch.float32, device=image.device)
    check(load().dodt_crop_and_resize(_ptr(image), B, H, W, C, _ptr(boxes), _ptr(box_ind), n,
                                      _ptr(n_dev), ch, cw, float(extrapolation_value), _ptr(out),
                                      _stream()), "dodt_crop_and_resize")
    return out


def _crop_no_ind(image, boxes, crop_size, extrapolation_value, out, n_dev):
    """Hot-path form: single image batch semantics (box_ind all zero), preallocated buffers."""
    _need_cuda(image, boxes, out, n_dev)
    B, H, W, C = image.shape
    ch, cw = int(crop_size[0]), int(crop_size[1])
    n = boxes.shape[0]
    if out is None:
        out = torch.empty((n, ch, cw, C), dtype=torch.float32, device=image.device)
    check(load().dodt_crop_and_resize(_ptr(image), B, H, W, C, _ptr(boxes), None, n, _ptr(n_dev),
                                      ch, cw, float(extrapolation_value), _ptr(out), _stream()),
          "dodt_crop_and_resize")
    return out


# ------------------------------------------------------------------------------------------ S4


def correlation_out_shape(H, W, kernel_size, max_displacement, stride_1, stride_2, pad):
    hwc = (ctypes.c_int32 * 3)()
    rc = load().dodt_correlation_out_shape(H, W, kernel_size, max_displacement, stride_1, stride_2,
                                           pad, hwc)
    if rc == _lib.DODT_ESHAPE:
        raise ValueError("Neighborhood and kernel don't fit in input.")
    check(rc, "dodt_correlation_out_shape")
    return int(hwc[0]), int(hwc[1]), int(hwc[2])


def correlation(a, b, kernel_size, max_displacement, stride_1, stride_2, pad, out=None,
                max_ctas=0):
    """max_ctas > 0: cap on the persistent CTAs of the launch (dodt_correlation_shared)."""
    _need_cuda(a, b, out)
    if a.dim() != 4:
        raise ValueError("input_a must have rank 4")
    if b.dim() != 4:
        raise ValueError("input_b must have rank 4")
    if a.shape != b.shape:
        raise ValueError("input_a and input_b must have the same shape")
    if kernel_size % 2 == 0:
        raise ValueError("kernel_size must be odd")
    if a.dtype != torch.float32 or b.dtype != torch.float32:
        raise TypeError("correlation expects float32 inputs")
    a = a.contiguous()
    b = b.contiguous()
    N, H, W, C = a.shape
    oh, ow, oc = correlation_out_shape(H, W, kernel_size, max_displacement, stride_1, stride_2, pad)
    if out is None:
        out = torch.empty((N, oh, ow, oc), dtype=torch.float32, device=a.device)
    check(load().dodt_correlation_shared(_ptr(a), _ptr(b), N, H, W, C, kernel_size,
                                         max_displacement, stride_1, stride_2, pad, _ptr(out),
                                         int(max_ctas), _stream()), "dodt_correlation")
    return out


def correlation_stream(maps, kernel_size, max_displacement, stride_1, stride_2, pad, outs=None,
                       max_ctas=0):
    """outs[j] = correlation(maps[j], maps[j + 1]) for the consecutive frames of a stream
    (dodt_correlation_stream): maps is a list of [1,H,W,C] float32 CUDA tensors; up to
    CORR_STREAM_MAX_PAIRS pairs share a launch and the map two pairs have in common is read once."""
    maps = list(maps)
    if len(maps) < 2:
        raise ValueError("correlation_stream needs at least two feature maps")
    _need_cuda(*maps)
    first = maps[0]
    if first.dim() != 4 or first.shape[0] != 1:
        raise ValueError("feature maps must be [1, H, W, C]")
    for m in maps:
        if m.shape != first.shape:
            raise ValueError("all feature maps must have the same shape")
        if m.dtype != torch.float32:
            raise TypeError("correlation expects float32 inputs")
    if kernel_size % 2 == 0:
        raise ValueError("kernel_size must be odd")
    maps = [m.contiguous() for m in maps]
    _, H, W, C = first.shape
    oh, ow, oc = correlation_out_shape(H, W, kernel_size, max_displacement, stride_1, stride_2, pad)
    if outs is None:
        outs = [torch.empty((1, oh, ow, oc), dtype=torch.float32, device=first.device)
                for _ in range(len(maps) - 1)]
    outs = list(outs)
    if len(outs) != len(maps) - 1:
        raise ValueError("one output per pair of consecutive maps")
    _need_cuda(*outs)
    for o in outs:
        if tuple(o.shape) != (1, oh, ow, oc) or not o.is_contiguous() or o.dtype != torch.float32:
            raise ValueError("outputs must be contiguous float32 [1, %d, %d, %d]" % (oh, ow, oc))
    mp = (ctypes.c_void_p * len(maps))(*[m.data_ptr() for m in maps])
    op = (ctypes.c_void_p * len(outs))(*[o.data_ptr() for o in outs])
    check(load().dodt_correlation_stream(mp, len(maps), op, H, W, C, kernel_size, max_displacement,
                                         stride_1, stride_2, pad, int(max_ctas), _stream()),
          "dodt_correlation_stream")
    return outs


def correlation_grad_workspace_bytes(N, H, W, C, kernel_size, max_displacement, stride_1, stride_2, pad):
    return int(load().dodt_correlation_grad_workspace_bytes(N, H, W, C, kernel_size, max_displacement,
                                                             stride_1, stride_2, pad))


def correlation_grad(grad, a, b, kernel_size, max_displacement, stride_1, stride_2, pad,
                     grad_a=None, grad_b=None, workspace=None, need_a=True, need_b=True):
    """CorrelationGrad (correlation_grad_kernel.cc:28-151): grad [N,oh,ow,oc], a, b [N,H,W,C] f32 ->
    (grad_a, grad_b) [N,H,W,C]; a gradient that is not needed is returned as None."""
    _need_cuda(grad, a, b, grad_a, grad_b, workspace)
    if a.dim() != 4:
        raise ValueError("input_a must have rank 4")
    if b.dim() != 4:
        raise ValueError("input_b must have rank 4")
    if a.shape != b.shape:
        raise ValueError("input_a and input_b must have the same shape")
    if kernel_size % 2 == 0:
        raise ValueError("kernel_size must be odd")
    if a.dtype != torch.float32 or b.dtype != torch.float32 or grad.dtype != torch.float32:
        raise TypeError("correlation_grad expects float32 tensors")
    a, b, grad = a.contiguous(), b.contiguous(), grad.contiguous()
    N, H, W, C = a.shape
    oh, ow, oc = correlation_out_shape(H, W, kernel_size, max_displacement, stride_1, stride_2, pad)
    if tuple(grad.shape) != (N, oh, ow, oc):
        raise ValueError("gradients must have the shape of the correlation output %r" % ((N, oh, ow, oc),))
    if need_a and grad_a is None:
        grad_a = torch.empty_like(a)
    if need_b and grad_b is None:
        grad_b = torch.empty_like(b)
    ws_bytes = correlation_grad_workspace_bytes(N, H, W, C, kernel_size, max_displacement, stride_1,
                                                stride_2, pad) if need_b else 0
    if ws_bytes and workspace is None:
        workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=a.device)
    check(load().dodt_correlation_grad(_ptr(grad), _ptr(a), _ptr(b), N, H, W, C, kernel_size,
                                       max_displacement, stride_1, stride_2, pad,
                                       _ptr(grad_a) if need_a else None,
                                       _ptr(grad_b) if need_b else None,
                                       _ptr(workspace) if workspace is not None else None,
                                       int(workspace.numel() * workspace.element_size()) if workspace is not None else 0,
                                       _stream()), "dodt_correlation_grad")
    return (grad_a if need_a else None), (grad_b if need_b else None)


# ------------------------------------------------------------------------------------------ S5


def nms_workspace_bytes(n):
    return int(load().dodt_nms_workspace_bytes(int(n)))


def nms(boxes, scores, max_out, iou_threshold, keep=None, n_keep=None, workspace=None,
        n_dev=None, max_windows=0, first_window=0):
    """boxes [n,4] f32, scores [n] f32 -> (keep [max_out] i32 padded with -1, n_keep [2] i32:
    number selected, 1 if the selection is complete), both on the device (no synchronisation).
    n_dev: optional device int32 candidate count; max_windows / first_window: see
    include/dodt_fe.h (bounded launch count, continuation of an incomplete selection)."""
    _need_cuda(boxes, scores, keep, n_keep, workspace, n_dev)
    if boxes.dim() != 2 or boxes.shape[1] != 4:
        raise ValueError("boxes must be 2-D [num_boxes, 4]")
    if scores.dim() != 1 or scores.shape[0] != boxes.shape[0]:
        raise ValueError("scores has incompatible shape")
    if boxes.dtype != torch.float32 or scores.dtype != torch.float32:
        raise TypeError("non_max_suppression expects float32 boxes and scores")
    boxes = boxes.contiguous()
    scores = scores.contiguous()
    n = boxes.shape[0]
    max_out = int(max_out)
    if max_out < 0:
        raise ValueError("max_output_size must be non-negative")
    dev = boxes.device
    if keep is None:
        keep = torch.empty((max_out,), dtype=torch.int32, device=dev)
    if n_keep is None:
        n_keep = torch.empty((2,), dtype=torch.int32, device=dev)
    if workspace is None:
        workspace = torch.empty(max(nms_workspace_bytes(n), 256), dtype=torch.uint8, device=dev)
    check(load().dodt_nms(_ptr(boxes), _ptr(scores), n, _ptr(n_dev), max_out, float(iou_threshold),
                          int(first_window), int(max_windows), _ptr(keep), _ptr(n_keep), _ptr(workspace),
                          workspace.numel(), _stream()), "dodt_nms")
    return keep, n_keep
