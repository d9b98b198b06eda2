"""Drop-in for avod/core/bev_generators/bev_slices.py (BevSlices) running on sm_100a kernels.

Same constructor and `generate_bev(source, point_cloud, ground_plane, area_extents, voxel_size)`
signature as the reference (bev_slices.py:14-31,33-38; built by
avod/builders/bev_generator_builder.py:4-12 and called from
avod/datasets/kitti/kitti_utils.py:122-126). Differences, all deliberate:

  * maps come back as float32 (the reference returns float64 and TensorFlow casts them to the
    float32 placeholder, avod/core/models/dt_rpn_model.py:183); each value equals
    np.float32(reference value), the density map bit-for-bit, the height maps bit-for-bit for the
    axis-aligned plane DODT hard-codes (wavedata tracking_utils.py:239);
  * float32 point clouds are treated as exact float64 values (the reference pipeline always hands
    float64, wavedata calib_utils.py:484-523);
  * a cloud with no point inside the density slice makes the reference crash with IndexError
    (voxel_grid_2d.py:100-102); here it raises IndexError when `strict` (default for NumPy input)
    and returns all-zero maps otherwise;
  * `kitti_utils` is accepted for signature compatibility but not called — the slice predicate of
    create_slice_filter / get_point_filter is evaluated inside the kernel.

The same pass can also emit the occupancy grid of the 0.2–2.0 m anchor-filter slice
(kitti_utils.py:212-277), which removes the reference's second read + voxelisation of the cloud
(its own TODO at avod/core/models/dt_rpn_model.py:943): see `generate_bev_and_voxel_grid`.
"""
import numpy as np
import torch

from . import ops
from ._lib import BEV_STATS_LEN, STAT_DENSITY, STAT_OOB, STAT_OVERFLOW
from .voxel_grid_2d import VoxelGrid2D


def _to_device_points(point_cloud, device):
    """(3, N) ndarray / tensor -> CUDA tensor with 16-byte aligned rows. Returns (tensor, was_numpy)."""
    was_numpy = not torch.is_tensor(point_cloud)
    if was_numpy:
        arr = np.asarray(point_cloud)
        if arr.dtype not in (np.float32, np.float64):
            arr = arr.astype(np.float64)
        t = torch.from_numpy(np.ascontiguousarray(arr))
    else:
        t = point_cloud
        if t.dtype not in (torch.float32, torch.float64):
            t = t.double()
    if t.dim() != 2 or t.shape[0] != 3:
        raise ValueError("Points have the wrong shape: {}".format(tuple(t.shape)))
    if t.is_cuda and t.stride(1) == 1:
        return t, was_numpy
    n = t.shape[1]
    per16 = 16 // t.element_size()
    pitch = (n + per16 - 1) // per16 * per16
    dev = torch.empty((3, max(pitch, per16)), dtype=t.dtype, device=device)
    dev[:, :n].copy_(t, non_blocking=True)
    return dev[:, :n], was_numpy


class BevBuffers:
    """Device buffers of one BEV configuration, reused frame after frame."""

    def __init__(self, nx, nz, num_slices, max_points, device, with_occ=True, debug=False):
        self.nx, self.nz, self.num_slices, self.max_points = nx, nz, num_slices, max_points
        self.maps = torch.empty((num_slices + 1, nz, nx), dtype=torch.float32, device=device)
        self.occ = torch.empty((nx, nz), dtype=torch.uint8, device=device) if with_occ else None
        self.stats = torch.empty((BEV_STATS_LEN,), dtype=torch.int32, device=device)
        self.workspace = torch.empty(
            max(ops.bev_workspace_bytes(max_points, num_slices, nx, nz), 256),
            dtype=torch.uint8, device=device)
        self.winner = torch.empty((max(num_slices, 1), nz, nx), dtype=torch.int32, device=device) \
            if debug else None
        self.counts = torch.empty((nz, nx), dtype=torch.int32, device=device) if debug else None


class BevSlices:
    """BEV maps created using slices of the point cloud (reference: bev_slices.py:8)."""

    NORM_VALUES = {
        'lidar': np.log(16),
    }

    def __init__(self, config, kitti_utils=None, device=None):
        self.height_lo = config.height_lo
        self.height_hi = config.height_hi
        self.num_slices = config.num_slices
        self.kitti_utils = kitti_utils
        self.height_per_division = (self.height_hi - self.height_lo) / self.num_slices
        self.device = torch.device("cuda" if device is None else device)
        self._buffers = {}

    # -- device API ---------------------------------------------------------------------------
    def buffers(self, nx, nz, n_points, with_occ, debug=False):
        key = (nx, nz, with_occ, debug)
        buf = self._buffers.get(key)
        if buf is None or buf.max_points < n_points:
            cap = max(n_points, 1 << 17)
            buf = BevBuffers(nx, nz, self.num_slices, cap, self.device, with_occ, debug)
            self._buffers[key] = buf
        return buf

    def generate_bev_device(self, source, points, ground_plane, area_extents, voxel_size,
                            with_occupancy=False, occ_lo=0.2, occ_hi=2.0, buffers=None,
                            debug=False):
        """points: (3, N) CUDA tensor. Returns the BevBuffers holding maps [(S+1), nz, nx] (height
        slices then density), the occupancy grid and the stats block; nothing is synchronised."""
        params = ops.make_bev_params(ground_plane, area_extents, voxel_size, self.height_lo,
                                     self.height_hi, self.num_slices, True, occ_lo, occ_hi,
                                     self.NORM_VALUES[source])
        nx, _, nz, _, _, _ = ops.bev_grid(area_extents, voxel_size)
        buf = buffers or self.buffers(nx, nz, points.shape[1], with_occupancy, debug)
        ops.bev_slices(points, params, buf.maps, buf.occ if with_occupancy else None, buf.stats,
                       buf.workspace, buf.winner, buf.counts)
        return buf

    # -- reference API ------------------------------------------------------------------------
    def generate_bev(self, source, point_cloud, ground_plane, area_extents, voxel_size,
                     strict=None):
        """Generates the BEV maps dictionary (reference: bev_slices.py:33-150).

        Args and return value as the reference: point_cloud (3, N), ground_plane [a, b, c, d],
        area_extents [[min_x, max_x], [min_y, max_y], [min_z, max_z]], voxel_size in m ->
        {'height_maps': [num_slices x (H, W)], 'density_map': (H, W)}. NumPy in -> NumPy out,
        CUDA tensor in -> CUDA tensors out (views of one [S+1, H, W] buffer, valid until the next
        call)."""
        points, was_numpy = _to_device_points(point_cloud, self.device)
        strict = was_numpy if strict is None else strict
        buf = self.generate_bev_device(source, points, ground_plane, area_extents, voxel_size)
        if strict:
            stats = buf.stats.cpu().numpy()
            _raise_like_reference(stats, self.num_slices)
        maps = buf.maps
        if was_numpy:
            host = maps.cpu().numpy()
            return {'height_maps': [host[i] for i in range(self.num_slices)],
                    'density_map': host[self.num_slices]}
        return {'height_maps': [maps[i] for i in range(self.num_slices)],
                'density_map': maps[self.num_slices]}

    def generate_bev_and_voxel_grid(self, source, point_cloud, ground_plane, area_extents,
                                    voxel_size, occ_lo=0.2, occ_hi=2.0):
        """generate_bev plus the VoxelGrid2D of kitti_utils.py:268-277
        (create_sliced_voxel_grid_2d_v2) from the same pass over the points."""
        points, was_numpy = _to_device_points(point_cloud, self.device)
        buf = self.generate_bev_device(source, points, ground_plane, area_extents, voxel_size,
                                       with_occupancy=True, occ_lo=occ_lo, occ_hi=occ_hi)
        grid = VoxelGrid2D.from_occupancy(buf.occ, voxel_size, area_extents)
        maps = buf.maps.cpu().numpy() if was_numpy else buf.maps
        bev = {'height_maps': [maps[i] for i in range(self.num_slices)],
               'density_map': maps[self.num_slices]}
        return bev, grid


def _raise_like_reference(stats, num_slices):
    if stats[STAT_OVERFLOW]:
        raise MemoryError("BEV touched-cell list overflowed its workspace")
    if stats[STAT_OOB]:
        # voxel_grid_2d.py:133-138
        raise ValueError("Extents are smaller than min_voxel_coord / max_voxel_coord")
    if stats[STAT_DENSITY] == 0:
        # voxel_grid_2d.py:100-102: unique_indices[-1] on an empty array
        raise IndexError("index -1 is out of bounds for axis 0 with size 0 "
                         "(no point inside the density slice)")
