"""Drop-in for avod/core/anchor_filter.py:64-119 (get_empty_anchor_filter_2d) on sm_100a kernels.

Same signature and result: anchors N x [x, y, z, dim_x, dim_y, dim_z], a voxel grid and an integer
density threshold -> boolean keep mask (N,). The integral image
(wavedata/.../integral_image_2d.py:17-37), the float32 corner rounding and index truncation
(anchor_filter.py:93-108, voxel_grid_2d.py:182-184) and the box-sum query
(integral_image_2d.py:71-85) all run on the device.

`voxel_grid_2d` may be a dodt_b200.VoxelGrid2D (occupancy already on the GPU — the normal case,
produced by the BEV pass) or any object with the reference's attributes (`leaf_layout_2d`,
`voxel_size`, `min_voxel_coord`, `num_divisions`), e.g. wavedata's own VoxelGrid2D, whose
leaf layout is uploaded.
"""
import numpy as np
import torch

from . import ops
from .voxel_grid_2d import VoxelGrid2D


def check_anchor_format(anchors):
    """avod/core/format_checker.py:81-105: TypeError unless N x 6 (or a single anchor of 6)."""
    shape = tuple(anchors.shape)
    if len(shape) == 2 and shape[1] != 6:
        raise TypeError('Given input does not have valid number of attributes. '
                        'Should be N x 6 for anchor.')
    if len(shape) == 1 and shape[0] != 6:
        raise TypeError('Given input does not have valid number of attributes. '
                        'Should be 6 for anchor.')
    if len(shape) not in (1, 2):
        raise TypeError('Given input is not of valid types.(i.e. np.ndarray or tf.Tensor)')


def _device_grid(voxel_grid_2d, device):
    if isinstance(voxel_grid_2d, VoxelGrid2D):
        return voxel_grid_2d
    layout = np.squeeze(np.asarray(voxel_grid_2d.leaf_layout_2d) + 1)
    if layout.ndim != 2:
        raise ValueError('Not a 2D image for integral image: input dim {}'.format(layout.ndim))
    grid = VoxelGrid2D(device)
    grid.voxel_size = voxel_grid_2d.voxel_size
    grid.min_voxel_coord = np.asarray(voxel_grid_2d.min_voxel_coord, dtype=np.float64)
    grid.num_divisions = np.asarray(voxel_grid_2d.num_divisions)
    grid.occ = torch.from_numpy(np.ascontiguousarray(layout != 0).astype(np.uint8)).to(grid.device)
    return grid


def get_empty_anchor_filter_2d(anchors, voxel_grid_2d, density_threshold=1):
    """Returns a filter for empty anchors from the given 2D anchor list.

    NumPy anchors -> NumPy bool mask; CUDA tensor anchors -> CUDA bool tensor (no sync)."""
    was_numpy = not torch.is_tensor(anchors)
    a = torch.from_numpy(np.ascontiguousarray(anchors)) if was_numpy else anchors
    check_anchor_format(a)
    if a.dim() == 1:
        a = a.reshape(1, 6)
    if a.dtype not in (torch.float32, torch.float64):
        a = a.double()
    device = a.device if a.is_cuda else torch.device("cuda")
    grid = _device_grid(voxel_grid_2d, device)
    a = a.to(grid.device, non_blocking=True)
    nx, nz = int(grid.num_divisions[0]), int(grid.num_divisions[2])
    keep = ops.anchor_filter_2d(a, grid.integral_image(), nx, nz, int(grid.min_voxel_coord[0]),
                                int(grid.min_voxel_coord[2]), grid.voxel_size, density_threshold)
    if was_numpy:
        return keep.cpu().numpy().astype(bool)
    return keep.bool()
