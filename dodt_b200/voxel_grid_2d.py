"""Device-resident drop-in for wavedata/wavedata/tools/core/voxel_grid_2d.py (VoxelGrid2D).

Keeps the reference's attribute names (voxel_size, min_voxel_coord, max_voxel_coord,
num_divisions, voxel_indices, heights, num_pts_in_voxel, leaf_layout_2d) and methods
(voxelize_2d, map_to_index). The occupancy lives on the GPU as a uint8 [nx, nz] grid
(1 = VOXEL_FILLED); the NumPy views the reference exposes are materialised lazily, so the
anchor filter (dodt_b200.anchor_filter) never leaves the device.
"""
import numpy as np
import torch

from . import ops
from ._lib import BEV_STATS_LEN, STAT_OOB


class VoxelGrid2D(object):
    VOXEL_EMPTY = -1
    VOXEL_FILLED = 0

    def __init__(self, device=None):
        self.voxel_size = 0.0
        self.min_voxel_coord = np.array([])
        self.max_voxel_coord = np.array([])
        self.num_divisions = np.array([0, 0, 0])
        self.points = []
        self.device = torch.device("cuda" if device is None else device)
        self.occ = None           # CUDA uint8 [nx, nz]
        self._ii = None           # cached integral image
        self._maps = None         # CUDA f32 [2, nz, nx]: raw heights, (unused density)
        self._counts = None       # CUDA i32 [nz, nx]
        self._winner = None       # CUDA i32 [1, nz, nx]
        self._sparse = None

    # ---- construction from the fused BEV pass -----------------------------------------------
    @classmethod
    def from_occupancy(cls, occ, voxel_size, area_extents):
        """Wrap an occupancy grid produced by dodt_bev_slices (S1) for the same extents."""
        self = cls(occ.device)
        self.voxel_size = voxel_size
        nx, _, nz, min_x, _, min_z = ops.bev_grid(area_extents, voxel_size)
        self.min_voxel_coord = np.array([float(min_x), 0.0, float(min_z)])
        self.max_voxel_coord = np.array([float(min_x + nx - 1), 0.0, float(min_z + nz - 1)])
        self.num_divisions = np.array([nx, 1, nz], dtype=np.int32)
        self.occ = occ
        return self

    # ---- reference API ------------------------------------------------------------------------
    def voxelize_2d(self, pts, voxel_size, extents=None, ground_plane=None,
                    create_leaf_layout=True):
        """Voxelises N x [x, y, z] points (reference: voxel_grid_2d.py:43-160): per (x, z) cell the
        number of points and the height of the cell's first point in (y-bin, input order)."""
        if torch.is_tensor(pts):
            p = pts if pts.dtype in (torch.float32, torch.float64) else pts.double()
        else:
            arr = np.asarray(pts)
            p = torch.from_numpy(np.ascontiguousarray(arr, dtype=arr.dtype if arr.dtype in
                                                      (np.float32, np.float64) else np.float64))
        if p.dim() != 2 or p.shape[1] != 3:
            raise ValueError("Points have the wrong shape: {}".format(tuple(p.shape)))
        p = p.to(self.device)
        soa = p.t().contiguous()                      # (3, N) as the kernel wants it
        self.voxel_size = voxel_size
        n = soa.shape[1]
        if n == 0:
            raise IndexError("index -1 is out of bounds for axis 0 with size 0")
        # extents of the data in voxel units (plumbing reductions; the binning itself is in the kernel)
        lo = soa.amin(dim=1).double().cpu().numpy()
        hi = soa.amax(dim=1).double().cpu().numpy()
        data_min = np.floor(lo / voxel_size)
        data_max = np.floor(hi / voxel_size)
        if extents is not None:
            ext_t = np.array(extents).transpose()
            if ext_t.shape != (2, 3):
                raise ValueError("Extents are the wrong shape {}".format(np.shape(extents)))
            self.min_voxel_coord = np.floor(ext_t[0] / voxel_size)
            self.max_voxel_coord = np.ceil((ext_t[1] / voxel_size) - 1)
            self.min_voxel_coord[1] = 0
            self.max_voxel_coord[1] = 0
            if not (self.min_voxel_coord[[0, 2]] <= data_min[[0, 2]]).all():
                raise ValueError("Extents are smaller than min_voxel_coord")
            if not (self.max_voxel_coord[[0, 2]] >= data_max[[0, 2]]).all():
                raise ValueError("Extents are smaller than max_voxel_coord")
        else:
            self.min_voxel_coord = np.array([data_min[0], 0.0, data_min[2]])
            self.max_voxel_coord = np.array([data_max[0], 0.0, data_max[2]])
        self.num_divisions = ((self.max_voxel_coord - self.min_voxel_coord) + 1).astype(np.int32)
        nx, nz = int(self.num_divisions[0]), int(self.num_divisions[2])
        vs = float(voxel_size)
        # a grid that reproduces exactly [min, max] under floor(min/v), ceil(max/v - 1); the y range
        # only has to bound the data (it orders points inside a cell)
        grid_ext = [(self.min_voxel_coord[0] + 0.5) * vs, (self.max_voxel_coord[0] + 0.5) * vs,
                    (data_min[1] + 0.5) * vs, (data_max[1] + 0.5) * vs,
                    (self.min_voxel_coord[2] + 0.5) * vs, (self.max_voxel_coord[2] + 0.5) * vs]
        got = ops.bev_grid(grid_ext, vs)
        want = (nx, int(data_max[1] - data_min[1]) + 1, nz, int(self.min_voxel_coord[0]),
                int(data_min[1]), int(self.min_voxel_coord[2]))
        if got != want:
            raise RuntimeError("internal: voxel grid extents did not round-trip %r != %r" % (got, want))
        params = ops.make_bev_params(ground_plane, grid_ext, vs, 0.0, 1.0, 1, filter_mode=False)
        dev = self.device
        self._maps = torch.empty((2, nz, nx), dtype=torch.float32, device=dev)
        self._counts = torch.empty((nz, nx), dtype=torch.int32, device=dev)
        self._winner = torch.empty((1, nz, nx), dtype=torch.int32, device=dev)
        self.occ = torch.empty((nx, nz), dtype=torch.uint8, device=dev)
        stats = torch.empty((BEV_STATS_LEN,), dtype=torch.int32, device=dev)
        ws = torch.empty(max(ops.bev_workspace_bytes(n, 1, nx, nz), 256), dtype=torch.uint8, device=dev)
        ops.bev_slices(soa, params, self._maps, self.occ, stats, ws, self._winner, self._counts)
        if int(stats[STAT_OOB].item()) != 0:
            raise ValueError("Extents are smaller than min/max_voxel_coord")
        self._points_soa = soa
        self._sparse = None
        self._ii = None

    def _materialise(self):
        """voxel_indices / heights / num_pts_in_voxel in the reference's order (ascending x, z)."""
        if self._sparse is not None:
            return self._sparse
        if self._counts is None:
            raise AttributeError("this VoxelGrid2D was built from an occupancy grid only")
        nx, nz = int(self.num_divisions[0]), int(self.num_divisions[2])
        counts = self._counts.cpu().numpy()          # [nz, nx], rotated: row r = iz nz-1-r
        winner = self._winner[0].cpu().numpy()
        grid_counts = np.flip(counts, axis=0).transpose()     # [nx, nz]
        grid_winner = np.flip(winner, axis=0).transpose()
        ix, iz = np.nonzero(grid_counts)                      # row-major => ascending (x, z)
        pts = self._points_soa.t().cpu().numpy()
        first = grid_winner[ix, iz]
        self._sparse = dict(
            voxel_indices=np.stack([ix, np.zeros_like(ix), iz], axis=1).astype(int),
            num_pts_in_voxel=grid_counts[ix, iz].astype(np.int64),
            first=first, first_points=pts[first])
        return self._sparse

    @property
    def voxel_indices(self):
        return self._materialise()["voxel_indices"]

    @property
    def num_pts_in_voxel(self):
        return self._materialise()["num_pts_in_voxel"]

    @property
    def heights(self):
        sp = self._materialise()
        nz = int(self.num_divisions[2])
        h = self._maps[0].cpu().numpy()
        vi = sp["voxel_indices"]
        return h[nz - 1 - vi[:, 2], vi[:, 0]]

    @property
    def leaf_layout_2d(self):
        """-1 empty / 0 filled, float64 (nx, 1, nz) as voxel_grid_2d.py:152-160."""
        occ = self.occ.cpu().numpy().astype(np.float64)
        return (occ - 1.0).reshape(occ.shape[0], 1, occ.shape[1])

    def map_to_index(self, map_index):
        """Map coordinates -> clipped grid indices (reference: voxel_grid_2d.py:162-186)."""
        if self.voxel_size == 0 or len(self.min_voxel_coord) == 0 or len(map_index) == 0:
            return []
        was_numpy = not torch.is_tensor(map_index)
        t = torch.as_tensor(np.asarray(map_index)) if was_numpy else map_index
        if t.dtype not in (torch.float32, torch.float64):
            t = t.double()
        out = ops.map_to_index(t.to(self.device), self.voxel_size, int(self.min_voxel_coord[0]),
                               int(self.min_voxel_coord[2]), int(self.num_divisions[0]),
                               int(self.num_divisions[2]))
        return out.cpu().numpy().astype(np.float64) if was_numpy else out

    # ---- used by the anchor filter -------------------------------------------------------------
    def integral_image(self):
        if self._ii is None:
            self._ii = ops.integral_image_2d(self.occ)
        return self._ii
