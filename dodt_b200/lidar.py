"""Drop-ins for the LiDAR ingest helpers of wavedata on sm_100a kernels:

  lidar_to_cam_frame(xyz_lidar, frame_calib)                 wavedata/.../core/calib_utils.py:484-523
  get_lidar_in_camera_view(velo, frame_calib, im_size=None)  wavedata/.../obj_detection/tracking_utils.py:115-148,152-203
  Oxts, coordinate_transform(oxts_cur, oxts_next)            avod/datasets/kitti/kitti_tracking_utils.py:129-216,
                                                             kitti_tracking_dataset.py:300-315 (host scalars)
  point_cloud_transform(point_cloud, trans, matrix)          kitti_tracking_dataset.py:317-328: frame t+tau's
                                                             scan in frame t's LiDAR coordinates

`frame_calib` is duck-typed like wavedata's FrameCalibrationData: `.r0_rect` (3x3),
`.tr_velodyne_to_cam` (3x4), `.p2` (3x4). The 4x4 products of the calibration matrices are formed
on the host with NumPy exactly as the reference does; the per-point work runs on the device.
NumPy in -> NumPy out (float64, like the reference); CUDA tensors in -> (points, count) on the
device without synchronisation, ready for BevSlices / dodt_bev_slices (n_dev).
"""
import numpy as np
import torch

from . import ops


def rectified_matrix(frame_calib):
    """R0_rect (padded to 4x4) . Tr_velo_to_cam (padded to 4x4), calib_utils.py:503-519."""
    r0 = np.pad(np.asarray(frame_calib.r0_rect, dtype=np.float64), ((0, 1), (0, 1)), 'constant')
    r0[3, 3] = 1
    tf = np.pad(np.asarray(frame_calib.tr_velodyne_to_cam, dtype=np.float64), ((0, 1), (0, 0)), 'constant')
    tf[3, 3] = 1
    return np.dot(r0, tf)


def _velo_tensor(velo):
    was_numpy = not torch.is_tensor(velo)
    t = torch.from_numpy(np.ascontiguousarray(velo, dtype=np.float32)) if was_numpy else velo
    if t.dim() != 2 or t.shape[1] not in (3, 4):
        raise ValueError("expected N x 3 (xyz) or N x 4 (xyz, intensity) points, got {}".format(tuple(t.shape)))
    t = t.to(torch.float32)
    if t.shape[1] == 3:
        t = torch.cat([t, torch.zeros_like(t[:, :1])], dim=1)
    return t.cuda().contiguous() if not t.is_cuda else t.contiguous(), was_numpy


def lidar_to_cam_frame(xyz_lidar, frame_calib):
    """N x 3 lidar points -> N x 3 points in the rectified camera-0 frame."""
    velo, was_numpy = _velo_tensor(xyz_lidar)
    pts, count = ops.lidar_to_camera(velo, rectified_matrix(frame_calib))
    out = pts[:, :velo.shape[0]].t()
    return out.cpu().numpy() if was_numpy else out


class Oxts:
    """One GPS/IMU record of a KITTI tracking oxts file (30 values per line; the first six are
    used): kitti_tracking_utils.py:129-216. Host scalar arithmetic, NumPy like the reference."""
    EARTH_RADIUS = 6378137.0

    def __init__(self, oxts_line):
        v = [float(t) for t in oxts_line.split()[:6]]
        self.latitude, self.longitude, self.altitude, self.roll, self.pitch, self.yaw = v

    def distance(self, other):
        """Haversine distance in metres (:169-187)."""
        lat1, lon1 = self.latitude * np.pi / 180.00, self.longitude * np.pi / 180.00
        lat2, lon2 = other.latitude * np.pi / 180.00, other.longitude * np.pi / 180.00
        a, b = lat2 - lat1, lon2 - lon1
        return abs(2 * self.EARTH_RADIUS * np.arcsin(np.sqrt(
            np.power(np.sin(a / 2), 2) + np.cos(lat1) * np.cos(lat2) * np.power(np.sin(b / 2), 2))))

    def displacement(self, other):
        """:189-196"""
        d = self.distance(other)
        dyaw, dpitch = self.yaw - other.yaw, self.pitch - other.pitch
        return np.array([d * np.cos(dyaw), d * np.sin(dyaw), d * np.sin(dpitch)])

    def get_delta(self, other, theta='yaw'):
        return getattr(self, theta) - getattr(other, theta)

    def get_rotate_matrix(self, other, axis='y'):
        """:141-167,198-207 — note the reference's naming: 'z' rotates by the pitch difference in the
        x-z plane, 'x' by the roll difference in the y-z plane, 'y' by the yaw difference in x-y."""
        t = self.get_delta(other, {'z': 'pitch', 'x': 'roll', 'y': 'yaw'}[axis])
        c, s = np.cos(t), np.sin(t)
        return {'z': np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]]),
                'x': np.array([[1, 0, 0], [0, c, -s], [0, s, c]]),
                'y': np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])}[axis]


def coordinate_transform(oxts_cur, oxts_next):
    """(translation [3], rotation [3, 3], yaw difference) that express frame t+tau in frame t's
    coordinates: kitti_tracking_dataset.py:300-315. Arguments: Oxts records (or their text lines)."""
    cur = oxts_cur if hasattr(oxts_cur, "displacement") else Oxts(oxts_cur)
    nxt = oxts_next if hasattr(oxts_next, "yaw") else Oxts(oxts_next)
    matrix = cur.get_rotate_matrix(nxt, 'z') @ cur.get_rotate_matrix(nxt, 'x') @ cur.get_rotate_matrix(nxt, 'y')
    return cur.displacement(nxt), matrix, cur.get_delta(nxt, theta='yaw')


def point_cloud_transform(point_cloud, trans, matrix):
    """kitti_tracking_dataset.py:317-328 on one scan: point_cloud (4, N) or (3, N) rows x, y, z
    [, intensity] of frame t+tau (float32, as calib_utils.read_lidar delivers them) -> the same
    layout with xyz <- float32((xyz + trans) @ matrix). NumPy in -> NumPy out; CUDA in -> CUDA out."""
    was_numpy = not torch.is_tensor(point_cloud)
    pc = torch.from_numpy(np.ascontiguousarray(point_cloud, dtype=np.float32)) if was_numpy else point_cloud
    if pc.dim() != 2 or pc.shape[0] not in (3, 4):
        raise ValueError("expected a (3, N) or (4, N) point cloud, got {}".format(tuple(pc.shape)))
    velo, _ = _velo_tensor(pc.t())
    aligned = ops.lidar_align(velo, trans, matrix)
    out = aligned[:, :pc.shape[0]].t().contiguous()
    return out.cpu().numpy() if was_numpy else out


def get_lidar_in_camera_view(velo, frame_calib, im_size=None, dtype=torch.float64, ego=None):
    """velo: N x 4 raw scan rows (x, y, z, intensity). Returns the (3, M) camera-frame cloud of the
    points in front of the camera that project strictly inside an image of im_size = [w, h]
    (all points if im_size is None). NumPy input: NumPy (3, M) float64. CUDA tensor input:
    (points (3, N) with the first M columns valid, count [1] int32 = M) on the device.
    ego = (trans, matrix) of coordinate_transform: the scan is frame t+tau's and is first moved into
    frame t's LiDAR frame (point_cloud_transform fused into the same pass)."""
    t, was_numpy = _velo_tensor(velo)
    pts, count = ops.lidar_to_camera(t, rectified_matrix(frame_calib),
                                     None if im_size is None else np.asarray(frame_calib.p2, dtype=np.float64),
                                     im_size, dtype=dtype, ego=ego)
    if was_numpy:
        m = int(count.item())
        return pts[:, :m].cpu().numpy()
    return pts, count
