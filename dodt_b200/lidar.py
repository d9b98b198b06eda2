"""Drop-ins for the LiDAR ingest helpers of wavedata on sm_100a kernels:

  lidar_to_cam_frame(xyz_lidar, frame_calib)                 wavedata/.../core/calib_utils.py:484-523
  get_lidar_in_camera_view(velo, frame_calib, im_size=None)  wavedata/.../obj_detection/tracking_utils.py:115-148,152-203

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


def get_lidar_in_camera_view(velo, frame_calib, im_size=None, dtype=torch.float64):
    """velo: N x 4 raw scan rows (x, y, z, intensity). Returns the (3, M) camera-frame cloud of the
    points in front of the camera that project strictly inside an image of im_size = [w, h]
    (all points if im_size is None). NumPy input: NumPy (3, M) float64. CUDA tensor input:
    (points (3, N) with the first M columns valid, count [1] int32 = M) on the device."""
    t, was_numpy = _velo_tensor(velo)
    pts, count = ops.lidar_to_camera(t, rectified_matrix(frame_calib),
                                     None if im_size is None else np.asarray(frame_calib.p2, dtype=np.float64),
                                     im_size, dtype=dtype)
    if was_numpy:
        m = int(count.item())
        return pts[:, :m].cpu().numpy()
    return pts, count
