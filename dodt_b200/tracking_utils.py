"""File-level drop-ins for the LiDAR ingest of DODT's tracking data set (SURVEY 8(f) rank 2 and the
on-disk formats either side of it), same names and call signatures as the reference:

  read_lidar(velo_dir, img_idx)                      wavedata/.../core/calib_utils.py:441-480
  read_tracking_calibration(calib_dir, video_id)     wavedata/.../core/calib_utils.py:155-214
  get_raw_lidar_point_cloud(name, velo_dir)          wavedata/.../obj_detection/tracking_utils.py:108-114
  get_lidar_in_camera_view(pts, name, calib_dir, im_size=None, min_intensity=None)      :115-148
  get_lidar_point_cloud(name, calib_dir, velo_dir, im_size=None, min_intensity=None)    :152-203
  get_road_plane(name, planes_dir)                   :205-248 (host; DODT fixes the plane)
  get_oxts(oxts_dir, sample_name)                    avod/datasets/kitti/kitti_tracking_dataset.py:215-223
  get_pair_point_clouds(sample_names, ...)           kitti_tracking_dataset.py:485-494 (load_samples):
                                                     raw scans -> frame t+tau moved into frame t's LiDAR
                                                     frame -> both rectified and frustum-cropped

Files are read on the host exactly as the reference reads them (np.fromfile of little-endian
float32 [n, 4]; space-separated calibration rows; one oxts line per frame); everything per point runs
on the device through dodt_b200.lidar -> dodt_lidar_to_camera[_aligned]. Frame names are the
reference's six characters: two digits of video id, four digits of frame id.
"""
import os
from types import SimpleNamespace

import numpy as np

from . import lidar


def _ids(name):
    assert len(name) == 6, print('Sample name incorrect!')
    return int(name[:2]), int(name[2:])


def read_lidar(velo_dir, img_idx):
    """(x, y, z, i) float32 arrays of one velodyne scan, or [] if the file does not exist."""
    path = velo_dir + "/%06d.bin" % img_idx
    if not os.path.exists(path):
        return []
    xyzi = np.fromfile(path, np.single).reshape(-1, 4)
    return xyzi[:, 0], xyzi[:, 1], xyzi[:, 2], xyzi[:, 3]


def read_tracking_calibration(calib_dir, video_id):
    """Calibration of one tracking video: .p0 .. .p3 (3x4), .r0_rect (3x3), .tr_velodyne_to_cam (3x4).
    Rows are 'Name: v v v ...' in the order P0, P1, P2, P3, R_rect, Tr_velo_cam; runs of spaces are
    separators (the reference strips the empty csv fields)."""
    with open(calib_dir + "/%04d.txt" % video_id, 'r') as f:
        rows = [[t for t in line.rstrip('\n').split(' ') if t != ''] for line in f]
    vals = [[float(v) for v in r[1:]] for r in rows[:6]]
    return SimpleNamespace(p0=np.reshape(vals[0], (3, 4)), p1=np.reshape(vals[1], (3, 4)),
                           p2=np.reshape(vals[2], (3, 4)), p3=np.reshape(vals[3], (3, 4)),
                           r0_rect=np.reshape(vals[4], (3, 3)), tr_velodyne_to_cam=np.reshape(vals[5], (3, 4)))


def get_raw_lidar_point_cloud(name, velo_dir):
    """(4, N) float32 rows x, y, z, intensity of the frame's raw scan (LiDAR frame)."""
    video_id, frame_id = _ids(name)
    x, y, z, i = read_lidar(velo_dir=velo_dir + '/' + str(video_id).zfill(4), img_idx=frame_id)
    return np.vstack((x, y, z, i))


def get_lidar_in_camera_view(pts, name, calib_dir, im_size=None, min_intensity=None, ego=None):
    """pts (4, N) raw scan rows -> (3, M) float64 points in the rectified camera frame; with
    im_size = [w, h] only the points in front of the camera that project strictly inside the image.
    ego = (trans, matrix): move the scan first (dodt_b200.lidar.point_cloud_transform fused in)."""
    video_id, _ = _ids(name)
    frame_calib = read_tracking_calibration(calib_dir, video_id)
    velo = np.ascontiguousarray(np.asarray(pts, dtype=np.float32).T)
    if velo.shape[1] == 3:
        velo = np.concatenate([velo, np.zeros((len(velo), 1), np.float32)], axis=1)
    if not im_size:
        return lidar.get_lidar_in_camera_view(velo, frame_calib, ego=ego)
    if min_intensity:
        # tracking_utils.py:143-148 combines the image filter (one flag per point in FRONT of the
        # camera) with an intensity filter over ALL points: NumPy raises unless no point lies behind
        # the camera. DODT never passes min_intensity; the same error is reported here.
        raise ValueError("operands could not be broadcast together: the reference's intensity filter "
                         "is defined over all points, its image filter over those with z > 0")
    return lidar.get_lidar_in_camera_view(velo, frame_calib, im_size=im_size, ego=ego)


def get_lidar_point_cloud(name, calib_dir, velo_dir, im_size=None, min_intensity=None):
    """The frame's scan in the rectified camera frame, optionally cropped to the image frustum:
    (3, M) float64 like the reference."""
    return get_lidar_in_camera_view(get_raw_lidar_point_cloud(name, velo_dir), name, calib_dir,
                                    im_size=im_size, min_intensity=min_intensity)


def get_road_plane(name, planes_dir):
    """Ground plane (a, b, c, d) of a frame. The reference parses the 4th line of the frame's plane
    file when it exists and then OVERRIDES the result for the tracking data sets with the fixed plane
    [0, -1, 0, 1.65] (tracking_utils.py:232-236), normal facing up (+y is down), normalised: that
    fixed plane is what every caller gets, whatever the file holds. A malformed existing file still
    raises like the reference's parse does."""
    video_id, frame_id = _ids(name)
    plane_file = planes_dir + '/%04d/%06d.txt' % (video_id, frame_id)
    if os.path.exists(plane_file):
        with open(plane_file, 'r') as f:
            [float(v) for v in f.readlines()[3].split()]
    plane = np.asarray([0, -1, 0, 1.65])
    if plane[1] > 0:
        plane = -plane
    return plane / np.linalg.norm(plane[0:3])


def get_oxts(oxts_dir, sample_name):
    """The frame's GPS/IMU record: line frame_id of <oxts_dir>/<video id>.txt."""
    video_id, frame_id = _ids(sample_name)
    with open(oxts_dir + '/%04d.txt' % video_id) as f:
        lines = [line.rstrip() for line in f.readlines()]
    return lidar.Oxts(lines[frame_id])


def get_pair_point_clouds(sample_names, calib_dir, velo_dir, oxts_dir, image_shapes):
    """The two clouds of a DODT sample [name_t, name_t+tau] as load_samples builds them: raw scans,
    the second moved into the first one's LiDAR frame with the OXTS records, both taken to the
    camera view of their image (image_shapes: (h, w) per frame). -> [(3, M0), (3, M1)] float64."""
    assert sample_names[0][:2] == sample_names[1][:2], print("sample couple from different video!")
    trans, matrix, _ = lidar.coordinate_transform(get_oxts(oxts_dir, sample_names[0]),
                                                  get_oxts(oxts_dir, sample_names[1]))
    out = []
    for k, name in enumerate(sample_names):
        raw = get_raw_lidar_point_cloud(name, velo_dir)
        out.append(get_lidar_in_camera_view(raw, name, calib_dir,
                                            im_size=[image_shapes[k][1], image_shapes[k][0]],
                                            ego=(trans, matrix) if k == len(sample_names) - 1 else None))
    return out
