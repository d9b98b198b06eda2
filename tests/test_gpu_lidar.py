"""GPU parity of the LiDAR ingest: raw KITTI scan -> camera-frame frustum
cloud, against the reference's output frozen in tests/golden/ and the oracle restatement; and the
device-resident chain ingest -> BEV maps without a host round trip."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import synth_ref as S
from oracle import np_oracle as O

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def frame(lib):
    g = np.load(os.path.join(GOLDEN, "lidar_kitti_000003.npz"))
    calib = SimpleNamespace(r0_rect=g["r0_rect"], tr_velodyne_to_cam=g["tr_velodyne_to_cam"], p2=g["p2"])
    return g, calib


def test_lidar_in_camera_view_equals_reference(frame):
    from dodt_b200 import lidar
    g, calib = frame
    fov = lidar.get_lidar_in_camera_view(g["velo"], calib, im_size=list(g["im_size"]))
    assert fov.dtype == np.float64 and fov.shape == g["fov"].shape     # the same points were kept
    np.testing.assert_allclose(fov, g["fov"], rtol=1e-12, atol=1e-12)  # np.dot order is BLAS-defined
    full = lidar.get_lidar_in_camera_view(g["velo"], calib)
    assert full.shape == (3, len(g["velo"]))
    np.testing.assert_allclose(full[:, ::7], g["full_every_7th"], rtol=1e-12, atol=1e-12)
    cam = lidar.lidar_to_cam_frame(g["velo"][:, :3], calib)
    np.testing.assert_allclose(cam, O.lidar_to_cam_frame(g["velo"][:, :3].astype(np.float64), g["r0_rect"],
                                                         g["tr_velodyne_to_cam"]), rtol=1e-12, atol=1e-12)


def test_lidar_edge_cases(frame):
    from dodt_b200 import lidar, ops
    g, calib = frame
    rng = np.random.default_rng(2)
    velo = np.concatenate([g["velo"][:5000], rng.uniform(-80, 80, (3000, 4)).astype(np.float32)])
    for im in ([1242, 375], [10, 10], [4000, 4000]):
        want = O.lidar_in_camera_view(velo, g["r0_rect"], g["tr_velodyne_to_cam"], g["p2"], im)
        got = lidar.get_lidar_in_camera_view(velo, calib, im_size=im)
        assert got.shape == want.shape
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12)
    empty = lidar.get_lidar_in_camera_view(np.zeros((0, 4), np.float32), calib, im_size=[1242, 375])
    assert empty.shape == (3, 0)
    with pytest.raises(ValueError):
        lidar.get_lidar_in_camera_view(np.zeros((7, 5), np.float32), calib)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.lidar_to_camera(torch.zeros(4, 4), np.eye(4))


def test_ego_motion_alignment_equals_reference(lib):
    """The frame-pair ingest: frame t+1's scan moved into frame t's LiDAR frame (OXTS records), then
    rectified and frustum-cropped, in one pass on the device == the outputs of the reference's
    point_cloud_transform + get_lidar_in_camera_view frozen in tests/golden/. The moved scan is
    float32((xyz + trans) @ matrix); np's matmul order is BLAS-defined, so a value may round the other
    way when its float64 result lies within 1e-16 of a float32 tie: at most 1 float32 ulp, almost never."""
    from dodt_b200 import lidar
    g = np.load(os.path.join(GOLDEN, "lidar_pair_000003_000004.npz"))
    calib = SimpleNamespace(r0_rect=g["r0_rect"], tr_velodyne_to_cam=g["tr_velodyne_to_cam"], p2=g["p2"])
    trans, matrix, _ = lidar.coordinate_transform(str(g["oxts_line0"]), str(g["oxts_line1"]))
    moved = lidar.point_cloud_transform(g["velo1"].T, trans, matrix)
    assert moved.dtype == np.float32 and moved.shape == g["moved"].shape
    assert (moved != g["moved"]).mean() < 1e-5
    np.testing.assert_allclose(moved, g["moved"], rtol=1.2e-7, atol=0)
    np.testing.assert_array_equal(lidar.point_cloud_transform(g["velo1"].T[:3], trans, matrix), moved[:3])
    fov = lidar.get_lidar_in_camera_view(g["velo1"], calib, im_size=list(g["im_size"]), ego=(trans, matrix))
    assert fov.shape == g["fov"].shape
    if np.array_equal(moved, g["moved"]):
        np.testing.assert_allclose(fov, g["fov"], rtol=1e-12, atol=1e-12)
    else:
        np.testing.assert_allclose(fov, g["fov"], rtol=1e-6, atol=1e-5)
    # device-resident: the count stays on the GPU, the points feed dodt_bev_slices(n_dev) directly
    pts, count = lidar.get_lidar_in_camera_view(torch.from_numpy(g["velo1"]).cuda(), calib,
                                                im_size=list(g["im_size"]), ego=(trans, matrix))
    assert int(count.item()) == g["fov"].shape[1]
    np.testing.assert_array_equal(pts[:, :int(count.item())].cpu().numpy(), fov)
    # against the oracle on other inputs
    rng = np.random.default_rng(5)
    velo = rng.uniform(-70, 70, (4000, 4)).astype(np.float32)
    t2, m2, _ = O.oxts_coordinate_transform([49.0, 8.4, 115, 0.01, -0.02, 1.0], [49.00001, 8.40002, 115, 0.02, -0.01, 0.97])
    want = O.point_cloud_transform(velo.T, t2, m2)
    got = lidar.point_cloud_transform(velo.T, t2, m2)
    np.testing.assert_allclose(got, want, rtol=1.2e-7, atol=0)
    assert (got != want).mean() < 1e-3


def test_file_level_drop_ins_equal_reference(frame, tmp_path):
    """wavedata's tracking_utils.get_lidar_point_cloud / get_lidar_in_camera_view and the frame-pair
    ingest of KittiTrackingDataset.load_samples, called by file name like the reference, on a KITTI
    directory rebuilt from the committed fixtures == the reference's frozen outputs."""
    from dodt_b200 import tracking_utils as T
    g, _ = frame
    gp = np.load(os.path.join(GOLDEN, "lidar_pair_000003_000004.npz"))
    mini = os.path.join(GOLDEN, "kitti_mini")
    velo_dir = str(tmp_path / "velodyne")
    os.makedirs(velo_dir + "/0000")
    g["velo"].tofile(velo_dir + "/0000/000003.bin")
    gp["velo1"].tofile(velo_dir + "/0000/000004.bin")
    fov = T.get_lidar_point_cloud("000003", mini + "/calib", velo_dir, im_size=[1242, 375])
    assert fov.shape == g["fov"].shape and fov.dtype == np.float64
    np.testing.assert_allclose(fov, g["fov"], rtol=1e-12, atol=1e-12)
    full = T.get_lidar_point_cloud("000003", mini + "/calib", velo_dir)
    np.testing.assert_allclose(full[:, ::7], g["full_every_7th"], rtol=1e-12, atol=1e-12)
    pair = T.get_pair_point_clouds(["000003", "000004"], mini + "/calib", velo_dir, mini + "/oxts",
                                   [(375, 1242), (375, 1242)])
    np.testing.assert_allclose(pair[0], g["fov"], rtol=1e-12, atol=1e-12)
    assert pair[1].shape == gp["fov"].shape
    np.testing.assert_allclose(pair[1], gp["fov"], rtol=1e-6, atol=1e-5)
    with pytest.raises(ValueError):
        T.get_lidar_point_cloud("000003", mini + "/calib", velo_dir, im_size=[1242, 375], min_intensity=0.1)


def test_ingest_to_bev_on_device(frame):
    """Raw scan -> frustum cloud -> six BEV maps + anchor keep mask with the point count staying on
    the device (dodt_bev_slices n_dev): equal to the reference's outputs for that frame."""
    import dodt_b200 as dd
    from dodt_b200 import lidar, ops
    g, calib = frame
    ref = np.load(os.path.join(GOLDEN, "s1s2_kitti_000003.npz"))
    velo = torch.from_numpy(g["velo"]).cuda()
    pts, count = lidar.get_lidar_in_camera_view(velo, calib, im_size=list(g["im_size"]))
    gen = dd.BevSlices(S.SlicesConfig())
    nx, _, nz, _, _, _ = ops.bev_grid(S.AREA_EXTENTS, S.VOXEL_SIZE)
    buf = gen.buffers(nx, nz, pts.shape[1], True)
    params = ops.make_bev_params(S.GROUND_PLANE, S.AREA_EXTENTS, S.VOXEL_SIZE, S.HEIGHT_LO, S.HEIGHT_HI,
                                 S.NUM_SLICES)
    ops.bev_slices(pts, params, buf.maps, buf.occ, buf.stats, buf.workspace, n_dev=count)
    maps = buf.maps.cpu().numpy()
    assert int(count.item()) == ref["points"].shape[1]
    for i in range(6):
        want = np.zeros((nz, nx))
        want[ref["map%d_r" % i], ref["map%d_c" % i]] = ref["map%d_v" % i]
        np.testing.assert_array_equal(maps[i], want.astype(np.float32), err_msg="map %d" % i)
    grid = dd.VoxelGrid2D.from_occupancy(buf.occ, S.VOXEL_SIZE, S.AREA_EXTENTS)
    keep = dd.get_empty_anchor_filter_2d(S.car_anchors(), grid, 1)
    np.testing.assert_array_equal(keep, np.unpackbits(ref["keep_packed"])[:89600].astype(bool))
