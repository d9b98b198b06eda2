"""GPU parity against the committed golden fixtures (tests/golden/, generated from the reference's
own Python and from independent implementations by oracle/make_golden.py): the CUDA path, called
through the reference-signature drop-ins, on the SAME inputs the reference saw.
"""
import os

import numpy as np
import pytest
import torch

from oracle import synth_ref as S

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def dd(lib):
    import dodt_b200
    return dodt_b200


def _dense(g, i, shape):
    m = np.zeros(shape)
    m[g["map%d_r" % i], g["map%d_c" % i]] = g["map%d_v" % i]
    return m


@pytest.mark.parametrize("name", ["s1s2_kitti_000003.npz", "s1s2_kitti_010005.npz", "s1s2_synth.npz"])
def test_s1_s2_equal_reference_outputs(dd, name):
    """Real KITTI frames (float64 clouds as wavedata produces them) + the degenerate synthetic one:
    six BEV maps == float32(reference), occupancy and the 89 600-anchor keep mask bit-exact."""
    g = np.load(os.path.join(GOLDEN, name))
    pc = g["points"]
    gen = dd.BevSlices(S.SlicesConfig())
    bev, grid = gen.generate_bev_and_voxel_grid('lidar', pc, S.GROUND_PLANE, S.AREA_EXTENTS, S.VOXEL_SIZE)
    maps = bev['height_maps'] + [bev['density_map']]
    assert len(maps) == 6
    for i, m in enumerate(maps):
        np.testing.assert_array_equal(m, _dense(g, i, m.shape).astype(np.float32), err_msg="map %d" % i)
    # the reference-signature entry point gives the same maps
    bev2 = gen.generate_bev('lidar', pc, S.GROUND_PLANE, S.AREA_EXTENTS, S.VOXEL_SIZE)
    for a, b in zip(bev2['height_maps'] + [bev2['density_map']], maps):
        np.testing.assert_array_equal(a, b)
    occ = grid.occ.cpu().numpy()
    want_occ = np.zeros_like(occ)
    want_occ[g["occ_x"], g["occ_z"]] = 1
    np.testing.assert_array_equal(occ, want_occ)
    anchors = S.car_anchors()
    keep = dd.get_empty_anchor_filter_2d(anchors, grid, 1)
    np.testing.assert_array_equal(keep, np.unpackbits(g["keep_packed"])[:len(anchors)].astype(bool))


def test_reference_unit_test_inputs(dd):
    g = np.load(os.path.join(GOLDEN, "s1_unit_vectors.npz"))
    vg = dd.VoxelGrid2D()
    vg.voxelize_2d(g["vg_pts"], 0.1)
    np.testing.assert_array_equal(vg.voxel_indices, g["vg_voxel_indices"])
    np.testing.assert_array_equal(vg.num_pts_in_voxel, g["vg_counts"])
    np.testing.assert_array_equal(vg.num_divisions, g["vg_num_divisions"])
    np.testing.assert_array_equal(vg.min_voxel_coord, g["vg_min"])
    np.testing.assert_allclose(vg.heights, g["vg_heights"], rtol=1e-6)
    vg = dd.VoxelGrid2D()
    vg.voxelize_2d(g["vg2_pts"], 0.1, np.array([[-50, 50], [-5, 5], [0, 70]]), ground_plane=[0, -1, 0, 1.65])
    np.testing.assert_array_equal(vg.voxel_indices, g["vg2_voxel_indices"])
    np.testing.assert_array_equal(vg.num_pts_in_voxel, g["vg2_counts"])
    np.testing.assert_array_equal(vg.heights, g["vg2_heights"].astype(np.float32))
    np.testing.assert_array_equal(vg.map_to_index(g["map_coords"]), g["map_index"])
    np.testing.assert_array_equal(vg.map_to_index(g["map_coords"].astype(np.float32)), g["map_index_f32"])
    from dodt_b200 import ops
    ii = ops.integral_image_2d(torch.from_numpy(g["ii_img"].astype(np.uint8)).cuda()).cpu().numpy()
    np.testing.assert_array_equal(ii, g["ii_image"].astype(np.int32))


def test_crop_and_resize_equals_grid_sample(dd):
    g = np.load(os.path.join(GOLDEN, "s3_grid_sample.npz"))
    ind = np.zeros(len(g["boxes"]), dtype=np.int32)
    for ch, cw in ((7, 7), (3, 3), (2, 5)):
        got = dd.crop_and_resize(g["image"], g["boxes"], ind, (ch, cw))
        np.testing.assert_allclose(got, g["crop_%dx%d" % (ch, cw)], rtol=1e-4, atol=2e-5)


def test_correlation_equals_shift_formulation(dd):
    g = np.load(os.path.join(GOLDEN, "s4_shift_formulation.npz"))
    got = dd.correlation(g["a"], g["b"], kernel_size=1, max_displacement=5, stride_1=1, stride_2=2, padding=5)
    np.testing.assert_allclose(got, g["out"], rtol=1e-5, atol=1e-7)


def test_nms_equals_torchvision(dd):
    g = np.load(os.path.join(GOLDEN, "s5_torchvision_nms.npz"))
    for k in range(3):
        boxes, scores, thr, keep = g["boxes%d" % k], g["scores%d" % k], float(g["thr%d" % k]), g["keep%d" % k]
        np.testing.assert_array_equal(dd.non_max_suppression(boxes, scores, len(boxes), thr), keep)
        np.testing.assert_array_equal(dd.non_max_suppression(boxes, scores, 17, thr), keep[:17])
