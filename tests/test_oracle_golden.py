"""CPU tests that PIN THE ORACLE: the NumPy / C restatements under oracle/ against

 (1) outputs of the reference's own Python, frozen in tests/golden/ by oracle/make_golden.py
     (two real KITTI tracking frames of the reference's unit-test dataset, a synthetic cloud with
     degenerate slices, and the inputs of the reference's unit tests),
 (2) the known answers the reference's unit tests assert, typed in here from
     wavedata/wavedata/tools/core/voxel_grid_2d_test.py, integral_image_2d_test.py,
     avod/core/anchor_filter_test.py and avod/core/box_list_ops_test.py,
 (3) independent implementations for the TensorFlow ops that are not in the checkout
     (torch grid_sample, torchvision nms, a float64 shifted-product correlation), frozen in
     tests/golden/.

No GPU, no /root/reference needed.
"""
import os

import numpy as np
import pytest

from oracle import anchor_helpers as A
from oracle import synth_ref as S
from oracle import c_oracle as CO
from oracle import np_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return np.load(os.path.join(GOLDEN, name))


def _dense(g, i, shape):
    m = np.zeros(shape)
    m[g["map%d_r" % i], g["map%d_c" % i]] = g["map%d_v" % i]
    return m


# ------------------------------------------------------------------------------------------ S1/S2


@pytest.mark.parametrize("name", ["s1s2_kitti_000003.npz", "s1s2_kitti_010005.npz", "s1s2_synth.npz"])
def test_s1_s2_against_reference_outputs(name):
    """BevSlices.generate_bev, create_sliced_voxel_grid_2d and get_empty_anchor_filter_2d of the
    reference on the stored cloud == the oracle on the same cloud, bit for bit (float64)."""
    g = _load(name)
    pc = g["points"]
    out = O.bev_slices(pc, S.GROUND_PLANE, S.AREA_EXTENTS, S.VOXEL_SIZE, S.HEIGHT_LO, S.HEIGHT_HI,
                       S.NUM_SLICES)
    maps = out["height_maps"] + [out["density_map"]]
    for i, m in enumerate(maps):
        np.testing.assert_array_equal(m, _dense(g, i, m.shape), err_msg="map %d" % i)
    occ, vox = O.occupancy_grid(pc, S.GROUND_PLANE, S.AREA_EXTENTS, S.VOXEL_SIZE)
    want_occ = np.zeros_like(occ)
    want_occ[g["occ_x"], g["occ_z"]] = 1
    np.testing.assert_array_equal(occ, want_occ)
    anchors = S.car_anchors()
    assert len(anchors) == int(g["n_anchors"]) == 89600
    keep = O.empty_anchor_filter_2d(anchors, occ, S.VOXEL_SIZE, vox["min_coord"][[0, 2]], 1)
    want = np.unpackbits(g["keep_packed"])[:len(anchors)].astype(bool)
    np.testing.assert_array_equal(keep, want)
    assert 0 < want.sum() < len(want)


def test_reference_unit_test_inputs():
    """Inputs of the reference's unit tests, outputs produced by the reference classes."""
    g = _load("s1_unit_vectors.npz")
    v = O.voxelize_2d(g["vg_pts"], 0.1)
    np.testing.assert_array_equal(v["voxel_indices"], g["vg_voxel_indices"])
    np.testing.assert_array_equal(v["heights"], g["vg_heights"])
    np.testing.assert_array_equal(v["counts"], g["vg_counts"])
    np.testing.assert_array_equal(v["num_divisions"], g["vg_num_divisions"])
    np.testing.assert_array_equal(v["min_coord"], g["vg_min"])
    ext = np.array([[-50, 50], [-5, 5], [0, 70]])
    v = O.voxelize_2d(g["vg2_pts"], 0.1, ext, ground_plane=[0, -1, 0, 1.65])
    np.testing.assert_array_equal(v["voxel_indices"], g["vg2_voxel_indices"])
    np.testing.assert_array_equal(v["heights"], g["vg2_heights"])
    np.testing.assert_array_equal(v["counts"], g["vg2_counts"])
    nd = v["num_divisions"][[0, 2]]
    mc = v["min_coord"][[0, 2]]
    np.testing.assert_array_equal(O.map_to_index(g["map_coords"], 0.1, mc, nd), g["map_index"])
    np.testing.assert_array_equal(O.map_to_index(g["map_coords"].astype(np.float32), 0.1, mc, nd),
                                  g["map_index_f32"])
    ii = O.integral_image_2d(g["ii_img"])
    np.testing.assert_array_equal(ii, g["ii_image"])
    np.testing.assert_array_equal(O.integral_query(ii, g["ii_boxes"]), g["ii_query"])
    np.testing.assert_array_equal(
        O.slice_filter(g["sf_pc"], [[-2, 2], [-5, 5], [-2, 2]], [0, 1, 0, 0], 0.2, 2.0), g["sf_filter"])
    np.testing.assert_array_equal(O.point_filter(g["pf_pc"], [[-2, 2], [-1, 1], [-2, 2]]),
                                  g["pf_extents_only"])
    np.testing.assert_array_equal(
        O.point_filter(g["pf_pc"], [[-2, 2], [-2, 2], [-2, 2]], [0, -1, 0, 1.65], 1.0), g["pf_plane"])


def test_voxel_grid_2d_known_answers():
    """wavedata/wavedata/tools/core/voxel_grid_2d_test.py:15-59 (leaf layout of the 12 corner
    points), :61-79 (bad extents raise, division counts), :81-118 (map_to_index)."""
    pts = np.array([[-39.99, 4.99, 0], [39.99, 4.99, 0], [-39.99, -4.99, 0], [39.99, -4.99, 0],
                    [-39.99, 4.99, 69.99], [39.99, 4.99, 69.99], [-39.99, -4.99, 69.99],
                    [39.99, -4.99, 69.99], [-39.99, 4.99, 69.99], [39.99, 4.99, 69.99],
                    [-39.99, -4.99, 69.99], [39.99, -4.99, 69.99]])
    v = O.voxelize_2d(pts, 0.1)
    assert (v["min_coord"] == [-400, 0, 0]).all()
    assert (v["num_divisions"] == [800, 1, 700]).all()
    filled = np.floor((pts * 10) + [400, 0, 0]).astype(np.int32)
    filled[:, 1] = 0
    expected = -1 * np.ones((800, 1, 700))
    for idx in filled:
        expected[tuple(idx)] = 0
    assert (O.leaf_layout_2d(v) == expected).all()

    rng = np.random.default_rng(0)
    points = (rng.random((70000, 3)) * [80, 8, 60]) - [40, 4, 0]
    with pytest.raises(ValueError):
        O.voxelize_2d(points, 0.1, np.array([[-30, 30], [-3, 3], [10, 60]]))
    with pytest.raises(ValueError):
        O.voxelize_2d(points[:, :2], 0.1)
    v = O.voxelize_2d(points, 0.1, np.array([[-50, 50], [-5, 5], [0, 70]]))
    assert (v["num_divisions"] == [1000, 1, 700]).all()
    assert O.leaf_layout_2d(v).shape == (1000, 1, 700)
    nd, mc = v["num_divisions"][[0, 2]], v["min_coord"][[0, 2]]
    for coords, want in [([[0, 0]], [500, 0]), (np.array([[0, 0]]) + 0.1, [501, 1]),
                         ([[-50, 0]], [0, 0]), ([[50, 70]], [1000, 700]), ([[60, 80]], [1000, 700])]:
        assert (O.map_to_index(np.array(coords, dtype=np.float64), 0.1, mc, nd) == want).all()


def test_integral_image_known_answers():
    """wavedata/wavedata/tools/core/integral_image_2d_test.py:9-48."""
    ii = O.integral_image_2d(np.ones((3, 3), dtype=np.float32))
    q = lambda b: list(O.integral_query(ii, np.array(b).T.astype(np.uint32)))
    assert q([[0, 0, 1, 1], [0, 0, 2, 2], [0, 0, 3, 3]]) == [1, 4, 9]
    assert q([[1, 1, 2, 2], [1, 1, 3, 3]]) == [1, 4]
    assert q([[0, 0, 3, 1]]) == [3]
    assert q([[0, 0, 2312, 162]]) == [9]
    with pytest.raises(ValueError):
        O.integral_image_2d(np.ones((3, 3, 3)))
    with pytest.raises(TypeError):
        O.integral_query(ii, np.zeros((4, 2), dtype=np.int64))
    with pytest.raises(ValueError):
        O.integral_query(ii, np.zeros((3, 2), dtype=np.uint32))


def test_anchor_filter_known_answers():
    """avod/core/anchor_filter_test.py:28-99: the 3D test's masks hold for the 2D filter because
    all boxes share one y layer."""
    pts = np.array([[0.51, -0.5, 1.1], [1.51, -0.5, 1.1]])
    vox = O.voxelize_2d(pts, 0.5, [(0., 2.), (-1., 0.), (0., 2.)])
    occ = (np.squeeze(O.leaf_layout_2d(vox), 1) + 1).astype(np.uint8)
    mc = vox["min_coord"][[0, 2]]
    boxes = np.array([[0.51, 0, 0.51, 1, 1, 1, 0], [0.51, 0, 0.51, 1, 1, 1, np.pi / 2.],
                      [0.51, 0, 1.1, 1, 1, 1, 0], [0.51, 0, 1.1, 1, 1, 1, np.pi / 2.],
                      [1.51, 0, 0.51, 1, 1, 1, 0], [1.51, 0, 0.51, 1, 1, 1, np.pi / 2.],
                      [1.51, 0, 1.1, 1, 1, 1, 0], [1.51, 0, 1.1, 1, 1, 1, np.pi / 2.]])
    got = O.empty_anchor_filter_2d(A.box_3d_to_anchor(boxes), occ, 0.5, mc, 1)
    assert list(got) == [False, False, True, True, False, False, True, True]
    boxes = np.array([[0.5, 0, 0.5, 2, 1, 1, 0], [0.5, 0, 0.5, 2, 1, 1, np.pi / 2.],
                      [0.5, 0, 1.5, 1, 2, 1, 0], [0.5, 0, 1.5, 1, 2, 1, np.pi / 2.],
                      [1.5, 0, 0.5, 2, 1, 1, 0], [1.5, 0, 0.5, 2, 1, 1, np.pi / 2.],
                      [1.5, 0, 1.5, 1, 2, 1, 0], [1.5, 0, 1.5, 1, 2, 1, np.pi / 2.]])
    got = O.empty_anchor_filter_2d(A.box_3d_to_anchor(boxes), occ, 0.5, mc, 1)
    assert list(got) == [False, True, True, True, False, True, True, True]


def test_degenerate_slices_match_survey_probe():
    """a slice with 0 or 1 points yields one pixel [H-1, nx/2] = (1.65 - lo_i)/hpd,
    while the density map still counts the dropped point (bev_slices.py:76-99)."""
    pc = np.array([[3.0, 10.0], [1.65 - 0.1, 1.65 - 1.0], [20.0, 30.0]])   # one point in slice 0, one in 2
    out = O.bev_slices(pc, S.GROUND_PLANE, S.AREA_EXTENTS, S.VOXEL_SIZE, S.HEIGHT_LO, S.HEIGHT_HI,
                       S.NUM_SLICES, return_debug=True)
    hpd = (S.HEIGHT_HI - S.HEIGHT_LO) / S.NUM_SLICES
    for i, m in enumerate(out["height_maps"]):
        r, c = np.nonzero(m)
        assert list(r) == [699] and list(c) == [400]
        assert m[699, 400] == (1.65 - (S.HEIGHT_LO + i * hpd)) / hpd
    assert np.count_nonzero(out["density_map"]) == 2
    assert out["slice_counts"].tolist() == [1, 0, 1, 0, 0, 2]


# ------------------------------------------------------------------------------------------ S3


def test_crop_and_resize_against_grid_sample():
    g = _load("s3_grid_sample.npz")
    img, boxes = g["image"], g["boxes"]
    ind = np.zeros(len(boxes), dtype=np.int32)
    for ch, cw in ((7, 7), (3, 3), (2, 5)):
        want = g["crop_%dx%d" % (ch, cw)]
        for impl in (O, CO):
            got = impl.crop_and_resize(img, boxes, ind, (ch, cw))
            assert got.dtype == np.float32 and got.shape == want.shape
            np.testing.assert_allclose(got, want, rtol=1e-4, atol=2e-5)


def test_crop_and_resize_semantics():
    """TF 1.3 crop_and_resize_op.cc: extrapolation outside [0, size-1], crop of 1 samples the box
    centre, flipped boxes flip the crop, out-of-range box_ind rows stay untouched (zero)."""
    rng = np.random.default_rng(5)
    img = rng.standard_normal((2, 9, 11, 3)).astype(np.float32)
    full = np.array([[0, 0, 1, 1]], dtype=np.float32)
    for impl in (O, CO):
        np.testing.assert_array_equal(impl.crop_and_resize(img, full, [1], (9, 11))[0], img[1])
        flipped = impl.crop_and_resize(img, np.array([[1, 1, 0, 0]], np.float32), [0], (9, 11))[0]
        np.testing.assert_array_equal(flipped, img[0, ::-1, ::-1])
        centre = impl.crop_and_resize(img, full, [0], (1, 1))[0, 0, 0]
        np.testing.assert_allclose(centre, img[0, 4, 5], rtol=1e-6)
        out = impl.crop_and_resize(img, np.array([[-0.5, -0.5, 1.5, 1.5]], np.float32), [0], (5, 5), -7.0)[0]
        assert (out[0] == -7.0).all() and (out[:, 0] == -7.0).all() and (out[-1] == -7.0).all()
        np.testing.assert_allclose(out[2, 2], img[0, 4, 5], rtol=1e-6)
        bad = impl.crop_and_resize(img, np.repeat(full, 2, 0), [5, 0], (2, 2))
        assert (bad[0] == 0).all() and (bad[1] != 0).any()
        assert impl.crop_and_resize(img, np.zeros((0, 4), np.float32), np.zeros(0, np.int32), (3, 3)).shape == (0, 3, 3, 3)


def test_c_oracle_matches_numpy_oracle_s3():
    rng = np.random.default_rng(6)
    img = np.abs(rng.standard_normal((1, 70, 80, 8))).astype(np.float32)
    c = rng.uniform(-0.1, 1.1, (500, 2))
    h = rng.uniform(0.0, 0.2, (500, 2))
    boxes = np.stack([c[:, 0] - h[:, 0], c[:, 1] - h[:, 1], c[:, 0] + h[:, 0], c[:, 1] + h[:, 1]], 1).astype(np.float32)
    ind = np.zeros(500, dtype=np.int32)
    for crop in ((7, 7), (3, 3), (1, 4)):
        np.testing.assert_array_equal(CO.crop_and_resize(img, boxes, ind, crop),
                                      O.crop_and_resize(img, boxes, ind, crop))


# ------------------------------------------------------------------------------------------ S4


def test_correlation_against_shift_formulation():
    g = _load("s4_shift_formulation.npz")
    for impl in (O, CO):
        got = impl.correlation(g["a"], g["b"], 1, 5, 1, 2, 5)
        assert got.shape == g["out"].shape and got.dtype == np.float32
        np.testing.assert_allclose(got, g["out"], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("kw", [dict(kernel_size=1, max_displacement=5, stride_1=1, stride_2=2, padding=5),
                                dict(kernel_size=3, max_displacement=4, stride_1=2, stride_2=2, padding=4),
                                dict(kernel_size=1, max_displacement=3, stride_1=1, stride_2=1, padding=2),
                                dict(kernel_size=1, max_displacement=20, stride_1=1, stride_2=2, padding=20)])
def test_c_oracle_matches_numpy_oracle_s4(kw):
    rng = np.random.default_rng(7)
    a = np.abs(rng.standard_normal((2, 14, 17, 40))).astype(np.float32)
    b = np.abs(rng.standard_normal((2, 14, 17, 40))).astype(np.float32)
    np.testing.assert_array_equal(CO.correlation(a, b, **kw), O.correlation(a, b, **kw))


def test_correlation_shape_and_errors():
    """avod/core/ops/correlation/correlation_kernel.cc:23,39-57."""
    assert O.correlation_out_shape(700, 800, 1, 5, 1, 2, 5) == (700, 800, 25)
    assert O.correlation_out_shape(64, 96, 1, 20, 1, 2, 20) == (64, 96, 441)
    assert O.correlation_out_shape(48, 64, 3, 4, 2, 2, 4) == (23, 31, 25)
    with pytest.raises(ValueError):
        O.correlation_out_shape(64, 64, 2, 4, 1, 2, 4)
    with pytest.raises(ValueError):
        O.correlation_out_shape(4, 64, 1, 8, 1, 2, 0)
    with pytest.raises(ValueError):
        O.correlation(np.zeros((1, 8, 8, 4), np.float32), np.zeros((1, 8, 9, 4), np.float32))


# ------------------------------------------------------------------------------------------ S5


def test_nms_against_torchvision():
    g = _load("s5_torchvision_nms.npz")
    for k in range(3):
        boxes, scores, thr, keep = g["boxes%d" % k], g["scores%d" % k], float(g["thr%d" % k]), g["keep%d" % k]
        for impl in (O, CO):
            got = impl.non_max_suppression(boxes, scores, len(boxes), thr)
            np.testing.assert_array_equal(got, keep)
            np.testing.assert_array_equal(impl.non_max_suppression(boxes, scores, 17, thr), keep[:17])


def test_iou_known_answers():
    """avod/core/box_list_ops_test.py:86-98 (test_iou)."""
    c1 = np.array([[4.0, 3.0, 7.0, 5.0], [5.0, 6.0, 10.0, 7.0]], dtype=np.float32)
    c2 = np.array([[3.0, 4.0, 6.0, 8.0], [14.0, 14.0, 15.0, 15.0], [0.0, 0.0, 20.0, 20.0]], dtype=np.float32)
    want = [[2.0 / 16.0, 0, 6.0 / 400.0], [1.0 / 16.0, 0.0, 5.0 / 400.0]]
    area = lambda b: (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    for i in range(2):
        got = O.iou_matrix_row(c1[i], area(c1[i:i + 1])[0], c2, area(c2))
        np.testing.assert_allclose(got, want[i], rtol=1e-6)


def test_nms_semantics():
    """TF 1.3 non_max_suppression_op.cc: strict '>' threshold, flipped corners, zero-area boxes
    never suppress, max_output_size caps, empty input."""
    b = np.array([[0, 0, 1, 1], [0, 0, 1, 0.5], [1, 1, 0, 0], [0.2, 0.2, 0.2, 0.9], [2, 2, 3, 3]], dtype=np.float32)
    s = np.array([0.9, 0.8, 0.7, 0.6, 0.5], dtype=np.float32)
    for impl in (O, CO):
        # IoU(0,1) = 0.5 exactly: not suppressed at thr 0.5, suppressed below; box 2 == box 0 flipped
        assert impl.non_max_suppression(b, s, 10, 0.5).tolist() == [0, 1, 3, 4]
        assert impl.non_max_suppression(b, s, 10, 0.49).tolist() == [0, 3, 4]
        assert impl.non_max_suppression(b, s, 2, 0.5).tolist() == [0, 1]
        assert impl.non_max_suppression(b, s, 0, 0.5).tolist() == []
        assert impl.non_max_suppression(np.zeros((0, 4), np.float32), np.zeros(0, np.float32), 5, 0.5).tolist() == []
        # equal scores: ascending index
        assert impl.non_max_suppression(b[[4, 0]], np.array([0.3, 0.3], np.float32), 5, 0.5).tolist() == [0, 1]


def test_c_oracle_matches_numpy_oracle_s5():
    rng = np.random.default_rng(8)
    a = S.car_anchors()[::37]
    _, boxes, scores = S.rpn_proposals(3, 0, a)
    for max_out, thr in ((1024, 0.8), (100, 0.01), (300, 0.5)):
        np.testing.assert_array_equal(CO.non_max_suppression(boxes, scores, max_out, thr),
                                      O.non_max_suppression(boxes, scores, max_out, thr))


def test_lidar_ingest_against_reference_outputs():
    """The reference's get_lidar_point_cloud on a real KITTI tracking scan (frozen by
    oracle/make_golden.py) == the oracle restatement: same points kept, same order, same float64
    coordinates (same NumPy, same BLAS)."""
    g = _load("lidar_kitti_000003.npz")
    fov = O.lidar_in_camera_view(g["velo"], g["r0_rect"], g["tr_velodyne_to_cam"], g["p2"], list(g["im_size"]))
    assert fov.shape == g["fov"].shape == (3, 20583)
    np.testing.assert_allclose(fov, g["fov"], rtol=1e-13, atol=1e-13)
    full = O.lidar_in_camera_view(g["velo"], g["r0_rect"], g["tr_velodyne_to_cam"], g["p2"])
    np.testing.assert_allclose(full[:, ::7], g["full_every_7th"], rtol=1e-13, atol=1e-13)
    np.testing.assert_array_equal(g["fov"], _load("s1s2_kitti_000003.npz")["points"])


def test_ego_motion_alignment_against_reference_outputs():
    """DODT's frame-pair ingest (KittiTrackingDataset.coordinate_transform / point_cloud_transform run
    from the checkout on fixture frames 000003 -> 000004, frozen by oracle/make_golden.py): the OXTS
    scalars and the moved float32 scan are reproduced bit for bit by the oracle and by the host
    mirror of the product package; the frustum crop of the moved scan keeps the same points."""
    from dodt_b200 import lidar
    g = _load("lidar_pair_000003_000004.npz")
    for trans, matrix, delta in (O.oxts_coordinate_transform(g["oxts"][0], g["oxts"][1]),
                                 lidar.coordinate_transform(str(g["oxts_line0"]), str(g["oxts_line1"]))):
        np.testing.assert_array_equal(trans, g["trans"])
        np.testing.assert_array_equal(matrix, g["matrix"])
        assert delta == g["delta"]
    moved = O.point_cloud_transform(g["velo1"].T, g["trans"], g["matrix"])
    assert moved.dtype == np.float32
    np.testing.assert_array_equal(moved, g["moved"])
    np.testing.assert_array_equal(moved[3], g["velo1"][:, 3])          # intensity untouched
    fov = O.lidar_in_camera_view(moved.T, g["r0_rect"], g["tr_velodyne_to_cam"], g["p2"], list(g["im_size"]))
    assert fov.shape == g["fov"].shape == (3, 5443)
    np.testing.assert_allclose(fov, g["fov"], rtol=1e-13, atol=1e-13)


def test_tracking_file_readers_against_reference_outputs(tmp_path):
    """dodt_b200.tracking_utils' host-side readers (KITTI calibration text, oxts line, velodyne .bin)
    on the reference's fixture files (tests/golden/kitti_mini: calib/0000.txt and the first six oxts
    lines, data files of the KITTI tracking set) == what the reference's readers returned when
    oracle/make_golden.py ran them."""
    from dodt_b200 import tracking_utils as T
    mini = os.path.join(GOLDEN, "kitti_mini")
    g, gp = _load("lidar_kitti_000003.npz"), _load("lidar_pair_000003_000004.npz")
    calib = T.read_tracking_calibration(mini + "/calib", 0)
    for k in ("p2", "r0_rect", "tr_velodyne_to_cam"):
        np.testing.assert_array_equal(getattr(calib, k), g[k])
    assert calib.p0.shape == calib.p1.shape == calib.p3.shape == (3, 4)
    for i, frame in enumerate(("000003", "000004")):
        rec = T.get_oxts(mini + "/oxts", frame)
        np.testing.assert_array_equal([rec.latitude, rec.longitude, rec.altitude, rec.roll, rec.pitch, rec.yaw],
                                      gp["oxts"][i])
    os.makedirs(tmp_path / "velodyne" / "0000")
    g["velo"].tofile(tmp_path / "velodyne" / "0000" / "000003.bin")
    raw = T.get_raw_lidar_point_cloud("000003", str(tmp_path / "velodyne"))
    assert raw.dtype == np.float32 and raw.shape == (4, len(g["velo"]))
    np.testing.assert_array_equal(raw.T, g["velo"])
    assert T.read_lidar(str(tmp_path / "velodyne" / "0000"), 99) == []
    np.testing.assert_array_equal(T.get_road_plane("000003", str(tmp_path / "planes")), [0.0, -1.0, 0.0, 1.65])


def test_kitti_like_cloud_has_the_surveyed_occupancy():
    """synth.point_cloud_kitti (bench.py --workload kitti) lands in the ranges SURVEY 8(d) measured on
    the reference's real KITTI tracking frames: 16-20 k points, 9.3-15.1 k of 89 600 anchors kept by
    the empty-anchor filter (the evenly spread benchmark cloud keeps ~60 k), a few thousand occupied
    cells in the 0.2-2.0 m slice."""
    a = S.car_anchors()
    for frame in (0, 5):
        pc = S.point_cloud_kitti(2, frame).astype(np.float64)
        assert 16000 <= pc.shape[1] <= 20500
        occ, vox = O.occupancy_grid(pc, S.GROUND_PLANE, S.AREA_EXTENTS, S.VOXEL_SIZE)
        keep = O.empty_anchor_filter_2d(a, occ, S.VOXEL_SIZE, vox["min_coord"][[0, 2]], 1)
        assert 9300 <= int(keep.sum()) <= 15100
        assert 1750 <= int(occ.sum()) <= 4000
