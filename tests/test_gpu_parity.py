"""GPU parity tests: every stage of the front end, called through the C ABI (via dodt_b200.ops /
the reference-signature drop-ins), against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): integer outputs — winner indices, per-cell counts, occupancy,
integral image, keep mask, NMS indices — bit-exact; float outputs within 1e-5 relative (the BEV
maps are in fact compared bit-exactly against np.float32(oracle)).
"""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-5


@pytest.fixture(scope="module")
def dd(lib):
    import dodt_b200
    return dodt_b200


def _cfg():
    from oracle import synth_ref as synth
    return synth


# ------------------------------------------------------------------------------------------ S1


def _run_bev(dd, pc, voxel=None, plane=None, extents=None, lo=None, hi=None, S=None, debug=True):
    s = _cfg()
    voxel = s.VOXEL_SIZE if voxel is None else voxel
    plane = s.GROUND_PLANE if plane is None else plane
    extents = s.AREA_EXTENTS if extents is None else extents
    cfg = s.SlicesConfig(s.HEIGHT_LO if lo is None else lo, s.HEIGHT_HI if hi is None else hi,
                         s.NUM_SLICES if S is None else S)
    gen = dd.BevSlices(cfg)
    from dodt_b200.bev_slices import _to_device_points
    pts, _ = _to_device_points(pc, gen.device)
    buf = gen.generate_bev_device('lidar', pts, plane, extents, voxel, with_occupancy=True,
                                  debug=debug)
    torch.cuda.synchronize()
    ref = O.bev_slices(np.asarray(pc, dtype=np.float64), plane, extents, voxel, cfg.height_lo,
                       cfg.height_hi, cfg.num_slices, return_debug=True)
    return buf, ref, cfg


def _check_bev(buf, ref, S):
    maps = buf.maps.cpu().numpy()
    stats = buf.stats.cpu().numpy()
    assert stats[19] == 0 and stats[20] == 0
    np.testing.assert_array_equal(stats[:S], ref["slice_counts"][:S])
    assert stats[16] == ref["slice_counts"][S]
    for i in range(S):
        want = ref["height_maps"][i].astype(np.float32)
        np.testing.assert_array_equal(maps[i], want, err_msg="height map %d" % i)
        if buf.winner is not None and ref["slice_counts"][i] > 1:
            np.testing.assert_array_equal(buf.winner[i].cpu().numpy(), ref["winner"][i])
    np.testing.assert_array_equal(maps[S], ref["density_map"].astype(np.float32))
    if buf.counts is not None:
        np.testing.assert_array_equal(buf.counts.cpu().numpy(), ref["counts"])


@pytest.mark.parametrize("frame", [0, 1])
def test_bev_synthetic_config_a_fp32(dd, frame):
    s = _cfg()
    pc = s.point_cloud(1, frame)                       # 120k fp32 points
    buf, ref, cfg = _run_bev(dd, pc)
    _check_bev(buf, ref, cfg.num_slices)
    occ, _ = O.occupancy_grid(pc.astype(np.float64), s.GROUND_PLANE, s.AREA_EXTENTS, s.VOXEL_SIZE)
    np.testing.assert_array_equal(buf.occ.cpu().numpy(), occ)


def test_bev_fp64_points_unaligned_n(dd):
    rng = np.random.default_rng(7)
    n = 50001
    pc = np.stack([rng.uniform(-45, 45, n), rng.uniform(-6, 4, n), rng.uniform(-5, 75, n)])
    buf, ref, cfg = _run_bev(dd, pc)                   # float64, many points outside the extents
    _check_bev(buf, ref, cfg.num_slices)


def test_bev_dense_config_e(dd):
    s = _cfg()
    pc = s.point_cloud(5, 0, n_points=500000)
    buf, ref, cfg = _run_bev(dd, pc, voxel=s.VOXEL_SIZE_DENSE)
    assert buf.maps.shape == (6, 1400, 1600)
    _check_bev(buf, ref, cfg.num_slices)


def test_bev_boundary_points(dd):
    """Points exactly on voxel edges, slice boundaries and the open extents."""
    s = _cfg()
    v = s.VOXEL_SIZE
    hpd = (s.HEIGHT_HI - s.HEIGHT_LO) / 5
    xs = np.array([k * v for k in range(-400, 400, 37)] + [-40.0, 40.0, -39.999999, 39.999999])
    ys = np.array([1.65 - (s.HEIGHT_LO + k * hpd) for k in range(6)] +
                  [np.nextafter(1.65 - (s.HEIGHT_LO + k * hpd), 10) for k in range(6)] +
                  [np.nextafter(1.65 - (s.HEIGHT_LO + k * hpd), -10) for k in range(6)] + [-5.0, 3.0])
    zs = np.array([k * v for k in range(0, 700, 53)] + [0.0, 70.0, 1e-9, 69.9999999])
    X, Y, Z = np.meshgrid(xs, ys, zs, indexing="ij")
    pc = np.stack([X.ravel(), Y.ravel(), Z.ravel()])
    buf, ref, cfg = _run_bev(dd, pc)
    _check_bev(buf, ref, cfg.num_slices)
    pc32 = pc.astype(np.float32)
    buf, ref, cfg = _run_bev(dd, pc32)
    _check_bev(buf, ref, cfg.num_slices)


def test_bev_degenerate_slices(dd):
    """bev_slices.py:76-99: slices holding 0 or 1 points fall back to one origin point."""
    pc = np.array([[1.0, 2.0, 2.05, -3.0, -3.02, 0.01],
                   [1.80, 1.2, 1.21, 0.4, 0.41, -0.25],    # heights -0.15 (slice 0, alone), 0.45/0.44, 1.25/1.24, 1.9 (slice 4, alone)
                   [10.0, 20.0, 20.03, 30.0, 30.0, 0.02]])
    buf, ref, cfg = _run_bev(dd, pc)
    _check_bev(buf, ref, cfg.num_slices)
    pc1 = pc[:, :1]                                          # a single point in the whole cloud
    buf, ref, cfg = _run_bev(dd, pc1)
    _check_bev(buf, ref, cfg.num_slices)
    maps = buf.maps.cpu().numpy()
    assert maps[4, 699, 400] < 0                             # the reference's negative origin value


def test_bev_hot_cell_contention(dd):
    """100k points in a handful of cells: warp-aggregated atomics must still pick the winner."""
    rng = np.random.default_rng(3)
    n = 100000
    pc = np.stack([rng.uniform(0.0, 0.35, n), 1.65 - rng.uniform(-0.2, 2.3, n),
                   rng.uniform(10.0, 10.25, n)]).astype(np.float32)
    buf, ref, cfg = _run_bev(dd, pc)
    _check_bev(buf, ref, cfg.num_slices)


def test_bev_general_plane_tolerance(dd):
    """A tilted plane: same cells and winners; heights within 1e-5 relative (np.dot's summation
    order is BLAS-dependent, so predicate ties on the plane itself are excluded by construction)."""
    s = _cfg()
    pc = s.point_cloud(1, 3, n_points=30000).astype(np.float64)
    plane = [0.01, -0.9998, 0.015, 1.62]
    buf, ref, cfg = _run_bev(dd, pc, plane=plane)
    maps = buf.maps.cpu().numpy()
    for i in range(5):
        np.testing.assert_allclose(maps[i], ref["height_maps"][i], rtol=RTOL, atol=1e-6)
    np.testing.assert_array_equal(maps[5], ref["density_map"].astype(np.float32))


def test_generate_bev_numpy_dropin(dd):
    s = _cfg()
    pc = s.point_cloud(1, 2, n_points=20000).astype(np.float64)
    gen = dd.BevSlices(s.SlicesConfig())
    out = gen.generate_bev('lidar', pc, s.GROUND_PLANE, s.AREA_EXTENTS, s.VOXEL_SIZE)
    ref = O.bev_slices(pc, s.GROUND_PLANE, s.AREA_EXTENTS, s.VOXEL_SIZE, s.HEIGHT_LO, s.HEIGHT_HI, 5)
    assert len(out['height_maps']) == 5 and out['density_map'].shape == (700, 800)
    for a, b in zip(out['height_maps'], ref['height_maps']):
        np.testing.assert_array_equal(a, b.astype(np.float32))
    with pytest.raises(ValueError):
        gen.generate_bev('lidar', pc[:2], s.GROUND_PLANE, s.AREA_EXTENTS, s.VOXEL_SIZE)
    with pytest.raises(IndexError):
        far = np.array([[100.0], [0.0], [100.0]])
        gen.generate_bev('lidar', far, s.GROUND_PLANE, s.AREA_EXTENTS, s.VOXEL_SIZE)


# ------------------------------------------------------------------------------------------ S2


def test_voxel_grid_2d_reference_vectors(dd):
    """wavedata/wavedata/tools/core/voxel_grid_2d_test.py:15-59,81-118."""
    pts = np.array([[-39.99, 4.99, 0], [39.99, 4.99, 0], [-39.99, -4.99, 0], [39.99, -4.99, 0],
                    [-39.99, 4.99, 69.99], [39.99, 4.99, 69.99], [-39.99, -4.99, 69.99],
                    [39.99, -4.99, 69.99], [-39.99, 4.99, 69.99], [39.99, 4.99, 69.99],
                    [-39.99, -4.99, 69.99], [39.99, -4.99, 69.99]])
    vg = dd.VoxelGrid2D()
    vg.voxelize_2d(pts, 0.1)
    assert (vg.min_voxel_coord == [-400, 0, 0]).all()
    assert (vg.max_voxel_coord == [399, 0, 699]).all()
    assert (vg.num_divisions == [800, 1, 700]).all()
    filled = np.floor((pts * 10) + [400, 0, 0]).astype(np.int32)
    filled[:, 1] = 0
    expected = -1 * np.ones((800, 1, 700))
    for idx in filled:
        expected[tuple(idx)] = 0
    assert (vg.leaf_layout_2d == expected).all()
    ref = O.voxelize_2d(pts, 0.1)
    np.testing.assert_array_equal(vg.voxel_indices, ref["voxel_indices"])
    np.testing.assert_array_equal(vg.num_pts_in_voxel, ref["counts"])
    np.testing.assert_allclose(vg.heights, ref["heights"], rtol=1e-6)

    rng = np.random.default_rng(0)
    points = (rng.random((70000, 3)) * [80, 8, 60]) - [40, 4, 0]
    with pytest.raises(ValueError):
        dd.VoxelGrid2D().voxelize_2d(points, 0.1, np.array([[-30, 30], [-3, 3], [10, 60]]))
    vg = dd.VoxelGrid2D()
    vg.voxelize_2d(points, 0.1, np.array([[-50, 50], [-5, 5], [0, 70]]))
    assert (vg.num_divisions == [1000, 1, 700]).all()
    assert vg.leaf_layout_2d.shape == (1000, 1, 700)
    ref = O.voxelize_2d(points, 0.1, np.array([[-50, 50], [-5, 5], [0, 70]]))
    np.testing.assert_array_equal(vg.leaf_layout_2d, O.leaf_layout_2d(ref))
    np.testing.assert_array_equal(vg.num_pts_in_voxel, ref["counts"])
    for coords, expected in [([[0, 0]], [500, 0]), (np.array([[0, 0]]) + 0.1, [501, 1]),
                             ([[-50, 0]], [0, 0]), ([[50, 70]], [1000, 700]),
                             ([[60, 80]], [1000, 700])]:
        assert (vg.map_to_index(np.array(coords, dtype=np.float64)) == expected).all()


def test_integral_image_reference_vectors(dd):
    """wavedata/wavedata/tools/core/integral_image_2d_test.py:9-48 through the device integral
    image (queries evaluated on the host from the device image)."""
    from dodt_b200 import ops
    occ = torch.ones((3, 3), dtype=torch.uint8, device="cuda")
    ii = ops.integral_image_2d(occ).cpu().numpy().astype(np.float64)
    np.testing.assert_array_equal(ii, O.integral_image_2d(np.ones((3, 3))))
    q = lambda b: O.integral_query(ii, np.array(b).T.astype(np.uint32))
    assert list(q([[0, 0, 1, 1], [0, 0, 2, 2], [0, 0, 3, 3]])) == [1, 4, 9]
    assert list(q([[1, 1, 2, 2], [1, 1, 3, 3]])) == [1, 4]
    assert q([[0, 0, 3, 1]])[0] == 3
    assert q([[0, 0, 2312, 162]])[0] == 9


@pytest.mark.parametrize("shape", [(800, 700), (1600, 1400), (37, 53), (16, 1), (1, 130), (1000, 701)])
def test_integral_image_random(dd, shape):
    from dodt_b200 import ops
    rng = np.random.default_rng(shape[0])
    occ = (rng.random(shape) < 0.07).astype(np.uint8)
    ii = ops.integral_image_2d(torch.from_numpy(occ).cuda()).cpu().numpy()
    np.testing.assert_array_equal(ii, O.integral_image_2d(occ).astype(np.int32))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_anchor_filter_config_b(dd, dtype):
    s = _cfg()
    pc = s.point_cloud(2, 0)
    gen = dd.BevSlices(s.SlicesConfig())
    _, grid = gen.generate_bev_and_voxel_grid('lidar', torch.from_numpy(pc).cuda(), s.GROUND_PLANE,
                                              s.AREA_EXTENTS, s.VOXEL_SIZE)
    anchors = s.car_anchors().astype(dtype)
    keep = dd.get_empty_anchor_filter_2d(anchors, grid, density_threshold=1)
    occ, vox = O.occupancy_grid(pc.astype(np.float64), s.GROUND_PLANE, s.AREA_EXTENTS, s.VOXEL_SIZE)
    want = O.empty_anchor_filter_2d(anchors, occ, s.VOXEL_SIZE, vox["min_coord"][[0, 2]], 1)
    assert keep.dtype == bool and keep.shape == (89600,)
    np.testing.assert_array_equal(keep, want)
    assert 0 < want.sum() < len(want)
    keep3 = dd.get_empty_anchor_filter_2d(anchors, grid, density_threshold=3)
    np.testing.assert_array_equal(
        keep3, O.empty_anchor_filter_2d(anchors, occ, s.VOXEL_SIZE, vox["min_coord"][[0, 2]], 3))


def test_anchor_filter_reference_3d_masks_in_2d(dd):
    """avod/core/anchor_filter_test.py:28-99 — the 3D test's masks hold for the 2D filter because
    every box shares one y layer."""
    pts = np.array([[0.51, -0.5, 1.1], [1.51, -0.5, 1.1]])
    vg = dd.VoxelGrid2D()
    vg.voxelize_2d(pts, 0.5, extents=[(0., 2.), (-1., 0.), (0., 2.)])
    from oracle.anchor_helpers import box_3d_to_anchor
    boxes = np.array([[0.51, 0, 0.51, 1, 1, 1, 0], [0.51, 0, 0.51, 1, 1, 1, np.pi / 2.],
                      [0.51, 0, 1.1, 1, 1, 1, 0], [0.51, 0, 1.1, 1, 1, 1, np.pi / 2.],
                      [1.51, 0, 0.51, 1, 1, 1, 0], [1.51, 0, 0.51, 1, 1, 1, np.pi / 2.],
                      [1.51, 0, 1.1, 1, 1, 1, 0], [1.51, 0, 1.1, 1, 1, 1, np.pi / 2.]])
    got = dd.get_empty_anchor_filter_2d(box_3d_to_anchor(boxes), vg, 1)
    assert list(got) == [False, False, True, True, False, False, True, True]
    boxes = np.array([[0.5, 0, 0.5, 2, 1, 1, 0], [0.5, 0, 0.5, 2, 1, 1, np.pi / 2.],
                      [0.5, 0, 1.5, 1, 2, 1, 0], [0.5, 0, 1.5, 1, 2, 1, np.pi / 2.],
                      [1.5, 0, 0.5, 2, 1, 1, 0], [1.5, 0, 0.5, 2, 1, 1, np.pi / 2.],
                      [1.5, 0, 1.5, 1, 2, 1, 0], [1.5, 0, 1.5, 1, 2, 1, np.pi / 2.]])
    got = dd.get_empty_anchor_filter_2d(box_3d_to_anchor(boxes), vg, 1)
    assert list(got) == [False, True, True, True, False, True, True, True]
    with pytest.raises(TypeError):
        dd.get_empty_anchor_filter_2d(np.zeros((4, 5)), vg, 1)


# ------------------------------------------------------------------------------------------ S3


def _rand_boxes(rng, n, spread=0.15):
    c = rng.uniform(-0.05, 1.05, (n, 2))
    h = rng.uniform(0.005, spread, (n, 2))
    b = np.stack([c[:, 0] - h[:, 0], c[:, 1] - h[:, 1], c[:, 0] + h[:, 0], c[:, 1] + h[:, 1]], 1)
    return b.astype(np.float32)


@pytest.mark.parametrize("C,crop,n", [(32, (7, 7), 1024), (25, (7, 7), 300), (1, (3, 3), 5000),
                                      (3, (1, 1), 64), (8, (1, 5), 64), (32, (2, 3), 33)])
def test_crop_and_resize_random(dd, C, crop, n):
    rng = np.random.default_rng(C * 100 + n)
    img = rng.standard_normal((2, 45, 67, C)).astype(np.float32)
    boxes = _rand_boxes(rng, n, 0.3)
    boxes[:5] = [[0, 0, 1, 1], [0.2, 0.2, 0.2, 0.2], [1, 1, 0, 0], [-1, -1, 2, 2], [0.5, 0, 0.5, 1]]
    ind = rng.integers(0, 2, n).astype(np.int32)
    got = dd.crop_and_resize(img, boxes, ind, crop)
    want = O.crop_and_resize(img, boxes, ind, crop)
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=1e-6)
    np.testing.assert_array_equal(got == 0, want == 0)          # same extrapolation decisions
    got = dd.crop_and_resize(img, boxes, ind, crop, extrapolation_value=-7.5)
    np.testing.assert_allclose(got, O.crop_and_resize(img, boxes, ind, crop, -7.5), rtol=RTOL, atol=1e-6)


def test_crop_and_resize_config_b_shapes(dd):
    """7x7 crops of [1,700,800,32] and [1,360,1200,32] for 1024 proposals; 3x3 crops of the
    1-channel bottlenecks for the anchors that survive the filter."""
    s = _cfg()
    rng = np.random.default_rng(11)
    anchors = s.car_anchors()
    sel = rng.choice(len(anchors), 1024, replace=False)
    bev_boxes, img_boxes = s.crop_boxes(anchors[sel])
    ind = np.zeros(1024, dtype=np.int32)
    bev = np.abs(rng.standard_normal((1, 700, 800, 32), dtype=np.float32))
    img = np.abs(rng.standard_normal((1, 360, 1200, 32), dtype=np.float32))
    for fmap, boxes in ((bev, bev_boxes), (img, img_boxes)):
        got = dd.crop_and_resize(fmap, boxes, ind, (7, 7))
        np.testing.assert_allclose(got, O.crop_and_resize(fmap, boxes, ind, (7, 7)), rtol=RTOL, atol=1e-6)
    sel = rng.choice(len(anchors), 12000, replace=False)
    bev_boxes, img_boxes = s.crop_boxes(anchors[sel])
    ind = np.zeros(12000, dtype=np.int32)
    got = dd.crop_and_resize(bev[..., :1].copy(), bev_boxes, ind, (3, 3))
    np.testing.assert_allclose(got, O.crop_and_resize(bev[..., :1], bev_boxes, ind, (3, 3)), rtol=RTOL, atol=1e-6)


def test_crop_and_resize_errors_and_empty(dd):
    img = np.zeros((1, 8, 8, 4), dtype=np.float32)
    out = dd.crop_and_resize(img, np.zeros((0, 4), np.float32), np.zeros((0,), np.int32), (3, 3))
    assert out.shape == (0, 3, 3, 4)
    with pytest.raises(ValueError):
        dd.crop_and_resize(img[0], np.zeros((1, 4), np.float32), np.zeros((1,), np.int32), (3, 3))
    with pytest.raises(ValueError):
        dd.crop_and_resize(img, np.zeros((1, 4), np.float32), np.zeros((1,), np.int32), (3, 3), method='nearest')


# ------------------------------------------------------------------------------------------ S4


@pytest.mark.parametrize("shape,kw", [
    ((1, 40, 72, 32), dict(kernel_size=1, max_displacement=5, stride_1=1, stride_2=2, padding=5)),   # DODT
    ((2, 19, 23, 32), dict(kernel_size=1, max_displacement=5, stride_1=1, stride_2=2, padding=5)),
    ((1, 30, 30, 8), dict(kernel_size=1, max_displacement=2, stride_1=1, stride_2=2, padding=2)),
    ((1, 24, 20, 16), dict(kernel_size=1, max_displacement=4, stride_1=1, stride_2=1, padding=4)),
    ((1, 21, 17, 6), dict(kernel_size=3, max_displacement=4, stride_1=2, stride_2=2, padding=4)),
    ((1, 16, 16, 3), dict(kernel_size=1, max_displacement=20, stride_1=1, stride_2=2, padding=20)),  # defaults
    ((2, 12, 14, 40), dict(kernel_size=3, max_displacement=3, stride_1=1, stride_2=1, padding=5)),
])
def test_correlation_small(dd, shape, kw):
    rng = np.random.default_rng(shape[1] * 7 + shape[3])
    a = np.abs(rng.standard_normal(shape)).astype(np.float32)
    b = np.abs(rng.standard_normal(shape)).astype(np.float32)
    got = dd.correlation(a, b, **kw)
    want = O.correlation(a, b, **kw)
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=1e-7)


def test_correlation_config_c_full_size(dd):
    """[1,700,800,32] x2 -> [1,700,800,25]; compared with the oracle on full rows near the borders
    and in the middle, plus structure checks on the whole map."""
    s = _cfg()
    f0, f1 = s.feature_pair(3, 0)
    kw = dict(kernel_size=1, max_displacement=5, stride_1=1, stride_2=2, padding=5)
    got = dd.correlation(torch.from_numpy(f0).cuda(), torch.from_numpy(f1).cuda(), **kw).cpu().numpy()
    assert got.shape == (1, 700, 800, 25)
    for lo, hi in ((0, 24), (338, 362), (676, 700)):
        # the oracle on a strip of input rows equals the full result on the rows whose +-4 row
        # neighbourhood lies inside the strip (or is cut by the true image border)
        want = O.correlation(f0[:, lo:hi], f1[:, lo:hi], **kw)
        v0 = lo if lo == 0 else lo + 4
        v1 = hi if hi == 700 else hi - 4
        np.testing.assert_allclose(got[:, v0:v1], want[:, v0 - lo:v1 - lo], rtol=RTOL, atol=1e-7)
    # f1 is f0 shifted by (+2 rows, -2 cols): displacement (p, o) = (+2, -2) -> k = (1+2)*5 + (−1+2) = 16
    centre = got[0, 50:650, 50:750]
    assert centre.argmax(axis=-1).ravel().tolist().count(16) > 0.7 * centre.shape[0] * centre.shape[1]


def test_correlation_errors(dd):
    a = np.zeros((1, 8, 8, 4), dtype=np.float32)
    with pytest.raises(ValueError):
        dd.correlation(a, a, kernel_size=2)
    with pytest.raises(ValueError):
        dd.correlation(a[0], a[0])
    with pytest.raises(ValueError):
        dd.correlation(a, a, kernel_size=1, max_displacement=20, padding=0)   # nothing fits


# ------------------------------------------------------------------------------------------ S5


def _nms_case(rng, n, cluster):
    """n boxes around `cluster` centres (heavy overlap when cluster << n), tie-free scores."""
    centres = rng.uniform(0.1, 0.9, (cluster, 2))
    which = rng.integers(0, cluster, n)
    c = centres[which] + rng.normal(0, 0.004, (n, 2))
    h = np.abs(rng.normal(0.03, 0.004, (n, 2))) + 0.002
    boxes = np.stack([c[:, 0] - h[:, 0], c[:, 1] - h[:, 1], c[:, 0] + h[:, 0], c[:, 1] + h[:, 1]], 1)
    flip = rng.random(n) < 0.3                     # TF accepts any corner order
    boxes[flip] = boxes[flip][:, [2, 3, 0, 1]]
    scores = rng.permutation(np.linspace(0.0, 1.0, n))
    return boxes.astype(np.float32), scores.astype(np.float32)


@pytest.mark.parametrize("n,cluster,max_out,thr", [
    (1, 1, 10, 0.5), (2, 1, 10, 0.5), (63, 5, 100, 0.5), (64, 64, 64, 0.8), (65, 3, 10, 0.3),
    (1000, 50, 100, 0.01), (1024, 1024, 100, 0.01), (1536, 100, 1024, 0.8), (1537, 20, 1024, 0.8),
    (5000, 40, 1024, 0.8), (5000, 5000, 300, 0.8), (12000, 300, 1024, 0.8), (12000, 12000, 1024, 0.8),
    (20000, 30, 4000, 0.5), (3000, 10, 50, 0.0), (4000, 4000, 5000, 1.0),
])
def test_nms_random(dd, n, cluster, max_out, thr):
    rng = np.random.default_rng(n * 31 + cluster)
    boxes, scores = _nms_case(rng, n, cluster)
    got = dd.non_max_suppression(boxes, scores, max_out, thr)
    want = O.non_max_suppression(boxes, scores, max_out, thr)
    assert got.dtype == np.int32
    np.testing.assert_array_equal(got, want)


def test_nms_config_b_and_stress(dd):
    """RPN NMS of config B: anchors kept by the filter, regressed, IoU 0.8, 1024 outputs; then the
    89.6k-anchor stress size and the final NMS (0.01, 100)."""
    s = _cfg()
    anchors = s.car_anchors()
    rng = np.random.default_rng(5)
    kept = np.sort(rng.choice(len(anchors), 12000, replace=False))
    _, boxes, scores = s.rpn_proposals(2, 0, anchors[kept])
    for max_out, thr in ((1024, 0.8), (300, 0.8)):
        np.testing.assert_array_equal(dd.non_max_suppression(boxes, scores, max_out, thr),
                                      O.non_max_suppression(boxes, scores, max_out, thr))
    _, boxes, scores = s.rpn_proposals(2, 1, anchors)
    np.testing.assert_array_equal(dd.non_max_suppression(boxes, scores, 1024, 0.8),
                                  O.non_max_suppression(boxes, scores, 1024, 0.8))
    top = O.non_max_suppression(boxes, scores, 1024, 0.8)
    np.testing.assert_array_equal(dd.non_max_suppression(boxes[top], scores[top], 100, 0.01),
                                  O.non_max_suppression(boxes[top], scores[top], 100, 0.01))


def test_nms_device_form_and_empty(dd):
    from dodt_b200 import ops
    boxes = torch.zeros((0, 4), device="cuda")
    keep, n_keep = ops.nms(boxes, torch.zeros((0,), device="cuda"), 8, 0.5)
    assert n_keep.cpu().tolist() == [0, 1] and (keep.cpu().numpy() == -1).all()
    b = torch.tensor([[0, 0, 1, 1], [0, 0.1, 1, 1.1], [0, -0.1, 1, 0.9], [0, 10, 1, 11]],
                     dtype=torch.float32, device="cuda")
    sc = torch.tensor([0.9, 0.75, 0.6, 0.95], device="cuda")
    keep, n_keep = ops.nms(b, sc, 3, 0.5)
    assert keep.cpu().tolist() == [3, 0, -1] and n_keep.cpu().tolist() == [2, 1]
    assert dd.non_max_suppression(b, sc, 3, 0.5).cpu().tolist() == [3, 0]
    # degenerate (zero-area) boxes never suppress and are never suppressed
    z = torch.tensor([[0.5, 0.5, 0.5, 0.5]] * 3, dtype=torch.float32, device="cuda")
    assert dd.non_max_suppression(z, torch.tensor([0.3, 0.2, 0.1], device="cuda"), 3, 0.5).cpu().tolist() == [0, 1, 2]


@pytest.mark.parametrize("n,levels", [(9000, 1), (9000, 3), (7000, 40), (3073, 2), (20000, 1500), (100, 1)])
def test_nms_score_ties_across_chunks(dd, n, levels):
    """Heavily tied scores: candidates are visited in (descending score, ascending index) order —
    the stable order of the oracle — also where a run of equal scores straddles the 3072-candidate
    chunks that the radix selection orders lazily (index digits of the 64-bit key decide)."""
    rng = np.random.default_rng(n + levels)
    boxes, _ = _nms_case(rng, n, max(n // 3, 1))
    scores = (rng.integers(0, levels, n) / max(levels, 1)).astype(np.float32)
    for max_out, thr in ((n, 0.5), (700, 0.8)):
        np.testing.assert_array_equal(dd.non_max_suppression(boxes, scores, max_out, thr),
                                      O.non_max_suppression(boxes, scores, max_out, thr))


def test_nms_special_scores_and_device_count(dd):
    """Negative scores, zeros of both signs and infinities order like the oracle's stable sort on
    -scores; a device-side candidate count cuts the input at any position relative to the chunks."""
    from dodt_b200 import ops
    rng = np.random.default_rng(77)
    n = 10000
    boxes, scores = _nms_case(rng, n, 2500)
    scores = (scores - 0.5).astype(np.float32)
    scores[rng.choice(n, 50, replace=False)] = 0.0
    scores[rng.choice(n, 5, replace=False)] = np.inf
    scores[rng.choice(n, 5, replace=False)] = -np.inf
    np.testing.assert_array_equal(dd.non_max_suppression(boxes, scores, 2000, 0.6),
                                  O.non_max_suppression(boxes, scores, 2000, 0.6))
    b, s = torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda()
    for cut in (1, 1535, 3072, 3073, 6144, 6145, 9999):
        n_dev = torch.tensor([cut], dtype=torch.int32, device="cuda")
        keep, n_keep = ops.nms(b, s, 1500, 0.6, n_dev=n_dev)
        want = O.non_max_suppression(boxes[:cut], scores[:cut], 1500, 0.6)
        assert n_keep.cpu().tolist() == [len(want), 1], cut
        np.testing.assert_array_equal(keep[:len(want)].cpu().numpy(), want)


@pytest.mark.parametrize("seed", range(10))
def test_bev_and_filter_adversarial_random_configs(dd, seed):
    """Randomised configurations (voxel size, slice count and range, extents, point dtype) with
    clouds built to sit on every rounding edge: points snapped to voxel multiples, to slice
    boundaries +- one ulp, to the open extents, duplicates, and points outside. S1 maps, winner
    indices, counts and occupancy must equal the oracle bit for bit; the S2 keep mask on random
    (partly out-of-range, partly negative) anchors likewise."""
    from oracle import synth_ref as synth
    from adversarial import adversarial_case
    rng = np.random.default_rng(1900 + seed)
    pc, voxel, ext, lo, hi, S = adversarial_case(seed)
    buf, ref, cfg = _run_bev(dd, pc, voxel=voxel, extents=ext, lo=lo, hi=hi, S=S)
    _check_bev(buf, ref, S)
    pc64 = np.asarray(pc, dtype=np.float64)
    try:
        occ, vox = O.occupancy_grid(pc64, synth.GROUND_PLANE, ext, voxel)
    except IndexError:
        return                                                    # no point in the 0.2-2.0 m slice
    np.testing.assert_array_equal(buf.occ.cpu().numpy(), occ)
    a = np.stack([rng.uniform(ext[0][0] - 6, ext[0][1] + 6, 4000), np.zeros(4000),
                  rng.uniform(ext[2][0] - 6, ext[2][1] + 6, 4000), rng.uniform(0.05, 6.0, 4000),
                  np.ones(4000), rng.uniform(0.05, 6.0, 4000)], 1)
    if seed % 3 == 0:
        a = a.astype(np.float32)
    grid = dd.VoxelGrid2D.from_occupancy(buf.occ, voxel, ext)
    for thr in (1, 2):
        np.testing.assert_array_equal(dd.get_empty_anchor_filter_2d(a, grid, thr),
                                      O.empty_anchor_filter_2d(a, occ, voxel, vox["min_coord"][[0, 2]], thr))
