"""Oracle against the LIVE reference (S1/S2 and the host-side anchor helpers): the reference's own
Python is imported from the read-only checkout through oracle/ref_shim.py and run on seeded random
inputs next to the restatement. Skipped where the checkout does not exist (the GPU box); the frozen
outputs under tests/golden/ (tests/test_oracle_golden.py) cover that case.
"""
import os

import numpy as np
import pytest

from oracle import anchor_helpers as A
from oracle import synth_ref as S
from oracle import np_oracle as O
from oracle import ref_shim

S_GROUND = S.GROUND_PLANE

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference checkout not present")


@pytest.fixture(scope="module", autouse=True)
def _ref():
    assert ref_shim.install()


def _cloud(seed, n, ground=0.4):
    return S.point_cloud(40 + seed, seed, n_points=n).astype(np.float64)


@pytest.mark.parametrize("seed,n", [(0, 20000), (1, 3000), (2, 60000)])
def test_bev_slices_live(seed, n):
    pc = _cloud(seed, n)
    gen = ref_shim.reference_bev_slices(S.HEIGHT_LO, S.HEIGHT_HI, S.NUM_SLICES)
    ref = gen.generate_bev('lidar', pc, S.GROUND_PLANE, S.AREA_EXTENTS, S.VOXEL_SIZE)
    got = O.bev_slices(pc, S.GROUND_PLANE, S.AREA_EXTENTS, S.VOXEL_SIZE, S.HEIGHT_LO, S.HEIGHT_HI,
                       S.NUM_SLICES)
    for a, b in zip(ref['height_maps'] + [ref['density_map']], got['height_maps'] + [got['density_map']]):
        np.testing.assert_array_equal(a, b)


def test_bev_slices_live_other_config():
    """3 slices, 0.25 m voxels, tilted plane: nothing in the restatement is specialised to config A."""
    pc = _cloud(3, 15000)
    plane = [0.01, -0.9995, 0.02, 1.7]
    ext = [[-30, 30], [-4, 3], [5, 65]]
    pc = pc[:, (pc[2] > 5.5) & (np.abs(pc[0]) < 29.5) & (pc[2] < 64.5)]
    gen = ref_shim.reference_bev_slices(0.0, 1.5, 3)
    ref = gen.generate_bev('lidar', pc, plane, ext, 0.25)
    got = O.bev_slices(pc, plane, ext, 0.25, 0.0, 1.5, 3)
    for a, b in zip(ref['height_maps'] + [ref['density_map']], got['height_maps'] + [got['density_map']]):
        np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("seed", range(10))
def test_adversarial_configs_live(seed):
    """The edge-case clouds of tests/adversarial.py (voxel edges, slice boundaries +- ulp, open
    extents, duplicates; random voxel size / slices / extents): oracle == live reference, so the
    GPU test that compares against the oracle on the same cases is pinned to the reference."""
    from adversarial import adversarial_case
    from avod.core import anchor_filter
    pc, voxel, ext, lo, hi, S = adversarial_case(seed)
    pc = np.asarray(pc, dtype=np.float64)     # the reference pipeline hands float64
    gen = ref_shim.reference_bev_slices(lo, hi, S)
    ref = gen.generate_bev('lidar', pc, S_GROUND, ext, voxel)
    got = O.bev_slices(pc, S_GROUND, ext, voxel, lo, hi, S)
    for a, b in zip(ref['height_maps'] + [ref['density_map']], got['height_maps'] + [got['density_map']]):
        np.testing.assert_array_equal(a, b)
    try:
        vg = ref_shim.reference_sliced_voxel_grid_2d(pc, S_GROUND, ext, voxel)
    except IndexError:
        with pytest.raises(IndexError):
            O.occupancy_grid(pc, S_GROUND, ext, voxel)
        return
    occ, vox = O.occupancy_grid(pc, S_GROUND, ext, voxel)
    np.testing.assert_array_equal(np.squeeze(vg.leaf_layout_2d, 1) + 1, occ)
    rng = np.random.default_rng(1900 + seed)
    a = np.stack([rng.uniform(ext[0][0] - 6, ext[0][1] + 6, 4000), np.zeros(4000),
                  rng.uniform(ext[2][0] - 6, ext[2][1] + 6, 4000), rng.uniform(0.05, 6.0, 4000),
                  np.ones(4000), rng.uniform(0.05, 6.0, 4000)], 1)
    for thr in (1, 2):
        np.testing.assert_array_equal(O.empty_anchor_filter_2d(a, occ, voxel, vox["min_coord"][[0, 2]], thr),
                                      anchor_filter.get_empty_anchor_filter_2d(a, vg, thr))


@pytest.mark.parametrize("seed,n,thr", [(0, 20000, 1), (4, 5000, 1), (5, 40000, 3)])
def test_anchor_filter_live(seed, n, thr):
    from avod.core import anchor_filter
    pc = _cloud(seed, n)
    vg = ref_shim.reference_sliced_voxel_grid_2d(pc, S.GROUND_PLANE, S.AREA_EXTENTS, S.VOXEL_SIZE)
    anchors = S.car_anchors()
    want = anchor_filter.get_empty_anchor_filter_2d(anchors, vg, thr)
    occ, vox = O.occupancy_grid(pc, S.GROUND_PLANE, S.AREA_EXTENTS, S.VOXEL_SIZE)
    np.testing.assert_array_equal(np.squeeze(vg.leaf_layout_2d) + 1, occ)
    got = O.empty_anchor_filter_2d(anchors, occ, S.VOXEL_SIZE, vox["min_coord"][[0, 2]], thr)
    np.testing.assert_array_equal(got, want)


def test_voxelize_2d_live():
    from wavedata.tools.core.voxel_grid_2d import VoxelGrid2D
    rng = np.random.default_rng(12)
    pts = (rng.random((30000, 3)) * [80, 8, 70]) - [40, 4, 0]
    for ext, plane in ((None, None), (np.array([[-50, 50], [-5, 5], [0, 70]]), [0, -1, 0, 1.65])):
        vg = VoxelGrid2D()
        vg.voxelize_2d(pts, 0.1, ext, plane)
        v = O.voxelize_2d(pts, 0.1, ext, plane)
        np.testing.assert_array_equal(v["voxel_indices"], vg.voxel_indices)
        np.testing.assert_array_equal(v["heights"], vg.heights)
        np.testing.assert_array_equal(v["counts"], vg.num_pts_in_voxel)
        np.testing.assert_array_equal(O.leaf_layout_2d(v), vg.leaf_layout_2d)


def test_anchor_helpers_live():
    """oracle.anchor_helpers (host-side inputs of S2/S3/S5) == the reference helpers."""
    from avod.core import anchor_encoder, anchor_projector, box_3d_encoder
    from avod.core.anchor_generators import grid_anchor_3d_generator as G
    boxes = G.tile_anchors_3d(S.AREA_EXTENTS, A.CAR_ANCHOR_SIZES, S.ANCHOR_STRIDE, S.GROUND_PLANE)
    np.testing.assert_array_equal(A.tile_anchors_3d(S.AREA_EXTENTS, A.CAR_ANCHOR_SIZES, S.ANCHOR_STRIDE,
                                                    S.GROUND_PLANE), boxes)
    anchors = box_3d_encoder.box_3d_to_anchor(boxes)
    np.testing.assert_array_equal(A.box_3d_to_anchor(boxes), anchors)
    sub = anchors[::97]
    for got, want in zip(A.project_to_bev(sub, S.BEV_EXTENTS), anchor_projector.project_to_bev(sub, S.BEV_EXTENTS)):
        np.testing.assert_array_equal(got, want)
    for got, want in zip(A.project_to_image_space(sub, A.KITTI_P2, S.IMAGE_SHAPE),
                         anchor_projector.project_to_image_space(sub, A.KITTI_P2, S.IMAGE_SHAPE)):
        np.testing.assert_array_equal(got, want)
    off = np.random.default_rng(3).normal(0, 0.1, sub.shape)
    np.testing.assert_array_equal(A.offset_to_anchor(sub, off), anchor_encoder.offset_to_anchor(sub, off))
    # anchor_projector.reorder_projected_boxes (:254-273) is a tf.stack of the columns
    # [y1, x1, y2, x2] <- [x1, y1, x2, y2]; it cannot run without TensorFlow
    b = np.random.default_rng(4).random((9, 4))
    np.testing.assert_array_equal(A.reorder_projected_boxes(b), b[:, [1, 0, 3, 2]])


def test_oxts_alignment_live():
    """Oxts arithmetic and point_cloud_transform of the checkout on random records / scans == the
    oracle restatement and the product's host mirror (needs the whole checkout, not the staged S1/S2)."""
    if not ref_shim.full_checkout():
        pytest.skip("needs avod.datasets of the full checkout")
    import types
    from avod.datasets.kitti.kitti_tracking_utils import Oxts
    from avod.datasets.kitti.kitti_tracking_dataset import KittiTrackingDataset
    from dodt_b200 import lidar
    rng = np.random.default_rng(11)
    for _ in range(20):
        rec = np.concatenate([[49.0 + rng.normal(0, 1e-4), 8.4 + rng.normal(0, 1e-4), 115.0], rng.normal(0, 0.05, 3)])
        nxt = rec + np.concatenate([rng.normal(0, 1e-5, 2), [0.0], rng.normal(0, 0.01, 3)])
        lines = [" ".join(repr(float(v)) for v in np.concatenate([r, np.zeros(24)])) for r in (rec, nxt)]
        ds = types.SimpleNamespace(get_oxts=lambda name: Oxts(lines[int(name)]))
        ds.coordinate_transform = types.MethodType(KittiTrackingDataset.coordinate_transform, ds)
        want = ds.coordinate_transform(["0", "1"])
        for got in (O.oxts_coordinate_transform(rec, nxt), lidar.coordinate_transform(lines[0], lines[1])):
            for a, b in zip(got, want):
                np.testing.assert_array_equal(np.asarray(a), np.asarray(b))
        pc = rng.uniform(-60, 60, (4, 500)).astype(np.float32)
        moved = types.MethodType(KittiTrackingDataset.point_cloud_transform, ds)([pc.copy(), pc.copy()], ["0", "1"])[1]
        np.testing.assert_array_equal(O.point_cloud_transform(pc, want[0], want[1]), moved)


def test_tracking_file_readers_live():
    """The product's calibration / oxts / velodyne readers == the reference's on its own fixtures."""
    if not ref_shim.full_checkout():
        pytest.skip("needs the fixture data of the full checkout")
    from wavedata.tools.core import calib_utils
    from wavedata.tools.obj_detection import tracking_utils
    from avod.datasets.kitti.kitti_tracking_utils import Oxts
    from dodt_b200 import tracking_utils as T
    base = os.path.join(ref_shim.REFERENCE_ROOT, "avod/tests/datasets/Kitti/tracking/training")
    for video in (0, 1):
        want, got = calib_utils.read_tracking_calibration(base + "/calib", video), \
            T.read_tracking_calibration(base + "/calib", video)
        for k in ("p0", "p1", "p2", "p3", "r0_rect", "tr_velodyne_to_cam"):
            np.testing.assert_array_equal(getattr(got, k), getattr(want, k))
    for name in ("000000", "000007", "010003"):
        np.testing.assert_array_equal(T.get_raw_lidar_point_cloud(name, base + "/velodyne"),
                                      tracking_utils.get_raw_lidar_point_cloud(name, base + "/velodyne"))
        line = open(base + "/oxts/%04d.txt" % int(name[:2])).read().splitlines()[int(name[2:])]
        a, b = T.get_oxts(base + "/oxts", name), Oxts(line)
        assert (a.latitude, a.longitude, a.altitude, a.roll, a.pitch, a.yaw) == \
            (b.latitude, b.longitude, b.altitude, b.roll, b.pitch, b.yaw)
        np.testing.assert_array_equal(T.get_road_plane(name, base + "/planes"),
                                      tracking_utils.get_road_plane(name, base + "/planes"))
