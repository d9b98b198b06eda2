"""Generates tests/golden/*.npz from the REFERENCE's own Python (imported from the read-only
checkout through oracle/ref_shim.py) and from independent implementations for the two TensorFlow
ops. TEST INFRASTRUCTURE; runs only where /root/reference exists (this container), the fixtures it
writes are committed and used everywhere else.

  python -m oracle.make_golden

Fixtures
  s1s2_kitti_<seq>_<frame>.npz   real KITTI tracking frames of the reference's unit-test dataset
        (avod/tests/datasets/Kitti/tracking): camera-frame FOV-cropped float64 cloud exactly as
        wavedata tracking_utils.get_lidar_point_cloud produces it, and the reference outputs:
        BevSlices.generate_bev (6 maps, stored sparse), the leaf layout of the 0.2-2.0 m slice
        (kitti_utils.create_sliced_voxel_grid_2d) and get_empty_anchor_filter_2d on the 89 600 Car
        anchors (GridAnchor3dGenerator + box_3d_to_anchor).
  lidar_kitti_000003.npz         raw velodyne scan + calibration matrices of that frame, the
        frustum-cropped cloud wavedata tracking_utils.get_lidar_point_cloud makes of it and every
        7th point of the uncropped camera-frame cloud.
  s1s2_synth.npz                 the same on a 30 k-point synthetic cloud incl. degenerate slices.
  s1_unit_vectors.npz            outputs of the reference's own unit-test inputs
        (voxel_grid_2d_test.py, integral_image_2d_test.py, kitti_utils_test.py, obj_utils_test.py).
  s3_grid_sample.npz             torch.nn.functional.grid_sample(align_corners=True) on in-bounds
        boxes — independent check of the crop_and_resize restatement (TF 1.3 is not available).
  s5_torchvision_nms.npz         torchvision.ops.nms on tie-free boxes — independent check of the
        non_max_suppression restatement (same "IoU > thr" rule).
  s4_shift_formulation.npz       correlation by explicit shifted products in float64.
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import anchor_helpers as A  # noqa: E402
from oracle import synth_ref as S  # noqa: E402
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def sparse(m):
    r, c = np.nonzero(m)
    return r.astype(np.int32), c.astype(np.int32), m[r, c]


def reference_s1s2(pc):
    """pc (3, N) float64 -> dict of reference outputs."""
    from avod.core import anchor_filter, box_3d_encoder
    from avod.core.anchor_generators import grid_anchor_3d_generator as G
    gen = ref_shim.reference_bev_slices(S.HEIGHT_LO, S.HEIGHT_HI, S.NUM_SLICES)
    bev = gen.generate_bev('lidar', pc, S.GROUND_PLANE, S.AREA_EXTENTS, S.VOXEL_SIZE)
    out = {}
    for i, m in enumerate(bev['height_maps'] + [bev['density_map']]):
        out["map%d_r" % i], out["map%d_c" % i], out["map%d_v" % i] = sparse(m)
    vg = ref_shim.reference_sliced_voxel_grid_2d(pc, S.GROUND_PLANE, S.AREA_EXTENTS, S.VOXEL_SIZE)
    occ = (np.squeeze(vg.leaf_layout_2d) + 1).astype(np.uint8)
    out["occ_x"], out["occ_z"] = [a.astype(np.int32) for a in np.nonzero(occ)]
    boxes = G.tile_anchors_3d(S.AREA_EXTENTS, A.CAR_ANCHOR_SIZES, S.ANCHOR_STRIDE, S.GROUND_PLANE)
    anchors = box_3d_encoder.box_3d_to_anchor(boxes)
    assert np.array_equal(anchors, S.car_anchors())
    keep = anchor_filter.get_empty_anchor_filter_2d(anchors, vg, 1)
    out["keep_packed"] = np.packbits(keep)
    out["n_anchors"] = np.int64(len(keep))
    out["points"] = pc
    return out


def kitti_frames():
    from wavedata.tools.obj_detection import tracking_utils
    base = os.path.join(ref_shim.REFERENCE_ROOT, "avod/tests/datasets/Kitti/tracking/training")
    for name in ("000003", "010005"):
        pc = tracking_utils.get_lidar_point_cloud(name, base + "/calib", base + "/velodyne",
                                                  im_size=[1242, 375])
        out = reference_s1s2(np.asarray(pc, dtype=np.float64))
        np.savez_compressed(os.path.join(OUT, "s1s2_kitti_%s.npz" % name), **out)
        print("kitti", name, pc.shape, "kept", int(np.unpackbits(out["keep_packed"])[:89600].sum()))


def lidar_frame():
    """Raw velodyne scan + calibration of one fixture frame and what the reference makes of it."""
    from wavedata.tools.core import calib_utils
    from wavedata.tools.obj_detection import tracking_utils
    base = os.path.join(ref_shim.REFERENCE_ROOT, "avod/tests/datasets/Kitti/tracking/training")
    name = "000003"
    calib = calib_utils.read_tracking_calibration(base + "/calib", 0)
    x, y, z, i = calib_utils.read_lidar(base + "/velodyne/0000", 3)
    velo = np.stack([x, y, z, i], axis=1).astype(np.float32)
    fov = tracking_utils.get_lidar_point_cloud(name, base + "/calib", base + "/velodyne", im_size=[1242, 375])
    full = tracking_utils.get_lidar_point_cloud(name, base + "/calib", base + "/velodyne")
    np.savez_compressed(os.path.join(OUT, "lidar_kitti_000003.npz"), velo=velo, p2=calib.p2,
                        r0_rect=calib.r0_rect, tr_velodyne_to_cam=calib.tr_velodyne_to_cam,
                        fov=np.asarray(fov), full_every_7th=np.asarray(full)[:, ::7], im_size=np.array([1242, 375]))
    print("lidar", velo.shape, "->", np.asarray(fov).shape)


def lidar_pair():
    """DODT's frame pair ingest: the scan of frame t+1 moved into frame t's LiDAR frame with the OXTS
    records (KittiTrackingDataset.coordinate_transform / point_cloud_transform, run unbound on a
    stand-in that only has oxts_dir — the real constructor needs the protobuf configs), then the
    frustum crop of KittiTrackingUtils.transfer_lidar_to_camera_view."""
    import types
    from wavedata.tools.core import calib_utils
    from wavedata.tools.obj_detection import tracking_utils
    from avod.datasets.kitti.kitti_tracking_dataset import KittiTrackingDataset
    base = os.path.join(ref_shim.REFERENCE_ROOT, "avod/tests/datasets/Kitti/tracking/training")
    names = ["000003", "000004"]
    ds = types.SimpleNamespace(oxts_dir=base + "/oxts")
    for m in ("get_oxts", "coordinate_transform", "point_cloud_transform"):
        setattr(ds, m, types.MethodType(getattr(KittiTrackingDataset, m), ds))
    trans, matrix, delta = ds.coordinate_transform(names)
    raw = [tracking_utils.get_raw_lidar_point_cloud(n, base + "/velodyne") for n in names]
    raw[1] = np.ascontiguousarray(raw[1][:, ::4])            # every 4th point keeps the fixture small
    velo1 = raw[1].T.copy()
    moved = ds.point_cloud_transform([raw[0], raw[1].copy()], names)[1]
    assert moved.dtype == np.float32
    fov = tracking_utils.get_lidar_in_camera_view(moved.copy(), names[1], base + "/calib", im_size=[1242, 375])
    calib = calib_utils.read_tracking_calibration(base + "/calib", 0)
    lines = [open(base + "/oxts/0000.txt").read().splitlines()[i] for i in (3, 4)]
    oxts = np.array([[float(v) for v in ln.split()[:6]] for ln in lines])
    np.savez_compressed(os.path.join(OUT, "lidar_pair_000003_000004.npz"), velo1=velo1, oxts=oxts,
                        oxts_line0=lines[0], oxts_line1=lines[1], trans=trans, matrix=matrix, delta=delta,
                        moved=np.asarray(moved), fov=np.asarray(fov), p2=calib.p2, r0_rect=calib.r0_rect,
                        tr_velodyne_to_cam=calib.tr_velodyne_to_cam, im_size=np.array([1242, 375]))
    print("lidar pair", velo1.shape, "trans", trans, "->", np.asarray(fov).shape)


def synth_frame():
    pc = S.point_cloud(7, 0, n_points=30000).astype(np.float64)
    # make slice 4 degenerate (a single point) and put points exactly on slice boundaries
    h = 1.65 - pc[1]
    pc = pc[:, ~((h > 1.8) & (h <= 2.3))]
    pc = np.concatenate([pc, np.array([[5.0, 1.65 - 2.0, 20.0], [1.0, 1.65 - 0.3, 10.0],
                                       [1.0, 1.65 - 0.8, 10.0]]).T], axis=1)
    out = reference_s1s2(pc)
    np.savez_compressed(os.path.join(OUT, "s1s2_synth.npz"), **out)
    print("synth", pc.shape)


def unit_vectors():
    from wavedata.tools.core.integral_image_2d import IntegralImage2D
    from wavedata.tools.core.voxel_grid_2d import VoxelGrid2D
    from wavedata.tools.obj_detection import obj_utils
    out = {}
    pts = np.array([[-39.99, 4.99, 0], [39.99, 4.99, 0], [-39.99, -4.99, 0], [39.99, -4.99, 0],
                    [-39.99, 4.99, 69.99], [39.99, 4.99, 69.99], [-39.99, -4.99, 69.99],
                    [39.99, -4.99, 69.99], [-39.99, 4.99, 69.99], [39.99, 4.99, 69.99],
                    [-39.99, -4.99, 69.99], [39.99, -4.99, 69.99]])
    vg = VoxelGrid2D()
    vg.voxelize_2d(pts, 0.1)
    out["vg_pts"] = pts
    out["vg_voxel_indices"] = vg.voxel_indices
    out["vg_heights"] = vg.heights
    out["vg_counts"] = vg.num_pts_in_voxel
    out["vg_num_divisions"] = vg.num_divisions
    out["vg_min"] = vg.min_voxel_coord
    rng = np.random.default_rng(0)
    points = (rng.random((5000, 3)) * [80, 8, 60]) - [40, 4, 0]
    ext = np.array([[-50, 50], [-5, 5], [0, 70]])
    vg = VoxelGrid2D()
    vg.voxelize_2d(points, 0.1, ext, ground_plane=[0, -1, 0, 1.65])
    out["vg2_pts"] = points
    out["vg2_voxel_indices"] = vg.voxel_indices
    out["vg2_heights"] = vg.heights
    out["vg2_counts"] = vg.num_pts_in_voxel
    coords = np.array([[0, 0], [0.1, 0.1], [-50, 0], [50, 70], [60, 80], [-12.34, 33.3]])
    out["map_coords"] = coords
    out["map_index"] = np.asarray(vg.map_to_index(coords))
    out["map_index_f32"] = np.asarray(vg.map_to_index(coords.astype(np.float32)))
    img = (rng.random((37, 53)) < 0.2).astype(np.float64)
    ii = IntegralImage2D(img)
    boxes = np.stack([rng.integers(0, 40, 64), rng.integers(0, 60, 64), rng.integers(0, 2400, 64),
                      rng.integers(0, 200, 64)]).astype(np.uint32)
    out["ii_img"] = img
    out["ii_boxes"] = boxes
    out["ii_query"] = ii.query(boxes)
    out["ii_image"] = ii._integral_image
    pc = np.array([[1.0, 1.0, 1.0], [0.0, 1.0, 3.0], [1.0, 1.0, 1.0]])
    out["sf_pc"] = pc
    out["sf_filter"] = ref_shim.KittiUtilsStandIn().create_slice_filter(
        pc, [[-2, 2], [-5, 5], [-2, 2]], [0, 1, 0, 0], 0.2, 2.0)
    cloud = rng.uniform(-3, 3, (3, 200))
    out["pf_pc"] = cloud
    out["pf_extents_only"] = obj_utils.get_point_filter(cloud, [[-2, 2], [-1, 1], [-2, 2]])
    out["pf_plane"] = obj_utils.get_point_filter(cloud, [[-2, 2], [-2, 2], [-2, 2]], [0, -1, 0, 1.65], 1.0)
    np.savez_compressed(os.path.join(OUT, "s1_unit_vectors.npz"), **out)
    print("unit vectors ok")


def independent_tf_ops():
    """Runs in a subprocess: torchvision does not import next to the stub tensorflow module."""
    code = r'''
import numpy as np, torch, torch.nn.functional as F, torchvision
rng = np.random.default_rng(11)
img = rng.standard_normal((1, 40, 56, 6)).astype(np.float32)
N = 96
c = rng.uniform(0.25, 0.75, (N, 2)); h = rng.uniform(0.02, 0.2, (N, 2))
boxes = np.stack([c[:,0]-h[:,0], c[:,1]-h[:,1], c[:,0]+h[:,0], c[:,1]+h[:,1]], 1).astype(np.float32)
res = {}
for ch, cw in ((7, 7), (3, 3), (2, 5)):
    ty = np.linspace(0, 1, ch); tx = np.linspace(0, 1, cw)
    ys = boxes[:, 0:1].astype(np.float64) + ty[None] * (boxes[:, 2:3] - boxes[:, 0:1]).astype(np.float64)
    xs = boxes[:, 1:2].astype(np.float64) + tx[None] * (boxes[:, 3:4] - boxes[:, 1:2]).astype(np.float64)
    grid = np.stack(np.broadcast_arrays(xs[:, None, :] * 2 - 1, ys[:, :, None] * 2 - 1), -1)
    t = F.grid_sample(torch.from_numpy(img).permute(0, 3, 1, 2).double().expand(N, -1, -1, -1),
                      torch.from_numpy(grid).double(), mode="bilinear", padding_mode="zeros", align_corners=True)
    res["crop_%dx%d" % (ch, cw)] = t.permute(0, 2, 3, 1).numpy()
np.savez_compressed(OUT + "/s3_grid_sample.npz", image=img, boxes=boxes, **res)
out = {}
for k, (n, spread, thr) in enumerate(((400, 0.02, 0.5), (3000, 0.004, 0.8), (1000, 0.01, 0.01))):
    centres = rng.uniform(0.1, 0.9, (40, 2))
    cc = centres[rng.integers(0, 40, n)] + rng.normal(0, spread, (n, 2))
    hh = np.abs(rng.normal(0.03, 0.004, (n, 2))) + 0.002
    b = np.stack([cc[:,0]-hh[:,0], cc[:,1]-hh[:,1], cc[:,0]+hh[:,0], cc[:,1]+hh[:,1]], 1).astype(np.float32)
    s = rng.permutation(np.linspace(0, 1, n)).astype(np.float32)
    keep = torchvision.ops.nms(torch.from_numpy(b), torch.from_numpy(s), thr).numpy()
    out["boxes%d" % k], out["scores%d" % k], out["thr%d" % k], out["keep%d" % k] = b, s, np.float32(thr), keep
np.savez_compressed(OUT + "/s5_torchvision_nms.npz", **out)
a = np.abs(rng.standard_normal((1, 18, 22, 32))).astype(np.float32)
b = np.abs(rng.standard_normal((1, 18, 22, 32))).astype(np.float32)
bp = np.pad(b, ((0, 0), (4, 4), (4, 4), (0, 0))).astype(np.float64)
ref = np.zeros((1, 18, 22, 25))
for k in range(25):
    p, o = k // 5 - 2, k % 5 - 2
    ref[..., k] = (a.astype(np.float64) * bp[:, 4 + 2 * p:4 + 2 * p + 18, 4 + 2 * o:4 + 2 * o + 22]).sum(-1) / 32
np.savez_compressed(OUT + "/s4_shift_formulation.npz", a=a, b=b, out=ref)
print("independent ops ok", torchvision.__version__)
'''
    subprocess.run([sys.executable, "-c", "OUT=%r\n" % OUT + code], check=True)


if __name__ == "__main__":
    assert ref_shim.install(), "the reference checkout is needed to generate golden vectors"
    os.makedirs(OUT, exist_ok=True)
    kitti_frames()
    lidar_frame()
    lidar_pair()
    synth_frame()
    unit_vectors()
    independent_tf_ops()
    for f in sorted(os.listdir(OUT)):
        print("%8d  %s" % (os.path.getsize(os.path.join(OUT, f)), f))
