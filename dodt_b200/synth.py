"""Seeded synthetic KITTI-shaped RAW inputs of the front end for bench.py and tests;
NumPy, host side: point clouds, feature maps and network-head outputs — what a frame slot consumes.
The host-decoded boxes the CPU oracle needs on top of these live in oracle/synth_ref.py.

Everything is a pure function of (config, frame): rng = default_rng(1000 * config + frame).
"""
import numpy as np

from .constants import CAR_ANCHOR_SIZES

GROUND_PLANE = [0, -1, 0, 1.65]                       # wavedata tracking_utils.py:239 (hard-coded)
AREA_EXTENTS = [[-40, 40], [-5, 3], [0, 70]]          # kitti_utils_config.area_extents
BEV_EXTENTS = [[-40, 40], [0, 70]]
VOXEL_SIZE = float(np.float32(0.1))                   # proto float fields are fp32
VOXEL_SIZE_DENSE = float(np.float32(0.05))
HEIGHT_LO = float(np.float32(-0.2))
HEIGHT_HI = float(np.float32(2.3))
NUM_SLICES = 5
ANCHOR_STRIDE = [0.5, 0.5]
IMAGE_SHAPE = (360, 1200)
BEV_SHAPE = (700, 800)


class SlicesConfig:
    """Stand-in for the `slices` protobuf message (avod/protos/kitti_utils.proto:31-32)."""

    def __init__(self, height_lo=HEIGHT_LO, height_hi=HEIGHT_HI, num_slices=NUM_SLICES):
        self.height_lo, self.height_hi, self.num_slices = height_lo, height_hi, num_slices


def point_cloud(config, frame, n_points=120000):
    """(3, n) float32 camera-frame points: x right, y down, z forward."""
    rng = np.random.default_rng(1000 * config + frame)
    r = 70.0 * np.sqrt(rng.random(n_points))
    th = rng.uniform(-0.7, 0.7, n_points)
    x = np.clip(r * np.sin(th), -39.99, 39.99)
    z = np.clip(r * np.cos(th), 0.01, 69.99)
    h = rng.uniform(-0.2, 2.3, n_points)
    ground = rng.random(n_points) < 0.40
    h[ground] = rng.normal(0.0, 0.03, int(ground.sum()))
    n_car = int(0.02 * n_points)
    car_idx = rng.choice(n_points, n_car, replace=False)
    centers = np.stack([rng.uniform(-30, 30, 30), rng.uniform(8, 60, 30)], axis=1)
    which = rng.integers(0, 30, n_car)
    x[car_idx] = centers[which, 0] + rng.uniform(-0.9, 0.9, n_car)
    z[car_idx] = centers[which, 1] + rng.uniform(-2.0, 2.0, n_car)
    h[car_idx] = rng.uniform(0.0, 1.5, n_car)
    y = 1.65 - h
    return np.stack([x, y, z]).astype(np.float32)


def feature_pair(config, frame, shape=(1, 700, 800, 32)):
    """Post-ReLU-like NHWC float32 features of frame t and t+1 (t shifted by (2,-2) px + noise)."""
    rng = np.random.default_rng(1000 * config + frame + 500000)
    f0 = np.abs(rng.standard_normal(shape, dtype=np.float32))
    f1 = np.roll(f0, shift=(2, -2), axis=(1, 2))
    f1 = np.abs(f1 + np.float32(0.1) * rng.standard_normal(shape, dtype=np.float32))
    return f0, f1.astype(np.float32)


def rpn_offsets(config, frame, n):
    """The RPN head's regression output for n anchors: float32 [n, 6] anchor-form offsets."""
    rng = np.random.default_rng(1000 * config + frame + 700000)
    return rng.normal(0.0, 0.1, (n, 6)).astype(np.float32)


def num_car_anchors(area_extents=AREA_EXTENTS):
    """Anchors of the Car grid: x columns x z rows x sizes x 2 rotations (89 600 for the car config;
    grid_anchor_3d_generator.py:61-85)."""
    nx = len(np.arange(area_extents[0][0] + ANCHOR_STRIDE[0] / 2.0, area_extents[0][1], step=ANCHOR_STRIDE[0]))
    nz = len(np.arange(area_extents[2][1] - ANCHOR_STRIDE[1] / 2.0, area_extents[2][0], step=-ANCHOR_STRIDE[1]))
    return nx * nz * len(CAR_ANCHOR_SIZES) * 2


def rpn_scores(config, frame, n):
    """The RPN head's objectness for n anchors: tie-free float32 scores (same random stream as
    rpn_offsets: the offsets are drawn first)."""
    rng = np.random.default_rng(1000 * config + frame + 700000)
    rng.normal(0.0, 0.1, (n, 6))
    return rng.permutation(np.linspace(0.01, 0.99, n)).astype(np.float32)


def frame_inputs(config, frame, n_points=120000, rpn_nms_size=1024):
    """Host inputs of one frame slot (dodt_b200.frontend.FrameSlot): synthetic sensor data and
    synthetic network-head outputs, a pure function of (config, frame)."""
    n = num_car_anchors()
    rng = np.random.default_rng(1000 * config + frame + 900000)
    bev_feat, _ = feature_pair(config, frame)                          # [1,700,800,32]
    img_feat = np.abs(rng.standard_normal((1,) + tuple(IMAGE_SHAPE) + (32,), dtype=np.float32))
    final_scores = rng.permutation(np.linspace(0.01, 0.99, rpn_nms_size)).astype(np.float32)
    return dict(
        points=point_cloud(config, frame, n_points),
        bev_feat=bev_feat, img_feat=img_feat,
        bev_1ch=np.ascontiguousarray(bev_feat[..., :1]) * np.float32(0.5),
        img_1ch=np.ascontiguousarray(img_feat[..., :1]) * np.float32(0.5),
        rpn_offsets=rpn_offsets(config, frame, n),
        rpn_scores=rpn_scores(config, frame, n),
        final_scores=final_scores)


def clustered_rpn_outputs(config, frame, anchors, n_targets=40, per_target=300, jitter=0.01):
    """RPN head outputs shaped like a trained network's: the `per_target` anchors nearest to each of
    `n_targets` objects regress onto (almost) the same box and carry the highest scores, so that at
    IoU 0.8 nearly all of the best n_targets * per_target candidates suppress one another and
    tf.image.non_max_suppression has to scan far down the score order to fill its 1024 outputs.
    anchors: (n, 6) float64 anchor grid. Returns the frame_inputs() keys it replaces:
    rpn_offsets (inverse of anchor_encoder.offset_to_anchor towards the targets), rpn_scores."""
    a = np.asarray(anchors)
    n = len(a)
    rng = np.random.default_rng(1000 * config + frame + 1100000)
    offsets = rng.normal(0.0, 0.1, (n, 6))
    tx = rng.uniform(-30, 30, n_targets)
    tz = rng.uniform(8, 60, n_targets)
    clustered = np.zeros(n, dtype=bool)
    for k in range(n_targets):
        d2 = (a[:, 0] - tx[k]) ** 2 + (a[:, 2] - tz[k]) ** 2
        near = np.argsort(d2, kind="stable")[:per_target]
        near = near[~clustered[near]]
        if len(near) == 0:
            continue
        clustered[near] = True
        target = np.array([tx[k], a[near[0], 1], tz[k], 3.9, 1.56, 1.6])
        o = np.empty((len(near), 6))
        o[:, 0:3] = (target[0:3] - a[near, 0:3]) / a[near, 3:6]
        o[:, 3:6] = np.log(target[3:6] / a[near, 3:6])
        offsets[near] = o + rng.normal(0.0, jitter, o.shape)
    n_cl = int(clustered.sum())
    scores = np.empty(n, dtype=np.float64)
    lin = np.linspace(0.01, 0.99, n)
    scores[clustered] = rng.permutation(lin[n - n_cl:])       # the clustered anchors score highest
    scores[~clustered] = rng.permutation(lin[:n - n_cl])
    return dict(rpn_offsets=offsets.astype(np.float32), rpn_scores=scores.astype(np.float32))


def point_cloud_kitti(config, frame, n_points=18500, n_objects=64):
    """(3, n) float32 camera-frame points with the OCCUPANCY of a real KITTI tracking frame after
    the camera-FOV crop (SURVEY 8(d) "realistic ranges": 16-20 k points, 1.7-3.3 k occupied cells in
    the 0.2-2.0 m filter slice, up to ~100 points in a cell, 9-15 k of the 89 600 anchors kept):
    ground returns on scan rings, and compact objects (cars, poles, walls, vegetation) that carry
    the points above 0.2 m — unlike point_cloud(), whose evenly spread points keep 60 k anchors."""
    rng = np.random.default_rng(1000 * config + frame + 1300000)
    n_ground = int(0.58 * n_points)
    n_obj = n_points - n_ground
    # ground: 64 beams, elevation -24.8..+2 deg, sensor 1.73 m above the road
    elev = np.deg2rad(np.linspace(-24.8, -1.2, 56))
    ring_r = 1.73 / np.tan(-elev)
    ring = rng.integers(0, len(ring_r), n_ground)
    r = ring_r[ring] * (1.0 + rng.normal(0, 0.004, n_ground))
    th = rng.uniform(-0.72, 0.72, n_ground)
    gx, gz = r * np.sin(th), r * np.cos(th)
    gh = rng.normal(0.0, 0.035, n_ground)
    # objects: centres in the field of view, a footprint of a few square metres, points ~ 1/range
    oz = rng.uniform(6, 68, n_objects)
    ox = rng.uniform(-1, 1, n_objects) * np.minimum(38.0, oz * 0.8)
    half = np.stack([rng.uniform(0.2, 1.6, n_objects), rng.uniform(0.1, 0.5, n_objects)], axis=1)
    top = rng.uniform(0.8, 2.2, n_objects)
    w = 1.0 / np.maximum(oz, 8.0) ** 1.5
    which = rng.choice(n_objects, n_obj, p=w / w.sum())
    px = ox[which] + rng.uniform(-1, 1, n_obj) * half[which, 0]
    pz = oz[which] + rng.uniform(-1, 1, n_obj) * half[which, 1]
    ph = rng.uniform(0.05, 1.0, n_obj) * top[which]
    x = np.clip(np.concatenate([gx, px]), -39.99, 39.99)
    z = np.clip(np.concatenate([gz, pz]), 0.01, 69.99)
    h = np.concatenate([gh, ph])
    order = rng.permutation(n_points)            # file order is not sorted by anything
    y = 1.65 - h
    return np.stack([x, y, z]).astype(np.float32)[:, order]
