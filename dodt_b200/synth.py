"""Seeded synthetic KITTI-shaped inputs for tests and bench.py; NumPy, host side.

Everything is a pure function of (config, frame): rng = default_rng(1000 * config + frame).
"""
import numpy as np

from . import anchors as A

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


def car_anchors(area_extents=AREA_EXTENTS, ground_plane=GROUND_PLANE):
    """The 89 600-anchor Car grid in anchor form (N, 6) float64 (dt_rpn_model.py:913,950)."""
    boxes = A.tile_anchors_3d(area_extents, A.CAR_ANCHOR_SIZES, ANCHOR_STRIDE, ground_plane)
    return A.box_3d_to_anchor(boxes)


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


def rpn_proposals(config, frame, anchors_kept):
    """Regressed anchors, their normalised BEV boxes [x1,z1,x2,z2] (what dt_rpn_model.py:573-591
    hands to NMS) and tie-free scores."""
    rng = np.random.default_rng(1000 * config + frame + 700000)
    n = len(anchors_kept)
    offsets = rng.normal(0.0, 0.1, (n, 6)).astype(np.float32).astype(np.float64)   # == rpn_offsets
    regressed = A.offset_to_anchor(anchors_kept, offsets)
    _, bev_norm = A.project_to_bev(regressed, BEV_EXTENTS)
    scores = rng.permutation(np.linspace(0.01, 0.99, n)).astype(np.float32)
    return regressed, bev_norm.astype(np.float32), scores


def crop_boxes(anchors, image_shape=IMAGE_SHAPE):
    """Normalised [y1,x1,y2,x2] float32 boxes on the BEV map and on the image for a set of anchors
    (dt_rpn_model.py:975-985)."""
    _, bev_norm = A.project_to_bev(anchors, BEV_EXTENTS)
    _, img_norm = A.project_to_image_space(anchors, A.KITTI_P2, image_shape)
    return (A.reorder_projected_boxes(bev_norm).astype(np.float32),
            A.reorder_projected_boxes(img_norm).astype(np.float32))


_ANCHOR_CACHE = None


def anchor_set():
    """(anchors (N,6) f64, their BEV boxes, their image boxes — both [y1,x1,y2,x2] f32)."""
    global _ANCHOR_CACHE
    if _ANCHOR_CACHE is None:
        a = car_anchors()
        _, bev_norm = A.project_to_bev(a, BEV_EXTENTS)
        _, img_norm = A.project_to_image_space(a, A.KITTI_P2, IMAGE_SHAPE)
        _ANCHOR_CACHE = (a, A.reorder_projected_boxes(bev_norm).astype(np.float32),
                         A.reorder_projected_boxes(img_norm).astype(np.float32))
    return _ANCHOR_CACHE


def frame_inputs(config, frame, n_points=120000, rpn_nms_size=1024):
    """Host inputs of one frame slot (dodt_b200.frontend.FrameSlot): synthetic sensor data and
    synthetic network-head outputs, a pure function of (config, frame). `rpn_offsets` is what the
    slot consumes; `rpn_boxes` / `rpn_img_boxes` are the same offsets decoded and projected on the
    host (the reference's NumPy chain), for the CPU oracle."""
    a, _, _ = anchor_set()
    rng = np.random.default_rng(1000 * config + frame + 900000)
    bev_feat, _ = feature_pair(config, frame)                          # [1,700,800,32]
    img_feat = np.abs(rng.standard_normal((1,) + tuple(IMAGE_SHAPE) + (32,), dtype=np.float32))
    regressed, bev_norm, scores = rpn_proposals(config, frame, a)
    _, img_norm = A.project_to_image_space(regressed, A.KITTI_P2, IMAGE_SHAPE)
    return dict(
        points=point_cloud(config, frame, n_points),
        bev_feat=bev_feat, img_feat=img_feat,
        bev_1ch=np.ascontiguousarray(bev_feat[..., :1]) * np.float32(0.5),
        img_1ch=np.ascontiguousarray(img_feat[..., :1]) * np.float32(0.5),
        rpn_offsets=rpn_offsets(config, frame, len(a)),
        rpn_boxes=A.reorder_projected_boxes(bev_norm).astype(np.float32),
        rpn_img_boxes=A.reorder_projected_boxes(img_norm).astype(np.float32),
        rpn_scores=scores,
        final_scores=rng.permutation(np.linspace(0.01, 0.99, rpn_nms_size)).astype(np.float32))


def clustered_rpn_outputs(config, frame, n_targets=40, per_target=300, jitter=0.01):
    """RPN head outputs shaped like a trained network's: the `per_target` anchors nearest to each of
    `n_targets` objects regress onto (almost) the same box and carry the highest scores, so that at
    IoU 0.8 nearly all of the best n_targets * per_target candidates suppress one another and
    tf.image.non_max_suppression has to scan far down the score order to fill its 1024 outputs.
    Returns the keys of frame_inputs() it replaces: rpn_offsets, rpn_scores, rpn_boxes,
    rpn_img_boxes."""
    a, _, _ = anchor_set()
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
    offsets = offsets.astype(np.float32)
    n_cl = int(clustered.sum())
    scores = np.empty(n, dtype=np.float64)
    lin = np.linspace(0.01, 0.99, n)
    scores[clustered] = rng.permutation(lin[n - n_cl:])       # the clustered anchors score highest
    scores[~clustered] = rng.permutation(lin[:n - n_cl])
    regressed = A.offset_to_anchor(a, offsets.astype(np.float64))
    _, bev_norm = A.project_to_bev(regressed, BEV_EXTENTS)
    _, img_norm = A.project_to_image_space(regressed, A.KITTI_P2, IMAGE_SHAPE)
    return dict(rpn_offsets=offsets, rpn_scores=scores.astype(np.float32),
                rpn_boxes=A.reorder_projected_boxes(bev_norm).astype(np.float32),
                rpn_img_boxes=A.reorder_projected_boxes(img_norm).astype(np.float32))
