"""Synthetic inputs WITH the host-decoded boxes the CPU oracle consumes — TEST INFRASTRUCTURE.

dodt_b200.synth produces the raw inputs of a frame slot; this module adds what the reference's
host code computes from them between the stages (anchor grid, anchor projections, decoded RPN
boxes; oracle/anchor_helpers.py). Everything of dodt_b200.synth is re-exported, so tests use this
module as a drop-in superset.
"""
import numpy as np

from dodt_b200.synth import *  # noqa: F401,F403
from dodt_b200 import synth as _raw
from dodt_b200.synth import (ANCHOR_STRIDE, AREA_EXTENTS, BEV_EXTENTS, GROUND_PLANE, IMAGE_SHAPE)

from . import anchor_helpers as A


def car_anchors(area_extents=AREA_EXTENTS, ground_plane=GROUND_PLANE):
    """The 89 600-anchor Car grid in anchor form (N, 6) float64 (dt_rpn_model.py:913,950)."""
    boxes = A.tile_anchors_3d(area_extents, A.CAR_ANCHOR_SIZES, ANCHOR_STRIDE, ground_plane)
    return A.box_3d_to_anchor(boxes)


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
        _ANCHOR_CACHE = (a,) + crop_boxes(a)
    return _ANCHOR_CACHE


def decoded_boxes(offsets, tf_float32=True):
    """rpn_boxes / rpn_img_boxes ([y1,x1,y2,x2] float32, all anchors) for float32 RPN offsets: the
    reference's chain offset_to_anchor -> project_to_bev / project_to_image_space -> reorder.
    tf_float32 (the frame runner's default, FrontEndConfig.decode_tf_float32): the float32 tf.Tensor
    branches the inference graph runs; False: the float64 NumPy branches. The image projection is
    float64 from the regressed anchor either way (a float32 matmul in the TF graph)."""
    a, _, _ = anchor_set()
    if tf_float32:
        regressed = A.offset_to_anchor_tf32(a, np.asarray(offsets, dtype=np.float32))
        _, bev_norm = A.project_to_bev_tf32(regressed, BEV_EXTENTS)
        regressed = regressed.astype(np.float64)
    else:
        regressed = A.offset_to_anchor(a, np.asarray(offsets, dtype=np.float32).astype(np.float64))
        _, bev_norm = A.project_to_bev(regressed, BEV_EXTENTS)
    _, img_norm = A.project_to_image_space(regressed, A.KITTI_P2, IMAGE_SHAPE)
    return dict(rpn_boxes=A.reorder_projected_boxes(bev_norm).astype(np.float32),
                rpn_img_boxes=A.reorder_projected_boxes(img_norm).astype(np.float32))


def frame_inputs(config, frame, n_points=120000, rpn_nms_size=1024):
    """dodt_b200.synth.frame_inputs plus `rpn_boxes` / `rpn_img_boxes`: the same offsets decoded and
    projected on the host (the reference's NumPy chain), for the CPU oracle."""
    inp = _raw.frame_inputs(config, frame, n_points, rpn_nms_size)
    inp.update(decoded_boxes(inp["rpn_offsets"]))
    return inp


def clustered_rpn_outputs(config, frame, n_targets=40, per_target=300, jitter=0.01):
    """dodt_b200.synth.clustered_rpn_outputs on the car anchor grid, plus the decoded boxes."""
    a, _, _ = anchor_set()
    out = _raw.clustered_rpn_outputs(config, frame, a, n_targets, per_target, jitter)
    out.update(decoded_boxes(out["rpn_offsets"]))
    return out
