"""The whole front end of one frame on the CPU through the oracle — TEST INFRASTRUCTURE ONLY
(used by bench.py's cpu_baseline / --impl reference legs and by the end-to-end parity test).

S1/S2 run the NumPy restatement of the reference's own NumPy code or, with use_reference=True
(where the read-only checkout or its staged copy oracle/_ref/py exists), that code ITSELF through
oracle/ref_shim.py; S3/S4/S5
run the C restatement (oracle/c_oracle.c). Same stage order and inputs as
dodt_b200.frontend.FrontEnd.enqueue.
"""
import time

import numpy as np

from . import anchor_helpers as A  # noqa: F401
from . import synth_ref as s

from . import c_oracle as CO
from . import np_oracle as O

anchors = s.anchor_set
frame_inputs = s.frame_inputs


def reference_s1_s2_available():
    """True when the reference's own S1/S2 Python can be imported (checkout or oracle/_ref/py)."""
    from . import ref_shim
    return ref_shim.available()


_REF_GEN = None


def _reference_s1_s2(pc, a):
    """S1 + S2 by the reference's OWN code: BevSlices.generate_bev (bev_slices.py:33-150),
    create_sliced_voxel_grid_2d (kitti_utils.py:212-277) and get_empty_anchor_filter_2d
    (anchor_filter.py:64-119), imported through oracle/ref_shim.py."""
    global _REF_GEN
    from . import ref_shim
    if _REF_GEN is None:
        _REF_GEN = ref_shim.reference_bev_slices(s.HEIGHT_LO, s.HEIGHT_HI, s.NUM_SLICES)
    from avod.core import anchor_filter
    t0 = time.perf_counter()
    bev = _REF_GEN.generate_bev("lidar", pc, s.GROUND_PLANE, np.asarray(s.AREA_EXTENTS), s.VOXEL_SIZE)
    t1 = time.perf_counter()
    vg = ref_shim.reference_sliced_voxel_grid_2d(pc, s.GROUND_PLANE, np.asarray(s.AREA_EXTENTS), s.VOXEL_SIZE)
    keep = anchor_filter.get_empty_anchor_filter_2d(a, vg, 1)
    t2 = time.perf_counter()
    occ = (np.squeeze(vg.leaf_layout_2d) + 1).astype(np.uint8)
    return bev, occ, keep, t1 - t0, t2 - t1


def run_frame(inp, prev_bev_feat, rpn_nms=(1024, 0.8), avod_nms=(100, 0.01), timings=None,
              k_boxes=None, prop_img_boxes=None, anchor_img_boxes=None, use_reference=False):
    """All stages of one frame; returns the outputs FrontEnd produces (for parity).

    The RPN decode (offset_to_anchor + projections) is the reference's NumPy chain, precomputed in
    inp["rpn_boxes"] / inp["rpn_img_boxes"]; k_boxes (BEV boxes of the KEPT anchors),
    prop_img_boxes (image boxes of the NMS survivors, or a callable top -> boxes) and
    anchor_img_boxes replace them when a test wants the later stages judged on exactly the boxes
    the device produced (those agree with NumPy to float64 rounding noise, see tests)."""
    a, a_bev, a_img = anchors()
    t = {}

    def tic():
        return time.perf_counter()

    pc = inp["points"].astype(np.float64)
    if use_reference:
        bev, occ, keep, t["S1"], t["S2"] = _reference_s1_s2(pc, a)
    else:
        t0 = tic()
        bev = O.bev_slices(pc, s.GROUND_PLANE, s.AREA_EXTENTS, s.VOXEL_SIZE, s.HEIGHT_LO, s.HEIGHT_HI,
                           s.NUM_SLICES)
        t["S1"] = tic() - t0
        t0 = tic()
        occ, vox = O.occupancy_grid(pc, s.GROUND_PLANE, s.AREA_EXTENTS, s.VOXEL_SIZE)
        keep = O.empty_anchor_filter_2d(a, occ, s.VOXEL_SIZE, vox["min_coord"][[0, 2]], 1)
        t["S2"] = tic() - t0
    kept = np.flatnonzero(keep)
    t0 = tic()
    zeros = np.zeros(len(kept), dtype=np.int32)
    rpn_bev_crops = CO.crop_and_resize(inp["bev_1ch"], a_bev[kept], zeros, (3, 3))
    if anchor_img_boxes is not None:
        a_img = anchor_img_boxes
    rpn_img_crops = CO.crop_and_resize(inp["img_1ch"], a_img[kept], zeros, (3, 3))
    t["S3_rpn"] = tic() - t0
    t0 = tic()
    k_scores = inp["rpn_scores"][kept]
    if k_boxes is None:
        k_boxes = inp["rpn_boxes"][kept]

    top = CO.non_max_suppression(k_boxes, k_scores, rpn_nms[0], rpn_nms[1])
    t["S5_rpn"] = tic() - t0
    t0 = tic()
    corr = CO.correlation(prev_bev_feat, inp["bev_feat"], 1, 5, 1, 2, 5)
    t["S4"] = tic() - t0
    t0 = tic()
    prop_bev = k_boxes[top]
    if prop_img_boxes is None:
        prop_img = inp["rpn_img_boxes"][kept][top]
    else:
        prop_img = prop_img_boxes(top) if callable(prop_img_boxes) else prop_img_boxes
    z = np.zeros(len(top), dtype=np.int32)
    bev_rois = CO.crop_and_resize(inp["bev_feat"], prop_bev, z, (7, 7))
    img_rois = CO.crop_and_resize(inp["img_feat"], prop_img, z, (7, 7))
    corr_rois = CO.crop_and_resize(corr, prop_bev, z, (7, 7))
    t["S3_avod"] = tic() - t0
    t0 = tic()
    final = CO.non_max_suppression(prop_bev, inp["final_scores"][:len(top)], avod_nms[0], avod_nms[1])
    t["S5_avod"] = tic() - t0
    if timings is not None:
        timings.update(t)
    return dict(bev=bev, occ=occ, keep=keep, kept=kept, rpn_bev_crops=rpn_bev_crops,
                rpn_img_crops=rpn_img_crops, top=top, corr=corr, bev_rois=bev_rois,
                img_rois=img_rois, corr_rois=corr_rois, final=final)


def _worker(args):
    """One frame in one process. BLAS is held to one thread per process: the pool already uses
    every core, and np.dot inside get_point_filter would otherwise oversubscribe them."""
    config, frame = args[0], args[1]
    use_reference = bool(args[2]) if len(args) > 2 else False
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    inp = frame_inputs(config, frame)
    prev, _ = s.feature_pair(config, frame + 1)
    t0 = time.perf_counter()
    timings = {}
    out = run_frame(inp, prev, timings=timings, use_reference=use_reference)
    dt = time.perf_counter() - t0
    return dt, timings, int(len(out["top"])), int(len(out["final"]))
