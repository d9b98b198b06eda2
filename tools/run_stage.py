"""Runs one stage of the front end a few times on synthetic config-B data (for ncu captures).
usage: python tools/run_stage.py {corr|corrstream|bev|filter|nms|crops|frame} [reps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dodt_b200 import ops, synth  # noqa: E402
from dodt_b200.frontend import FrontEnd  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "corr"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
fe = FrontEnd()
slots = [fe.new_slot(), fe.new_slot()]
for i, s in enumerate(slots):
    inp = synth.frame_inputs(2, i)
    for k, dst in s.input_tensors().items():
        src = torch.from_numpy(np.ascontiguousarray(inp[k]))
        if k == "points":
            dst[:, :src.shape[1]].copy_(src)
            s.n_points = src.shape[1]
        else:
            dst.copy_(src)
fe.enqueue(slots[1], slots[0])
torch.cuda.synchronize()
c, s, p = fe.cfg, slots[1], slots[0]
for _ in range(reps):
    if what == "corrstream":   # the frame runner's launch: CORR_PAIRS (8) consecutive pairs, one kernel
        if "ring" not in globals():
            n_pairs = int(os.environ.get("CORR_PAIRS", "8"))
            ring = [torch.rand_like(s.bev_feat) for _ in range(n_pairs + 1)]
            outs = [torch.empty_like(s.corr) for _ in range(n_pairs)]
        ops.correlation_stream(ring, 1, c.corr_max_displacement, 1, c.corr_stride_2, c.corr_padding, outs=outs,
                               max_ctas=int(os.environ.get("CORR_CTAS", "0")))
    elif what == "corr":
        ops.correlation(p.bev_feat, s.bev_feat, 1, c.corr_max_displacement, 1, c.corr_stride_2, c.corr_padding, out=s.corr)
    elif what == "bev":
        ops.bev_slices(s.points[:, :s.n_points], fe.bev_params, s.maps, s.occ, s.stats, s.ws_bev)
    elif what == "filter":
        fe.enqueue_s2(s)
    elif what == "nms":
        ops.nms(s.k_rpn_boxes, s.k_rpn_scores, c.rpn_nms_size, c.rpn_nms_iou, keep=s.top_idx, n_keep=s.n_top,
                workspace=s.ws_nms_rpn, n_dev=s.n_kept, max_windows=c.nms_max_windows)
    elif what == "final":
        ops.nms(s.prop_bev_boxes, s.final_scores, c.avod_nms_size, c.avod_nms_iou, keep=s.final_idx,
                n_keep=s.n_final, workspace=s.ws_nms_final, n_dev=s.n_top)
    elif what == "crops":
        ops.crop_and_resize(s.bev_feat, s.prop_bev_boxes, None, c.avod_crop, 0.0, out=s.bev_rois, n_dev=s.n_top)
        ops.crop_and_resize(s.bev_1ch, s.k_bev_boxes, None, c.rpn_crop, 0.0, out=s.rpn_bev_crops, n_dev=s.n_kept)
    else:
        fe.enqueue(s, p)
torch.cuda.synchronize()
print("ok", what, reps)
