"""Frame-stream runner: the whole DODT proposal front end of one frame (pair) enqueued on the
device without a host round trip, captured once into a CUDA graph and replayed per frame.

Stage order and data flow follow the reference's inference loop:

  points ─S1 bev_slices──► BEV maps [6,700,800] ─► (network input, out of scope)
          └─ occupancy ─S2 integral image + anchor filter ─► keep mask ─► compaction (kept_idx, n)
  kept anchors ─S3a crop_and_resize 3x3 on the 1-ch BEV / image bottlenecks  (dt_rpn_model.py:418-428)
  RPN head outputs of the kept anchors ─S5a NMS 0.8 / 1024 ─► proposals     (dt_rpn_model.py:587-597)
  BEV features t, t+1 ─S4 correlation ─► corr map [1,700,800,25]            (dt_rpn_model.py:324-331)
  proposals ─S3b crop_and_resize 7x7 on BEV / image / corr maps            (dt_avod_model.py:253-273)
  second-stage outputs ─S5b NMS 0.01 / 100 ─► detections                   (dt_avod_model.py:609-613)

The networks between the stages are out of scope; their outputs (RPN scores and regressed boxes,
final scores) are per-frame INPUTS of a slot, exactly as the BEV/image feature maps are.
S4 runs on a second stream so that the small latency-bound kernels of S1/S2/S5 overlap it.
"""
from dataclasses import dataclass, field

import numpy as np
import torch

from . import ops, synth
from .constants import CAR_ANCHOR_SIZES, KITTI_P2
from ._lib import BEV_STATS_LEN


@dataclass
class FrontEndConfig:
    area_extents: list = field(default_factory=lambda: [list(e) for e in synth.AREA_EXTENTS])
    ground_plane: list = field(default_factory=lambda: list(synth.GROUND_PLANE))
    voxel_size: float = synth.VOXEL_SIZE
    height_lo: float = synth.HEIGHT_LO
    height_hi: float = synth.HEIGHT_HI
    num_slices: int = synth.NUM_SLICES
    anchor_3d_sizes: list = field(default_factory=lambda: [list(v) for v in CAR_ANCHOR_SIZES])
    anchor_stride: list = field(default_factory=lambda: list(synth.ANCHOR_STRIDE))
    occ_lo: float = 0.2
    occ_hi: float = 2.0
    density_threshold: int = 1
    max_points: int = 131072
    image_shape: tuple = synth.IMAGE_SHAPE          # (360, 1200)
    stereo_calib_p2: list = field(default_factory=lambda: [float(v) for v in KITTI_P2.reshape(-1)])
    feat_channels: int = 32
    rpn_crop: tuple = (3, 3)                        # rpn_proposal_roi_crop_size
    rpn_nms_size: int = 1024                        # rpn_train_nms_size
    rpn_nms_iou: float = 0.8
    avod_crop: tuple = (7, 7)                       # avod_proposal_roi_crop_size
    avod_nms_size: int = 100
    avod_nms_iou: float = 0.01
    corr_max_displacement: int = 5
    corr_padding: int = 5
    corr_stride_2: int = 2
    nms_max_windows: int = 2                        # windows (1536 candidates each, a multiple of 2) reserved
                                                    # for the RPN NMS in a graph; a frame that needs more
                                                    # reports n_top[1] == 0 and is finished by
                                                    # FrontEnd.complete_frame
    decode_tf_float32: bool = True                  # RPN decode + BEV projection as the reference's inference
                                                    # graph computes them (float32 tf.Tensor branches,
                                                    # dt_rpn_model.py:568-591); False: the float64 NumPy
                                                    # branches of the same functions
    corr_stream_priority: int = 0                   # CUDA stream priorities of the graph's S4 branch and of
    chain_stream_priority: int = -1                 # its per-frame chains (0 = lowest, -1, -2 .. higher).
                                                    # The block scheduler serves the chains' short kernels
                                                    # before the correlation's long-lived CTAs: 78.4 instead
                                                    # of 82.0 us per frame (0 / -1: nothing)
    corr_pairs_per_launch: int = 8                  # consecutive pairs per frame-stream S4 launch
    corr_max_ctas: int = 0                          # CTA cap of the correlation launch (0 = none: two
                                                    # persistent CTAs per SM). 148 (one per SM, the other
                                                    # half of each SM left to neighbouring frames) is ~2 %
                                                    # faster in the frame pipeline when the hardware
                                                    # spreads the 148 CTAs one per SM, but about one run
                                                    # in ten it co-locates some of them and the launch
                                                    # takes 1.5x (330 vs 224 us for four pairs)


def _layout(specs, align=256):
    """[(name, shape, dtype)] -> ({name: (offset, shape, dtype)}, total bytes), 256-byte aligned."""
    out, off = {}, 0
    for name, shape, dtype in specs:
        nbytes = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        out[name] = (off, tuple(shape), dtype)
        off = (off + nbytes + align - 1) // align * align
    return out, off


def _views(buf, layout):
    views = {}
    for name, (off, shape, dtype) in layout.items():
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        views[name] = buf[off:off + n].view(dtype).view(shape)
    return views


class FrameSlot:
    """Device buffers of one in-flight frame: inputs (filled by the caller) and outputs.

    The per-frame SENSOR-side inputs (point cloud, RPN / second-stage head outputs) live in one
    contiguous buffer and the results handed back to the host in another, so that a frame costs
    one host->device and one device->host copy besides the feature maps (`HostFrame` mirrors the
    layouts in pinned memory)."""

    def __init__(self, fe):
        c, dev = fe.cfg, fe.device
        f32, i32 = torch.float32, torch.int32
        H, W = fe.nz, fe.nx
        ih, iw = c.image_shape
        nA, C = fe.num_anchors, c.feat_channels
        e = lambda *s, dtype=f32: torch.empty(s, dtype=dtype, device=dev)
        # ---- inputs
        self.sensor_buf = torch.empty(fe.sensor_bytes, dtype=torch.uint8, device=dev)
        v = _views(self.sensor_buf, fe.sensor_layout)
        self.points = v["points"]              # (3, max_points)
        self.rpn_offsets = v["rpn_offsets"]    # RPN regression output, anchor-form offsets per anchor
        self.rpn_scores = v["rpn_scores"]
        self.final_scores = v["final_scores"]
        self.frame_id = v["frame_id"]          # (sequence, frame) of the frame in this slot
        self.n_points = 0
        self.bev_feat = e(1, H, W, C)
        self.img_feat = e(1, ih, iw, C)
        self.bev_1ch = e(1, H, W, 1)
        self.img_1ch = e(1, ih, iw, 1)
        # ---- results (one buffer)
        self.result_buf = torch.empty(fe.result_bytes, dtype=torch.uint8, device=dev)
        r = _views(self.result_buf, fe.result_layout)
        self.n_kept, self.n_top, self.n_final = r["n_kept"], r["n_top"], r["n_final"]
        self.stats, self.top_idx, self.final_idx = r["stats"], r["top_idx"], r["final_idx"]
        self.det_row = r["det_row"]            # row of the shard's DetectionBlock this frame wrote
        # ---- intermediates
        self.maps = e(c.num_slices + 1, H, W)
        self.occ = e(fe.nx, fe.nz, dtype=torch.uint8)
        self.ii = e(fe.nx + 1, fe.nz + 1, dtype=i32)
        self.keep = e(nA, dtype=torch.uint8)
        self.kept_idx = e(nA, dtype=i32)
        self.k_bev_boxes = e(nA, 4)
        self.k_img_boxes = e(nA, 4)
        self.k_rpn_boxes = e(nA, 4)
        self.k_rpn_scores = e(nA)
        self.rpn_bev_crops = e(nA, c.rpn_crop[0], c.rpn_crop[1], 1)
        self.rpn_img_crops = e(nA, c.rpn_crop[0], c.rpn_crop[1], 1)
        self.prop_bev_boxes = e(c.rpn_nms_size, 4)
        self.prop_img_boxes = e(c.rpn_nms_size, 4)
        self.corr = e(1, H, W, fe.corr_channels)
        self.bev_rois = e(c.rpn_nms_size, c.avod_crop[0], c.avod_crop[1], C)
        self.img_rois = e(c.rpn_nms_size, c.avod_crop[0], c.avod_crop[1], C)
        self.corr_rois = e(c.rpn_nms_size, c.avod_crop[0], c.avod_crop[1], fe.corr_channels)
        # ---- workspaces
        u8 = lambda n: torch.empty(max(int(n), 256), dtype=torch.uint8, device=dev)
        self.ws_bev = u8(ops.bev_workspace_bytes(c.max_points, c.num_slices, fe.nx, fe.nz))
        # zero before first use; the banded integral image and the fused filter leave them zero
        self.ws_ii = torch.zeros(max(ops.integral_banded_workspace_bytes(fe.nx, fe.nz), 256),
                                 dtype=torch.uint8, device=dev)
        self.ws_fused = ops.anchor_filter_fused_workspace(nA, dev)
        self.ws_nms_rpn = u8(ops.nms_workspace_bytes(nA))
        self.ws_nms_final = u8(ops.nms_workspace_bytes(c.rpn_nms_size))

    def input_tensors(self):
        return dict(points=self.points, bev_feat=self.bev_feat, img_feat=self.img_feat,
                    bev_1ch=self.bev_1ch, img_1ch=self.img_1ch, rpn_offsets=self.rpn_offsets,
                    rpn_scores=self.rpn_scores,
                    final_scores=self.final_scores)

    def result_tensors(self):
        """What a step hands back to the host (detection lists; crops stay on the device)."""
        return dict(n_kept=self.n_kept, top_idx=self.top_idx, n_top=self.n_top,
                    final_idx=self.final_idx, n_final=self.n_final, stats=self.stats)


FEATURE_KEYS = ("bev_feat", "img_feat", "bev_1ch", "img_1ch")   # network outputs


class HostFrame:
    """Pinned host mirror of one FrameSlot's inputs and results (same packed layouts)."""

    def __init__(self, fe):
        self.fe = fe
        self.sensor_buf = torch.empty(fe.sensor_bytes, dtype=torch.uint8).pin_memory()
        self.sensor = _views(self.sensor_buf, fe.sensor_layout)
        c = fe.cfg
        ih, iw = c.image_shape
        shapes = dict(bev_feat=(1, fe.nz, fe.nx, c.feat_channels), img_feat=(1, ih, iw, c.feat_channels),
                      bev_1ch=(1, fe.nz, fe.nx, 1), img_1ch=(1, ih, iw, 1))
        self.features = {k: torch.empty(s, dtype=torch.float32).pin_memory() for k, s in shapes.items()}
        self.result_buf = torch.empty(fe.result_bytes, dtype=torch.uint8).pin_memory()
        self.results = _views(self.result_buf, fe.result_layout)
        self.n_points = 0

    def fill(self, inputs, sequence=0, frame=0):
        """inputs: dict of NumPy arrays as dodt_b200.synth.frame_inputs returns them."""
        self.sensor["frame_id"][0] = int(sequence)
        self.sensor["frame_id"][1] = int(frame)
        for k, v in inputs.items():
            t = torch.from_numpy(np.ascontiguousarray(v))
            if k == "points":
                self.n_points = t.shape[1]
                self.sensor["points"][:, :self.n_points].copy_(t)
            elif k in self.features:
                self.features[k].copy_(t)
            elif k in self.sensor:
                self.sensor[k].copy_(t)       # other keys (host-decoded boxes for the oracle) are not inputs
        return self

    @property
    def h2d_bytes(self):
        return self.sensor_buf.numel() + sum(v.numel() * 4 for v in self.features.values())

    def upload(self, slot, features=True):
        """Enqueue the host->device copies of this frame on the current stream."""
        slot.sensor_buf.copy_(self.sensor_buf, non_blocking=True)
        slot.n_points = self.n_points
        if features:
            for k, v in self.features.items():
                getattr(slot, k).copy_(v, non_blocking=True)

    def download(self, slot):
        """Enqueue the device->host copy of the frame's results on the current stream."""
        self.result_buf.copy_(slot.result_buf, non_blocking=True)


class FrontEnd:
    def __init__(self, cfg=None, anchors=None, device=None):
        self.cfg = cfg or FrontEndConfig()
        c = self.cfg
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None \
            else torch.device(device)
        self.nx, _, self.nz, self.min_x, _, self.min_z = ops.bev_grid(c.area_extents, c.voxel_size)
        with torch.cuda.device(self.device):
            # S2 evaluates the config's anchor grid from the anchor index (no table read); a caller's own
            # anchors are read from their table
            self.anchor_grid = None if anchors is not None else \
                (c.area_extents, c.anchor_3d_sizes, c.anchor_stride, c.ground_plane)
            if anchors is None:
                # the anchor grid of the config, generated on the device (bit-identical to
                # box_3d_to_anchor(tile_anchors_3d(...)))
                self.anchors = ops.grid_anchors(c.area_extents, c.anchor_3d_sizes, c.anchor_stride,
                                                c.ground_plane, self.device)
            else:
                self.anchors = torch.from_numpy(np.ascontiguousarray(anchors, dtype=np.float64)).to(self.device)
            self.num_anchors = int(self.anchors.shape[0])
            self.bev_extents4 = [c.area_extents[0][0], c.area_extents[0][1],
                                 c.area_extents[2][0], c.area_extents[2][1]]
            # the anchors' own projections (crop boxes of the RPN stage, dt_rpn_model.py:975-985)
            self.anchor_bev_boxes = ops.project_to_bev(self.anchors, self.bev_extents4, tf_order=True)
            self.anchor_img_boxes = ops.project_to_image_space(self.anchors, c.stereo_calib_p2,
                                                               c.image_shape, tf_order=True)
        _, _, self.corr_channels = ops.correlation_out_shape(
            self.nz, self.nx, 1, c.corr_max_displacement, 1, c.corr_stride_2, c.corr_padding)
        self.bev_params = ops.make_bev_params(c.ground_plane, c.area_extents, c.voxel_size,
                                              c.height_lo, c.height_hi, c.num_slices, True,
                                              c.occ_lo, c.occ_hi)
        # stream priorities of the graph's branches (CUDA: lower number = higher priority; kernel
        # nodes captured from a stream keep its priority)
        self.side_stream = torch.cuda.Stream(device=self.device, priority=c.corr_stream_priority)
        self.branch_streams = []   # one per frame of an enqueue_group
        f32, i32 = torch.float32, torch.int32
        nA = self.num_anchors
        self.sensor_layout, self.sensor_bytes = _layout([
            ("points", (3, c.max_points), f32), ("rpn_offsets", (nA, 6), f32),
            ("rpn_scores", (nA,), f32),
            ("final_scores", (c.rpn_nms_size,), f32), ("frame_id", (2,), i32)])
        self.result_layout, self.result_bytes = _layout([
            ("n_kept", (1,), i32), ("n_top", (2,), i32), ("n_final", (2,), i32),
            ("det_row", (1,), i32),
            ("stats", (BEV_STATS_LEN,), i32), ("top_idx", (c.rpn_nms_size,), i32),
            ("final_idx", (c.avod_nms_size,), i32)])

    def new_slot(self):
        return FrameSlot(self)

    # ---------------------------------------------------------------------------------------
    def enqueue(self, slot, prev_slot, block=None, skip=()):
        """Enqueue every stage of `slot`'s frame on the current stream (+ the side stream for S4);
        `prev_slot.bev_feat` is frame t of the correlation pair, `slot.bev_feat` frame t+1.
        block: optional dodt_b200.shard.DetectionBlock on this device — the frame's final
        detections are appended to it (the list that a sharded run gathers once per shard).
        skip: stage names left out (profiling only: "S1", "S2", "S3a", "S5a", "S4", "S3b", "S5b").
        Capturable into a CUDA graph; returns the number of library kernels launched."""
        return self.enqueue_group([slot], prev_slot, block, skip)

    def enqueue_group(self, slots, prev_slot, block=None, skip=()):
        """Enqueue k CONSECUTIVE frames of a stream (slots[0] follows prev_slot, slots[j] follows
        slots[j-1]). The k correlations go out as ONE frame-stream launch on the side stream
        (dodt_correlation_stream: the feature map that pair j and pair j+1 share is read from HBM
        once); the per-frame chains S1..S5a run on a branch stream each and join the correlation
        before their S3b. Same results as k calls of `enqueue`."""
        c = self.cfg
        k = len(slots)
        before = ops.launch_count()
        main = torch.cuda.current_stream()
        n_branch = k if c.chain_stream_priority != 0 else k - 1   # a priority needs a stream of its own
        while len(self.branch_streams) < n_branch:
            self.branch_streams.append(torch.cuda.Stream(device=self.device, priority=c.chain_stream_priority))
        lanes = self.branch_streams[:k] if c.chain_stream_priority != 0 else [main] + self.branch_streams[:k - 1]
        # fork: S4 on the side stream (independent of the point clouds), frame j on lane j
        self.side_stream.wait_stream(main)
        for st in lanes:
            if st is not main:
                st.wait_stream(main)
        # S4: launches of up to corr_pairs_per_launch consecutive pairs, back to back on the side
        # stream; frame j waits for the launch that holds its pair only
        P = max(1, min(int(c.corr_pairs_per_launch), 8))
        corr_done = [None] * k
        if "S4" not in skip:
            with torch.cuda.stream(self.side_stream):
                for j0 in range(0, k, P):
                    grp = slots[j0:j0 + P]
                    prev = prev_slot if j0 == 0 else slots[j0 - 1]
                    if len(grp) == 1:
                        ops.correlation(prev.bev_feat, grp[0].bev_feat, 1, c.corr_max_displacement, 1,
                                        c.corr_stride_2, c.corr_padding, out=grp[0].corr,
                                        max_ctas=c.corr_max_ctas)
                    else:
                        ops.correlation_stream([prev.bev_feat] + [s.bev_feat for s in grp], 1,
                                               c.corr_max_displacement, 1, c.corr_stride_2, c.corr_padding,
                                               outs=[s.corr for s in grp], max_ctas=c.corr_max_ctas)
                    ev = torch.cuda.Event()
                    ev.record(self.side_stream)
                    for j in range(j0, min(j0 + P, k)):
                        corr_done[j] = ev
        for s, st in zip(slots, lanes):
            with torch.cuda.stream(st):
                self._enqueue_pre(s, skip)
        for j, (s, st) in enumerate(zip(slots, lanes)):
            with torch.cuda.stream(st):
                if corr_done[j] is not None:
                    st.wait_event(corr_done[j])           # S3b: the corr crop needs this frame's S4
                self._enqueue_post(s, block, skip)
        main.wait_stream(self.side_stream)
        for st in lanes:
            if st is not main:
                main.wait_stream(st)
        return ops.launch_count() - before

    def _enqueue_proposals(self, s):
        """What follows the RPN NMS: proposal boxes on the BEV map and (decoded again, eight fp64
        corner projections each) on the image, for the survivors only."""
        c = self.cfg
        # one launch: the BEV boxes come out of the same arithmetic as k_rpn_boxes (same bits as a
        # gather of the survivors' rows)
        ops.rpn_decode(self.anchors, s.rpn_offsets, s.kept_idx, s.n_top, self.bev_extents4,
                       c.stereo_calib_p2, c.image_shape, s.prop_bev_boxes, s.prop_img_boxes, idx2=s.top_idx,
                       tf_float32=c.decode_tf_float32)

    # ---------------------------------------------------------------------------------------
    def rpn_nms_complete(self, frame):
        """True if the RPN NMS of the frame finished inside the windows the graph reserves
        (FrontEndConfig.nms_max_windows). frame: a FrameSlot (reads 8 bytes from the device,
        synchronising the current stream) or a HostFrame whose results were downloaded."""
        if isinstance(frame, HostFrame):
            return int(frame.results["n_top"][1]) == 1
        return int(frame.n_top.cpu()[1]) == 1

    def complete_frame(self, slot, block=None):
        """Finish a frame whose RPN NMS did not complete inside the reserved windows: clustered
        detections at IoU 0.8 can leave fewer than rpn_nms_size survivors among the top
        nms_max_windows * 1536 candidates, and tf.image.non_max_suppression always scans on
        (dt_rpn_model.py:587-591). Resumes the selection where the graph left it (same workspace,
        first_window), then redoes everything downstream of it — proposal boxes, the 7x7 crops, the
        final NMS — and REPLACES the frame's row of the detection block. Eager, on the current
        stream, after the frame's graph has finished; the correlation map of the slot must still be
        the frame's. Returns True if work was done."""
        c = self.cfg
        if self.rpn_nms_complete(slot):
            return False
        ops.nms(slot.k_rpn_boxes, slot.k_rpn_scores, c.rpn_nms_size, c.rpn_nms_iou, keep=slot.top_idx,
                n_keep=slot.n_top, workspace=slot.ws_nms_rpn, n_dev=slot.n_kept,
                first_window=c.nms_max_windows, max_windows=0)
        self._enqueue_proposals(slot)
        self._enqueue_post(slot, block, (), rewrite=block is not None)
        return True

    def enqueue_s2(self, s):
        """S2 of one slot (needs its occupancy grid from S1)."""
        c = self.cfg
        # two launches: the band-local integral image (+ band offsets by its last CTA), then
        # filter + ordered compaction + crop boxes / scores of the kept anchors + the RPN decode
        # of the kept anchors (dt_rpn_model.py:573-591: regressed anchors projected into the BEV
        # map) in one kernel
        bandoff, band_rows = ops.integral_image_2d_banded(s.occ, s.ii, s.ws_ii)
        ops.anchor_filter_fused(None if self.anchor_grid is not None else self.anchors, s.ii, self.nx, self.nz,
                                self.min_x, self.min_z,
                                c.voxel_size, c.density_threshold, s.keep, s.kept_idx, s.n_kept,
                                s.ws_fused, bandoff=bandoff, band_rows=band_rows,
                                anchor_bev_boxes=self.anchor_bev_boxes, k_bev_boxes=s.k_bev_boxes,
                                anchor_img_boxes=self.anchor_img_boxes, k_img_boxes=s.k_img_boxes,
                                rpn_scores=s.rpn_scores, k_scores=s.k_rpn_scores,
                                rpn_offsets=s.rpn_offsets, bev_extents=self.bev_extents4,
                                k_rpn_boxes=s.k_rpn_boxes, tf_float32=c.decode_tf_float32, grid=self.anchor_grid)

    def _enqueue_pre(self, s, skip):
        c = self.cfg
        if "S1" not in skip:
            ops.bev_slices(s.points[:, :s.n_points], self.bev_params, s.maps, s.occ, s.stats, s.ws_bev)
        if "S2" not in skip:
            self.enqueue_s2(s)
        if "S3a" not in skip:
            ops.crop_and_resize_multi([(s.bev_1ch, s.k_bev_boxes, s.rpn_bev_crops),
                                       (s.img_1ch, s.k_img_boxes, s.rpn_img_crops)],
                                      c.rpn_crop, 0.0, n_dev=s.n_kept)
        if "S5a" not in skip:
            ops.nms(s.k_rpn_boxes, s.k_rpn_scores, c.rpn_nms_size, c.rpn_nms_iou, keep=s.top_idx,
                    n_keep=s.n_top, workspace=s.ws_nms_rpn, n_dev=s.n_kept,
                    max_windows=c.nms_max_windows)
            self._enqueue_proposals(s)

    def _enqueue_post(self, s, block, skip, rewrite=False):
        c = self.cfg
        if "S3b" not in skip:
            ops.crop_and_resize_multi([(s.bev_feat, s.prop_bev_boxes, s.bev_rois),
                                       (s.img_feat, s.prop_img_boxes, s.img_rois),
                                       (s.corr, s.prop_bev_boxes, s.corr_rois)],
                                      c.avod_crop, 0.0, n_dev=s.n_top)
        if "S5b" not in skip:
            ops.nms(s.prop_bev_boxes, s.final_scores, c.avod_nms_size, c.avod_nms_iou,
                    keep=s.final_idx, n_keep=s.n_final, workspace=s.ws_nms_final, n_dev=s.n_top)
        if block is not None:
            ops.emit_detections(s.prop_bev_boxes, s.final_scores, s.final_idx, s.n_final, block,
                                frame_id=s.frame_id, row_io=s.det_row, rewrite=rewrite)

    def capture(self, slot, prev_slot, block=None, skip=()):
        """Warm up eagerly once (sets kernel attributes), then capture `enqueue` into a graph.
        Returns (graph, kernels per replay)."""
        return self.capture_group([slot], prev_slot, block, skip)

    def capture_group(self, slots, prev_slot, block=None, skip=()):
        """`enqueue_group` as one CUDA graph. Returns (graph, kernels per replay)."""
        self.enqueue_group(slots, prev_slot, block)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            launches = self.enqueue_group(slots, prev_slot, block, skip)
        return graph, launches

    # ---------------------------------------------------------------------------------------
    def algorithmic_bytes(self, n_points, n_kept, n_top):
        """Bytes per frame of SURVEY §8(d) / BASELINE.md (the roofline numerator), per stage."""
        c = self.cfg
        H, W, S, C = self.nz, self.nx, c.num_slices, c.feat_channels
        ih, iw = c.image_shape
        nA = self.num_anchors
        s1 = 16 * n_points + 4 * (S + 1) * H * W
        s2 = H * W + 4 * (self.nx + 1) * (self.nz + 1) + 65 * nA
        rc = c.rpn_crop[0] * c.rpn_crop[1]
        ac = c.avod_crop[0] * c.avod_crop[1]
        s3a = 2 * (n_kept * rc * 1 * 20 + 16 * n_kept)
        s3b = n_top * ac * (2 * C + self.corr_channels) * 20 + 3 * 16 * n_top
        s4 = 2 * H * W * C * 4 + H * W * self.corr_channels * 4
        s5 = (20 * n_kept + 4 * c.rpn_nms_size) + (20 * n_top + 4 * c.avod_nms_size)
        return dict(S1=s1, S2=s2, S3_rpn=s3a, S3_avod=s3b, S4=s4, S5=s5,
                    total=s1 + s2 + s3a + s3b + s4 + s5)
