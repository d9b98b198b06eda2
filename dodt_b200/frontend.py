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

from . import anchors as A
from . import ops, synth
from ._lib import BEV_STATS_LEN


@dataclass
class FrontEndConfig:
    area_extents: list = field(default_factory=lambda: [list(e) for e in synth.AREA_EXTENTS])
    ground_plane: list = field(default_factory=lambda: list(synth.GROUND_PLANE))
    voxel_size: float = synth.VOXEL_SIZE
    height_lo: float = synth.HEIGHT_LO
    height_hi: float = synth.HEIGHT_HI
    num_slices: int = synth.NUM_SLICES
    occ_lo: float = 0.2
    occ_hi: float = 2.0
    density_threshold: int = 1
    max_points: int = 131072
    image_shape: tuple = synth.IMAGE_SHAPE          # (360, 1200)
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
    nms_max_windows: int = 4                        # launches reserved for the RPN NMS in a graph


class FrameSlot:
    """Device buffers of one in-flight frame: inputs (filled by the caller) and outputs."""

    def __init__(self, fe):
        c, dev = fe.cfg, fe.device
        f32, i32 = torch.float32, torch.int32
        H, W = fe.nz, fe.nx
        ih, iw = c.image_shape
        nA, C = fe.num_anchors, c.feat_channels
        e = lambda *s, dtype=f32: torch.empty(s, dtype=dtype, device=dev)
        # ---- inputs
        self.points = e(3, c.max_points)
        self.n_points = 0
        self.bev_feat = e(1, H, W, C)
        self.img_feat = e(1, ih, iw, C)
        self.bev_1ch = e(1, H, W, 1)
        self.img_1ch = e(1, ih, iw, 1)
        self.rpn_boxes = e(nA, 4)          # regressed BEV boxes [z1,x1,z2,x2] normalised, per anchor
        self.rpn_img_boxes = e(nA, 4)      # regressed image boxes [y1,x1,y2,x2] normalised
        self.rpn_scores = e(nA)
        self.final_scores = e(c.rpn_nms_size)
        # ---- outputs / intermediates
        self.maps = e(c.num_slices + 1, H, W)
        self.occ = e(fe.nx, fe.nz, dtype=torch.uint8)
        self.stats = e(BEV_STATS_LEN, dtype=i32)
        self.ii = e(fe.nx + 1, fe.nz + 1, dtype=i32)
        self.keep = e(nA, dtype=torch.uint8)
        self.kept_idx = e(nA, dtype=i32)
        self.n_kept = e(1, dtype=i32)
        self.k_bev_boxes = e(nA, 4)
        self.k_img_boxes = e(nA, 4)
        self.k_rpn_boxes = e(nA, 4)
        self.k_rpn_img_boxes = e(nA, 4)
        self.k_rpn_scores = e(nA)
        self.rpn_bev_crops = e(nA, c.rpn_crop[0], c.rpn_crop[1], 1)
        self.rpn_img_crops = e(nA, c.rpn_crop[0], c.rpn_crop[1], 1)
        self.top_idx = e(c.rpn_nms_size, dtype=i32)
        self.n_top = e(2, dtype=i32)
        self.prop_bev_boxes = e(c.rpn_nms_size, 4)
        self.prop_img_boxes = e(c.rpn_nms_size, 4)
        self.corr = e(1, H, W, fe.corr_channels)
        self.bev_rois = e(c.rpn_nms_size, c.avod_crop[0], c.avod_crop[1], C)
        self.img_rois = e(c.rpn_nms_size, c.avod_crop[0], c.avod_crop[1], C)
        self.corr_rois = e(c.rpn_nms_size, c.avod_crop[0], c.avod_crop[1], fe.corr_channels)
        self.final_idx = e(c.avod_nms_size, dtype=i32)
        self.n_final = e(2, dtype=i32)
        # ---- workspaces
        u8 = lambda n: torch.empty(max(int(n), 256), dtype=torch.uint8, device=dev)
        self.ws_bev = u8(ops.bev_workspace_bytes(c.max_points, c.num_slices, fe.nx, fe.nz))
        self.ws_ii = u8(ops.integral_workspace_bytes(fe.nx, fe.nz))
        self.ws_compact = u8(ops.load().dodt_compact_workspace_bytes(nA))
        self.ws_nms_rpn = u8(ops.nms_workspace_bytes(nA))
        self.ws_nms_final = u8(ops.nms_workspace_bytes(c.rpn_nms_size))

    def input_tensors(self):
        return dict(points=self.points, bev_feat=self.bev_feat, img_feat=self.img_feat,
                    bev_1ch=self.bev_1ch, img_1ch=self.img_1ch, rpn_boxes=self.rpn_boxes,
                    rpn_img_boxes=self.rpn_img_boxes, rpn_scores=self.rpn_scores,
                    final_scores=self.final_scores)

    def result_tensors(self):
        """What a step hands back to the host (detection lists; crops stay on the device)."""
        return dict(n_kept=self.n_kept, top_idx=self.top_idx, n_top=self.n_top,
                    final_idx=self.final_idx, n_final=self.n_final, stats=self.stats)


class FrontEnd:
    def __init__(self, cfg=None, anchors=None, device=None):
        self.cfg = cfg or FrontEndConfig()
        c = self.cfg
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None \
            else torch.device(device)
        self.nx, _, self.nz, self.min_x, _, self.min_z = ops.bev_grid(c.area_extents, c.voxel_size)
        anchors = synth.car_anchors(c.area_extents, c.ground_plane) if anchors is None else anchors
        self.anchors_np = np.ascontiguousarray(anchors, dtype=np.float64)
        self.num_anchors = len(self.anchors_np)
        bev_extents = [c.area_extents[0], c.area_extents[2]]
        _, bev_norm = A.project_to_bev(self.anchors_np, bev_extents)
        _, img_norm = A.project_to_image_space(self.anchors_np, A.KITTI_P2, c.image_shape)
        to_dev = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dt).to(self.device)
        self.anchors = to_dev(self.anchors_np, torch.float64)
        self.anchor_bev_boxes = to_dev(A.reorder_projected_boxes(bev_norm), torch.float32)
        self.anchor_img_boxes = to_dev(A.reorder_projected_boxes(img_norm), torch.float32)
        _, _, self.corr_channels = ops.correlation_out_shape(
            self.nz, self.nx, 1, c.corr_max_displacement, 1, c.corr_stride_2, c.corr_padding)
        self.bev_params = ops.make_bev_params(c.ground_plane, c.area_extents, c.voxel_size,
                                              c.height_lo, c.height_hi, c.num_slices, True,
                                              c.occ_lo, c.occ_hi)
        self.side_stream = torch.cuda.Stream(device=self.device)

    def new_slot(self):
        return FrameSlot(self)

    # ---------------------------------------------------------------------------------------
    def enqueue(self, slot, prev_slot):
        """Enqueue every stage of `slot`'s frame on the current stream (+ the side stream for S4);
        `prev_slot.bev_feat` is frame t of the correlation pair, `slot.bev_feat` frame t+1.
        Capturable into a CUDA graph; returns the number of library kernels launched."""
        c, s = self.cfg, slot
        before = ops.launch_count()
        main = torch.cuda.current_stream()
        # S4 on the side stream (independent of the point cloud)
        self.side_stream.wait_stream(main)
        with torch.cuda.stream(self.side_stream):
            ops.correlation(prev_slot.bev_feat, s.bev_feat, 1, c.corr_max_displacement, 1,
                            c.corr_stride_2, c.corr_padding, out=s.corr)
        # S1
        ops.bev_slices(s.points[:, :s.n_points], self.bev_params, s.maps, s.occ, s.stats, s.ws_bev)
        # S2
        ops.integral_image_2d(s.occ, s.ii, s.ws_ii)
        ops.anchor_filter_2d(self.anchors, s.ii, self.nx, self.nz, self.min_x, self.min_z,
                             c.voxel_size, c.density_threshold, keep=s.keep)
        ops.compact_mask(s.keep, s.kept_idx, s.n_kept, s.ws_compact)
        ops.gather_rows_multi([(self.anchor_bev_boxes, s.k_bev_boxes),
                               (self.anchor_img_boxes, s.k_img_boxes),
                               (s.rpn_boxes, s.k_rpn_boxes),
                               (s.rpn_img_boxes, s.k_rpn_img_boxes),
                               (s.rpn_scores, s.k_rpn_scores)], s.kept_idx, s.n_kept)
        # S3a
        ops.crop_and_resize_multi([(s.bev_1ch, s.k_bev_boxes, s.rpn_bev_crops),
                                   (s.img_1ch, s.k_img_boxes, s.rpn_img_crops)],
                                  c.rpn_crop, 0.0, n_dev=s.n_kept)
        # S5a
        ops.nms(s.k_rpn_boxes, s.k_rpn_scores, c.rpn_nms_size, c.rpn_nms_iou, keep=s.top_idx,
                n_keep=s.n_top, workspace=s.ws_nms_rpn, n_dev=s.n_kept,
                max_windows=c.nms_max_windows)
        ops.gather_rows_multi([(s.k_rpn_boxes, s.prop_bev_boxes),
                               (s.k_rpn_img_boxes, s.prop_img_boxes)], s.top_idx, s.n_top)
        # S3b (the corr crop needs S4)
        main.wait_stream(self.side_stream)
        ops.crop_and_resize_multi([(s.bev_feat, s.prop_bev_boxes, s.bev_rois),
                                   (s.img_feat, s.prop_img_boxes, s.img_rois),
                                   (s.corr, s.prop_bev_boxes, s.corr_rois)],
                                  c.avod_crop, 0.0, n_dev=s.n_top)
        # S5b
        ops.nms(s.prop_bev_boxes, s.final_scores, c.avod_nms_size, c.avod_nms_iou,
                keep=s.final_idx, n_keep=s.n_final, workspace=s.ws_nms_final, n_dev=s.n_top)
        return ops.launch_count() - before

    def capture(self, slot, prev_slot):
        """Warm up eagerly once (sets kernel attributes), then capture `enqueue` into a graph.
        Returns (graph, kernels per replay)."""
        self.enqueue(slot, prev_slot)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            launches = self.enqueue(slot, prev_slot)
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
