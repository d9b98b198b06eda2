"""Tracking-association IoU and the greedy IoU linker on the device: the
consumer of the front end's detection lists.

Drop-ins for
  wavedata/wavedata/tools/obj_detection/evaluation.py:44-92   three_d_iou(box, boxes)
  avod/experiments/video_detection.py:58-67                   iou_3d(box3d_1, box3d_2)
  avod/experiments/video_detection.py:235-277                 track_iou(...)
The reference scores every (track, detection) pair with a Python call that rasterises two
rectangles with PIL; here the pairs of a frame are ONE kernel launch (dodt_three_d_iou_matrix:
exact clipping of the rotated bases, float64), which agrees with the rasterised value to the
discretisation error of the reference (<= 0.01 IoU on car-sized boxes; see tests).
"""
import numpy as np
import torch

from . import ops
from ._lib import check, load


def three_d_iou_matrix(boxes_a, boxes_b, out=None):
    """IoU of every box of boxes_a [na, 7] with every box of boxes_b [nb, 7] (CUDA float64 rows
    [ry, l, h, w, tx, ty, tz]) -> [na, nb] float64 on the device, no synchronisation."""
    ops._need_cuda(boxes_a, boxes_b, out)
    if boxes_a.dim() != 2 or boxes_b.dim() != 2 or boxes_a.shape[1] != 7 or boxes_b.shape[1] != 7:
        raise ValueError("boxes must be [n, 7] rows [ry, l, h, w, tx, ty, tz]")
    if boxes_a.dtype != torch.float64 or boxes_b.dtype != torch.float64:
        raise TypeError("three_d_iou_matrix expects float64 boxes")
    a, b = boxes_a.contiguous(), boxes_b.contiguous()
    na, nb = a.shape[0], b.shape[0]
    if out is None:
        out = torch.empty((na, nb), dtype=torch.float64, device=a.device)
    check(load().dodt_three_d_iou_matrix(ops._ptr(a), na, ops._ptr(b), nb, ops._ptr(out), ops._stream()),
          "dodt_three_d_iou_matrix")
    return out


def three_d_iou(box, boxes):
    """Signature and return conventions of wavedata evaluation.three_d_iou: box [7], boxes [7] or
    [n, 7] as NumPy arrays (or CUDA tensors); returns a float for one box, else an array [n]."""
    as_numpy = not torch.is_tensor(box)
    dev = torch.device("cuda", torch.cuda.current_device()) if as_numpy else box.device
    tb = torch.as_tensor(np.asarray(box, dtype=np.float64) if as_numpy else box, dtype=torch.float64, device=dev)
    to = torch.as_tensor(np.asarray(boxes, dtype=np.float64) if not torch.is_tensor(boxes) else boxes,
                         dtype=torch.float64, device=dev)
    if to.dim() == 1:
        to = to[None]
    iou = three_d_iou_matrix(tb.reshape(1, 7), to)[0]
    if as_numpy:
        iou = iou.cpu().numpy()
    return iou[0] if iou.shape[0] == 1 else iou


def iou_3d(box3d_1, box3d_2, scale_first=4.0):
    """avod/experiments/video_detection.py:58-67: boxes [l, w, h, tx, ty, tz, ry] reordered to
    [ry, l, h, w, tx, ty, tz]; that file inflates the FIRST box's dimensions four-fold before the
    test (scale_first = 4; video_detection_iou.py:164-173 does not: pass 1)."""
    b1 = np.asarray(box3d_1, dtype=np.float64)[[-1, 0, 2, 1, 3, 4, 5]].copy()
    b1[1:4] = scale_first * b1[1:4]
    b2 = np.asarray(box3d_2, dtype=np.float64)
    b2 = b2[[-1, 0, 2, 1, 3, 4, 5]] if b2.ndim == 1 else b2[:, [-1, 0, 2, 1, 3, 4, 5]]
    return three_d_iou(b1, b2)


def track_iou(detections, sigma_l, sigma_h, sigma_iou, t_min, box_key="boxes3d", to_iou_box=None):
    """The greedy IoU linker of avod/experiments/video_detection.py:235-277 with the pair scores of
    a frame computed in one launch: IoU of the last detection of every active track with every
    detection of the frame (three_d_iou), then the reference's loop — for each track in order the
    best remaining detection is taken if its IoU exceeds sigma_iou, unmatched tracks are finished
    when max_score >= sigma_h and length >= t_min, unmatched detections start tracks.
    detections: per frame a list of dicts with 'scores', 'frame_id' and det[box_key] = [l, w, h, tx,
    ty, tz, ry]; to_iou_box maps that row to [ry, l, h, w, tx, ty, tz] (default: the reordering of
    video_detection_iou.iou_3d without inflation)."""
    if to_iou_box is None:
        def to_iou_box(b):
            return np.asarray(b, dtype=np.float64)[[-1, 0, 2, 1, 3, 4, 5]]
    dev = torch.device("cuda", torch.cuda.current_device())
    tracks_active, tracks_finished = [], []
    for detections_frame in detections:
        if detections_frame == []:
            continue
        dets = [det for det in detections_frame if det['scores'] >= sigma_l]
        alive = list(range(len(dets)))            # indices of dets not yet taken
        iou = None
        if tracks_active and dets:
            ta = torch.as_tensor(np.stack([to_iou_box(t['trajectory'][-1][box_key]) for t in tracks_active]), device=dev)
            td = torch.as_tensor(np.stack([to_iou_box(d[box_key]) for d in dets]), device=dev)
            iou = three_d_iou_matrix(ta, td).cpu().numpy()
        updated_tracks = []
        for ti, track in enumerate(tracks_active):
            matched = False
            if alive:
                row = iou[ti, alive]
                best = int(np.argmax(row))
                if row[best] > sigma_iou:
                    det = dets[alive[best]]
                    track['trajectory'].append(det)
                    track['max_score'] = max(track['max_score'], det['scores'])
                    updated_tracks.append(track)
                    del alive[best]
                    matched = True
            if not matched and track['max_score'] >= sigma_h and len(track['trajectory']) >= t_min:
                tracks_finished.append(track)
        new_tracks = [{'trajectory': [dets[k]], 'max_score': dets[k]['scores'], 'start_frame': dets[k]['frame_id']}
                      for k in alive]
        tracks_active = updated_tracks + new_tracks
    tracks_finished += [t for t in tracks_active if t['max_score'] >= sigma_h and len(t['trajectory']) >= t_min]
    return tracks_finished
