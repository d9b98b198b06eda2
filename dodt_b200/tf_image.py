"""Drop-ins for the two TensorFlow 1.3 image ops on the DODT proposal path, with TF's argument
names and order, running on sm_100a kernels:

  crop_and_resize(image, boxes, box_ind, crop_size, method='bilinear', extrapolation_value=0)
      call sites: avod/core/models/dt_rpn_model.py:418-428, dt_avod_model.py:253-273
  non_max_suppression(boxes, scores, max_output_size, iou_threshold=0.5)
      call sites: avod/core/models/dt_rpn_model.py:587-591, dt_avod_model.py:609-613

NumPy arguments are uploaded and the result comes back as NumPy; CUDA tensors stay on the device.
"""
import numpy as np
import torch

from . import ops
from ._lib import NMS_CHUNK_WINDOWS, NMS_WINDOW


def _dev(x, dtype):
    was_numpy = not torch.is_tensor(x)
    t = torch.as_tensor(np.asarray(x)) if was_numpy else x
    if t.dtype != dtype:
        t = t.to(dtype)
    if not t.is_cuda:
        t = t.cuda(non_blocking=True)
    return t, was_numpy


def crop_and_resize(image, boxes, box_ind, crop_size, method='bilinear',
                    extrapolation_value=0, name=None):
    """tf.image.crop_and_resize: image [batch, H, W, C] float32 NHWC, boxes [n, 4] normalised
    [y1, x1, y2, x2], box_ind [n] int32, crop_size [crop_h, crop_w] -> [n, crop_h, crop_w, C]."""
    if method != 'bilinear':
        raise ValueError("method must be 'bilinear' (the only method TensorFlow 1.3 supports)")
    img, np_in = _dev(image, torch.float32)
    bx, _ = _dev(boxes, torch.float32)
    bi, _ = _dev(box_ind, torch.int32)
    out = ops.crop_and_resize(img, bx, bi, crop_size, float(extrapolation_value))
    return out.cpu().numpy() if np_in else out


def non_max_suppression(boxes, scores, max_output_size, iou_threshold=0.5, name=None):
    """tf.image.non_max_suppression: boxes [n, 4], scores [n] -> int32 [m] selected indices in
    descending score order, m <= max_output_size. The output length is data dependent, so this
    wrapper synchronises once to read it; use ops.nms for the fixed-shape device form."""
    bx, np_in = _dev(boxes, torch.float32)
    sc, _ = _dev(scores, torch.float32)
    # Candidates are ordered lazily in score order; almost every selection completes within the
    # first few windows, so windows are enqueued in growing batches and the completion flag is
    # read between batches (the output length needs that read anyway).
    n = bx.shape[0]
    workspace = torch.empty(max(ops.nms_workspace_bytes(n), 256), dtype=torch.uint8, device=bx.device)
    keep = n_keep = None
    first, batch = 0, 2 * NMS_CHUNK_WINDOWS
    while True:
        keep, n_keep = ops.nms(bx, sc, int(max_output_size), float(iou_threshold), keep=keep,
                               n_keep=n_keep, workspace=workspace, first_window=first,
                               max_windows=batch)
        m, complete = n_keep.cpu().tolist()
        first += batch
        if complete or first * NMS_WINDOW >= n:
            break
        batch *= 4
    sel = keep[:m]
    return sel.cpu().numpy() if np_in else sel
