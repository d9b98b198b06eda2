"""Drop-in for avod/core/corr_layers/correlation.py:7-27 (the `correlation` wrapper of the
reference's TensorFlow custom op, avod/core/ops/correlation/correlation_op.cc:53-62), forward only.

Same signature and defaults; inputs are NHWC float32 [batch, H, W, C], output is
[batch, out_h, out_w, (2*(max_displacement // stride_2) + 1)**2] with the reference's shape rule
(correlation_kernel.cc:39-57). Errors follow the op: even kernel_size, rank != 4 and an empty
output raise ValueError (InvalidArgument in TF).
"""
import numpy as np
import torch

from . import ops


def correlation(input_a, input_b, kernel_size=1, max_displacement=20, stride_1=1, stride_2=2,
                padding=20):
    was_numpy = not torch.is_tensor(input_a)
    a = torch.as_tensor(np.asarray(input_a, dtype=np.float32)) if was_numpy else input_a
    b = torch.as_tensor(np.asarray(input_b, dtype=np.float32)) if not torch.is_tensor(input_b) \
        else input_b
    if not a.is_cuda:
        a = a.cuda(non_blocking=True)
    if not b.is_cuda:
        b = b.cuda(non_blocking=True)
    out = ops.correlation(a, b, int(kernel_size), int(max_displacement), int(stride_1),
                          int(stride_2), int(padding))
    return out.cpu().numpy() if was_numpy else out
