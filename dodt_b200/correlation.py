"""Drop-in for avod/core/corr_layers/correlation.py (the wrappers of the reference's TensorFlow
custom ops, avod/core/ops/correlation/correlation_op.cc:53-83): `correlation` (:7-27, forward) and
`correlation_grad` (the CorrelationGrad op that `_correlation_grad`, :30-48, registers as the
gradient of "Correlation").

Same signatures and defaults; inputs are NHWC float32 [batch, H, W, C], the forward output is
[batch, out_h, out_w, (2*(max_displacement // stride_2) + 1)**2] with the reference's shape rule
(correlation_kernel.cc:39-57). Errors follow the ops: even kernel_size, rank != 4 and an empty
output raise ValueError (InvalidArgument in TF). `correlation` is differentiable when its inputs
are CUDA tensors that require grad (torch.autograd plays the role of tf.RegisterGradient).
"""
import numpy as np
import torch

from . import ops


def _as_cuda(x):
    t = x if torch.is_tensor(x) else torch.as_tensor(np.asarray(x, dtype=np.float32))
    return t if t.is_cuda else t.cuda(non_blocking=True)


class _Correlation(torch.autograd.Function):
    """forward = dodt_correlation, backward = dodt_correlation_grad
    (correlation.py:30-48: returns backprops_a, backprops_b)."""

    @staticmethod
    def forward(ctx, a, b, kernel_size, max_displacement, stride_1, stride_2, padding):
        ctx.save_for_backward(a, b)
        ctx.attrs = (kernel_size, max_displacement, stride_1, stride_2, padding)
        return ops.correlation(a, b, kernel_size, max_displacement, stride_1, stride_2, padding)

    @staticmethod
    def backward(ctx, gradients):
        a, b = ctx.saved_tensors
        ga, gb = ops.correlation_grad(gradients, a, b, *ctx.attrs, need_a=ctx.needs_input_grad[0],
                                      need_b=ctx.needs_input_grad[1])
        return ga, gb, None, None, None, None, None


def correlation(input_a, input_b, kernel_size=1, max_displacement=20, stride_1=1, stride_2=2,
                padding=20):
    was_numpy = not torch.is_tensor(input_a)
    a, b = _as_cuda(input_a), _as_cuda(input_b)
    attrs = (int(kernel_size), int(max_displacement), int(stride_1), int(stride_2), int(padding))
    if torch.is_grad_enabled() and (a.requires_grad or b.requires_grad):
        return _Correlation.apply(a, b, *attrs)
    out = ops.correlation(a, b, *attrs)
    return out.cpu().numpy() if was_numpy else out


def correlation_stream(feature_maps, kernel_size=1, max_displacement=20, stride_1=1, stride_2=2,
                       padding=20):
    """`correlation` over the consecutive key frames of a sequence: returns
    [correlation(f[0], f[1]), correlation(f[1], f[2]), ...], each bit-identical to the pairwise call.
    The reference runs the op once per sample pair (dt_rpn_model.py:324-331 inside the inference
    loop) and therefore reads every BEV feature map twice; here neighbouring pairs share a launch
    and the common map comes from HBM once. feature_maps: sequence of [1, H, W, C] arrays/tensors."""
    was_numpy = not torch.is_tensor(feature_maps[0])
    maps = [_as_cuda(m) for m in feature_maps]
    outs = ops.correlation_stream(maps, int(kernel_size), int(max_displacement), int(stride_1),
                                  int(stride_2), int(padding))
    return [o.cpu().numpy() for o in outs] if was_numpy else outs


def correlation_grad(gradients, input_a, input_b, kernel_size=1, max_displacement=20, stride_1=1,
                     stride_2=2, pad=20):
    """The CorrelationGrad op (argument order of correlation.py:37-44): returns
    (backprops_a, backprops_b), each shaped like the inputs."""
    was_numpy = not torch.is_tensor(input_a)
    g, a, b = _as_cuda(gradients), _as_cuda(input_a), _as_cuda(input_b)
    ga, gb = ops.correlation_grad(g, a, b, int(kernel_size), int(max_displacement), int(stride_1),
                                  int(stride_2), int(pad))
    if was_numpy:
        return ga.cpu().numpy(), gb.cpu().numpy()
    return ga, gb
