"""TEST INFRASTRUCTURE ONLY (see oracle/np_oracle.py header): an independent float64 restatement of
the correlation FORWARD built from torch slicing, so that torch.autograd yields the gradients the
reference's CorrelationGrad op (avod/core/ops/correlation/correlation_grad_kernel.cu.cc:20-189)
must produce. Used to pin oracle/c_oracle.c:oracle_correlation_grad, which cannot be pinned by
reference outputs (the op is a GPU-only TensorFlow 1.3 custom op).

  python -m oracle.corr_autograd     regenerates tests/golden/s4_grad_autograd.npz
"""
import os

import numpy as np
import torch

from .np_oracle import correlation_out_shape


def forward_f64(a, b, kernel_size, max_displacement, stride_1, stride_2, padding):
    """correlation_kernel.cu.cc:45-110 on torch tensors [N,H,W,C] (any dtype; use float64): inputs
    are padded by `padding` plus a zero margin that makes every displaced patch addressable."""
    ks, md, s1, s2, pad = kernel_size, max_displacement, stride_1, stride_2, padding
    N, H, W, C = a.shape
    oh, ow, oc = correlation_out_shape(H, W, ks, md, s1, s2, pad)
    r = md // s2
    wn = 2 * r + 1
    ex = r * s2 + md + ks + s1 * max(oh, ow)
    P = pad + ex
    ap = torch.nn.functional.pad(a, (0, 0, P, P, P, P))
    bp = torch.nn.functional.pad(b, (0, 0, P, P, P, P))
    outs = []
    for k in range(oc):
        s2o, s2p = (k % wn - r) * s2, (k // wn - r) * s2
        acc = 0
        for j in range(ks):
            for i in range(ks):
                y0, x0 = md + ex + j, md + ex + i
                pa = ap[:, y0:y0 + s1 * oh:s1, x0:x0 + s1 * ow:s1]
                pb = bp[:, y0 + s2p:y0 + s2p + s1 * oh:s1, x0 + s2o:x0 + s2o + s1 * ow:s1]
                acc = acc + (pa * pb).sum(-1)
        outs.append(acc / (ks * ks * C))
    return torch.stack(outs, dim=-1)


def gradients_f64(a, b, g, **kw):
    """(forward, d/da, d/db) of sum(g * correlation(a, b)) in float64; NumPy in, NumPy out."""
    ta = torch.from_numpy(np.asarray(a)).double().requires_grad_(True)
    tb = torch.from_numpy(np.asarray(b)).double().requires_grad_(True)
    out = forward_f64(ta, tb, **kw)
    out.backward(torch.from_numpy(np.asarray(g)).double())
    return out.detach().numpy(), ta.grad.numpy(), tb.grad.numpy()


GOLDEN_CASES = [
    ((1, 14, 18, 8), dict(kernel_size=1, max_displacement=5, stride_1=1, stride_2=2, padding=5)),   # DODT
    ((2, 9, 11, 4), dict(kernel_size=3, max_displacement=4, stride_1=2, stride_2=2, padding=4)),
    ((1, 10, 12, 8), dict(kernel_size=1, max_displacement=4, stride_1=1, stride_2=2, padding=6)),
]


def make_golden(path):
    rng = np.random.default_rng(2024)
    out = {}
    for i, (shape, kw) in enumerate(GOLDEN_CASES):
        a = rng.standard_normal(shape).astype(np.float32)
        b = rng.standard_normal(shape).astype(np.float32)
        oh, ow, oc = correlation_out_shape(shape[1], shape[2], kw["kernel_size"], kw["max_displacement"],
                                           kw["stride_1"], kw["stride_2"], kw["padding"])
        g = rng.standard_normal((shape[0], oh, ow, oc)).astype(np.float32)
        f, ga, gb = gradients_f64(a, b, g, **kw)
        out.update({"a%d" % i: a, "b%d" % i: b, "g%d" % i: g, "out%d" % i: f, "ga%d" % i: ga, "gb%d" % i: gb,
                    "attrs%d" % i: np.array([kw["kernel_size"], kw["max_displacement"], kw["stride_1"],
                                             kw["stride_2"], kw["padding"]], dtype=np.int32)})
    np.savez_compressed(path, **out)
    return path


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    p = make_golden(os.path.join(here, "..", "tests", "golden", "s4_grad_autograd.npz"))
    print(os.path.getsize(p), p, "torch", torch.__version__)
