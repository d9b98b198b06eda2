"""S4 backward: the CorrelationGrad op
(avod/core/ops/correlation/correlation_grad_kernel.cu.cc:20-189).

Pin of the oracle (CPU, no GPU): the reference op is a TF 1.3 GPU custom op that cannot be built
here ("parity unpinned by reference outputs"), so oracle/c_oracle.c's loop-
for-loop restatement is checked against what the gradient IS: torch.autograd through an
independent float64 restatement of the FORWARD formula (correlation_kernel.cu.cc:45-110).
GPU parity: dodt_correlation_grad against the oracle, bit for bit (same fused multiply-adds in
the same order), through the C ABI.
"""
import numpy as np
import pytest
import torch

from oracle import c_oracle as CO
from oracle import corr_autograd as AG
from oracle import np_oracle as O

CASES = [
    ((1, 20, 24, 8), dict(kernel_size=1, max_displacement=5, stride_1=1, stride_2=2, padding=5)),   # DODT
    ((2, 11, 13, 16), dict(kernel_size=1, max_displacement=5, stride_1=1, stride_2=2, padding=5)),
    ((1, 12, 12, 8), dict(kernel_size=1, max_displacement=2, stride_1=1, stride_2=2, padding=2)),   # r = 1
    ((1, 14, 12, 8), dict(kernel_size=1, max_displacement=4, stride_1=1, stride_2=2, padding=6)),   # shift < 0
    ((1, 14, 15, 8), dict(kernel_size=1, max_displacement=5, stride_1=1, stride_2=2, padding=4)),   # shift > 0
    ((1, 12, 10, 5), dict(kernel_size=1, max_displacement=4, stride_1=1, stride_2=1, padding=4)),
    ((1, 15, 13, 3), dict(kernel_size=3, max_displacement=4, stride_1=2, stride_2=2, padding=4)),
    ((2, 9, 10, 4), dict(kernel_size=3, max_displacement=3, stride_1=1, stride_2=1, padding=5)),
    ((1, 8, 8, 2), dict(kernel_size=1, max_displacement=20, stride_1=1, stride_2=2, padding=20)),   # defaults
]


def _inputs(shape, kw, seed=0):
    rng = np.random.default_rng(seed + shape[1] * 31 + shape[3])
    a = rng.standard_normal(shape).astype(np.float32)
    b = rng.standard_normal(shape).astype(np.float32)
    oh, ow, oc = O.correlation_out_shape(shape[1], shape[2], kw["kernel_size"], kw["max_displacement"],
                                         kw["stride_1"], kw["stride_2"], kw["padding"])
    g = rng.standard_normal((shape[0], oh, ow, oc)).astype(np.float32)
    return a, b, g


@pytest.mark.parametrize("shape,kw", CASES)
def test_oracle_grad_is_the_gradient_of_the_forward(shape, kw):
    a, b, g = _inputs(shape, kw)
    out, want_a, want_b = AG.gradients_f64(a, b, g, **kw)
    # the differentiated forward is the forward the oracle (and the kernels) compute
    np.testing.assert_allclose(out, CO.correlation(a, b, **kw), rtol=1e-5, atol=1e-6)
    ga, gb = CO.correlation_grad(g, a, b, **kw)
    assert ga.shape == a.shape and gb.shape == b.shape and ga.dtype == np.float32
    scale = max(np.abs(want_a).max(), 1e-30)
    np.testing.assert_allclose(ga, want_a, rtol=1e-5, atol=2e-6 * scale)
    np.testing.assert_allclose(gb, want_b, rtol=1e-5, atol=2e-6 * scale)


def _golden():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "s4_grad_autograd.npz"))


def test_oracle_grad_against_frozen_autograd_vectors():
    """tests/golden/s4_grad_autograd.npz (python -m oracle.corr_autograd): float64 autograd
    gradients frozen with their inputs; the C oracle must reproduce them."""
    gold = _golden()
    for i in range(len(AG.GOLDEN_CASES)):
        ks, md, s1, s2, pad = (int(v) for v in gold["attrs%d" % i])
        kw = dict(kernel_size=ks, max_displacement=md, stride_1=s1, stride_2=s2, padding=pad)
        a, b, g = gold["a%d" % i], gold["b%d" % i], gold["g%d" % i]
        np.testing.assert_allclose(CO.correlation(a, b, **kw), gold["out%d" % i], rtol=1e-5, atol=1e-6)
        ga, gb = CO.correlation_grad(g, a, b, **kw)
        scale = np.abs(gold["ga%d" % i]).max()
        np.testing.assert_allclose(ga, gold["ga%d" % i], rtol=1e-5, atol=2e-6 * scale)
        np.testing.assert_allclose(gb, gold["gb%d" % i], rtol=1e-5, atol=2e-6 * scale)


def test_oracle_grad_adjoint_identity():
    """<G, corr(A, B)> is bilinear: <gA, A> = <gB, B> = <G, corr(A,B)> (float64 accumulation)."""
    shape, kw = CASES[0]
    a, b, g = _inputs(shape, kw, seed=5)
    ga, gb = CO.correlation_grad(g, a, b, **kw)
    f = CO.correlation(a, b, **kw)
    want = np.sum(g.astype(np.float64) * f)
    assert abs(np.sum(ga.astype(np.float64) * a) - want) < 1e-5 * abs(want) + 1e-6
    assert abs(np.sum(gb.astype(np.float64) * b) - want) < 1e-5 * abs(want) + 1e-6


# ------------------------------------------------------------------------------------------ GPU


@pytest.fixture(scope="module")
def dd():
    import dodt_b200
    return dodt_b200


GPU_CASES = CASES + [
    ((1, 40, 72, 32), dict(kernel_size=1, max_displacement=5, stride_1=1, stride_2=2, padding=5)),
    ((2, 19, 150, 32), dict(kernel_size=1, max_displacement=5, stride_1=1, stride_2=2, padding=5)),
    ((1, 33, 70, 24), dict(kernel_size=1, max_displacement=3, stride_1=1, stride_2=2, padding=3)),
    ((1, 17, 66, 16), dict(kernel_size=1, max_displacement=5, stride_1=1, stride_2=2, padding=3)),
    ((1, 17, 66, 16), dict(kernel_size=1, max_displacement=4, stride_1=1, stride_2=2, padding=7)),
]


@pytest.mark.gpu
@pytest.mark.parametrize("shape,kw", GPU_CASES)
def test_gpu_correlation_grad_matches_oracle(dd, shape, kw):
    a, b, g = _inputs(shape, kw, seed=1)
    want_a, want_b = CO.correlation_grad(g, a, b, **kw)
    kw2 = dict(kw)
    kw2["pad"] = kw2.pop("padding")
    got_a, got_b = dd.correlation_grad(g, a, b, **kw2)
    np.testing.assert_array_equal(got_a, want_a)
    np.testing.assert_array_equal(got_b, want_b)


@pytest.mark.gpu
def test_gpu_correlation_grad_against_frozen_autograd_vectors(dd):
    gold = _golden()
    for i in range(len(AG.GOLDEN_CASES)):
        ks, md, s1, s2, pad = (int(v) for v in gold["attrs%d" % i])
        ga, gb = dd.correlation_grad(gold["g%d" % i], gold["a%d" % i], gold["b%d" % i], ks, md, s1, s2, pad)
        scale = np.abs(gold["ga%d" % i]).max()
        np.testing.assert_allclose(ga, gold["ga%d" % i], rtol=1e-5, atol=2e-6 * scale)
        np.testing.assert_allclose(gb, gold["gb%d" % i], rtol=1e-5, atol=2e-6 * scale)


@pytest.mark.gpu
def test_gpu_correlation_autograd_and_partial_grads(dd):
    shape, kw = GPU_CASES[-5]
    a, b, g = _inputs(shape, kw, seed=2)
    want_a, want_b = CO.correlation_grad(g, a, b, **kw)
    ta = torch.from_numpy(a).cuda().requires_grad_(True)
    tb = torch.from_numpy(b).cuda().requires_grad_(True)
    out = dd.correlation(ta, tb, **kw)
    out.backward(torch.from_numpy(g).cuda())
    np.testing.assert_array_equal(ta.grad.cpu().numpy(), want_a)
    np.testing.assert_array_equal(tb.grad.cpu().numpy(), want_b)
    # only one input requires grad: the other gradient is not computed
    tb2 = torch.from_numpy(b).cuda().requires_grad_(True)
    out = dd.correlation(torch.from_numpy(a).cuda(), tb2, **kw)
    out.backward(torch.from_numpy(g).cuda())
    np.testing.assert_array_equal(tb2.grad.cpu().numpy(), want_b)


@pytest.mark.gpu
def test_gpu_correlation_grad_full_size_properties(dd):
    """Config C size [1,700,800,32]: the oracle on border / interior strips, and the adjoint
    identity <gA, A> = <G, corr(A, B)> = <gB, B> over the whole map."""
    from oracle import synth_ref as synth
    f0, f1 = synth.feature_pair(3, 0)
    kw = dict(kernel_size=1, max_displacement=5, stride_1=1, stride_2=2, padding=5)
    rng = np.random.default_rng(11)
    g = rng.standard_normal((1, 700, 800, 25)).astype(np.float32)
    ta, tb, tg = (torch.from_numpy(x).cuda() for x in (f0, f1, g))
    ga, gb = dd.correlation_grad(tg, ta, tb, 1, 5, 1, 2, 5)
    out = dd.correlation(ta, tb, **kw)
    want = float((tg.double() * out.double()).sum())
    assert abs(float((ga.double() * ta.double()).sum()) - want) < 1e-5 * abs(want)
    assert abs(float((gb.double() * tb.double()).sum()) - want) < 1e-5 * abs(want)
    ga, gb = ga.cpu().numpy(), gb.cpu().numpy()
    for lo, hi in ((0, 20), (340, 360), (680, 700)):
        wa, wb = CO.correlation_grad(g[:, lo:hi], f0[:, lo:hi], f1[:, lo:hi], **kw)
        v0 = lo if lo == 0 else lo + 4
        v1 = hi if hi == 700 else hi - 4
        np.testing.assert_array_equal(ga[:, v0:v1], wa[:, v0 - lo:v1 - lo])
        np.testing.assert_array_equal(gb[:, v0:v1], wb[:, v0 - lo:v1 - lo])


@pytest.mark.gpu
def test_gpu_correlation_grad_errors(dd):
    a = np.zeros((1, 8, 8, 4), dtype=np.float32)
    g = np.zeros((1, 8, 8, 25), dtype=np.float32)
    with pytest.raises(ValueError):
        dd.correlation_grad(g, a, a, kernel_size=2, max_displacement=5, pad=5)
    with pytest.raises(ValueError):
        dd.correlation_grad(g, a[0], a[0], max_displacement=5, pad=5)
    with pytest.raises(ValueError):
        dd.correlation_grad(g[:, :4], a, a, max_displacement=5, pad=5)     # gradient of another shape
