"""GPU parity of the device-side anchor geometry against oracle.anchor_helpers —
the NumPy restatement of the reference helpers that tests/test_oracle_vs_reference.py pins to the
LIVE reference — and against the answers the reference's unit tests assert."""
import numpy as np
import pytest
import torch

from oracle import anchor_helpers as A
from oracle import synth_ref as S

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(lib):
    from dodt_b200 import ops
    return ops


def test_grid_anchors_bit_exact(ops):
    """box_3d_to_anchor(tile_anchors_3d(...)): the 89 600 Car anchors and odd grids, bit for bit."""
    got = ops.grid_anchors(S.AREA_EXTENTS, A.CAR_ANCHOR_SIZES, S.ANCHOR_STRIDE, S.GROUND_PLANE).cpu().numpy()
    np.testing.assert_array_equal(got, S.car_anchors())
    for ext, sizes, stride, plane in (
            ([[-3.3, 7.1], [-1, 0], [0.2, 9.9]], [[3.9, 1.6, 1.5]], [0.7, 0.3], [0.01, -0.99, 0.02, 1.7]),
            ([[-40, 40], [-5, 3], [0, 70]], [[0.8, 0.6, 1.73], [1.76, 0.6, 1.73], [4.2, 1.7, 1.5]], [0.5, 1.0], [0, -1, 0, 1.65])):
        want = A.box_3d_to_anchor(A.tile_anchors_3d(ext, sizes, stride, plane))
        got = ops.grid_anchors(ext, sizes, stride, plane).cpu().numpy()
        np.testing.assert_array_equal(got, want)
    # grid_anchor_3d_generator_test.py:32-70
    want = np.array([[-0.5, 0., 0.5, 1., 1., 1.], [-0.5, 0., 0.5, 1., 1., 1.], [-0.5, 0., 0.5, 2., 1., 1.],
                     [-0.5, 0., 0.5, 1., 1., 2.], [0.5, 0., 0.5, 1., 1., 1.], [0.5, 0., 0.5, 1., 1., 1.],
                     [0.5, 0., 0.5, 2., 1., 1.], [0.5, 0., 0.5, 1., 1., 2.]])
    got = ops.grid_anchors([(-1., 1.), (-1., 0.), (0., 1.)], [[1., 1., 1.], [2., 1., 1.]], [1, 1], [0., -1., 0., 0.])
    np.testing.assert_allclose(got.cpu().numpy(), want, atol=1e-3)
    assert ops.grid_anchors([(0., 0.), (-1., 0.), (0., 2.)], [[1., 1., 1.]], [1, 1], [0., -1., 0., 0.]).shape == (0, 6)


def test_project_to_bev(ops):
    """anchor_projector_test.py:15-100 known answers, then the Car grid bit for bit."""
    anchors = torch.tensor([[1, 0, 3, 2, 0, 6], [3, 0, 3, 2, 0, 2]], dtype=torch.float64, device="cuda")
    metres, norm = ops.project_to_bev(anchors, [0, 5, 0, 10], want_metres=True)
    np.testing.assert_allclose(metres.cpu().numpy(), [[0, 4, 2, 10], [2, 6, 4, 8]], rtol=1e-5)
    np.testing.assert_allclose(norm.cpu().numpy(), np.array([[0, 4, 2, 10], [2, 6, 4, 8]]) / [5, 10, 5, 10], rtol=1e-5)
    anchors = torch.tensor([[0, 0, 0, 10, 0, 2]], dtype=torch.float64, device="cuda")
    metres, _ = ops.project_to_bev(anchors, [-3, 3, 0, 10], want_metres=True)
    np.testing.assert_allclose(metres.cpu().numpy(), [[-2, 9, 8, 11]], rtol=1e-5)
    a = S.car_anchors()
    corners, norm = A.project_to_bev(a, S.BEV_EXTENTS)
    dev = torch.from_numpy(a).cuda()
    g_metres, g_norm = ops.project_to_bev(dev, [-40, 40, 0, 70], want_metres=True)
    np.testing.assert_array_equal(g_norm.cpu().numpy(), norm.astype(np.float32))
    np.testing.assert_array_equal(g_metres.cpu().numpy(), corners.astype(np.float32))
    np.testing.assert_array_equal(ops.project_to_bev(dev, [-40, 40, 0, 70], tf_order=True).cpu().numpy(),
                                  A.reorder_projected_boxes(norm).astype(np.float32))
    # float32 anchors (tensors in the reference) are promoted exactly
    np.testing.assert_array_equal(ops.project_to_bev(dev.float(), [-40, 40, 0, 70]).cpu().numpy(),
                                  A.project_to_bev(a.astype(np.float32).astype(np.float64), S.BEV_EXTENTS)[1].astype(np.float32))
    with pytest.raises(TypeError):
        ops.project_to_bev(dev[:, :5], [-40, 40, 0, 70])


def test_project_to_image_space(ops):
    a = S.car_anchors()[::7]
    pixels, norm = A.project_to_image_space(a, A.KITTI_P2, S.IMAGE_SHAPE)
    g_pixels, g_norm = ops.project_to_image_space(torch.from_numpy(a).cuda(), A.KITTI_P2, S.IMAGE_SHAPE, want_pixels=True)
    # np.dot's summation order is BLAS-defined: float64 noise, i.e. at most one float32 ulp
    np.testing.assert_allclose(g_pixels.cpu().numpy(), pixels, rtol=2e-7, atol=1e-4)
    np.testing.assert_allclose(g_norm.cpu().numpy(), norm, rtol=2e-7, atol=1e-7)
    assert (g_norm.cpu().numpy() == norm).mean() > 0.99
    np.testing.assert_allclose(ops.project_to_image_space(torch.from_numpy(a).cuda(), A.KITTI_P2, S.IMAGE_SHAPE,
                                                          tf_order=True).cpu().numpy(),
                               A.reorder_projected_boxes(norm), rtol=2e-7, atol=1e-7)


def test_offset_to_anchor(ops):
    rng = np.random.default_rng(4)
    a = S.car_anchors()[::11]
    off = rng.normal(0, 0.2, a.shape)
    want = A.offset_to_anchor(a, off)
    for o in (off, off.astype(np.float32)):
        got = ops.offset_to_anchor(torch.from_numpy(a).cuda(), torch.from_numpy(o).cuda()).cpu().numpy()
        ref = A.offset_to_anchor(a, o.astype(np.float64))
        np.testing.assert_array_equal(got[:, :3], ref[:, :3])                 # products and sums: exact
        np.testing.assert_allclose(got[:, 3:], ref[:, 3:], rtol=1e-14)        # exp / log: libm vs CUDA
    assert want.shape == got.shape


def test_rpn_decode_chain_matches_host(ops):
    """The chain the reference runs between the RPN head and NMS (dt_rpn_model.py:573-591):
    offset_to_anchor -> project_to_bev; device boxes == host boxes, so NMS picks the same set."""
    import dodt_b200 as dd
    from oracle import np_oracle as O
    a = S.car_anchors()[::5]
    regressed, bev_norm, scores = S.rpn_proposals(8, 1, a)
    off = np.random.default_rng(1000 * 8 + 1 + 700000).normal(0.0, 0.1, (len(a), 6))
    dev_reg = ops.offset_to_anchor(torch.from_numpy(a).cuda(), torch.from_numpy(off).cuda())
    dev_boxes = ops.project_to_bev(dev_reg, [-40, 40, 0, 70])
    np.testing.assert_allclose(dev_boxes.cpu().numpy(), bev_norm, rtol=1e-6, atol=1e-7)
    keep = dd.non_max_suppression(dev_boxes, torch.from_numpy(scores).cuda(), 300, 0.8).cpu().numpy()
    np.testing.assert_array_equal(keep, O.non_max_suppression(dev_boxes.cpu().numpy(), scores, 300, 0.8))


def test_rpn_decode_tf_float32_branch(ops):
    """decode_f32: the tf.Tensor branches the reference's inference graph runs on float32 placeholders
    (dt_rpn_model.py:568-591), op by op in float32 with correctly rounded exp / log == the float32
    restatement in oracle/anchor_helpers.py, bit for bit (a float64 exp / log result within 1e-16 of
    a float32 tie could round the other way: not observed); the float64 NumPy branch differs from it
    by a few float32 ulps, which is why the runner follows the graph."""
    import dodt_b200 as dd
    from oracle import np_oracle as O
    a = S.car_anchors()
    rng = np.random.default_rng(21)
    off = rng.normal(0.0, 0.1, (len(a), 6)).astype(np.float32)
    reg = A.offset_to_anchor_tf32(a, off)
    _, bev = A.project_to_bev_tf32(reg, S.BEV_EXTENTS)
    want_bev = A.reorder_projected_boxes(bev)
    _, img = A.project_to_image_space(reg.astype(np.float64), A.KITTI_P2, S.IMAGE_SHAPE)
    want_img = A.reorder_projected_boxes(img)
    n = len(a)
    idx = torch.arange(n, dtype=torch.int32, device="cuda")
    cnt = torch.tensor([n], dtype=torch.int32, device="cuda")
    got_bev = torch.empty((n, 4), dtype=torch.float32, device="cuda")
    got_img = torch.empty((n, 4), dtype=torch.float32, device="cuda")
    t_a, t_off = torch.from_numpy(a).cuda(), torch.from_numpy(off).cuda()
    ops.rpn_decode(t_a, t_off, idx, cnt, [-40, 40, 0, 70], A.KITTI_P2.reshape(-1), S.IMAGE_SHAPE, got_bev, got_img,
                   tf_float32=True)
    np.testing.assert_array_equal(got_bev.cpu().numpy(), want_bev)
    np.testing.assert_allclose(got_img.cpu().numpy(), want_img, rtol=1e-6, atol=1e-7)
    f64_bev = torch.empty_like(got_bev)
    ops.rpn_decode(t_a, t_off, idx, cnt, [-40, 40, 0, 70], A.KITTI_P2.reshape(-1), S.IMAGE_SHAPE, f64_bev, None)
    diff = (f64_bev != got_bev).float().mean().item()
    assert 0.0 < diff < 0.9                                   # the two branches are different roundings
    np.testing.assert_allclose(f64_bev.cpu().numpy(), want_bev, rtol=2e-5, atol=2e-6)
    # NMS on the graph's boxes: the device picks what the oracle picks on the oracle's boxes
    scores = rng.permutation(np.linspace(0.01, 0.99, n)).astype(np.float32)
    sub = slice(0, n, 7)
    keep = dd.non_max_suppression(got_bev[sub], torch.from_numpy(scores[sub]).cuda(), 1024, 0.8).cpu().numpy()
    np.testing.assert_array_equal(keep, O.non_max_suppression(want_bev[sub], scores[sub], 1024, 0.8))
