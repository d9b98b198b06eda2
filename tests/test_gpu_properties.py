"""Size-independent properties of the stages at BASELINE.json's FULL sizes (configs B, C and E), where
the CPU oracle would take minutes: exact scalings and shifts of the correlation, partition of unity
of the bilinear crop, the independent-set / maximality / idempotence properties of greedy NMS on
89 600 boxes, checksums of the BEV counts and of the integral image. Everything goes through the
reference-signature drop-ins (-> C ABI -> kernels)."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O
from oracle import synth_ref as S

pytestmark = pytest.mark.gpu
KW = dict(kernel_size=1, max_displacement=5, stride_1=1, stride_2=2, padding=5)


@pytest.fixture(scope="module")
def dd(lib):
    import dodt_b200
    return dodt_b200


def test_correlation_full_size_scaling_shift_and_symmetry(dd):
    """[1,700,800,32] pair (config C). (1) scaling an input by a power of two scales every output
    exactly; (2) additivity in the second input within 1e-5; (3) moving BOTH maps by an even number
    of pixels moves the output with them, bit for bit, wherever no tap crosses the border (zero
    padding = the reference's PadData); (4) corr(A, B)[y, x, (p, o)] is the same sum as
    corr(B, A)[y + 2p, x + 2o, (-p, -o)] with the channels added in the same order: bit-identical."""
    f0, f1 = S.feature_pair(4, 1)
    a, b = torch.from_numpy(f0).cuda(), torch.from_numpy(f1).cuda()
    base = dd.correlation(a, b, **KW)
    assert base.shape == (1, 700, 800, 25)
    assert torch.equal(dd.correlation(a * 4.0, b, **KW), base * 4.0)
    assert torch.equal(dd.correlation(a, b * 0.5, **KW), base * 0.5)
    b2 = torch.from_numpy(np.abs(np.random.default_rng(3).standard_normal(f1.shape)).astype(np.float32)).cuda()
    both = dd.correlation(a, b + b2, **KW)
    np.testing.assert_allclose(both.cpu().numpy(), (base + dd.correlation(a, b2, **KW)).cpu().numpy(),
                               rtol=1e-5, atol=1e-6)
    dy, dx = 6, 10
    sa, sb = torch.roll(a, (dy, dx), (1, 2)), torch.roll(b, (dy, dx), (1, 2))
    moved = dd.correlation(sa, sb, **KW)
    assert torch.equal(moved[:, dy + 4:700 - 4, dx + 4:800 - 4], base[:, 4:700 - 4 - dy, 4:800 - 4 - dx])
    swapped = dd.correlation(b, a, **KW)
    for p, o in ((-2, -2), (-1, 2), (0, 0), (2, 1), (1, -2)):
        k, k_rev = (p + 2) * 5 + (o + 2), (-p + 2) * 5 + (-o + 2)
        y0, y1, x0, x1 = max(0, -2 * p), min(700, 700 - 2 * p), max(0, -2 * o), min(800, 800 - 2 * o)
        assert torch.equal(base[0, y0:y1, x0:x1, k], swapped[0, y0 + 2 * p:y1 + 2 * p, x0 + 2 * o:x1 + 2 * o, k_rev])


def test_correlation_stream_full_size_equals_pairwise(dd):
    """Five consecutive config-C maps in ONE frame-stream launch == four pairwise calls, bit for bit."""
    maps = [torch.from_numpy(S.feature_pair(5, k)[0]).cuda() for k in range(5)]
    outs = dd.correlation_stream(maps, **KW)
    assert len(outs) == 4
    for k in range(4):
        assert torch.equal(outs[k], dd.correlation(maps[k], maps[k + 1], **KW))


def test_crop_and_resize_full_size_partition_of_unity_and_identity(dd):
    """1024 proposal boxes x 7x7 on the config-B/C maps: (1) a constant map gives that constant at
    every sample inside the map and the extrapolation value outside, exactly (the bilinear weights of
    the TF formula cancel in top + (bottom - top) * lerp); (2) crops are linear in the image within
    1e-5; (3) the box [0, 0, 1, 1] with crop size = map size samples every pixel centre: identity."""
    rng = np.random.default_rng(12)
    n = 1024
    c = rng.uniform(-0.05, 1.05, (n, 2))
    h = rng.uniform(0.01, 0.08, (n, 2))
    boxes = np.stack([c[:, 0] - h[:, 0], c[:, 1] - h[:, 1], c[:, 0] + h[:, 0], c[:, 1] + h[:, 1]], 1).astype(np.float32)
    t_boxes = torch.from_numpy(boxes).cuda()
    ind = torch.zeros(n, dtype=torch.int32, device="cuda")
    for (H, W, C) in ((700, 800, 32), (360, 1200, 32), (700, 800, 25)):
        const = torch.full((1, H, W, C), 3.25, device="cuda")
        out = dd.crop_and_resize(const, t_boxes, ind, (7, 7), extrapolation_value=-1.0)
        assert out.shape == (n, 7, 7, C)
        ys = boxes[:, 0:1] * (H - 1) + np.arange(7, dtype=np.float32)[None] * ((boxes[:, 2:3] - boxes[:, 0:1]) * (H - 1) / 6)
        xs = boxes[:, 1:2] * (W - 1) + np.arange(7, dtype=np.float32)[None] * ((boxes[:, 3:4] - boxes[:, 1:2]) * (W - 1) / 6)
        sure_in = ((ys > 0.01) & (ys < H - 1.01))[:, :, None] & ((xs > 0.01) & (xs < W - 1.01))[:, None, :]
        sure_out = ((ys < -0.01) | (ys > H - 0.99))[:, :, None] | ((xs < -0.01) | (xs > W - 0.99))[:, None, :]
        got = out.cpu().numpy()
        assert sure_in.sum() > 20000 and sure_out.sum() > 1000
        assert (got[sure_in] == 3.25).all() and (got[sure_out] == -1.0).all()
        assert np.isin(got, (3.25, -1.0)).all()
    img = torch.from_numpy(np.abs(np.random.default_rng(13).standard_normal((1, 360, 1200, 32))).astype(np.float32)).cuda()
    img2 = torch.roll(img, 7, 2)
    lhs = dd.crop_and_resize(img + 2.0 * img2, t_boxes, ind, (7, 7))
    rhs = dd.crop_and_resize(img, t_boxes, ind, (7, 7)) + 2.0 * dd.crop_and_resize(img2, t_boxes, ind, (7, 7))
    np.testing.assert_allclose(lhs.cpu().numpy(), rhs.cpu().numpy(), rtol=1e-5, atol=1e-5)
    small = img[:, :90, :128, :8].contiguous()
    whole = dd.crop_and_resize(small, torch.tensor([[0.0, 0.0, 1.0, 1.0]], device="cuda"),
                               torch.zeros(1, dtype=torch.int32, device="cuda"), (90, 128))
    np.testing.assert_allclose(whole.cpu().numpy(), small.cpu().numpy(), rtol=1e-5, atol=1e-6)


def _iou(b, others):
    """TF 1.3 non_max_suppression_op.cc IoU of box b against rows of others ([y1,x1,y2,x2])."""
    ymin, xmin = np.minimum(others[:, 0], others[:, 2]), np.minimum(others[:, 1], others[:, 3])
    ymax, xmax = np.maximum(others[:, 0], others[:, 2]), np.maximum(others[:, 1], others[:, 3])
    bymin, bxmin, bymax, bxmax = min(b[0], b[2]), min(b[1], b[3]), max(b[0], b[2]), max(b[1], b[3])
    area = (ymax - ymin) * (xmax - xmin)
    barea = (bymax - bymin) * (bxmax - bxmin)
    ih = np.maximum(np.minimum(ymax, bymax) - np.maximum(ymin, bymin), 0)
    iw = np.maximum(np.minimum(xmax, bxmax) - np.maximum(xmin, bxmin), 0)
    inter = ih * iw
    return np.where((area > 0) & (barea > 0), inter / (area + barea - inter), 0)


def test_nms_full_size_independent_maximal_idempotent(dd):
    """All 89 600 regressed anchors (the stress size of SURVEY 8(d)), IoU 0.8, 1024 outputs:
    the kept boxes come in descending score order, no two of them overlap by more than the threshold,
    every better-scored box that was NOT kept is suppressed by a kept box that beats it (greedy
    maximality, checked for all candidates ahead of the last kept one), and NMS of the kept set
    keeps all of it (idempotence)."""
    anchors = S.car_anchors()
    _, boxes, scores = S.rpn_proposals(6, 2, anchors)
    assert len(boxes) == 89600
    max_out = 1024
    # clustered scores (the best-scored anchors crowd around 40 spots, as real RPN outputs do around
    # objects): at IoU 0.5 thousands of candidates are rejected before 1024 are found
    rng = np.random.default_rng(77)
    spots = np.stack([rng.uniform(-30, 30, 40), rng.uniform(5, 60, 40)], 1)
    dist = np.sqrt(((anchors[:, None, [0, 2]] - spots[None]) ** 2).sum(-1)).min(1) + rng.normal(0, 0.3, len(anchors))
    clustered = np.empty(len(anchors), np.float32)
    clustered[np.argsort(dist, kind="stable")] = np.linspace(0.99, 0.01, len(anchors)).astype(np.float32)
    assert len(np.unique(clustered)) == len(anchors)
    random_scores = scores
    for thr, scores, min_rejected in ((0.8, random_scores, 1), (0.5, random_scores, 50), (0.8, clustered, 30),
                                      (0.5, clustered, 2000)):
        keep = dd.non_max_suppression(boxes, scores, max_out, thr)
        assert keep.dtype == np.int32 and len(keep) == max_out and len(set(keep.tolist())) == max_out
        ks = scores[keep]
        assert (np.diff(ks) < 0).all()
        kb = boxes[keep].astype(np.float32)
        for i in range(1, max_out):
            assert (_iou(kb[i], kb[:i]) <= thr + 1e-6).all()
        order = np.argsort(-scores, kind="stable")
        rank_of_last = int(np.flatnonzero(order == keep[-1])[0])
        kept_set = set(keep.tolist())
        rejected = [j for j in order[:rank_of_last] if j not in kept_set]
        assert len(rejected) == rank_of_last + 1 - max_out and len(rejected) >= min_rejected, len(rejected)
        for j in rejected:
            better = kb[ks > scores[j]]
            assert (_iou(boxes[j], better) > thr - 1e-6).any()
        again = dd.non_max_suppression(kb, ks, max_out, thr)
        np.testing.assert_array_equal(again, np.arange(max_out, dtype=np.int32))
    # second-stage NMS on the kept boxes (0.01, 100): the same three properties at its own threshold
    final = dd.non_max_suppression(kb, ks, 100, 0.01)
    fb = kb[final]
    for i in range(1, len(final)):
        assert (_iou(fb[i], fb[:i]) <= 0.01 + 1e-6).all()
    np.testing.assert_array_equal(dd.non_max_suppression(fb, ks[final], 100, 0.01), np.arange(len(final)))


def test_bev_and_integral_image_checksums_config_e(dd):
    """configs[4] (500 k points, 1400x1600 grid): the per-cell counts add up to the number of points in
    the density slice; a cell holds a height in slice s iff some point of slice s fell into it (number
    of non-zero winners = number of occupied cells per slice, from the per-cell winner indices: all
    distinct point indices); every winner lies in its cell; the bottom-right entry of the integral
    image is the number of occupied cells and the image is monotone along both axes; the keep mask
    shrinks as the density threshold grows."""
    from dodt_b200 import ops
    from dodt_b200.bev_slices import _to_device_points
    pc = S.point_cloud(5, 3, n_points=500000)
    gen = dd.BevSlices(S.SlicesConfig())
    pts, _ = _to_device_points(pc, gen.device)
    buf = gen.generate_bev_device('lidar', pts, S.GROUND_PLANE, S.AREA_EXTENTS, S.VOXEL_SIZE_DENSE,
                                  with_occupancy=True, debug=True)
    stats = buf.stats.cpu().numpy()
    counts = buf.counts.cpu().numpy()
    assert counts.shape == (1400, 1600) and counts.sum() == stats[16] and stats[19] == 0 and stats[20] == 0
    assert stats[:5].sum() == stats[16]                      # the five slices partition the density range
    winner = buf.winner.cpu().numpy()
    h = 1.65 - pc[1].astype(np.float64)
    for s in range(5):
        w = winner[s][winner[s] >= 0]
        assert len(np.unique(w)) == len(w)
        lo = np.float32(S.HEIGHT_LO) + s * 0.5
        assert ((h[w] > lo - 1e-6) & (h[w] < lo + 0.5 + 1e-6)).all()
        rows, cols = np.nonzero(winner[s] >= 0)
        v = float(S.VOXEL_SIZE_DENSE)
        assert (np.floor(pc[0][winner[s][rows, cols]].astype(np.float64) / v) + 800 == cols).all()
        assert (1399 - np.floor(pc[2][winner[s][rows, cols]].astype(np.float64) / v) == rows).all()
    assert ((counts > 0) == (buf.maps[5].cpu().numpy() > 0)).all()
    occ = buf.occ
    ii = ops.integral_image_2d(occ).cpu().numpy()
    assert ii.shape == (1601, 1401) and ii[-1, -1] == int(occ.sum().item())
    assert (np.diff(ii, axis=0) >= 0).all() and (np.diff(ii, axis=1) >= 0).all() and not ii[0].any() and not ii[:, 0].any()
    grid = dd.VoxelGrid2D.from_occupancy(occ, S.VOXEL_SIZE_DENSE, S.AREA_EXTENTS)
    anchors = S.car_anchors()
    k1 = dd.get_empty_anchor_filter_2d(anchors, grid, 1)
    k5 = dd.get_empty_anchor_filter_2d(anchors, grid, 5)
    k50 = dd.get_empty_anchor_filter_2d(anchors, grid, 50)
    assert (k5 <= k1).all() and (k50 <= k5).all() and k50.sum() < k5.sum() < k1.sum() < len(anchors)
