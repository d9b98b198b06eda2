"""S2 as the frame runner launches it (round 2): dodt_integral_image_2d_banded +
dodt_anchor_filter_fused against the oracle (avod/core/anchor_filter.py:64-119 restated) and against
the unfused library calls they replace (integral image, filter, compaction, gathers, RPN decode) —
bit for bit, through the C ABI, at config B's and config E's grid sizes and on ragged sizes."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _occ(rng, nx, nz, density):
    return (rng.random((nx, nz)) < density).astype(np.uint8)


@pytest.mark.parametrize("nx,nz,density", [(800, 700, 0.005), (1600, 1400, 0.002), (37, 53, 0.2), (16, 700, 0.01),
                                           (17, 5, 0.5), (800, 700, 0.0)])
def test_banded_integral_image_equals_full(nx, nz, density):
    from dodt_b200 import ops
    from oracle import np_oracle as O
    rng = np.random.default_rng(nx * 7 + nz)
    occ = _occ(rng, nx, nz, density)
    t_occ = torch.from_numpy(occ).cuda()
    ws = torch.zeros(max(ops.integral_banded_workspace_bytes(nx, nz), 256), dtype=torch.uint8, device="cuda")
    ii_local = torch.empty((nx + 1, nz + 1), dtype=torch.int32, device="cuda")
    for _ in range(2):          # the second call checks that the first one re-armed the workspace
        ii_local.fill_(-1)
        bandoff, band_rows = ops.integral_image_2d_banded(t_occ, ii_local, ws)
        torch.cuda.synchronize()
        full = ii_local.cpu().numpy().astype(np.int64)
        off = bandoff.cpu().numpy()
        rows = (np.arange(1, nx + 1) - 1) // band_rows
        full[1:, 1:] += off[rows]
        want = O.integral_image_2d(occ.astype(np.float64))
        np.testing.assert_array_equal(full, want.astype(np.int64))
    np.testing.assert_array_equal(ops.integral_image_2d(t_occ).cpu().numpy(), want.astype(np.int32))


@pytest.mark.parametrize("n_anchors,seed", [(89600, 0), (89600, 1), (1, 2), (1023, 3), (1025, 4), (5000, 5)])
def test_fused_filter_equals_unfused_chain_and_oracle(n_anchors, seed):
    from dodt_b200 import ops
    from oracle import synth_ref as synth
    from oracle import np_oracle as O
    rng = np.random.default_rng(100 + seed)
    nx, nz, voxel = 800, 700, synth.VOXEL_SIZE
    occ = _occ(rng, nx, nz, [0.004, 0.05, 0.3, 0.0005, 0.01, 0.0][seed])
    anchors = synth.car_anchors()
    if n_anchors != len(anchors):
        anchors = anchors[rng.choice(len(anchors), n_anchors, replace=False)]
    thr = 1 if seed != 1 else 3
    dev = "cuda"
    t_occ = torch.from_numpy(occ).cuda()
    t_anchors = torch.from_numpy(anchors).cuda()
    n = len(anchors)
    bev_boxes = ops.project_to_bev(t_anchors, [-40, 40, 0, 70], tf_order=True)
    img_boxes = torch.from_numpy(rng.random((n, 4)).astype(np.float32)).cuda()
    scores = torch.from_numpy(rng.random(n).astype(np.float32)).cuda()
    offsets = torch.from_numpy(rng.normal(0, 0.1, (n, 6)).astype(np.float32)).cuda()
    e = lambda *s, dtype=torch.float32: torch.full(s, -5, dtype=dtype, device=dev)
    keep, kept_idx, n_kept = e(n, dtype=torch.uint8), e(n, dtype=torch.int32), e(1, dtype=torch.int32)
    k_bev, k_img, k_sc, k_rpn = e(n, 4), e(n, 4), e(n), e(n, 4)
    ws_ii = torch.zeros(max(ops.integral_banded_workspace_bytes(nx, nz), 256), dtype=torch.uint8, device=dev)
    ws = ops.anchor_filter_fused_workspace(n, dev)
    ii_local = torch.empty((nx + 1, nz + 1), dtype=torch.int32, device=dev)
    min_x, min_z = -400, 0
    for rep in range(2):        # twice: the workspace must be left re-armed
        bandoff, band_rows = ops.integral_image_2d_banded(t_occ, ii_local, ws_ii)
        ops.anchor_filter_fused(t_anchors, ii_local, nx, nz, min_x, min_z, voxel, thr, keep, kept_idx, n_kept, ws,
                                bandoff=bandoff, band_rows=band_rows, anchor_bev_boxes=bev_boxes, k_bev_boxes=k_bev,
                                anchor_img_boxes=img_boxes, k_img_boxes=k_img, rpn_scores=scores, k_scores=k_sc,
                                rpn_offsets=offsets, bev_extents=[-40, 40, 0, 70], k_rpn_boxes=k_rpn)
        torch.cuda.synchronize()
        assert not ws.any(), "the fused kernel must leave its workspace zero"
        want = O.empty_anchor_filter_2d(anchors, occ, voxel, np.array([min_x, min_z]), thr)
        np.testing.assert_array_equal(keep.cpu().numpy().astype(bool), want)
        idx = np.flatnonzero(want)
        assert int(n_kept.item()) == len(idx)
        np.testing.assert_array_equal(kept_idx[:len(idx)].cpu().numpy(), idx)
        np.testing.assert_array_equal(k_bev[:len(idx)].cpu().numpy(), bev_boxes.cpu().numpy()[idx])
        np.testing.assert_array_equal(k_img[:len(idx)].cpu().numpy(), img_boxes.cpu().numpy()[idx])
        np.testing.assert_array_equal(k_sc[:len(idx)].cpu().numpy(), scores.cpu().numpy()[idx])
    # the unfused library chain on the full integral image gives the same bits
    ii = ops.integral_image_2d(t_occ)
    keep2 = ops.anchor_filter_2d(t_anchors, ii, nx, nz, min_x, min_z, voxel, thr)
    idx2, cnt2 = ops.compact_mask(keep2)
    assert torch.equal(keep2, keep) and int(cnt2.item()) == len(idx)
    want_rpn = torch.empty((n, 4), dtype=torch.float32, device=dev)
    ops.rpn_decode(t_anchors, offsets, idx2, cnt2, [-40, 40, 0, 70], synth.A.KITTI_P2.reshape(-1), synth.IMAGE_SHAPE,
                   want_rpn, None)
    assert torch.equal(k_rpn[:len(idx)], want_rpn[:len(idx)])
    # the float32 tf.Tensor-branch decode (the frame runner's default) through both entry points
    ops.rpn_decode(t_anchors, offsets, idx2, cnt2, [-40, 40, 0, 70], synth.A.KITTI_P2.reshape(-1), synth.IMAGE_SHAPE,
                   want_rpn, None, tf_float32=True)
    ops.anchor_filter_fused(t_anchors, ii_local, nx, nz, min_x, min_z, voxel, thr, keep, kept_idx, n_kept, ws,
                            bandoff=bandoff, band_rows=band_rows, rpn_offsets=offsets, bev_extents=[-40, 40, 0, 70],
                            k_rpn_boxes=k_rpn, tf_float32=True)
    assert torch.equal(k_rpn[:len(idx)], want_rpn[:len(idx)])
    if n_anchors == 89600:
        # the anchors as a function of the index (no table read): same keep mask, order and decoded boxes
        keep_g, kept_g, n_g, k_g = torch.empty_like(keep), torch.empty_like(kept_idx), torch.empty_like(n_kept), \
            torch.empty_like(k_rpn)
        for f32 in (True, False):
            ops.anchor_filter_fused(None, ii_local, nx, nz, min_x, min_z, voxel, thr, keep_g, kept_g, n_g, ws,
                                    bandoff=bandoff, band_rows=band_rows, rpn_offsets=offsets,
                                    bev_extents=[-40, 40, 0, 70], k_rpn_boxes=k_g, tf_float32=f32,
                                    grid=(synth.AREA_EXTENTS, synth.A.CAR_ANCHOR_SIZES, synth.ANCHOR_STRIDE,
                                          synth.GROUND_PLANE))
            ops.anchor_filter_fused(t_anchors, ii_local, nx, nz, min_x, min_z, voxel, thr, keep, kept_idx, n_kept, ws,
                                    bandoff=bandoff, band_rows=band_rows, rpn_offsets=offsets,
                                    bev_extents=[-40, 40, 0, 70], k_rpn_boxes=k_rpn, tf_float32=f32)
            assert torch.equal(keep_g, keep) and torch.equal(n_g, n_kept)
            assert torch.equal(kept_g[:len(idx)], kept_idx[:len(idx)]) and torch.equal(k_g[:len(idx)], k_rpn[:len(idx)])
        with pytest.raises(Exception):     # a grid of another size than n
            ops.anchor_filter_fused(None, ii_local, nx, nz, min_x, min_z, voxel, thr, keep_g[:1000], kept_g, n_g, ws,
                                    bandoff=bandoff, band_rows=band_rows,
                                    grid=(synth.AREA_EXTENTS, synth.A.CAR_ANCHOR_SIZES, synth.ANCHOR_STRIDE,
                                          synth.GROUND_PLANE))
        ops.anchor_filter_fused(t_anchors, ii_local, nx, nz, min_x, min_z, voxel, thr, keep, kept_idx, n_kept, ws,
                                bandoff=bandoff, band_rows=band_rows, rpn_offsets=offsets, bev_extents=[-40, 40, 0, 70],
                                k_rpn_boxes=k_rpn, tf_float32=True)
    reg = synth.A.offset_to_anchor_tf32(anchors[idx], offsets.cpu().numpy()[idx])
    np.testing.assert_array_equal(k_rpn[:len(idx)].cpu().numpy(),
                                  synth.A.reorder_projected_boxes(synth.A.project_to_bev_tf32(reg, synth.BEV_EXTENTS)[1]))
    # ... and with the FULL image handed to the fused kernel (bandoff = None)
    keep.fill_(9)
    ops.anchor_filter_fused(t_anchors, ii, nx, nz, min_x, min_z, voxel, thr, keep, kept_idx, n_kept, ws)
    assert torch.equal(keep2, keep) and int(n_kept.item()) == len(idx)
