"""Frame-stream correlation (dodt_correlation_stream) and the k-frame group of the frame runner:
same bits as the pairwise / per-frame calls."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dd():
    import dodt_b200
    return dodt_b200


@pytest.mark.parametrize("n_maps,shape,kw", [
    (3, (1, 40, 72, 32), dict(kernel_size=1, max_displacement=5, stride_1=1, stride_2=2, padding=5)),
    (5, (1, 33, 70, 16), dict(kernel_size=1, max_displacement=5, stride_1=1, stride_2=2, padding=5)),
    (11, (1, 36, 64, 8), dict(kernel_size=1, max_displacement=2, stride_1=1, stride_2=2, padding=2)),  # > 8 pairs
    (4, (1, 16, 18, 6), dict(kernel_size=3, max_displacement=4, stride_1=2, stride_2=2, padding=4)),   # generic path
    (2, (1, 40, 72, 32), dict(kernel_size=1, max_displacement=5, stride_1=1, stride_2=2, padding=5)),
])
def test_stream_equals_pairwise_and_oracle(dd, n_maps, shape, kw):
    from oracle import c_oracle as CO
    rng = np.random.default_rng(n_maps * 13 + shape[3])
    maps = [np.abs(rng.standard_normal(shape)).astype(np.float32) for _ in range(n_maps)]
    got = dd.correlation_stream(maps, **kw)
    assert len(got) == n_maps - 1
    for j in range(n_maps - 1):
        np.testing.assert_array_equal(got[j], dd.correlation(maps[j], maps[j + 1], **kw))
    np.testing.assert_allclose(got[0], CO.correlation(maps[0], maps[1], **kw), rtol=1e-5, atol=1e-7)


def test_stream_errors(dd):
    m = np.zeros((1, 8, 8, 8), np.float32)
    with pytest.raises(ValueError):
        dd.correlation_stream([m])
    with pytest.raises(ValueError):
        dd.correlation_stream([m, np.zeros((1, 8, 9, 8), np.float32)])
    with pytest.raises(ValueError):
        dd.correlation_stream([m, m], kernel_size=2)


def test_group_of_frames_equals_frame_by_frame():
    """FrontEnd.enqueue_group (one S4 launch for k frames, per-frame chains on branch streams,
    eager and as a CUDA graph) leaves every slot with the results of k enqueue() calls."""
    from oracle import synth_ref as synth
    from dodt_b200.frontend import FrontEnd, HostFrame
    fe = FrontEnd()
    k = 3
    ref_slots = [fe.new_slot() for _ in range(k + 1)]
    grp_slots = [fe.new_slot() for _ in range(k + 1)]
    for i in range(k + 1):
        h = HostFrame(fe).fill(synth.frame_inputs(2, 40 + i))
        h.upload(ref_slots[i])
        h.upload(grp_slots[i])
    for s in ref_slots[1:]:
        s.result_buf.zero_()     # layout padding and unused stats words are never written
    for i in range(1, k + 1):
        fe.enqueue(ref_slots[i], ref_slots[i - 1])
    torch.cuda.synchronize()
    graph, launches = fe.capture_group(grp_slots[1:], grp_slots[0])
    for s in grp_slots[1:]:
        s.result_buf.zero_()
        s.corr.zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert launches > 0
    for r, g in zip(ref_slots[1:], grp_slots[1:]):
        assert torch.equal(r.result_buf, g.result_buf)
        assert torch.equal(r.corr, g.corr)
        n_top = int(r.n_top[0])
        assert torch.equal(r.corr_rois[:n_top], g.corr_rois[:n_top])
        assert torch.equal(r.bev_rois[:n_top], g.bev_rois[:n_top])
