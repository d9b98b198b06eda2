"""End-to-end parity of the frame-stream runner (dodt_b200.frontend.FrontEnd: all stages of a
frame enqueued without a host round trip, eager and as a CUDA-graph replay) against the CPU
oracle's whole-frame restatement (oracle/cpu_frontend.py) on the same synthetic frame."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _upload(slot, inp):
    for k, dst in slot.input_tensors().items():
        src = torch.from_numpy(np.ascontiguousarray(inp[k]))
        if k == "points":
            dst[:, :src.shape[1]].copy_(src)
            slot.n_points = src.shape[1]
        else:
            dst.copy_(src)


def _reference_frame(fe, slot, inp, prev_bev_feat):
    """The CPU oracle's frame, with the RPN-decode boxes taken from the device after checking that
    they equal the host NumPy chain to float64 rounding noise (exp/log and np.dot are library
    defined; everything downstream is then compared exactly / to 1e-5)."""
    from oracle import cpu_frontend
    n_kept = int(slot.n_kept.item())
    kept = slot.kept_idx[:n_kept].cpu().numpy()
    k_boxes = slot.k_rpn_boxes[:n_kept].cpu().numpy()
    np.testing.assert_allclose(k_boxes, inp["rpn_boxes"][kept], rtol=1e-6, atol=1e-7)
    assert (k_boxes == inp["rpn_boxes"][kept]).mean() > 0.98

    def prop_img(top):   # image boxes exist on the device for the NMS survivors only
        got = slot.prop_img_boxes[:len(top)].cpu().numpy()
        np.testing.assert_allclose(got, inp["rpn_img_boxes"][kept][top], rtol=1e-6, atol=1e-7)
        return got
    a_img = fe.anchor_img_boxes.cpu().numpy()
    np.testing.assert_allclose(a_img, cpu_frontend.anchors()[2], rtol=1e-6, atol=1e-7)
    np.testing.assert_array_equal(fe.anchors.cpu().numpy(), cpu_frontend.anchors()[0])
    np.testing.assert_array_equal(fe.anchor_bev_boxes.cpu().numpy(), cpu_frontend.anchors()[1])
    return cpu_frontend.run_frame(inp, prev_bev_feat, k_boxes=k_boxes, prop_img_boxes=prop_img,
                                  anchor_img_boxes=a_img)


def _check(slot, ref):
    n_kept = int(slot.n_kept.item())
    assert n_kept == len(ref["kept"])
    np.testing.assert_array_equal(slot.keep.cpu().numpy().astype(bool), ref["keep"])
    np.testing.assert_array_equal(slot.kept_idx[:n_kept].cpu().numpy(), ref["kept"])
    maps = slot.maps.cpu().numpy()
    for i in range(5):
        np.testing.assert_array_equal(maps[i], ref["bev"]["height_maps"][i].astype(np.float32))
    np.testing.assert_array_equal(maps[5], ref["bev"]["density_map"].astype(np.float32))
    np.testing.assert_array_equal(slot.occ.cpu().numpy(), ref["occ"])
    np.testing.assert_allclose(slot.rpn_bev_crops[:n_kept].cpu().numpy(), ref["rpn_bev_crops"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(slot.rpn_img_crops[:n_kept].cpu().numpy(), ref["rpn_img_crops"], rtol=1e-5, atol=1e-6)
    n_top, complete = slot.n_top.cpu().tolist()
    assert complete == 1 and n_top == len(ref["top"])
    np.testing.assert_array_equal(slot.top_idx[:n_top].cpu().numpy(), ref["top"])
    np.testing.assert_allclose(slot.corr.cpu().numpy(), ref["corr"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(slot.bev_rois[:n_top].cpu().numpy(), ref["bev_rois"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(slot.img_rois[:n_top].cpu().numpy(), ref["img_rois"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(slot.corr_rois[:n_top].cpu().numpy(), ref["corr_rois"], rtol=2e-5, atol=1e-6)
    n_final, complete = slot.n_final.cpu().tolist()
    assert complete == 1 and n_final == len(ref["final"])
    np.testing.assert_array_equal(slot.final_idx[:n_final].cpu().numpy(), ref["final"])


def test_frontend_frame_eager_and_graph(lib):
    from oracle import synth_ref as synth
    from dodt_b200.frontend import FrontEnd
    from oracle import cpu_frontend
    fe = FrontEnd()
    slots = [fe.new_slot(), fe.new_slot()]
    inputs = [synth.frame_inputs(2, 40), synth.frame_inputs(2, 41)]
    for s, inp in zip(slots, inputs):
        _upload(s, inp)
    launches = fe.enqueue(slots[1], slots[0])
    torch.cuda.synchronize()
    assert launches > 10
    ref = _reference_frame(fe, slots[1], inputs[1], inputs[0]["bev_feat"])
    _check(slots[1], ref)
    # the same frame as a CUDA-graph replay, after scribbling over the outputs
    graph, n = fe.capture(slots[1], slots[0])
    for t in (slots[1].maps, slots[1].corr, slots[1].bev_rois, slots[1].rpn_bev_crops):
        t.fill_(-1.0)
    slots[1].top_idx.fill_(-7)
    graph.replay()
    torch.cuda.synchronize()
    _check(slots[1], ref)
    # replay on new data without re-capturing
    _upload(slots[1], synth.frame_inputs(2, 42))
    graph.replay()
    torch.cuda.synchronize()
    ref2 = _reference_frame(fe, slots[1], synth.frame_inputs(2, 42), inputs[0]["bev_feat"])
    _check(slots[1], ref2)


def test_compact_and_gather(lib):
    from dodt_b200 import ops
    rng = np.random.default_rng(0)
    for n in (1, 5, 4095, 4096, 4097, 89600, 300001):
        keep = (rng.random(n) < 0.37).astype(np.uint8)
        idx, count = ops.compact_mask(torch.from_numpy(keep).cuda())
        want = np.flatnonzero(keep)
        assert int(count.item()) == len(want)
        np.testing.assert_array_equal(idx[:len(want)].cpu().numpy(), want)
        src = rng.standard_normal((n, 4)).astype(np.float32)
        got = ops.gather_rows(torch.from_numpy(src).cuda(), idx, count)
        np.testing.assert_array_equal(got[:len(want)].cpu().numpy(), src[want])
    idx, count = ops.compact_mask(torch.zeros(1000, dtype=torch.uint8, device="cuda"))
    assert int(count.item()) == 0


def test_nms_device_count_and_window_cap(lib):
    """n_dev limits the candidates; max_windows bounds the launches and reports completeness."""
    from dodt_b200 import ops
    from oracle import np_oracle as O
    rng = np.random.default_rng(1)
    n = 6000
    c = rng.uniform(0.1, 0.9, (n, 2))
    boxes = np.concatenate([c - 0.02, c + 0.02], 1).astype(np.float32)
    scores = rng.permutation(np.linspace(0, 1, n)).astype(np.float32)
    b, s = torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda()
    n_dev = torch.tensor([2500], dtype=torch.int32, device="cuda")
    keep, n_keep = ops.nms(b, s, 4000, 0.3, n_dev=n_dev)
    want = O.non_max_suppression(boxes[:2500], scores[:2500], 4000, 0.3)
    assert n_keep.cpu().tolist() == [len(want), 1]
    np.testing.assert_array_equal(keep[:len(want)].cpu().numpy(), want)
    keep, n_keep = ops.nms(b, s, 4000, 0.3, max_windows=1)      # 6000 candidates need 4 windows
    got_n, complete = n_keep.cpu().tolist()
    assert complete == 0
    want = O.non_max_suppression(boxes, scores, 4000, 0.3)
    np.testing.assert_array_equal(keep[:got_n].cpu().numpy(), want[:got_n])
    keep, n_keep = ops.nms(b, s, 50, 0.3, max_windows=1)        # 50 found inside the first window
    assert n_keep.cpu().tolist() == [50, 1]
    np.testing.assert_array_equal(keep.cpu().numpy(), want[:50])


def test_host_frame_and_detection_block(lib):
    """HostFrame (packed pinned inputs/results) drives a graph-captured frame whose final detections
    are appended to a shard DetectionBlock by dodt_emit_detections; the unpacked block equals the
    oracle's final NMS selection (box, score, index), frame after frame, and overflow is dropped."""
    from dodt_b200 import shard
    from oracle import synth_ref as synth
    from dodt_b200.frontend import FrontEnd, HostFrame
    from oracle import cpu_frontend
    fe = FrontEnd()
    slots = [fe.new_slot(), fe.new_slot()]
    inputs = [synth.frame_inputs(2, 50), synth.frame_inputs(2, 51)]
    hosts = [HostFrame(fe).fill(inp, sequence=3, frame=50 + i) for i, inp in enumerate(inputs)]
    for h, s in zip(hosts, slots):
        h.upload(s)
    block = shard.DetectionBlock(3, fe.cfg.avod_nms_size, fe.device)
    graph, _ = fe.capture(slots[1], slots[0], block)
    block.reset()
    graph.replay()
    hosts[1].download(slots[1])
    torch.cuda.synchronize()
    ref = _reference_frame(fe, slots[1], inputs[1], inputs[0]["bev_feat"])
    _check(slots[1], ref)
    got = shard.gather_detections(block)
    assert list(got) == [(3, 51)]
    rows = got[(3, 51)].numpy()
    kept = ref["kept"]
    prop = slots[1].k_rpn_boxes[:len(kept)].cpu().numpy()[ref["top"]]
    np.testing.assert_array_equal(rows[:, 5].astype(np.int64), ref["final"])
    np.testing.assert_array_equal(rows[:, :4], prop[ref["final"]])
    np.testing.assert_array_equal(rows[:, 4], inputs[1]["final_scores"][ref["final"]])
    # packed results on the host
    res = hosts[1].results
    assert res["n_final"].tolist() == [len(ref["final"]), 1]
    np.testing.assert_array_equal(res["final_idx"][:len(ref["final"])].numpy(), ref["final"])
    assert int(res["n_kept"][0]) == len(kept)
    # a second frame through the same graph lands in the next row; capacity overflow is dropped
    hosts[1].fill(inputs[1], sequence=3, frame=52).upload(slots[1])
    for _ in range(4):
        graph.replay()
    torch.cuda.synchronize()
    assert int(block.cursor.item()) == 5
    ids = block.frame_ids.cpu().numpy()
    assert ids.tolist() == [[3, 51], [3, 52], [3, 52]]


def test_frontend_resumes_incomplete_rpn_nms(lib):
    """Clustered RPN outputs (many anchors regressed onto the same objects, the best scores among
    them): at IoU 0.8 the 1024 proposals are NOT found inside the windows the frame graph reserves
    for the RPN NMS, the graph reports n_top[1] == 0, and FrontEnd.complete_frame resumes the
    selection and redoes what follows it. The reference always scans to completion
    (tf.image.non_max_suppression, dt_rpn_model.py:587-591): the finished frame must equal the
    oracle's, and its row of the detection block must have been replaced."""
    from dodt_b200 import shard
    from oracle import synth_ref as synth
    from dodt_b200.frontend import FrontEnd, HostFrame
    fe = FrontEnd()
    slots = [fe.new_slot(), fe.new_slot()]
    inputs = [synth.frame_inputs(2, 60), synth.frame_inputs(2, 61)]
    inputs[1].update(synth.clustered_rpn_outputs(2, 61, n_targets=60))
    hosts = [HostFrame(fe).fill(inp, sequence=5, frame=60 + i) for i, inp in enumerate(inputs)]
    for h, s in zip(hosts, slots):
        h.upload(s)
    block = shard.DetectionBlock(4, fe.cfg.avod_nms_size, fe.device)
    graph, _ = fe.capture(slots[1], slots[0], block)
    block.reset()
    graph.replay()
    hosts[1].download(slots[1])
    torch.cuda.synchronize()
    n_top, complete = slots[1].n_top.cpu().tolist()
    assert complete == 0 and n_top < fe.cfg.rpn_nms_size, "the scene must not fit the reserved windows"
    assert not fe.rpn_nms_complete(hosts[1]) and not fe.rpn_nms_complete(slots[1])
    truncated = shard.gather_detections(block)[(5, 61)].clone()
    assert fe.complete_frame(slots[1], block) is True
    torch.cuda.synchronize()
    assert fe.rpn_nms_complete(slots[1])
    assert fe.complete_frame(slots[1], block) is False          # nothing left to do
    ref = _reference_frame(fe, slots[1], inputs[1], inputs[0]["bev_feat"])
    assert len(ref["top"]) == fe.cfg.rpn_nms_size
    _check(slots[1], ref)
    # the block still holds ONE row for the frame, now with the complete frame's detections
    assert int(block.cursor.item()) == 1
    got = shard.gather_detections(block)
    assert list(got) == [(5, 61)]
    rows = got[(5, 61)].numpy()
    np.testing.assert_array_equal(rows[:, 5].astype(np.int64), ref["final"])
    prop = slots[1].k_rpn_boxes[:len(ref["kept"])].cpu().numpy()[ref["top"]]
    np.testing.assert_array_equal(rows[:, :4], prop[ref["final"]])
    assert truncated.shape != rows.shape or not np.array_equal(truncated.numpy(), rows)
    # an ordinary frame through the same graph afterwards is complete without help
    hosts[1].fill(synth.frame_inputs(2, 62), sequence=5, frame=62).upload(slots[1])
    graph.replay()
    torch.cuda.synchronize()
    assert fe.rpn_nms_complete(slots[1])


def test_frontend_frame_realistic_occupancy(lib):
    """The same whole-frame parity on a cloud with a real KITTI frame's occupancy
    (synth.point_cloud_kitti: ~18.5 k points, ~12 k of the 89 600 anchors kept) — the regime the
    reference runs in, five times fewer kept anchors than the evenly spread benchmark cloud."""
    from oracle import synth_ref as synth
    from dodt_b200.frontend import FrontEnd
    fe = FrontEnd()
    slots = [fe.new_slot(), fe.new_slot()]
    inputs = [synth.frame_inputs(2, 70), synth.frame_inputs(2, 71)]
    for k, inp in enumerate(inputs):
        inp["points"] = synth.point_cloud_kitti(2, 70 + k)
    for s, inp in zip(slots, inputs):
        _upload(s, inp)
    graph, _ = fe.capture(slots[1], slots[0])
    graph.replay()
    torch.cuda.synchronize()
    n_kept = int(slots[1].n_kept.item())
    assert 8000 < n_kept < 16000
    assert fe.rpn_nms_complete(slots[1])
    ref = _reference_frame(fe, slots[1], inputs[1], inputs[0]["bev_feat"])
    _check(slots[1], ref)
