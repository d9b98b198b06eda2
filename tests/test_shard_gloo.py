"""The N>1 path on CPU: two gloo ranks shard sequences (and one long stream) as dodt_b200.shard
plans them, each runs its frames independently (here through the CPU oracle's NMS on small seeded
inputs, standing in for the GPU front end), and ONE all_gather of the fixed-size detection blocks
gives every rank the full, duplicate-free set — identical to a single-process run.
"""
import os
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dodt_b200 import shard


def test_assign_and_split():
    assert shard.assign_sequences(8, 8, 3) == [3]
    assert shard.assign_sequences(8, 2, 1) == [1, 3, 5, 7]
    assert shard.assign_sequences(3, 4, 3) == []
    with pytest.raises(ValueError):
        shard.assign_sequences(8, 2, 2)
    for n, w in ((300, 8), (7, 4), (2, 4), (0, 2), (1235, 3)):
        chunks = [shard.split_stream(n, w, r) for r in range(w)]
        assert chunks[0][0] == 0 and chunks[-1][1] == n
        for (f0, e0, _), (f1, e1, h1) in zip(chunks, chunks[1:]):
            assert e0 == f1
            assert h1 == (f1 - 1 if (f1 > 0 and e1 > f1) else None)
        sizes = [e - f for f, e, _ in chunks]
        assert max(sizes) - min(sizes) <= 1
        assert chunks[0][2] is None
    lens = [300, 297, 310, 154, 233, 447, 270, 800]
    frames = sum(shard.plan(lens, 8, r).n_frames for r in range(8))
    assert frames == sum(lens)
    assert sum(shard.plan(lens, 2, r).n_frames for r in range(2)) == sum(lens)
    p = shard.plan([10], 4, 2)
    assert p.items == [(0, 6, 8, 5)] and p.n_frames == 2


def _frame_detections(seq, frame):
    """Deterministic stand-in for one frame of the front end: seeded boxes -> oracle NMS."""
    from oracle import np_oracle as O
    rng = np.random.default_rng(1000 * seq + frame)
    n = int(rng.integers(5, 60))
    c = rng.uniform(0.2, 0.8, (n, 2))
    h = rng.uniform(0.02, 0.1, (n, 2))
    boxes = np.concatenate([c - h, c + h], 1).astype(np.float32)
    scores = rng.permutation(np.linspace(0.1, 0.9, n)).astype(np.float32)
    keep = O.non_max_suppression(boxes, scores, 16, 0.3)
    return boxes[keep], scores[keep], keep


def _run_shard(rank, world, lens, max_det=16):
    p = shard.plan(lens, world, rank)
    max_frames = max(shard.plan(lens, world, r).n_frames for r in range(world))
    block = shard.DetectionBlock(max_frames, max_det, "cpu")
    for seq, first, end, _halo in p.items:
        for f in range(first, end):
            b, s, k = _frame_detections(seq, f)
            block.append(seq, f, torch.from_numpy(b), torch.from_numpy(s), torch.from_numpy(k),
                         torch.tensor([len(k)], dtype=torch.int32))
    return shard.gather_detections(block)


def _worker(rank, world, init_file, lens, out_dir):
    dist.init_process_group("gloo", init_method="file://" + init_file, rank=rank, world_size=world)
    try:
        got = _run_shard(rank, world, lens)
        torch.save({k: v for k, v in got.items()}, os.path.join(out_dir, "rank%d.pt" % rank))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("lens", [[12, 9, 14, 7], [23]])
def test_two_rank_gather_equals_single_process(lens):
    want = _run_shard(0, 1, lens)
    assert len(want) == sum(lens)
    with tempfile.TemporaryDirectory() as d:
        init_file = os.path.join(d, "rdzv")
        mp.spawn(_worker, args=(2, init_file, lens, d), nprocs=2, join=True)
        for r in range(2):
            got = torch.load(os.path.join(d, "rank%d.pt" % r))
            assert sorted(got) == sorted(want)
            for k in want:
                assert torch.equal(got[k], want[k]), k
    for (seq, frame), rows in want.items():
        b, s, k = _frame_detections(seq, frame)
        np.testing.assert_array_equal(rows[:, :4].numpy(), b)
        np.testing.assert_array_equal(rows[:, 4].numpy(), s)
        np.testing.assert_array_equal(rows[:, 5].numpy().astype(np.int64), k)


def test_detection_block_overflow_and_duplicates():
    blk = shard.DetectionBlock(1, 4, "cpu")
    z4, z = torch.zeros(2, 4), torch.zeros(2)
    blk.append(0, 0, z4, z, z, 2)
    with pytest.raises(MemoryError):
        blk.append(0, 1, z4, z, z, 2)
    blk2 = shard.DetectionBlock(2, 4, "cpu")
    blk2.append(0, 0, z4, z, z, 2)
    blk2.append(0, 0, z4, z, z, 2)
    with pytest.raises(RuntimeError):
        shard.gather_detections(blk2)
