"""Multi-GPU partitioning of the front end: one process per GPU, whole KITTI-tracking sequences (or
contiguous chunks of one long stream) per rank, NO data-path collective; the only communication is
one all_gather of the per-frame detection lists at the end of a shard.

Reference precedent: the index split over forked workers of
scripts/preprocessing/gen_tracking_mini_batches.py:22-69 (split_indices / os.fork) — frames are
independent except that the correlation needs the previous frame of the SAME sequence
(avod/core/models/dt_rpn_model.py:328-330; pairs never cross videos,
avod/datasets/kitti/kitti_tracking_dataset.py:480).

Works with any torch.distributed backend: NCCL over NVLink on the GPU box (CUDA tensors), gloo in
the CPU tests (host tensors). Nothing here computes; it is bookkeeping around the kernels.
"""
from dataclasses import dataclass

import torch
import torch.distributed as dist

# columns of one gathered detection row: the normalised BEV box the NMS ran on, its score and the
# index of the proposal it came from (the reference's per-frame text rows,
# avod/core/dt_evaluator.py:1098-1147, are built from these by the network heads that are out of scope)
DET_COLS = 6


def assign_sequences(n_sequences, world_size, rank):
    """Round-robin whole sequences: sequence s runs on rank s mod world_size."""
    if not 0 <= rank < world_size:
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    return list(range(rank, n_sequences, world_size))


def split_stream(n_frames, world_size, rank):
    """One long stream: contiguous chunks, sizes differing by at most one. Returns (first, last+1,
    halo) where halo is the frame before `first` whose BEV features must be recomputed on this rank
    to correlate with `first` (None for the first chunk)."""
    if not 0 <= rank < world_size:
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    base, extra = divmod(n_frames, world_size)
    first = rank * base + min(rank, extra)
    count = base + (1 if rank < extra else 0)
    halo = first - 1 if (first > 0 and count > 0) else None
    return first, first + count, halo


@dataclass
class ShardPlan:
    """What one rank runs: a list of (sequence id, first frame, end frame, halo frame)."""
    rank: int
    world_size: int
    items: list

    @property
    def n_frames(self):
        return sum(e - f for _, f, e, _ in self.items)


def plan(sequence_lengths, world_size, rank):
    """Sequences are dealt round-robin while there are at least as many sequences as ranks;
    otherwise every sequence is cut into world_size chunks (with the one-frame halo)."""
    n = len(sequence_lengths)
    if n >= world_size:
        items = [(s, 0, int(sequence_lengths[s]), None) for s in assign_sequences(n, world_size, rank)]
    else:
        items = []
        for s, length in enumerate(sequence_lengths):
            first, end, halo = split_stream(int(length), world_size, rank)
            if end > first:
                items.append((s, first, end, halo))
    return ShardPlan(rank, world_size, items)


class DetectionBlock:
    """Fixed-size padded detection lists of one shard: rows [frames, max_det, DET_COLS] f32,
    counts [frames] i32, frame_ids [frames, 2] i32 (sequence, frame). Fixed shapes so that the
    gather is ONE collective of equal-size blocks, whatever each frame detected: the three arrays
    are views of ONE contiguous 4-byte-word buffer (`buf`), which is what crosses NVLink."""

    def __init__(self, max_frames, max_det, device):
        self.max_frames, self.max_det = int(max_frames), int(max_det)
        n_rows = self.max_frames * self.max_det * DET_COLS
        self._n_rows, self._n_counts, self._n_ids = n_rows, self.max_frames, 2 * self.max_frames
        self.buf = torch.zeros((n_rows + self._n_counts + self._n_ids,), dtype=torch.float32, device=device)
        self.rows, self.counts, self.frame_ids = self.views(self.buf)
        self.frame_ids.fill_(-1)
        self.cursor = torch.zeros((1,), dtype=torch.int32, device=device)   # device-side row cursor
        self.n = 0
        self._gather_buf = None

    def views(self, buf):
        """(rows, counts, frame_ids) views of a buffer laid out like `buf`."""
        a, b = self._n_rows, self._n_rows + self._n_counts
        rows = buf[:a].view(self.max_frames, self.max_det, DET_COLS)
        counts = buf[a:b].view(torch.int32)
        ids = buf[b:b + self._n_ids].view(torch.int32).view(self.max_frames, 2)
        return rows, counts, ids

    def gather_buffer(self, world):
        """The receive buffer of the gather, allocated once (outside any timed region)."""
        if self._gather_buf is None or self._gather_buf.shape[0] != world:
            self._gather_buf = torch.empty((world, self.buf.numel()), dtype=self.buf.dtype,
                                           device=self.buf.device)
        return self._gather_buf

    def reset(self):
        """Forget every stored frame (device-side cursor included); no synchronisation."""
        self.frame_ids.fill_(-1)
        self.counts.zero_()
        self.cursor.zero_()
        self.n = 0

    def append(self, sequence, frame, boxes, scores, indices, count):
        """boxes [k,4], scores [k], indices [k] (tensors on the block's device), count int or 0-d /
        1-element tensor (no synchronisation when it is a tensor)."""
        i = self.n
        if i >= self.rows.shape[0]:
            raise MemoryError("detection block is full")
        k = min(boxes.shape[0], self.rows.shape[1])
        self.rows[i, :k, 0:4] = boxes[:k]
        self.rows[i, :k, 4] = scores[:k]
        self.rows[i, :k, 5] = indices[:k].to(torch.float32)
        if torch.is_tensor(count):
            self.counts[i:i + 1] = count.reshape(-1)[:1].to(torch.int32)
        else:
            self.counts[i] = int(count)
        self.frame_ids[i, 0] = int(sequence)
        self.frame_ids[i, 1] = int(frame)
        self.n += 1


def all_gather_blocks(block, group=None):
    """The ONE collective of a shard: a single all_gather_into_tensor of every rank's packed block
    (rows, counts and frame ids travel in one buffer; the receive buffer is preallocated by
    `block.gather_buffer`). Returns [(rows, counts, frame_ids)] per rank as views, on the block's
    device, without synchronising; with no process group (or a world of one) it is the local block."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
        return [(block.rows, block.counts, block.frame_ids)]
    world = dist.get_world_size(group)
    out = block.gather_buffer(world)
    dist.all_gather_into_tensor(out.view(-1), block.buf, group=group)
    return [block.views(out[r]) for r in range(world)]


def unpack_blocks(parts):
    """[(rows, counts, frame_ids)] -> {(sequence, frame): rows [count, DET_COLS]} (host tensors)."""
    out = {}
    for rows, counts, ids in parts:
        rows, counts, ids = rows.cpu(), counts.cpu(), ids.cpu()
        for i in range(ids.shape[0]):
            seq, frame = int(ids[i, 0]), int(ids[i, 1])
            if seq < 0:
                continue
            key = (seq, frame)
            if key in out:
                raise RuntimeError("frame %r was processed by two ranks" % (key,))
            out[key] = rows[i, :int(counts[i])].clone()
    return out


def gather_detections(block, group=None):
    """all_gather of every rank's DetectionBlock, unpacked: {(sequence, frame): rows} on every rank."""
    return unpack_blocks(all_gather_blocks(block, group))
