"""S4 over consecutive frames of a stream: k pairs (f0,f1), (f1,f2), ... as k launches vs one
batch-k launch on overlapping views of a contiguous [k+1,H,W,C] ring (shared maps are then fetched
from HBM once; tile order is batch-interleaved). usage: python tools/time_corr_pairs.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dodt_b200 import ops
torch.manual_seed(0)
H, W, C = 700, 800, 32
for k in (1, 2, 3, 4, 6):
    rings = [torch.rand(k + 1, H, W, C, device="cuda") for _ in range(3)]
    outs = [torch.empty(k, H, W, 25, device="cuda") for _ in range(3)]
    if k == 2:
        ref = outs[0].clone()
        for j in range(k):
            ops.correlation(rings[0][j:j + 1], rings[0][j + 1:j + 2], 1, 5, 1, 2, 5, out=ref[j:j + 1])
        ops.correlation(rings[0][0:k], rings[0][1:k + 1], 1, 5, 1, 2, 5, out=outs[0])
        torch.cuda.synchronize()
        print("batched == separate:", bool(torch.equal(ref, outs[0])))
    for ctas in (0, 148):
        for i in range(3):
            ops.correlation(rings[i][0:k], rings[i][1:k + 1], 1, 5, 1, 2, 5, out=outs[i], max_ctas=ctas)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 24
        e0.record()
        for i in range(reps):
            ops.correlation(rings[i % 3][0:k], rings[i % 3][1:k + 1], 1, 5, 1, 2, 5, out=outs[i % 3], max_ctas=ctas)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps / k
        print("k=%d ctas=%3d  %.1f us per pair  %.0f GB/s algorithmic" % (k, ctas, us, 199.36e6 / us / 1e3))
    del rings, outs
