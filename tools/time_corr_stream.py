"""S4 timing: one pair and the k-pair frame-stream launch, CUDA events over graph-free back-to-back
launches on rotating buffers (inputs larger than L2). With --diag the diagnostic library is loaded,
so DODT_CORR_FEED=0 (round-1 cp.async kernel) can be compared with the default TMA-fed kernel.
usage: python tools/time_corr_stream.py [--diag] [--pairs 1,4,8]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if "--diag" in sys.argv:
    from dodt_b200 import _lib
    _lib.use_diag_library()
import torch  # noqa: E402

from dodt_b200 import ops  # noqa: E402

pairs = [1, 4, 8]
ctas = 0
for i, a in enumerate(sys.argv):
    if a == "--pairs":
        pairs = [int(v) for v in sys.argv[i + 1].split(",")]
    if a == "--ctas":
        ctas = int(sys.argv[i + 1])
torch.manual_seed(0)
H, W, C = 700, 800, 32
for k in pairs:
    n_sets = 3 if k >= 4 else 6
    maps = [[torch.rand(1, H, W, C, device="cuda") for _ in range(k + 1)] for _ in range(n_sets)]
    outs = [[torch.empty(1, H, W, 25, device="cuda") for _ in range(k)] for _ in range(n_sets)]

    def launch(i):
        if k == 1:
            ops.correlation(maps[i][0], maps[i][1], 1, 5, 1, 2, 5, out=outs[i][0], max_ctas=ctas)
        else:
            ops.correlation_stream(maps[i], 1, 5, 1, 2, 5, outs=outs[i], max_ctas=ctas)
    for i in range(n_sets):
        launch(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 30
    e0.record()
    for i in range(reps):
        launch(i % n_sets)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps / k
    print("pairs per launch %d: %.1f us per pair, %.0f GB/s algorithmic (%.3f of 6554)" %
          (k, us, 199.36e6 / us / 1e3, 199.36e6 / us / 1e3 / 6554))
    del maps, outs
