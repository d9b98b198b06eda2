"""Times dodt_correlation_grad at the config-C size [1,700,800,32] -> 25 displacements: both
gradients, and each alone (CUDA events, three rotating buffer sets = 1 GB > L2).
usage: python tools/time_corr_grad.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dodt_b200 import ops
torch.manual_seed(0)
N, H, W, C, D2 = 1, 700, 800, 32, 25
ws_bytes = ops.correlation_grad_workspace_bytes(N, H, W, C, 1, 5, 1, 2, 5)
bufs = [dict(a=torch.rand(N, H, W, C, device="cuda"), b=torch.rand(N, H, W, C, device="cuda"),
             g=torch.randn(N, H, W, D2, device="cuda"), ga=torch.empty(N, H, W, C, device="cuda"),
             gb=torch.empty(N, H, W, C, device="cuda"),
             ws=torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")) for _ in range(3)]
IN, G = N * H * W * C * 4, N * H * W * D2 * 4


def run(x, need_a, need_b):
    ops.correlation_grad(x["g"], x["a"], x["b"], 1, 5, 1, 2, 5, grad_a=x["ga"], grad_b=x["gb"],
                         workspace=x["ws"], need_a=need_a, need_b=need_b)


for name, na, nb, nbytes in (("both", True, True, G + 4 * IN), ("grad_a", True, False, G + 2 * IN),
                             ("grad_b", False, True, G + 2 * IN)):
    for x in bufs:
        run(x, na, nb)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 30
    e0.record()
    for i in range(reps):
        run(bufs[i % 3], na, nb)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    print("%-7s %.1f us  %.0f GB/s algorithmic (%.1f MB)" % (name, us, nbytes / us / 1e3, nbytes / 1e6))

# per-kernel durations (CUPTI through torch.profiler)
try:
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(6):
            run(bufs[i % 3], True, True)
        torch.cuda.synchronize()
    for ev in prof.key_averages():
        print("  %-60s n=%d  avg %.1f us" % (ev.key[:60], ev.count, ev.device_time_total / max(ev.count, 1)))
except Exception as exc:   # noqa: BLE001
    print("profiler unavailable:", exc)
