"""Times dodt_correlation at the config-C size (DODT_CORR_IMPL / DODT_CORR_PREFETCH select variants).
usage: python tools/time_corr.py [max_ctas]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dodt_b200 import ops
    from oracle import synth_ref as synth
torch.manual_seed(0)
max_ctas = int(sys.argv[1]) if len(sys.argv) > 1 else 0
bufs = [(torch.rand(1, 700, 800, 32, device="cuda"), torch.rand(1, 700, 800, 32, device="cuda"),
         torch.empty(1, 700, 800, 25, device="cuda")) for _ in range(3)]
for a, b, o in bufs:
    ops.correlation(a, b, 1, 5, 1, 2, 5, out=o, max_ctas=max_ctas)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 60
e0.record()
for i in range(reps):
    a, b, o = bufs[i % 3]
    ops.correlation(a, b, 1, 5, 1, 2, 5, out=o, max_ctas=max_ctas)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / reps
print("impl=%s prefetch=%s ctas=%d  %.1f us  %.0f GB/s (algorithmic 199.36 MB)" % (
    os.environ.get("DODT_CORR_IMPL", "default"), os.environ.get("DODT_CORR_PREFETCH", "0"), max_ctas, us, 199.36e6 / us / 1e3))
