"""Throughput of the multi-stream frame pipeline with one stage left out at a time: the marginal
cost of each stage in frames/s terms (device-resident inputs, as bench.py's `value`).
usage: python tools/stage_ablation.py [slots] [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if os.environ.get("ABL_DIAG"):       # the diagnostic library: DODT_* knobs read the environment
    from dodt_b200 import _lib
    _lib.use_diag_library()
from dodt_b200 import synth  # noqa: E402
from dodt_b200.frontend import FrontEnd, HostFrame  # noqa: E402

n_slots = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
from dodt_b200.frontend import FrontEndConfig  # noqa: E402
cfg = FrontEndConfig()
if os.environ.get("ABL_CORR_CTAS"):
    cfg.corr_max_ctas = int(os.environ["ABL_CORR_CTAS"])
if os.environ.get("ABL_NMS_WINDOWS"):
    cfg.nms_max_windows = int(os.environ["ABL_NMS_WINDOWS"])
if os.environ.get("ABL_PRIO"):        # "chain,corr" stream priorities, e.g. -1,0
    cfg.chain_stream_priority, cfg.corr_stream_priority = (int(v) for v in os.environ["ABL_PRIO"].split(","))
fe = FrontEnd(cfg)
slots = [fe.new_slot() for _ in range(n_slots)]
for i, s in enumerate(slots):
    HostFrame(fe).fill(synth.frame_inputs(2, i)).upload(s)
torch.cuda.synchronize()
streams = [torch.cuda.Stream() for _ in range(n_slots)]
main = torch.cuda.current_stream()


GROUP = int(os.environ.get("ABL_GROUP", "1"))     # consecutive frames per graph (one S4 launch)
assert n_slots % GROUP == 0 and steps % GROUP == 0


def run(skip):
    n_groups = n_slots // GROUP
    graphs = [fe.capture_group(slots[g * GROUP:(g + 1) * GROUP], slots[g * GROUP - 1], None, skip)[0]
              for g in range(n_groups)]

    def rr(n):
        for st in streams:
            st.wait_stream(main)
        for i in range(n // GROUP):
            with torch.cuda.stream(streams[i % n_groups]):
                graphs[i % n_groups].replay()
        for st in streams:
            main.wait_stream(st)
    rr(3 * n_slots)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    rr(steps)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / steps


base = run(())
print("all stages: %.1f us/frame (%.0f frames/s)" % (base, 1e6 / base), flush=True)
if len(sys.argv) > 3 and sys.argv[3] == "base":
    sys.exit(0)
for st in ("S1", "S2", "S3a", "S5a", "S4", "S3b", "S5b"):
    t = run((st,))
    print("without %-3s: %6.1f us/frame  (stage costs %5.1f us of throughput)" % (st, t, base - t))
t = run(("S1", "S2", "S3a", "S5a", "S3b", "S5b"))
print("only S4    : %6.1f us/frame" % t)
t = run(("S4",))
