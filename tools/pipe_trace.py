"""Kernel timeline of the multi-stream frame pipeline (CUPTI via torch.profiler): writes
gpurun_out/pipe_trace.json with (name, stream, start us, duration us) of every kernel of ~4 rounds
over the resident slots, for offline analysis of what overlaps with what.
usage: python tools/pipe_trace.py [slots]"""
import json
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dodt_b200 import synth  # noqa: E402
from dodt_b200.frontend import FrontEnd, FrontEndConfig, HostFrame  # noqa: E402

n_slots = int(sys.argv[1]) if len(sys.argv) > 1 else 12
cfg = FrontEndConfig()
if os.environ.get("ABL_CORR_CTAS"):
    cfg.corr_max_ctas = int(os.environ["ABL_CORR_CTAS"])
fe = FrontEnd(cfg)
slots = [fe.new_slot() for _ in range(n_slots)]
for i, s in enumerate(slots):
    HostFrame(fe).fill(synth.frame_inputs(2, i)).upload(s)
torch.cuda.synchronize()
streams = [torch.cuda.Stream() for _ in range(n_slots)]
main = torch.cuda.current_stream()
graphs = [fe.capture(slots[i], slots[i - 1], None)[0] for i in range(n_slots)]


def rr(n):
    for st in streams:
        st.wait_stream(main)
    for i in range(n):
        with torch.cuda.stream(streams[i % n_slots]):
            graphs[i % n_slots].replay()
    for st in streams:
        main.wait_stream(st)


rr(10 * n_slots)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    rr(6 * n_slots)
    torch.cuda.synchronize()
os.makedirs("gpurun_out", exist_ok=True)
prof.export_chrome_trace("gpurun_out/pipe_trace_raw.json")
raw = json.load(open("gpurun_out/pipe_trace_raw.json"))
ev = [dict(name=e["name"][:80], stream=e.get("args", {}).get("stream"), ts=e["ts"], dur=e["dur"],
           grid=e.get("args", {}).get("grid"), block=e.get("args", {}).get("block"))
      for e in raw["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
json.dump(ev, open("gpurun_out/pipe_trace.json", "w"))
os.remove("gpurun_out/pipe_trace_raw.json")
print("kernels traced:", len(ev), "span us:", ev[-1]["ts"] + ev[-1]["dur"] - ev[0]["ts"],
      "-> us/frame", (ev[-1]["ts"] + ev[-1]["dur"] - ev[0]["ts"]) / (6 * n_slots))
