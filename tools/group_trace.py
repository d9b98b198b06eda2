"""Kernel timeline (CUPTI via torch.profiler) of the frame pipeline as bench.py runs it: graphs of
ABL_GROUP consecutive frames, one replay stream per group, `slots` resident slots. Prints what the S4
launches look like inside the pipeline (duration, gaps between consecutive launches, how much of the
time none / one / two are resident) and how many other kernels run at once, inside and outside S4.
usage: ABL_GROUP=8 python tools/group_trace.py [slots] [skip,stages]"""
import collections
import json
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dodt_b200 import synth  # noqa: E402
from dodt_b200.frontend import FrontEnd, FrontEndConfig, HostFrame  # noqa: E402

n_slots = int(sys.argv[1]) if len(sys.argv) > 1 else 32
skip = tuple(s for s in (sys.argv[2].split(",") if len(sys.argv) > 2 else []) if s)
GROUP = int(os.environ.get("ABL_GROUP", "8"))
cfg = FrontEndConfig()
if os.environ.get("ABL_PRIO"):
    cfg.chain_stream_priority, cfg.corr_stream_priority = (int(v) for v in os.environ["ABL_PRIO"].split(","))
fe = FrontEnd(cfg)
slots = [fe.new_slot() for _ in range(n_slots)]
for i, s in enumerate(slots):
    HostFrame(fe).fill(synth.frame_inputs(2, i)).upload(s)
torch.cuda.synchronize()
n_groups = n_slots // GROUP
streams = [torch.cuda.Stream() for _ in range(n_groups)]
main = torch.cuda.current_stream()
graphs = [fe.capture_group(slots[g * GROUP:(g + 1) * GROUP], slots[g * GROUP - 1], None, skip)[0]
          for g in range(n_groups)]


def rr(sweeps):
    for st in streams:
        st.wait_stream(main)
    for i in range(sweeps * n_groups):
        with torch.cuda.stream(streams[i % n_groups]):
            graphs[i % n_groups].replay()
    for st in streams:
        main.wait_stream(st)


rr(6)
torch.cuda.synchronize()
SWEEPS = 5
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    rr(SWEEPS)
    torch.cuda.synchronize()
os.makedirs("gpurun_out", exist_ok=True)
prof.export_chrome_trace("gpurun_out/group_trace_raw.json")
raw = json.load(open("gpurun_out/group_trace_raw.json"))
os.remove("gpurun_out/group_trace_raw.json")
ev = sorted(({"name": e["name"], "ts": e["ts"], "dur": e["dur"]} for e in raw["traceEvents"] if e.get("cat") == "kernel"),
            key=lambda e: e["ts"])
frames = SWEEPS * n_slots
t0, t1 = ev[0]["ts"], max(e["ts"] + e["dur"] for e in ev)
T = t1 - t0
print("%d kernels, span %.0f us, %.1f us/frame (under the profiler)" % (len(ev), T, T / frames))
short = lambda n: n.split("::")[-1].split("(")[0].split("<")[0][:28]
tot = collections.defaultdict(list)
for e in ev:
    tot[short(e["name"])].append(e["dur"])
for n, d in sorted(tot.items(), key=lambda x: -sum(x[1]))[:14]:
    print("  %-28s n=%4d avg %7.1f max %7.1f sum/frame %6.1f" % (n, len(d), sum(d) / len(d), max(d), sum(d) / frames))
corr = [e for e in ev if "corr_" in e["name"]]
other = [e for e in ev if "corr_" not in e["name"]]
if corr:
    durs = [e["dur"] for e in corr]
    starts = [e["ts"] for e in corr]
    ends = sorted(e["ts"] + e["dur"] for e in corr)
    print("S4 launches: n=%d dur avg %.1f min %.1f max %.1f us; start-to-start avg %.1f us" %
          (len(corr), sum(durs) / len(durs), min(durs), max(durs), (starts[-1] - starts[0]) / max(1, len(starts) - 1)))


def concurrency(evs, cap=12):
    pts = sorted([(e["ts"], 1) for e in evs] + [(e["ts"] + e["dur"], -1) for e in evs])
    cur, last, hist = 0, t0, collections.Counter()
    for t, d in pts:
        hist[min(cur, cap)] += t - last
        last = t
        cur += d
    hist[0] += t1 - last
    return {k: round(v / T, 3) for k, v in sorted(hist.items())}


print("S4 kernels resident at once (time share):", concurrency(corr, 4))
print("other kernels running at once (time share):", concurrency(other))
# other-kernel concurrency while >= 1 S4 kernel is resident vs while none is
marks = sorted([(e["ts"], 0, 1) for e in corr] + [(e["ts"] + e["dur"], 0, -1) for e in corr] +
               [(e["ts"], 1, 1) for e in other] + [(e["ts"] + e["dur"], 1, -1) for e in other])
c4 = co = 0
last = t0
acc = {True: [0.0, 0.0], False: [0.0, 0.0]}      # [time, other-kernel-time integral]
for t, kind, d in marks:
    acc[c4 > 0][0] += t - last
    acc[c4 > 0][1] += (t - last) * co
    last = t
    if kind == 0:
        c4 += d
    else:
        co += d
for k in (True, False):
    tt, integ = acc[k]
    print("while %s S4 kernel is resident: %.0f us (%.0f%% of the span), mean other kernels running %.2f, "
          "other-kernel time %.0f us" % ("an" if k else "no", tt, 100 * tt / T, integ / tt if tt else 0, integ))
