"""Summarises gpurun_out/pipe_trace.json (tools/pipe_trace.py): per-kernel totals, how many
correlation kernels run at once, and the kernel-to-kernel gaps on one frame stream."""
import collections
import json
import sys

ev = json.load(open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/pipe_trace.json"))
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 72
t0 = ev[0]["ts"]
T = ev[-1]["ts"] + ev[-1]["dur"] - t0
print(len(ev), "kernels, span %.0f us, %.1f us/frame" % (T, T / frames))
st = collections.defaultdict(list)
for e in ev:
    st[e["name"].split("::")[-1].split("(")[0][:40]].append(e["dur"])
for n, d in sorted(st.items(), key=lambda x: -sum(x[1]))[:12]:
    print("%-40s n=%4d avg %7.1f  max %7.1f  sum/frame %6.1f" % (n, len(d), sum(d) / len(d), max(d), sum(d) / frames))
corr = [e for e in ev if "corr_async" in e["name"]]
pts = sorted([(e["ts"], 1) for e in corr] + [(e["ts"] + e["dur"], -1) for e in corr])
cur, last, hist = 0, pts[0][0], collections.Counter()
for t, d in pts:
    hist[cur] += t - last
    last = t
    cur += d
print("corr kernels running at once (time share):", {k: round(v / T, 3) for k, v in sorted(hist.items())})
# all-kernel concurrency
pts = sorted([(e["ts"], 1) for e in ev] + [(e["ts"] + e["dur"], -1) for e in ev])
cur, last, hist = 0, pts[0][0], collections.Counter()
for t, d in pts:
    hist[min(cur, 8)] += t - last
    last = t
    cur += d
print("kernels running at once (time share):", {k: round(v / T, 3) for k, v in sorted(hist.items())})
counts = collections.Counter(e["stream"] for e in ev)
big = [s for s, c in counts.items() if c == max(counts.values())]
mine = [e for e in ev if e["stream"] == sorted(big)[len(big) // 2]]
prev_end = None
for e in mine[len(mine) // 2:len(mine) // 2 + 34]:
    gap = (e["ts"] - prev_end) if prev_end else 0
    print("%9.1f  +%7.1f gap  dur %6.1f  %s grid=%s" % (e["ts"] - t0, gap, e["dur"],
                                                       e["name"].split("::")[-1].split("(")[0][:28], e["grid"]))
    prev_end = e["ts"] + e["dur"]
