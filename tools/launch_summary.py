"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list into a per-kernel table.
usage: python tools/launch_summary.py gpurun_out/launches.csv profiles/rNN_launches.json "<command>" """
import collections
import csv
import json
import sys

src, dst, cmd = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
h = rows[0]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= vi:
        continue
    a = agg.setdefault(r[ki], [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", "")) / 1e3
tot = sum(a[1] for a in agg.values())
out = {"command": cmd, "note": "ncu per-launch times are cold-cache and serialised: compare SHARES, not absolutes",
       "launches": len(rows) - 1, "total_us": round(tot, 1),
       "kernels": [{"kernel": k, "launches": c, "avg_us": round(t / c, 2), "share_pct": round(100 * t / tot, 1)}
                   for k, (c, t) in agg.items()]}
json.dump(out, open(dst, "w"), indent=1)
for k in out["kernels"]:
    print("%-90s n=%4d avg=%8.2f us share=%5.1f%%" % (k["kernel"][:90], k["launches"], k["avg_us"], k["share_pct"]))
