"""Per-kernel table of one frame from an `ncu --set full --page raw --csv` export.
usage: python tools/frame_ncu_table.py raw.csv out.md "<command that produced the capture>" """
import csv
import sys

raw, out, cmd = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
rows = list(csv.reader(open(raw)))
h, data = rows[0], rows[2:]
col = {n: i for i, n in enumerate(h)}


def f(r, name, default=0.0):
    try:
        return float(r[col[name]].replace(",", ""))
    except Exception:
        return default


stall_cols = [n for n in h if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio")]
PEAK = 6554.2
lines = ["# ncu `--set full` capture of one frame (round 1)", "", "`%s`" % cmd, "",
         "ncu serialises the kernels and replays each ~40 times with its own cache control, so the "
         "times are per-kernel and colder than in the pipelined run: read the SHARES and the "
         "per-kernel rates, not the sum. DRAM GB/s = (dram read + write bytes) / duration; %% of the "
         "measured copy bandwidth (%.0f GB/s, MEASURED_PEAKS.json)." % PEAK, "",
         "| # | kernel | grid x block | regs | time us | DRAM rd MB | DRAM wr MB | DRAM GB/s | % peak | warps act % | issue % | L2 tput % | top stalls (per issue) |",
         "|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
tot = sum(f(r, "gpu__time_duration.sum") for r in data)
for k, r in enumerate(data):
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("dodt::<unnamed>::", "")
    t = f(r, "gpu__time_duration.sum")
    rd, wr = f(r, "dram__bytes_read.sum"), f(r, "dram__bytes_write.sum")
    unit_rd = rows[1][col["dram__bytes_read.sum"]]
    scale = {"Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6, "Gbyte": 1e3}.get(unit_rd, 1.0)
    unit_wr = rows[1][col["dram__bytes_write.sum"]]
    scale_w = {"Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6, "Gbyte": 1e3}.get(unit_wr, 1.0)
    rd, wr = rd * scale, wr * scale_w
    unit_t = rows[1][col["gpu__time_duration.sum"]]
    t_us = t * {"us": 1.0, "ns": 1e-3, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(unit_t, 1.0)
    gbs = (rd + wr) / t_us * 1e3 if t_us else 0.0
    stalls = sorted(((f(r, c), c[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")])
                     for c in stall_cols if c.split("_stalled_")[1].split("_per_")[0] not in ("selected",)), reverse=True)[:3]
    lines.append("| %d | `%s` | %d x %d | %d | %.1f | %.2f | %.2f | %.0f | %.0f | %.0f | %.0f | %.0f | %s |" % (
        k, name, f(r, "launch__grid_size"), f(r, "launch__block_size"), f(r, "launch__registers_per_thread"),
        t_us, rd, wr, gbs, 100 * gbs / PEAK, f(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        f(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        f(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        ", ".join("%s %.2f" % (n, v) for v, n in stalls)))
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[-len(data) - 2:]))
