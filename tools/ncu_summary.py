"""Summarise an ncu export: tools/ncu_summary.py raw.csv src.csv"""
import collections, csv, re, sys
raw, src = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__cycles_elapsed.max',
        'smsp__cycles_active.avg', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active']
keys += [h for h in hdr if 'issue_stalled' in h and 'per_issue_active' in h]
for w in keys:
    if w in hdr:
        i = hdr.index(w)
        print('%-88s %-8s %s' % (w.replace('smsp__average_warps_issue_stalled_', 'stall:'), units[i], data[0][i]))
rows = list(csv.reader(open(src)))
hidx = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
h = rows[hidx[0]]
end = hidx[1] - 1 if len(hidx) > 1 else len(rows)
d = rows[hidx[0] + 1:end]
ci = {x: i for i, x in enumerate(h)}
def g(r, k):
    try:
        return int(r[ci[k]] or 0)
    except Exception:
        return 0
ops, samp = collections.Counter(), collections.Counter()
for r in d:
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[ci['Source']])
    op = m.group(2).split('.')[0] if m else '?'
    ops[op] += g(r, 'Instructions Executed'); samp[op] += g(r, '# Samples')
ti, ts = sum(ops.values()), sum(samp.values())
print('total inst', ti, 'samples', ts)
for op, c in ops.most_common(12):
    print('  %-10s inst=%9d (%4.1f%%) samples=%6d (%4.1f%%)' % (op, c, 100 * c / ti, samp[op], 100 * samp[op] / max(ts, 1)))
for reason in ['stall_long_sb', 'stall_short_sb', 'stall_wait', 'stall_mio', 'stall_dispatch', 'stall_barrier', 'stall_math', 'stall_lg', 'stall_not_selected']:
    tot = sum(g(r, reason) for r in d)
    top = sorted(range(len(d)), key=lambda i: -g(d[i], reason))[:3]
    print('== %s %d (%.1f%%)' % (reason, tot, 100 * tot / max(ts, 1)))
    for i in top:
        print('     %5d %-56s %d' % (i, d[i][ci['Source']][:56], g(d[i], reason)))
