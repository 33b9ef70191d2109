#!/usr/bin/env python
"""Summarise an ncu capture (raw page csv + source page csv): headline metrics and a per-segment SASS breakdown."""
import csv, sys
raw, src, nwarps = sys.argv[1], sys.argv[2], float(sys.argv[3])
rows = list(csv.reader(open(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warp_latency_per_inst_issued.ratio', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active']
for h, u, v in zip(hdr, units, vals):
    if h in keep or 'issue_stalled' in h and 'per_issue_active' in h:
        print("%-80s %-14s %s" % (h, u, v))
rows = list(csv.reader(open(src)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        n = int(r[ix['Instructions Executed']])
    except Exception:
        continue
    data.append((r[ix['Source']], n, int(r[ix['# Samples']]), int(r[ix['L1 Wavefronts Shared']] or 0),
                 int(r[ix['L1 Wavefronts Shared Ideal']] or 0), float(r[ix['Avg. Threads Executed']] or 0)))
print("\ntotal warp instructions %d = %.1f per warp; samples %d" % (sum(d[1] for d in data), sum(d[1] for d in data) / nwarps, sum(d[2] for d in data)))
segs, cur = [], None
for i, (s, n, sm, w, wi, at) in enumerate(data):
    k = n / nwarps
    if cur is None or abs(k - cur[0]) > 0.15 * max(cur[0], 0.3):
        cur = [k, i, i, 0, 0, 0, 0]
        segs.append(cur)
    cur[2] = i; cur[3] += n; cur[4] += sm; cur[5] += w; cur[6] += wi
print("exec/warp  sass-range  n  instr/warp  samples  smemWF/warp  idealWF/warp  avg-threads")
for k, a, b, n, sm, w, wi in segs:
    if n / nwarps < 0.5 and sm < 20:
        continue
    at = sum(d[5] for d in data[a:b + 1]) / (b - a + 1)
    print("%6.2f  %4d-%4d %4d  %7.1f %6d  %7.1f %7.1f  %5.1f" % (k, a, b, b - a + 1, n / nwarps, sm, w / nwarps, wi / nwarps, at))
if len(sys.argv) > 4:
    for i, d in enumerate(data):
        print("%4d %7.2f %5d %6.2f %6.2f %5.1f  %s" % (i, d[1] / nwarps, d[2], d[3] / nwarps, d[4] / nwarps, d[5], d[0][:100]))
