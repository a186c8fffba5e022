#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) or an ncu launch-list csv.  Usage:
   python tools/ncu_summary.py raw  <file.ncu-rep>
   python tools/ncu_summary.py list <launches.csv>"""
import csv, subprocess, sys
from collections import defaultdict

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum', 'lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__occupancy_limit_registers', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.max', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_barrier_per_warp_active.pct', 'smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct',
        'sm__inst_executed.sum', 'dram__cycles_active.avg', 'dram__cycles_elapsed.avg']

def raw(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('==', r[hdr.index('Kernel Name')])
        for w in WANT:
            if w in hdr:
                print(f'  {w} = {r[hdr.index(w)]} {units[hdr.index(w)]}')

def lst(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
    hdr = rows[hi]; kn = hdr.index('Kernel Name'); mv = hdr.index('Metric Value')
    d = defaultdict(list)
    for r in rows[hi + 1:]:
        if len(r) > mv:
            d[r[kn][:70]].append(float(r[mv].replace(',', '')))
    tot = sum(sum(v) for v in d.values())
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        print(f'{k:72s} n={len(v):4d} avg={sum(v)/len(v)/1e3:9.1f} us  share={sum(v)/tot*100:5.1f}%')

if __name__ == '__main__':
    (raw if sys.argv[1] == 'raw' else lst)(sys.argv[2])
