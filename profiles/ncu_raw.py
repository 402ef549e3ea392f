#!/usr/bin/env python
"""Key raw-page metrics of every kernel in an ncu report:  ncu -i X.ncu-rep --page raw --csv | python profiles/ncu_raw.py"""
import csv
import sys
KEYS = ('gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active')
rows = list(csv.reader(sys.stdin))
h, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(h, r))
    print('====', d.get('Kernel Name', '?').split('(')[0])
    for k in KEYS:
        if k in d:
            print('  %-75s %s %s' % (k, d[k], units[h.index(k)]))
    for k in h:
        if 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and k in d:
            try:
                if float(d[k]) >= 0.3:
                    print('  %-75s %s' % (k, d[k]))
            except ValueError:
                pass
