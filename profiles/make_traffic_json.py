#!/usr/bin/env python
"""DRAM bytes, warp instructions and time per launch of every libnrq kernel from an `ncu --set full` report:

    ncu -i X.ncu-rep --page raw --csv | python profiles/make_traffic_json.py "<what was profiled>" > X_traffic.json

bench.py reads the committed file for `roofline.traffic` and the per-stage instruction-issue figures."""
import csv
import json
import sys

rows = list(csv.reader(sys.stdin))
h, units = rows[0], rows[1]
scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}
out = {'source': sys.argv[1] if len(sys.argv) > 1 else 'ncu --set full --clock-control none', 'kernels': {}}
for r in rows[2:]:
    d = dict(zip(h, r))
    name = d['Kernel Name'].split('(')[0].replace('void ', '').split('<')[0]

    def val(key, conv=None):
        v = float(d[key].replace(',', ''))
        return v * (conv or {}).get(units[h.index(key)], 1.0)
    out['kernels'][name] = {
        'dram_read_bytes': int(val('dram__bytes_read.sum', scale)),
        'dram_write_bytes': int(val('dram__bytes_write.sum', scale)),
        'warp_instructions': int(val('smsp__inst_executed.sum')),
        'issue_active_pct': val('smsp__issue_active.avg.pct_of_peak_sustained_active'),
        'ncu_ms': val('gpu__time_duration.sum', {'us': 1e-3, 'ms': 1.0, 'ns': 1e-6, 's': 1e3}),
    }
json.dump(out, sys.stdout, indent=1)
print()
