#!/usr/bin/env python
"""Per-source-line share of executed warp instructions and stall samples from an ncu report.

    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > X_src.csv
    python profiles/hot_lines.py X_src.csv [top N]

The CSV interleaves, per kernel and per source file, one row per source line (with totals) followed by the
SASS rows attributed to it; only the source-line rows are used.
"""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    kernel, fname, header = None, None, None
    per = defaultdict(lambda: defaultdict(lambda: [0, 0, '']))  # kernel -> (file, line) -> [inst, samples, text]
    with open(path, newline='') as fp:
        for row in csv.reader(fp):
            if not row:
                continue
            if row[0] == 'File Path':
                fname = row[1].split('/')[-1]
                continue
            if row[0] == 'Function Name':
                kernel = row[1].split('(')[0]
                continue
            if row[0] == 'Line No':
                header = row
                i_inst = header.index('Instructions Executed')
                i_smp = header.index('# Samples')
                continue
            if header is None or row[0] in ('', '-') or not row[0].isdigit():
                continue
            try:
                inst = int(row[i_inst])
                smp = int(row[i_smp])
            except ValueError:
                continue
            e = per[kernel][(fname, int(row[0]))]
            e[0] += inst
            e[1] += smp
            e[2] = row[1].strip()
    for kernel, lines in per.items():
        ti = sum(v[0] for v in lines.values()) or 1
        ts = sum(v[1] for v in lines.values()) or 1
        print('==== %s: %d warp instructions, %d stall samples' % (kernel, ti, ts))
        for (f, ln), v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
            print('%5.1f%% inst %5.1f%% smp  %s:%d  %s' % (100.0 * v[0] / ti, 100.0 * v[1] / ts, f, ln, v[2][:90]))


if __name__ == '__main__':
    main()
