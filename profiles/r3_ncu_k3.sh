#!/bin/bash
# instruction counts / issue utilisation of k_event_table for libnrq variants (suffixes as arguments, "-" = product)
cd /root/repo
tag=$1; shift
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes_read.sum,dram__bytes_write.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,sm__warps_active.avg.pct_of_peak_sustained_active
for v in "$@"; do
  lib=/root/repo/nanoraw_b200/libnrq_$v.so; [ "$v" = "-" ] && lib=/root/repo/nanoraw_b200/libnrq.so
  NRQ_LIB_PATH=$lib ncu --metrics $M --clock-control none -k regex:k_event_table --launch-skip 3 -c 1 --csv \
    --log-file gpurun_out/${tag}_$v.csv python bench.py --reads 100000 --unique 1000 --steps 1 --warmup 1 \
    --no-cpu-baseline --fast5-files 0 --e2e-steps 1 > gpurun_out/${tag}_$v.log 2>&1
  echo "== $v"; python - gpurun_out/${tag}_$v.csv <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]
for r in rows[1:]:
    d = dict(zip(h, r))
    print('  %-80s %s %s' % (d['Metric Name'], d['Metric Value'], d['Metric Unit']))
PY
done
