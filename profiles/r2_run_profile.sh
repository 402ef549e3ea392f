#!/bin/bash
# final evidence of the round: bench lines (both arms), ncu launch list, ncu --set full of one 100 k-read step
cd /root/repo
tag=${1:-r2z}
K='regex:k_align|k_scan_counts|k_norm_stats|k_speculate|k_resegment|k_event_table'
python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.json 2>> gpurun_out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --fast5-files 0 --e2e-steps 1 > gpurun_out/${tag}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 7 -c 7 -o gpurun_out/${tag}_prof -f \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --fast5-files 0 --e2e-steps 1 > gpurun_out/${tag}_ncu.log 2>&1
ncu -i gpurun_out/${tag}_prof.ncu-rep --page raw --csv 2>/dev/null > gpurun_out/${tag}_raw.csv
python profiles/ncu_raw.py < gpurun_out/${tag}_raw.csv > gpurun_out/${tag}_ncu_full_100k_summary.txt
python profiles/make_traffic_json.py "ncu --set full --clock-control none, one step of bench.py (100 k reads), $tag" < gpurun_out/${tag}_raw.csv > gpurun_out/${tag}_traffic_100k.json
ncu -i gpurun_out/${tag}_prof.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${tag}_src.csv 2>/dev/null
python profiles/hot_lines.py gpurun_out/${tag}_src.csv 40 > gpurun_out/${tag}_hot_lines.txt
rm -f gpurun_out/${tag}_src.csv gpurun_out/${tag}_raw.csv
cuobjdump -sass nanoraw_b200/libnrq.so | grep -c "UBLKCP\|UTMALDG" > gpurun_out/${tag}_sass_bulk_count.txt
cat gpurun_out/${tag}_ncu_full_100k_summary.txt | grep "====\|time_duration\|bytes_read\|bytes_write\|inst_executed.sum\|issue_active"
