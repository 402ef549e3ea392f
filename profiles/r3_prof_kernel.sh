#!/bin/bash
# ncu --set full with source of ONE kernel's largest launch at <reads> reads:  r3_prof_kernel.sh <tag> <kernel regex> <reads> <skip> [lib suffix]
cd /root/repo
tag=$1; K=$2; reads=$3; skip=$4; v=${5:--}
lib=/root/repo/nanoraw_b200/libnrq_$v.so; [ "$v" = "-" ] && lib=/root/repo/nanoraw_b200/libnrq.so
NRQ_LIB_PATH=$lib ncu --set full --clock-control none --import-source on -k "regex:$K" --launch-skip $skip -c 1 -o gpurun_out/${tag}_prof -f \
  python bench.py --reads $reads --unique 1000 --steps 1 --warmup 1 --no-cpu-baseline --fast5-files 0 --e2e-steps 1 > gpurun_out/${tag}_ncu.log 2>&1
ncu -i gpurun_out/${tag}_prof.ncu-rep --page raw --csv 2>/dev/null | python profiles/ncu_raw.py > gpurun_out/${tag}_summary.txt
ncu -i gpurun_out/${tag}_prof.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${tag}_src.csv 2>/dev/null
python profiles/hot_lines.py gpurun_out/${tag}_src.csv 70 > gpurun_out/${tag}_hot_lines.txt
rm -f gpurun_out/${tag}_prof.ncu-rep
cat gpurun_out/${tag}_summary.txt
