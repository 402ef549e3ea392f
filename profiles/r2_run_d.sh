#!/bin/bash
cd /root/repo
python -m pytest tests -m gpu -q -x > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -4 gpurun_out/r2d_pytest.log
rm -f gpurun_out/r2d_summary.txt
bash profiles/r2_sweep.sh r2d 20000 - w8 w7 slab p6
bash profiles/r2_sweep.sh r2d100 100000 - w8
