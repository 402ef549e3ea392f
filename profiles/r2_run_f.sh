#!/bin/bash
cd /root/repo
python -m pytest tests -m gpu -q -x > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -3 gpurun_out/r2f_pytest.log
rm -f gpurun_out/r2f_summary.txt gpurun_out/r2f100_summary.txt
bash profiles/r2_sweep.sh r2f 20000 - w8 r1 bulk
bash profiles/r2_sweep.sh r2f100 100000 - w8
