#!/bin/bash
# round 2, third session: GPU suite, then the K3 deferred-round variant against the build without it (100 k reads)
cd /root/repo
python -m pytest tests -m gpu -x -q > gpurun_out/r3a_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r3a_pytest.log
tail -3 gpurun_out/r3a_pytest.log
rm -f gpurun_out/r3a_summary.txt
bash profiles/r2_sweep.sh r3a 100000 - nodefer
