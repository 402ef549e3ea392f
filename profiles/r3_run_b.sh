#!/bin/bash
cd /root/repo
tag=$1; shift
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
rm -f gpurun_out/${tag}_summary.txt
bash profiles/r2_sweep.sh $tag 100000 "$@"
