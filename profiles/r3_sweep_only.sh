#!/bin/bash
cd /root/repo
tag=$1; shift
rm -f gpurun_out/${tag}_summary.txt
bash profiles/r2_sweep.sh $tag 100000 "$@"
