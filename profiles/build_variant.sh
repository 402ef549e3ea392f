#!/bin/bash
# build a tuning variant of libnrq next to the product library:
#   profiles/build_variant.sh <suffix> [-DNAME=VALUE ...]   ->  nanoraw_b200/libnrq_<suffix>.so
# run it with NRQ_LIB_PATH=/root/repo/nanoraw_b200/libnrq_<suffix>.so python bench.py ...
set -e
cd "$(dirname "$0")/.."
sfx=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false --shared \
  -Xcompiler -fPIC -Xptxas -v "$@" -o nanoraw_b200/libnrq_$sfx.so nanoraw_b200/csrc/nrq_api.cu 2> /tmp/ptxas_$sfx.txt
grep -A2 "Compiling entry function '_ZN3nrq1[13]k_\(event_tableE\|resegment\)" /tmp/ptxas_$sfx.txt | grep -v Compiling
