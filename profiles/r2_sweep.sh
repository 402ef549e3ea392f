#!/bin/bash
# on the GPU box: bench each tuning variant of libnrq (suffixes as arguments, "-" = the product library)
#   profiles/r2_sweep.sh <tag> <reads> <suffix> [<suffix> ...]
cd /root/repo
tag=$1; reads=$2; shift 2
for v in "$@"; do
  lib=/root/repo/nanoraw_b200/libnrq_$v.so; [ "$v" = "-" ] && lib=/root/repo/nanoraw_b200/libnrq.so
  NRQ_LIB_PATH=$lib python bench.py --reads $reads --unique 1000 --steps 4 --warmup 3 --no-cpu-baseline --fast5-files 0 \
    --e2e-steps 1 > gpurun_out/${tag}_$v.json 2> gpurun_out/${tag}_$v.err
  python - "$v" gpurun_out/${tag}_$v.json <<'PY' | tee -a gpurun_out/${tag}_summary.txt
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read())
    print(sys.argv[1], '%.4g' % d['value'], '%.2f ms' % d['ms_per_step'],
          {k: round(v['ms'], 3) for k, v in d['roofline']['stages'].items()},
          {k: round(v, 3) for k, v in d['roofline']['stages']['k_resegment']['parts_ms'].items()},
          'failed', d['config']['failed_reads'], 'mismatch', d['config']['oracle_spot_check_mismatches'])
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
done
