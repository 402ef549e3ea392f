#!/bin/bash
# r3_sweep_cfg.sh <tag> <config> <reads> <suffix>...: bench variants on another BASELINE config
cd /root/repo
tag=$1; cfg=$2; reads=$3; shift 3
for v in "$@"; do
  lib=/root/repo/nanoraw_b200/libnrq_$v.so; [ "$v" = "-" ] && lib=/root/repo/nanoraw_b200/libnrq.so
  NRQ_LIB_PATH=$lib python bench.py --config $cfg --reads $reads --unique 1000 --steps 3 --warmup 2 --no-cpu-baseline --fast5-files 0 \
    --e2e-steps 1 > gpurun_out/${tag}_${cfg}_$v.json 2> gpurun_out/${tag}_${cfg}_$v.err
  python - "$v" gpurun_out/${tag}_${cfg}_$v.json <<'PY' | tee -a gpurun_out/${tag}_summary.txt
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read())
    print(sys.argv[1], d['config']['workload'][:40], '%.4g' % d['value'], '%.2f ms' % d['ms_per_step'],
          {k: round(v['ms'], 3) for k, v in d['roofline']['stages'].items()},
          {k: round(v, 3) for k, v in d['roofline']['stages']['k_resegment']['parts_ms'].items()},
          'failed', d['config']['failed_reads'], 'mismatch', d['config']['oracle_spot_check_mismatches'])
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
done
