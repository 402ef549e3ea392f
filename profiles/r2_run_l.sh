#!/bin/bash
# FAST5-in / FAST5-out leg of bench.py at several file counts and device batch sizes (native libnrqio path)
cd /root/repo
tag=${1:-r2t}
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    f=d.get('e2e_fast5') or {}
    print(sys.argv[1], 'e2e_fast5 %.4g'%f.get('value',0), 'reads/s %.0f'%f.get('reads_per_s',0), 'failed', f.get('failed_reads'), 'sec %.2f'%f.get('seconds',0), 'files', f.get('files'), 'batch', f.get('batch_reads'), 'setup %.1f'%f.get('setup_seconds',0))
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
}
nproc
for cfg in "8192 2048" "16384 2048" "16384 1024" "16384 4096"; do
  set -- $cfg
  NANORAW_B200_BATCH_READS=$2 python bench.py --reads 20000 --unique 1000 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --fast5-files $1 > gpurun_out/${tag}_fast5_$1_$2.json 2> gpurun_out/${tag}_fast5_$1_$2.err
  show gpurun_out/${tag}_fast5_$1_$2.json
done
