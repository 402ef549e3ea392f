#!/bin/bash
# GPU tests, then the FAST5 leg of bench.py with its host-work breakdown
cd /root/repo
tag=${1:-r2x}
python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
python bench.py --reads 20000 --unique 1000 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${tag}_fast5.json 2> gpurun_out/${tag}_fast5.err
python - gpurun_out/${tag}_fast5.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read()); f=d['e2e_fast5']
    print('e2e_fast5 %.4g reads/s %.0f failed %s' % (f['value'], f['reads_per_s'], f['failed_reads']))
    print({k:v for k,v in f['host_work'].items() if k!='what'})
except Exception as e:
    print('FAILED', e)
PY
tail -3 gpurun_out/${tag}_fast5.err
