#!/bin/bash
# r2k+: native FAST5 file functions (libnrqio nrqio_f5_*): GPU tests, then the FAST5-in / FAST5-out leg of bench.py
# with the native path and, for comparison, with NRQ_FAST5_NATIVE=0 (hdf5_min.py on the I/O processes)
cd /root/repo
tag=${1:-r2s}
python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    f=d.get('e2e_fast5') or {}
    print(sys.argv[1], 'value %.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], 'e2e_fast5 %.4g'%f.get('value',0), 'reads/s %.0f'%f.get('reads_per_s',0), 'failed', f.get('failed_reads'), 'sec %.2f'%f.get('seconds',0), 'files', f.get('files'), 'batch', f.get('batch_reads'))
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
}
nproc
for native in 1 0; do
  NRQ_FAST5_NATIVE=$native python bench.py --reads 20000 --unique 1000 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${tag}_fast5_native$native.json 2> gpurun_out/${tag}_fast5_native$native.err
  show gpurun_out/${tag}_fast5_native$native.json
done
