#!/bin/bash
# r2a: first run of the speculative search
cd /root/repo
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -15 gpurun_out/r2a_pytest.log
for fl in 0 4; do
  python bench.py --reads 20000 --unique 1000 --steps 4 --warmup 3 --no-cpu-baseline --fast5-files 0 --e2e-steps 1 --debug-flags $fl > gpurun_out/r2a_bench20k_f$fl.json 2> gpurun_out/r2a_bench20k_f$fl.err
  python - gpurun_out/r2a_bench20k_f$fl.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read())
print(sys.argv[1], '%.4g'%d['value'], '%.2f ms'%d['ms_per_step'], {k:round(v['ms'],3) for k,v in d['roofline']['stages'].items()}, d['roofline']['stages']['k_resegment'].get('parts_ms'), 'failed', d['config']['failed_reads'], 'mismatch', d['config']['oracle_spot_check_mismatches'])
PY
done
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --fast5-files 0 --e2e-steps 2 > gpurun_out/r2a_bench100k.json 2> gpurun_out/r2a_bench100k.err
python - gpurun_out/r2a_bench100k.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read())
print(sys.argv[1], '%.4g'%d['value'], '%.2f ms'%d['ms_per_step'], {k:round(v['ms'],3) for k,v in d['roofline']['stages'].items()}, d['roofline']['stages']['k_resegment'].get('parts_ms'), 'failed', d['config']['failed_reads'], 'mismatch', d['config']['oracle_spot_check_mismatches'], 'e2e %.4g'%d['e2e']['value'])
PY
