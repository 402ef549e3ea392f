#!/bin/bash
# r2b: speculative search with its own scratch; tests, benches, one ncu --set full capture at 8 k reads
cd /root/repo
tag=${1:-r2b}
python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -25 gpurun_out/${tag}_pytest.log
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    print(sys.argv[1], '%.4g'%d['value'], '%.2f ms'%d['ms_per_step'], {k:round(v['ms'],3) for k,v in d['roofline']['stages'].items()}, d['roofline']['stages']['k_resegment'].get('parts_ms'), 'failed', d['config']['failed_reads'], 'mismatch', d['config']['oracle_spot_check_mismatches'], 'e2e %.4g'%d['e2e']['value'])
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
}
for fl in 0 4; do
  python bench.py --reads 20000 --unique 1000 --steps 4 --warmup 3 --no-cpu-baseline --fast5-files 0 --e2e-steps 1 --debug-flags $fl > gpurun_out/${tag}_bench20k_f$fl.json 2> gpurun_out/${tag}_bench20k_f$fl.err
  show gpurun_out/${tag}_bench20k_f$fl.json
done
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --fast5-files 0 --e2e-steps 2 > gpurun_out/${tag}_bench100k.json 2> gpurun_out/${tag}_bench100k.err
show gpurun_out/${tag}_bench100k.json
if [ "$2" = "ncu" ]; then
  ncu --set full --clock-control none --import-source on -k regex:'k_speculate|k_resegment' -c 2 -o gpurun_out/${tag}_prof -f \
    python bench.py --reads 8000 --unique 1000 --steps 1 --warmup 0 --no-cpu-baseline --fast5-files 0 --e2e-steps 1 > gpurun_out/${tag}_ncu.log 2>&1
  ncu -i gpurun_out/${tag}_prof.ncu-rep --page raw --csv 2>/dev/null | python profiles/ncu_raw.py > gpurun_out/${tag}_ncu_summary.txt
  ncu -i gpurun_out/${tag}_prof.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${tag}_src.csv 2>/dev/null
  python profiles/hot_lines.py gpurun_out/${tag}_src.csv 45 > gpurun_out/${tag}_hot_lines.txt
  cat gpurun_out/${tag}_ncu_summary.txt
fi
