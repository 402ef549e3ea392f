#!/bin/bash
cd /root/repo
tag=${1:-r2g}
python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
( time python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err ) 2>&1 | grep real
python - gpurun_out/${tag}_bench.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read())
print('value %.4g ms %.2f e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']))
print({k:round(v['ms'],3) for k,v in d['roofline']['stages'].items()}, d['roofline']['stages']['k_resegment'].get('parts_ms'))
for k in ('cpu_baseline','cpu_baseline_fast5','e2e_fast5','tie_order_divergence'):
    v=d.get(k); 
    if v: print(k, {a:(('%.4g'%b) if isinstance(b,float) else b) for a,b in v.items() if a not in ('what','sample','against')})
PY
tail -5 gpurun_out/${tag}_bench.err
( time python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.json 2>> gpurun_out/${tag}_bench.err ) 2>&1 | grep real
python profiles/copy_probe.py > gpurun_out/${tag}_copy_probe_1gpu.json 2>> gpurun_out/${tag}_bench.err; cat gpurun_out/${tag}_copy_probe_1gpu.json
