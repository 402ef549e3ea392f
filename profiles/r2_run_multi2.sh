#!/bin/bash
# profiles/r2_run_multi2.sh <tag> <N>: bench.py on N GPUs of one box (no copy probe: see r2m_copy_probe_*), with the
# FAST5-in / FAST5-out leg on every rank
cd /root/repo
tag=$1; N=$2
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${tag}_bench_${N}gpu.json 2> gpurun_out/${tag}_bench_${N}gpu.err ) 2>&1 | grep real
python - gpurun_out/${tag}_bench_${N}gpu.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    print('N', d['n_gpus'], 'value %.4g ms %.2f e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']))
    v=d.get('e2e_fast5'); print('e2e_fast5', {a:(('%.4g'%b) if isinstance(b,float) else b) for a,b in v.items() if a not in ('what',)} if v else None)
except Exception as e:
    print('FAILED', e)
PY
tail -3 gpurun_out/${tag}_bench_${N}gpu.err
nproc
