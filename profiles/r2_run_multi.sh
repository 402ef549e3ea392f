#!/bin/bash
# profiles/r2_run_multi.sh <tag> <N>: copy-only probe and bench.py on N GPUs of one box
cd /root/repo
tag=$1; N=$2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  profiles/copy_probe.py > gpurun_out/${tag}_copy_probe_${N}gpu.json 2> gpurun_out/${tag}_copy_probe_${N}gpu.err
cat gpurun_out/${tag}_copy_probe_${N}gpu.json
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
nproc; nvidia-smi topo -m 2>/dev/null | head -12 > gpurun_out/${tag}_topo_${N}gpu.txt
