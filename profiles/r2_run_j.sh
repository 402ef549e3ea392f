#!/bin/bash
cd /root/repo
tag=${1:-r2j}
python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    print(sys.argv[1], '%.4g'%d['value'], '%.2f ms'%d['ms_per_step'], {k:round(v['ms'],3) for k,v in d['roofline']['stages'].items()}, d['roofline']['stages']['k_resegment'].get('parts_ms'), 'failed', d['config']['failed_reads'], 'mismatch', d['config']['oracle_spot_check_mismatches'], 'e2e %.4g'%d['e2e']['value'])
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
}
for parts in 1 16; do
  NRQ_EV_PARTS=$parts python bench.py --steps 5 --warmup 3 --no-cpu-baseline --fast5-files 0 --e2e-steps 2 > gpurun_out/${tag}_std_parts$parts.json 2> gpurun_out/${tag}_std.err
  show gpurun_out/${tag}_std_parts$parts.json
  NRQ_EV_PARTS=$parts python bench.py --config long --reads 10000 --unique 500 --steps 5 --warmup 3 --no-cpu-baseline --fast5-files 0 --e2e-steps 2 --e2e-reps 1 > gpurun_out/${tag}_long_parts$parts.json 2> gpurun_out/${tag}_long.err
  show gpurun_out/${tag}_long_parts$parts.json
done
python - <<'PY' 2>&1 | tail -4
import numpy as np, torch
from nanoraw_b200 import synth
from nanoraw_b200.batch import ReadBatch
from nanoraw_b200.engine import ResquiggleEngine, ResquiggleParams
eng = ResquiggleEngine()
genome = synth.make_genome(1_000_000, seed=9)
reads = synth.make_reads(8, synth.SynthConfig(span_mean=100_000, span_sd=5_000, seed_base=45000), genome=genome)
for n in (1, 8):
    db = eng.upload(ReadBatch.from_reads(reads[:n]))
    pc = ResquiggleParams().to_c()
    out = eng.alloc_results(db); ws = eng.workspace(db.n_reads, db.total_align, pc)
    for _ in range(3): eng.run_device(db, pc, out=out, workspace=ws)
    torch.cuda.synchronize()
    ms = [eng.run_device_profiled(db, pc, out, ws) for _ in range(5)]
    print('long reads in the batch:', n, 'stage ms (scan, stats, speculate, walk, events):', np.round(np.mean(ms, 0), 3), 'total %.3f ms' % np.mean(ms, 0).sum())
PY
