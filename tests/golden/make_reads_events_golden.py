"""Generate tests/golden/reads_events_v1.npz with the REAL reference: get_reads_events (nanoraw_helper.py:446-484)
over the same corrected reads as aggregate_v1.npz (same generator, same seeds).  Per (chromosome, strand) track:
the positions that have events, and the event means of each position.  The reference orders the means of one
position by an unstable argsort of the positions; the fixture stores them sorted by value, as a multiset.
Build container only."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from nanoraw_b200 import synth  # noqa: E402
from oracle import run_reference as rr  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden', 'reads_events_v1.npz')
GENOME_LEN = 4000


def main():
    ref, ref_nh = rr.load_reference(stable_sort=True)
    cfg = synth.SynthConfig(span_mean=350, span_sd=80, seed_base=8100)
    genome = synth.make_genome(GENOME_LEN, seed=31)
    levels = synth.kmer_levels(cfg)
    files = []
    for i in range(28):
        read = synth.make_read(genome, levels, cfg, i)
        read.chrom = 'chrA' if i % 4 else 'chrB'
        name = rr.make_mem_fast5('mem://rev_golden/%d' % i, read)
        rr.run_resquiggle_read(ref, name, read)
        files.append(name)
    blob = {}
    with np.errstate(all='raise'):
        rrc = ref_nh.parse_fast5s(files, 'RawGenomeCorrected_000', ['BaseCalled_template'])
        for k, (key, reads) in enumerate(sorted(rrc.items())):
            events = ref_nh.get_reads_events(reads, key[1] == '-')
            pos = np.array(sorted(events), dtype=np.int64)
            vals = [np.sort(np.asarray(events[p], dtype=np.float64)) for p in pos]
            blob['t%d/pos' % k] = pos
            blob['t%d/off' % k] = np.concatenate([[0], np.cumsum([len(v) for v in vals])]).astype(np.int64)
            blob['t%d/values' % k] = np.concatenate(vals)
    blob['meta'] = np.frombuffer(json.dumps({'tracks': [list(k) for k in sorted(rrc)]}).encode(), dtype=np.uint8)
    np.savez_compressed(OUT, **blob)
    print('wrote', OUT, os.path.getsize(OUT), 'bytes; tracks', sorted(rrc))


if __name__ == '__main__':
    main()
