"""Generate tests/golden/aggregate_v1.npz with the REAL reference: its resquiggle_read writes corrected groups
into in-memory FAST5 files, then its parse_fast5s / get_coverage / get_base_means / get_base_sds /
get_base_lengths (nanoraw_helper.py:110-152, 298-444) aggregate them.  Build container only."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from nanoraw_b200 import synth  # noqa: E402
from oracle import run_reference as rr  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden', 'aggregate_v1.npz')
GENOME_LEN = 4000


def main():
    ref, ref_nh = rr.load_reference(stable_sort=True)
    cfg = synth.SynthConfig(span_mean=350, span_sd=80, seed_base=8100)
    genome = synth.make_genome(GENOME_LEN, seed=31)
    levels = synth.kmer_levels(cfg)
    files, blob, meta = [], {}, []
    for i in range(28):
        read = synth.make_read(genome, levels, cfg, i)
        read.chrom = 'chrA' if i % 4 else 'chrB'
        name = rr.make_mem_fast5('mem://agg_golden/%d' % i, read)
        rr.run_resquiggle_read(ref, name, read)
        files.append(name)
    with np.errstate(all='raise'):
        rrc = ref_nh.parse_fast5s(files, 'RawGenomeCorrected_000', ['BaseCalled_template'])
        cov = ref_nh.get_coverage(rrc)
        sizes = {'chrA': GENOME_LEN, 'chrB': GENOME_LEN}
        have = {c: GENOME_LEN for c in sizes if (c, '+') in rrc and (c, '-') in rrc}
        means = ref_nh.get_base_means(rrc, sizes)
        sds = ref_nh.get_base_sds(rrc, have)
        lens = ref_nh.get_base_lengths(rrc, have)
    from nanoraw_b200 import fast5_io
    for k, (key, reads) in enumerate(sorted(rrc.items())):
        for j, r in enumerate(reads):
            ev = fast5_io.MemFile(r.fn)['/Analyses/' + r.corr_group + '/Events'][()]
            tag = 't%d/r%d/' % (k, j)
            blob[tag + 'events'] = ev
            meta.append({'track': k, 'chrom': key[0], 'strand': key[1], 'start': int(r.start),
                         'end': int(r.end), 'tag': tag, 'read_start_rel_to_raw': int(r.read_start_rel_to_raw)})
        blob['t%d/coverage' % k] = cov[key]
        blob['t%d/means' % k] = means[key]
        if key in sds:
            blob['t%d/sds' % k] = sds[key]
            blob['t%d/lengths' % k] = lens[key]
    blob['meta'] = np.frombuffer(json.dumps({'genome_len': GENOME_LEN, 'reads': meta,
                                             'tracks': [list(k) for k in sorted(rrc)]}).encode(), dtype=np.uint8)
    np.savez_compressed(OUT, **blob)
    print('wrote', OUT, os.path.getsize(OUT), 'bytes; tracks', sorted(rrc), 'reads', len(meta))


if __name__ == '__main__':
    main()
