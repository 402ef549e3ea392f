"""Generate tests/golden/frontend_v1.json by running the REAL reference's front-end functions
(parse_sam_record, get_read_data + fix_stay_states, fix_raw_starts_for_clipped_bases) through
oracle/run_reference.py.  Build container only:  python tests/golden/make_frontend_golden.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from nanoraw_b200 import fast5_io, synth  # noqa: E402
from oracle import run_reference as rr  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden', 'frontend_v1.json')


def events_fast5(name, read, events, version, start_time, unicode_states):
    f = fast5_io.MemFile(name, 'w')
    rd = f.create_group('Raw/Reads/Read_7')
    rd.create_dataset('Signal', data=np.asarray(read.raw, dtype=np.int16))
    rd.attrs['read_id'] = 'golden-read'
    rd.attrs['start_time'] = np.uint64(start_time)
    ch = f.create_group('UniqueGlobalKey/channel_id')
    for k, v in (('offset', 10.0), ('range', 1400.0), ('digitisation', 8192.0),
                 ('sampling_rate', 4000.0)):
        ch.attrs[k] = np.float64(v)
    ch.attrs['channel_number'] = '7'
    grp = f.create_group('Analyses/Basecall_1D_000')
    if version is not None:
        grp.attrs['version'] = version
    ev = events
    if unicode_states:
        dt = [(n, 'U5' if n == 'model_state' else t) for n, t in synth.BASECALL_EVENT_DTYPE]
        ev = events.astype(dt)
    f.create_group('Analyses/Basecall_1D_000/BaseCalled_template').create_dataset('Events', data=ev)
    f.close()
    return name


def main():
    ref, ref_nh = rr.load_reference()
    cfg = synth.SynthConfig(span_mean=300, span_sd=20, seed_base=900)
    genome = synth.make_genome(40_000, seed=21)
    levels = synth.kmer_levels(cfg)
    gi = {'synth_chr1': genome.tobytes().decode()}
    out = {'genome_seed': 21, 'genome_len': 40_000, 'sam': [], 'events': [], 'clip': []}

    # ---- parse_sam_record
    sam_cases = []
    for i in range(8):
        read = synth.make_read(genome, levels, cfg, i)
        sc = [(0, 0), (4, 0), (0, 6), (3, 5)][i % 4]
        hc = (2, 0) if i == 5 else ((0, 3) if i == 6 else (0, 0))
        sam_cases.append(synth.to_sam_record(read, soft_clip=sc, hard_clip=hc, seed=i))
    # hand-made records: leading / trailing non-match operations, '=' / 'X', 'N', 'P'
    base = synth.make_read(genome, levels, synth.SynthConfig(
        span_mean=120, span_sd=1, del_rate=0, ins_rate=0, mismatch_rate=0, seed_base=901), 0)
    rec = synth.to_sam_record(base)
    n = len(rec['seq'])
    variants = [
        '2I%dM' % (n - 2), '3D%dM' % n, '%dM2I' % (n - 2), '%dM4D' % n,
        '2S2I%dM3D' % (n - 4), '10=2X%dM' % (n - 12), '20M5N%dM' % (n - 20),
        '5M2P%dM' % (n - 7), '%dM' % n]
    for cig in variants:
        r2 = dict(rec)
        r2['cigar'] = cig
        sam_cases.append(r2)
    sam_cases.append(dict(rec, cigar='*'))
    for rec in sam_cases:
        entry = {'record': rec}
        try:
            vals, loc, cs, ce = ref.parse_sam_record(rec, gi)
            entry['read_row'] = ''.join(a for a, _ in vals)
            entry['genome_row'] = ''.join(b for _, b in vals)
            entry['loc'] = [int(loc.Start), loc.Strand, loc.Chrom]
            entry['clip'] = [int(cs), int(ce)]
        except Exception as e:  # noqa: BLE001
            entry['error'] = '%s: %s' % (type(e).__name__, e)
        out['sam'].append(entry)

    # ---- get_read_data (+ fix_stay_states)
    ev_specs = [
        dict(stay_rate=0.0), dict(stay_rate=0.15), dict(stay_rate=0.3, lead_stays=3),
        dict(stay_rate=0.2, trail_stays=4), dict(stay_rate=0.2, lead_stays=2, trail_stays=2),
        dict(stay_rate=0.1, version='0.8.4'), dict(stay_rate=0.1, version=None),
        dict(stay_rate=0.1, start_time=1), dict(stay_rate=0.1, start_time=40),
        dict(stay_rate=0.1, zero_len=True), dict(all_stays=True),
    ]
    for i, spec in enumerate(ev_specs):
        read = synth.make_read(genome, levels, cfg, 100 + i)
        ev = synth.to_basecall_events(read, stay_rate=spec.get('stay_rate', 0.1),
                                      lead_stays=spec.get('lead_stays', 0),
                                      trail_stays=spec.get('trail_stays', 0), seed=i)
        if spec.get('zero_len'):
            ev['length'][5] = 0.0
        if spec.get('all_stays'):
            ev['move'][1:] = 0
        version = spec.get('version', '1.1.0')
        start_time = spec.get('start_time', 0)
        name = events_fast5('mem://fe_golden/%d' % i, read, ev, version, start_time, True)
        entry = {'spec': {k: v for k, v in spec.items()}, 'index': 100 + i, 'seed': i,
                 'version': version, 'start_time': start_time}
        try:
            rs, starts, calls, chan, rid = ref.get_read_data(
                name, 'Basecall_1D_000', 'BaseCalled_template')
            entry['read_start'] = int(rs)
            entry['starts'] = [int(v) for v in starts]
            entry['basecalls'] = ''.join(str(c) for c in calls)
            entry['sampling_rate'] = int(chan.sampling_rate)
            entry['read_id'] = str(rid)
        except Exception as e:  # noqa: BLE001
            entry['error'] = '%s: %s' % (type(e).__name__, e)
        out['events'].append(entry)

    # ---- fix_raw_starts_for_clipped_bases
    starts = np.cumsum(np.concatenate(([0], np.arange(3, 23))))
    for cs, ce in ((0, 0), (3, 0), (0, 4), (2, 5)):
        s2, rs = ref.fix_raw_starts_for_clipped_bases(cs, ce, starts.copy(), 100)
        out['clip'].append({'clip': [cs, ce], 'starts_in': [int(v) for v in starts],
                            'starts': [int(v) for v in s2], 'read_start': int(rs)})
    with open(OUT, 'w') as fp:
        json.dump(out, fp, indent=0)
    print('wrote', OUT, os.path.getsize(OUT), 'bytes;',
          len(out['sam']), 'sam,', len(out['events']), 'events cases')
    for e in out['sam']:
        print(' sam', e['record']['cigar'][:30], e.get('error', 'ok'), e.get('clip'))
    for e in out['events']:
        print(' ev ', e['spec'], e.get('error', 'ok'))


if __name__ == '__main__':
    main()
