"""Live check of the oracle against the real reference (build container only:
skipped where /root/reference does not exist)."""
import numpy as np
import pytest

from nanoraw_b200 import synth
from oracle import resquiggle_oracle as orc
from oracle import run_reference as rr

pytestmark = pytest.mark.skipif(not rr.reference_available(),
                                reason='reference sources not present')


@pytest.fixture(scope='module')
def refs():
    return {'default': rr.load_reference(False), 'stable': rr.load_reference(True)}


@pytest.mark.parametrize('kind', ['default', 'stable'])
@pytest.mark.parametrize('cfg_name', ['standard', 'dense', 'lowdwell'])
def test_fresh_reads(refs, kind, cfg_name):
    cfg = {
        'standard': synth.SynthConfig(span_mean=1500, span_sd=300, seed_base=31000),
        'dense': synth.SynthConfig(span_mean=1200, span_sd=200, del_rate=0.12,
                                   ins_rate=0.12, clip_bases=(10, 80), seed_base=32000),
        'lowdwell': synth.SynthConfig(span_mean=1000, span_sd=200, dwell_mean=6.3,
                                      seed_base=33000),
    }[cfg_name]
    genome = synth.make_genome(200_000, seed=3)
    levels = synth.kmer_levels(cfg)
    ref_rsq, ref_nh = refs[kind]
    for i in range(4):
        read = synth.make_read(genome, levels, cfg, i)
        norm, sv, groups = rr.run_hot_path(ref_rsq, ref_nh, read)
        out = orc.resquiggle_read(
            read.raw, read.read_start_rel_to_raw, read.starts,
            read.read_align.decode(), read.genome_align.decode(), sort_kind=kind)
        assert np.array_equal(norm, out['norm_signal'])
        assert tuple(sv) == tuple(out['scale_values'])
        assert [(g.start, g.end, list(g.cpts), list(g.indels)) for g in groups] == \
            [(g.start, g.end, list(g.cpts), list(g.indels)) for g in out['indel_groups']]
        name = rr.make_mem_fast5('mem://live/%s/%s/%d' % (kind, cfg_name, i), read)
        sub = rr.run_resquiggle_read(ref_rsq, name, read)
        ev = sub['Events'].value
        for col in ('norm_mean', 'norm_stdev', 'start', 'length', 'base'):
            assert np.array_equal(ev[col], out['events'][col])


def test_verbatim_clip_loop_same_values():
    cfg = synth.SynthConfig(span_mean=500, span_sd=10, spike_rate=0.01, seed_base=34000)
    read = synth.make_reads(1, cfg, genome=synth.make_genome(50_000, seed=5))[0]
    a, sva = orc.normalize_raw_signal(read.raw, read.read_start_rel_to_raw,
                                      read.n_samples, 'median', None, 5)
    b, svb = orc.normalize_raw_signal(read.raw, read.read_start_rel_to_raw,
                                      read.n_samples, 'median', None, 5,
                                      verbatim_clip_loop=True)
    assert np.array_equal(a, b) and tuple(sva) == tuple(svb)
