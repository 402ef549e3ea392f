"""Host-side logic that needs no GPU: CLI parser, process arithmetic, failed-read report,
message table, CSR packing and sharding, status -> exception mapping."""
import os
import re

import numpy as np
import pytest

from nanoraw_b200 import messages, option_parsers, pipeline, synth
from nanoraw_b200.batch import ReadBatch, shard_bounds

REF = '/root/reference/nanoraw'


def test_parser_defaults_match_reference():
    """defaults of option_parsers.py:9-136 / shell_tests.sh:30-34 style invocation"""
    p = option_parsers.get_resquiggle_parser()
    a = p.parse_args(['reads/', 'genome.fa', '--graphmap-executable', './graphmap'])
    assert (a.processes, a.align_processes, a.alignment_batch_size) == (2, 1, 500)
    assert a.normalization_type == 'median' and a.outlier_threshold == 5
    assert a.corrected_group == 'RawGenomeCorrected_000' and a.basecall_group == 'Basecall_1D_000'
    assert a.basecall_subgroups == ['BaseCalled_template']
    assert a.timeout is None and a.cpts_limit is None
    assert not (a.skip_event_stdev or a.overwrite or a.recursive or a.quiet)
    assert a.fast5_pattern is None and a.failed_reads_filename is None
    b = p.parse_args(['reads/', 'genome.fa', '--bwa-mem-executable', 'bwa', '--timeout', '60',
                      '--cpts-limit', '100', '--normalization-type', 'pA_raw', '--2d',
                      '--processes', '4', '--overwrite', '--outlier-threshold', '-1'])
    assert b.basecall_subgroups == ['BaseCalled_template', 'BaseCalled_complement']
    assert (b.timeout, b.cpts_limit, b.normalization_type, b.outlier_threshold) == (60, 100, 'pA_raw', -1.0)
    with pytest.raises(SystemExit):
        p.parse_args(['reads/', 'genome.fa', '--normalization-type', 'robust_median'])
    with pytest.raises(SystemExit):
        p.parse_args(['reads/', 'genome.fa', '--2d', '--basecall-subgroups', 'x'])


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference sources not present')
def test_parser_flags_are_the_reference_flags():
    src = open(os.path.join(REF, 'option_parsers.py')).read()
    body = src[src.index('def get_resquiggle_parser'):src.index('def get_max_cov_parser')]
    names = re.findall(r'add_argument\(\*?(\w+_opt)\[0\]', body)
    flags = set()
    for n in names:
        m = re.search(n + r"\s*=\s*\((.*?),\s*\{", src, re.S)
        flags.update(re.findall(r"'([^']+)'", m.group(1)))
    mine = set()
    for action in option_parsers.get_resquiggle_parser()._actions:
        mine.update(action.option_strings or [action.dest])
    assert flags == mine


def test_process_arithmetic():
    p = option_parsers.get_resquiggle_parser()
    a = p.parse_args(['d', 'g', '--processes', '1'])
    assert pipeline.resolve_processes(a) == (1, 1)       # num_proc = max(2, 1)
    a = p.parse_args(['d', 'g', '--processes', '16', '--align-processes', '2'])
    assert pipeline.resolve_processes(a) == (4, 8)
    a = p.parse_args(['d', 'g', '--processes', '16', '--resquiggle-processes', '3',
                      '--align-threads-per-process', '5'])
    assert pipeline.resolve_processes(a) == (5, 3)


def test_failed_reads_report_format():
    summary, table = pipeline.format_failed_reads(
        {'New segments include zero length events.': ['BaseCalled_template :: a.fast5',
                                                      'BaseCalled_template :: b.fast5'],
         'Alignment not produced.': ['BaseCalled_template:::c.fast5']})
    assert summary == ('Failed reads summary:\n'
                       '\tNew segments include zero length events. :\t2\n'
                       '\tAlignment not produced. :\t1\n')
    assert table.splitlines()[0] == ('New segments include zero length events.\t'
                                     'BaseCalled_template :: a.fast5,BaseCalled_template :: b.fast5')


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference sources not present')
def test_messages_are_verbatim_reference_text():
    ref = open(os.path.join(REF, 'resquiggle.py')).read()
    flat = re.sub(r"""['"]\s*\+\s*\\?\s*\n?\s*['"]""", '', ref)
    for key, val in vars(messages).items():
        if key.isupper() and isinstance(val, str) and key != 'BANNER':
            assert val in flat, key


def test_csr_packing_and_slicing():
    cfg = synth.SynthConfig(span_mean=150, span_sd=30, seed_base=970)
    reads = synth.make_reads(7, cfg, genome=synth.make_genome(20_000, seed=2))
    b = ReadBatch.from_reads(reads)
    assert b.n_reads == 7 and b.raw.dtype == np.int16 and b.starts.dtype == np.int32
    for r, read in enumerate(reads):
        assert np.array_equal(b.raw[b.raw_off[r]:b.raw_off[r + 1]], read.raw)
        assert np.array_equal(b.starts[b.starts_off[r]:b.starts_off[r + 1]], read.starts)
        assert b.read_align[b.align_off[r]:b.align_off[r + 1]].tobytes() == read.read_align
        assert b.n_samples[r] == read.n_samples
    s = b.slice(2, 5)
    assert s.n_reads == 3 and s.raw_off[0] == 0
    assert np.array_equal(s.raw[s.raw_off[1]:s.raw_off[2]], reads[3].raw)
    with pytest.raises(ValueError):
        ReadBatch.from_arrays([reads[0].raw], [0], [reads[0].starts], [b'AC'], [b'A'])


def test_shards_cover_all_reads_and_balance_samples():
    rng = np.random.default_rng(0)
    lens = rng.integers(1000, 200_000, size=1000)
    off = np.concatenate(([0], np.cumsum(lens)))
    for world in (1, 2, 3, 8):
        bounds = [shard_bounds(off, r, world) for r in range(world)]
        assert bounds[0][0] == 0 and bounds[-1][1] == 1000
        assert all(bounds[i][1] == bounds[i + 1][0] for i in range(world - 1))
        loads = [off[hi] - off[lo] for lo, hi in bounds]
        assert max(loads) - min(loads) <= 2 * lens.max()


def test_status_exception_mapping():
    from nanoraw_b200 import _lib
    from nanoraw_b200.csrc import build
    build.build()
    from nanoraw_b200.errors import error_string, exception_for
    assert exception_for(_lib.ST_OK) is None
    assert isinstance(exception_for(_lib.ST_SEQ_MISMATCH), ValueError)
    assert isinstance(exception_for(_lib.ST_FPE_DIVIDE), FloatingPointError)
    assert error_string(_lib.ST_CPTS_LIMIT) == \
        'RuntimeError: Reached maximum number of changepoints for a single indel'


def test_pinned_batcher_packs_like_readbatch_and_grows():
    from nanoraw_b200.batch import PinnedBatcher
    cfg = synth.SynthConfig(span_mean=150, span_sd=30, seed_base=971)
    reads = synth.make_reads(9, cfg, genome=synth.make_genome(20_000, seed=2))
    pb = PinnedBatcher(max_reads=2, raw_capacity=500, pinned=False)   # forces every buffer to grow
    for rd in reads:
        pb.add(rd.raw, rd.read_start_rel_to_raw, rd.starts, rd.read_align, rd.genome_align)
    want = ReadBatch.from_reads(reads)
    assert pb.n_reads == want.n_reads and pb.total_align == want.total_align
    for f in ReadBatch.FIELDS:
        n = len(getattr(want, f))
        assert np.array_equal(getattr(pb, f)[:n], getattr(want, f)), f
    pb.reset()
    assert pb.n_reads == 0 and pb.total_align == 0
    pb.add(reads[3].raw, 5, reads[3].starts, reads[3].read_align, reads[3].genome_align, given=(1.5, 2.0))
    assert pb.has_given and pb.scale_values_in[0] == 1.5 and np.isnan(pb.scale_values_in[2])
    with pytest.raises(ValueError):
        pb.add(reads[0].raw, 0, reads[0].starts, b'AC', b'A')


def test_worker_reports_a_device_failure_against_the_batch_it_hit(monkeypatch):
    """ADVICE (round 1): when the device batch of the PREVIOUS feed raises, exactly that batch's reads are reported,
    once each, and the batch whose file reads had just been started goes through afterwards (resquiggle.py:434-438
    per read -> per batch here).  The engine is a stand-in: no device is needed for the bookkeeping."""
    import queue
    from nanoraw_b200 import fast5_io, resquiggle as rsq, synth
    monkeypatch.setattr(rsq, 'DEVICE_BATCH_READS', 3)
    monkeypatch.setattr(rsq, '_BATCHER', [])
    cfg = synth.SynthConfig(span_mean=300, span_sd=30, seed_base=91000)
    genome = synth.make_genome(20_000, seed=5)
    names = []
    basecalls_q, failed_q = queue.Queue(), queue.Queue()
    for i in range(8):
        read = synth.make_read(genome, synth.kmer_levels(cfg), cfg, i)
        name = 'mem://failsafe/read_%d.fast5' % i
        f = fast5_io.open_fast5(name, 'w')
        rd = f.create_group('Raw/Reads/Read_1')
        rd.create_dataset('Signal', data=read.raw)
        ch = f.create_group('UniqueGlobalKey/channel_id')
        for k, v in (('offset', 1.0), ('range', 1400.0), ('digitisation', 8192.0), ('sampling_rate', 4000.0)):
            ch.attrs[k] = np.float64(v)
        ch.attrs['channel_number'] = '1'
        f.close()
        names.append(name)
        info = rsq.readInfo('r%d' % i, 'BaseCalled_template', 0, 0, 0, 0, 0, 0)
        basecalls_q.put((name, [((read.read_align, read.genome_align), rsq.genomeLoc(0, '+', 'chr'), read.starts,
                                 read.read_start_rel_to_raw, info)]))
    basecalls_q.put((None, None))

    class Result(object):
        def __init__(self, n):
            self.status = np.zeros(n, dtype=np.int32)

    class Engine(object):
        calls = 0

        def run_host(self, batcher, params, skip=()):
            Engine.calls += 1
            if Engine.calls == 2:
                raise RuntimeError('device fell over')
            return Result(batcher.n_reads)

    real = rsq.BatchPipeline

    def pipeline(*a, **kw):
        kw.update(in_place=False, engine=Engine())
        return real(*a, **kw)

    monkeypatch.setattr(rsq, 'BatchPipeline', pipeline)
    rsq.resquiggle_worker(basecalls_q, failed_q, 'Basecall_1D_000', 'RawGenomeCorrected_000', 'median', 5,
                          None, None, True, None)
    failed = [failed_q.get() for _ in range(failed_q.qsize())]
    # batches of 3, 3, 2 reads: the second one fails, whole and only
    assert Engine.calls == 3
    assert sorted(failed) == sorted(('device fell over', 'BaseCalled_template :: ' + n) for n in names[3:6])
