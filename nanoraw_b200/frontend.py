"""Front end of genome_resquiggle: what turns a basecalled FAST5 file plus a mapper's SAM record
into the inputs of the re-squiggle hot path ("next" row f-2).

Behaviourally equivalent to nanoraw/resquiggle.py:447-494 and :583-1035 (cited per function),
written independently: sequences are handled as byte arrays and the event bookkeeping is
vectorised.  The alignment itself is produced by an external mapper (graphmap / bwa mem), as
in the reference.  Alignments are carried as two equal-length strings (read row, genome row)
instead of a list of base pairs; `as_pairs` gives the reference's representation.
"""
import os
import queue
import re
from collections import defaultdict, namedtuple
from subprocess import call
from tempfile import NamedTemporaryFile

import numpy as np

from . import fast5_io, messages as msg
from . import nanoraw_helper as nh

readInfo = namedtuple('readInfo', ('ID', 'Subgroup', 'ClipStart', 'ClipEnd',
                                   'Insertions', 'Deletions', 'Matches', 'Mismatches'))
genomeLoc = namedtuple('genomeLoc', ('Start', 'Strand', 'Chrom'))

SAM_FIELDS = ('qName', 'flag', 'rName', 'pos', 'mapq', 'cigar', 'rNext', 'pNext', 'tLen',
              'seq', 'qual')
_CIGAR = re.compile(r'(\d+)([MIDNSHP=X])')
_CONSUMES_REF = frozenset('MDN=X')
_CONSUMES_READ = frozenset('MIP=X')
_ALIGNED = frozenset('M=X')
_READ_ONLY = frozenset('IP')
GAP = ord('-')


def as_pairs(rows):
    """(read row, genome row) -> [(read_base, genome_base)] as the reference builds it."""
    return list(zip(*rows))


# ------------------------------------------------------------------ basecall events -> starts
def _trim_stays(moves):
    """index range [first, last] of the events kept once leading / trailing stays are cut."""
    moved = np.flatnonzero(moves)
    # the reference walks forward from event 0 and gives up once it would reach the last state
    if len(moved) == 0 or moved[0] >= len(moves) - 1:
        raise RuntimeError(msg.ALL_STAYS)
    return int(moved[0]), int(moved[-1])


def fix_stay_states(called_dat, starts_rel_to_read, basecalls, read_start_rel_to_raw):
    """resquiggle.py:767-805.  An event whose `move` is 0 re-reports the previous k-mer; its
    boundary is dropped so that every remaining segment is one base.  Leading / trailing stay
    runs are clipped off the read entirely (their samples move into read_start_rel_to_raw)."""
    moves = np.asarray(called_dat['move'])[1:] > 0
    first, last = _trim_stays(moves)
    starts = starts_rel_to_read[first:len(starts_rel_to_read) - (len(moves) - 1 - last)]
    calls = basecalls[first:len(basecalls) - (len(moves) - 1 - last)]
    moves = moves[first:last + 1]
    if first > 0:
        read_start_rel_to_raw += starts[0]
        starts = starts - starts[0]
    # boundary j+1 survives iff event j+1 moved; the two outer boundaries always do
    keep = np.concatenate(([True], moves[:-1], [True], moves[-1:]))
    return starts[keep], calls[keep[:-1]], read_start_rel_to_raw


def _albacore_after_1_0(version):
    if isinstance(version, bytes):
        version = version.decode()
    parts = [int(p) if p.isdigit() else p for p in str(version).split('.')]
    return parts > [1, 0]


def get_read_data(fast5_fn, basecall_group, basecall_subgroup, native=True):
    """resquiggle.py:807-907 -> (read_start_rel_to_raw, starts_rel_to_read, basecalls,
    channel_info, read_id).  The event arithmetic and the stay removal (:848-898, :767-805) run in libnrqio
    (nrqio_events_to_starts); native=False is the numpy mirror the tests hold it against."""
    try:
        f5 = fast5_io.open_fast5(fast5_fn, 'r')
    except Exception:
        raise NotImplementedError(msg.OPEN_FOR_ALIGNMENT)
    try:
        group = f5['/Analyses/' + basecall_group]
        version = group.attrs['version'] if 'version' in group.attrs else '0.0'
        events = fast5_io.read_dataset(f5['/Analyses/%s/%s/Events' % (
            basecall_group, basecall_subgroup)])
    except Exception:
        raise RuntimeError(msg.NO_EVENTS)
    try:
        raw_attrs = dict(list(f5['/Raw/Reads/'].values())[0].attrs.items())
    except Exception:
        raise RuntimeError(msg.NO_RAW)
    try:
        channel = nh.get_channel_info(f5)
        f5.close()
    except Exception:
        raise RuntimeError(msg.CHANNEL_INFO)

    rate = channel.sampling_rate
    if native:
        from . import _libio
        st, starts, bases, read_start = _libio.events_to_starts(
            events, _albacore_after_1_0(version), int(rate), int(raw_attrs['start_time']))
        if st != _libio.FE_OK:
            text = _libio.fe_message(st)
            raise (RuntimeError if st == _libio.FE_ALL_STAYS else NotImplementedError)(text)
        return read_start, starts, np.frombuffer(bases, dtype='S1').astype('U1'), channel, raw_attrs['read_id']
    first_sample = np.round(events['start'][0].astype(np.float64) * rate).astype(np.uint64)
    raw_t0 = raw_attrs['start_time']
    if first_sample >= raw_t0:
        read_start = int(first_sample - raw_t0)
    elif first_sample >= raw_t0 - 2:  # float noise in the stored start time (:852-855)
        read_start = 0
    else:
        raise NotImplementedError(msg.EVENTS_BEFORE_RAW)

    if _albacore_after_1_0(version):
        # lengths are more precise than the float32 start column (:872-877)
        n_obs = np.round(events['length'] * rate).astype('int_')
        starts = np.concatenate(([0], np.cumsum(n_obs)))
    else:
        edges = np.append(events['start'], events['start'][-1] + events['length'][-1])
        starts = np.round(edges.astype(np.float64) * rate).astype('int_') - first_sample
    states = events['model_state']
    basecalls = np.array([(s.decode() if isinstance(s, bytes) else s)[2] for s in states])

    if min(len(starts), len(basecalls), len(states)) <= 1:
        raise NotImplementedError(msg.TOO_FEW_SEGMENTS)
    if np.diff(starts).min() < 1:
        raise NotImplementedError(msg.ZERO_LENGTH_INPUT)
    starts, basecalls, read_start = fix_stay_states(events, starts, basecalls, read_start)
    return read_start, starts, basecalls, channel, raw_attrs['read_id']


def get_read_data_py(fast5_fn, basecall_group, basecall_subgroup):
    """the numpy mirror of resquiggle.py:807-907 (test reference for the native path)"""
    return get_read_data(fast5_fn, basecall_group, basecall_subgroup, native=False)


# ------------------------------------------------------------------ SAM record -> alignment
def parse_sam_record_native(record, genome_index):
    """parse_sam_record through libnrqio (nrqio_sam_to_rows); same return value, same exceptions"""
    from . import _libio
    st, rr, gr, clip, _ = _libio.sam_to_rows(record['cigar'], int(record['flag']), int(record['pos']),
                                             record['seq'], genome_index[record['rName']])
    if st == _libio.FE_BAD_CIGAR:
        raise RuntimeError(msg.BAD_CIGAR)
    if st == _libio.FE_SEQ_CIGAR_DISAGREE:
        raise AssertionError(msg.SEQ_CIGAR_DISAGREE)
    if st != _libio.FE_OK:
        raise RuntimeError(_libio.fe_message(st))
    reverse = bool(int(record['flag']) & 0x10)
    return ((rr.decode(), gr.decode()), genomeLoc(int(record['pos']) - 1, '-' if reverse else '+', record['rName']),
            clip[0], clip[1])


def parse_sam_record(record, genome_index):
    """resquiggle.py:583-660 -> ((read row, genome row), genomeLoc, start_clipped, end_clipped).

    Works in read orientation: for a reverse-strand hit the CIGAR is reversed and both sequences
    reverse-complemented.  The reference adds `cigar[0][0]` (not the trailing operation's own
    length) to the end clip when it strips a trailing non-match operation (:636); that is kept.
    """
    ops = [(int(n), op) for n, op in _CIGAR.findall(record['cigar'])]
    if not ops:
        raise RuntimeError(msg.BAD_CIGAR)
    reverse = bool(int(record['flag']) & 0x10)
    read = record['seq']
    if reverse:
        ops.reverse()
        read = nh.rev_comp(read)
    clip = [0, 0]
    for side, idx in ((0, 0), (1, -1)):  # hard clips first, both ends
        if ops[idx][1] == 'H':
            clip[side] += ops.pop(idx)[0]
    if ops[0][1] == 'S':
        n = ops.pop(0)[0]
        clip[0] += n
        read = read[n:]
    if ops[-1][1] == 'S':
        n = ops.pop()[0]
        clip[1] += n
        read = read[:-n]

    pos = int(record['pos'])
    ref_len = sum(n for n, op in ops if op in _CONSUMES_REF)
    ref = genome_index[record['rName']][pos - 1:pos - 1 + ref_len]
    if reverse:
        ref = nh.rev_comp(ref)

    while ops[0][1] not in _ALIGNED:
        n, op = ops.pop(0)
        if op in _READ_ONLY:
            ref = ref[n:]
        else:
            read = read[n:]
            clip[0] += n
    while ops[-1][1] not in _ALIGNED:
        first_len = ops[0][0]
        n, op = ops.pop()
        if op in _READ_ONLY:
            ref = ref[:-n]
        else:
            read = read[:-n]
            clip[1] += first_len
    assert len(read) == sum(n for n, op in ops if op in _CONSUMES_READ), msg.SEQ_CIGAR_DISAGREE

    read_row, ref_row = [], []
    ri = gi = 0
    for n, op in ops:
        if op in _ALIGNED:
            # zip(qSeq[:n], tSeq[:n]) (:646): as many pairs as the shorter of the two has (the reference sequence
            # runs short when a leading / trailing insertion was trimmed from it, :626, :632)
            pairs = min(len(read[ri:ri + n]), len(ref[gi:gi + n]))
            read_row.append(read[ri:ri + pairs])
            ref_row.append(ref[gi:gi + pairs])
            ri += n
            gi += n
        elif op in _READ_ONLY:
            read_row.append(read[ri:ri + n])
            ref_row.append('-' * len(read[ri:ri + n]))
            ri += n
        else:
            read_row.append('-' * len(ref[gi:gi + n]))
            ref_row.append(ref[gi:gi + n])
            gi += n
    rows = (''.join(read_row), ''.join(ref_row))
    return rows, genomeLoc(pos - 1, '-' if reverse else '+', record['rName']), clip[0], clip[1]


def parse_sam_output(align_output, batch_reads_data, genome_index):
    """resquiggle.py:662-693: keep the highest-MAPQ record per read, parse each."""
    best = dict.fromkeys(batch_reads_data)
    for line in align_output:
        if isinstance(line, bytes):
            line = line.decode()
        if line.startswith('@'):
            continue
        rec = dict(zip(SAM_FIELDS, line.strip().split()))
        if len(rec) < len(SAM_FIELDS) or rec['rName'] == '*':
            continue
        name = rec['qName'].replace('|||', ' ')
        if best[name] is None or int(best[name]['mapq']) < int(rec['mapq']):
            best[name] = rec
    failed, parsed = [], {}
    for name, rec in best.items():
        if rec is None:
            failed.append((msg.NO_ALIGNMENT, name))
            continue
        try:
            parsed[name] = parse_sam_record_native(rec, genome_index)
        except Exception as e:  # noqa: BLE001 -- per-read failure, like the reference
            failed.append((str(e), name))
    return failed, parsed


def fix_raw_starts_for_clipped_bases(start_clipped_bases, end_clipped_bases,
                                     starts_rel_to_read, read_start_rel_to_raw):
    """resquiggle.py:447-460: drop the event boundaries of soft/hard-clipped bases."""
    if start_clipped_bases > 0:
        shift = starts_rel_to_read[start_clipped_bases]
        starts_rel_to_read = starts_rel_to_read[start_clipped_bases:] - shift
        read_start_rel_to_raw += shift
    if end_clipped_bases > 0:
        starts_rel_to_read = starts_rel_to_read[:-end_clipped_bases]
    return starts_rel_to_read, read_start_rel_to_raw


def alignment_tallies(rows):
    """(insertions, deletions, matches, mismatches) as resquiggle.py:476-485 counts them."""
    r = np.frombuffer(rows[0].encode(), dtype=np.uint8)
    g = np.frombuffer(rows[1].encode(), dtype=np.uint8)
    n_del = int((r == GAP).sum())
    n_ins = int(((g == GAP) & (r != GAP)).sum())
    both = (r != GAP) & (g != GAP)
    n_match = int((both & (r == g)).sum())
    return n_ins, n_del, n_match, int(both.sum()) - n_match


def fix_all_clipped_bases(batch_align_data, batch_reads_data):
    """resquiggle.py:462-494 -> [(fast5_fn, (rows, genome_loc, starts, read_start, readInfo))]:
    the items the alignment workers put on the queue for the re-squiggle workers."""
    out = []
    for name, (rows, genome_loc, clip_start, clip_end) in batch_align_data.items():
        read_start, starts, _, _, read_id = batch_reads_data[name]
        starts, read_start = fix_raw_starts_for_clipped_bases(clip_start, clip_end, starts, read_start)
        subgroup, fast5_fn = name.split(':::')
        info = readInfo(read_id, subgroup, clip_start, clip_end, *alignment_tallies(rows))
        out.append((fast5_fn, (rows, genome_loc, starts, read_start, info)))
    return out


# ------------------------------------------------------------------ external mapper
def prep_graphmap_options(genome_fn, read_fn, out_fn, output_format, num_align_ps):
    """resquiggle.py:695-698"""
    return ['align', '-r', genome_fn, '-d', read_fn, '-o', out_fn, '-L', output_format,
            '-t', str(num_align_ps)]


def prep_bwa_mem_options(genome_fn, read_fn, num_align_ps):
    """resquiggle.py:700-702"""
    return ['mem', '-x', 'ont2d', '-v', '1', '-t', str(num_align_ps), genome_fn, read_fn]


def align_to_genome(batch_reads_data, genome_fn, mapper_exe, mapper_type, genome_index,
                    num_align_ps, output_format='sam'):
    """resquiggle.py:704-765: reads -> temp FASTA -> mapper subprocess -> parsed records."""
    fasta = ''.join('>%s\n%s\n' % (name.replace(' ', '|||'), ''.join(data[2]))
                    for name, data in batch_reads_data.items())
    if mapper_type not in ('graphmap', 'bwa_mem'):
        raise RuntimeError(msg.MAPPER_UNSUPPORTED)
    with NamedTemporaryFile(suffix='.fasta', mode='w') as reads_fp, \
            NamedTemporaryFile(mode='w+') as out_fp, open(os.devnull, 'w') as devnull:
        reads_fp.write(fasta)
        reads_fp.flush()
        try:
            if mapper_type == 'graphmap':
                call([mapper_exe] + prep_graphmap_options(
                    genome_fn, reads_fp.name, out_fp.name, output_format, num_align_ps),
                    stdout=devnull, stderr=devnull)
            else:
                call([mapper_exe] + prep_bwa_mem_options(genome_fn, reads_fp.name, num_align_ps),
                     stdout=out_fp, stderr=devnull)
            out_fp.seek(0)
            sam_lines = out_fp.readlines()
        except Exception:
            return [(msg.MAPPER_FAILED, name) for name in batch_reads_data], []
    if output_format != 'sam':
        raise RuntimeError(msg.FORMAT_UNSUPPORTED)
    failed, parsed = parse_sam_output(sam_lines, batch_reads_data, genome_index)
    return failed, fix_all_clipped_bases(parsed, batch_reads_data)


def align_and_parse(fast5s_to_process, genome_fn, mapper_exe, mapper_type, genome_index,
                    basecall_group, basecall_subgroups, num_align_ps):
    """resquiggle.py:909-940"""
    reads, failed = {}, []
    for fast5_fn in fast5s_to_process:
        for subgroup in basecall_subgroups:
            name = subgroup + ':::' + fast5_fn
            try:
                reads[name] = get_read_data(fast5_fn, basecall_group, subgroup)
            except Exception as e:  # noqa: BLE001
                failed.append((str(e), name))
    align_failed, aligned = align_to_genome(
        reads, genome_fn, mapper_exe, mapper_type, genome_index, num_align_ps)
    by_file = defaultdict(list)  # both strands of a 2D read travel together (:930-934)
    for fast5_fn, item in aligned:
        by_file[fast5_fn].append(item)
    return failed + align_failed, by_file


# ------------------------------------------------------------------ FAST5 preparation
def prep_fast5(fast5_fn, basecall_group, corrected_group, overwrite, in_place):
    """resquiggle.py:942-988.  None when the file is ready, else (error, fast5_fn).  Without
    --overwrite a file that already holds the corrected group is skipped, which is what makes
    an interrupted run resumable."""
    if not in_place:
        return msg.NOT_IN_PLACE, fast5_fn
    if not fast5_fn.startswith('mem://') and not os.access(fast5_fn, os.W_OK):
        return msg.NOT_WRITABLE, fast5_fn
    try:
        with fast5_io.open_fast5(fast5_fn, 'r') as f5:
            if '/Analyses/' + basecall_group not in f5:
                return msg.NO_BASECALLS, fast5_fn
            if not overwrite and '/Analyses/' + corrected_group in f5:
                return msg.ALREADY_CORRECTED, fast5_fn
    except Exception:
        return msg.CORRUPT_FILE, fast5_fn
    try:
        with fast5_io.open_fast5(fast5_fn, 'r+') as f5:
            analyses = f5['/Analyses']
            if corrected_group in analyses:
                del analyses[corrected_group]
            grp = analyses.create_group(corrected_group)
            grp.attrs['nanoraw_version'] = nh.NANORAW_VERSION
            grp.attrs['basecall_group'] = basecall_group
    except Exception:
        return msg.CREATE_GROUP, fast5_fn
    return None


def align_reads(fast5_batch, genome_fn, mapper_exe, mapper_type, genome_index, basecall_group,
                basecall_subgroups, corrected_group, basecalls_q, overwrite, num_align_ps,
                in_place=True):
    """resquiggle.py:990-1014"""
    failed, ready = [], []
    for fast5_fn in fast5_batch:
        problem = prep_fast5(fast5_fn, basecall_group, corrected_group, overwrite, in_place)
        (ready.append(fast5_fn) if problem is None else failed.append(problem))
    align_failed, by_file = align_and_parse(
        ready, genome_fn, mapper_exe, mapper_type, genome_index, basecall_group,
        basecall_subgroups, num_align_ps)
    for fast5_fn, items in by_file.items():
        basecalls_q.put((fast5_fn, items))
    return failed + align_failed


def alignment_worker(fast5_q, basecalls_q, failed_reads_q, genome_fn, mapper_exe, mapper_type,
                     basecall_group, basecall_subgroups, corrected_group, overwrite,
                     num_align_ps):
    """resquiggle.py:1016-1035"""
    genome_index = nh.parse_fasta(genome_fn)
    while not fast5_q.empty():
        try:
            fast5_batch = fast5_q.get(block=False)
        except queue.Empty:
            break
        for failure in align_reads(
                fast5_batch, genome_fn, mapper_exe, mapper_type, genome_index, basecall_group,
                basecall_subgroups, corrected_group, basecalls_q, overwrite, num_align_ps):
            failed_reads_q.put(failure)
