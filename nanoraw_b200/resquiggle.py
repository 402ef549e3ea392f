"""Host side of the re-squiggle hot path: the call surface of nanoraw/resquiggle.py:55-440.

Same names, argument order and meaning, error strings, queue items and RawGenomeCorrected
FAST5 layout as the reference (lines cited per function).  What changes is the body of the
re-squiggle worker: instead of one numpy pass per read it packs the reads it dequeues into
ragged CSR batches, runs them through libnrq's CUDA kernels (nanoraw_b200.engine) and scatters
the per-read results to the FAST5 writer.  Per-read failures come back as status codes and are
turned into the exceptions / strings the reference produces (nanoraw_b200.errors).  There is
no CPU implementation of the signal path: without a GPU these functions fail loudly.

The front end (FAST5 events, SAM parsing, mapper call) lives in nanoraw_b200.frontend and the
orchestration / CLI entry in nanoraw_b200.pipeline; their public names are re-exported here so
that `resquiggle.<name>` resolves as it does in the reference.  The reference's unused m5
parser (resquiggle.py:496-581, dead code: the output format is always 'sam', :706) is not
reproduced.
"""
import os
import queue
import sys
import time
from collections import namedtuple

import numpy as np

from . import fast5_io
from . import messages as msg
from . import nanoraw_helper as nh
from .errors import exception_for
from .frontend import (  # noqa: F401  (re-exported call surface)
    SAM_FIELDS, align_and_parse, align_reads, align_to_genome, alignment_worker,
    fix_all_clipped_bases, fix_raw_starts_for_clipped_bases, fix_stay_states, genomeLoc,
    get_read_data, parse_sam_output, parse_sam_record, prep_bwa_mem_options, prep_fast5,
    prep_graphmap_options, readInfo)

VERBOSE = False

# reads per device batch in the re-squiggle worker (additive knob, not a reference option)
DEVICE_BATCH_READS = int(os.environ.get('NANORAW_B200_BATCH_READS', '2048'))

indelStats = namedtuple('indelStats', ('start', 'end', 'diff'))  # :29-30
indelGroupStats = namedtuple('indelGroupStats', ('start', 'end', 'cpts', 'indels'))  # :31-32

EVENT_DTYPE = [('norm_mean', '<f8'), ('norm_stdev', '<f8'),
               ('start', '<u4'), ('length', '<u4'), ('base', 'S1')]  # :69-70



def _rows(alignVals):
    """alignVals -> (read row, genome row) as str.  Accepts the reference's list of
    (read_base, genome_base) pairs or an already split (read_align, genome_align) tuple."""
    if isinstance(alignVals, tuple) and len(alignVals) == 2 and not isinstance(alignVals[0], tuple):
        ra, ga = alignVals
        if isinstance(ra, bytes):
            ra, ga = ra.decode(), ga.decode()
        return ra, ga
    return ''.join(r for r, _ in alignVals), ''.join(g for _, g in alignVals)


#################################################
########## Raw Signal Re-squiggle Code ##########
#################################################

def write_new_fast5_group(
        filename, genome_location, read_info,
        read_start_rel_to_raw, new_segs, align_seq, alignVals,
        old_segs, norm_signal, scale_values, corrected_group,
        basecall_subgroup, norm_type, outlier_thresh, compute_sd, event_data=None):
    """resquiggle.py:55-137.  `event_data` (additive) lets the batch worker hand over the event
    table the device already produced; when absent the per-base statistics of `norm_signal` are
    computed on the GPU (nrq_event_table, numpy summation order)."""
    read_align, genome_align = _rows(alignVals)
    try:
        if event_data is None:
            from .engine import get_engine
            means, sds = get_engine().event_stats(norm_signal, new_segs, compute_sd)
            event_data = np.empty(len(new_segs) - 1, dtype=EVENT_DTYPE)
            event_data['norm_mean'] = means
            event_data['norm_stdev'] = sds
            event_data['start'] = new_segs[:-1]
            event_data['length'] = np.diff(new_segs)
            event_data['base'] = np.frombuffer(align_seq.encode(), dtype='S1')
        np_read_align = np.frombuffer(read_align.encode(), dtype='S1')
        np_genome_align = np.frombuffer(genome_align.encode(), dtype='S1')
    except Exception:
        raise NotImplementedError(msg.EVENTS)

    try:
        read_data = fast5_io.open_fast5(filename, 'r+')
    except Exception:
        raise NotImplementedError(msg.OPEN_FOR_WRITE)

    # layout: docs/resquiggle.rst:11-55, resquiggle.py:89-124
    subgroup_attrs = (
        ('shift', scale_values.shift), ('scale', scale_values.scale),
        ('lower_lim', scale_values.lower_lim), ('upper_lim', scale_values.upper_lim),
        ('norm_type', norm_type), ('outlier_threshold', outlier_thresh))
    alignment_attrs = (
        ('mapped_start', genome_location.Start), ('mapped_strand', genome_location.Strand),
        ('mapped_chrom', genome_location.Chrom),
        ('clipped_bases_start', read_info.ClipStart), ('clipped_bases_end', read_info.ClipEnd),
        ('num_insertions', read_info.Insertions), ('num_deletions', read_info.Deletions),
        ('num_matches', read_info.Matches), ('num_mismatches', read_info.Mismatches))
    # read_segments keeps the basecaller's segmentation so the correction can be plotted
    alignment_datasets = (('read_alignment', np_read_align),
                          ('genome_alignment', np_genome_align), ('read_segments', old_segs))
    try:
        subgroup = read_data['/Analyses/' + corrected_group].create_group(basecall_subgroup)
        for key, value in subgroup_attrs:
            subgroup.attrs[key] = value
        alignment = subgroup.create_group('Alignment')
        for key, value in alignment_attrs:
            alignment.attrs[key] = value
        for name, data in alignment_datasets:
            alignment.create_dataset(name, data=data, compression='gzip')
        events = subgroup.create_dataset('Events', data=event_data, compression='gzip')
        events.attrs['read_start_rel_to_raw'] = read_start_rel_to_raw
    except Exception:
        raise NotImplementedError(msg.WRITE_BACK)

    try:
        read_data.flush()
        read_data.close()
    except Exception:
        raise NotImplementedError(msg.CLOSE_AFTER_WRITE)
    return


def write_packed_fast5_group(
        filename, genome_location, read_info, read_start_rel_to_raw, packed, scale_values, corrected_group,
        basecall_subgroup, norm_type, outlier_thresh):
    """write_new_fast5_group (resquiggle.py:80-135) for a read whose four datasets already exist as gzip chunks
    (libnrqio, nrqio_pack_corrected): `packed` = (n_bases, n_align, n_segs, Events chunk, read_alignment chunk,
    genome_alignment chunk, read_segments chunk).  Same groups, attributes, dataset names, dtypes and filter as
    the reference writes; same error messages."""
    n_bases, n_align, n_segs, ev_chunk, ra_chunk, ga_chunk, seg_chunk = packed
    try:
        read_data = fast5_io.open_fast5(filename, 'r+')
    except Exception:
        raise NotImplementedError(msg.OPEN_FOR_WRITE)
    try:
        subgroup = read_data['/Analyses/' + corrected_group].create_group(basecall_subgroup)
        for key, value in (('shift', scale_values.shift), ('scale', scale_values.scale),
                           ('lower_lim', scale_values.lower_lim), ('upper_lim', scale_values.upper_lim),
                           ('norm_type', norm_type), ('outlier_threshold', outlier_thresh)):
            subgroup.attrs[key] = value
        alignment = subgroup.create_group('Alignment')
        for key, value in (
                ('mapped_start', genome_location.Start), ('mapped_strand', genome_location.Strand),
                ('mapped_chrom', genome_location.Chrom),
                ('clipped_bases_start', read_info.ClipStart), ('clipped_bases_end', read_info.ClipEnd),
                ('num_insertions', read_info.Insertions), ('num_deletions', read_info.Deletions),
                ('num_matches', read_info.Matches), ('num_mismatches', read_info.Mismatches)):
            alignment.attrs[key] = value
        fast5_io.create_packed_dataset(alignment, 'read_alignment', (n_align,), 'S1', ra_chunk)
        fast5_io.create_packed_dataset(alignment, 'genome_alignment', (n_align,), 'S1', ga_chunk)
        fast5_io.create_packed_dataset(alignment, 'read_segments', (n_segs,), '<i8', seg_chunk)
        events = fast5_io.create_packed_dataset(subgroup, 'Events', (n_bases,), np.dtype(EVENT_DTYPE), ev_chunk)
        events.attrs['read_start_rel_to_raw'] = read_start_rel_to_raw
    except Exception:
        raise NotImplementedError(msg.WRITE_BACK)
    try:
        read_data.flush()
        read_data.close()
    except Exception:
        raise NotImplementedError(msg.CLOSE_AFTER_WRITE)
    return


def get_indel_groups(
        alignVals, align_segs, raw_signal, min_seg_len, timeout,
        num_cpts_limit):
    """resquiggle.py:139-316 on the GPU (nrq_find_indels + nrq_resegment).  `raw_signal` is the
    normalised float64 signal, as in the reference.  Change-point ties are visited in the
    deterministic order (value desc, index desc); see DESIGN.md."""
    from .engine import get_engine
    read_align, genome_align = _rows(alignVals)
    status, groups = get_engine().indel_groups(
        read_align.encode(), genome_align.encode(), np.asarray(align_segs),
        raw_signal, min_seg_len, timeout, num_cpts_limit)
    exc = exception_for(status)
    if exc is not None:
        raise exc
    return [indelGroupStats(s, e, cpts, [indelStats(*i) for i in indels])
            for s, e, cpts, indels in groups]


def running_differences(norm_signal, min_seg_len=4):
    """The trace `plot_correction` draws under the signal (plot_commands.py:225-229): the same running
    difference get_cpts ranks (resquiggle.py:239-243), for a window of normalised signal."""
    from .engine import get_engine
    return get_engine().running_diffs(norm_signal, min_seg_len)


def _read_raw_inputs(fast5_fn, norm_type, basecall_group, subgroup):
    """what resquiggle_read pulls out of the FAST5 file (resquiggle.py:327-347)."""
    try:
        fast5_data = fast5_io.open_fast5(fast5_fn, 'r')
        channel_info = nh.get_channel_info(fast5_data)
        raw_read = list(fast5_data['/Raw/Reads/'].values())[0]
        all_raw_signal = fast5_io.read_dataset(raw_read['Signal'])
        event_means, event_kmers = None, None
        if norm_type == 'pA':
            event_data = fast5_io.read_dataset(fast5_data[
                '/Analyses/' + basecall_group + '/' + subgroup + '/Events'])
            event_means = event_data['mean']
            event_kmers = event_data['model_state']
        fast5_data.close()
    except Exception:
        raise NotImplementedError(msg.OPEN_FOR_RESQUIGGLE)
    return channel_info, all_raw_signal, event_means, event_kmers


_BATCHER = []


def _batcher(which=0):
    """the worker process's pinned staging buffers (grow-only, reused for every device batch).  The native FAST5
    path fills batch n + 1 while batch n is on the device, so it alternates between two of them."""
    from .batch import PinnedBatcher
    while len(_BATCHER) <= which:
        _BATCHER.append(PinnedBatcher(max_reads=DEVICE_BATCH_READS))
    return _BATCHER[which]


_IO_THREAD = []


def _io_thread():
    """one background thread per worker process: it sits in libnrqio's batch file functions (which run on their own
    threads and hold no GIL) while the main thread drives the device"""
    if not _IO_THREAD:
        from concurrent.futures import ThreadPoolExecutor
        _IO_THREAD.append(ThreadPoolExecutor(max_workers=2, thread_name_prefix='nrq-fast5'))
    return _IO_THREAD[0]


class _Later(object):
    """results of a background call, read like the iterators fast5_jobs.start_jobs returns"""

    def __init__(self, future):
        self._future = future

    def __iter__(self):
        return iter(self._future.result())


class _LoadedBatch(object):
    """a device batch whose inputs libnrqio has put into a PinnedBatcher (BatchPipeline._load_native)"""
    __slots__ = ('ready', 'failed', 'unreadable', 'batcher')

    def __init__(self, ready, failed, unreadable, batcher):
        self.ready, self.failed, self.unreadable, self.batcher = ready, failed, unreadable, batcher


class _PendingRead(object):
    __slots__ = ('fast5_fn', 'alignVals', 'genome_loc', 'starts', 'read_start', 'read_info',
                 'raw', 'given', 'rows')

    def __init__(self, fast5_fn, align_data):
        (self.alignVals, self.genome_loc, self.starts, self.read_start,
         self.read_info) = align_data
        self.fast5_fn = fast5_fn
        self.raw = None
        self.given = None
        # the two alignment rows as bytes, which is what the device batch takes (the reference's list of pairs and
        # str rows are converted once, here)
        av = self.alignVals
        if isinstance(av, tuple) and len(av) == 2 and isinstance(av[0], bytes) and isinstance(av[1], bytes):
            self.rows = av
        else:
            ra, ga = _rows(av)
            self.rows = (ra.encode(), ga.encode())


class BatchPipeline(object):
    """The batched body of resquiggle_read (resquiggle.py:318-401) as a three-deep pipeline over device batches:
    while the FAST5 reads of batch n + 1 run on the I/O processes, batch n goes through the device and its
    RawGenomeCorrected write-backs are queued behind them, and batch n - 1's write-backs are being collected.
    `feed(pending)` takes the next list of _PendingRead and returns the failures [(error string, read)] that have
    become known; `drain()` finishes what is in flight.  A file is never handled by two jobs at once (the
    one-writer-per-file rule, :422-424): a batch that shares a file with unfinished work waits for it."""

    def __init__(self, norm_type, outlier_thresh, timeout, num_cpts_limit, basecall_group, corrected_group,
                 compute_sd, pore_model, min_event_obs=4, in_place=True, engine=None):
        self.norm_type, self.outlier_thresh = norm_type, outlier_thresh
        self.timeout, self.num_cpts_limit = timeout, num_cpts_limit
        self.basecall_group, self.corrected_group = basecall_group, corrected_group
        self.compute_sd, self.pore_model = compute_sd, pore_model
        self.min_event_obs, self.in_place, self.engine = min_event_obs, in_place, engine
        self._reading = None   # (pending, iterator over read results)
        self._writing = []     # [(reads being written, iterator over write results, file names)]
        self._n_batches = 0
        # wall seconds spent in each stage of the native file path (libnrqio), summed over batches: with
        # NRQ_F5_THREADS=1 and NRQ_PACK_THREADS=1 these are the single-core costs per read that bound the worker
        self.stage_seconds = {'open': 0.0, 'fill': 0.0, 'inflate': 0.0, 'device': 0.0, 'pack': 0.0, 'append': 0.0}

    def _native(self, pending):
        """file work of this batch through libnrqio (nrq_io.h: nrqio_f5_*) rather than hdf5_min.py on the I/O
        processes: files on disk, write-back wanted, and a normalisation that needs no basecall events"""
        return (self.in_place and self.norm_type in ('median', 'none', 'pA_raw')
                and os.environ.get('NRQ_FAST5_NATIVE', '1') != '0'
                and not any(str(pr.fast5_fn).startswith('mem://') for pr in pending))

    # ---- stage 1: FAST5 reads of a batch, on the I/O processes (or libnrqio's threads)
    def _start_reads(self, pending):
        from . import fast5_jobs
        if self._native(pending):
            self._n_batches += 1
            return pending, _Later(_io_thread().submit(self._load_native, pending, _batcher(self._n_batches & 1)))
        jobs = [(pr.fast5_fn, self.norm_type, self.basecall_group, pr.read_info.Subgroup) for pr in pending]
        return pending, fast5_jobs.start_jobs(fast5_jobs.read_job, jobs)

    def _load_native(self, pending, batcher):
        """resquiggle.py:327-347 for a whole batch: libnrqio opens the files, finds the signals and inflates them
        straight into the pinned raw array of `batcher`; the other CSR arrays are filled from the alignment data.
        A file outside libnrqio's HDF5 slice is read by hdf5_min.py instead; a read that cannot be loaded fails with
        the message the reference gives."""
        from . import _libio, fast5_jobs
        n = len(pending)
        t_0 = time.perf_counter()
        fb = _libio.F5RawBatch([pr.fast5_fn for pr in pending])
        self.stage_seconds['open'] += time.perf_counter() - t_0
        try:
            lens = fb.n_samples.copy()
            channel = fb.channel.copy()
            slow, failed = {}, []

            def python_read(i):
                got = fast5_jobs.read_job((pending[i].fast5_fn, self.norm_type, self.basecall_group,
                                           pending[i].read_info.Subgroup))
                if got[0] != 'ok':
                    return got[1]
                chan, raw = got[1], np.ascontiguousarray(got[2], dtype=np.int16)
                slow[i] = raw
                lens[i] = len(raw)
                channel[i] = (chan.offset, chan.range, chan.digitisation, chan.sampling_rate)
                return None

            keep = np.ones(n, dtype=bool)
            for i in np.nonzero(fb.status != _libio.F5_OK)[0]:
                err = python_read(int(i))
                if err is not None:
                    failed.append((err, pending[i]))
                    keep[i] = False
            for i, pr in enumerate(pending):  # what PinnedBatcher.add refuses, per read
                if keep[i] and (len(pr.rows[0]) != len(pr.rows[1]) or
                                (len(pr.starts) and int(pr.starts[-1]) > np.iinfo(np.int32).max)):
                    failed.append(('alignment rows differ in length' if len(pr.rows[0]) != len(pr.rows[1])
                                   else 'read longer than 2^31 samples', pr))
                    keep[i] = False
            idx = np.nonzero(keep)[0]
            ready = [pending[i] for i in idx]
            given = None
            if self.norm_type != 'median':
                given = [nh.given_shift_scale(self.norm_type, nh.channelInfo(
                    channel[i][0], channel[i][1], channel[i][2], None, int(channel[i][3]))) for i in idx]
            t_0 = time.perf_counter()
            raw_off = batcher.fill(
                lens[idx], [int(pr.read_start) for pr in ready], [pr.starts for pr in ready],
                [pr.rows[0] for pr in ready], [pr.rows[1] for pr in ready], given)
            dst_off = np.full(n, -1, dtype=np.int64)
            dst_off[idx] = raw_off[:-1]
            for i in slow:
                dst_off[i] = -1
            t_1 = time.perf_counter()
            status = fb.read_into(batcher.raw, dst_off)
            self.stage_seconds['fill'] += t_1 - t_0
            self.stage_seconds['inflate'] += time.perf_counter() - t_1
        finally:
            fb.close()
        unreadable = {}
        for k, i in enumerate(idx):
            lo, hi = int(raw_off[k]), int(raw_off[k + 1])
            if i not in slow and status[i] != _libio.F5_OK:  # e.g. a filter libnrqio does not undo: hdf5_min's turn
                err = python_read(int(i))
                if err is None and len(slow[i]) != hi - lo:
                    err = msg.OPEN_FOR_RESQUIGGLE
                if err is not None:
                    slow.pop(i, None)
                    batcher.raw[lo:hi] = 0
                    unreadable[k] = err
            if i in slow:
                batcher.raw[lo:hi] = slow[i]
        if self.norm_type != 'median':
            for pr, g in zip(ready, given):
                pr.given = g
        return [_LoadedBatch(ready, failed, unreadable, batcher)]

    # ---- stage 2: one device batch (a1..a7), then the write-backs are queued
    def _process(self, pending, loaded):
        from . import fast5_jobs
        from .engine import ResquiggleParams, get_engine
        if loaded and isinstance(loaded[0], _LoadedBatch):
            return self._process_native(loaded[0])
        norm_type = self.norm_type
        failed, ready = [], []
        for pr, got in zip(pending, loaded):
            try:
                if got[0] != 'ok':
                    raise NotImplementedError(got[1])
                chan, raw, ev_means, ev_kmers = got[1:]
                pr.raw = np.ascontiguousarray(raw, dtype=np.int16)
                pr.given = nh.given_shift_scale(norm_type, chan, self.pore_model, ev_means, ev_kmers)
                ready.append(pr)
            except Exception as e:  # noqa: BLE001 -- mirrors the worker's blanket except (:434)
                failed.append((str(e), pr))
        if not ready:
            return failed
        engine = self.engine or get_engine()
        batcher = _batcher()
        batcher.reset()
        for pr in ready:  # straight into pinned CSR staging: no intermediate copies
            batcher.add(pr.raw, int(pr.read_start), pr.starts, pr.rows[0], pr.rows[1],
                        None if norm_type == 'median' else (float(pr.given[0]), float(pr.given[1])))
        params = ResquiggleParams(
            norm_type='median' if norm_type == 'median' else 'given', outlier_thresh=self.outlier_thresh,
            compute_sd=self.compute_sd, num_cpts_limit=self.num_cpts_limit, timeout=self.timeout,
            min_event_obs=self.min_event_obs)
        # new_segs / ev_start (both recoverable from ev_length) and the accounting arrays stay on the device
        res = engine.run_host(batcher, params, skip=('new_segs', 'ev_start', 'region_samples', 'n_indels')
                              if self.in_place else ())
        # the four datasets of every read as gzip chunks, straight from the pinned result arrays (libnrqio threads):
        # no per-read event tables are built on this side, and the I/O processes neither deflate nor see numpy arrays
        chunks = off = ln = None
        if self.in_place:
            from . import _libio
            chunks, off, ln = _libio.pack_corrected(res.c_batch, res.c_results, len(ready))
        writes, writers = [], []
        for r, pr in enumerate(ready):
            pr.raw = None
            exc = exception_for(res.status[r])
            if exc is not None:
                failed.append((str(exc), pr))
                continue
            if not self.in_place:
                continue
            try:
                sv = res.scale_values_of(r)
                if norm_type != 'median':
                    sv = (pr.given[0], pr.given[1], sv[2], sv[3])
                packed = (int(res.n_bases[r]), len(pr.rows[0]), len(pr.starts)) + tuple(
                    chunks[off[4 * r + j]:off[4 * r + j] + ln[4 * r + j]].tobytes() for j in range(4))
                writes.append((
                    pr.fast5_fn, pr.genome_loc, pr.read_info, pr.read_start, packed,
                    nh.scaleValues(*sv), self.corrected_group, pr.read_info.Subgroup, norm_type,
                    self.outlier_thresh))
                writers.append(pr)
            except Exception as e:  # noqa: BLE001
                failed.append((str(e), pr))
        if writes:
            self._writing.append((writers, fast5_jobs.start_jobs(fast5_jobs.write_packed_job, writes),
                                  set(pr.fast5_fn for pr in writers)))
        return failed

    def _process_native(self, lb):
        """the device batch of a _LoadedBatch, then its write-backs: libnrqio packs the gzip chunks from the pinned
        result arrays and appends the RawGenomeCorrected groups to the files (nrqio_f5_write_corrected_batch) on
        the background thread, while the next batch goes through the device"""
        from . import _libio
        from .engine import ResquiggleParams, get_engine
        failed, ready, batcher = list(lb.failed), lb.ready, lb.batcher
        if not ready:
            return failed
        norm_type = self.norm_type
        engine = self.engine or get_engine()
        params = ResquiggleParams(
            norm_type='median' if norm_type == 'median' else 'given', outlier_thresh=self.outlier_thresh,
            compute_sd=self.compute_sd, num_cpts_limit=self.num_cpts_limit, timeout=self.timeout,
            min_event_obs=self.min_event_obs)
        t_0 = time.perf_counter()
        res = engine.run_host(batcher, params, skip=('new_segs', 'ev_start', 'region_samples', 'n_indels'))
        t_1 = time.perf_counter()
        chunks, off, ln = _libio.pack_corrected(res.c_batch, res.c_results, len(ready))
        self.stage_seconds['device'] += t_1 - t_0
        self.stage_seconds['pack'] += time.perf_counter() - t_1
        good = np.asarray(res.status[:len(ready)]) == 0
        for r in lb.unreadable:
            good[r] = False
            failed.append((lb.unreadable[r], ready[r]))
        for r in np.flatnonzero(np.asarray(res.status[:len(ready)]) != 0):
            if int(r) not in lb.unreadable:
                failed.append((str(exception_for(res.status[r])), ready[r]))
        sel = np.flatnonzero(good)
        if not len(sel):
            return failed
        writers = [ready[r] for r in sel]
        sv = np.array(res.scale_values, dtype=np.float64).reshape(-1, 4)[sel]
        if norm_type != 'median':
            sv[:, :2] = np.asarray([pr.given for pr in writers], dtype=np.float64)
        # attributes go out as int64 / float64 / str from libnrqio; anything else (numpy ints of another width, a
        # missing clip limit, ...) is typed by the Python writer's h5py rules, as the reference's would be
        def plain_int(v):
            return type(v) is int or isinstance(v, np.int64)

        ints = np.zeros((len(writers), 8), dtype=np.int64)
        odd = np.zeros(len(writers), dtype=np.int32)
        no_limits = np.isnan(sv).any(axis=1)
        subgroup = writers[0].read_info.Subgroup
        for k, pr in enumerate(writers):
            vals = (pr.genome_loc.Start, pr.read_info.ClipStart, pr.read_info.ClipEnd, pr.read_info.Insertions,
                    pr.read_info.Deletions, pr.read_info.Matches, pr.read_info.Mismatches, pr.read_start)
            try:
                if not all(plain_int(v) for v in vals) or no_limits[k] or \
                        not isinstance(pr.genome_loc.Strand, str) or not isinstance(pr.genome_loc.Chrom, str) or \
                        pr.read_info.Subgroup != subgroup:
                    odd[k] = 1
                else:
                    ints[k] = vals
            except (TypeError, ValueError, OverflowError):
                odd[k] = 1
        if norm_type == 'none' or type(self.outlier_thresh) not in (int, float):
            odd[:] = 1
        job = dict(
            writers=writers, sv=sv, ints=ints, odd=odd, n_bases=np.array(res.n_bases, dtype=np.int32)[sel],
            n_align=np.diff(np.asarray(res.align_off))[sel], n_segs=np.diff(np.asarray(batcher.starts_off[:len(ready) + 1]))[sel],
            chunks=chunks, off=off.reshape(-1, 4)[sel].copy(), ln=ln.reshape(-1, 4)[sel].copy())
        self._writing.append((writers, _Later(_io_thread().submit(self._write_native, job)),
                              set(pr.fast5_fn for pr in writers)))
        return failed

    def _write_native(self, job):
        """-> [None | error message] per writer (resquiggle.py:80-135 and its error strings)"""
        from . import _libio, fast5_jobs
        writers, sv, chunks, off, ln = job['writers'], job['sv'], job['chunks'], job['off'], job['ln']
        odd = job['odd']
        t_0 = time.perf_counter()
        if not odd.all():
            status = _libio.f5_write_corrected(
                [pr.fast5_fn for pr in writers], self.corrected_group, writers[0].read_info.Subgroup, self.norm_type,
                self.outlier_thresh, [str(pr.genome_loc.Chrom) for pr in writers],
                [str(pr.genome_loc.Strand) for pr in writers], sv, job['ints'], job['n_bases'], job['n_align'],
                job['n_segs'], chunks, off, ln, skip=odd)
        else:
            status = np.zeros(len(writers), dtype=np.int32)
        self.stage_seconds['append'] += time.perf_counter() - t_0
        errs = []
        for k, pr in enumerate(writers):
            st = int(status[k])
            if st == _libio.F5_OK and not odd[k]:
                errs.append(None)
            elif st == _libio.F5_EXISTS:
                errs.append(msg.WRITE_BACK)
            elif st == _libio.F5_IO:
                errs.append(msg.OPEN_FOR_WRITE)
            else:  # outside libnrqio's HDF5 slice (or another subgroup name): hdf5_min.py rewrites the file
                packed = (int(job['n_bases'][k]), int(job['n_align'][k]), int(job['n_segs'][k])) + tuple(
                    chunks[off[k, j]:off[k, j] + ln[k, j]].tobytes() for j in range(4))
                errs.append(fast5_jobs.write_packed_job((
                    pr.fast5_fn, pr.genome_loc, pr.read_info, pr.read_start, packed, nh.scaleValues(*sv[k]),
                    self.corrected_group, pr.read_info.Subgroup, self.norm_type, self.outlier_thresh)))
        return errs

    # ---- stage 3: results of the write-backs
    def _collect_writes(self, keep=0):
        failed = []
        while len(self._writing) > keep:
            writers, results, _ = self._writing.pop(0)
            failed += [(err, pr) for pr, err in zip(writers, results) if err is not None]
        return failed

    def _advance(self):
        """push the batch whose reads are in flight through the device.  An infrastructure failure (device error,
        I/O pool failure) fails exactly the reads of THAT batch, as the worker's blanket except does for one read
        (:434-438); the batch whose file reads have just been started is not touched by it."""
        failed = []
        if self._reading is not None:
            pending, loaded = self._reading
            self._reading = None
            try:
                got = self._process(pending, list(loaded))
                failed += got
            except Exception as e:  # noqa: BLE001
                # reads of this batch whose write-backs were already queued are reported by _collect_writes
                queued = set(id(pr) for writers, _, _ in self._writing for pr in writers)
                failed += [(str(e), pr) for pr in pending if id(pr) not in queued]
        return failed

    def feed(self, pending):
        failed = []
        files = set(pr.fast5_fn for pr in pending)
        busy = set(pr.fast5_fn for pr in self._reading[0]) if self._reading else set()
        for _, _, names in self._writing:
            busy |= names
        if files & busy:  # (only with several subgroups per file)
            failed += self.drain()
        nxt = self._start_reads(pending)      # queued behind the write-backs already in flight
        failed += self._advance()             # the previous batch: device, then its write-backs are queued
        failed += self._collect_writes(keep=1)
        self._reading = nxt
        return failed

    def drain(self):
        return self._advance() + self._collect_writes()


def resquiggle_reads_batch(
        pending, norm_type, outlier_thresh, timeout, num_cpts_limit,
        basecall_group, corrected_group, compute_sd, pore_model,
        min_event_obs=4, in_place=True, engine=None):
    """One device batch, start to finish: FAST5 read -> device (a1..a7) -> FAST5 write for a list of
    _PendingRead.  Returns [(error string, read)] for the reads that failed, with the reference's messages."""
    pipe = BatchPipeline(norm_type, outlier_thresh, timeout, num_cpts_limit, basecall_group, corrected_group,
                         compute_sd, pore_model, min_event_obs, in_place, engine)
    return pipe.feed(pending) + pipe.drain()


def resquiggle_read(
        fast5_fn, read_start_rel_to_raw, starts_rel_to_read,
        norm_type, outlier_thresh, alignVals,
        timeout, num_cpts_limit, genome_loc, read_info,
        basecall_group, corrected_group, compute_sd, pore_model,
        min_event_obs=4, in_place=True):
    """resquiggle.py:318-401 for one read (a device batch of one).  Raises what the reference
    raises."""
    pr = _PendingRead(fast5_fn, (alignVals, genome_loc, starts_rel_to_read,
                                 read_start_rel_to_raw, read_info))
    failed = resquiggle_reads_batch(
        [pr], norm_type, outlier_thresh, timeout, num_cpts_limit, basecall_group,
        corrected_group, compute_sd, pore_model, min_event_obs, in_place)
    if failed:
        raise _exception_from_message(failed[0][0])
    return


def _exception_from_message(text):
    """rebuild the reference's exception type from its message (types per resquiggle.py)."""
    from . import _lib
    from .errors import STATUS_EXCEPTIONS
    for code, cls in STATUS_EXCEPTIONS.items():
        if _lib.status_message(code) == text:
            return cls(text)
    return NotImplementedError(text)


def resquiggle_worker(
        basecalls_q, failed_reads_q, basecall_group, corrected_group,
        norm_type, outlier_thresh, timeout, num_cpts_limit, compute_sd,
        pore_model):
    """resquiggle.py:403-440 with a batching body: queue items (fast5_fn, [align_data per
    subgroup]) are gathered until DEVICE_BATCH_READS reads are pending, then processed as one
    device batch.  Subgroups of one file are never in the same batch, so a file is never open
    twice (the reference's one-writer-per-file rule, :422-424)."""
    num_processed = 0
    done = False
    pending = []
    carry = []  # later subgroups of files already seen, for following batches
    pipe = BatchPipeline(norm_type, outlier_thresh, timeout, num_cpts_limit, basecall_group, corrected_group,
                         compute_sd, pore_model)
    while not done or carry:
        pending, carry_next = [], []
        seen_files = set()
        for pr in carry:
            (carry_next if pr.fast5_fn in seen_files else pending).append(pr)
            seen_files.add(pr.fast5_fn)
        while not done and len(pending) < DEVICE_BATCH_READS:
            try:
                fast5_fn, sgs_align_data = basecalls_q.get(timeout=1)
            except queue.Empty:
                if pending:
                    break
                continue
            # None values placed in queue when all files have been processed
            if fast5_fn is None:
                done = True
                break
            num_processed += 1
            if VERBOSE and num_processed % 100 == 0:
                sys.stderr.write('.')
                sys.stderr.flush()
            for i, align_data in enumerate(sgs_align_data):
                pr = _PendingRead(fast5_fn, align_data)
                (pending if i == 0 else carry_next).append(pr)
        carry = carry_next
        if not pending:
            continue
        # device / pool failures are turned into per-read failures of the batch they hit inside the pipeline
        # (BatchPipeline._advance); what can still raise here is starting the file reads of `pending` itself
        try:
            failed = pipe.feed(pending)
        except Exception as e:  # noqa: BLE001
            failed = [(str(e), pr) for pr in pending]
        for text, pr in failed:
            failed_reads_q.put((text, pr.read_info.Subgroup + ' :: ' + pr.fast5_fn))
    failed = pipe.drain()
    for text, pr in failed:
        failed_reads_q.put((text, pr.read_info.Subgroup + ' :: ' + pr.fast5_fn))
    return




# orchestration and CLI entry (resquiggle.py:1037-1203) live in nanoraw_b200.pipeline
from .pipeline import (  # noqa: E402,F401
    ALIGN_BATCH_MULTIPLIER, args_and_main, resquiggle_all_reads, resquiggle_main)

if __name__ == '__main__':
    args_and_main()
