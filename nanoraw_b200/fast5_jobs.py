"""FAST5 reads and write-backs of one device batch, spread over host processes.

The reference opens each FAST5 file inside its re-squiggle worker, one read at a time
(resquiggle.py:328-347 read, :80-135 write).  Here a GPU worker handles hundreds of reads per
device batch (about 10 ms of kernels), so the file work of a batch -- HDF5 parsing, inflate of the
raw signal, deflate of the new event table, about 12 ms per read -- is what bounds the worker; it
runs on a pool of I/O processes, and resquiggle.BatchPipeline keeps that pool busy: the reads of the
next batch are queued before the current batch goes through the device.  Nothing GPU-related is
imported in those processes.

NRQ_IO_PROCS sets the pool size (default: the host cores, at most 32; 0 or 1 = do the work inline).
One file is only ever handled by one job at a time (the one-writer-per-file rule, :422-424):
a batch never holds two subgroups of the same file.
"""
import multiprocessing as mp
import os
from concurrent.futures import ProcessPoolExecutor

_POOL = []


def io_procs():
    env = os.environ.get('NRQ_IO_PROCS')
    if env is not None:
        return max(0, int(env))
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        cores = os.cpu_count() or 1
    return min(32, cores)


def _init_worker(backend):
    from . import fast5_io
    fast5_io._BACKEND[0] = backend


def pool():
    """the process pool, or None when the work is to be done inline"""
    n = io_procs()
    if n <= 1:
        return None
    if not _POOL:
        from . import fast5_io
        _POOL.append(ProcessPoolExecutor(max_workers=n, mp_context=mp.get_context('spawn'),
                                         initializer=_init_worker, initargs=(fast5_io._BACKEND[0],)))
    return _POOL[0]


def shutdown():
    while _POOL:
        _POOL.pop().shutdown(wait=True)


def read_job(args):
    """-> ('ok', channel_info, raw int16, event means, event k-mers) | ('err', message)"""
    fast5_fn, norm_type, basecall_group, subgroup = args
    from . import resquiggle
    try:
        return ('ok',) + tuple(resquiggle._read_raw_inputs(fast5_fn, norm_type, basecall_group, subgroup))
    except Exception as e:  # noqa: BLE001 -- the worker's blanket except (resquiggle.py:434)
        return ('err', str(e))


def write_job(args):
    """write_new_fast5_group(*args[:-1], event_data=args[-1]) -> None | error message"""
    from . import resquiggle
    try:
        resquiggle.write_new_fast5_group(*args[:-1], event_data=args[-1])
        return None
    except Exception as e:  # noqa: BLE001
        return str(e)


def write_packed_job(args):
    """write_packed_fast5_group(*args) -> None | error message"""
    from . import resquiggle
    try:
        resquiggle.write_packed_fast5_group(*args)
        return None
    except Exception as e:  # noqa: BLE001
        return str(e)


def start_jobs(fn, jobs):
    """submit fn over jobs and return at once: an iterable of the results, in order, that blocks as it is read"""
    p = pool()
    if not jobs:
        return []
    if p is None or any(str(j[0]).startswith('mem://') for j in jobs):
        return [fn(j) for j in jobs]
    return p.map(fn, jobs, chunksize=max(1, len(jobs) // (4 * io_procs()) or 1))


def map_jobs(fn, jobs):
    """results of fn over jobs, in order; in-memory FAST5 names (mem://) live in this process only"""
    p = pool()
    if p is None or any(str(j[0]).startswith('mem://') for j in jobs):
        return [fn(j) for j in jobs]
    return list(p.map(fn, jobs, chunksize=max(1, len(jobs) // (4 * io_procs()) or 1)))
