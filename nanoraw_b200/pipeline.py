"""Orchestration of `nanoraw genome_resquiggle` (reference: resquiggle.py:1037-1203): file
discovery, process / thread arithmetic, the three queues, worker start-up and the failed-read
summary.  Alignment workers are CPU processes driving the external mapper; re-squiggle workers
are GPU workers, one process per worker, assigned round-robin to the visible devices.  Reads
reach a GPU through the shared queue, so they shard across devices with no collective."""
import multiprocessing as mp
import os
import queue
import sys
from collections import defaultdict
from glob import glob
from time import sleep

from . import messages as msg
from . import nanoraw_helper as nh
from . import option_parsers
from .frontend import alignment_worker

ALIGN_BATCH_MULTIPLIER = 5  # parsed reads allowed to wait for the GPU, in alignment batches (:27)


def _gpu_worker(device_index, worker_args, n_workers=1):
    import torch
    from . import fast5_jobs, resquiggle
    resquiggle.VERBOSE = worker_args[-1]
    if torch.cuda.is_available():
        torch.cuda.set_device(device_index % torch.cuda.device_count())
    # the host cores are shared between the GPU workers' FAST5 I/O pools
    cores = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    os.environ.setdefault('NRQ_IO_PROCS', str(max(1, min(32, cores // max(n_workers, 1)))))
    try:
        resquiggle.resquiggle_worker(*worker_args[:-1])
    finally:
        fast5_jobs.shutdown()


def _drain(q, into):
    try:
        while True:
            err, name = q.get(block=False)
            into[err].append(name)
    except queue.Empty:
        pass


def resquiggle_all_reads(
        fast5_fns, genome_fn, mapper_exe, mapper_type,
        basecall_group, basecall_subgroups, corrected_group, norm_type,
        outlier_thresh, timeout, num_cpts_limit, overwrite,
        align_batch_size, num_align_ps, align_threads_per_proc,
        num_resquiggle_ps, compute_sd, pore_model):
    """resquiggle.py:1037-1114 -> {error string: [read ids]}"""
    from . import resquiggle
    ctx = mp.get_context('spawn')  # CUDA contexts do not survive fork
    manager = ctx.Manager()
    fast5_q = manager.Queue()
    basecalls_q = manager.Queue(align_batch_size * ALIGN_BATCH_MULTIPLIER)
    failed_reads_q = manager.Queue()
    fast5_fns = list(fast5_fns)
    for lo in range(0, len(fast5_fns), align_batch_size):
        fast5_q.put(fast5_fns[lo:lo + align_batch_size])

    aligners = [ctx.Process(target=alignment_worker, args=(
        fast5_q, basecalls_q, failed_reads_q, genome_fn, mapper_exe, mapper_type,
        basecall_group, basecall_subgroups, corrected_group, overwrite, align_threads_per_proc))
        for _ in range(num_align_ps)]
    worker_args = (basecalls_q, failed_reads_q, basecall_group, corrected_group, norm_type,
                   outlier_thresh, timeout, num_cpts_limit, compute_sd, pore_model,
                   resquiggle.VERBOSE)
    squigglers = [ctx.Process(target=_gpu_worker, args=(i, worker_args, num_resquiggle_ps))
                  for i in range(num_resquiggle_ps)]
    for p in aligners + squigglers:
        p.start()

    if resquiggle.VERBOSE:
        sys.stderr.write(
            'Correcting %d files with %d subgroup(s)/read(s) each (Will print a dot for each '
            '100 files completed).\n' % (len(fast5_fns), len(basecall_subgroups)))
    failed_reads = defaultdict(list)
    while any(p.is_alive() for p in aligners):
        _drain(failed_reads_q, failed_reads)
        sleep(0.2)
    for _ in squigglers:  # one end-of-input marker per re-squiggle worker (:1095-1096)
        basecalls_q.put((None, None))
    while any(p.is_alive() for p in squigglers):
        _drain(failed_reads_q, failed_reads)
        sleep(0.2)
    _drain(failed_reads_q, failed_reads)
    if resquiggle.VERBOSE:
        sys.stderr.write('\n')
    return dict(failed_reads)


def _die(text):
    sys.stderr.write('%s\n%s\n%s\n' % (msg.BANNER, text, msg.BANNER))
    sys.exit()


def list_fast5_files(args):
    """resquiggle.py:1134-1143"""
    if args.fast5_pattern or args.recursive:
        pattern = '*/*.fast5' if args.recursive else args.fast5_pattern
        return glob(os.path.join(args.fast5_basedir, pattern))
    return [path for path in (os.path.join(args.fast5_basedir, fn)
                              for fn in os.listdir(args.fast5_basedir)) if os.path.isfile(path)]


def resolve_processes(args):
    """resquiggle.py:1162-1170 -> (align threads per process, re-squiggle processes)"""
    num_proc = max(2, args.processes)
    threads = args.align_threads_per_process
    if threads is None:
        threads = max(int(num_proc / (2 * args.align_processes)), 1)
    squigglers = args.resquiggle_processes
    if squigglers is None:
        squigglers = int(num_proc / 2)
    return threads, squigglers


def format_failed_reads(failed_reads):
    """the stderr summary and the --failed-reads-filename body (resquiggle.py:1189-1196)"""
    summary = 'Failed reads summary:\n' + '\n'.join(
        '\t%s :\t%d' % (err, len(names)) for err, names in failed_reads.items()) + '\n'
    table = '\n'.join(err + '\t' + ','.join(names) for err, names in failed_reads.items()) + '\n'
    return summary, table


def resquiggle_main(args):
    """resquiggle.py:1116-1198"""
    from . import resquiggle
    resquiggle.VERBOSE = not args.quiet
    if args.graphmap_executable is not None:
        mapper_exe, mapper_type = args.graphmap_executable, 'graphmap'
    elif args.bwa_mem_executable is not None:
        mapper_exe, mapper_type = args.bwa_mem_executable, 'bwa_mem'
    else:
        _die(msg.NEED_MAPPER)

    if resquiggle.VERBOSE:
        sys.stderr.write('Getting file list.\n')
    try:
        files = list_fast5_files(args)
    except OSError:
        _die(msg.DIR_INACCESSIBLE)
    if len(files) < 1:
        _die(msg.NO_FILES)

    # a non-positive threshold switches outlier clipping off (:1159-1160)
    outlier_thresh = args.outlier_threshold if args.outlier_threshold > 0 else None
    align_threads, num_resquiggle_ps = resolve_processes(args)
    pore_model = None
    if args.normalization_type == 'pA':
        pore_model = nh.parse_pore_model(args.pore_model_filename)

    failed_reads = resquiggle_all_reads(
        files, args.genome_fasta, mapper_exe, mapper_type, args.basecall_group,
        args.basecall_subgroups, args.corrected_group, args.normalization_type, outlier_thresh,
        args.timeout, args.cpts_limit, args.overwrite, args.alignment_batch_size,
        args.align_processes, align_threads, num_resquiggle_ps, not args.skip_event_stdev,
        pore_model)
    summary, table = format_failed_reads(failed_reads)
    sys.stderr.write(summary)
    if args.failed_reads_filename is not None:
        with open(args.failed_reads_filename, 'w') as fp:
            fp.write(table)


def args_and_main():
    resquiggle_main(option_parsers.get_resquiggle_parser().parse_args())
