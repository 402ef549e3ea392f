#!/usr/bin/env python
"""bench.py -- throughput of the genome_resquiggle hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

A step is one pass of the whole a1..a7 path (indel scan, median/MAD order statistics,
re-segmentation, event table) over one batch of synthetic reads.  At N = 1 the workload is
BASELINE.json configs[1]: 100k synthetic ~10 kb reads resident in HBM.  For N > 1 every rank
holds its own 100k reads (reads are independent: no collective on the data path; weak scaling).

One JSON line on stdout (rank 0).  `value` is kernel-only samples/s with inputs resident in
HBM; `e2e` is the same metric through the C-ABI host-buffer call (nrq_resquiggle_batch_host)
with pinned host inputs/outputs and both copies inside the timed region; `roofline` is the
dominant kernel's algorithmic bytes / its CUDA-event time against the measured HBM peak;
`cpu_baseline` is the oracle port timed on this box's host cores on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from nanoraw_b200 import synth  # noqa: E402

METRIC = 'resquiggled raw samples/sec'
UNIT = 'samples/s'
STAGES = ('k_align_scan', 'k_norm_stats', 'k_resegment', 'k_event_table')


def log(*a):
    print(*a, file=sys.stderr, flush=True)


_JSON_OUT = [None]


def claim_stdout():
    """stdout carries the one JSON line and nothing else: keep a private handle on it and point fd 1 at stderr,
    so that whatever a library prints (NCCL's version banner, warnings of child processes) cannot get in"""
    if _JSON_OUT[0] is None:
        sys.stdout.flush()
        _JSON_OUT[0] = os.fdopen(os.dup(1), 'w')
        os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT[0] or sys.stdout
    out.write(json.dumps(line) + '\n')
    out.flush()


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def workload_config(name, seed_base):
    cfg = {'standard': synth.config_standard, 'short': synth.config_short,
           'long': synth.config_long, 'dense': synth.config_indel_dense}[name](seed_base)
    return cfg


# ----------------------------------------------------------------------------- CPU legs
def _oracle_one(read):
    from oracle import resquiggle_oracle as orc
    t0 = time.perf_counter()
    try:
        orc.resquiggle_read(read.raw, read.read_start_rel_to_raw, read.starts,
                            read.read_align.decode(), read.genome_align.decode(),
                            sort_kind='default', verbatim_clip_loop=True)
        ok = True
    except Exception:
        ok = False
    return read.n_samples if ok else 0, time.perf_counter() - t0


def time_oracle(reads, procs):
    """samples/s of the oracle port over `reads` with a process pool of `procs`."""
    import multiprocessing as mp
    if procs <= 1:
        t0 = time.perf_counter()
        res = [_oracle_one(r) for r in reads]
        dt = time.perf_counter() - t0
    else:
        with mp.get_context('fork').Pool(procs) as pool:
            pool.map(_oracle_one, reads[:procs])  # spin the workers up outside the timing
            t0 = time.perf_counter()
            res = pool.map(_oracle_one, reads, chunksize=1)
            dt = time.perf_counter() - t0
    samples = sum(s for s, _ in res)
    cpu_s = sum(t for _, t in res)
    return samples / dt, dt, samples, cpu_s


def reference_arm(args):
    """The reference's own CPU implementation of the path: the Python-3 oracle port (the Py2
    original cannot run in this image), all host cores, bounded sample per step."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = host_cores()
    cfg = workload_config(args.config, 2000)
    n = max(4, 4 * cores if args.ref_reads is None else args.ref_reads)
    genome = synth.make_genome(cfg.genome_len)
    levels = synth.kmer_levels(cfg)
    reads = [synth.make_read(genome, levels, cfg, i) for i in range(n)]
    for _ in range(args.warmup):
        time_oracle(reads[:max(cores, 1)], cores)
    tot_s, tot_t = 0, 0.0
    for _ in range(args.steps):
        _, dt, samples, _ = time_oracle(reads, cores)
        tot_s += samples
        tot_t += dt
    value = tot_s / tot_t
    sample = '%d synthetic %s reads per step (%d samples), oracle port with verbatim Python clip loop' % (
        n, args.config, tot_s // max(args.steps, 1))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * tot_t / max(args.steps, 1),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic', 'reads_per_s': n * args.steps / tot_t,
        'config': {'workload': workload_name(args), 'sample': sample},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)


def workload_name(args):
    return ('BASELINE.json configs[1]: %d synthetic ~%s reads per GPU preloaded in HBM '
            '(config "%s", %d unique reads x device-side +-2 DAC noise replicas)' % (
                args.reads, {'standard': '10 kb / ~89k-sample', 'short': '4.5 kb / ~40k-sample',
                             'long': '100 kb / ~0.9M-sample', 'dense': '10 kb indel-dense'}[args.config],
                args.config, min(args.unique, args.reads)))


# ----------------------------------------------------------------------------- clocks
class ClockSampler(object):
    FIELDS = ('timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
              'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
              'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []   # (wall-clock arrival time, csv line)
        self.window = [None, None]

    def mark_begin(self):
        self.window[0] = time.time()

    def mark_end(self):
        self.window[1] = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '--query-gpu=' + self.FIELDS, '--format=csv,noheader,nounits',
                 '-i', str(self.index), '-lms', '200'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        lo, hi = self.window
        inside = [ln for t, ln in self.lines if lo is not None and hi is not None and lo <= t <= hi + 0.25]
        scope = 'timed region'
        if len(inside) < 2:  # region shorter than the sampling period: use every sample under load
            inside = [ln for _, ln in self.lines]
            scope = 'warm-up + timed region'
        for ln in inside:
            parts = [p.strip() for p in ln.split(',')]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower() == 'active':
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None,
                'sm_max_mhz': float(max(smax)) if smax else None,
                'power_w_max': float(max(power)) if power else None,
                'samples': len(sm), 'scope': scope, 'reasons': sorted(reasons)}


# ----------------------------------------------------------------------------- GPU arm
def build_device_batch(engine, host_batch, reps, rank):
    """Tile the unique reads `reps` times on the device; replica k > 0 gets its own +-2 DAC
    noise so every read is distinct work (replica 0 stays exact for the oracle spot check)."""
    import torch
    from nanoraw_b200.batch import ReadBatch
    from nanoraw_b200.engine import DeviceBatch
    dev = engine.device
    U = host_batch.n_reads
    t = {}
    small = engine.upload(host_batch)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    for name in ('raw', 'starts', 'read_align', 'genome_align'):
        src = small.t[name]
        big = torch.empty(src.numel() * reps, dtype=src.dtype, device=dev)
        for k in range(reps):
            dst = big[k * src.numel():(k + 1) * src.numel()]
            dst.copy_(src)
            if name == 'raw' and k > 0:
                noise = torch.randint(-2, 3, (src.numel(),), device=dev, generator=gen, dtype=torch.int16)
                dst.add_(noise)
        t[name] = big
    for name, data in (('raw_off', 'raw'), ('starts_off', 'starts'), ('align_off', 'read_align')):
        off = small.t[name]
        n = small.t[data].numel()
        parts = [off[:-1] + k * n for k in range(reps)] + [torch.tensor([reps * n], dtype=off.dtype, device=dev)]
        t[name] = torch.cat(parts)
    t['read_start'] = small.t['read_start'].repeat(reps)
    total_align = host_batch.total_align * reps
    n_samples = torch.from_numpy(host_batch.n_samples).to(dev).repeat(reps)
    db = DeviceBatch.from_tensors(t, U * reps, total_align, dev, host=None)
    del small
    return db, n_samples


def spot_check(engine, host_batch, reads, out_dev, n_check=2):
    """replica 0 of the big batch is the unperturbed unique set: compare a couple of reads with
    the oracle (not timed; the real parity suite is tests/ -m gpu)."""
    from oracle import resquiggle_oracle as orc
    import torch
    bad = 0
    for r in range(min(n_check, len(reads))):
        read = reads[r]
        want = orc.resquiggle_read(read.raw, read.read_start_rel_to_raw, read.starts,
                                   read.read_align.decode(), read.genome_align.decode())
        o = int(host_batch.align_off[r])
        nb = int(out_dev.n_bases[r].item())
        segs = out_dev.new_segs[o + r:o + r + nb + 1].cpu().numpy()
        mean = out_dev.norm_mean[o:o + nb].cpu().numpy()
        ref_mean = want['events']['norm_mean']
        if not (np.array_equal(segs, want['new_segs']) and
                np.all(np.abs(mean - ref_mean) <= 1e-9 * np.maximum(1.0, np.abs(ref_mean)))):
            bad += 1
    return bad


def gpu_arm(args):
    import torch
    import torch.distributed as dist
    from nanoraw_b200 import _lib, distributed
    from nanoraw_b200.batch import ReadBatch
    from nanoraw_b200.engine import ResquiggleEngine, ResquiggleParams, DeviceResults

    # stdout carries the one JSON line and nothing else: NCCL's own messages (its version banner under
    # NCCL_DEBUG=VERSION/WARN/INFO) go to stderr
    os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
    rank, world, local = distributed.init('nccl')  # one process per GPU; no collective on the data path
    torch.cuda.set_device(local)
    engine = ResquiggleEngine()
    lib = engine.lib
    dev = engine.device
    # the device entry point cannot see read lengths: say that the long-read workload holds reads to stream in parts
    params = ResquiggleParams(debug_flags=args.debug_flags, long_read_parts=16 if args.config == 'long' else 0)
    pc = params.to_c()

    # ---- synthetic reads: `unique` per rank on the host, tiled to `reads` on the device
    unique = min(args.unique, args.reads)
    reps = max(1, args.reads // unique)
    n_reads = unique * reps
    cfg = workload_config(args.config, 2000 + rank * unique)
    t0 = time.perf_counter()
    reads = synth.make_reads_parallel(unique, cfg)
    host_batch = ReadBatch.from_reads(reads)
    log('[rank %d] generated %d unique reads in %.1fs (%d samples)' % (
        rank, unique, time.perf_counter() - t0, int(host_batch.n_samples.sum())))
    db, n_samples = build_device_batch(engine, host_batch, reps, rank)
    out = DeviceResults(db.n_reads, db.total_align, dev)
    ws = engine.workspace(db.n_reads, db.total_align, pc)
    torch.cuda.synchronize()
    log('[rank %d] device batch: %d reads, inputs %.2f GB, outputs %.2f GB, workspace %.2f GB' % (
        rank, db.n_reads, db.device_bytes() / 1e9, out.device_bytes() / 1e9, ws.numel() / 1e9))

    def barrier():
        distributed.barrier()
        torch.cuda.synchronize()

    # ---- warm-up, then the timed kernel-only region
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        engine.run_device(db, pc, out=out, workspace=ws)
    torch.cuda.synchronize()
    bad = spot_check(engine, host_batch, reads, out) if rank == 0 else 0
    barrier()
    sampler.mark_begin()
    launches0 = lib.nrq_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_ms5 = np.zeros(5)
    ev0.record()
    for _ in range(args.steps):
        stage_ms5 += np.array(engine.run_device_profiled(db, pc, out, ws))
    ev1.record()
    barrier()
    sampler.mark_end()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = lib.nrq_launch_count() - launches0
    clocks = sampler.stop()
    stage_ms5 /= max(args.steps, 1)
    # the re-segmentation stage is two kernels: speculative first searches (k_speculate), then the walk (k_resegment)
    stage_ms = np.array([stage_ms5[0], stage_ms5[1], stage_ms5[2] + stage_ms5[3], stage_ms5[4]])

    ok = (out.status[:db.n_reads] == 0)
    samples_ok = int((n_samples * ok).sum().item())
    reads_ok = int(ok.sum().item())
    # accounting totals for the roofline (per launch = per step)
    tot = {
        'S': int(n_samples.sum().item()),
        'Br1': int(db.t['starts'].numel()),
        'A': int(db.total_align),
        'Bg': int(out.n_bases[:db.n_reads].sum().item()),
        'I': int(out.n_indels[:db.n_reads].sum().item()),
        'S_reg': int(out.region_samples[:db.n_reads].sum().item()),
        'R': db.n_reads,
    }

    # ---- e2e: host buffers through the C-ABI host entry point, copies inside the timed region
    # one call carries about 2 GB of input (the unique reads repeated), the batch size a host with pinned
    # staging would use; the call is synchronous, so pipeline fill and drain are inside every step
    in_bytes = sum(getattr(host_batch, f).nbytes for f in ReadBatch.FIELDS)
    e2e_reps = max(1, min(args.e2e_reps, int(2.2e9 // max(in_bytes, 1))))
    e2e_batch = host_batch if e2e_reps == 1 else ReadBatch.from_reads(reads * e2e_reps)
    e2e_samples, e2e_ms, h2d, d2h = e2e_leg(args, lib, _lib, e2e_batch, pc, local, barrier)

    # ---- reduce over ranks: max time, summed work
    (elapsed_ms, e2e_ms), (samples_all, reads_all, e2e_samples_all, launches_all) = distributed.combine(
        [elapsed_ms, e2e_ms], [samples_ok, reads_ok, e2e_samples, launches], device=dev)

    if rank == 0:
        ms_per_step = elapsed_ms / max(args.steps, 1)
        value = samples_all / (ms_per_step / 1e3)
        peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
        if os.path.exists(peaks_path):
            peak = float(json.load(open(peaks_path))['hbm_gbs'])
            peak_src = 'MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)'
        else:
            peak, peak_src = 6650.0, 'fallback (B200_PROFILING.md)'
        stage_bytes = algorithmic_bytes(tot)
        stages = {}
        for name, ms in zip(STAGES, stage_ms):
            gbs = stage_bytes[name] / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
            stages[name] = {'ms': float(ms), 'algorithmic_bytes': stage_bytes[name],
                            'gbs': gbs, 'frac': gbs / peak}
        stages['k_resegment']['parts_ms'] = {'k_speculate': float(stage_ms5[2]), 'k_resegment': float(stage_ms5[3])}
        dom = max(stages, key=lambda k: stages[k]['ms'])
        path_gbs = stage_bytes['path'] / (float(stage_ms.sum()) * 1e-3) / 1e9
        traffic, traffic_src = measured_traffic(args, dom)
        roofline = {'bound': 'hbm', 'kernel': dom, 'achieved': stages[dom]['gbs'], 'peak': peak,
                    'unit': 'GB/s', 'frac': stages[dom]['frac'], 'traffic': traffic,
                    'traffic_source': traffic_src,
                    'peak_source': peak_src, 'share_of_step': stages[dom]['ms'] / float(stage_ms.sum()),
                    'stages': stages,
                    'path': {'algorithmic_bytes': stage_bytes['path'], 'achieved': path_gbs,
                             'frac': path_gbs / peak},
                    'issue': issue_figures(args, stages, clocks)}
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'reads_per_s': reads_all / (ms_per_step / 1e3),
            'config': {'workload': workload_name(args), 'reads_per_gpu': db.n_reads,
                       'samples_per_gpu': tot['S'], 'l2': 'inputs (%.1f GB per GPU) larger than L2' % (
                           db.device_bytes() / 1e9),
                       'e2e': 'nrq_resquiggle_batch_host_submit / _wait, %d reads per step, pinned host buffers, step i + 1 '
                              'submitted while step i comes back (two result buffer sets); FAST5 I/O is the '
                              'separate e2e_fast5 leg' % e2e_batch.n_reads,
                       'failed_reads': int(db.n_reads * world - reads_all),
                       'oracle_spot_check_mismatches': bad},
            'e2e': {'value': e2e_samples_all / (e2e_ms / 1e3) if e2e_ms > 0 else None, 'unit': UNIT,
                    'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h},
            'gpu_launches': int(launches_all),
            'clocks': clocks,
            'roofline': roofline,
        }
        if world == 1 and not args.no_cpu_baseline:
            line['cpu_baseline'] = cpu_baseline_leg(args, reads)
            if args.divergence_reads > 0:
                line['tie_order_divergence'] = divergence_leg(args, engine, reads)
    # ---- FAST5 in, FAST5 out (BASELINE.json configs[2] at this N): every rank on its own files and I/O processes
    if args.fast5_files > 0:
        if world > 1:
            os.environ['NRQ_IO_PROCS'] = str(max(2, host_cores() // world))
            os.environ.setdefault('NRQ_F5_THREADS', str(max(2, host_cores() // world)))
            os.environ.setdefault('NRQ_PACK_THREADS', str(max(2, min(16, host_cores() // world))))
            # the host cores are shared by all ranks: fewer files per rank keep the leg's duration about the same
            from nanoraw_b200 import resquiggle as _rsq
            args.fast5_files = max(2 * _rsq.DEVICE_BATCH_READS, args.fast5_files // world)
        f5 = fast5_leg(args, reads, barrier)
        (dt,), agg = distributed.combine([f5['seconds']], [f5['ok_samples'], f5['files'] - f5['failed_reads'],
                                                           f5['failed_reads']], device=dev)
        if rank == 0:
            f5.update({'value': agg[0] / dt, 'reads_per_s': agg[1] / dt,
                       'files': args.fast5_files * world, 'failed_reads': int(agg[2]), 'seconds': dt,
                       'n_gpus': world, 'host_cores': host_cores()})
            del f5['ok_samples']
            line['e2e_fast5'] = f5
            if world == 1 and not args.no_cpu_baseline:
                line['cpu_baseline_fast5'] = cpu_fast5_leg(args, reads)
    if rank == 0:
        emit(line)
    if world > 1:
        distributed.barrier()
        dist.destroy_process_group()


TRAFFIC_FILE = 'r3z_traffic_100k.json'
STAGE_KERNELS = {'k_align_scan': ('k_align_count', 'k_scan_counts', 'k_align_fill'), 'k_norm_stats': ('k_norm_stats',),
                 'k_resegment': ('k_speculate', 'k_resegment'), 'k_event_table': ('k_event_table',)}


def ncu_capture(args):
    """per-kernel figures of the committed `ncu --set full` capture of this same workload (profiles/), or None
    when the workload differs"""
    path = os.path.join(ROOT, 'profiles', TRAFFIC_FILE)
    if not (os.path.exists(path) and args.reads == 100_000 and args.config == 'standard'):
        return None
    return json.load(open(path))['kernels']


def measured_traffic(args, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`"""
    data = ncu_capture(args)
    if data is None:
        return None, None
    total = sum(data[n]['dram_read_bytes'] + data[n]['dram_write_bytes'] for n in STAGE_KERNELS[kernel] if n in data)
    return int(total), 'profiles/%s (ncu --set full, same workload)' % TRAFFIC_FILE


def issue_figures(args, stages, clocks):
    """The stages are bound by instruction issue rather than by DRAM: warp instructions per launch (ncu capture of
    the same workload) over the stage's CUDA-event time, against 4 schedulers x 1 instruction / clock per SM."""
    data = ncu_capture(args)
    if data is None or not clocks.get('sm_mhz'):
        return None
    import torch
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    peak = sms * 4 * clocks['sm_mhz'] * 1e6
    out = {'peak_warp_inst_per_s': peak, 'sms': sms, 'sm_mhz': clocks['sm_mhz'], 'stages': {}}
    for name, st in stages.items():
        inst = sum(data[n].get('warp_instructions', 0) for n in STAGE_KERNELS[name] if n in data)
        if inst and st['ms'] > 0:
            rate = inst / (st['ms'] * 1e-3)
            out['stages'][name] = {'warp_instructions': inst, 'achieved': rate, 'frac': rate / peak}
    return out


def algorithmic_bytes(t):
    """Compulsory bytes per launch (every input read once, every output written once); DESIGN.md
    section 'Roofline accounting' states the per-unit figures."""
    return {
        # alignment rows read twice is NOT compulsory: count once; indel int4 + bases written
        'k_align_scan': 2 * t['A'] + 16 * t['I'] + t['Bg'] + 12 * t['R'],
        'k_norm_stats': 2 * t['S'] + 32 * t['R'],
        'k_resegment': 2 * t['S_reg'] + 4 * t['Br1'] + 16 * t['I'] + 4 * (t['Bg'] + t['R']) + 32 * t['R'],
        'k_event_table': 2 * t['S'] + 4 * (t['Bg'] + t['R']) + 24 * t['Bg'] + 32 * t['R'],
        # 2S + 4(Br+1) + alignment bytes + Bg + 25 Bg + 32 per read
        'path': 2 * t['S'] + 4 * t['Br1'] + 2 * t['A'] + t['Bg'] + 25 * t['Bg'] + 32 * t['R'],
    }


def e2e_leg(args, lib, _lib, host_batch, pc, device_index, barrier):
    import ctypes as C
    import torch
    from nanoraw_b200.batch import ReadBatch
    R, A = host_batch.n_reads, host_batch.total_align
    keep = []

    def pinned(arr):
        t = torch.from_numpy(np.ascontiguousarray(arr)).pin_memory()
        keep.append(t)
        return t

    hb = _lib.NrqBatch()
    hb.n_reads = R
    h2d = 0
    for f in ReadBatch.FIELDS:
        t = pinned(getattr(host_batch, f))
        h2d += t.numel() * t.element_size()
        setattr(hb, f, t.data_ptr())
    # what resquiggle_read hands to the FAST5 writer: scale values + the event table (+ status).  new_segs and the
    # `start` column are the running sum of `length` (nrqio_pack_corrected rebuilds start on the host), n_indels and
    # region_samples are accounting aids: none of them is copied back
    outs = {'status': (R, np.int32), 'n_bases': (R, np.int32),
            'scale_values': (4 * R, np.float64),
            'norm_mean': (A, np.float64), 'norm_stdev': (A, np.float64),
            'ev_length': (A, np.uint32), 'ev_base': (A, np.uint8)}
    # two sets of host result buffers: step i + 1 is submitted (nrq_resquiggle_batch_host_submit) while step i is
    # still coming back, as a worker that keeps reads coming does, so the copy / kernel pipeline never drains
    # between steps; every step's inputs go host -> device and its results device -> host inside the timed region
    hrs, host_outs = [], []
    d2h = 0
    for _ in range(2):
        hr = _lib.NrqResults()
        host_out = {}
        for k, (n, dt) in outs.items():
            t = pinned(np.zeros(n, dtype=dt).view(np.int32 if dt == np.uint32 else dt))
            host_out[k] = t
            setattr(hr, k, t.data_ptr())
        hrs.append(hr)
        host_outs.append(host_out)
    d2h = sum(t.numel() * t.element_size() for t in host_outs[0].values())
    h = lib.nrq_create(device_index)
    if not h:
        raise RuntimeError('nrq_create failed: ' + lib.nrq_last_error().decode())

    def run_steps(n_steps, what):
        prev = None
        for i in range(n_steps):
            ticket = lib.nrq_resquiggle_batch_host_submit(h, C.byref(hb), C.byref(pc), C.byref(hrs[i % 2]))
            if ticket < 0:
                _lib.check(int(ticket), what)
            if prev is not None:
                _lib.check(lib.nrq_resquiggle_batch_host_wait(h, prev), what)  # step i - 1 is in host memory
            prev = ticket
        if prev is not None:
            _lib.check(lib.nrq_resquiggle_batch_host_wait(h, prev), what)

    run_steps(max(args.warmup, 2), 'e2e warm-up')
    barrier()
    t0 = time.perf_counter()
    run_steps(args.e2e_steps, 'e2e step')
    e2e_ms = (time.perf_counter() - t0) * 1e3 / max(args.e2e_steps, 1)
    lib.nrq_destroy(h)
    status = host_outs[(args.e2e_steps - 1) % 2]['status'].numpy()
    samples = int(host_batch.n_samples[status == 0].sum())
    return samples, e2e_ms, h2d, d2h


def _libio_threads():
    from nanoraw_b200 import _libio
    return _libio.f5_threads()


def host_work_per_read(jobs, pend, reads_per_s):
    """What bounds the FAST5 leg: the host work per read of each stage of the worker, measured by running a sample
    of the leg's own files through the same pipeline on ONE thread (NRQ_F5_THREADS = NRQ_PACK_THREADS = 1), and
    the reads/s the host cores of this rank could do at that cost -- the leg's equivalent of a roofline."""
    from nanoraw_b200 import fast5_jobs
    from nanoraw_b200 import resquiggle as rsq
    fast5_jobs.map_jobs(synth.drop_corrected_job, jobs)          # back to the state prep_fast5 leaves
    saved = {k: os.environ.get(k) for k in ('NRQ_F5_THREADS', 'NRQ_PACK_THREADS')}
    os.environ['NRQ_F5_THREADS'] = os.environ['NRQ_PACK_THREADS'] = '1'
    try:
        pipe = rsq.BatchPipeline('median', 5, None, None, 'Basecall_1D_000', 'RawGenomeCorrected_000', True, None)
        failed = pipe.feed(pend) + pipe.drain()
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    n = max(len(pend) - len(failed), 1)
    ms = {k: round(1e3 * v / n, 4) for k, v in pipe.stage_seconds.items()}
    cores = max(1, host_cores() // max(1, int(os.environ.get('WORLD_SIZE', '1'))))
    cpu_ms = sum(v for k, v in ms.items() if k != 'device')
    bound = 1e3 * cores / cpu_ms if cpu_ms > 0 else None
    return {'ms_per_read_one_thread': ms, 'sample_reads': n, 'cores_of_this_rank': cores,
            'bound_reads_per_s': bound, 'frac_of_bound': (reads_per_s / bound) if bound else None,
            'what': 'open = file read + HDF5 walk, fill = CSR arrays of the batch (Python), inflate = raw signal into '
                    'pinned staging, device = nrq_resquiggle_batch_host, pack = four gzip chunks per read, append = '
                    'RawGenomeCorrected group written into the file; bound = cores / (sum without device)'}


def fast5_leg(args, reads, barrier=None):
    """The re-squiggle worker's whole body (resquiggle.py:403-440 -> 318-401) on FAST5 files on disk: open each
    file and read its raw signal, one device batch through the C ABI, RawGenomeCorrected written back into each
    file.  File access is libnrqio's FAST5 functions (include/nrq_io.h) on NRQ_F5_THREADS host threads.  Alignments are
    the synthetic ones (the mapper stays external), handed over as the queue items the alignment workers make."""
    import shutil
    import tempfile
    from nanoraw_b200 import fast5_io, fast5_jobs
    from nanoraw_b200 import resquiggle as rsq
    from nanoraw_b200.frontend import genomeLoc, readInfo
    n = args.fast5_files
    tmp = tempfile.mkdtemp(prefix='nrq_fast5_')
    try:
        jobs = [(os.path.join(tmp, 'read_%05d.fast5' % i), reads[i % len(reads)]) for i in range(n)]
        t0 = time.perf_counter()
        sizes = fast5_jobs.map_jobs(synth.write_basecalled_fast5_job, jobs)
        setup_s = time.perf_counter() - t0

        def pending_reads():
            out = []
            for fn, read in jobs:
                info = readInfo(os.path.basename(fn), 'BaseCalled_template', 0, 0, 0, 0, 0, 0)
                loc = genomeLoc(int(read.genome_start), read.strand, read.chrom)
                out.append(rsq._PendingRead(fn, ((read.read_align, read.genome_align), loc, read.starts,
                                                 int(read.read_start_rel_to_raw), info)))
            return out

        def one_pass():
            pend = pending_reads()
            pipe = rsq.BatchPipeline('median', 5, None, None, 'Basecall_1D_000', 'RawGenomeCorrected_000', True, None)
            failed = []
            for lo in range(0, len(pend), rsq.DEVICE_BATCH_READS):  # as resquiggle_worker feeds its batches
                failed += pipe.feed(pend[lo:lo + rsq.DEVICE_BATCH_READS])
            return failed + pipe.drain()

        one_pass()                                   # warm-up: pool start-up, pinned staging, kernels
        fast5_jobs.map_jobs(synth.drop_corrected_job, jobs)      # back to the state prep_fast5 leaves (not timed)
        if barrier is not None:
            barrier()
        t0 = time.perf_counter()
        failed = one_pass()
        dt = time.perf_counter() - t0
        samples_of = {fn: read.n_samples for fn, read in jobs}
        ok_samples = sum(samples_of.values()) - sum(samples_of[pr.fast5_fn] for _, pr in failed)
        host = host_work_per_read(jobs[:min(n, 128)], pending_reads()[:min(n, 128)], (n - len(failed)) / dt)
        # spot check of what landed on disk
        f = fast5_io.open_fast5(jobs[0][0], 'r')
        ev = f['/Analyses/RawGenomeCorrected_000/BaseCalled_template/Events'][()]
        f.close()
        return {'value': ok_samples / dt, 'ok_samples': ok_samples, 'unit': UNIT,
                'reads_per_s': (n - len(failed)) / dt, 'files': n, 'batch_reads': rsq.DEVICE_BATCH_READS,
                'failed_reads': len(failed), 'seconds': dt, 'io_procs': fast5_jobs.io_procs(),
                'bytes_per_file_in': int(np.mean(sizes)), 'events_rows_file0': int(len(ev)),
                'setup_seconds': setup_s,
                'f5_threads': _libio_threads(), 'host_work': host,
                'what': 'resquiggle.BatchPipeline (the re-squiggle worker body): FAST5 open + raw signal inflated '
                        'into pinned staging (libnrqio threads), device batch, datasets packed into gzip chunks and '
                        'the RawGenomeCorrected group appended to the same file (libnrqio threads; hdf5_min.py '
                        'for files outside its HDF5 slice); mapper and basecall event parsing not included'}
    finally:
        fast5_jobs.shutdown()
        shutil.rmtree(tmp, ignore_errors=True)


def _cpu_fast5_one(job):
    """the reference's per-read path with its file work (resquiggle.py:318-401 + :55-137) on one core: FAST5 open
    and raw signal read, the oracle port, RawGenomeCorrected written with zlib level 4 as h5py does"""
    from oracle import resquiggle_oracle as orc
    from nanoraw_b200 import resquiggle as rsq
    from nanoraw_b200 import nanoraw_helper as nh
    from nanoraw_b200.frontend import genomeLoc, readInfo
    fn, read = job
    t0 = time.perf_counter()
    try:
        _, raw, _, _ = rsq._read_raw_inputs(fn, 'median', 'Basecall_1D_000', 'BaseCalled_template')
        out = orc.resquiggle_read(raw, read.read_start_rel_to_raw, read.starts, read.read_align.decode(),
                                  read.genome_align.decode(), sort_kind='default', verbatim_clip_loop=True)
        ev = np.empty(len(out['events']['start']), dtype=np.dtype(rsq.EVENT_DTYPE))
        for col in ('norm_mean', 'norm_stdev', 'start', 'length', 'base'):
            ev[col] = out['events'][col]
        rsq.write_new_fast5_group(
            fn, genomeLoc(int(read.genome_start), read.strand, read.chrom),
            readInfo(os.path.basename(fn), 'BaseCalled_template', 0, 0, 0, 0, 0, 0),
            int(read.read_start_rel_to_raw), out['new_segs'], ev['base'].tobytes().decode(),
            (read.read_align, read.genome_align), read.starts, None, nh.scaleValues(*out['scale_values']),
            'RawGenomeCorrected_000', 'BaseCalled_template', 'median', 5, True, event_data=ev)
        ok = True
    except Exception:
        ok = False
    return read.n_samples if ok else 0, time.perf_counter() - t0


def cpu_fast5_leg(args, reads):
    """CPU figure that includes the file work, next to e2e_fast5: a bounded sample of the same files"""
    import multiprocessing as mp
    import shutil
    import tempfile
    cores = host_cores()
    n = min(len(reads), max(8, 8 * cores if args.cpu_reads is None else args.cpu_reads))
    tmp = tempfile.mkdtemp(prefix='nrq_cpu_fast5_')
    try:
        jobs = [(os.path.join(tmp, 'read_%05d.fast5' % i), reads[i]) for i in range(n)]
        with mp.get_context('fork').Pool(cores) as pool:
            pool.map(synth.write_basecalled_fast5_job, jobs)
            t0 = time.perf_counter()
            res = pool.map(_cpu_fast5_one, jobs, chunksize=1)
            dt = time.perf_counter() - t0
        samples = sum(s for s, _ in res)
        cpu_s = sum(t for _, t in res)
        return {'value': samples / dt, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                'sample': '%d FAST5 files of the same workload (%d samples), %.1f s wall, %.1f CPU-s: file read + '
                          'oracle port + RawGenomeCorrected write (zlib level 4, as h5py)' % (n, samples, dt, cpu_s),
                'per_core': samples / cpu_s if cpu_s > 0 else None}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def _divergence_one(args):
    """per read: boundaries of the deterministic order (what the GPU must equal) and of numpy's default argsort"""
    from oracle import resquiggle_oracle as orc
    read, gpu_segs = args
    try:
        d = orc.resquiggle_read(read.raw, read.read_start_rel_to_raw, read.starts, read.read_align.decode(),
                                read.genome_align.decode(), sort_kind='default')['new_segs']
    except Exception:
        return 0, 0, 0
    if len(d) != len(gpu_segs):
        return len(d), len(d), 1
    diff = int((np.asarray(d) != np.asarray(gpu_segs)).sum())
    return len(d), diff, int(diff > 0)


def divergence_leg(args, engine, reads):
    """BASELINE.json: 'any float64 near-tie divergence logged and bounded'.  The reference visits running
    differences in numpy's default (unstable) argsort order (resquiggle.py:246); the device path is held to the
    deterministic order.  Fraction of boundaries of the GPU output that differ from the reference order on THIS
    host's numpy, over the first --divergence-reads reads of the workload."""
    import multiprocessing as mp
    from nanoraw_b200.batch import ReadBatch
    from nanoraw_b200.engine import ResquiggleParams
    n = min(len(reads), args.divergence_reads)
    sub = reads[:n]
    res = engine.run(ReadBatch.from_reads(sub), ResquiggleParams())
    jobs = [(rd, res.new_segs_of(r)) for r, rd in enumerate(sub) if res.ok(r)]
    with mp.get_context('fork').Pool(host_cores()) as pool:
        out = pool.map(_divergence_one, jobs, chunksize=4)
    total = sum(a for a, _, _ in out)
    diff = sum(b for _, b, _ in out)
    return {'reads': len(jobs), 'boundaries': total, 'differing_boundaries': diff,
            'fraction': diff / total if total else None, 'reads_affected': sum(c for _, _, c in out),
            'against': "oracle port with sort_kind='default' (numpy %s argsort on this host); the GPU output "
                       "equals sort_kind='stable' bit for bit (tests/)" % np.__version__}


def cpu_baseline_leg(args, reads):
    cores = host_cores()
    n = min(len(reads), max(8, 8 * cores if args.cpu_reads is None else args.cpu_reads))
    value, dt, samples, cpu_s = time_oracle(reads[:n], cores)
    return {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': '%d reads of the same workload (%d samples), %.1f s wall, %.1f CPU-s; oracle port '
                      'with verbatim Python clip loop' % (n, samples, dt, cpu_s),
            'per_core': samples / cpu_s if cpu_s > 0 else None}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=('ours', 'reference'))
    ap.add_argument('--reads', type=int, default=100_000, help='reads per GPU')
    ap.add_argument('--unique', type=int, default=4000, help='distinct synthetic reads generated on the host')
    ap.add_argument('--config', default='standard', choices=('standard', 'short', 'long', 'dense'))
    ap.add_argument('--e2e-steps', type=int, default=None, help='steps of the e2e leg (default: --steps)')
    ap.add_argument('--e2e-reps', type=int, default=4, help='unique reads repeated this often in one e2e call')
    ap.add_argument('--cpu-reads', type=int, default=None)
    ap.add_argument('--ref-reads', type=int, default=None)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--debug-flags', type=int, default=0, help='nrq_params.debug_flags (4 = no speculative search)')
    ap.add_argument('--fast5-files', type=int, default=8192,
                    help='FAST5 files on disk per GPU for the e2e_fast5 leg (0 = skip)')
    ap.add_argument('--divergence-reads', type=int, default=1024,
                    help='reads compared with the reference\'s default argsort order (N = 1; 0 = skip)')
    args = ap.parse_args()
    if args.e2e_steps is None:
        args.e2e_steps = max(args.steps, 1)
    claim_stdout()
    if args.impl == 'reference':
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == '__main__':
    main()
