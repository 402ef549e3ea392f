"""world_size-2 run of the multi-GPU host logic on CPU (gloo): sharding by samples, the counter
reduction bench.py uses, and the status gather.  No CUDA involved."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as tmp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    from nanoraw_b200 import distributed, synth
    from nanoraw_b200.batch import ReadBatch
    r, w, _ = distributed.init('gloo')
    assert (r, w) == (rank, world)
    cfg = synth.SynthConfig(span_mean=120, span_sd=40, seed_base=980)
    reads = synth.make_reads(23, cfg, genome=synth.make_genome(20_000, seed=4))
    whole = ReadBatch.from_reads(reads)
    mine = distributed.shard_batch(whole, rank, world)
    # every rank "processes" its shard: here just counts samples; read r fails on odd sample counts
    status = (mine.n_samples % 2).astype(np.int32)
    elapsed_ms = 10.0 * (rank + 1)
    (t_max,), (samples, n_reads) = distributed.combine(
        [elapsed_ms], [int(mine.n_samples.sum()), mine.n_reads])
    all_status = distributed.gather_statuses(status)
    distributed.barrier()
    if rank == 0:
        np.save(os.path.join(out_dir, 'result.npy'),
                np.array([t_max, samples, n_reads, int(whole.n_samples.sum()), whole.n_reads]))
        np.save(os.path.join(out_dir, 'status.npy'), all_status)
        np.save(os.path.join(out_dir, 'expect.npy'), (whole.n_samples % 2).astype(np.int32))
    import torch.distributed as dist
    dist.destroy_process_group()


def test_two_rank_sharding_and_reduction(tmp_path):
    port = _free_port()
    tmp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    t_max, samples, n_reads, want_samples, want_reads = np.load(tmp_path / 'result.npy')
    assert t_max == 20.0                      # max over ranks, not the sum or rank 0's
    assert samples == want_samples and n_reads == want_reads
    assert np.array_equal(np.load(tmp_path / 'status.npy'), np.load(tmp_path / 'expect.npy'))
