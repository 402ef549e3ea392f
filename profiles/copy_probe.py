#!/usr/bin/env python
"""Copy-only probe of the box's host<->device fabric: the e2e leg's buffers (same sizes, pinned), both directions
busy at once on two streams, NO kernels.  What N GPUs can move together is the ceiling of `e2e` at that N.

    python profiles/copy_probe.py                                   # N = 1
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        profiles/copy_probe.py

One JSON line (rank 0): per-direction and combined GB/s, summed over ranks, max time over ranks."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist


def main():
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')  # stdout carries the JSON line only
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    h2d_bytes = int(float(os.environ.get('PROBE_H2D_GB', '1.92')) * 1e9)
    d2h_bytes = int(float(os.environ.get('PROBE_D2H_GB', '2.07')) * 1e9)
    steps = int(os.environ.get('PROBE_STEPS', '8'))
    hin = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
    hout = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
    din = torch.empty(h2d_bytes, dtype=torch.uint8, device='cuda')
    dout = torch.zeros(d2h_bytes, dtype=torch.uint8, device='cuda')
    s_up, s_down = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(n, up=True, down=True):
        barrier()
        t0 = time.perf_counter()
        for _ in range(n):
            if up:
                with torch.cuda.stream(s_up):
                    din.copy_(hin, non_blocking=True)
            if down:
                with torch.cuda.stream(s_down):
                    hout.copy_(dout, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        v = torch.tensor([dt], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
        return float(v.item())

    run(2)
    out = {'n_gpus': world, 'steps': steps, 'h2d_bytes_per_step': h2d_bytes, 'd2h_bytes_per_step': d2h_bytes}
    for name, up, down in (('h2d_only', True, False), ('d2h_only', False, True), ('both', True, True)):
        dt = run(steps, up, down)
        moved = steps * world * ((h2d_bytes if up else 0) + (d2h_bytes if down else 0))
        out[name + '_gbs'] = moved / dt / 1e9
    # what the e2e leg's samples/s would be if the copies were all there is (5.0 B/sample in + 5.4 out at 89 k samples/read)
    out['e2e_ceiling_samples_per_s_at_499KB_per_read'] = out['both_gbs'] * 1e9 / 499e3 * 88.7e3
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
