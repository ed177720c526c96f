"""Platform ceiling of the end-to-end leg: pinned cudaMemcpyAsync host->device and device->host at the same time on
EVERY rank at once (torchrun --nproc-per-node N), the traffic pattern of the flow's host-buffer API (40 B/sample in,
44 B/sample out).  Prints one JSON line: aggregate GB/s per direction and the samples/s it would carry at 84 B/sample.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 profiles/microbench/pcie_ceiling.py
"""
import json
import os
import time

import torch
import torch.distributed as dist

world = int(os.environ.get('WORLD_SIZE', '1'))
rank = int(os.environ.get('RANK', '0'))
local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
n_in, n_out = 100_000_000, 110_000_000          # floats: 10^7 samples x 10 in, x 11 out
h_in = torch.empty(n_in, dtype=torch.float32, pin_memory=True).normal_()
h_out = torch.empty(n_out, dtype=torch.float32, pin_memory=True)
d_in = torch.empty(n_in, dtype=torch.float32, device=dev)
d_out = torch.empty(n_out, dtype=torch.float32, device=dev).normal_()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def both():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def sync_all():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


for _ in range(2):
    both()
sync_all()
reps = 5
t0 = time.perf_counter()
for _ in range(reps):
    both()
sync_all()
dt = (time.perf_counter() - t0) / reps
t = torch.tensor([dt], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
dt = float(t.item())
if rank == 0:
    print(json.dumps({'n_gpus': world, 'h2d_gbs_total': world * n_in * 4 / dt / 1e9, 'd2h_gbs_total': world * n_out * 4 / dt / 1e9,
                      'samples_per_s_at_84B': world * 10_000_000 / dt, 'seconds_per_round': dt,
                      'what': 'pinned cudaMemcpyAsync both directions at once on all ranks, 400 MB in + 440 MB out per rank per round'}))
if world > 1:
    dist.destroy_process_group()
