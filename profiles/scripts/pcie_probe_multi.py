"""Host<->device copy bandwidth with ALL ranks copying at once (torchrun, one rank per GPU, pinned memory):
the host-side bound of the end-to-end (host pointer) numbers at N GPUs.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 profiles/scripts/pcie_probe_multi.py"""
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
lr = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
N = 1 << 27  # 1 GiB of doubles
h = [torch.empty(N, dtype=torch.float64, pin_memory=True) for _ in range(2)]
d = [torch.empty(N, dtype=torch.float64, device=dev) for _ in range(2)]
s = [torch.cuda.Stream() for _ in range(2)]
for t in h:
    t.zero_()


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def run(fn, nbytes, label):
    fn()
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dt = torch.tensor([(time.perf_counter() - t0) / 3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("%-40s %7.1f GB/s per GPU, %7.1f GB/s aggregate over %d GPUs" %
              (label, nbytes / dt.item() / 1e9, world * nbytes / dt.item() / 1e9, world), flush=True)


B = N * 8
run(lambda: d[0].copy_(h[0], non_blocking=True), B, "H2D, every rank at once")
run(lambda: h[0].copy_(d[0], non_blocking=True), B, "D2H, every rank at once")


def bidir():
    with torch.cuda.stream(s[0]):
        d[0].copy_(h[0], non_blocking=True)
    with torch.cuda.stream(s[1]):
        h[1].copy_(d[1], non_blocking=True)


run(bidir, 2 * B, "H2D + D2H concurrent, every rank")
if world > 1:
    dist.destroy_process_group()
