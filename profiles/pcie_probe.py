"""Concurrent device->host bandwidth of the GPUs of one box (torchrun, one rank per GPU): the ceiling of the
host-buffer (e2e) path, measured with plain pinned-memory copies of one cfg2 result (328 MB)."""
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xrsdkit_b200 import batch as xb
xb.bind_to_gpu_numa(local)
n = 10000 * 4096
d = torch.empty(n, dtype=torch.float64, device="cuda")
h = torch.empty(n, dtype=torch.float64).pin_memory()


def run(active):
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if active:
        for _ in range(10):
            h.copy_(d, non_blocking=True)
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gbs = torch.tensor([10 * n * 8 / dt / 1e9 if active else 0.0], device="cuda")
    out = [torch.zeros_like(gbs) for _ in range(world)]
    dist.all_gather(out, gbs)
    return [round(float(o.item()), 1) for o in out]


run(True)
allr = run(True)
solo = run(rank == 0)
if rank == 0:
    print("D2H GB/s per GPU, all %d GPUs copying at once: %s  (sum %.1f)" % (world, allr, sum(allr)))
    print("D2H GB/s, GPU 0 alone: %.1f" % solo[0])
dist.destroy_process_group()
