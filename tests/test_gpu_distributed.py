"""Multi-GPU path on real devices: shard the batch over 2 ranks (NCCL), evaluate, one final gather.
Skipped on single-GPU boxes (the plumbing itself is covered on CPU by tests/test_distributed_cpu.py)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q_out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from helpers import cfg2_system, cfg2_params, cfg4_system, cfg4_params
        from xrsdkit_b200 import batch
        from xrsdkit_b200.system import System
        ok = True
        for make, pars, n, nq in ((cfg2_system, cfg2_params, 1001, 1024), (cfg4_system, cfg4_params, 33, 2048)):
            s = make(System)
            q = np.linspace(0.01, 0.9, nq)
            lay = batch.layout_of(s, q)
            P = pars(lay, n)
            full = batch.compute_intensity_sharded(lay, q, P, gather=True)
            lo, hi = batch.shard_bounds(n, rank, world)
            local = batch.compute_intensity_batch(lay, q, P[lo:hi], device_out=True)
            ok = ok and full.shape == (n, nq) and bool((full[lo:hi] == local).all())
            if rank == 0:
                ref = batch.compute_intensity_batch(lay, q, P, device_out=True)
                ok = ok and bool((full == ref).all())
        q_out.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_sharded_intensity_with_nccl_gather():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert sorted(q.get(timeout=5) for _ in range(2)) == [(0, True), (1, True)]
