"""world_size-2 gloo test of the multi-rank plumbing (shard -> local work -> one all_gather)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from xrsdkit_b200 import batch


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(n * 3, dtype=torch.float64).reshape(n, 3)
        lo, hi = batch.shard_bounds(n, rank, world)
        local = full[lo:hi] * 2.0                     # stand-in for the per-rank kernel output
        out = batch.gather_rows(local, n)
        assert out.shape == (n, 3) and torch.equal(out, full * 2.0)
        t = torch.tensor([float(hi - lo)])            # max-over-ranks timing reduction used by bench.py
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert t.item() == max(batch.shard_bounds(n, r, world)[1] - batch.shard_bounds(n, r, world)[0] for r in range(world))
        q.put((rank, "ok"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [7, 10])
def test_shard_and_gather_world2(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(q.get(timeout=5) for _ in range(2)) == [(0, "ok"), (1, "ok")]
