"""Batched public API: many parameter sets of one system layout per launch, sharded over GPUs.

    lay = batch.layout_of(system, q)                       # lower once
    I = batch.compute_intensity_batch(lay, q, params)      # [B, n_q]
    chi2 = batch.evaluate_residual_batch(lay, q, Iobs, params)
    res = batch.fit_batch(lay, q, Iobs, params0)

Multi-GPU: the batch is independent per row, so each rank of a torch.distributed job takes a
contiguous slice (shard_bounds) and the results meet in one all_gather at the end.
"""
import numpy as np

from . import engine as _engine, fitting, lowering


def layout_of(system, q):
    return lowering.lower_system(system, q)


def shard_bounds(n, rank, world):
    """Contiguous [lo, hi) slice of n items for `rank` of `world`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def compute_intensity_batch(layout, q, params, device_out=False, eng=None):
    """I[b, :] for every row of params ([B, n_params]).  numpy out unless device_out."""
    eng = eng or _engine.get_engine()
    I = eng.intensity(layout, q, params)
    return I if device_out else I.cpu().numpy()


def evaluate_residual_batch(layout, q, Iobs, params, dI=None, error_weighted=True, logI_weighted=True,
                            q_range=(0., float('inf')), device_out=False, eng=None):
    eng = eng or _engine.get_engine()
    obj = eng.objective(error_weighted, logI_weighted, q_range)
    chi2 = eng.residual(layout, obj, q, params, Iobs, dI)
    return chi2 if device_out else chi2.cpu().numpy()


def fit_batch(layout, q, Iobs, params0, **kw):
    return fitting.lm_fit_batch(layout, q, Iobs, params0, **kw)


def gather_rows(local, n_total, group=None):
    """all_gather of per-rank row blocks produced with shard_bounds() into the full [n_total, ...] tensor.
    Works on any backend (nccl on GPUs, gloo on CPU); blocks may differ by one row, so they are padded."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world == 1:
        return local
    rows = max(shard_bounds(n_total, r, world)[1] - shard_bounds(n_total, r, world)[0] for r in range(world))
    pad = torch.zeros((rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = torch.empty((world * rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    parts = []
    for r in range(world):
        lo, hi = shard_bounds(n_total, r, world)
        parts.append(out[r * rows:r * rows + (hi - lo)])
    return torch.cat(parts, dim=0)


def compute_intensity_sharded(layout, q, params, group=None, gather=True, eng=None):
    """Each rank evaluates its slice of `params` (full matrix given on every rank); optional final gather."""
    import torch.distributed as dist
    rank, world = (dist.get_rank(group), dist.get_world_size(group)) if dist.is_initialized() else (0, 1)
    lo, hi = shard_bounds(len(params), rank, world)
    eng = eng or _engine.get_engine()
    local = eng.intensity(layout, q, params[lo:hi])
    if gather and world > 1:
        return gather_rows(local, len(params), group)
    return local
