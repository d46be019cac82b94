"""Batched public API: many parameter sets of one system layout per launch, sharded over GPUs.

    lay = batch.layout_of(system, q)                       # lower once
    I = batch.compute_intensity_batch(lay, q, params)      # [B, n_q]
    chi2 = batch.evaluate_residual_batch(lay, q, Iobs, params)
    res = batch.fit_batch(lay, q, Iobs, params0)

Multi-GPU: the batch is independent per row, so each rank of a torch.distributed job takes a
contiguous slice (shard_bounds) and the results meet in one all_gather at the end.
"""
import os

from . import engine as _engine, fitting, lowering


def layout_of(system, q):
    return lowering.lower_system(system, q)


def shard_bounds(n, rank, world):
    """Contiguous [lo, hi) slice of n items for `rank` of `world`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def bind_to_gpu_numa(device_index):
    """Pin this process to the CPUs NVML reports as local to the GPU (its NUMA node), so that pinned host
    buffers allocated afterwards sit next to the GPU's PCIe root.  Matters for the host-buffer path when
    several ranks copy 100s of MB per step at once.  Returns the CPU set, or None when NVML is unavailable."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = set(w * 64 + b for w, word in enumerate(mask) for b in range(64) if (word >> b) & 1)
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return allowed
    except Exception:
        pass
    return None


def compute_intensity_batch(layout, q, params, device_out=False, eng=None, precision="fp64"):
    """I[b, :] for every row of params ([B, n_params]).  numpy out unless device_out.  precision="fp32": the optional
    single-precision path (size-distribution loops in FP32, within 1e-5 of the default; no crystalline populations)."""
    eng = eng or _engine.get_engine()
    I = eng.intensity(layout, q, params, precision=precision)
    return I if device_out else I.cpu().numpy()


def evaluate_residual_batch(layout, q, Iobs, params, dI=None, error_weighted=True, logI_weighted=True,
                            q_range=(0., float('inf')), device_out=False, eng=None):
    eng = eng or _engine.get_engine()
    obj = eng.objective(error_weighted, logI_weighted, q_range)
    chi2 = eng.residual(layout, obj, q, params, Iobs, dI)
    return chi2 if device_out else chi2.cpu().numpy()


def rescale_I0_batch(layout, q, I, params, eng=None):
    """I0 rescaling of models/predict.py:372-378 for a batch: factor[b] = sum(I[b]) / sum(I_computed[b]) multiplies the
    noise I0 and every population's I0 of row b (all populations are linear in their I0, so the rescaled systems
    reproduce sum(I) exactly up to rounding).  I: [B, n_q] or [n_q].  Returns (rescaled params, factors) as numpy."""
    eng = eng or _engine.get_engine()
    d_p = eng.to_device(params)
    if d_p.dim() == 1:
        d_p = d_p.unsqueeze(0)
    d_p = d_p.clone()
    I_comp = eng.intensity(layout, q, d_p)
    d_I = eng.to_device(I)
    factor = d_I.reshape(-1, d_I.shape[-1]).sum(dim=1) / I_comp.sum(dim=1)
    slots = [int(p.p_I0) for p in layout.c.pops[:layout.c.n_pops] if p.p_I0 >= 0]
    d_p[:, slots] *= factor[:, None]
    return d_p.cpu().numpy(), factor.cpu().numpy()


def fit_batch(layout, q, Iobs, params0, method='lm', **kw):
    """Fit every row of params0.  method='lm' (default): batched Levenberg-Marquardt;
    method='nelder-mead': the reference's optimiser, batched, for like-for-like comparison."""
    if method in ('nelder-mead', 'nelder', 'nm'):
        return fitting.nm_fit_batch(layout, q, Iobs, params0, **kw)
    if method != 'lm':
        raise ValueError("unknown fit method %r" % (method,))
    return fitting.lm_fit_batch(layout, q, Iobs, params0, **kw)


def scan_start(layout, q, Iobs, params0, name, rel_window=0.1, n_points=21, **kw):
    """Coarse scan of one parameter (by name, e.g. 'x__a': the lattice edge) before a fit; returns the improved start
    parameters (device tensor) for fit_batch.  See fitting.scan_start."""
    return fitting.scan_start(layout, q, Iobs, params0, layout.slot(name), rel_window=rel_window, n_points=n_points, **kw)


def fit_mixed(groups, eng=None, **kw):
    """Fit a batch that mixes system layouts (BASELINE config 5: sphere and crystal systems side by side).

    `groups` is a list of dicts with the keys of fit_batch (layout, q, Iobs, params0 and optionally free_slots, dI,
    ...): the systems of one layout each.  Every group runs its Levenberg-Marquardt loop on its own CUDA stream,
    driven by its own host thread, so the launches of the groups interleave on the GPU (the tail of one group's
    launch is filled by the other's CTAs) and one group's flag reads do not stall the other.  Returns the
    FitResults (device tensors) in the order of `groups`."""
    import threading
    eng = eng or _engine.get_engine()
    torch = eng.torch
    if len(groups) == 1:
        return [fitting.lm_fit_batch(eng=eng, return_device=True, **dict(kw, **groups[0]))]
    cur = torch.cuda.current_stream(eng.device)
    results, errors = [None] * len(groups), []

    def work(k, g, stream):
        try:
            torch.cuda.set_device(eng.device)
            with torch.cuda.stream(stream):
                stream.wait_stream(cur)
                results[k] = fitting.lm_fit_batch(eng=eng, return_device=True, **dict(kw, **g))
            stream.synchronize()
        except BaseException as e:      # re-raised in the calling thread
            errors.append(e)

    # the streams live as long as the engine: the caching allocator keeps its blocks per stream, so fresh streams for
    # every call would send every fit's workspaces (GBs) through cudaMalloc again
    pool = eng.__dict__.setdefault("_mixed_streams", [])
    while len(pool) < len(groups):
        pool.append(torch.cuda.Stream(device=eng.device))
    threads = [threading.Thread(target=work, args=(k, g, pool[k])) for k, g in enumerate(groups)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return results


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


_pinned_cache = {}


def compute_intensity_pinned(layout, q_host, params_host, out_host, chunks=None, eng=None):
    """Host-buffer path: params_host [B, n_params] and out_host [B, n_q] are PINNED host tensors.

    The batch is cut into `chunks` slices that alternate between two CUDA streams, so the
    host->device copy of slice c+1 and the device->host copy of slice c-1 overlap the kernels of
    slice c.  Returns out_host after both streams have drained."""
    eng = eng or _engine.get_engine()
    torch = eng.torch
    B, n_q = params_host.shape[0], q_host.shape[0]
    if chunks is None:       # measured on cfg2 (10^4 x 4096): 2 slices 49 GB/s of D2H, 4: 52, 8: 54.0, 16: 54.6, 32: 54.4
        chunks = int(os.environ.get("XRSD_PINNED_CHUNKS", 0)) or min(16, max(1, B // 512))
    chunks = max(1, min(chunks, B))
    rows = (B + chunks - 1) // chunks
    key = (str(eng.device), rows, n_q, params_host.shape[1])
    if key not in _pinned_cache:
        _pinned_cache.clear()
        _pinned_cache[key] = dict(
            streams=[torch.cuda.Stream(device=eng.device) for _ in range(2)],
            d_p=[torch.empty((rows, params_host.shape[1]), dtype=torch.float64, device=eng.device) for _ in range(2)],
            d_I=[torch.empty((rows, n_q), dtype=torch.float64, device=eng.device) for _ in range(2)],
            d_q=torch.empty(n_q, dtype=torch.float64, device=eng.device))
    ws = _pinned_cache[key]
    cur = torch.cuda.current_stream(eng.device)
    ws["d_q"].copy_(q_host, non_blocking=True)
    max_cells = bad = None
    if layout.n_xtal:
        # one candidate-box bound for the whole batch; the slices build their tables without reading the status
        # back (that read would drain the pipeline once per slice): overflow flags are OR-ed on the device and
        # looked at once, after the last slice
        max_cells = lowering.max_cells(layout, params_host.numpy())
        bad = ws.setdefault("bad", torch.zeros(2, dtype=torch.int32, device=eng.device))    # one flag per stream
        bad.zero_()
        ws.setdefault("tables", [None, None])
    for s in ws["streams"]:
        s.wait_stream(cur)
    for c in range(chunks):
        lo, hi = c * rows, min(B, (c + 1) * rows)
        if lo >= hi:
            break
        k = c % 2
        with torch.cuda.stream(ws["streams"][k]):
            d_p = ws["d_p"][k][:hi - lo]
            d_p.copy_(params_host[lo:hi], non_blocking=True)
            tables = None
            if layout.n_xtal:
                reuse = ws["tables"][k] if hi - lo == rows else None
                tables = eng.build_tables(layout, d_p, hi - lo, max_cells=max_cells, check=False, reuse=reuse)
                if hi - lo == rows:
                    ws["tables"][k] = tables
                bad[k:k + 1].copy_(torch.maximum(bad[k:k + 1], tables.status.max().reshape(1)))
            d_I = ws["d_I"][k][:hi - lo]
            eng.intensity(layout, ws["d_q"], d_p, tables=tables, out=d_I)
            out_host[lo:hi].copy_(d_I, non_blocking=True)
    for s in ws["streams"]:
        cur.wait_stream(s)
    cur.synchronize()
    if bad is not None and int(bad.max().item()) != 0:
        # a table outgrew the default capacities: redo the batch through the checked, self-growing path
        d_p = eng.to_device(params_host)
        tables = eng.build_tables(layout, d_p, B, params_host=params_host.numpy())
        out_host.copy_(eng.intensity(layout, ws["d_q"], d_p, tables=tables))
    return out_host
