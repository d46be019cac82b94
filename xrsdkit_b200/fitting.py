"""Batched Levenberg-Marquardt fitting on the GPU.

Replaces the optimiser loop of the reference's fit() (system/__init__.py:220-278: lmfit
Nelder-Mead, one full compute_intensity per simplex evaluation with the reflection table
rebuilt every time).  The objective is the reference's scalar chi^2 (evaluate_residual); its
residual vector r_i = sqrt(w_i / sum w) (f(I_model) - f(I_obs)), f = log or identity, is what
LM linearises.  One iteration for the whole batch is five launches:

    tables(params) -> jacobian (n_free+1 evaluations per system, fused to JTJ/JTr/chi2)
    -> lm_propose (damped solve, projection on the bounds) -> tables(trial) + residual(trial)
    -> lm_update (accept / reject, damping update, convergence flags)

No host synchronisation is needed inside the loop except the reflection-table capacity check.
"""
import ctypes as C

import numpy as np

from . import abi, engine as _engine, lowering


class FitResult(object):
    def __init__(self, params, chi2_init, chi2, n_iter, n_accept, active):
        self.params, self.chi2_init, self.chi2 = params, chi2_init, chi2
        self.n_iter, self.n_accept, self.active = n_iter, n_accept, active

    @property
    def converged(self):
        return ~self.active


def lm_fit_batch(layout, q, Iobs, params0, free_slots=None, dI=None, error_weighted=True, logI_weighted=True,
                 q_range=(0., float('inf')), max_iter=30, ftol=1e-8, lambda0=1e-2, up=8.0, down=4.0,
                 lambda_max=1e10, rel_step=1e-6, lo=None, hi=None, eng=None, return_device=False, ties=None):
    """Fit every row of params0 ([batch, n_params]) to the matching row of Iobs ([batch, n_q] or [n_q]).

    Returns a FitResult of device tensors (return_device=True) or numpy arrays."""
    eng = eng or _engine.get_engine()
    torch = eng.torch
    lib = eng.lib
    if free_slots is None:
        free_slots = layout.free_slots()
    if ties is None:
        ties = layout.ties()
    free_slots = [s for s in free_slots if s not in [d for d, _ in ties]]      # a tied parameter is not varied
    tie_c, n_tie = eng.tie_array(ties)
    nf = len(free_slots)
    if nf == 0:
        raise ValueError("no free parameters to fit")
    if nf > abi.MAX_FREE:
        raise NotImplementedError("at most %d free parameters per system" % abi.MAX_FREE)
    d_q = eng.to_device(q)
    host_p0 = None if isinstance(params0, torch.Tensor) else np.atleast_2d(np.asarray(params0, dtype=np.float64))
    d_p = eng.to_device(params0 if host_p0 is None else host_p0).clone()
    if d_p.dim() == 1:
        d_p = d_p.unsqueeze(0)
    batch, n_q = d_p.shape[0], d_q.shape[0]
    for dst, src in ties:                       # lmfit evaluates the expression before the first objective call
        d_p[:, dst] = d_p[:, src]
    d_I, d_dI, ld_obs = eng._obs(Iobs, dI, batch, n_q)
    obj = eng.objective(error_weighted, logI_weighted, q_range, has_dI=d_dI is not None)
    blo, bhi = layout.bounds(free_slots)
    if lo is not None:
        blo = np.maximum(blo, lo)
    if hi is not None:
        bhi = np.minimum(bhi, hi)
    d_lo = eng.to_device(np.broadcast_to(blo, (batch, nf)).copy())
    d_hi = eng.to_device(np.broadcast_to(bhi, (batch, nf)).copy())
    d_lambda = torch.full((batch,), float(lambda0), dtype=torch.float64, device=eng.device)
    d_active = torch.ones(batch, dtype=torch.uint8, device=eng.device)
    d_nacc = torch.zeros(batch, dtype=torch.int32, device=eng.device)
    d_trial = torch.empty_like(d_p)
    d_chi2_t = torch.empty(batch, dtype=torch.float64, device=eng.device)
    jout = (torch.empty((batch, nf, nf), dtype=torch.float64, device=eng.device),
            torch.empty((batch, nf), dtype=torch.float64, device=eng.device),
            torch.empty(batch, dtype=torch.float64, device=eng.device))
    idx = (C.c_int32 * nf)(*free_slots)
    # candidate-box bound that holds for the whole fit: lattice edges can grow up to their upper
    # bound or, if unbounded, 1.5x the start
    max_cells = None
    if layout.n_xtal:
        hp = host_p0 if host_p0 is not None else d_p.cpu().numpy()
        grown = hp.copy()
        for j, s in enumerate(free_slots):
            grown[:, s] = np.where(np.isfinite(bhi[j]), np.minimum(bhi[j], 1.5 * np.abs(hp[:, s])), 1.5 * np.abs(hp[:, s]))
        max_cells = max(lowering.max_cells(layout, hp), lowering.max_cells(layout, grown))
        max_cells = min(max_cells, 90000)
    tables = eng.build_tables(layout, d_p, batch, max_cells=max_cells) if layout.n_xtal else None
    chi2_init = None
    t_tables = None
    res_ws = eng.residual_workspace(batch)
    stream = eng._stream()
    n_iter = 0
    for it in range(max_iter):
        n_iter = it + 1
        _, _, d_chi2 = eng.jacobian(layout, obj, d_q, d_p, d_I, d_dI, ld_obs, free_slots, tables, rel_step, out=jout,
                                    active=d_active if it > 0 else None, ties=ties)
        if chi2_init is None:
            chi2_init = d_chi2.clone()
        abi.check(lib.xrsd_lm_propose(d_p.data_ptr(), d_p.shape[1], batch, idx, nf, tie_c, n_tie, jout[0].data_ptr(), jout[1].data_ptr(),
                                      d_lambda.data_ptr(), d_lo.data_ptr(), d_hi.data_ptr(), d_active.data_ptr(),
                                      d_trial.data_ptr(), stream))
        # in-loop table builds do not read their status back (no host sync); it is checked every 5 iterations
        t_tables = eng.build_tables(layout, d_trial, batch, max_cells=max_cells, reuse=t_tables,
                                    check=t_tables is None, active=d_active if t_tables is not None else None) if layout.n_xtal else None
        abi.check(lib.xrsd_residual_batch(C.byref(layout.c), C.byref(obj), d_q.data_ptr(), n_q, d_trial.data_ptr(),
                                          d_trial.shape[1], batch, C.byref(t_tables.c) if t_tables is not None else None,
                                          d_I.data_ptr(), d_dI.data_ptr() if d_dI is not None else None, ld_obs,
                                          d_chi2_t.data_ptr(), res_ws.data_ptr(), d_active.data_ptr(), stream))
        abi.check(lib.xrsd_lm_update(d_p.data_ptr(), d_p.shape[1], batch, d_trial.data_ptr(), d_chi2.data_ptr(),
                                     d_chi2_t.data_ptr(), d_lambda.data_ptr(), d_active.data_ptr(), d_nacc.data_ptr(),
                                     float(up), float(down), float(ftol), float(lambda_max), stream))
        eng.launches += 3
        if layout.n_xtal:
            # accepted systems now sit at `trial`: their tables are the trial tables
            tables = eng.build_tables(layout, d_p, batch, max_cells=max_cells, reuse=tables, check=False, active=d_active)
        if it % 5 == 4:
            flags = torch.stack([d_active.any().to(torch.int32),
                                 tables.status.max() if tables is not None else torch.zeros((), dtype=torch.int32, device=eng.device),
                                 t_tables.status.max() if t_tables is not None else torch.zeros((), dtype=torch.int32, device=eng.device)])
            any_active, st_a, st_b = [int(v) for v in flags.cpu().tolist()]      # the only host sync of the loop
            if st_a or st_b:
                # a table outgrew its storage since the last check: rebuild both sets with the checked
                # (self-growing) path and resynchronise chi2 with the accepted parameters
                tables = eng.build_tables(layout, d_p, batch, max_cells=max_cells)
                t_tables = None
                d_chi2.copy_(eng.residual(layout, obj, d_q, d_p, d_I, d_dI, tables=tables))
            if not any_active:
                break
    # final objective at the accepted parameters (tables rebuilt for every system: converged ones were
    # skipped by the masked in-loop builds after their last accepted step)
    if layout.n_xtal:
        tables = eng.build_tables(layout, d_p, batch, max_cells=max_cells, reuse=tables)
    chi2 = eng.residual(layout, obj, d_q, d_p, d_I, d_dI, tables=tables)
    res = FitResult(d_p, chi2_init, chi2, n_iter, d_nacc, d_active.bool())
    if return_device:
        return res
    return FitResult(d_p.cpu().numpy(), chi2_init.cpu().numpy(), chi2.cpu().numpy(), n_iter,
                     d_nacc.cpu().numpy(), d_active.bool().cpu().numpy())


def fit_system_inplace(system, q, I, dI=None):
    """fit() backend: optimise `system`'s free parameters in place and fill its fit_report with the
    keys the reference writes (system/__init__.py:268-275)."""
    eng = _engine.get_engine()
    lay = lowering.lower_system(system, q)
    rep = system.fit_report
    free = lay.free_slots()
    if dI is not None and np.ndim(dI) == 0:
        dI = np.full(q.shape, float(dI))      # the reference's own examples pass a scalar here
    if not free:
        obj0 = system.evaluate_residual(q, I, dI)
        rep.update(converged=True, initial_objective=obj0, final_objective=obj0)
        return system
    res = lm_fit_batch(lay, q, I, lay.values[None, :], free, dI=dI, error_weighted=rep['error_weighted'],
                       logI_weighted=rep['logI_weighted'], q_range=rep['q_range'], max_iter=60, eng=eng)
    tied = [d for d, _ in lay.ties()]
    for slot, rec in enumerate(lay.records):
        if slot in free or slot in tied:
            rec['value'] = float(res.params[0, slot])
    I_opt = system.compute_intensity(q)
    I_bg = I - I_opt
    rep['converged'] = bool(res.converged[0])
    rep['initial_objective'] = float(res.chi2_init[0])
    rep['final_objective'] = float(res.chi2[0])
    rep['fit_snr'] = float(np.mean(I_opt) / np.std(I_bg))
    return system
