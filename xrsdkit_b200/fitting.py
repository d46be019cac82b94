"""Batched Levenberg-Marquardt fitting on the GPU.

Replaces the optimiser loop of the reference's fit() (system/__init__.py:220-278: lmfit
Nelder-Mead, one full compute_intensity per simplex evaluation with the reflection table
rebuilt every time).  The objective is the reference's scalar chi^2 (evaluate_residual); its
residual vector r_i = sqrt(w_i / sum w) (f(I_model) - f(I_obs)), f = log or identity, is what
LM linearises.  lm_fit_batch describes the loop: every launch of an iteration is one of the
library's own kernels (Jacobian variants + reduction for the systems whose last step was
accepted, damped solve with a relative trust region, shared reflection tables, trial pattern,
fused chi2 / accept-reject, accepted rows taken over), the host reads one flag word every few
iterations.  nm_fit_batch is the reference's own optimiser (Nelder-Mead), batched, for
like-for-like comparison and for parameters LM cannot differentiate (atomic coordinates).
"""
import ctypes as C
import os

import numpy as np

from . import abi, engine as _engine, lowering


class FitResult(object):
    """params / chi2_init / chi2 / n_accept / active / reason per system.  reason: 0 = stopped by the iteration cap,
    1 = converged (ftol / xtol), 2 = no acceptable step left (damping beyond lambda_max), 3 = objective not finite."""

    def __init__(self, params, chi2_init, chi2, n_iter, n_accept, active, reason=None):
        self.params, self.chi2_init, self.chi2 = params, chi2_init, chi2
        self.n_iter, self.n_accept, self.active = n_iter, n_accept, active
        self.reason = reason

    @property
    def converged(self):
        """True where the optimiser stopped on its tolerance (the reference reports lmfit's `success`)."""
        if self.reason is None:
            return ~self.active
        return self.reason == 1


def _apply_ties(params, ties):
    """The constraint terms of SystemLayout.ties() on a parameter matrix (numpy array or torch tensor, [..., n_params]),
    in order -- the host-side twin of apply_ties() in csrc/fit_kernels.cu: linear terms dst = scale * src + offset
    (dst given as -(slot + 1): dst += ...), and the words of postfix expression programs over a value stack."""
    if not ties:
        return
    is_np = isinstance(params, np.ndarray)
    xp = np
    if not is_np:
        import torch as xp
    f1 = {'neg': lambda a: -a, 'sqrt': xp.sqrt, 'exp': xp.exp, 'log': xp.log, 'log10': xp.log10, 'sin': xp.sin, 'cos': xp.cos,
          'tan': xp.tan, 'abs': xp.abs, 'asin': xp.arcsin if is_np else xp.asin, 'acos': xp.arccos if is_np else xp.acos,
          'atan': xp.arctan if is_np else xp.atan}
    f2 = {'add': lambda a, b: a + b, 'sub': lambda a, b: a - b, 'mul': lambda a, b: a * b, 'div': lambda a, b: a / b,
          'pow': lambda a, b: a ** b, 'min': xp.minimum, 'max': xp.maximum}
    one = params[..., 0] * 0.0 + 1.0            # constants are pushed as arrays so that every function applies
    stack = []
    for t in ties:
        scale, off = (t[2], t[3]) if len(t) >= 4 else (1.0, 0.0)
        if t[0] <= -abi.TIE_WORD:
            op = abi.TIE_OPS[-t[0] - abi.TIE_WORD]
            if op == 'pushc':
                stack.append(one * scale)
            elif op == 'pushp':
                stack.append(params[..., t[1]] * 1.0)
            elif op in f2:
                b = stack.pop()
                a = stack.pop()
                stack.append(f2[op](a, b))
            else:
                stack.append(f1[op](stack.pop()))
        elif t[1] < 0:                          # store the value of the expression
            params[..., t[0]] = stack.pop() * scale + off
            stack = []
        elif t[0] < 0:
            params[..., -t[0] - 1] += params[..., t[1]] * scale + off
        else:
            params[..., t[0]] = params[..., t[1]] * scale + off


def lm_fit_batch(layout, q, Iobs, params0, free_slots=None, dI=None, error_weighted=True, logI_weighted=True,
                 q_range=(0., float('inf')), max_iter=30, ftol=1e-8, xtol=1e-10, lambda0=1e-2, up=8.0, down=4.0,
                 lambda_max=1e10, rel_step=1e-6, lo=None, hi=None, eng=None, return_device=False, ties=None,
                 resident_limit=24e9, min_group_rows=2048, table_growth=1.09, loop_cells=None, max_rel_step=0.25,
                 check_every=5, share_tables=True):
    """Fit every row of params0 ([batch, n_params]) to the matching row of Iobs ([batch, n_q] or [n_q]).

    One iteration = one trial step per system; every launch is masked, and nothing but the kernels of
    include/xrsd_b200.h runs inside the loop:

        jacobian        (systems whose last step was accepted: variants + reduction -> JTJ, JTr, chi2)
        lm_propose      (damped solve, relative trust region, projection on the bounds -> trial parameters)
        tables(trial)   (crystalline layouts; shared tables make this a key lookup after the first iterations)
        trial pattern   (patterns-only variants launch: pattern + running totals in the trial workspace)
        lm_trial_update (chi2 of the stored trial pattern, accept / reject, damping, flags)
        lm_accept       (accepted systems: trial pattern -> base rows, trial tables -> the system's tables)

    A rejected step leaves the parameters, and therefore the Jacobian, as they are: the next iteration only solves
    again with more damping and evaluates the new trial point.  The host reads three flag words every `check_every`
    iterations (all converged? a table outgrown?).  Nothing in an iteration depends on host state (the iteration number
    lives on the device, xrsd_lm_propose counts it), so a resident fit issues each iteration through ONE native call
    (xrsd_lm_iteration).

    When the pattern slabs of the batch fit in `resident_limit` bytes they stay on the device for the whole fit;
    a larger batch is fitted in row groups that fit (systems are independent; a group of `min_group_rows` systems
    or more still fills the GPU many times over); if even such a group does not fit, the loop re-evaluates the
    base inside every Jacobian (shared 4 GB scratch) and the trial point through the residual kernel.

    Crystalline layouts whose tables cannot be shared build the trial tables with the candidate box of the current
    parameters grown by `table_growth` (`loop_cells` sets the first box explicitly); lattice steps beyond it are
    turned back and the box follows the accepted parameters (share_tables=False forces that path for any layout).
    `max_rel_step` is the relative trust region of a step (see xrsd_lm_propose; on BASELINE config 5 -- start values
    10 % off -- 0.25 takes the crystal systems from 30 % to 90 % success within 30 iterations,
    profiles/r02b_fit_quality.json).

    Returns a FitResult of device tensors (return_device=True) or numpy arrays."""
    eng = eng or _engine.get_engine()
    torch = eng.torch
    lib = eng.lib
    if free_slots is None:
        free_slots = layout.free_slots()
    if ties is None:
        ties = layout.ties()
    free_slots = [s for s in free_slots if s not in lowering.tied_slots(ties)]      # a tied parameter is not varied
    tie_c, coef_c, n_tie = eng.tie_array(ties)
    nf = len(free_slots)
    if nf == 0:
        raise ValueError("no free parameters to fit")
    if nf > abi.MAX_FREE:
        raise lowering.CapacityError("at most %d free parameters per system (XRSD_MAX_FREE); this fit has %d" % (abi.MAX_FREE, nf))
    if layout.coordinate_slots() & set(free_slots):
        # Free atomic coordinates (u_i, v_i, w_i of a polyatomic cell) act through the reflection tables, which the
        # Jacobian variants hold fixed: their finite-difference columns would be identically zero and LM
        # would never move them.  Such fits go to the batched Nelder-Mead mode, which rebuilds the tables for every
        # evaluation -- the reference's own optimiser.
        return nm_fit_batch(layout, q, Iobs, params0, free_slots=free_slots, dI=dI, error_weighted=error_weighted,
                            logI_weighted=logI_weighted, q_range=q_range, eng=eng, return_device=return_device, ties=ties)
    resident_limit = float(os.environ.get("XRSD_LM_RESIDENT_LIMIT", resident_limit))
    n_rows, n_pts = (int(np.shape(params0)[0]) if np.ndim(params0) == 2 else 1), int(np.shape(q)[0])
    n_pops = int(layout.c.n_pops)

    def resident_bytes(rows):
        return int(lib.xrsd_jacobian_workspace_bytes(n_pts, rows, nf)) + int(lib.xrsd_jacobian_workspace_bytes(n_pts, rows, 0))

    if resident_bytes(n_rows) > resident_limit:
        fit_rows = int(resident_limit // (resident_bytes(n_rows) / float(n_rows)))
        if fit_rows >= max(1, min_group_rows):
            n_groups = -(-n_rows // fit_rows)
            step = -(-n_rows // n_groups)

            def rows_of(x, a, b):      # per-system arrays are cut, shared ones ([n_q], scalars, None) pass through
                return x[a:b] if x is not None and np.ndim(x) == 2 and np.shape(x)[0] == n_rows else x

            parts = [lm_fit_batch(layout, q, rows_of(Iobs, a, min(a + step, n_rows)), params0[a:min(a + step, n_rows)],
                                  free_slots=free_slots, dI=rows_of(dI, a, min(a + step, n_rows)),
                                  error_weighted=error_weighted, logI_weighted=logI_weighted, q_range=q_range,
                                  max_iter=max_iter, ftol=ftol, xtol=xtol, lambda0=lambda0, up=up, down=down, lambda_max=lambda_max,
                                  rel_step=rel_step, lo=rows_of(lo, a, min(a + step, n_rows)),
                                  hi=rows_of(hi, a, min(a + step, n_rows)), eng=eng, return_device=return_device, ties=ties,
                                  resident_limit=resident_limit, min_group_rows=min_group_rows, table_growth=table_growth,
                                  loop_cells=loop_cells, max_rel_step=max_rel_step, check_every=check_every,
                                  share_tables=share_tables)
                     for a in range(0, n_rows, step)]
            cat = torch.cat if return_device else np.concatenate
            return FitResult(cat([r.params for r in parts]), cat([r.chi2_init for r in parts]), cat([r.chi2 for r in parts]),
                             max(r.n_iter for r in parts), cat([r.n_accept for r in parts]), cat([r.active for r in parts]),
                             cat([r.reason for r in parts]))
    d_q = eng.to_device(q)
    host_p0 = None if isinstance(params0, torch.Tensor) else np.atleast_2d(np.asarray(params0, dtype=np.float64))
    d_p = eng.to_device(params0 if host_p0 is None else host_p0).clone()
    if d_p.dim() == 1:
        d_p = d_p.unsqueeze(0)
    batch, n_q = d_p.shape[0], d_q.shape[0]
    _apply_ties(d_p, ties)                      # lmfit evaluates the expression before the first objective call
    d_I, d_dI, ld_obs = eng._obs(Iobs, dI, batch, n_q)
    obj_raw = eng.objective(error_weighted, logI_weighted, q_range, has_dI=d_dI is not None)
    obj, d_fo, d_w = eng.prepare_objective(obj_raw, d_q, d_I, d_dI, ld_obs, batch)      # log / weights / mask once per fit
    blo, bhi = layout.bounds(free_slots)
    if lo is not None:
        blo = np.maximum(blo, lo)
    if hi is not None:
        bhi = np.minimum(bhi, hi)
    d_lo = eng.to_device(np.broadcast_to(blo, (batch, nf)).copy())
    d_hi = eng.to_device(np.broadcast_to(bhi, (batch, nf)).copy())
    u8 = torch.uint8
    d_lambda = torch.full((batch,), float(lambda0), dtype=torch.float64, device=eng.device)
    d_active = torch.ones(batch, dtype=u8, device=eng.device)
    d_needjac = torch.ones(batch, dtype=u8, device=eng.device)
    d_accepted = torch.zeros(batch, dtype=u8, device=eng.device)
    d_reason = torch.zeros(batch, dtype=u8, device=eng.device)
    d_nacc = torch.zeros(batch, dtype=torch.int32, device=eng.device)
    d_trial = d_p.clone()
    d_chi2_t = torch.empty(batch, dtype=torch.float64, device=eng.device)
    jout = (torch.empty((batch, nf, nf), dtype=torch.float64, device=eng.device),
            torch.empty((batch, nf), dtype=torch.float64, device=eng.device),
            torch.full((batch,), float('inf'), dtype=torch.float64, device=eng.device))
    # an I0 that starts at exactly 0 has no derivative in I_pop / I0: such parameters get a finite-difference variant
    hp = host_p0 if host_p0 is not None else d_p.cpu().numpy()
    jac_slots = [(-(s + 1) if (eng.jacobian_variants(layout, [s], ties) == 1 and bool(np.any(hp[:, s] == 0.0))) else s)
                 for s in free_slots]
    idx = (C.c_int32 * nf)(*free_slots)
    # candidate-box bound that holds for the whole fit: lattice edges can grow up to their upper
    # bound or, if unbounded, 1.5x the start
    max_cells = None
    tables = t_tables = None
    shareable = False
    if layout.n_xtal:
        grown = hp.copy()
        for j, s in enumerate(free_slots):
            grown[:, s] = np.where(np.isfinite(bhi[j]), np.minimum(bhi[j], 1.5 * np.abs(hp[:, s])), 1.5 * np.abs(hp[:, s]))
        max_cells = max(lowering.max_cells(layout, hp), lowering.max_cells(layout, grown))
        shareable = bool(share_tables) and eng.tables_shareable(layout)
        if not shareable:
            max_cells = min(max_cells, _engine.SMEM_CELLS)
        # Tables that cannot be shared are rebuilt for every trial point, and the builder's shared memory (2 B per
        # candidate cell) sets its occupancy: sized for 1.5x growth it runs one CTA per SM, twice as slow as with the
        # box the systems actually need.  Those builds therefore use the box of the current parameters grown by
        # `table_growth`; a trial step that does not fit is turned back like any bad step (chi2 = inf: lambda grows,
        # the next step is shorter), and when that has happened the periodic flag read below re-centres the box on
        # the parameters accepted so far, so a lattice that really has to grow by more than that does so in stages.

        def near_cells(now):
            near = now.copy()
            for j, s in enumerate(free_slots):
                near[:, s] = np.minimum(grown[:, s], float(table_growth) * np.abs(now[:, s]))
            return min(max_cells, max(lowering.max_cells(layout, now), lowering.max_cells(layout, near)))

        if shareable:
            loop_cells = max_cells
        else:
            loop_cells = near_cells(hp) if loop_cells is None else min(max_cells, int(loop_cells))
        # accepted point and trial point are two views of one store: they share the shared rows
        store = eng.new_table_store(layout, batch, max_cells, n_views=2, share=shareable)
        tables = eng.build_tables(layout, d_p, batch, max_cells=max_cells, reuse=store.views[0], share=shareable)
        store = tables.store
        t_tables = store.views[1]
    n_iter = 0
    n_var = eng.jacobian_variants(layout, jac_slots, ties)
    persist = resident_bytes(batch) <= resident_limit
    key0, key1 = ("lm", id(d_p), 0), ("lm", id(d_p), 1)
    res_ws = None if persist else eng.residual_workspace(batch, n_q)
    tight = bool(layout.n_xtal) and loop_cells < max_cells
    d_flags = torch.zeros(3, dtype=torch.int32, device=eng.device)      # [0] trial-table status bits met, [1] last iteration that
                                                                          # left a system running, [2] iteration counter (device side)
    state = dict(base_valid=False)

    def iteration():
        """The launches of one iteration (see the docstring).  Nothing in here depends on the iteration number or
        reads anything back, so from the second iteration on the same sequence is replayed as a CUDA graph."""
        eng.jacobian(layout, obj, d_q, d_p, d_fo, d_w, ld_obs, jac_slots, tables, rel_step, out=jout, active=d_needjac,
                     ties=ties, resident=key0 if persist else None, base_ready=persist and state['base_valid'])
        state['base_valid'] = persist
        if state.get('chi2_init') is None:          # first iteration only (never inside a graph capture)
            state['chi2_init'] = jout[2].clone()
        abi.check(lib.xrsd_lm_propose(d_p.data_ptr(), d_p.shape[1], batch, idx, nf, tie_c, coef_c, n_tie, jout[0].data_ptr(),
                                      jout[1].data_ptr(), d_lambda.data_ptr(), d_lo.data_ptr(), d_hi.data_ptr(),
                                      d_active.data_ptr(), d_trial.data_ptr(), float(max_rel_step),
                                      d_flags.data_ptr() + 8, eng._stream()))
        if layout.n_xtal:       # in-loop builds do not read their status back; lm_trial_update looks at it on the device
            eng.build_tables(layout, d_trial, batch, max_cells=loop_cells, reuse=t_tables, check=False, active=d_active)
        trial_pat, ld_pat = None, 0
        if persist:     # pattern and running totals of the trial point, kept in the second resident workspace
            eng.jacobian(layout, obj, d_q, d_trial, None, None, 0, [], t_tables, rel_step, active=d_active, ties=ties,
                         resident=key1, patterns_only=True)
            slabs1 = eng.jacobian_slabs(key1, layout, batch, n_q, 1)
            trial_pat, ld_pat = slabs1.data_ptr(), slabs1.stride(0)
        else:
            abi.check(lib.xrsd_residual_batch(C.byref(layout.c), C.byref(obj), d_q.data_ptr(), n_q, d_trial.data_ptr(),
                                              d_trial.shape[1], batch, C.byref(t_tables.c) if t_tables is not None else None,
                                              d_fo.data_ptr(), d_w.data_ptr(), ld_obs, d_chi2_t.data_ptr(), res_ws.data_ptr(),
                                              d_active.data_ptr(), eng._stream()))
        abi.check(lib.xrsd_lm_trial_update(C.byref(obj), d_q.data_ptr(), n_q, trial_pat, ld_pat, batch, d_fo.data_ptr(),
                                           d_w.data_ptr(), ld_obs, d_p.data_ptr(), d_p.shape[1], d_trial.data_ptr(),
                                           jout[2].data_ptr(), d_chi2_t.data_ptr(), d_lambda.data_ptr(), d_active.data_ptr(),
                                           d_accepted.data_ptr(), d_needjac.data_ptr(), d_reason.data_ptr(), d_nacc.data_ptr(),
                                           C.byref(t_tables.c) if t_tables is not None else None, layout.n_xtal,
                                           d_flags.data_ptr(), 0, float(up), float(down), float(ftol), float(xtol),
                                           float(lambda_max), eng._stream()))
        if persist or layout.n_xtal:
            ws0 = eng._jac_resident[key0][0] if persist else None
            ws1 = eng._jac_resident[key1][0] if persist else None
            abi.check(lib.xrsd_lm_accept(d_accepted.data_ptr(), batch,
                                         ws0.data_ptr() if persist else None, (n_var + n_pops) * n_q,
                                         ws1.data_ptr() if persist else None, (1 + n_pops) * n_q, n_q, n_var, n_pops,
                                         C.byref(tables.c) if tables is not None else None,
                                         C.byref(t_tables.c) if t_tables is not None else None, layout.n_xtal, eng._stream()))
        eng.launches += 3 + (0 if persist else 1)

    # Resident fits run the iteration natively: ONE call (xrsd_lm_iteration) issues the launches listed above from C, so a
    # small batch, whose iteration is a millisecond of GPU work, does not pay the host time of eight ctypes calls.
    # (Replaying the iteration as a CUDA graph was tried: capture + instantiation per fit cost more than they saved.)
    native = None
    if persist:
        ws0 = eng.resident_workspace(key0, n_q, batch, nf)
        ws1 = eng.resident_workspace(key1, n_q, batch, 0)
        state['chi2_init'] = torch.empty(batch, dtype=torch.float64, device=eng.device)
        native = abi.LmState()
        native.layout, native.obj = C.pointer(layout.c), C.pointer(obj)
        native.d_q, native.n_q = d_q.data_ptr(), n_q
        native.d_params, native.ld_params, native.batch, native.d_trial = d_p.data_ptr(), d_p.shape[1], batch, d_trial.data_ptr()
        native.d_fo, native.d_w, native.ld_obs = d_fo.data_ptr(), d_w.data_ptr(), ld_obs
        jidx = (C.c_int32 * nf)(*jac_slots)
        native.free_idx, native.jac_idx, native.n_free = C.cast(idx, C.POINTER(C.c_int32)), C.cast(jidx, C.POINTER(C.c_int32)), nf
        native.ties = C.cast(tie_c, C.POINTER(C.c_int32)) if n_tie else None
        native.tie_coef = C.cast(coef_c, C.POINTER(C.c_double)) if n_tie else None
        native.n_tie = n_tie
        native.rel_step, native.max_rel_step = float(rel_step), float(max_rel_step)
        native.up, native.down, native.ftol, native.xtol, native.lambda_max = float(up), float(down), float(ftol), float(xtol), float(lambda_max)
        native.d_JTJ, native.d_JTr, native.d_chi2, native.d_chi2_trial = jout[0].data_ptr(), jout[1].data_ptr(), jout[2].data_ptr(), d_chi2_t.data_ptr()
        native.d_lambda, native.d_lo, native.d_hi = d_lambda.data_ptr(), d_lo.data_ptr(), d_hi.data_ptr()
        native.d_active, native.d_accepted, native.d_needjac, native.d_reason = (d_active.data_ptr(), d_accepted.data_ptr(),
                                                                                   d_needjac.data_ptr(), d_reason.data_ptr())
        native.d_n_accept, native.d_flags = d_nacc.data_ptr(), d_flags.data_ptr()
        native.d_ws_base, native.d_ws_trial = ws0.data_ptr(), ws1.data_ptr()
        native.d_chi2_init = state['chi2_init'].data_ptr()
        native.n_var, native.n_pops = n_var, n_pops

        def refresh_native():
            native.tables = C.pointer(tables.c) if tables is not None else None
            native.trial_tables = C.pointer(t_tables.c) if t_tables is not None else None
            native.loop_cells = min(int(loop_cells), _engine.SMEM_CELLS) if layout.n_xtal else 0
        refresh_native()
        per_iter = 6 + (2 if layout.n_xtal else 0)
    try:
        for it in range(max_iter):
            n_iter = it + 1
            if native is not None:
                abi.check(lib.xrsd_lm_iteration(C.byref(native), 1 if it else 0, eng._stream()))
                eng.launches += per_iter
            else:
                iteration()
            if it % check_every == check_every - 1:
                # the only host sync of the loop: [0] status bits met, [1] the last iteration that left a system running
                st_seen, last_running, n_done = [int(v) for v in d_flags.cpu().tolist()]
                any_active = last_running == n_done
                if st_seen & abi.TABLE_GRID_OVERFLOW and tight:
                    # an active system proposed a cell beyond the near box and was turned back: re-centre the box
                    loop_cells = near_cells(d_p.cpu().numpy())
                    tight = loop_cells < max_cells
                    if native is not None:
                        refresh_native()
                if st_seen & (abi.TABLE_LIST_OVERFLOW | abi.TABLE_SHELL_OVERFLOW):
                    # a trial table outgrew the store's capacities (those trial points were turned back): move to a
                    # larger store -- the checked build of the accepted points grows it further if it has to
                    store = eng.new_table_store(layout, batch, max_cells, n_views=2, share=shareable,
                                                cap_shell=store.cap_shell * (4 if st_seen & abi.TABLE_SHELL_OVERFLOW else 1),
                                                cap_list=4096 if st_seen & abi.TABLE_LIST_OVERFLOW else store.cap_list_hint)
                    tables = eng.build_tables(layout, d_p, batch, max_cells=max_cells, reuse=store.views[0], share=shareable)
                    store = tables.store
                    t_tables = store.views[1]
                    if native is not None:
                        refresh_native()
                if st_seen:
                    d_flags[0:1].zero_()
                if not any_active:
                    break
        # final objective at the accepted parameters (raw objective: the same call a user makes)
        if layout.n_xtal:
            tables = eng.build_tables(layout, d_p, batch, max_cells=max_cells, reuse=tables, share=shareable)
        chi2 = eng.residual(layout, obj, d_q, d_p, d_fo, d_w, tables=tables)
    finally:
        eng.release_resident(key0, key1)
    chi2_init = state['chi2_init']
    res = FitResult(d_p, chi2_init, chi2, n_iter, d_nacc, d_active.bool(), d_reason)
    if return_device:
        return res
    return FitResult(d_p.cpu().numpy(), chi2_init.cpu().numpy(), chi2.cpu().numpy(), n_iter,
                     d_nacc.cpu().numpy(), d_active.bool().cpu().numpy(), d_reason.cpu().numpy())


def scan_start(layout, q, Iobs, params0, slot, rel_window=0.1, n_points=21, dI=None, error_weighted=True,
               logI_weighted=True, q_range=(0., float('inf')), eng=None, ties=None):
    """Coarse one-dimensional scan of ONE parameter before a fit: every system's objective is evaluated with the
    parameter at n_points values in [1 - rel_window, 1 + rel_window] x its start value (one residual launch per grid
    value for the whole batch), and the best value becomes the start.  Meant for the lattice edge of a crystalline
    population: a start that is several peak widths off puts the objective into another basin (on BASELINE config 5,
    starts whose `a` is >= 8 % off converge in 57 % of the cases, those within 6 % in > 98 %), and no local optimiser --
    this LM loop or the reference's Nelder-Mead -- gets out of it.  The reference has no such step.  Returns the new
    start parameters as a device tensor [batch, n_params]."""
    eng = eng or _engine.get_engine()
    torch = eng.torch
    if ties is None:
        ties = layout.ties()
    d_q = eng.to_device(q)
    d_p = eng.to_device(params0).clone()
    if d_p.dim() == 1:
        d_p = d_p.unsqueeze(0)
    batch, n_q = d_p.shape[0], d_q.shape[0]
    d_I, d_dI, _ = eng._obs(Iobs, dI, batch, n_q)
    obj = eng.objective(error_weighted, logI_weighted, q_range, has_dI=d_dI is not None)
    best = torch.full((batch,), float('inf'), dtype=torch.float64, device=eng.device)
    best_val = d_p[:, slot].clone()
    tables = None
    hp = d_p.cpu().numpy()
    max_cells = None
    if layout.n_xtal:
        hi = hp.copy()
        hi[:, slot] *= 1.0 + float(rel_window)
        max_cells = max(lowering.max_cells(layout, hp), lowering.max_cells(layout, hi))
    for f in np.linspace(1.0 - float(rel_window), 1.0 + float(rel_window), int(n_points)):
        P = d_p.clone()
        P[:, slot] *= float(f)
        _apply_ties(P, ties)
        if layout.n_xtal:
            tables = eng.build_tables(layout, P, batch, max_cells=max_cells, reuse=tables, check=tables is None)
        chi2 = eng.residual(layout, obj, d_q, P, d_I, d_dI, tables=tables)
        chi2 = torch.where(torch.isnan(chi2), torch.full_like(chi2, float('inf')), chi2)
        better = chi2 < best
        best = torch.where(better, chi2, best)
        best_val = torch.where(better, P[:, slot], best_val)
    d_p[:, slot] = best_val
    _apply_ties(d_p, ties)
    return d_p


def fit_system_inplace(system, q, I, dI=None, method='lm'):
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
    kw = dict(dI=dI, error_weighted=rep['error_weighted'], logI_weighted=rep['logI_weighted'], q_range=rep['q_range'], eng=eng)
    if method in ('nelder-mead', 'nelder', 'nm'):
        res = nm_fit_batch(lay, q, I, lay.values[None, :], free, **kw)
    elif method == 'lm':
        res = lm_fit_batch(lay, q, I, lay.values[None, :], free, max_iter=60, **kw)
    else:
        raise ValueError("unknown fit method %r" % (method,))
    tied = lowering.tied_slots(lay.ties())
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


# ---------------------------------------------------------------------------------------------
# Nelder-Mead, batched: the reference's own optimiser (lmfit.minimize(method='nelder-mead'),
# system/__init__.py:262-266), for like-for-like comparison with the LM fitter above.
# ---------------------------------------------------------------------------------------------
class _BoundTransform(object):
    """lmfit's MINUIT-style mapping between bounded parameter values and the unbounded internal
    variables the simplex moves in: both bounds  x = lo + (sin(u) + 1)(hi - lo)/2;
    lower only  x = lo - 1 + sqrt(u^2 + 1);  upper only  x = hi + 1 - sqrt(u^2 + 1)."""

    def __init__(self, torch, lo, hi, device):
        self.t = torch
        self.lo = torch.as_tensor(lo, dtype=torch.float64, device=device)
        self.hi = torch.as_tensor(hi, dtype=torch.float64, device=device)
        fin_lo, fin_hi = torch.isfinite(self.lo), torch.isfinite(self.hi)
        self.both, self.only_lo, self.only_hi = fin_lo & fin_hi, fin_lo & ~fin_hi, ~fin_lo & fin_hi

    def to_internal(self, x):
        t = self.t
        lo = torch_nan_to(self.t, self.lo)
        hi = torch_nan_to(self.t, self.hi)
        u = x.clone()
        arg = (2 * (x - lo) / (hi - lo) - 1).clamp(-1, 1)
        u = t.where(self.both, t.asin(arg), u)
        u = t.where(self.only_lo, t.sqrt(((x - lo + 1) ** 2 - 1).clamp(min=0)), u)
        u = t.where(self.only_hi, t.sqrt(((hi - x + 1) ** 2 - 1).clamp(min=0)), u)
        return u

    def to_external(self, u):
        t = self.t
        lo = torch_nan_to(self.t, self.lo)
        hi = torch_nan_to(self.t, self.hi)
        x = u.clone()
        x = t.where(self.both, lo + (t.sin(u) + 1) * (hi - lo) / 2, x)
        x = t.where(self.only_lo, lo - 1 + t.sqrt(u * u + 1), x)
        x = t.where(self.only_hi, hi + 1 - t.sqrt(u * u + 1), x)
        return x


def torch_nan_to(torch, v):
    return torch.where(torch.isfinite(v), v, torch.zeros_like(v))


def nm_fit_batch(layout, q, Iobs, params0, free_slots=None, dI=None, error_weighted=True, logI_weighted=True,
                 q_range=(0., float('inf')), max_iter=None, xatol=1e-4, fatol=1e-4, eng=None, return_device=False,
                 ties=None):
    """Batched Nelder-Mead on the reference's scalar objective (scipy's algorithm with its default
    coefficients and initial simplex, lmfit's bound transform).  Every objective evaluation of the whole
    batch is one residual launch (plus a masked table build for crystalline layouts); the simplex
    bookkeeping is elementwise tensor work on the device."""
    eng = eng or _engine.get_engine()
    torch = eng.torch
    if free_slots is None:
        free_slots = layout.free_slots()
    if ties is None:
        ties = layout.ties()
    free_slots = [s for s in free_slots if s not in lowering.tied_slots(ties)]
    nf = len(free_slots)
    if nf == 0:
        raise ValueError("no free parameters to fit")
    d_q = eng.to_device(q)
    host_p0 = None if isinstance(params0, torch.Tensor) else np.atleast_2d(np.asarray(params0, dtype=np.float64))
    d_p = eng.to_device(params0 if host_p0 is None else host_p0).clone()
    if d_p.dim() == 1:
        d_p = d_p.unsqueeze(0)
    batch, n_q = d_p.shape[0], d_q.shape[0]
    _apply_ties(d_p, ties)
    d_I, d_dI, _ = eng._obs(Iobs, dI, batch, n_q)
    obj = eng.objective(error_weighted, logI_weighted, q_range, has_dI=d_dI is not None)
    blo, bhi = layout.bounds(free_slots)
    tr = _BoundTransform(torch, blo, bhi, eng.device)
    fidx = torch.as_tensor(free_slots, dtype=torch.long, device=eng.device)
    max_cells = None
    if layout.n_xtal:
        hp = d_p.cpu().numpy()
        max_cells = min(90000, int(lowering.max_cells(layout, 1.5 * np.abs(hp))))
    state = dict(tables=None, n_eval=0)
    inf = torch.full((batch,), float('inf'), dtype=torch.float64, device=eng.device)

    def evaluate(u, mask=None):
        """chi2 of every system at internal coordinates u [batch, nf]; systems outside `mask` give +inf."""
        p = d_p.clone()
        p[:, fidx] = tr.to_external(u)
        _apply_ties(p, ties)
        act = None if mask is None else mask.to(torch.uint8)
        if layout.n_xtal:
            state['tables'] = eng.build_tables(layout, p, batch, max_cells=max_cells, reuse=state['tables'],
                                               check=state['tables'] is None, active=act if state['tables'] is not None else None)
        out = inf.clone()
        eng.residual(layout, obj, d_q, p, d_I, d_dI, tables=state['tables'], active=act, out=out)
        state['n_eval'] += 1
        out = torch.where(torch.isnan(out), inf, out)
        return out if mask is None else torch.where(mask, out, inf)

    u0 = tr.to_internal(d_p[:, fidx])
    U = u0.unsqueeze(1).repeat(1, nf + 1, 1)
    for k in range(nf):                                   # scipy: nonzdelt = 0.05, zdelt = 0.00025
        col = U[:, k + 1, k]
        U[:, k + 1, k] = torch.where(col != 0, 1.05 * col, torch.full_like(col, 0.00025))
    F = torch.stack([evaluate(U[:, k]) for k in range(nf + 1)], dim=1)
    chi2_init = F[:, 0].clone()
    if max_iter is None:
        max_iter = 200 * nf                               # scipy's default maxiter
    done = torch.zeros(batch, dtype=torch.bool, device=eng.device)
    n_iter = 0
    rows = torch.arange(batch, device=eng.device)
    for it in range(max_iter):
        n_iter = it + 1
        order = torch.argsort(F, dim=1)
        F = torch.gather(F, 1, order)
        U = torch.gather(U, 1, order.unsqueeze(-1).expand(-1, -1, nf))
        small_x = (U[:, 1:] - U[:, :1]).abs().amax(dim=(1, 2)) <= xatol
        small_f = (F[:, 1:] - F[:, :1]).abs().amax(dim=1) <= fatol
        done = done | (small_x & small_f)
        if it % 10 == 9 and bool(done.all().item()):
            break
        act = ~done
        xbar = U[:, :-1].mean(dim=1)
        xw = U[:, -1]
        xr = 2 * xbar - xw
        fr = evaluate(xr, act)
        c_exp = act & (fr < F[:, 0])
        c_acc = act & ~c_exp & (fr < F[:, -2])
        c_oc = act & ~c_exp & ~c_acc & (fr < F[:, -1])
        c_ic = act & ~c_exp & ~c_acc & ~c_oc
        x2 = torch.where(c_exp.unsqueeze(1), 3 * xbar - 2 * xw,
                         torch.where(c_oc.unsqueeze(1), 1.5 * xbar - 0.5 * xw, 0.5 * xbar + 0.5 * xw))
        f2 = evaluate(x2, c_exp | c_oc | c_ic)
        take2 = (c_exp & (f2 < fr)) | (c_oc & (f2 <= fr)) | (c_ic & (f2 < F[:, -1]))
        take_r = (c_exp & ~take2) | c_acc
        shrink = (c_oc | c_ic) & ~take2
        new_x = torch.where(take2.unsqueeze(1), x2, torch.where(take_r.unsqueeze(1), xr, xw))
        new_f = torch.where(take2, f2, torch.where(take_r, fr, F[:, -1]))
        U[:, -1] = new_x
        F[:, -1] = new_f
        if bool(shrink.any().item()):
            for k in range(1, nf + 1):
                xs = U[:, 0] + 0.5 * (U[:, k] - U[:, 0])
                fs = evaluate(xs, shrink)
                U[:, k] = torch.where(shrink.unsqueeze(1), xs, U[:, k])
                F[:, k] = torch.where(shrink, fs, F[:, k])
    best = torch.argmin(F, dim=1)
    ub = U[rows, best]
    p = d_p.clone()
    p[:, fidx] = tr.to_external(ub)
    _apply_ties(p, ties)
    chi2 = F[rows, best]
    res = FitResult(p, chi2_init, chi2, n_iter, torch.full((batch,), state['n_eval'], dtype=torch.int32, device=eng.device), ~done,
                    done.to(torch.uint8))
    if return_device:
        return res
    return FitResult(p.cpu().numpy(), chi2_init.cpu().numpy(), chi2.cpu().numpy(), n_iter, res.n_accept.cpu().numpy(),
                     (~done).cpu().numpy(), done.to(torch.uint8).cpu().numpy())
