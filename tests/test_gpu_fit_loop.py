"""GPU tests of the fit loop's building blocks added in round 2: prepared objective streams, population-local
Jacobian variants, the fused trial-update / accept kernels, stop reasons, and the corner cases the reference's
objective defines (empty fit mask, parameters starting at zero)."""
import ctypes as C

import numpy as np
import pytest

from helpers import cfg2_system, cfg2_params, cfg4_system, cfg4_params

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from xrsdkit_b200 import engine
    return engine.get_engine()


@pytest.mark.parametrize("log_mode,err_mode,with_dI", [(True, True, False), (False, True, True), (True, False, False), (True, True, True)])
def test_prepared_objective_equals_raw_objective(eng, log_mode, err_mode, with_dI):
    """xrsd_objective_prepare + prepared = 1 gives the residual, the normal equations and the stored-pattern chi2 of the
    raw objective (system/__init__.py:134-187): masks (I_obs <= 0, q_range), default dI = sqrt(I), given dI."""
    import torch
    from xrsdkit_b200.system import System
    from xrsdkit_b200 import batch
    s = cfg2_system(System)
    q = np.linspace(0.01, 0.5, 1100)
    lay = batch.layout_of(s, q)
    B = 7
    P = cfg2_params(lay, B, seed=91)
    rs = np.random.RandomState(5)
    Iobs = batch.compute_intensity_batch(lay, q, P) * (1 + 0.05 * rs.randn(B, len(q)))
    Iobs[:, 11] = 0.0
    Iobs[3, 300:310] = -2.0
    dI = (0.02 * np.abs(Iobs) + 1e-3) if with_dI else None
    free = [lay.slot(k) for k in ("noise__I0", "p__I0", "p__r", "p__sigma")]
    d_q, d_p, d_I = eng.to_device(q), eng.to_device(P * 1.02), eng.to_device(Iobs)
    d_dI = eng.to_device(dI) if with_dI else None
    obj = eng.objective(err_mode, log_mode, (0.04, 0.46), has_dI=with_dI)
    objp, fo, w = eng.prepare_objective(obj, d_q, d_I, d_dI, len(q), B)
    assert objp.prepared == 1 and obj.prepared == 0
    assert bool((w[:, 11] == -1).all()) and bool((w[3, 300:310] == -1).all()) and bool((w[:, 0] == -1).all())
    want = eng.residual(lay, obj, d_q, d_p, d_I, d_dI).clone()
    got = eng.residual(lay, objp, d_q, d_p, fo, w).clone()
    np.testing.assert_allclose(got.cpu().numpy(), want.cpu().numpy(), rtol=1e-13)
    a = [t.clone() for t in eng.jacobian(lay, obj, d_q, d_p, d_I, d_dI, len(q), free, None)]
    b = [t.clone() for t in eng.jacobian(lay, objp, d_q, d_p, fo, w, len(q), free, None)]
    for x, y in zip(a, b):
        np.testing.assert_allclose(y.cpu().numpy(), x.cpu().numpy(), rtol=1e-11, atol=1e-300)
    pat = eng.intensity(lay, d_q, d_p)
    np.testing.assert_allclose(eng.pattern_chi2(objp, d_q, pat, fo, w, len(q)).cpu().numpy(), want.cpu().numpy(), rtol=1e-13)
    # one shared observed pattern: the streams are [n_q] and broadcast (ld = 0)
    objs, fo1, w1 = eng.prepare_objective(obj, d_q, d_I[0].contiguous(), d_dI[0].contiguous() if with_dI else None, 0, B)
    assert tuple(fo1.shape) == (len(q),)
    np.testing.assert_allclose(eng.residual(lay, objs, d_q, d_p, fo1, w1).cpu().numpy(),
                               eng.residual(lay, obj, d_q, d_p, d_I[0].contiguous(), d_dI[0].contiguous() if with_dI else None).cpu().numpy(),
                               rtol=1e-13)


def test_variant_slabs_of_one_population_parameters(eng):
    """A finite-difference variant whose parameter belongs to one population evaluates that population alone
    (xrsd_jacobian_n_variants): its slab equals the population's own pattern at the perturbed value, and the normal
    equations agree with a Jacobian that takes whole-pattern differences (the parameter tied into a second population
    forces those)."""
    from xrsdkit_b200.system import System
    from xrsdkit_b200 import batch
    s = cfg4_system(System)
    q = np.linspace(0.01, 1.0, 1400)
    lay = batch.layout_of(s, q)
    B = 5
    P = cfg4_params(lay, B, seed=47)
    names = ("noise__I0", "x__I0", "x__a", "x__r", "gp__I0", "gp__rg")
    free = [lay.slot(k) for k in names]
    obj = eng.objective(True, True, (0.0, float("inf")))
    d_q, d_p = eng.to_device(q), eng.to_device(P)
    d_I = eng.intensity(lay, d_q, d_p) * (1 + 0.02 * eng.to_device(np.random.RandomState(2).randn(B, len(q))))
    tables = eng.build_tables(lay, d_p, B, params_host=P)
    key = ("t", "local")
    out = [t.clone() for t in eng.jacobian(lay, obj, d_q, d_p, d_I, None, len(q), free, tables, resident=key)]
    n_var = eng.jacobian_variants(lay, free)
    assert n_var == 4                                   # base, a, r, rg
    slabs = eng.jacobian_slabs(key, lay, B, len(q), n_var).clone()
    eng.release_resident(key)
    for v, name, pop in ((1, "x__a", 1), (2, "x__r", 1), (3, "gp__rg", 2)):
        P1 = P.copy()
        P1[:, lay.slot(name)] *= 1 + 1e-6
        tot, parts = eng.intensity_components(lay, d_q, eng.to_device(P1), tables=tables)      # tables held fixed
        # (the components are differences of running totals: their absolute error is that of the total)
        np.testing.assert_allclose(slabs[:, v].cpu().numpy(), parts[:, pop].cpu().numpy(), rtol=1e-12,
                                   atol=4e-16 * float(tot.max()))
    # reference values: the same forward differences taken on WHOLE patterns, tables held fixed
    Ib = eng.intensity(lay, d_q, d_p, tables=tables)
    w = d_I / d_I.sum(dim=1, keepdim=True)              # error_weighted with dI = sqrt(I): weights ~ I_obs, normalised
    r = Ib.log() - d_I.log()
    cols = []
    for name in names:
        h = 1e-6 * np.abs(P[:, lay.slot(name)])
        Pp = P.copy()
        Pp[:, lay.slot(name)] += h
        hh = eng.to_device(Pp[:, lay.slot(name)] - P[:, lay.slot(name)])
        dlog = eng.torch.log1p((eng.intensity(lay, d_q, eng.to_device(Pp), tables=tables) - Ib) / Ib)
        cols.append(dlog / hh[:, None])
    J = eng.torch.stack(cols, dim=2)                    # [B, n_q, nf]
    JTJ = eng.torch.einsum("bq,bqi,bqj->bij", w, J, J)
    JTr = eng.torch.einsum("bq,bqi,bq->bi", w, J, r)
    scale = JTJ.diagonal(dim1=1, dim2=2).sqrt()
    np.testing.assert_allclose((out[0] / (scale[:, :, None] * scale[:, None, :])).cpu().numpy(),
                               (JTJ / (scale[:, :, None] * scale[:, None, :])).cpu().numpy(), atol=1e-5)
    np.testing.assert_allclose((out[1] / scale).cpu().numpy(), (JTr / scale).cpu().numpy(), atol=1e-5 * float((JTr / scale).abs().max()))


def test_hard_sphere_variants_with_a_resident_base(eng):
    """r_hard and v_fraction act through S(q) alone: with a resident base (XRSD_JAC_BASE_READY, every LM iteration after
    the first) their variant slabs are I_pop * S'(q) / S(q) from the stored running totals instead of a fresh
    size-distribution loop.  The slabs and the normal equations must equal those of the full evaluation."""
    from xrsdkit_b200.system import System
    from xrsdkit_b200 import batch
    s = cfg2_system(System)
    for q in (np.linspace(0.0, 0.5, 1111), np.sort(np.random.RandomState(4).uniform(0.004, 0.6, 900))):
        lay = batch.layout_of(s, q)
        B = 7
        P = cfg2_params(lay, B, seed=5)
        names = ("noise__I0", "p__I0", "p__r", "p__sigma", "p__r_hard", "p__v_fraction")
        free = [lay.slot(k) for k in names]
        obj = eng.objective(True, True, (0.0, float("inf")))
        d_q, d_p = eng.to_device(q), eng.to_device(P)
        d_I = eng.intensity(lay, d_q, d_p) * (1 + 0.02 * eng.to_device(np.random.RandomState(2).randn(B, len(q))))
        key = ("t", "hs")
        full = [t.clone() for t in eng.jacobian(lay, obj, d_q, d_p, d_I, None, len(q), free, None, resident=key)]
        n_var = eng.jacobian_variants(lay, free)
        assert n_var == 5                                   # base, r, sigma, r_hard, v_fraction
        slabs_full = eng.jacobian_slabs(key, lay, B, len(q), n_var).clone()
        ready = [t.clone() for t in eng.jacobian(lay, obj, d_q, d_p, d_I, None, len(q), free, None, resident=key, base_ready=True)]
        slabs_ready = eng.jacobian_slabs(key, lay, B, len(q), n_var).clone()
        eng.release_resident(key)
        a, b = slabs_full.cpu().numpy(), slabs_ready.cpu().numpy()
        fin = np.isfinite(a[:, 3:5])
        assert np.array_equal(fin, np.isfinite(b[:, 3:5]))
        # the population's term comes out of a difference of running totals: absolute error of the total's rounding
        tol = 1e-12 * np.abs(a[:, 3:5]) + 8e-16 * np.nanmax(np.abs(a[:, 0]), axis=1)[:, None, None]
        assert np.all(np.abs(a[:, 3:5] - b[:, 3:5])[fin] <= tol[fin])
        np.testing.assert_array_equal(a[:, 1:3], b[:, 1:3])                 # r and sigma: the same evaluation in both modes
        scale = full[0].diagonal(dim1=1, dim2=2).sqrt()
        np.testing.assert_allclose((ready[0] / (scale[:, :, None] * scale[:, None, :])).cpu().numpy(),
                                   (full[0] / (scale[:, :, None] * scale[:, None, :])).cpu().numpy(), atol=1e-6)
        np.testing.assert_allclose((ready[1] / scale).cpu().numpy(), (full[1] / scale).cpu().numpy(),
                                   atol=1e-6 * float((full[1] / scale).abs().max()))


def test_empty_fit_mask_gives_zero_like_the_reference(eng):
    """No q-point inside q_range (or no positive observation): compute_chi2 sums over nothing and returns 0.0
    (tools/__init__.py:104-125); the kernels must not return 0/0."""
    from xrsdkit_b200.system import System
    from xrsdkit_b200 import batch
    from oracle import xrsd_oracle as orc
    s = cfg2_system(System)
    q = np.linspace(0.01, 0.5, 600)
    lay = batch.layout_of(s, q)
    P = cfg2_params(lay, 3, seed=5)
    Iobs = batch.compute_intensity_batch(lay, q, P)
    assert orc.residual(q, Iobs[0], Iobs[0] * 1.1, q_range=(0.6, 0.7)) == 0.0
    chi2 = batch.evaluate_residual_batch(lay, q, Iobs * 1.1, P, q_range=(0.6, 0.7))
    assert chi2.tolist() == [0.0, 0.0, 0.0]
    neg = -np.abs(Iobs)
    assert batch.evaluate_residual_batch(lay, q, neg, P).tolist() == [0.0, 0.0, 0.0]
    obj = eng.objective(True, True, (0.6, 0.7))
    d_q, d_p, d_I = eng.to_device(q), eng.to_device(P), eng.to_device(Iobs)
    assert eng.pattern_chi2(obj, d_q, d_I, d_I, None, len(q)).tolist() == [0.0, 0.0, 0.0]
    free = [lay.slot("p__I0"), lay.slot("p__r")]
    JTJ, JTr, c = eng.jacobian(lay, obj, d_q, d_p, d_I, None, len(q), free, None)
    assert c.tolist() == [0.0, 0.0, 0.0] and float(JTJ.abs().max()) == 0.0 and float(JTr.abs().max()) == 0.0
    s.set_q_range(0.6, 0.7)
    assert s.evaluate_residual(q, Iobs[0]) == 0.0 == s.evaluate_residual(q, Iobs[0], I_comp=Iobs[0] * 1.2)
    # a fit over an empty mask has nothing to improve: it stops at once, parameters untouched
    res = batch.fit_batch(lay, q, Iobs, P * 1.05, free_slots=free, q_range=(0.6, 0.7), max_iter=5)
    np.testing.assert_array_equal(res.params, P * 1.05)
    assert not res.active.any() and res.chi2.tolist() == [0.0, 0.0, 0.0]


def test_lm_moves_a_population_that_starts_at_zero_intensity(eng):
    """A population (or noise level) that starts at I0 = 0 -- its lower bound and a legal value -- has no derivative in
    I_pop / I0; the loop gives such parameters a finite-difference variant and never parks an I0 on the bound."""
    from xrsdkit_b200.system import System
    from xrsdkit_b200 import batch
    s = cfg4_system(System)
    q = np.linspace(0.01, 1.0, 2048)
    lay = batch.layout_of(s, q)
    n = 12
    truth = cfg4_params(lay, n, seed=33)
    Iobs = batch.compute_intensity_batch(lay, q, truth)
    free = [lay.slot(k) for k in ("noise__I0", "x__I0", "gp__I0", "gp__rg")]
    start = truth.copy()
    start[:, lay.slot("gp__I0")] = 0.0
    start[::2, lay.slot("noise__I0")] = 0.0
    start[:, lay.slot("gp__rg")] *= 1.03
    res = batch.fit_batch(lay, q, Iobs, start, free_slots=free, max_iter=60)
    assert np.all(res.params[:, lay.slot("gp__I0")] > 0)
    np.testing.assert_allclose(res.params[:, free], truth[:, free], rtol=2e-3)
    assert np.all(res.chi2 < 1e-6 * res.chi2_init)
    # ... and a population that should vanish shrinks towards the bound without landing on it
    gone = truth.copy()
    gone[:, lay.slot("gp__I0")] = 0.0
    Iobs0 = batch.compute_intensity_batch(lay, q, gone)
    res0 = batch.fit_batch(lay, q, Iobs0, truth, free_slots=free, max_iter=60)
    g = res0.params[:, lay.slot("gp__I0")]
    assert np.all(g > 0) and np.all(g < 1e-4 * truth[:, lay.slot("gp__I0")])      # (the trust region allows -25 % per step)


def test_stop_reasons_and_converged_flag(eng):
    """FitResult.reason: 1 = tolerance met, 2 = no acceptable step left, 0 = iteration cap; `converged` is the first
    only (the reference reports the optimiser's success flag, system/__init__.py:268-275)."""
    from xrsdkit_b200.system import System
    from xrsdkit_b200 import batch
    s = cfg2_system(System)
    q = np.linspace(0.01, 0.5, 1024)
    lay = batch.layout_of(s, q)
    n = 24
    truth = cfg2_params(lay, n, seed=12)
    rs = np.random.RandomState(3)
    Iobs = batch.compute_intensity_batch(lay, q, truth) * (1 + 0.02 * rs.randn(n, len(q)))
    free = [lay.slot(k) for k in ("noise__I0", "p__I0", "p__r", "p__sigma")]
    start = truth.copy()
    start[:, free] *= 1 + 0.05 * rs.uniform(-1, 1, (n, len(free)))
    capped = batch.fit_batch(lay, q, Iobs, start, free_slots=free, max_iter=2)
    assert capped.active.all() and not capped.converged.any() and (capped.reason == 0).all()
    done = batch.fit_batch(lay, q, Iobs, start, free_slots=free, max_iter=200)
    assert not done.active.any() and set(done.reason.tolist()) <= {1, 2}
    assert done.converged.mean() > 0.7, done.reason
    np.testing.assert_array_equal(done.converged, done.reason == 1)
    # a NaN objective from the start (an error estimate that is not a number) is not a convergence
    dI = np.sqrt(np.abs(Iobs))
    dI[0, 100] = np.nan
    r = batch.fit_batch(lay, q, Iobs, start, free_slots=free, max_iter=200, dI=dI)
    assert r.reason[0] == 3 and not r.converged[0] and r.converged[1:].any()


def test_fit_that_starts_at_the_optimum_of_exact_data_stops_at_once(eng):
    """chi2 = 0 at the start (I_obs produced by the same kernel at the start parameters): nothing to improve, no step
    is taken, the parameters come back bit for bit and the systems count as converged."""
    from xrsdkit_b200.system import System
    from xrsdkit_b200 import batch, fitting
    s = cfg2_system(System)
    q = np.linspace(0.01, 0.5, 512)
    lay = batch.layout_of(s, q)
    truth = cfg2_params(lay, 4, seed=2)
    Iobs = batch.compute_intensity_batch(lay, q, truth)
    free = [lay.slot("p__r"), lay.slot("p__sigma")]
    res = fitting.lm_fit_batch(lay, q, Iobs, truth, free, max_iter=12, eng=eng)      # starts AT the optimum of noiseless data
    assert np.all(res.n_accept == 0) and np.all(res.chi2 == 0.0) and res.converged.all() and res.n_iter <= 5
    np.testing.assert_array_equal(res.params, truth)


def test_free_atomic_coordinate_is_fitted(eng):
    """u_1 of a two-atom cubic cell, freed: the coordinate acts through the reflection tables only, so the default
    fit (method 'lm') hands such systems to the batched Nelder-Mead mode instead of holding the parameter forever
    (the finite-difference column of a table parameter is identically zero).  The coordinate must be recovered."""
    from xrsdkit_b200.system import System
    from xrsdkit_b200 import batch

    def V(x, fixed=False):
        return {"value": x, "fixed": fixed}

    s = System(x=dict(structure="crystalline", form="polyatomic",
                      settings={"lattice": "P_cubic", "space_group": "Pm-3m", "q_max": 4.5, "n_atoms": 2,
                                "symbol_0": "Cs", "symbol_1": "Cl", "use_symmetry": False},
                      parameters=dict(a=V(4.12), hwhm_g=V(0.02), hwhm_l=V(0.01), I0=V(100.0),
                                      u_1=V(0.5), v_1=V(0.5), w_1=V(0.5))),
               noise={"model": "flat", "parameters": {"I0": V(0.01)}}, source_wavelength=0.8265617)
    q = np.linspace(0.8, 4.5, 900)
    lay = batch.layout_of(s, q)
    u1, I0 = lay.slot("x__u_1"), lay.slot("x__I0")
    assert u1 in lay.coordinate_slots()
    n = 6
    truth = np.tile(lay.values, (n, 1))
    truth[:, u1] = np.linspace(0.30, 0.42, n)        # (u and 1 - u give the same pattern: stay clear of 1/2)
    Iobs = batch.compute_intensity_batch(lay, q, truth)
    start = truth.copy()
    start[:, u1] += 0.02
    start[:, I0] *= 1.1
    res = batch.fit_batch(lay, q, Iobs, start, free_slots=[u1, I0])
    assert np.all(res.chi2 < 1e-4 * res.chi2_init), res.chi2 / res.chi2_init
    np.testing.assert_allclose(res.params[:, u1], truth[:, u1], atol=2e-4)
    np.testing.assert_allclose(res.params[:, I0], truth[:, I0], rtol=2e-3)


def test_layout_refuses_another_q_grid(eng):
    """A layout carries state of the grid it was lowered for (q_step, q0_is_zero, the q[-1] fallback of q_max): the
    engine refuses a different grid instead of evaluating it with the wrong assumptions."""
    from xrsdkit_b200.system import System
    from xrsdkit_b200 import batch
    s = cfg2_system(System)
    q = np.linspace(0.01, 0.5, 512)
    lay = batch.layout_of(s, q)
    P = cfg2_params(lay, 2)
    batch.compute_intensity_batch(lay, q, P)
    batch.compute_intensity_batch(lay, eng.to_device(q), P, device_out=True)
    for other in (np.linspace(0.01, 0.5, 513), np.linspace(0.01, 0.6, 512), np.sort(np.random.RandomState(0).uniform(0.01, 0.5, 512))):
        with pytest.raises(ValueError, match="another q grid"):
            batch.compute_intensity_batch(lay, other, P)
        with pytest.raises(ValueError, match="another q grid"):
            batch.evaluate_residual_batch(lay, eng.to_device(other), np.ones((2, 512)), P)
    # a layout lowered for a non-uniform grid makes no assumption about spacing: any grid with the same flags is fine
    qn = np.sort(np.random.RandomState(1).uniform(0.01, 0.5, 300))
    layn = batch.layout_of(s, qn)
    a = batch.compute_intensity_batch(layn, q, P)
    b = batch.compute_intensity_batch(lay, q, P)
    np.testing.assert_allclose(a, b, rtol=1e-12)
    with pytest.raises(ValueError, match="another q grid"):
        batch.compute_intensity_batch(layn, np.linspace(0.0, 0.5, 100), P)          # q[0] == 0 changes q0_is_zero


def test_rescale_I0_to_measured_pattern(eng, ref_tables):
    """System.rescale_I0_to = the I0 rescaling at the end of the reference's system_from_prediction
    (models/predict.py:372-378): every I0 times sum(I) / sum(I_computed), checked against that arithmetic on the oracle."""
    from oracle import xrsd_oracle as orc
    from xrsdkit_b200.system import System
    from xrsdkit_b200 import batch
    s = cfg4_system(System)
    q = np.linspace(0.01, 1.0, 1200)
    I_meas = 37.5 * orc.system_intensity(q, s.to_dict(), **ref_tables) * (1 + 0.01 * np.random.RandomState(0).randn(len(q)))
    want = np.sum(I_meas) / np.sum(orc.system_intensity(q, s.to_dict(), **ref_tables))
    before = {k: r["value"] for k, r in s.flatten_params().items()}
    f = s.rescale_I0_to(q, I_meas)
    assert f == pytest.approx(want, rel=1e-12)
    for k, r in s.flatten_params().items():
        assert r["value"] == pytest.approx(before[k] * (want if k.endswith("__I0") else 1.0), rel=1e-12), k
    assert np.sum(s.compute_intensity(q)) == pytest.approx(np.sum(I_meas), rel=1e-11)
    lay = batch.layout_of(cfg4_system(System), q)
    P = cfg4_params(lay, 5, seed=3)
    scaled, fac = batch.rescale_I0_batch(lay, q, I_meas, P)
    Ic = batch.compute_intensity_batch(lay, q, scaled)
    np.testing.assert_allclose(Ic.sum(axis=1), np.sum(I_meas), rtol=1e-11)
    i0 = [lay.slot(k) for k in ("noise__I0", "x__I0", "gp__I0")]
    np.testing.assert_allclose(scaled[:, i0], P[:, i0] * fac[:, None], rtol=1e-14)
    rest = [j for j in range(lay.n_params) if j not in i0]
    np.testing.assert_array_equal(scaled[:, rest], P[:, rest])


def test_reductions_are_bit_identical_from_run_to_run(eng):
    """chi2 of the residual kernel and JTJ / JTr / chi2 of the Jacobian reduction: the spans (chunks) of a system leave
    partial sums that the last-arriving one adds up in span order, so 20 repeats of a small batch -- every pattern cut
    into several spans that finish in a different order each time -- give the same bits."""
    import torch
    from xrsdkit_b200.system import System
    from xrsdkit_b200 import batch
    for make, params, names in ((cfg4_system, cfg4_params, ("noise__I0", "x__I0", "x__a", "x__hwhm_g", "x__r", "gp__I0", "gp__rg")),
                                (cfg2_system, cfg2_params, ("noise__I0", "p__I0", "p__r", "p__sigma", "p__r_hard"))):
        s = make(System)
        q = np.linspace(0.01, 1.0 if make is cfg4_system else 0.5, 8192)
        lay = batch.layout_of(s, q)
        B = 3
        P = params(lay, B, seed=8)
        free = [lay.slot(k) for k in names]
        obj = eng.objective(True, True, (0.0, float("inf")))
        d_q, d_p = eng.to_device(q), eng.to_device(P * 1.01)
        d_I = eng.intensity(lay, d_q, eng.to_device(P)) * (1 + 0.02 * eng.to_device(np.random.RandomState(1).randn(B, len(q))))
        tables = eng.build_tables(lay, d_p, B, params_host=P * 1.01) if lay.n_xtal else None
        first = None
        for rep in range(20):
            chi2 = eng.residual(lay, obj, d_q, d_p, d_I, tables=tables).clone()
            JTJ, JTr, c2 = [t.clone() for t in eng.jacobian(lay, obj, d_q, d_p, d_I, None, len(q), free, tables)]
            got = [chi2, JTJ, JTr, c2]
            if first is None:
                first = got
                assert torch.isfinite(JTJ).all() and float(JTJ.abs().max()) > 0
                np.testing.assert_allclose(c2.cpu().numpy(), chi2.cpu().numpy(), rtol=1e-12)
            for a, b in zip(first, got):
                assert torch.equal(a.view(torch.int64), b.view(torch.int64)), rep
