"""GPU parity tests: the CUDA engine (through the C ABI) against the reference's outputs
(tests/golden, produced by the unmodified reference) and against the oracle on seeded inputs.
Tolerance: 1e-10 relative (BASELINE.json FP64 gate); reflection lists bit-exact."""
import numpy as np
import pytest

from helpers import TOL_F64, rel_err, cfg2_system, cfg2_params, cfg3_system, cfg3_params, cfg4_system, cfg4_params, params_to_sysdict

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from xrsdkit_b200 import engine
    return engine.get_engine()


def test_library_is_the_loaded_native_code(eng):
    import os
    from xrsdkit_b200 import abi
    assert os.path.exists(abi.LIB_PATH)
    with open("/proc/self/maps") as f:
        assert "libxrsd_b200.so" in f.read()


def test_golden_systems_intensity(golden_systems, eng):
    """Every golden System (reference test constructions, BASELINE configs, lattice/profile spread and
    164 fitted fixture systems): total I(q), each population's I(q) and the noise term."""
    from xrsdkit_b200.system import System
    meta, arr = golden_systems
    worst = 0.0
    for m in meta:
        q = arr[m["q"]]
        s = System(**m["system"])
        got = s.compute_intensity(q)
        err, same = rel_err(got, arr[m["key"] + "_I"])
        assert same, m["name"]
        assert err <= TOL_F64, (m["name"], err)
        worst = max(worst, err)
        for pn, pop in s.populations.items():
            gp = pop.compute_intensity(q, s.sample_metadata["source_wavelength"])
            err, same = rel_err(gp, arr[m["key"] + "_pop_" + pn])
            assert same and err <= TOL_F64, (m["name"], pn, err)
        err, same = rel_err(s.noise_model.compute_intensity(q), arr[m["key"] + "_noise"])
        assert same and err <= TOL_F64, (m["name"], "noise", err)
    print("worst relative error over %d golden systems: %.2e" % (len(meta), worst))


def test_golden_elementary_functions(golden_kernels, eng):
    from xrsdkit_b200 import scattering
    from xrsdkit_b200.tools import peak_math
    g = golden_kernels
    q = g["q"]
    for row, ref in zip(g["rnormal_args"], g["rnormal_I"]):
        err, same = rel_err(scattering.spherical_normal_intensity(q, *row), ref)
        assert same and err <= TOL_F64, (row, err)
    for row, ref in zip(g["gp_args"], g["gp_I"]):
        err, same = rel_err(scattering.guinier_porod_intensity(q, *row), ref)
        assert same and err <= TOL_F64, (row, err)
    for r, ref in zip(g["sphere_r"], g["sphere_ff"]):
        # ff^2 near the zeros of ff is compared on the scale of the pattern (floor), cf. SURVEY H2
        err, same = rel_err(scattering.spherical_ff_squared(q, r), ref ** 2, floor_frac=1e-6)
        assert same and err <= TOL_F64, (r, err)
    for row, ref in zip(g["hs_args"], g["hs_S"]):
        # mask the ill-conditioned small-qd points: the reference itself is only accurate to
        # ~2e-16*(24/qd^6-ish cancellation) there: compare where qd >= 0.3
        ok = q * 2 * row[0] >= 0.3
        err, same = rel_err(scattering.hard_sphere_sf(q, *row)[ok], ref[ok])
        assert same and err <= TOL_F64, (row, err)
    x = g["px"]
    for (hg, hl), ref in zip(g["voigt_args"], g["voigt"]):
        # the argument (x/sigma/sqrt2) is rounded differently by 1 ulp at most: on the flank of a
        # Gaussian core that is a relative change of 2 z^2 ulp, so allow 1e-10 there as everywhere
        err, same = rel_err(peak_math.voigt(x, hg, hl), ref, floor_frac=0.0)
        assert same and err <= TOL_F64, (hg, hl, err)
    for h, ref in zip(g["hwhm"], g["gauss"]):
        got = peak_math.gaussian(x, h)
        big = ref > 1e-280      # denormal tail: relative comparison is meaningless there
        err, _ = rel_err(got[big], ref[big], floor_frac=0.0)
        assert err <= TOL_F64 and np.all(got[~big] <= 1e-279), (h, err)
    for h, ref in zip(g["hwhm"], g["lorentz"]):
        err, same = rel_err(peak_math.lorentzian(x, h), ref, floor_frac=0.0)
        assert same and err <= TOL_F64, (h, err)


def test_faddeeva_against_scipy(eng):
    """Re w(x+iy) over the whole argument range the Voigt profile can reach."""
    from scipy.special import wofz
    from xrsdkit_b200.tools import peak_math
    x = np.concatenate([np.linspace(0, 9, 3001), np.logspace(0.95, 7, 1500)])
    for y in (1e-10, 1e-8, 1e-6, 1e-4, 1e-2, 0.1, 0.83, 1.0, 3.0, 7.0, 7.99, 8.0, 20.0, 1e3, 1e6):
        # voigt(x', hg, hl) = Re w((x'+i hl)/(sigma sqrt2))/(sigma sqrt(2 pi)), sigma = hg/sqrt(2 ln2)
        hg = 1.0
        sigma = hg / np.sqrt(2 * np.log(2))
        got = peak_math.voigt(x * sigma * np.sqrt(2), hg, y * sigma * np.sqrt(2)) * sigma * np.sqrt(2 * np.pi)
        ref = np.real(wofz(x + 1j * y))
        err = np.max(np.abs(got - ref) / np.abs(ref))
        assert err <= 2e-12, (y, err)


def test_reflection_tables_bit_exact(golden_tables, golden_definitions, eng):
    """hkl lists and multiplicities of the reference (greedy symmetrisation) reproduced exactly."""
    from xrsdkit_b200.system import System
    from xrsdkit_b200 import lowering
    for case in golden_tables:
        pars = dict((k, {"value": v}) for k, v in case["latparams"].items())
        pars.update(r={"value": 5.0})
        s = System(x=dict(structure="crystalline", form="spherical",
                          settings={"lattice": case["lattice"], "space_group": case["space_group"],
                                    "q_min": case["q_min"], "q_max": case["q_max"], "use_symmetry": case["use_symmetry"]},
                          parameters=pars), source_wavelength=0.8265617)
        lay = lowering.lower_system(s, np.array([0.1, 1.0]))
        d_p = eng.to_device(lay.values[None, :])
        if len(case["hkl"]) > 4096:
            continue
        T = eng.build_tables(lay, d_p, 1, params_host=lay.values[None, :], export_reflections=True)
        n = int(T.n_refl[0].item())
        got = T.refl_hkl[0, :n].cpu().numpy()
        assert int(T.n_all[0].item()) == case["n_all"], (case["lattice"], case["space_group"])
        assert n == len(case["hkl"]), (case["lattice"], case["space_group"], n, len(case["hkl"]))
        assert got[:, :3].tolist() == case["hkl"], (case["lattice"], case["space_group"])
        assert got[:, 3].tolist() == case["mult"], (case["lattice"], case["space_group"])
        # shells conserve the total multiplicity unless extinct shells were pruned
        ns = int(T.n_shell[0].item())
        assert 0 < ns <= n


def test_reflection_tables_random_lattice_parameters_vs_oracle(golden_definitions, eng):
    """Seeded random cells per lattice class against the oracle's O(N^2) restatement."""
    from oracle import xrsd_oracle as orc
    from xrsdkit_b200.system import System
    from xrsdkit_b200 import lowering
    rs = np.random.RandomState(5)
    pgs = golden_definitions["sg_point_groups"]
    cases = [("F_cubic", "Fm-3m"), ("I_cubic", "Im-3m"), ("P_cubic", "Pm-3m"), ("hcp", "P6(3)/mmc"),
             ("hexagonal", "P6/mmm"), ("triclinic", "P-1"), ("P_orthorhombic", "Pmmm"), ("rhombohedral", "R-3m")]
    for lat, sg in cases:
        for _ in range(3):
            lp = dict(a=rs.uniform(20, 45), b=rs.uniform(20, 45), c=rs.uniform(20, 45),
                      alpha=rs.uniform(70, 110), beta=rs.uniform(70, 110), gamma=rs.uniform(70, 110))
            lp = dict((k, lp[k]) for k in orc.lattice_param_names(lat))
            q_max = rs.uniform(0.5, 0.9)
            hkl, mult, _ = orc.reflection_table(lat, lp, 0.0, q_max, pgs[sg])
            pars = dict((k, {"value": v}) for k, v in lp.items())
            s = System(x=dict(structure="crystalline", form="spherical",
                              settings={"lattice": lat, "space_group": sg, "q_max": q_max}, parameters=pars),
                       source_wavelength=0.8265617)
            lay = lowering.lower_system(s, np.array([0.1, 1.0]))
            T = eng.build_tables(lay, eng.to_device(lay.values[None, :]), 1, params_host=lay.values[None, :],
                                 export_reflections=True)
            n = int(T.n_refl[0].item())
            got = T.refl_hkl[0, :n].cpu().numpy()
            assert got[:, :3].tolist() == hkl.astype(int).tolist(), (lat, lp)
            assert got[:, 3].tolist() == np.asarray(mult).astype(int).tolist(), (lat, lp)


def test_cfg2_batch_vs_oracle(ref_tables, eng):
    """BASELINE cfg2 (polydisperse hard spheres), 48 seeded parameter sets on the full 4096-pt grid."""
    from oracle import xrsd_oracle as orc
    from xrsdkit_b200.system import System
    from xrsdkit_b200 import batch
    s = cfg2_system(System)
    q = np.linspace(0.01, 0.5, 4096)
    lay = batch.layout_of(s, q)
    P = cfg2_params(lay, 48)
    I = batch.compute_intensity_batch(lay, q, P)
    assert I.shape == (48, 4096)
    for b in range(48):
        ref = orc.system_intensity(q, params_to_sysdict(s, lay, P[b]), **ref_tables)
        err, same = rel_err(I[b], ref)
        assert same and err <= TOL_F64, (b, err, P[b])


def test_cfg3_cfg4_batch_vs_oracle(ref_tables, eng):
    """BASELINE cfg3/cfg4 (FCC Fm-3m voigt, + Guinier-Porod + noise): seeded lattice parameters, each
    with its own reflection table, local and radial structure-factor modes."""
    from oracle import xrsd_oracle as orc
    from xrsdkit_b200.system import System
    from xrsdkit_b200 import batch
    q = np.linspace(0.01, 1.0, 2048)
    for make, pars, n in ((lambda S: cfg3_system(S, "local"), cfg3_params, 6),
                          (lambda S: cfg3_system(S, "radial"), cfg3_params, 3), (cfg4_system, cfg4_params, 4)):
        s = make(System)
        lay = batch.layout_of(s, q)
        P = pars(lay, n)
        I = batch.compute_intensity_batch(lay, q, P)
        for b in range(n):
            ref = orc.system_intensity(q, params_to_sysdict(s, lay, P[b]), **ref_tables)
            err, same = rel_err(I[b], ref)
            assert same and err <= TOL_F64, (b, err, P[b])


def test_batch_properties_full_size(eng):
    """Size-independent properties at BASELINE sizes: batched == one-by-one, idempotence (bit-exact),
    row-permutation equivariance (bit-exact), linearity in I0."""
    from xrsdkit_b200.system import System
    from xrsdkit_b200 import batch
    s = cfg2_system(System)
    q = np.linspace(0.01, 0.5, 4096)
    lay = batch.layout_of(s, q)
    P = cfg2_params(lay, 10000)
    I = batch.compute_intensity_batch(lay, q, P, device_out=True)
    I2 = batch.compute_intensity_batch(lay, q, P, device_out=True)
    assert bool((I == I2).all())                                     # idempotent, deterministic
    assert bool(eng.torch.isfinite(I).all())
    perm = np.random.RandomState(0).permutation(10000)
    Ip = batch.compute_intensity_batch(lay, q, P[perm], device_out=True)
    assert bool((Ip == I[eng.torch.as_tensor(perm, device=I.device)]).all())
    for b in (0, 777, 9999):
        # a single-system launch uses shorter q spans, so its size-quadrature phases are not carried from span
        # to span by rotation: equal to rounding, not bit for bit
        one = batch.compute_intensity_batch(lay, q, P[b:b + 1])
        np.testing.assert_allclose(one[0], I[b].cpu().numpy(), rtol=1e-12, atol=0)
    # linearity: doubling the population I0 doubles (I - noise)
    P2 = P[:256].copy()
    P2[:, lay.slot("p__I0")] *= 2
    a = batch.compute_intensity_batch(lay, q, P[:256]) - 0.01
    c = batch.compute_intensity_batch(lay, q, P2) - 0.01
    np.testing.assert_allclose(c, 2 * a, rtol=1e-12, atol=1e-13)


def test_ragged_and_edge_inputs(eng, ref_tables):
    from oracle import xrsd_oracle as orc
    from xrsdkit_b200.system import System
    s = cfg2_system(System)
    assert s.compute_intensity(np.zeros(0)).shape == (0,)
    for n in (1, 2, 255, 257, 513, 1025):                             # spans that do not fill a CTA tile
        q = np.linspace(0.02, 0.4, n) if n > 1 else np.array([0.1])
        ref = orc.system_intensity(q, s.to_dict(), **ref_tables)
        err, same = rel_err(s.compute_intensity(q), ref)
        assert same and err <= TOL_F64, (n, err)
    # q[0] == 0 is special-cased by the reference; a zero elsewhere propagates NaN
    q = np.array([0.0, 0.1, 0.0, 0.2])
    with np.errstate(all="ignore"):
        ref = orc.system_intensity(q, s.to_dict(), **ref_tables)
    got = s.compute_intensity(q)
    assert np.isnan(ref[2]) and np.isnan(got[2])
    err, same = rel_err(got, ref)
    assert same and err <= TOL_F64
    # empty reflection table -> zeros (scattering/__init__.py:277-278)
    x = System(x=dict(structure="crystalline", form="spherical", settings={"lattice": "P_cubic", "q_max": 0.05},
                      parameters={"a": {"value": 10.0}}), source_wavelength=1.0)
    np.testing.assert_array_equal(x.compute_intensity(np.linspace(0.01, 0.05, 50)), np.full(50, 1e-3))
    # narrow Gaussian: I(0) normalisation underflows -> non-finite everywhere, as the reference
    gsys = System(x=dict(structure="crystalline", form="spherical",
                         settings={"lattice": "F_cubic", "space_group": "Fm-3m", "q_max": 0.6, "profile": "gaussian"},
                         parameters={"a": {"value": 40.0}, "hwhm": {"value": 0.002}}), source_wavelength=0.8265617)
    qg = np.linspace(0.05, 0.6, 300)
    with np.errstate(all="ignore"):
        refg = orc.system_intensity(qg, gsys.to_dict(), **ref_tables)
    gotg = gsys.compute_intensity(qg)
    assert not np.isfinite(refg).any() and not np.isfinite(gotg).any()


def test_error_behaviour(eng):
    from xrsdkit_b200.system import System, Population
    with pytest.raises(ValueError):
        System(x=dict(structure="crystalline", form="spherical", settings={"lattice": "F_cubic"},
                      parameters={"a": {"value": 40.0}}), source_wavelength=0).compute_intensity(np.linspace(0.1, 1, 10))
    with pytest.raises(ValueError):
        System(x=dict(structure="crystalline", form="spherical", settings={"lattice": "F_cubic", "space_group": "Pm-3m"},
                      parameters={"a": {"value": 40.0}}), source_wavelength=1.0).compute_intensity(np.linspace(0.1, 1, 10))
    with pytest.raises(ValueError):
        Population("diffuse", "spherical", {}, {"nope": {"value": 1.0}})
    with pytest.raises(KeyError):
        System(p=dict(structure="diffuse", form="atomic", settings={"symbol": "Xx"})).compute_intensity(np.linspace(0.1, 1, 10))


def test_host_buffer_c_abi_entry_point(eng, golden_systems):
    """xrsd_intensity_batch_host: the call a non-Python caller binds (HOST pointers in and out)."""
    import ctypes as C
    from xrsdkit_b200 import abi, batch
    from xrsdkit_b200.system import System
    lib = abi.load()
    q = np.linspace(0.01, 1.0, 1111)
    for make, pars, n in ((cfg2_system, cfg2_params, 37), (cfg4_system, cfg4_params, 9)):
        s = make(System)
        lay = batch.layout_of(s, q)
        P = np.ascontiguousarray(pars(lay, n))
        out = np.full((n, q.size + 3), -7.0)                     # row stride larger than n_q
        rc = lib.xrsd_intensity_batch_host(C.byref(lay.c), q.ctypes.data_as(C.c_void_p), q.size,
                                           P.ctypes.data_as(C.c_void_p), P.shape[1], n,
                                           out.ctypes.data_as(C.c_void_p), out.shape[1])
        assert rc == 0, lib.xrsd_last_error()
        ref = batch.compute_intensity_batch(lay, q, P)
        np.testing.assert_array_equal(out[:, :q.size], ref)
        assert np.all(out[:, q.size:] == -7.0)
    lib.xrsd_release_workspace()


def test_host_buffer_entry_point_from_several_threads(eng):
    """The host-buffer call shares one process-wide workspace: concurrent callers (ctypes drops the GIL) serialise
    on its mutex and every caller gets its own result -- small calls (mapped pinned memory, no copies), larger
    ones (staged copies) and crystalline layouts mixed."""
    import ctypes as C
    import threading
    from xrsdkit_b200 import abi, batch
    from xrsdkit_b200.system import System
    lib = abi.load()
    jobs = []
    for k, (make, pars, n, n_q) in enumerate(((cfg2_system, cfg2_params, 3, 900), (cfg2_system, cfg2_params, 200, 2048),
                                               (cfg4_system, cfg4_params, 4, 1500), (cfg2_system, cfg2_params, 1, 4096))):
        s = make(System)
        q = np.linspace(0.01, 0.5 + 0.1 * k, n_q)
        lay = batch.layout_of(s, q)
        P = np.ascontiguousarray(pars(lay, n, seed=80 + k))
        jobs.append((lay, q, P, batch.compute_intensity_batch(lay, q, P)))
    errors = []

    def work(lay, q, P, want):
        try:
            for _ in range(25):
                out = np.empty((len(P), q.size))
                rc = lib.xrsd_intensity_batch_host(C.byref(lay.c), q.ctypes.data_as(C.c_void_p), q.size,
                                                   P.ctypes.data_as(C.c_void_p), P.shape[1], len(P),
                                                   out.ctypes.data_as(C.c_void_p), q.size)
                assert rc == 0, lib.xrsd_last_error()
                np.testing.assert_array_equal(out, want)
        except Exception as e:      # noqa: BLE001 -- reported in the main thread
            errors.append(repr(e))

    threads = [threading.Thread(target=work, args=j) for j in jobs for _ in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:3]
    lib.xrsd_release_workspace()


def test_mixed_layout_collection_in_batches(eng, golden_systems):
    """All fixture systems evaluated through the layout-grouping loader: one launch per distinct layout."""
    from xrsdkit_b200.system import System
    from xrsdkit_b200.tools import ymltools
    meta, arr = golden_systems
    picks = [m for m in meta if m["q"] == "q_fixture"]
    systems = [System(**m["system"]) for m in picks]
    q = arr["q_fixture"]
    n0 = eng.launches
    I = ymltools.compute_intensity_many(systems, q)
    n_groups = len(ymltools.group_by_layout(systems, q))
    assert n_groups < len(systems) / 3
    for row, m in zip(I, picks):
        err, same = rel_err(row, arr[m["key"] + "_I"])
        assert same and err <= TOL_F64, (m["name"], err)
    assert eng.launches - n0 <= 2 * n_groups + 2


def test_far_field_interpolation_over_extreme_widths(ref_tables, eng):
    """The crystalline sum interpolates the far shells.  Peak widths from 1e-8 to 5e-2,
    Voigt and Lorentzian, sorted and unsorted q (unsorted disables the interpolation): all against the oracle."""
    from oracle import xrsd_oracle as orc
    from xrsdkit_b200.system import System
    rs = np.random.RandomState(77)
    q = np.linspace(0.02, 0.9, 1536)
    q_shuffled = q[rs.permutation(q.size)]
    for trial in range(14):
        a = rs.uniform(25, 55)
        prof = "voigt" if trial % 3 else "lorentzian"
        pars = {"a": {"value": a}, "r": {"value": a * rs.uniform(0.2, 0.35)}, "I0": {"value": 3.0}}
        if prof == "voigt":
            pars["hwhm_g"] = {"value": 10 ** rs.uniform(-8, np.log10(5e-2))}
            pars["hwhm_l"] = {"value": 10 ** rs.uniform(-8, np.log10(5e-2))}
        else:
            pars["hwhm"] = {"value": 10 ** rs.uniform(-6, np.log10(5e-2))}
        s = System(x=dict(structure="crystalline", form="spherical",
                          settings={"lattice": "F_cubic", "space_group": "Fm-3m", "q_max": 0.9, "profile": prof,
                                    "structure_factor_mode": "radial" if trial % 2 else "local"},
                          parameters=pars), source_wavelength=0.8265617)
        ref = orc.system_intensity(q, s.to_dict(), **ref_tables)
        got = s.compute_intensity(q)
        err, same = rel_err(got, ref, floor_frac=0.0)
        assert same and err <= TOL_F64, (trial, prof, pars, err)
        # the same points in random order take the direct path; the two paths must agree with each other
        got_u = s.compute_intensity(q_shuffled)
        ref_u = orc.system_intensity(q_shuffled, s.to_dict(), **ref_tables)
        err_u, same_u = rel_err(got_u, ref_u, floor_frac=0.0)
        assert same_u and err_u <= TOL_F64, (trial, prof, pars, err_u)


def test_voigt_with_zero_lorentzian_width(ref_tables, eng):
    """hwhm_l == 0: the engine takes the Gaussian closed form for such a Voigt profile (csrc/system_eval.cuh), the
    reference calls wofz on the real axis (tools/peak_math.py:38-47).  Interpolated path (sorted q), direct path
    (shuffled q) and the stand-alone profile call.  With hwhm_g = 0.02 the I(0) normalisation of the reference is
    carried by the extinct reflections (weights ~1e-31 of the allowed ones, Gaussian tails of the allowed ones underflow):
    those shells are kept for such a system (lowering.py, prune_extinct)."""
    from oracle import xrsd_oracle as orc
    from xrsdkit_b200.system import System
    from xrsdkit_b200.tools import peak_math
    rs = np.random.RandomState(5)
    q = np.linspace(0.02, 0.9, 1536)
    q_shuffled = q[rs.permutation(q.size)]
    for hl in (0.0, 1e-9):          # 1e-9 is the reference's lower bound for hwhm_l (definitions.py:292)
        for hg in (2e-2, 5e-2):
            s = System(x=dict(structure="crystalline", form="spherical",
                              settings={"lattice": "F_cubic", "space_group": "Fm-3m", "q_max": 0.9, "profile": "voigt"},
                              parameters={"a": {"value": 41.0}, "r": {"value": 11.0}, "I0": {"value": 2.0},
                                          "hwhm_g": {"value": hg}, "hwhm_l": {"value": hl}}),
                       noise={"model": "flat", "parameters": {"I0": {"value": 1e-3}}}, source_wavelength=0.8265617)
            for grid in (q, q_shuffled):
                ref = orc.system_intensity(grid, s.to_dict(), **ref_tables)
                err, same = rel_err(s.compute_intensity(grid), ref, floor_frac=0.0)
                assert same and err <= TOL_F64, (hl, hg, err)
    x = np.linspace(-0.2, 0.2, 4001)
    from scipy.special import wofz
    sigma = 4e-3 / np.sqrt(2 * np.log(2))
    want = wofz(x / (sigma * np.sqrt(2))).real / (sigma * np.sqrt(2 * np.pi))
    got = peak_math.voigt(x, 4e-3, 0.0)
    ok = want > 1e-300
    assert np.max(np.abs(got[ok] - want[ok]) / want[ok]) <= TOL_F64


def test_single_precision_path(eng):
    """The optional FP32 path of BASELINE.json's north_star (<= 1e-5 relative): there is no reference for it, so it
    is held against the FP64 engine -- on the cfg2 parameter distribution (polydisperse spheres with hard-sphere
    S(q): no zeros, plain relative error), on quadratures of other widths / steps, on a layout that mixes several
    non-crystalline populations, and on a non-uniform q grid.  Layouts with crystalline populations are refused."""
    from xrsdkit_b200 import batch
    from xrsdkit_b200.system import System
    s = cfg2_system(System)
    q = np.linspace(0.01, 0.5, 4096)
    lay = batch.layout_of(s, q)
    P = cfg2_params(lay, 512, seed=31)
    I64 = batch.compute_intensity_batch(lay, q, P)
    I32 = batch.compute_intensity_batch(lay, q, P, precision="fp32")
    rel = np.abs(I32 - I64) / np.abs(I64)
    assert np.isfinite(I32).all() and rel.max() <= 1e-5, rel.max()
    assert rel.max() > 0.0                                   # it really is another arithmetic
    # other quadratures, q down to where every radius takes the small-argument branch, and a non-uniform grid
    rs = np.random.RandomState(9)
    for width, step, grid in ((2.5, 0.02, np.linspace(1e-3, 0.6, 1500)), (4.0, 0.1, np.sort(rs.uniform(0.005, 0.8, 1200))),
                              (1.0, 0.05, np.linspace(0.2, 2.0, 777))):
        t = System(a=dict(structure="diffuse", form="spherical",
                          settings={"distribution": "r_normal", "sampling_width": width, "sampling_step": step},
                          parameters={"r": {"value": 35.0}, "sigma": {"value": 0.12}, "I0": {"value": 3.0}}),
                   b=dict(structure="diffuse", form="guinier_porod", parameters={"rg": {"value": 6.0}, "I0": {"value": 0.5}}),
                   c=dict(structure="disordered", form="spherical",
                          settings={"distribution": "r_normal", "sampling_width": 2.5, "sampling_step": 0.1},
                          parameters={"r": {"value": 12.0}, "sigma": {"value": 0.2}, "r_hard": {"value": 14.0},
                                      "v_fraction": {"value": 0.2}, "I0": {"value": 1.0}}),
                   noise={"model": "flat", "parameters": {"I0": {"value": 1e-4}}})
        lt = batch.layout_of(t, grid)
        a64 = batch.compute_intensity_batch(lt, grid, lt.values[None, :])
        a32 = batch.compute_intensity_batch(lt, grid, lt.values[None, :], precision="fp32")
        assert np.max(np.abs(a32 - a64) / np.abs(a64)) <= 1e-5, (width, step)
    x = cfg3_system(System)
    qx = np.linspace(0.05, 1.0, 600)
    with pytest.raises(ValueError):
        batch.compute_intensity_batch(batch.layout_of(x, qx), qx, batch.layout_of(x, qx).values[None, :], precision="fp32")


def test_components_in_one_launch(golden_systems, eng):
    """Per-population decomposition (noise, each population, total) from one launch against the reference's
    separate per-population evaluations."""
    from xrsdkit_b200.system import System
    meta, arr = golden_systems
    for m in meta[:54:3] + meta[60::25]:
        q = arr[m["q"]]
        s = System(**m["system"])
        n0 = eng.launches
        parts = s.compute_intensity_components(q)
        assert eng.launches - n0 <= 3          # (table key + table build) + one evaluation launch
        err, same = rel_err(parts["total"], arr[m["key"] + "_I"])
        assert same and err <= TOL_F64, m["name"]
        scale = np.nanmax(np.abs(arr[m["key"] + "_I"]))
        np.testing.assert_allclose(parts["noise"], arr[m["key"] + "_noise"], rtol=1e-10, atol=1e-13 * scale)
        for pn in s.populations:
            ref = arr[m["key"] + "_pop_" + pn]
            fin = np.isfinite(ref)
            # a difference of running totals: absolute accuracy is that of the total
            np.testing.assert_allclose(parts[pn][fin], ref[fin], rtol=1e-10, atol=1e-12 * scale, err_msg=m["name"] + ":" + pn)


def test_crystal_batch_properties_full_size(eng):
    """cfg3 / cfg4 at BASELINE sizes (8192 q, 2000 systems, each with its own reflection table):
    determinism, row-permutation equivariance, linearity in the crystal's I0, and I(q) > 0."""
    from xrsdkit_b200.system import System
    from xrsdkit_b200 import batch
    q = np.linspace(0.01, 1.0, 8192)
    s = cfg4_system(System)
    lay = batch.layout_of(s, q)
    P = cfg4_params(lay, 2000)
    I = batch.compute_intensity_batch(lay, q, P, device_out=True)
    assert bool((I == batch.compute_intensity_batch(lay, q, P, device_out=True)).all())
    assert bool(eng.torch.isfinite(I).all()) and bool((I > 0).all())
    perm = np.random.RandomState(1).permutation(2000)
    Ip = batch.compute_intensity_batch(lay, q, P[perm], device_out=True)
    assert bool((Ip == I[eng.torch.as_tensor(perm, device=I.device)]).all())
    # population-wise linearity through the one-launch decomposition: total = noise + crystal + gp
    tot, parts = eng.intensity_components(lay, q, P[:64])
    np.testing.assert_allclose(parts.sum(dim=1).cpu().numpy(), tot.cpu().numpy(), rtol=1e-13)
    P2 = P[:64].copy()
    P2[:, lay.slot("x__I0")] *= 3.0
    _, parts2 = eng.intensity_components(lay, q, P2)
    np.testing.assert_allclose(parts2[:, 1].cpu().numpy(), 3.0 * parts[:, 1].cpu().numpy(), rtol=1e-12, atol=1e-300)
    # parts are differences of running totals: the untouched population agrees to the rounding of the totals
    np.testing.assert_allclose(parts2[:, 2].cpu().numpy(), parts[:, 2].cpu().numpy(), rtol=0,
                               atol=1e-13 * float(tot.max()) * 3)


def test_all_753_fixture_systems(eng):
    """The reference's whole corpus of fitted systems (tests/test_data/dataset_1 and _2, 753 files: diffuse /
    disordered / crystalline mixes, both noise models, sigma from 8e-10 to 1.95, hwhm from 1e-9): I(q) one by
    one through System.compute_intensity, and once more through the layout-grouping batch loader."""
    import gzip
    import json
    import os
    from conftest import GOLDEN, _unjson
    from xrsdkit_b200.system import System
    from xrsdkit_b200.tools import ymltools
    with gzip.open(os.path.join(GOLDEN, "fixtures_all.json.gz"), "rt") as f:
        meta = _unjson(json.load(f))
    arr = np.load(os.path.join(GOLDEN, "fixtures_all.npz"))
    q = arr["q"]
    assert len(meta) == 753
    systems, worst = [], 0.0
    for i, m in enumerate(meta):
        s = System(**m["system"])
        systems.append(s)
        err, same = rel_err(s.compute_intensity(q), arr["I%04d" % i])
        assert same and err <= TOL_F64, (m["name"], err)
        worst = max(worst, err)
    I = ymltools.compute_intensity_many(systems, q)
    for i in range(len(meta)):
        err, same = rel_err(I[i], arr["I%04d" % i])
        assert same and err <= TOL_F64, (meta[i]["name"], err)
    print("753 fixture systems: worst relative error %.2e" % worst)


def test_large_cell_without_usable_symmetry(ref_tables, eng):
    """> 4096 reflections in the reduced list (orthorhombic cell, point group mmm: the reference defines no
    operations for it): the table builder moves its sort / shell arrays to global memory; thousands of shells
    take the chunked direct sum.  Against the oracle, intensity and reflection count."""
    from oracle import xrsd_oracle as orc
    from xrsdkit_b200.system import System
    from xrsdkit_b200 import lowering
    s = System(x=dict(structure="crystalline", form="spherical",
                      settings={"lattice": "P_orthorhombic", "space_group": "Pmmm", "q_max": 0.9, "profile": "voigt"},
                      parameters={"a": {"value": 60.0}, "b": {"value": 70.0}, "c": {"value": 80.0}, "r": {"value": 12.0},
                                  "hwhm_g": {"value": 0.004}, "hwhm_l": {"value": 0.003}}),
               source_wavelength=0.8265617)
    q = np.linspace(0.05, 0.9, 500)
    lay = lowering.lower_system(s, q)
    T = eng.build_tables(lay, eng.to_device(lay.values[None, :]), 1, params_host=lay.values[None, :])
    hkl, mult, _ = orc.reflection_table("P_orthorhombic", dict(a=60.0, b=70.0, c=80.0), 0.0, 0.9, "mmm")
    assert hkl.shape[0] > 4096
    assert int(T.n_refl[0].item()) == hkl.shape[0] == int(T.n_all[0].item())
    assert int(T.n_shell[0].item()) > 400          # 8 sign combinations share each |G|
    ref = orc.system_intensity(q, s.to_dict(), **ref_tables)
    err, same = rel_err(s.compute_intensity(q), ref)
    assert same and err <= TOL_F64, err


def test_reflection_table_long_candidate_list_with_symmetry(golden_definitions, eng):
    """F_cubic a = 80, q_max = 1: 8870 candidates, more than the alive-list capacity, so the greedy reduction walks
    the candidate box itself (the path cfg3-sized cells never take); reduced list and multiplicities bit-exact
    against the oracle's O(N^2) restatement."""
    from oracle import xrsd_oracle as orc
    from xrsdkit_b200.system import System
    from xrsdkit_b200 import lowering
    pgs = golden_definitions["sg_point_groups"]
    hkl, mult, _ = orc.reflection_table("F_cubic", dict(a=80.0), 0.0, 1.0, pgs["Fm-3m"])
    s = System(x=dict(structure="crystalline", form="spherical",
                      settings={"lattice": "F_cubic", "space_group": "Fm-3m", "q_max": 1.0},
                      parameters={"a": {"value": 80.0}, "r": {"value": 20.0}}), source_wavelength=0.8265617)
    lay = lowering.lower_system(s, np.array([0.1, 1.0]))
    T = eng.build_tables(lay, eng.to_device(lay.values[None, :]), 1, params_host=lay.values[None, :], export_reflections=True)
    assert int(T.n_all[0].item()) == int(np.sum(mult)) > 5 * 1024
    n = int(T.n_refl[0].item())
    got = T.refl_hkl[0, :n].cpu().numpy()
    assert got[:, :3].tolist() == hkl.astype(int).tolist()
    assert got[:, 3].tolist() == np.asarray(mult).astype(int).tolist()


def test_pinned_host_pipeline_matches_the_device_path(eng):
    """batch.compute_intensity_pinned (slices alternating over two streams, tables built without per-slice status
    read-back) against compute_intensity_batch: a sphere layout, a crystal layout, and a crystal layout whose
    tables outgrow the default capacities (the flags OR-ed on the device send the batch through the checked path)."""
    import torch
    from helpers import cfg2_system, cfg2_params, cfg4_system, cfg4_params
    from xrsdkit_b200.system import System
    from xrsdkit_b200 import batch
    cases = []
    s = cfg2_system(System)
    q = np.linspace(0.01, 0.5, 777)
    lay = batch.layout_of(s, q)
    cases.append((lay, q, cfg2_params(lay, 1500)))
    s = cfg4_system(System)
    q = np.linspace(0.01, 1.0, 1100)
    lay = batch.layout_of(s, q)
    cases.append((lay, q, cfg4_params(lay, 1300)))
    big = System(x=dict(structure="crystalline", form="spherical",
                        settings={"lattice": "P_orthorhombic", "space_group": "Pmmm", "q_max": 0.9},
                        parameters={"a": {"value": 60.0}, "b": {"value": 70.0}, "c": {"value": 80.0}, "r": {"value": 20.0}}),
                 source_wavelength=0.8265617)
    q = np.linspace(0.05, 0.9, 600)
    lay = batch.layout_of(big, q)
    P = np.repeat(lay.values[None, :], 3, axis=0)
    P[:, lay.slot("x__a")] = [60.0, 61.0, 59.0]
    cases.append((lay, q, P))
    for lay, q, P in cases:
        want = batch.compute_intensity_batch(lay, q, P)
        p_host = torch.as_tensor(P).pin_memory()
        q_host = torch.as_tensor(q).pin_memory()
        out = torch.empty((len(P), len(q)), dtype=torch.float64).pin_memory()
        for chunks in (None, 3):
            out.zero_()
            batch.compute_intensity_pinned(lay, q_host, p_host, out, chunks=chunks, eng=eng)
            np.testing.assert_allclose(out.numpy(), want, rtol=1e-12, atol=0)


def test_whole_pattern_ctas_equal_split_spans_at_full_size(ref_tables, eng):
    """Direct mode takes a whole 8192-point pattern per CTA once the batch fills the GPU and 512-point spans for a
    handful of systems: both must give the same bits (the span length only changes who computes what), and the
    full-size result must still match the oracle."""
    from oracle import xrsd_oracle as orc
    from helpers import cfg4_system, cfg4_params, params_to_sysdict
    from xrsdkit_b200.system import System
    from xrsdkit_b200 import batch
    s = cfg4_system(System)
    q = np.linspace(0.01, 1.0, 8192)
    lay = batch.layout_of(s, q)
    P = cfg4_params(lay, 5000, seed=77)
    I = batch.compute_intensity_batch(lay, q, P)
    rows = [0, 17, 999, 4999]
    assert np.array_equal(I[rows], batch.compute_intensity_batch(lay, q, P[rows]))
    for r in rows[:2]:
        err, same = rel_err(I[r], orc.system_intensity(q, params_to_sysdict(s, lay, P[r]), **ref_tables))
        assert same and err <= TOL_F64, (r, err)
