"""Shared helpers for the parity tests."""
import numpy as np

TOL_F64 = 1e-10   # BASELINE.json: FP64 path <= 1e-10 relative to the reference's numpy result


def rel_err(got, ref, floor_frac=1e-9):
    """max |got-ref| / (|ref| + floor_frac*max|ref|) over finite reference points, and whether the
    non-finite patterns agree.  The small floor only matters at exact zeros of the reference."""
    got = np.asarray(got, dtype=float)
    ref = np.asarray(ref, dtype=float)
    fin = np.isfinite(ref)
    same = np.array_equal(np.isfinite(got), fin)
    if not fin.any():
        return 0.0, same
    scale = np.max(np.abs(ref[fin]))
    if scale == 0.0:   # identically zero reference (e.g. a population with I0 = 0)
        return (0.0 if not np.any(got[fin]) else float("inf")), same
    err = np.abs(got[fin] - ref[fin]) / (np.abs(ref[fin]) + floor_frac * scale)
    return float(err.max()), same


def cfg2_system(System):
    return System(
        p=dict(structure="disordered", form="spherical",
               settings={"distribution": "r_normal", "sampling_width": 2.5, "sampling_step": 0.1},
               parameters={"r": {"value": 20.0}, "sigma": {"value": 0.1}, "r_hard": {"value": 25.0},
                           "v_fraction": {"value": 0.3}, "I0": {"value": 10.0}}),
        noise={"model": "flat", "parameters": {"I0": {"value": 0.01}}})


def cfg2_params(layout, n, seed=1002):
    """Synthetic cfg2 parameter sets: r0~U(10,40), sigma~U(.03,.3), r_hard=r0*U(1,1.5),
    v~U(.05,.5), I0~logU(1,1e3); flat noise 0.01."""
    rs = np.random.RandomState(seed)
    P = np.tile(layout.values, (n, 1))
    r0 = rs.uniform(10, 40, n)
    P[:, layout.slot("p__r")] = r0
    P[:, layout.slot("p__sigma")] = rs.uniform(0.03, 0.3, n)
    P[:, layout.slot("p__r_hard")] = r0 * rs.uniform(1.0, 1.5, n)
    P[:, layout.slot("p__v_fraction")] = rs.uniform(0.05, 0.5, n)
    P[:, layout.slot("p__I0")] = 10 ** rs.uniform(0, 3, n)
    P[:, layout.slot("noise__I0")] = 0.01
    return P


def cfg3_system(System, mode="local"):
    return System(
        x=dict(structure="crystalline", form="spherical",
               settings={"lattice": "F_cubic", "space_group": "Fm-3m", "q_max": 1.0, "profile": "voigt",
                         "structure_factor_mode": mode},
               parameters={"a": {"value": 60.0}, "r": {"value": 20.0}, "hwhm_g": {"value": 0.002},
                           "hwhm_l": {"value": 0.002}}),
        source_wavelength=0.8265617)


def cfg3_params(layout, n, seed=1003):
    rs = np.random.RandomState(seed)
    P = np.tile(layout.values, (n, 1))
    a = rs.uniform(40, 80, n)
    P[:, layout.slot("x__a")] = a
    P[:, layout.slot("x__r")] = a * rs.uniform(0.2, 0.35, n)
    P[:, layout.slot("x__hwhm_g")] = 10 ** rs.uniform(np.log10(5e-4), np.log10(5e-3), n)
    P[:, layout.slot("x__hwhm_l")] = 10 ** rs.uniform(np.log10(5e-4), np.log10(5e-3), n)
    return P


def cfg4_system(System):
    return System(
        x=dict(structure="crystalline", form="spherical",
               settings={"lattice": "F_cubic", "space_group": "Fm-3m", "q_max": 1.0, "profile": "voigt"},
               parameters={"a": {"value": 60.0}, "r": {"value": 15.0}, "hwhm_g": {"value": 0.002},
                           "hwhm_l": {"value": 0.002}, "I0": {"value": 1.0}}),
        gp=dict(structure="diffuse", form="guinier_porod", parameters={"rg": {"value": 8.0}, "I0": {"value": 10.0}}),
        noise={"model": "flat", "parameters": {"I0": {"value": 0.01}}},
        source_wavelength=0.8265617)


def cfg4_params(layout, n, seed=1004):
    rs = np.random.RandomState(seed)
    P = cfg3_params(layout, n, seed)
    P[:, layout.slot("x__I0")] = 10 ** rs.uniform(-1, 1, n)
    P[:, layout.slot("gp__rg")] = rs.uniform(3, 15, n)
    P[:, layout.slot("gp__I0")] = 10 ** rs.uniform(0, 2, n)
    P[:, layout.slot("noise__I0")] = 10 ** rs.uniform(-3, -1, n)
    return P


def params_to_sysdict(system, layout, row):
    """System.to_dict() with the parameter values of one row of a batch matrix."""
    d = system.to_dict()
    for name, v in zip(layout.names, row):
        pop, par = name.split("__")
        d[pop]["parameters"][par]["value"] = float(v)
    return d
