"""BENCH INFRASTRUCTURE ONLY -- the reference's CPU path on the host cores, for bench.py's `--impl reference`
arm and `cpu_baseline` leg.  Imports nothing from the product package (xrsdkit_b200) and never loads its .so.

Two back ends, chosen at run time:
  kind "reference": the UNMODIFIED reference package found under oracle/_ref/ (a copy of /root/reference/xrsdkit
                    made by __graft_entry__.build(); git-ignored, travels to the GPU box with the snapshot),
                    imported through oracle/ref_shim.py and driven through its own public API --
                    System(**dict), System.compute_intensity(q), System.evaluate_residual(q, I)
                    (system/__init__.py:112-187).
  kind "port":      oracle/xrsd_oracle.py (the pinned numpy restatement) when no copy of the reference is present.

Workloads are the BASELINE.json configs as plain data: the fully populated System dicts the reference itself
produces (tests/golden/workloads.json, written by oracle/make_workloads.py) and the seeded parameter generators
of tests/helpers.py (numpy only), so both bench arms evaluate the same parameter sets.
"""
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_COPY = os.path.join(HERE, "_ref")

Q_GRIDS = {"cfg1": (0.001, 0.5, 2000), "cfg2": (0.01, 0.5, 4096), "cfg3": (0.01, 1.0, 8192), "cfg4": (0.01, 1.0, 8192)}
TEMPLATE_OF = {"cfg1": "cfg1", "cfg2": "cfg2", "cfg3": "cfg3", "cfg4": "cfg4", "cfg5": "cfg4", "cfg5s": "cfg2"}
# free parameters of the fit workloads
FREE_NAMES = {"cfg4": ("noise__I0", "x__I0", "x__a", "x__hwhm_g", "x__hwhm_l", "x__r", "gp__I0", "gp__rg"),
              "cfg2": ("noise__I0", "p__I0", "p__r", "p__sigma", "p__r_hard", "p__v_fraction")}


def _unjson(o):
    if isinstance(o, dict):
        if set(o.keys()) == {"__float__"}:
            return float(o["__float__"])
        return {k: _unjson(v) for k, v in o.items()}
    if isinstance(o, list):
        return [_unjson(v) for v in o]
    return o


class PlainLayout(object):
    """names / values / slot() of a System in the reference's flatten_params() order -- what the parameter
    generators of tests/helpers.py need, without the product's lowering."""

    def __init__(self, names, values):
        self.names, self.values = list(names), np.array(values, dtype=float)
        self.n_params = len(self.names)

    def slot(self, key):
        return self.names.index(key)


def load_workload(name):
    """(system dict template, PlainLayout, q) of a BASELINE.json workload."""
    with open(os.path.join(ROOT, "tests", "golden", "workloads.json")) as f:
        rec = _unjson(json.load(f))[TEMPLATE_OF[name]]
    lo, hi, n = Q_GRIDS[TEMPLATE_OF[name]]
    return rec["system"], PlainLayout(rec["names"], rec["values"]), np.linspace(lo, hi, n)


def workload_params(name, layout, n, seed=None):
    """The seeded synthetic parameter sets of a workload (same generators as the GPU arm and the tests)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    gen = {"cfg2": helpers.cfg2_params, "cfg3": helpers.cfg3_params, "cfg4": helpers.cfg4_params}[TEMPLATE_OF[name]]
    return gen(layout, n) if seed is None else gen(layout, n, seed)


def reference_available():
    return os.path.isdir(os.path.join(REF_COPY, "xrsdkit"))


class Backend(object):
    """One worker's evaluator: intensity(row) and objective(row, I_obs) for parameter rows of a workload."""

    def __init__(self, workload, force_port=False):
        for v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
            os.environ[v] = "1"
        if ROOT not in sys.path:
            sys.path.insert(0, ROOT)
        self.sysdict, self.layout, self.q = load_workload(workload)
        self.kind = "reference" if (reference_available() and not force_port) else "port"
        if self.kind == "reference":
            from oracle import ref_shim
            ref_shim.load(REF_COPY)
            from xrsdkit.system import System
            self.system = System(**self.sysdict)
            self.flat = self.system.flatten_params()          # shared, mutable parameter records (system/__init__.py:211-218)
            assert list(self.flat.keys()) == self.layout.names
        else:
            from oracle import xrsd_oracle as orc
            with open(os.path.join(ROOT, "tests", "golden", "definitions.json")) as f:
                defs = _unjson(json.load(f))
            self.orc = orc
            self.tables = dict(atomic_table=defs["atomic_params"], sg_point_groups=defs["sg_point_groups"],
                               lattice_space_groups=defs["lattice_space_groups"])

    def _set(self, row):
        if self.kind == "reference":
            for name, v in zip(self.layout.names, row):
                self.flat[name]["value"] = float(v)
        else:
            for name, v in zip(self.layout.names, row):
                pop, par = name.split("__")
                self.sysdict[pop]["parameters"][par]["value"] = float(v)

    def intensity(self, row):
        self._set(row)
        if self.kind == "reference":
            return self.system.compute_intensity(self.q)
        return self.orc.system_intensity(self.q, self.sysdict, **self.tables)

    def objective(self, row, I_obs):
        self._set(row)
        if self.kind == "reference":
            return float(self.system.evaluate_residual(self.q, I_obs))
        return float(self.orc.residual(self.q, I_obs, self.orc.system_intensity(self.q, self.sysdict, **self.tables)))


_B = {}


def _init(workload, force_port):
    _B["be"] = Backend(workload, force_port)
    _B["P"] = workload_params(workload, _B["be"].layout, 4096 if TEMPLATE_OF[workload] == "cfg2" else 256)


def _eval(args):
    lo, hi = args
    be, P = _B["be"], _B["P"]
    acc = 0.0
    for b in range(lo, hi):
        acc += float(be.intensity(P[b % len(P)])[0])
    return hi - lo, acc


def _fit(args):
    """One fit the way the reference does it (system/__init__.py:220-278: Nelder-Mead on the scalar objective, one
    full compute_intensity per evaluation), with scipy's Nelder-Mead standing in for lmfit's wrapper of the same
    routine (lmfit is not installable here).  Returns (seconds, evaluations, initial objective, final objective,
    fitted row)."""
    start_row, iobs_row, free, maxfev = args
    from scipy.optimize import minimize
    be = _B["be"]
    start_row = np.array(start_row, dtype=float)

    def objective(x):
        if np.any(x <= 0):
            return 1e30
        row = start_row.copy()
        row[free] = x
        v = be.objective(row, iobs_row)
        return v if np.isfinite(v) else 1e30

    t0 = time.perf_counter()
    f0 = objective(start_row[free])
    r = minimize(objective, start_row[free], method="Nelder-Mead", options=dict(xatol=1e-4, fatol=1e-4, maxfev=maxfev))
    out = start_row.copy()
    out[free] = r.x
    return time.perf_counter() - t0, int(r.nfev), float(f0), float(r.fun), out


def _kind(_):
    return _B["be"].kind


class CpuArm(object):
    """A pool of host processes (one per core), each holding one Backend."""

    def __init__(self, workload, cores, force_port=False):
        self.cores, self.workload = cores, workload
        self.pool = mp.get_context("spawn").Pool(cores, initializer=_init, initargs=(workload, force_port))
        self.pool.map(_eval, [(0, 1)] * cores)          # import + first-touch warm-up
        self.kind = self.pool.map(_kind, [0])[0]

    def step(self, sets_per_core, offset=0):
        t0 = time.perf_counter()
        jobs = [(offset + c * sets_per_core, offset + (c + 1) * sets_per_core) for c in range(self.cores)]
        done = sum(n for n, _ in self.pool.map(_eval, jobs))
        return done, time.perf_counter() - t0

    def fits(self, start, iobs, free, maxfev=1500):
        """One Nelder-Mead fit per row (rows are spread over the cores); returns fits/s and the per-fit records."""
        t0 = time.perf_counter()
        recs = self.pool.map(_fit, [(start[i], iobs[i], list(free), maxfev) for i in range(len(start))], chunksize=1)
        return len(start) / (time.perf_counter() - t0), recs

    def close(self):
        self.pool.close()
        self.pool.join()


def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def describe(kind):
    return ("the unmodified reference (oracle/_ref/xrsdkit, System.compute_intensity)" if kind == "reference"
            else "numpy oracle port of the reference (oracle/xrsd_oracle.py)")
