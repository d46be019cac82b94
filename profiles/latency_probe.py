"""Single-pattern latency of the host path (BASELINE configs[0] and a cfg3 pattern), run on the GPU box:

    python profiles/latency_probe.py                # default build / settings
    XRSD_HOST_ZEROCOPY=0 python profiles/latency_probe.py   # copies staged through pinned memory instead of zero-copy

Prints one JSON line: microseconds per System.compute_intensity call (numpy in, numpy out), the share of the C call
(xrsd_intensity_batch_host) in it, and the largest relative difference from the numpy oracle.
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def per_call(f, n):
    for _ in range(min(n, max(20, n // 10))):
        f()
    t0 = time.perf_counter()
    for _ in range(n):
        f()
    return (time.perf_counter() - t0) / n * 1e6


def main():
    from helpers import cfg3_system
    from oracle import xrsd_oracle as xo
    from xrsdkit_b200 import definitions as d, engine, lowering
    from xrsdkit_b200.system import System
    tabs = dict(atomic_table=d.atomic_params, sg_point_groups=d.sg_point_groups, lattice_space_groups=d.lattice_space_groups)
    eng = engine.get_engine()
    out = {"zero_copy": os.environ.get("XRSD_HOST_ZEROCOPY", "1") != "0"}
    cases = (("cfg1", System(p=dict(structure="diffuse", form="spherical", parameters={"r": {"value": 20.0}})),
              np.linspace(0.001, 0.5, 2000), 3000),
             ("cfg2_one", None, np.linspace(0.01, 0.5, 4096), 2000),
             ("cfg3_one", cfg3_system(System), np.linspace(0.01, 1.0, 8192), 500))
    for name, s, q, n in cases:
        if s is None:
            from helpers import cfg2_system
            s = cfg2_system(System)
        lay = lowering.lower_system(s, q)
        vals = np.ascontiguousarray(lay.values)
        buf = np.empty(q.size)

        def c_call():
            rc = eng.lib.xrsd_intensity_batch_host(C.byref(lay.c), q.ctypes.data, q.size, vals.ctypes.data, vals.size, 1,
                                                   buf.ctypes.data, q.size)
            assert rc == 0
        want = xo.system_intensity(q, s.to_dict(), **tabs)
        got = s.compute_intensity(q)
        out[name] = {"us_per_call": per_call(lambda: s.compute_intensity(q), n), "c_call_us": per_call(c_call, n),
                     "lowering_us": per_call(lambda: lowering.lower_system(s, q), n),
                     "oracle_us_1core": per_call(lambda: xo.system_intensity(q, s.to_dict(), **tabs), 1 if name == "cfg3_one" else 200),
                     "max_rel_diff_vs_oracle": float(np.max(np.abs(got - want) / np.abs(want)))}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
