"""TEST INFRASTRUCTURE ONLY -- golden vectors for cells whose candidate hkl box is beyond the table builder's
shared memory (F_cubic with the reference's default q_max = 1.0 and cell edges of 150 and 206 A, the largest edge in
the reference's fixture corpus).

    python oracle/make_golden_large.py        # -> tests/golden/large_cell.npz   (about 15 minutes, one core)

These come from the ORACLE (oracle/xrsd_oracle.py, pinned to the unmodified reference by tests/test_oracle_golden.py
on every case the reference can run), not from the reference itself: its symmetrize_points forms an N x 3 x N float64
tensor (scattering/symmetries.py:70-72) -- 78 GB at a = 150 (N = 56 k reflections), 470 GB at a = 206 (N = 140 k).
The oracle evaluates the same distances in row slabs (row_chunk), which changes no value.
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import xrsd_oracle as orc  # noqa: E402

WL = 0.8265617


def sysdict(a):
    return {"x": {"structure": "crystalline", "form": "spherical",
                  "settings": {"lattice": "F_cubic", "space_group": "Fm-3m", "texture": "random",
                               "integration_mode": "spherical", "use_symmetry": True, "structure_factor_mode": "local",
                               "profile": "voigt", "polarization_correction": True, "lorentz_correction": True,
                               "distribution": "single", "q_min": 0.0, "q_max": 1.0},
                  "parameters": {"I0": {"value": 1.0}, "a": {"value": a}, "hwhm_g": {"value": 2e-3},
                                 "hwhm_l": {"value": 2e-3}, "r": {"value": 0.3 * a}}},
            "noise": {"model": "flat", "parameters": {"I0": {"value": 1e-3}}},
            "sample_metadata": {"source_wavelength": WL}}


def main():
    with open(os.path.join(ROOT, "tests", "golden", "definitions.json")) as f:
        defs = json.load(f)
    tables = dict(atomic_table=defs["atomic_params"], sg_point_groups=defs["sg_point_groups"],
                  lattice_space_groups=defs["lattice_space_groups"])
    q = np.linspace(0.02, 1.0, 300)
    out = {"q": q}
    for a in (150.0, 206.0):
        t0 = time.perf_counter()
        hkl, mult, _ = orc.reflection_table("F_cubic", {"a": a}, 0.0, 1.0, "m-3m", True, row_chunk=1024)
        out["hkl_%d" % a] = hkl.astype(np.int32)
        out["mult_%d" % a] = np.asarray(mult).astype(np.int32)
        print("a = %g: %d reduced reflections, sum of multiplicities %d, %.0f s" %
              (a, len(hkl), int(np.sum(mult)), time.perf_counter() - t0), flush=True)
        t0 = time.perf_counter()
        out["I_%d" % a] = orc.system_intensity(q, sysdict(a), **tables)
        print("   I(q) %.0f s" % (time.perf_counter() - t0), flush=True)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "large_cell.npz"), **out)


if __name__ == "__main__":
    main()
