"""TEST / BENCH INFRASTRUCTURE ONLY -- dump the benchmark workloads' Systems as the UNMODIFIED reference builds them.

Run in the build container (where /root/reference exists):

    python oracle/make_workloads.py          # -> tests/golden/workloads.json

For each BASELINE.json workload (cfg1..cfg4; cfg5 fits cfg2- and cfg4-like systems) the file holds the fully
populated System.to_dict() of the reference (defaults filled by the reference's own Population / NoiseModel)
and the parameter names in the reference's flatten_params() order (system/__init__.py:211-218).  bench.py's
`--impl reference` arm builds its parameter matrices from these with plain numpy -- it imports nothing from the
product package -- and tests/test_oracle_golden.py checks that the product lowers the same systems to the same
parameter order.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref_shim  # noqa: E402

ref_shim.load()
from xrsdkit.system import System  # noqa: E402
from helpers import cfg2_system, cfg3_system, cfg4_system  # noqa: E402
from oracle.make_golden import jsonable  # noqa: E402


def cfg1_system(S):
    return S(p=dict(structure="diffuse", form="spherical", parameters={"r": {"value": 20.0}, "I0": {"value": 1.0}}))


def main():
    out = {}
    for name, make in (("cfg1", cfg1_system), ("cfg2", cfg2_system), ("cfg3", cfg3_system), ("cfg4", cfg4_system)):
        s = make(System)
        flat = s.flatten_params()
        out[name] = {"system": jsonable(s.to_dict()), "names": list(flat.keys()),
                     "values": [float(r["value"]) for r in flat.values()]}
    path = os.path.join(ROOT, "tests", "golden", "workloads.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=False)
    print("wrote", path)


if __name__ == "__main__":
    main()
