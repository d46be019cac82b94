"""Print the parity of the CUDA engine against every golden system (run on the GPU box)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import importlib.util as iu
sp = iu.spec_from_file_location("cf", os.path.join(ROOT, "tests", "conftest.py")); cf = iu.module_from_spec(sp); sp.loader.exec_module(cf)
from xrsdkit_b200.system import System

meta = cf._unjson(json.load(open(os.path.join(ROOT, "tests/golden/systems.json"))))
arr = np.load(os.path.join(ROOT, "tests/golden/systems.npz"))
worst = []
for m in meta:
    q = arr[m["q"]]; ref = arr[m["key"] + "_I"]
    try:
        t0 = time.time()
        s = System(**m["system"])
        got = s.compute_intensity(q)
        dt = time.time() - t0
    except Exception as e:
        print("%s %-40s EXC %r" % (m["key"], m["name"][:40], e)); continue
    fin = np.isfinite(ref)
    same_nan = np.array_equal(np.isfinite(got), fin)
    scale = np.max(np.abs(ref[fin])) if fin.any() else 1.0
    rel = np.abs(got[fin] - ref[fin]) / np.maximum(np.abs(ref[fin]), 1e-300)
    relf = np.abs(got[fin] - ref[fin]) / (np.abs(ref[fin]) + 1e-9 * scale)
    i = int(np.argmax(rel)) if rel.size else 0
    print("%s %-44s n=%5d rel=%.2e (q=%.4g ref=%.3e) relfloor=%.2e nanpattern=%s %.0fms" % (
        m["key"], m["name"][-44:], len(q), rel.max() if rel.size else 0, q[fin][i] if rel.size else 0, ref[fin][i] if rel.size else 0,
        relf.max() if relf.size else 0, same_nan, dt * 1e3))
    worst.append((rel.max() if rel.size else 0, m["name"]))
worst.sort(reverse=True)
print("WORST", worst[:10])
