"""csrc/cheb_matrices.cuh (generated): the matrices that take Chebyshev node values to the even / odd power-series
coefficients the crystalline kernels evaluate (far-field interpolant, Voigt near-branch table).  CPU checks of the
committed header: it is what the generator writes, and E(u) + t O(u) built from it IS the Chebyshev interpolant."""
import os
import re

import numpy as np
import pytest
from numpy.polynomial import chebyshev as Ch

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "xrsdkit_b200", "csrc")


def _arrays():
    text = open(os.path.join(CSRC, "cheb_matrices.cuh")).read()
    out = {}
    for name, n, body in re.findall(r"const double (\w+)\[(\d+)\] = \{(.*?)\};", text, flags=re.S):
        vals = np.array([float(x) for x in body.replace("\n", " ").split(",") if x.strip()])
        assert len(vals) == int(n), name
        out[name] = vals
    return out


@pytest.mark.parametrize("name,N", [("XRSD_FAR", 15), ("XRSD_VT", 14)])
def test_power_series_rows_reproduce_the_chebyshev_interpolant(name, N):
    arr = _arrays()
    M, nodes = arr[name + "_M"].reshape(N, N), arr[name + "_NODES"]
    assert np.allclose(nodes, np.cos(np.pi * (np.arange(N) + 0.5) / N), rtol=0, atol=5e-16)
    assert np.abs(M).sum(axis=1).max() < 150.0          # the conditioning the header's accuracy argument rests on
    rs = np.random.RandomState(3)
    t = np.linspace(-1.0, 1.0, 301)
    u = 2.0 * t * t - 1.0
    for _ in range(20):
        # a smooth function of the kind both interpolants see: poles a few interval half-widths away
        poles = rs.uniform(4.0, 40.0, 6) * rs.choice([-1.0, 1.0], 6)
        w = rs.uniform(0.1, 3.0, 6)

        def f(x):
            return sum(wi / ((x - p) ** 2 + 0.01) for wi, p in zip(w, poles))

        B = f(nodes)
        k = np.arange(N)
        a = 2.0 / N * np.cos(np.pi * np.outer(k, np.arange(N) + 0.5) / N) @ B
        a[0] *= 0.5
        ref = Ch.chebval(t, a)
        c = M @ B                                         # rows e_0, o_0, e_1, o_1, ...
        e, o = c[0::2], c[1::2]
        E = np.polyval(e[::-1], u)
        O = np.polyval(o[::-1], u)
        got = E + t * O
        assert np.max(np.abs(got - ref) / np.abs(ref)) < 5e-14


def test_header_is_what_the_generator_writes():
    mp = pytest.importorskip("mpmath")          # noqa: F841
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_cheb_matrices", os.path.join(CSRC, "gen_cheb_matrices.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    arr = _arrays()
    for name, N in (("XRSD_FAR", 15), ("XRSD_VT", 14)):
        rows, nodes = gen.matrix(N)
        want = np.array([[float(x) for x in r] for r in rows]).ravel()
        assert np.array_equal(want, arr[name + "_M"])
        assert np.array_equal(np.array([float(x) for x in nodes]), arr[name + "_NODES"])
