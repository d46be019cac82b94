"""CPU checks of the arithmetic of the Voigt evaluations (csrc/faddeeva.cuh).

csrc/cheb_matrices.cuh (generated): the matrices that take Chebyshev node values to the even / odd power-series
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


# ---- numpy restatements of the two Voigt evaluations of csrc/faddeeva.cuh, against scipy's wofz (the function the
# reference calls, tools/peak_math.py:46).  They pin the ARITHMETIC the kernels use (the -m gpu tests pin the kernels).
_C = [1.0, 0.5, 0.75, 1.875, 6.5625, 29.53125, 162.421875, 1055.7421875, 7918.06640625, 67303.564453125,
      639383.8623046875, 6713530.554199219, 77205601.37329102]          # (2k-1)!! / 2^k


def _rew_far_raw(x, y):
    """sqrt(pi) Re w(x + iy) for |z|^2 >= 64 as rew_far_raw sums it: Clenshaw's backward recurrence over
    u_k+1 = al u_k - b u_k-1, truncated at k = 5 / 7 / 12 by |z|^2."""
    z2 = x * x + y * y
    t = 1.0 / z2
    b = t * t
    al = (t + t) - 4.0 * (y * y) * b
    out = np.empty_like(x)
    for lo, hi, n in ((1024.0, np.inf, 5), (256.0, 1024.0, 7), (64.0, 256.0, 12)):
        m = (z2 >= lo) & (z2 < hi)
        B1 = np.zeros(m.sum())
        B2 = np.zeros(m.sum())
        for k in range(n, -1, -1):
            B1, B2 = al[m] * B1 + (-b[m] * B2 + _C[k]), B1
        out[m] = y * (B2 * b[m] + B1 * t[m])
    return out


@pytest.mark.parametrize("y", [1e-10, 1e-6, 1e-3, 0.1, 1.0, 3.0, 7.9, 8.0, 20.0, 1e3])
def test_far_branch_series_matches_wofz(y):
    from scipy.special import wofz
    rs = np.random.RandomState(11)
    x = np.concatenate([np.linspace(0.0, 40.0, 4001), 10 ** rs.uniform(0.9, 6.0, 4000)])
    x = x[x * x + y * y >= 64.0]
    ref = np.real(wofz(x + 1j * y)) * np.sqrt(np.pi)
    assert np.max(np.abs(_rew_far_raw(x, y) - ref) / ref) < 2e-13       # the truncation at |z| = 8


@pytest.mark.parametrize("y", [1e-4, 1e-3, 1e-2, 0.1, 0.5, 1.0, 2.0, 4.0, 7.99])
def test_near_branch_table_matches_wofz(y):
    """The per-pattern table of the near branch: 16 intervals of 1/2 in x, 14 nodes each, rows (e_0, o_0, ..., e_6, o_6)
    from XRSD_VT_M, evaluated as E(u) + t O(u) (voigt_table / eo14 / rew_table)."""
    from scipy.special import wofz
    arr = _arrays()
    M, nodes = arr["XRSD_VT_M"].reshape(14, 14), arr["XRSD_VT_NODES"]
    x = np.linspace(0.0, 8.0, 6001)[:-1]
    x = x[x * x + y * y < 64.0]
    k = np.floor(2.0 * x).astype(int)
    t = 2.0 * (2.0 * x - k) - 1.0
    u = 2.0 * t * t - 1.0
    got = np.empty_like(x)
    for kk in np.unique(k):
        c = M @ np.real(wofz(0.5 * (kk + 0.5 + 0.5 * nodes) + 1j * y))
        m = k == kk
        got[m] = np.polyval(c[0::2][::-1], u[m]) + t[m] * np.polyval(c[1::2][::-1], u[m])
    ref = np.real(wofz(x + 1j * y))
    assert np.max(np.abs(got - ref) / ref) < 2e-13
