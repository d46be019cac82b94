"""TEST INFRASTRUCTURE ONLY -- CPU (numpy) restatement of xrsdkit's intensity path.

This module is the parity oracle for the CUDA engine in xrsdkit_b200.  It restates,
function by function, what the reference computes (file:line given per function,
paths relative to the reference root) with the same float64 operation order where
that order matters.  It is *pinned*: tests/test_oracle_golden.py checks it against
golden vectors produced by running the unmodified reference in the build container
(oracle/make_golden.py -> tests/golden/*.json|npz) and against the known-answer
values recorded in SURVEY.md section 8c.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import it.  The product package never does.

Third-party arithmetic on the path: scipy.special.wofz (Faddeeva, reference
tools/peak_math.py:2,46; scipy is unpinned in the reference's requirements.txt:3,
the container has scipy 1.18.1).  The oracle calls the same scipy routine; the CUDA
engine carries its own implementation (csrc/faddeeva.cuh).

The fit driver (reference system/__init__.py:220-278) needs lmfit, which is neither
vendored nor installed: fit-level parity is unpinned and defined on outcomes only
(see DESIGN.md).
"""
import math

import numpy as np
from scipy.special import wofz

TWO_PI = 2.0 * np.pi

# --------------------------------------------------------------------------------------
# form factors
# --------------------------------------------------------------------------------------

def sphere_ff(q, r):
    """Sphere amplitude 3(sin x - x cos x)/x^3, x=q*r.  reference: scattering/form_factors.py:20-28.
    Only a zero in q[0] is special-cased (-> 1); zeros elsewhere give NaN, as in the reference."""
    q = np.asarray(q, dtype=float)
    x = q * r
    if q[0] == 0:
        out = np.zeros(q.shape)
        out[0] = 1.0
        xt = x[1:]
        out[1:] = 3.0 * (np.sin(xt) - xt * np.cos(xt)) * xt ** -3
        return out
    return 3.0 * (np.sin(x) - x * np.cos(x)) * x ** -3


def atomic_ff(q, Z, a, b):
    """Normalised 4-Gaussian atomic form factor.  reference: scattering/form_factors.py:8-18."""
    q = np.asarray(q, dtype=float)
    g = q * 1.0 / (2 * np.pi)
    s = g * 1.0 / 2
    s2 = s ** 2
    acc = np.sum([ai * np.exp(-1 * bi * s2) for ai, bi in zip(a, b)], axis=0)
    return (Z - 41.78214 * s2 * acc) / Z


def guinier_porod(q, rg, D):
    """Guinier-Porod intensity, I(0)=1.  reference: scattering/__init__.py:69-108."""
    q = np.asarray(q, dtype=float)
    q_splice = 1.0 / rg * np.sqrt(3.0 / 2 * D)
    lo = q <= q_splice
    hi = q > q_splice
    porod = 1.0 * np.exp(-1.0 / 2 * D) * (3.0 / 2 * D) ** (1.0 / 2 * D) * 1.0 / (rg ** D)
    out = np.zeros(q.shape)
    if lo.any():
        out[lo] = 1.0 * np.exp(-1.0 / 3 * q[lo] ** 2 * rg ** 2)
    if hi.any():
        out[hi] = porod * 1.0 / (q[hi] ** D)
    return out


def normal_radius_grid(r0, sigma, width, step):
    """(r_min, r_max, dr) of the size quadrature.  reference: tools/__init__.py:168-173."""
    sigma_r = sigma * r0
    dr = sigma_r * step
    r_min = np.max([r0 - width * sigma_r, dr])
    r_max = r0 + width * sigma_r
    return r_min, r_max, dr


def sphere_normal_intensity(q, r0, sigma, width=3.5, step=0.05):
    """Normally distributed sphere radii, I(0)=1.  reference: scattering/__init__.py:111-161."""
    q = np.asarray(q, dtype=float)
    if sigma < 1.0e-9:
        v0 = float(4) / 3 * np.pi * r0 ** 3
        norm = v0 ** 2
        acc = norm * sphere_ff(q, r0) ** 2
    else:
        acc = np.zeros(q.shape)
        r_min, r_max, dr = normal_radius_grid(r0, sigma, width, step)
        sigma_r = sigma * r0
        norm = 0
        for ri in np.arange(r_min, r_max, dr):
            vi = float(4) / 3 * np.pi * ri ** 3
            rho = 1.0 / (np.sqrt(2 * np.pi) * sigma_r) * np.exp(-1 * (r0 - ri) ** 2 / (2 * sigma_r ** 2))
            wi = vi ** 2 * rho * dr
            norm += wi
            acc += wi * sphere_ff(q, ri) ** 2
    return acc / norm


# --------------------------------------------------------------------------------------
# disordered structure factor
# --------------------------------------------------------------------------------------

def hard_sphere_sf(q, r_hard, v_fraction):
    """Percus-Yevick S(q)/S(0).  reference: scattering/structure_factors.py:3-47."""
    q = np.asarray(q, dtype=float)
    p = v_fraction
    d = 2 * r_hard
    qd = q * d
    qd2, qd3, qd4, qd6 = qd ** 2, qd ** 3, qd ** 4, qd ** 6
    s, c = np.sin(qd), np.cos(qd)
    l1 = (1 + 2 * p) ** 2 / (1 - p) ** 4
    l2 = -1 * (1 + p / 2) ** 2 / (1 - p) ** 4
    nc0 = -2 * p * (4 * l1 + 18 * p * l2 + p * l1)

    def _nc(qd, qd2, qd3, qd4, qd6, s, c):
        return -24 * p * (
            l1 * ((s - qd * c) / qd3)
            - 6 * p * l2 * ((qd2 * c - 2 * qd * s - 2 * c + 2) / qd4)
            - p * l1 / 2 * ((qd4 * c - 4 * qd3 * s - 12 * qd2 * c + 24 * qd * s + 24 * c - 24) / qd6)
        )

    if q[0] == 0.0:
        nc = np.zeros(len(q))
        nc[0] = nc0
        t = slice(1, None)
        nc[t] = _nc(qd[t], qd2[t], qd3[t], qd4[t], qd6[t], s[t], c[t])
    else:
        nc = _nc(qd, qd2, qd3, qd4, qd6, s, c)
    return (1 / (1 - nc)) / (1 / (1 - nc0))


# --------------------------------------------------------------------------------------
# peak profiles
# --------------------------------------------------------------------------------------

def gaussian(x, hwhm):
    """reference: tools/peak_math.py:24-29."""
    return np.sqrt(np.log(2) / np.pi) / hwhm * np.exp(-(x / hwhm) ** 2 * np.log(2))


def lorentzian(x, hwhm):
    """reference: tools/peak_math.py:31-36."""
    return hwhm / np.pi / (x ** 2 + hwhm ** 2)


def voigt(x, hwhm_g, hwhm_l):
    """reference: tools/peak_math.py:38-47 (scipy.special.wofz)."""
    sigma = hwhm_g / np.sqrt(2 * np.log(2))
    gamma = hwhm_l
    return np.real(wofz((x + 1j * gamma) / sigma / np.sqrt(2))) / sigma / np.sqrt(2 * np.pi)


def make_profile(kind, values):
    """Profile closure pk(q, q_center).  reference: scattering/__init__.py:33-38, peak_math.py:6-22."""
    if kind == "voigt":
        hg, hl = values["hwhm_g"], values["hwhm_l"]
        return lambda q, qc: voigt(q - qc, hg, hl)
    if kind == "gaussian":
        h = values["hwhm"]
        return lambda q, qc: gaussian(q - qc, h)
    if kind == "lorentzian":
        h = values["hwhm"]
        return lambda q, qc: lorentzian(q - qc, h)
    raise KeyError(kind)


# --------------------------------------------------------------------------------------
# lattice geometry (reference: definitions.py:462-570)
# --------------------------------------------------------------------------------------

_CUBIC = ("P_cubic", "I_cubic", "F_cubic", "diamond")
_ORTHO = ("P_orthorhombic", "C_orthorhombic", "A_orthorhombic", "I_orthorhombic", "F_orthorhombic")


def lattice_vectors(lattice, a=None, b=None, c=None, alpha=None, beta=None, gamma=None):
    """Cartesian cell vectors.  reference: definitions.py:504-551."""
    if lattice in _CUBIC:
        return [a, 0.0, 0.0], [0.0, a, 0.0], [0.0, 0.0, a]
    if lattice == "hcp":
        return [a, 0.0, 0.0], [0.5 * a, np.sqrt(3.0) / 2 * a, 0.0], [0.0, 0.0, np.sqrt(8.0 / 3.0) * a]
    if lattice == "hexagonal":
        return [a, 0.0, 0.0], [0.5 * a, np.sqrt(3.0) / 2 * a, 0.0], [0.0, 0.0, c]
    if lattice in _ORTHO:
        return [a, 0.0, 0.0], [0.0, b, 0.0], [0.0, 0.0, c]
    if lattice in ("P_tetragonal", "I_tetragonal"):
        return [a, 0.0, 0.0], [0.0, a, 0.0], [0.0, 0.0, c]
    if lattice in ("P_monoclinic", "C_monoclinic"):
        br = float(beta) * np.pi / 180.0
        return [a, 0.0, 0.0], [0.0, b, 0.0], [c * np.cos(br), 0.0, c * np.sin(br)]
    if lattice == "rhombohedral":
        ar = float(alpha) * np.pi / 180.0
        ca = np.cos(ar)
        omega = a ** 3 * np.sqrt(1.0 - ca ** 2 - ca ** 2 - ca ** 2 + 2 * ca * ca * ca)
        cy = a * (ca - ca ** 2) / np.sin(ar)
        cz = omega / (a ** 2 * np.sin(ar))
        return [a, 0.0, 0.0], [a * ca, a * np.sin(ar), 0.0], [a * ca, cy, cz]
    if lattice == "triclinic":
        ar = float(alpha) * np.pi / 180.0
        br = float(beta) * np.pi / 180.0
        gr = float(gamma) * np.pi / 180.0
        omega = a * b * c * np.sqrt(
            1.0 - np.cos(ar) ** 2 - np.cos(br) ** 2 - np.cos(gr) ** 2
            + 2 * np.cos(ar) * np.cos(br) * np.cos(gr))
        cy = c * (np.cos(ar) - np.cos(br) * np.cos(gr)) / np.sin(gr)
        cz = omega / (a * b * np.sin(gr))
        return [a, 0.0, 0.0], [b * np.cos(gr), b * np.sin(gr), 0.0], [c * np.cos(br), cy, cz]
    raise KeyError(lattice)


def reciprocal_vectors(a1, a2, a3):
    """Crystallographic (no 2 pi) reciprocal vectors.  reference: definitions.py:553-570."""
    x1, x2, x3 = np.cross(a2, a3), np.cross(a3, a1), np.cross(a1, a2)
    vol = np.dot(a1, x1)
    return x1 / vol, x2 / vol, x3 / vol


def basis_sites(lattice):
    """Centring / basis sites of the conventional cell.  reference: definitions.py:462-502."""
    if lattice in ("triclinic", "P_monoclinic", "P_orthorhombic", "P_tetragonal",
                   "rhombohedral", "hexagonal", "P_cubic"):
        return np.array([[0.0, 0.0, 0.0]])
    if lattice in ("C_monoclinic", "C_orthorhombic"):
        return np.array([[0.0, 0.0, 0.0], [0.0, 0.5, 0.5]])
    if lattice == "A_orthorhombic":
        return np.array([[0.0, 0.0, 0.0], [0.5, 0.5, 0.0]])
    if lattice in ("I_orthorhombic", "I_tetragonal", "I_cubic"):
        return np.array([[0.0, 0.0, 0.0], [0.5, 0.5, 0.5]])
    if lattice in ("F_orthorhombic", "F_cubic"):
        return np.array([[0.0, 0.0, 0.0], [0.5, 0.5, 0.0], [0.5, 0.0, 0.5], [0.0, 0.5, 0.5]])
    if lattice == "hcp":
        return np.array([[0.0, 0.0, 0.0], [2.0 / 3, 1.0 / 3, 0.5]])
    if lattice == "diamond":
        return np.array([[0.0, 0.0, 0.0], [0.5, 0.5, 0.0], [0.5, 0.0, 0.5], [0.0, 0.5, 0.5],
                         [0.25, 0.25, 0.25], [0.75, 0.75, 0.25], [0.75, 0.25, 0.75], [0.25, 0.75, 0.75]])
    raise KeyError(lattice)


# lattice-parameter names per lattice (reference: definitions.py:262-288)
def lattice_param_names(lattice):
    if lattice in ("P_cubic", "I_cubic", "F_cubic", "diamond", "hcp"):
        return ["a"]
    if lattice in ("hexagonal", "P_tetragonal", "I_tetragonal"):
        return ["a", "c"]
    if lattice == "rhombohedral":
        return ["a", "alpha"]
    if lattice in ("P_orthorhombic", "C_orthorhombic", "I_orthorhombic", "F_orthorhombic"):
        return ["a", "b", "c"]
    if lattice in ("P_monoclinic", "C_monoclinic"):
        return ["a", "b", "c", "beta"]
    if lattice == "triclinic":
        return ["a", "b", "c", "alpha", "beta", "gamma"]
    return []  # e.g. A_orthorhombic: the reference defines no lattice parameters for it


# --------------------------------------------------------------------------------------
# symmetry reduction (reference: scattering/symmetries.py)
# --------------------------------------------------------------------------------------

def _m(rows):
    return np.array(rows, dtype=float)


SYM_OPS = {
    "inversion": _m([[-1, 0, 0], [0, -1, 0], [0, 0, -1]]),
    "mirror_x": _m([[-1, 0, 0], [0, 1, 0], [0, 0, 1]]),
    "mirror_y": _m([[1, 0, 0], [0, -1, 0], [0, 0, 1]]),
    "mirror_z": _m([[1, 0, 0], [0, 1, 0], [0, 0, -1]]),
    "mirror_x_y": _m([[0, -1, 0], [-1, 0, 0], [0, 0, 1]]),
    "mirror_y_z": _m([[1, 0, 0], [0, 0, -1], [0, -1, 0]]),
    "mirror_z_x": _m([[0, 0, -1], [0, 1, 0], [-1, 0, 0]]),
    "mirror_nx_y": _m([[0, 1, 0], [1, 0, 0], [0, 0, 1]]),
    "mirror_ny_z": _m([[1, 0, 0], [0, 0, 1], [0, 1, 0]]),
    "mirror_nz_x": _m([[0, 0, 1], [0, 1, 0], [1, 0, 0]]),
}

# operations the reference applies per point group (symmetries.py:27-42); every point group
# not listed here has `None` in the reference and reduces nothing.
POINT_GROUP_OPS = {
    "1": [],
    "-1": ["inversion"],
    "m-3m": ["mirror_x", "mirror_y", "mirror_z", "mirror_x_y", "mirror_y_z", "mirror_z_x",
             "mirror_nx_y", "mirror_ny_z", "mirror_nz_x"],
    "6/mmm": ["inversion", "mirror_z", "mirror_x_y", "mirror_nx_y"],
}


def reduce_by_symmetry(all_hkl, rlat, point_group, symprec=1.0e-6, row_chunk=None):
    """Greedy, order-dependent mirror reduction.  reference: scattering/symmetries.py:44-92.

    `point_group` is the point-group symbol (or None/'' for "no space group").  The
    N x N distance matrix is formed exactly like the reference (a N x 3 x N difference
    tensor, norm over axis 1); `row_chunk` only bounds memory by processing rows in
    slabs, which does not change any value.
    """
    hkl = np.array(all_hkl, dtype=float, copy=True)
    n = hkl.shape[0]
    mult = np.ones(n, dtype=int)
    pts = np.dot(hkl, rlat.T)
    span = np.max(hkl, axis=0) - np.min(hkl, axis=0)
    rank = hkl[:, 0] * (span[1] + 1) * (span[2] + 1) + hkl[:, 1] * (span[2] + 1) + hkl[:, 2]
    op_names = POINT_GROUP_OPS.get(point_group) if point_group else None
    for name in (op_names or []):
        op = SYM_OPS[name]
        sym = np.dot(op, pts.T).T
        m = pts.shape[0]
        if row_chunk is None or row_chunk >= m:
            dist = np.linalg.norm(pts[:, :, np.newaxis] - sym.T, axis=1)
            near = np.argmin(dist, axis=1)
            dmin = np.min(dist, axis=1)
        else:
            near = np.empty(m, dtype=int)
            dmin = np.empty(m)
            for lo in range(0, m, row_chunk):
                hi = min(m, lo + row_chunk)
                dist = np.linalg.norm(pts[lo:hi, :, np.newaxis] - sym.T, axis=1)
                near[lo:hi] = np.argmin(dist, axis=1)
                dmin[lo:hi] = np.min(dist, axis=1)
        drop = (near != np.arange(m)) & (dmin < symprec) & (rank < rank[near])
        if drop.any():
            keep = ~drop
            mult[near[drop]] += mult[drop]
            mult, hkl, pts, rank = mult[keep], hkl[keep, :], pts[keep, :], rank[keep]
    return hkl, mult


# --------------------------------------------------------------------------------------
# crystalline powder diffraction (reference: scattering/__init__.py:165-333)
# --------------------------------------------------------------------------------------

def enumerate_hkl(a_vecs, b_vecs, q_min, q_max):
    """All hkl with G_min < |G| <= G_max, l outer / k / h inner.  reference: scattering/__init__.py:248-278."""
    a1, a2, a3 = a_vecs
    b1, b2, b3 = b_vecs
    g_max = 1.0 / (2 * np.pi / q_max)
    g_min = 1.0 / (2 * np.pi / q_min) if q_min > 0.0 else 0
    n1 = np.ceil(g_max * np.linalg.norm(a1) / np.dot(b1, a1))
    n2 = np.ceil(g_max * np.linalg.norm(a2) / np.dot(b2, a2))
    n3 = np.ceil(g_max * np.linalg.norm(a3) / np.dot(b3, a3))
    hs, ks, ls = np.arange(-1 * n1 + 1, n1), np.arange(-1 * n2 + 1, n2), np.arange(-1 * n3 + 1, n3)
    # same arithmetic as np.linalg.norm(np.dot((h,k,l),(b1,b2,b3))) per candidate, vectorised:
    # the oracle only needs identical membership decisions, checked against the reference.
    L, K, H = np.meshgrid(ls, ks, hs, indexing="ij")
    cand = np.stack([H.ravel(), K.ravel(), L.ravel()], axis=1)
    bmat = np.array([b1, b2, b3])
    g = np.empty(cand.shape[0])
    for i in range(cand.shape[0]):  # keep numpy's per-row dot/norm arithmetic (ties at G_max)
        g[i] = np.linalg.norm(np.dot(cand[i], bmat))
    sel = (g_min < g) & (g <= g_max)
    return cand[sel]


def reflection_table(lattice, latparams, q_min, q_max, point_group, use_symmetry=True, row_chunk=2048):
    """(reduced_hkl, mult, b-matrix) as the reference builds them (scattering/__init__.py:240-284)."""
    a1, a2, a3 = lattice_vectors(lattice, **latparams)
    b1, b2, b3 = reciprocal_vectors(a1, a2, a3)
    all_hkl = enumerate_hkl((a1, a2, a3), (b1, b2, b3), q_min, q_max)
    bmat = np.array([b1, b2, b3])
    if all_hkl.shape[0] == 0:
        return all_hkl.reshape(0, 3), np.ones(0), bmat
    if use_symmetry:
        hkl, mult = reduce_by_symmetry(all_hkl, bmat, point_group, row_chunk=row_chunk)
    else:
        hkl, mult = all_hkl, np.ones(all_hkl.shape[0])
    return hkl, mult, bmat


def powder_intensity(q, wavelength, lattice, latparams, coords, ff_funcs, pk_func,
                     q_min=0.0, q_max=None, point_group=None, sf_mode="local",
                     polz=True, lorentz=True, use_symmetry=True, row_chunk=2048, table=None):
    """Isotropic integrated diffraction, I(0)=1.  reference: scattering/__init__.py:165-333.

    `point_group` replaces the reference's space_group argument (the space-group ->
    point-group table, definitions.py:425-460, is applied by the caller).  `table`
    optionally supplies a prebuilt (hkl, mult, bmat) so the "table-cached" CPU cost
    can be timed separately.
    """
    q = np.asarray(q, dtype=float)
    if not wavelength:
        raise ValueError("cannot compute diffraction with source wavelength of {}".format(wavelength))
    if not q_max:
        q_max = q[-1]
    n_q = len(q)
    theta = np.arcsin(wavelength * q / (4.0 * np.pi))
    pz = np.ones(n_q)
    if polz:
        pz = 0.5 * (1.0 + np.cos(2.0 * theta) ** 2)
    q0 = np.array([0.0])
    if table is None:
        table = reflection_table(lattice, latparams, q_min, q_max, point_group, use_symmetry, row_chunk)
    hkl, mult, bmat = table
    if hkl.shape[0] == 0 or not hkl.any():
        return np.zeros(n_q)

    absq = 2 * np.pi * np.linalg.norm(np.dot(hkl, bmat), axis=1)
    th_hkl = np.arcsin(wavelength * absq / (4.0 * np.pi))
    ltz = np.ones(th_hkl.shape[0])
    if lorentz:
        ltz = 1.0 / (np.sin(th_hkl) * np.sin(2 * th_hkl))

    shells = list(set(absq))
    pk_q = {}
    pk_0 = {}
    for qs in shells:
        pk_q[qs] = pk_func(q, qs)
        pk_0[qs] = pk_func(q0, qs)[0]

    n_r = hkl.shape[0]
    sf = np.zeros((n_r, n_q), dtype=complex)
    sf0 = np.zeros(n_r, dtype=complex)
    sites = basis_sites(lattice)
    shells_arr = np.array(shells)
    for xyz, ff in zip(coords, ff_funcs):
        ff0 = ff(q0)[0]
        if sf_mode == "radial":
            ffq = ff(q)
            for i in range(n_r):
                for site in sites:
                    phase = np.exp(2j * np.pi * np.dot(site + xyz, hkl[i, :]))
                    sf[i, :] += ffq * phase
                    sf0[i] += ff0 * phase
        elif sf_mode == "local":
            ff_at = dict(zip(shells, ff(shells_arr)))
            for i in range(n_r):
                for site in sites:
                    phase = np.exp(2j * np.pi * np.dot(site + xyz, hkl[i, :]))
                    sf[i, :] += ff_at[absq[i]] * phase
                    sf0[i] += ff0 * phase

    acc = np.zeros(n_q)
    norm = 0.0
    for i in range(n_r):
        acc += mult[i] * ltz[i] * (sf[i, :] * sf[i, :].conjugate()).real * pk_q[absq[i]]
        norm += mult[i] * ltz[i] * (sf0[i] * sf0[i].conjugate()).real * pk_0[absq[i]]
    return pz * acc / norm


# --------------------------------------------------------------------------------------
# dispatchers (reference: scattering/__init__.py:12-66, system/noise.py:50-63,
#              system/__init__.py:112-187, tools/__init__.py:104-125)
# --------------------------------------------------------------------------------------

def _val(p):
    return p["value"] if isinstance(p, dict) else p


def population_intensity(q, wavelength, structure, form, settings, parameters,
                         atomic_table=None, sg_point_groups=None, lattice_space_groups=None):
    """One population's I(q).  reference: scattering/__init__.py:12-66.

    `settings` must be fully populated (the reference's Population does that);
    `parameters` maps name -> value or {'value': ...}.  `atomic_table` maps element
    symbol -> {'Z','a','b'}; `sg_point_groups` maps space-group symbol -> point group.
    """
    P = {k: _val(v) for k, v in parameters.items()}
    q = np.asarray(q, dtype=float)

    def _atom(sym):
        rec = atomic_table[sym]
        return lambda qq, rec=rec: atomic_ff(qq, rec["Z"], rec["a"], rec["b"])

    if structure == "crystalline":
        coords = [[0.0, 0.0, 0.0]]
        if form == "spherical":
            ffs = [lambda qq, r=P["r"]: sphere_ff(qq, r)]
        elif form == "atomic":
            ffs = [_atom(settings["symbol"])]
        if form == "polyatomic":
            n_at = settings["n_atoms"]
            coords = [[P["u_%d" % i], P["v_%d" % i], P["w_%d" % i]] for i in range(n_at)]
            ffs = [_atom(settings["symbol_%d" % i]) for i in range(n_at)]
        lattice = settings["lattice"]
        latparams = {nm: P[nm] for nm in lattice_param_names(lattice)}
        pk = make_profile(settings["profile"], P)
        sg = settings["space_group"]
        if sg and lattice_space_groups is not None and sg not in lattice_space_groups[lattice].values():
            raise ValueError("space group {} not valid for {} lattice".format(sg, lattice))
        pg = sg_point_groups[sg] if sg else None
        ixtal = powder_intensity(
            q, wavelength, lattice, latparams, coords, ffs, pk,
            q_min=settings["q_min"], q_max=settings["q_max"], point_group=pg,
            sf_mode=settings["structure_factor_mode"],
            polz=settings["polarization_correction"], lorentz=settings["lorentz_correction"],
            use_symmetry=settings["use_symmetry"])
        return P["I0"] * ixtal
    if form == "atomic":
        ff2 = _atom(settings["symbol"])(q) ** 2
    if form == "spherical":
        if settings["distribution"] == "single":
            ff2 = sphere_ff(q, P["r"]) ** 2
        if settings["distribution"] == "r_normal":
            ff2 = sphere_normal_intensity(q, P["r"], P["sigma"],
                                          settings["sampling_width"], settings["sampling_step"])
    if form == "guinier_porod":
        ff2 = guinier_porod(q, P["rg"], P["D"])
    if structure == "disordered":
        return P["I0"] * hard_sphere_sf(q, P["r_hard"], P["v_fraction"]) * ff2
    if structure == "diffuse":
        return P["I0"] * ff2


def noise_intensity(q, model, parameters):
    """reference: system/noise.py:50-63."""
    P = {k: _val(v) for k, v in parameters.items()}
    n_q = len(q)
    out = np.zeros(n_q)
    if model == "flat":
        out += P["I0"] * np.ones(n_q)
    elif model == "low_q_scatter":
        beam = P["I0"] * (1.0 - P["I0_flat_fraction"])
        out += beam * guinier_porod(q, P["effective_rg"], P["effective_D"])
        out += P["I0"] * P["I0_flat_fraction"] * np.ones(n_q)
    else:
        raise ValueError("unsupported noise specification: {}".format(model))
    return out


_NON_POPULATION_KEYS = ("noise", "features", "fit_report", "sample_metadata")


def system_intensity(q, sysdict, **tables):
    """Noise + sum over populations in dict order.  reference: system/__init__.py:112-130.
    `sysdict` follows System.to_dict() (system/__init__.py:45-53)."""
    q = np.asarray(q, dtype=float)
    noise = sysdict.get("noise", {"model": "flat", "parameters": {"I0": 1.0e-3}})
    total = noise_intensity(q, noise["model"], noise["parameters"])
    wl = sysdict.get("sample_metadata", {}).get("source_wavelength", 1.0)
    for name, pop in sysdict.items():
        if name in _NON_POPULATION_KEYS:
            continue
        total += population_intensity(q, wl, pop["structure"], pop["form"],
                                      pop["settings"], pop["parameters"], **tables)
    return total


def chi2(y1, y2, weights=None):
    """reference: tools/__init__.py:104-125."""
    if weights is None:
        return np.sum((y1 - y2) ** 2)
    weights = weights / np.sum(weights)
    return np.sum((y1 - y2) ** 2 * weights)


def residual(q, I, I_comp, dI=None, error_weighted=True, logI_weighted=True,
             q_range=(0.0, float("inf"))):
    """Scalar fit objective.  reference: system/__init__.py:134-187 (note wts *= dI**2)."""
    q = np.asarray(q, dtype=float)
    I = np.asarray(I, dtype=float)
    fit = (I > 0) & (q >= q_range[0]) & (q <= q_range[1])
    wts = np.ones(len(q))
    if error_weighted:
        if dI is None:
            dI = np.empty(I.shape)
            dI.fill(np.nan)
            dI[fit] = np.sqrt(I[fit])
        wts = wts * dI ** 2
    if logI_weighted:
        fit = fit & (I_comp > 0)
        return chi2(np.log(I_comp[fit]), np.log(I[fit]), wts[fit])
    return chi2(I_comp[fit], I[fit], wts[fit])


# --------------------------------------------------------------------------------------
# pattern featuriser (reference: tools/profiler.py:64-221, tools/peak_math.py:87-123,
#                     tools/__init__.py:83-102); called by fit() at system/__init__.py:276
# --------------------------------------------------------------------------------------

PROFILE_KEYS = (
    "Imax_over_Imean", "Ilowq_over_Imean", "Imax_sharpness", "I_fluctuation", "logI_fluctuation",
    "logI_max_over_std", "r_fftIcentroid", "q_Icentroid", "q_logIcentroid", "pearson_q", "pearson_q2",
    "pearson_expq", "pearson_invexpq", "q_best_hump", "q_best_trough", "best_hump_qwidth",
    "best_trough_qwidth", "q_best_hump_log", "q_best_trough_log", "best_hump_qwidth_log",
    "best_trough_qwidth_log")


def pearson_r(u, v):
    """reference: tools/__init__.py:83-102 (centred sums, no 1/n factors)."""
    du = u - np.mean(u)
    dv = v - np.mean(v)
    return np.sum(du * dv) / (np.sqrt(np.sum(du ** 2)) * np.sqrt(np.sum(dv ** 2)))


def _trapezoid_centroid(x, y):
    """(sum of x_mid*dx*y_mid) / (sum of dx*y_mid); reference: profiler.py:124-137."""
    dx = np.diff(x)
    ym = 0.5 * (y[1:] + y[:-1])
    xm = 0.5 * (x[1:] + x[:-1])
    return np.sum(xm * dx * ym) / np.sum(dx * ym)


def _turning_sum(y):
    """Sum of |first differences| at the first point and wherever the difference changes
    sign against its predecessor; reference: profiler.py:140-150."""
    d = np.diff(y)
    keep = np.ones(len(d), dtype=bool)
    keep[1:] = d[1:] * d[:-1] < 0
    return np.sum(np.abs(d[keep]))


def hump_trough_scores(x, y, w=50):
    """reference: tools/peak_math.py:87-123.  For every point, the Pearson correlation of the
    values in a +-w window with the squared distance from the point; hump score = -y*r,
    trough score = std(window)*r."""
    n = len(y)
    hump = np.zeros(n)
    trough = np.zeros(n)
    for i in range(n):
        lo, hi = max(0, i - w), min(n, i + w + 1)
        yw = y[lo:hi]
        r = pearson_r(yw, (x[lo:hi] - x[i]) ** 2)
        hump[i] = -1 * y[i] * r
        trough[i] = np.std(yw) * r
    return hump, trough


def _parabola(x, y):
    """Vertex abscissa and focal width of the degree-2 least-squares fit (np.polyfit as in the
    reference, profiler.py:181-196)."""
    p = np.polyfit(x, y, 2)
    return -1 * p[1] / (2 * p[0]), abs(1.0 / p[0])


def profile_pattern(q, I):
    """The 21 features of reference tools/profiler.py:64-221, as an array in PROFILE_KEYS order."""
    q = np.asarray(q, dtype=float)
    I = np.asarray(I, dtype=float)
    with np.errstate(all="ignore"):
        i_max = int(np.argmax(I))
        I_max, I_min = I[i_max], I[int(np.argmin(I))]
        I_mean, I_std = np.mean(I), np.std(I)
        q_mean, q_std = np.mean(q), np.std(q)
        Is = (I - I_mean) / I_std
        qs = (q - q_mean) / q_std
        low = q < q[0] + 0.1 * (q[-1] - q[0])
        pos = I > 0
        q_p, qs_p, Is_p = q[pos], qs[pos], Is[pos]
        L = np.log(I[pos])
        L_max, L_min, L_std, L_mean = np.max(L), np.min(L), np.std(L), np.mean(L)
        Ls = (L - L_mean) / L_std
        near_max = (q > 0.9 * q[i_max]) & (q < 1.1 * q[i_max])

        amp = np.abs(np.fft.fft(I))
        r = np.fft.fftfreq(q.shape[-1])
        rp = r > 0

        hump, trough = hump_trough_scores(qs_p, Ls)
        i_h, i_t = int(np.argmax(hump)), int(np.argmax(trough))
        sel_h = (q_p > q_p[i_h] - 0.1 * q_std) & (q_p < q_p[i_h] + 0.1 * q_std)
        sel_t = (q_p > q_p[i_t] - 0.1 * q_std) & (q_p < q_p[i_t] + 0.1 * q_std)
        vh, wh = _parabola(qs_p[sel_h], Is_p[sel_h])
        vt, wt = _parabola(qs_p[sel_t], Is_p[sel_t])
        vhl, whl = _parabola(qs_p[sel_h], Ls[sel_h])
        vtl, wtl = _parabola(qs_p[sel_t], Ls[sel_t])

        return np.array([
            I_max / I_mean,
            np.mean(I[low]) / I_mean,
            I_max / np.mean(I[near_max]),
            _turning_sum(I) / (I_max - I_min),
            _turning_sum(L) / (L_max - L_min),
            L_max / L_std,
            _trapezoid_centroid(r[rp], amp[rp]),
            _trapezoid_centroid(q, I),
            _trapezoid_centroid(q_p, L),
            pearson_r(q, I), pearson_r(q ** 2, I), pearson_r(np.exp(q), I), pearson_r(np.exp(-1 * q), I),
            vh * q_std + q_mean, vt * q_std + q_mean, wh * q_std, wt * q_std,
            vhl * q_std + q_mean, vtl * q_std + q_mean, whl * q_std, wtl * q_std])
