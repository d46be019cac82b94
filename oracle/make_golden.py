"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/* by running the UNMODIFIED reference.

Run in the build container (where /root/reference exists):

    python oracle/make_golden.py

The reference is imported through oracle/ref_shim.py; every number written below is the
output of the reference's own functions (numpy 2.3.5 / scipy 1.18.1 at generation time).
The fixtures travel with the repo; /root/reference does not.

Files written
  tests/golden/definitions.json   the reference's definition tables (space groups, point
                                  groups, defaults, atomic table) as data
  tests/golden/kernels.npz        known-answer vectors of the elementary functions
  tests/golden/tables.json        reflection tables (hkl, mult) for lattice/space-group cases
  tests/golden/systems.json       System dicts (fully populated, from System.to_dict())
  tests/golden/systems.npz        q grids, I(q), per-population I(q) and residuals for them
"""
import glob
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

ref_shim.load()
import yaml  # noqa: E402
from xrsdkit import definitions as xd  # noqa: E402
from xrsdkit import scattering as xs  # noqa: E402
from xrsdkit.scattering import form_factors, structure_factors, symmetries  # noqa: E402
from xrsdkit.system import System, Population  # noqa: E402
from xrsdkit.tools import peak_math  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
WL = 0.8265617


def jsonable(o):
    if isinstance(o, dict):
        return {str(k): jsonable(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [jsonable(v) for v in o]
    if isinstance(o, (np.floating,)):
        return float(o)
    if isinstance(o, (np.integer,)):
        return int(o)
    if isinstance(o, (np.bool_,)):
        return bool(o)
    if isinstance(o, np.ndarray):
        return jsonable(o.tolist())
    if isinstance(o, float) and (o != o or o in (float("inf"), float("-inf"))):
        return {"__float__": repr(o)}
    return o


def dump_definitions():
    d = dict(
        structure_names=xd.structure_names,
        form_factor_names=xd.form_factor_names,
        noise_model_names=xd.noise_model_names,
        structure_settings=xd.structure_settings,
        form_settings=xd.form_settings,
        form_factor_params=xd.form_factor_params,
        noise_params=xd.noise_params,
        all_lattices=xd.all_lattices,
        bravais_lattices=xd.bravais_lattices,
        crystal_point_groups=xd.crystal_point_groups,
        all_point_groups=xd.all_point_groups,
        lattice_space_groups=xd.lattice_space_groups,
        sg_point_groups=xd.sg_point_groups,
        atomic_params=xd.atomic_params,
        symmetry_ops_defined=[k for k, v in symmetries.symmetry_operations.items() if v is not None],
    )
    # structure_params / additional params for every lattice and profile, as the reference returns them
    sp = {}
    for lat in xd.all_lattices:
        for prof in ("voigt", "gaussian", "lorentzian"):
            sp["%s|%s" % (lat, prof)] = xd.structure_params("crystalline", {"lattice": lat, "profile": prof})
    d["structure_params_crystalline"] = sp
    d["structure_params_disordered"] = xd.structure_params("disordered", {"interaction": "hard_spheres"})
    d["additional_polyatomic_3"] = xd.additional_form_factor_params("polyatomic", {"n_atoms": 3})
    d["additional_r_normal"] = xd.additional_form_factor_params("spherical", {"distribution": "r_normal"})
    lv = {}
    lp = dict(a=11.0, b=13.0, c=17.0, alpha=81.0, beta=97.0, gamma=103.0)
    for lat in xd.all_lattices:
        a1, a2, a3 = xd.lattice_vectors(lat, **lp)
        b1, b2, b3 = xd.reciprocal_lattice_vectors(a1, a2, a3)
        lv[lat] = dict(a=[list(map(float, a1)), list(map(float, a2)), list(map(float, a3))],
                       b=[b1.tolist(), b2.tolist(), b3.tolist()],
                       sites=xd.lattice_coords(lat).tolist())
    d["lattice_geometry_probe"] = dict(latparams=lp, vectors=lv)
    with open(os.path.join(OUT, "definitions.json"), "w") as f:
        json.dump(jsonable(d), f, indent=0, sort_keys=True)


def dump_kernels():
    rs = np.random.RandomState(7)
    q5 = np.array([0.01, 0.05, 0.1, 0.25, 0.5])
    x5 = np.array([-0.01, -0.002, 0, 0.001, 0.05])
    q = np.linspace(0.003, 0.7, 400)
    q0 = np.concatenate([[0.0], q])
    out = dict(q5=q5, x5=x5, q=q, q0=q0)
    out["kat1_sphere_ff_r20"] = form_factors.spherical_ff(q5, 20)
    out["kat2_rnormal"] = xs.spherical_normal_intensity(q5, 20, 0.1, 2.5, 0.1)
    out["kat3_gp"] = xs.guinier_porod_intensity(q5, 10, 4)
    out["kat4_hs"] = structure_factors.hard_sphere_sf(q5, 25, 0.3)
    out["kat5_al"] = form_factors.atomic_ff_normalized(q5, "Al")
    out["kat6_voigt"] = peak_math.voigt(x5, 0.002, 0.0018)
    out["kat7_gauss"] = peak_math.gaussian(x5, 0.002)
    out["kat8_lorentz"] = peak_math.lorentzian(x5, 0.002)
    # wider sweeps
    rr = rs.uniform(3, 80, 12)
    out["sphere_r"] = rr
    out["sphere_ff"] = np.array([form_factors.spherical_ff(q, r) for r in rr])
    out["sphere_ff_q0"] = np.array([form_factors.spherical_ff(q0, r) for r in rr[:3]])
    rn = np.array([[20, 0.1, 2.5, 0.1], [35.5, 0.3, 3.5, 0.05], [12.0, 0.05, 3.5, 0.05], [50.0, 1.5, 3.5, 0.05],
                   [8.0, 0.6, 2.0, 0.2], [25.0, 5e-10, 3.5, 0.05], [17.3, 0.2875, 3.5, 0.05]])
    out["rnormal_args"] = rn
    out["rnormal_I"] = np.array([xs.spherical_normal_intensity(q, *row) for row in rn])
    gp = np.array([[10, 4], [3.3, 2.5], [50, 3.9], [7.0, 0.0], [100.0, 4.0], [5.0, 1.0]])
    out["gp_args"] = gp
    out["gp_I"] = np.array([xs.guinier_porod_intensity(q, *row) for row in gp])
    hs = np.array([[25, 0.3], [10, 0.05], [40.0, 0.5], [15.0, 0.7405], [60.0, 0.01]])
    out["hs_args"] = hs
    out["hs_S"] = np.array([structure_factors.hard_sphere_sf(q, *row) for row in hs])
    out["hs_S_q0"] = np.array([structure_factors.hard_sphere_sf(q0, *row) for row in hs[:2]])
    syms = ["H", "C", "Al", "Au", "U", "Si"]
    out["atomic_syms"] = np.array(syms)
    qa = np.linspace(0.0, 6.0, 200)
    out["qa"] = qa
    out["atomic_ff"] = np.array([form_factors.atomic_ff_normalized(qa, s) for s in syms])
    # profiles over a log-ish sweep of offsets and widths
    xs_ = np.concatenate([-np.logspace(-7, 0.3, 60)[::-1], [0.0], np.logspace(-7, 0.3, 60)])
    out["px"] = xs_
    vg = []
    vargs = []
    for hg in (1e-9, 1e-6, 1e-4, 2e-3, 6e-3, 0.05):
        for hl in (1e-9, 1e-6, 1e-4, 1.8e-3, 1.5e-2, 0.1):
            vargs.append([hg, hl])
            vg.append(peak_math.voigt(xs_, hg, hl))
    out["voigt_args"] = np.array(vargs)
    out["voigt"] = np.array(vg)
    hh = np.array([1e-6, 1e-4, 2e-3, 0.03, 0.1])
    out["hwhm"] = hh
    out["gauss"] = np.array([peak_math.gaussian(xs_, h) for h in hh])
    out["lorentz"] = np.array([peak_math.lorentzian(xs_, h) for h in hh])
    np.savez_compressed(os.path.join(OUT, "kernels.npz"), **out)


TABLE_CASES = [
    ("F_cubic", "Fm-3m", dict(a=60.0), 0.0, 1.0, True),
    ("F_cubic", "Fm-3m", dict(a=4.046), 0.0, 5.0, True),
    ("F_cubic", "Fm-3m", dict(a=73.219), 0.0, 0.8, True),
    ("F_cubic", "", dict(a=40.0), 0.0, 0.9, True),
    ("F_cubic", "Fm-3m", dict(a=40.0), 0.0, 0.9, False),
    ("F_cubic", "Fm-3m", dict(a=60.0), 0.35, 1.0, True),
    ("F_cubic", "F23", dict(a=40.0), 0.0, 0.9, True),
    ("P_cubic", "Pm-3m", dict(a=33.3), 0.0, 1.0, True),
    ("I_cubic", "Im-3m", dict(a=35.0), 0.0, 1.0, True),
    ("diamond", "Fd-3m", dict(a=30.0), 0.0, 1.0, True),
    ("hcp", "P6(3)/mmc", dict(a=40.0), 0.0, 1.0, True),
    ("hcp", "P6(3)/mmc", dict(a=120.0), 0.0, 0.2, True),
    ("hexagonal", "P6/mmm", dict(a=40.0, c=50.0), 0.0, 1.0, True),
    ("hexagonal", "P-3m1", dict(a=31.0, c=44.0), 0.0, 0.9, True),
    ("triclinic", "P-1", dict(a=30.0, b=35.0, c=40.0, alpha=80.0, beta=95.0, gamma=100.0), 0.0, 1.0, True),
    ("triclinic", "P1", dict(a=30.0, b=35.0, c=40.0, alpha=80.0, beta=95.0, gamma=100.0), 0.0, 0.8, True),
    ("P_tetragonal", "P4/mmm", dict(a=30.0, c=40.0), 0.0, 1.0, True),
    ("I_tetragonal", "I4/mmm", dict(a=25.0, c=33.0), 0.0, 1.0, True),
    ("P_orthorhombic", "Pmmm", dict(a=20.0, b=27.0, c=35.0), 0.0, 1.0, True),
    ("F_orthorhombic", "Fmmm", dict(a=40.0, b=47.0, c=55.0), 0.0, 0.7, True),
    ("C_orthorhombic", "Cmcm", dict(a=30.0, b=37.0, c=45.0), 0.0, 0.8, True),
    ("P_monoclinic", "P2/m", dict(a=20.0, b=27.0, c=35.0, beta=105.0), 0.0, 1.0, True),
    ("C_monoclinic", "C2/m", dict(a=20.0, b=27.0, c=35.0, beta=99.0), 0.0, 1.0, True),
    ("rhombohedral", "R-3m", dict(a=30.0, alpha=75.0), 0.0, 1.0, True),
    ("rhombohedral", "R-3m", dict(a=30.0, alpha=90.0), 0.0, 1.0, True),
]


def ref_table(lattice, sg, latparams, q_min, q_max, use_sym):
    """The reference's own hkl enumeration + reduction (scattering/__init__.py:240-284), verbatim calls."""
    a1, a2, a3 = xd.lattice_vectors(lattice, **latparams)
    b1, b2, b3 = xd.reciprocal_lattice_vectors(a1, a2, a3)
    g_max = 1.0 / (2 * np.pi / q_max)
    g_min = 1.0 / (2 * np.pi / q_min) if q_min > 0.0 else 0
    n1 = np.ceil(g_max * np.linalg.norm(a1) / np.dot(b1, a1))
    n2 = np.ceil(g_max * np.linalg.norm(a2) / np.dot(b2, a2))
    n3 = np.ceil(g_max * np.linalg.norm(a3) / np.dot(b3, a3))
    hr, kr, lr = np.arange(-1 * n1 + 1, n1), np.arange(-1 * n2 + 1, n2), np.arange(-1 * n3 + 1, n3)
    all_hkl = np.array([(h, k, l) for l in lr for k in kr for h in hr
                        if (g_min < np.linalg.norm(np.dot((h, k, l), (b1, b2, b3))) <= g_max)])
    if use_sym:
        red, mult = symmetries.symmetrize_points(all_hkl, np.array([b1, b2, b3]), sg)
    else:
        red, mult = all_hkl, np.ones(all_hkl.shape[0])
    return int(n1), int(n2), int(n3), all_hkl, red, mult


def dump_tables():
    cases = []
    for lattice, sg, lp, q_min, q_max, use_sym in TABLE_CASES:
        n1, n2, n3, all_hkl, red, mult = ref_table(lattice, sg, lp, q_min, q_max, use_sym)
        cases.append(dict(lattice=lattice, space_group=sg, latparams=lp, q_min=q_min, q_max=q_max,
                          use_symmetry=use_sym, n=[n1, n2, n3], n_all=int(all_hkl.shape[0]),
                          hkl=red.astype(int).tolist(), mult=np.asarray(mult).astype(int).tolist()))
        print("table", lattice, sg, all_hkl.shape[0], "->", red.shape[0])
    with open(os.path.join(OUT, "tables.json"), "w") as f:
        json.dump(cases, f, separators=(",", ":"))


def _sys_cases():
    """(name, System, q) triples: the reference's own test constructions, the BASELINE configs,
    and a spread over structures / forms / lattices / profiles."""
    cases = []
    V = lambda v: {"value": v}  # noqa: E731
    # --- reference tests/test_diffraction.py:10-61
    fcc_al = Population("crystalline", "atomic",
                        {"lattice": "F_cubic", "space_group": "Fm-3m", "q_max": 5.0,
                         "structure_factor_mode": "local", "symbol": "Al"},
                        dict(a=V(4.046), hwhm_g=V(0.002), hwhm_l=V(0.0018)))
    hcp = Population("crystalline", "spherical",
                     {"lattice": "hcp", "space_group": "P6(3)/mmc", "q_max": 0.2,
                      "structure_factor_mode": "radial"},
                     dict(a=V(120.0), hwhm_g=V(0.002), hwhm_l=V(0.002), r=V(40.0)))
    glassy = Population("disordered", "atomic", {"interaction": "hard_spheres", "symbol": "Al"},
                        dict(r_hard=V(4.046 * np.sqrt(2) / 4), v_fraction=V(0.6), I0=V(1.0e5)))
    q_al = np.arange(1.0, 5.0, 0.004)
    cases.append(("ref_fcc_Al", System(fcc_Al=fcc_al.to_dict(), sample_metadata={"source_wavelength": WL}), q_al))
    cases.append(("ref_glassy_Al", System(glassy_Al=glassy.to_dict()), q_al))
    cases.append(("ref_mixed_Al", System(glassy_Al=glassy.to_dict(), fcc_Al=fcc_al.to_dict(),
                                         sample_metadata={"source_wavelength": WL}), q_al))
    cases.append(("ref_hcp_spheres", System(hcp_spheres=hcp.to_dict(), sample_metadata={"source_wavelength": WL}),
                  np.arange(0.02, 0.6, 0.001)))
    # --- reference tests/test_fit_gui.py:14-32 style
    cases.append(("ref_np_rnormal", System(
        nanoparticles=dict(structure="diffuse", form="spherical", settings={"distribution": "r_normal"},
                           parameters={"r": V(40.0), "sigma": V(0.1), "I0": V(1000.0)}),
        noise={"model": "flat", "parameters": {"I0": V(0.1)}}), np.linspace(0.005, 0.6, 560)))
    # --- BASELINE cfg1 / cfg2 single members / cfg3 / cfg4
    cases.append(("cfg1", System(p=dict(structure="diffuse", form="spherical", parameters={"r": V(20.0)})),
                  np.linspace(0.001, 0.5, 2000)))
    cases.append(("cfg2_member", System(
        p=dict(structure="disordered", form="spherical",
               settings={"distribution": "r_normal", "sampling_width": 2.5, "sampling_step": 0.1},
               parameters={"r": V(23.7), "sigma": V(0.17), "r_hard": V(29.1), "v_fraction": V(0.31), "I0": V(55.0)}),
        noise={"model": "flat", "parameters": {"I0": V(0.01)}}), np.linspace(0.01, 0.5, 4096)))
    xt = dict(structure="crystalline", form="spherical",
              settings={"lattice": "F_cubic", "space_group": "Fm-3m", "q_max": 1.0, "profile": "voigt"},
              parameters={"a": V(60.0), "r": V(20.0), "hwhm_g": V(0.002), "hwhm_l": V(0.002)})
    cases.append(("cfg3", System(x=xt, source_wavelength=WL), np.linspace(0.01, 1.0, 8192)))
    xt_rad = json.loads(json.dumps(xt))
    xt_rad["settings"]["structure_factor_mode"] = "radial"
    cases.append(("cfg3_radial", System(x=xt_rad, source_wavelength=WL), np.linspace(0.01, 1.0, 2048)))
    xt4 = json.loads(json.dumps(xt))
    xt4["parameters"].update(I0=V(2.5), a=V(52.3), r=V(14.0), hwhm_g=V(0.0011), hwhm_l=V(0.0031))
    cases.append(("cfg4_member", System(
        x=xt4, gp=dict(structure="diffuse", form="guinier_porod", parameters={"rg": V(7.5), "I0": V(20.0)}),
        noise={"model": "flat", "parameters": {"I0": V(0.02)}}, source_wavelength=WL),
        np.linspace(0.01, 1.0, 8192)))
    # --- spread: profiles, corrections, lattices
    for prof, pars in (("gaussian", {"hwhm": V(0.004)}), ("lorentzian", {"hwhm": V(0.003)}),
                       ("voigt", {"hwhm_g": V(1e-6), "hwhm_l": V(4e-3)}),
                       ("voigt", {"hwhm_g": V(5e-3), "hwhm_l": V(1e-6)})):
        pp = dict(a=V(45.0), r=V(12.0))
        pp.update(pars)
        tag = prof + "_" + "_".join("%g" % v["value"] for v in pars.values())
        cases.append(("fcc_" + tag, System(x=dict(structure="crystalline", form="spherical",
                      settings={"lattice": "F_cubic", "space_group": "Fm-3m", "q_max": 0.8, "profile": prof},
                      parameters=pp), source_wavelength=WL), np.linspace(0.02, 0.8, 1500)))
    cases.append(("fcc_nocorr", System(x=dict(structure="crystalline", form="spherical",
                  settings={"lattice": "F_cubic", "space_group": "Fm-3m", "q_max": 0.8,
                            "polarization_correction": False, "lorentz_correction": False, "q_min": 0.2},
                  parameters=dict(a=V(45.0), r=V(12.0), hwhm_g=V(0.003), hwhm_l=V(0.001))),
                  source_wavelength=1.54), np.linspace(0.02, 0.8, 1500)))
    cases.append(("fcc_qmax_none", System(x=dict(structure="crystalline", form="spherical",
                  settings={"lattice": "F_cubic", "space_group": "Fm-3m", "q_max": 0.0, "use_symmetry": False},
                  parameters=dict(a=V(45.0), r=V(12.0), hwhm_g=V(0.003), hwhm_l=V(0.001))),
                  source_wavelength=WL), np.linspace(0.02, 0.55, 700)))
    lat_cases = [
        ("P_cubic", "Pm-3m", dict(a=V(33.3))), ("I_cubic", "Im-3m", dict(a=V(35.0))),
        ("diamond", "Fd-3m", dict(a=V(30.0))), ("hcp", "P6(3)/mmc", dict(a=V(40.0))),
        ("hexagonal", "P6/mmm", dict(a=V(40.0), c=V(50.0))),
        ("P_tetragonal", "P4/mmm", dict(a=V(30.0), c=V(40.0))),
        ("I_tetragonal", "", dict(a=V(25.0), c=V(33.0))),
        ("P_orthorhombic", "Pmmm", dict(a=V(20.0), b=V(27.0), c=V(35.0))),
        ("F_orthorhombic", "Fmmm", dict(a=V(40.0), b=V(47.0), c=V(55.0))),
        ("C_orthorhombic", "Cmcm", dict(a=V(30.0), b=V(37.0), c=V(45.0))),
        ("I_orthorhombic", "Immm", dict(a=V(30.0), b=V(37.0), c=V(45.0))),
        ("P_monoclinic", "P2/m", dict(a=V(20.0), b=V(27.0), c=V(35.0), beta=V(105.0))),
        ("C_monoclinic", "C2/m", dict(a=V(20.0), b=V(27.0), c=V(35.0), beta=V(99.0))),
        ("rhombohedral", "R-3m", dict(a=V(30.0), alpha=V(75.0))),
        ("triclinic", "P-1", dict(a=V(30.0), b=V(35.0), c=V(40.0), alpha=V(80.0), beta=V(95.0), gamma=V(100.0))),
    ]
    for lat, sg, lp in lat_cases:
        for mode in ("local", "radial"):
            pp = dict(r=V(9.0), hwhm_g=V(0.004), hwhm_l=V(0.002), I0=V(3.0))
            pp.update(lp)
            cases.append(("lat_%s_%s" % (lat, mode), System(x=dict(structure="crystalline", form="spherical",
                          settings={"lattice": lat, "space_group": sg, "q_max": 0.7, "structure_factor_mode": mode},
                          parameters=pp), source_wavelength=WL), np.linspace(0.03, 0.75, 600)))
    # atomic + polyatomic crystals
    cases.append(("atomic_bcc_Fe", System(x=dict(structure="crystalline", form="atomic",
                  settings={"lattice": "I_cubic", "space_group": "Im-3m", "q_max": 6.0, "symbol": "Fe",
                            "structure_factor_mode": "radial"},
                  parameters=dict(a=V(2.87), hwhm_g=V(0.01), hwhm_l=V(0.005))), source_wavelength=WL),
                  np.linspace(1.0, 6.0, 1000)))
    cases.append(("poly_NaCl", System(x=dict(structure="crystalline", form="polyatomic",
                  settings={"lattice": "F_cubic", "space_group": "Fm-3m", "q_max": 5.0, "n_atoms": 2,
                            "symbol_0": "Na", "symbol_1": "Cl"},
                  parameters=dict(a=V(5.64), hwhm_g=V(0.01), hwhm_l=V(0.004),
                                  u_1=V(0.5), v_1=V(0.5), w_1=V(0.5))), source_wavelength=WL),
                  np.linspace(0.8, 5.0, 1000)))
    cases.append(("poly_3_radial", System(x=dict(structure="crystalline", form="polyatomic",
                  settings={"lattice": "P_tetragonal", "space_group": "P4/mmm", "q_max": 4.0, "n_atoms": 3,
                            "symbol_0": "Ti", "symbol_1": "O", "symbol_2": "O", "structure_factor_mode": "radial"},
                  parameters=dict(a=V(4.59), c=V(2.96), hwhm_g=V(0.008), hwhm_l=V(0.006),
                                  u_1=V(0.305), v_1=V(0.305), w_1=V(0.0), u_2=V(-0.305), v_2=V(-0.305), w_2=V(0.0))),
                  source_wavelength=WL), np.linspace(0.8, 4.2, 900)))
    # non-crystalline spread
    cases.append(("diffuse_atomic_C", System(p=dict(structure="diffuse", form="atomic", settings={"symbol": "C"},
                  parameters={"I0": V(7.0)})), np.linspace(0.0, 5.0, 500)))
    cases.append(("disordered_gp", System(p=dict(structure="disordered", form="guinier_porod",
                  parameters={"rg": V(12.0), "D": V(3.2), "r_hard": V(18.0), "v_fraction": V(0.2), "I0": V(4.0)}),
                  noise={"model": "low_q_scatter", "parameters": {"I0": V(50.0), "I0_flat_fraction": V(0.02),
                                                                  "effective_rg": V(150.0), "effective_D": V(3.7)}}),
                  np.linspace(0.004, 0.6, 560)))
    cases.append(("disordered_single_q0", System(p=dict(structure="disordered", form="spherical",
                  parameters={"r": V(22.0), "r_hard": V(25.0), "v_fraction": V(0.3), "I0": V(2.0)})),
                  np.linspace(0.0, 0.5, 300)))
    cases.append(("rnormal_trunc", System(p=dict(structure="diffuse", form="spherical",
                  settings={"distribution": "r_normal"}, parameters={"r": V(15.0), "sigma": V(0.9), "I0": V(10.0)})),
                  np.linspace(0.004, 0.6, 560)))
    cases.append(("rnormal_sigma0", System(p=dict(structure="diffuse", form="spherical",
                  settings={"distribution": "r_normal"}, parameters={"r": V(15.0), "sigma": V(8e-10), "I0": V(10.0)})),
                  np.linspace(0.004, 0.6, 560)))
    return cases


def _fixture_yaml_cases(max_per_kind=12):
    """A spread of the reference's fitted-system fixtures (tests/test_data/dataset_*), by kind."""
    paths = sorted(glob.glob(os.path.join(ref_shim.REFERENCE_ROOT, "tests", "test_data", "dataset_*", "*", "*.yml")))
    seen = {}
    picked = []
    for p in paths:
        with open(p) as f:
            d = yaml.safe_load(f)
        try:
            s = System(**d)
        except Exception:
            continue
        kinds = []
        for nm, pop in s.populations.items():
            k = pop.structure + "/" + pop.form
            if pop.form == "spherical":
                k += "/" + pop.settings.get("distribution", "")
            if pop.structure == "crystalline":
                k += "/" + pop.settings["lattice"] + "/" + pop.settings["space_group"] + "/" + pop.settings["structure_factor_mode"]
            kinds.append(k)
        key = s.noise_model.model + "|" + "+".join(sorted(kinds))
        if seen.get(key, 0) >= max_per_kind:
            continue
        seen[key] = seen.get(key, 0) + 1
        rel = os.path.relpath(p, os.path.join(ref_shim.REFERENCE_ROOT, "tests", "test_data"))
        picked.append(("yml:" + rel, s))
    return picked


def dump_systems():
    meta = []
    arrays = {}
    cases = _sys_cases()
    qfix = np.loadtxt(os.path.join(ref_shim.REFERENCE_ROOT, "tests", "test_data", "solution_saxs",
                                   "spheres", "spheres_0.dat"), skiprows=1)[:, 0]
    arrays["q_fixture"] = qfix
    for name, s in _fixture_yaml_cases():
        cases.append((name, s, None))
    rs = np.random.RandomState(11)
    for i, (name, s, q) in enumerate(cases):
        key = "c%03d" % i
        qkey = "q_fixture" if q is None else key + "_q"
        qq = qfix if q is None else np.asarray(q, dtype=float)
        if q is not None:
            arrays[qkey] = qq
        with np.errstate(all="ignore"):
            I = s.compute_intensity(qq)
            arrays[key + "_I"] = I
            pops = {}
            for pn, pop in s.populations.items():
                pops[pn] = pop.compute_intensity(qq, s.sample_metadata["source_wavelength"])
                arrays[key + "_pop_" + pn] = pops[pn]
            arrays[key + "_noise"] = s.noise_model.compute_intensity(qq)
            # residuals against a perturbed copy of the pattern, all four weighting modes
            Iobs = I * (1 + 0.02 * rs.randn(len(qq)))
            arrays[key + "_Iobs"] = Iobs
            res = []
            saved = dict(s.fit_report)
            for ew in (True, False):
                for lw in (True, False):
                    s.set_error_weighted(ew)
                    s.set_logI_weighted(lw)
                    res.append(s.evaluate_residual(qq, Iobs, None, 1.01 * I + 1e-4))
            s.set_error_weighted(True)
            s.set_logI_weighted(True)
            s.set_q_range(qq[len(qq) // 8], qq[-len(qq) // 8])
            res.append(s.evaluate_residual(qq, Iobs, 0.1 * np.sqrt(np.abs(Iobs)) + 1e-3, 1.01 * I + 1e-4))
            s.fit_report.update(saved)
            arrays[key + "_res"] = np.array(res, dtype=float)
        meta.append(dict(key=key, name=name, q=qkey, system=jsonable(s.to_dict())))
        print("system", key, name, len(qq), float(np.nansum(I)))
    with open(os.path.join(OUT, "systems.json"), "w") as f:
        json.dump(meta, f, separators=(",", ":"))
    np.savez_compressed(os.path.join(OUT, "systems.npz"), **arrays)


if __name__ == "__main__" and "--all-fixtures" not in sys.argv and "--features" not in sys.argv:
    os.makedirs(OUT, exist_ok=True)
    dump_definitions()
    dump_kernels()
    dump_tables()
    dump_systems()
    for fn in sorted(os.listdir(OUT)):
        print(fn, os.path.getsize(os.path.join(OUT, fn)))


def dump_all_fixtures():
    """tests/golden/fixtures_all.{json.gz,npz}: every fitted System of the reference's
    tests/test_data/dataset_*/ (753 files) with the reference's I(q) on the 560-point fixture grid.
    Run with `python oracle/make_golden.py --all-fixtures` (about 25 s)."""
    import gzip
    paths = sorted(glob.glob(os.path.join(ref_shim.REFERENCE_ROOT, "tests", "test_data", "dataset_*", "*", "*.yml")))
    q = np.loadtxt(os.path.join(ref_shim.REFERENCE_ROOT, "tests", "test_data", "solution_saxs", "spheres", "spheres_0.dat"),
                   skiprows=1)[:, 0]
    meta, arrs = [], {"q": q}
    for i, p in enumerate(paths):
        with open(p) as f:
            s = System(**yaml.safe_load(f))
        with np.errstate(all="ignore"):
            arrs["I%04d" % i] = s.compute_intensity(q)
        d = jsonable(s.to_dict())
        d.pop("features", None)
        meta.append(dict(name=os.path.relpath(p, os.path.join(ref_shim.REFERENCE_ROOT, "tests", "test_data")), system=d))
    with gzip.open(os.path.join(OUT, "fixtures_all.json.gz"), "wt") as f:
        json.dump(meta, f, separators=(",", ":"))
    np.savez_compressed(os.path.join(OUT, "fixtures_all.npz"), **arrs)


if __name__ == "__main__" and "--all-fixtures" in sys.argv:
    dump_all_fixtures()


def dump_features():
    """tests/golden/features.npz: the reference's profile_pattern (tools/profiler.py:64) on the seven
    measured patterns of tests/test_data/solution_saxs (stored with their q, I) and on the clean and
    noisy I(q) of every case in systems.npz.  Run with `python oracle/make_golden.py --features`."""
    import warnings
    from xrsdkit.tools import profiler
    warnings.simplefilter("ignore")
    arrs = {}
    files = sorted(glob.glob(os.path.join(ref_shim.REFERENCE_ROOT, "tests", "test_data", "solution_saxs", "*", "*.dat")))
    for i, p in enumerate(files):
        a = np.loadtxt(p)
        arrs["meas%d_q" % i], arrs["meas%d_I" % i] = a[:, 0], a[:, 1]
        arrs["meas%d_f" % i] = np.array(list(profiler.profile_pattern(a[:, 0], a[:, 1]).values()), dtype=float)
    Z = np.load(os.path.join(OUT, "systems.npz"))
    with open(os.path.join(OUT, "systems.json")) as f:
        meta = json.load(f)
    n = 0
    for m in meta:
        q = Z[m["q"]]
        for tag in ("I", "Iobs"):
            I = Z[m["key"] + "_" + tag]
            if not np.all(np.isfinite(I)) or np.sum(I > 0) < 8:
                continue
            with np.errstate(all="ignore"):
                try:
                    f = np.array(list(profiler.profile_pattern(q, I).values()), dtype=float)
                except Exception as e:           # degenerate patterns the reference itself cannot profile
                    print("skip", m["key"], tag, type(e).__name__)
                    continue
            arrs[m["key"] + "_f" + tag] = f
            n += 1
    np.savez_compressed(os.path.join(OUT, "features.npz"), **arrs)
    print("features:", len(files), "measured +", n, "computed patterns")


if __name__ == "__main__" and "--features" in sys.argv:
    dump_features()
