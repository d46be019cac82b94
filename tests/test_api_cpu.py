"""Host-side logic that needs no GPU: the System / Population / NoiseModel object model against the
reference's own dict output, lowering to the flat descriptors, sharding arithmetic."""
import copy

import numpy as np
import pytest

from xrsdkit_b200 import abi, batch, lowering
from xrsdkit_b200.system import System, Population, NoiseModel, unflatten_params


def test_dict_round_trip_equals_reference(golden_systems):
    """System(**d).to_dict() reproduces the reference's fully-populated dict (defaults, key sets, values)."""
    meta, _ = golden_systems
    for m in meta:
        d = m["system"]
        s = System(**copy.deepcopy(d))
        out = s.to_dict()
        assert set(out.keys()) == set(d.keys()), m["name"]
        for k in d:
            assert out[k] == d[k], (m["name"], k)
        assert System.from_dict(out).to_dict() == out
        assert s.clone().to_dict() == out


def test_defaults_from_minimal_construction(golden_systems):
    """Building from the user-level arguments only gives the same populated dict as the reference did."""
    meta, _ = golden_systems
    by = {m["name"]: m for m in meta}
    s = System(p=dict(structure="diffuse", form="spherical", parameters={"r": {"value": 20.0}}))
    assert s.to_dict() == by["cfg1"]["system"]
    x = dict(structure="crystalline", form="spherical",
             settings={"lattice": "F_cubic", "space_group": "Fm-3m", "q_max": 1.0, "profile": "voigt"},
             parameters={"a": {"value": 60.0}, "r": {"value": 20.0}, "hwhm_g": {"value": 0.002}, "hwhm_l": {"value": 0.002}})
    assert System(x=x, source_wavelength=0.8265617).to_dict() == by["cfg3"]["system"]


def test_population_settings_and_parameter_rules():
    p = Population("crystalline", "spherical", {"lattice": "hexagonal", "profile": "gaussian"})
    assert list(p.parameters.keys()) == ["I0", "a", "r", "c", "hwhm"]      # same order as the reference
    p.update_settings({"profile": "voigt", "lattice": "F_cubic"})
    assert list(p.parameters.keys()) == ["I0", "a", "r", "hwhm_g", "hwhm_l"]
    assert p.settings["q_max"] == 1.0 and p.settings["structure_factor_mode"] == "local"
    with pytest.raises(ValueError):
        p.update_parameters({"sigma": {"value": 0.1}})
    with pytest.raises(ValueError):
        p.update_settings({"distribution": "r_normal"})
    with pytest.raises(ValueError):
        Population("diffuse", "polyatomic")
    q = Population("diffuse", "spherical", {"distribution": "r_normal"}, {"sigma": {"value": 0.2}})
    assert q.settings["sampling_width"] == 3.5 and q.settings["sampling_step"] == 0.05
    assert q.parameters["sigma"] == {"value": 0.2, "fixed": False, "bounds": [0., 2.], "constraint_expr": None}
    n = NoiseModel("low_q_scatter")
    assert list(n.parameters.keys()) == ["I0", "I0_flat_fraction", "effective_rg", "effective_D"]
    with pytest.raises(ValueError):
        n.update_parameters({"rg": {"value": 1.0}})
    n.set_model("flat")
    assert list(n.parameters.keys()) == ["I0"]


def test_system_parameter_plumbing():
    s = System(a=dict(structure="diffuse", form="guinier_porod"), b=dict(structure="disordered", form="spherical"))
    flat = s.flatten_params()
    assert list(flat.keys()) == ["noise__I0", "a__I0", "a__rg", "a__D", "b__I0", "b__r_hard", "b__v_fraction", "b__r"]
    flat["a__rg"]["value"] = 3.0                     # shared by reference, like xrsdkit
    assert s.populations["a"].parameters["rg"]["value"] == 3.0
    s.update_params_from_dict(unflatten_params({"b__r": {"value": 7.0}, "noise__I0": {"value": 0.5}}))
    assert s.populations["b"].parameters["r"]["value"] == 7.0 and s.noise_model.parameters["I0"]["value"] == 0.5
    s.set_q_range(0.1, 0.4)
    s.set_error_weighted(0)
    assert s.fit_report["q_range"] == [0.1, 0.4] and s.fit_report["error_weighted"] is False
    s.remove_population("a")
    s.add_population("c", "diffuse", "atomic", {"symbol": "Au"})
    assert list(s.populations.keys()) == ["b", "c"]
    assert System(source_wavelength=0.7).sample_metadata["source_wavelength"] == 0.7


def test_lowering_slots_and_descriptors():
    s = System(
        x=dict(structure="crystalline", form="polyatomic",
               settings={"lattice": "F_cubic", "space_group": "Fm-3m", "q_max": 0.0, "n_atoms": 2,
                         "symbol_0": "Na", "symbol_1": "Cl", "structure_factor_mode": "radial"},
               parameters={"a": {"value": 5.64}, "u_1": {"value": 0.5}}),
        g=dict(structure="disordered", form="spherical", settings={"distribution": "r_normal"}),
        noise={"model": "low_q_scatter"}, source_wavelength=0.8265617)
    q = np.linspace(0.0, 4.0, 50)
    lay = lowering.lower_system(s, q)
    assert lay.names == list(s.flatten_params().keys())
    assert lay.c.n_pops == 3 and lay.c.n_xtal == 1 and lay.c.q0_is_zero == 1
    nz, x, g = lay.c.pops[0], lay.c.pops[1], lay.c.pops[2]
    assert nz.kind == abi.POP_NOISE_LOW_Q and nz.p_rg == lay.slot("noise__effective_rg")
    assert x.kind == abi.POP_CRYSTALLINE and x.n_species == 2 and x.n_sites == 4 and x.n_ops == 9 and x.sf_radial == 1
    assert x.q_max == 4.0                                   # falsy q_max -> q[-1]
    assert x.atom_Z[0] == 11 and x.atom_Z[1] == 17 and x.p_uvw[1][0] == lay.slot("x__u_1")
    assert x.p_lat[0] == lay.slot("x__a") and x.p_lat[1] == -1
    assert g.kind == abi.POP_NONCRYSTALLINE and g.disordered == 1 and g.form == abi.FORM_SPHERE_R_NORMAL
    assert g.sampling_width == 3.5 and g.p_sigma == lay.slot("g__sigma")
    # lmfit-style free set: not fixed, no constraint expression
    assert lay.slot("x__occupancy_0") not in lay.free_slots() and lay.slot("x__a") in lay.free_slots()
    lo, hi = lay.bounds([lay.slot("g__v_fraction"), lay.slot("x__a")])
    assert lo.tolist() == [0.01, 0.1] and hi[0] == 0.7405 and np.isinf(hi[1])
    # candidate box bound: a = 5.64, q_max = 4 -> n = ceil(4*5.64/2pi)+1 = 5 -> 9^3
    assert lowering.max_cells(lay, lay.values[None, :]) == 9 ** 3


def test_lowering_errors():
    with pytest.raises(ValueError):      # falsy wavelength with a crystalline population
        lowering.lower_system(System(x=dict(structure="crystalline", form="spherical", parameters={"a": {"value": 9.}}),
                                     source_wavelength=0), np.array([0.1]))
    with pytest.raises(ValueError):      # space group of another lattice
        lowering.lower_system(System(x=dict(structure="crystalline", form="spherical",
                                            settings={"lattice": "hcp", "space_group": "Fm-3m"})), np.array([0.1]))
    with pytest.raises(KeyError):
        lowering.lower_system(System(x=dict(structure="diffuse", form="atomic", settings={"symbol": "Qq"})), np.array([0.1]))
    # fixed capacities of the engine (include/xrsd_b200.h): a ValueError that names the limit
    with pytest.raises(ValueError, match="XRSD_MAX_SPECIES"):
        lowering.lower_system(System(x=dict(structure="crystalline", form="polyatomic", settings={"n_atoms": 5})),
                              np.array([0.1]))
    many = dict(("p%d" % i, dict(structure="diffuse", form="spherical")) for i in range(6))
    with pytest.raises(ValueError, match="XRSD_MAX_POPS"):
        lowering.lower_system(System(**many), np.array([0.1]))
    poly = dict(("x%d" % i, dict(structure="crystalline", form="polyatomic", settings={"n_atoms": 4, "lattice": "triclinic"}))
                for i in range(4))
    with pytest.raises(ValueError, match="XRSD_MAX_PARAMS"):
        lowering.lower_system(System(**poly), np.array([0.1]))
    assert issubclass(lowering.CapacityError, NotImplementedError)


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 10000, 1000003):
        for world in (1, 2, 3, 8):
            spans = [batch.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_compute_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        System(p=dict(structure="diffuse", form="spherical")).compute_intensity(np.linspace(0.01, 0.5, 10))


def test_yaml_round_trip_and_grouping(tmp_path, golden_systems):
    """Files in the reference's YAML format load unchanged; systems are grouped by layout for batching."""
    import yaml
    from xrsdkit_b200.tools import ymltools
    meta, arr = golden_systems
    picks = [m for m in meta if m["name"].startswith("yml:")][:40]
    paths = []
    for i, m in enumerate(picks):
        p = tmp_path / ("s%03d.yml" % i)
        with open(p, "w") as f:
            yaml.dump(m["system"], f)               # what xrsdkit's save_sys_to_yaml writes
        paths.append(str(p))
    q = arr["q_fixture"]
    systems, groups = ymltools.load_batches(paths, q)
    assert [s.to_dict() for s in systems] == [m["system"] for m in picks]
    assert sorted(i for _, _, idx in groups for i in idx) == list(range(len(picks)))
    assert len(groups) < len(picks)                 # fitted systems of one experiment share layouts
    for lay, params, idx in groups:
        assert params.shape == (len(idx), lay.n_params)
        for row, i in zip(params, idx):
            vals = [rec["value"] for rec in systems[i].flatten_params().values()]
            np.testing.assert_array_equal(row, np.array(vals, dtype=float))
    out = tmp_path / "again.yml"
    ymltools.save_sys_to_yaml(str(out), systems[0])
    assert ymltools.load_sys_from_yaml(str(out)).to_dict() == systems[0].to_dict()


def test_affine_constraint_expressions_are_parsed_and_others_rejected():
    """constraint_expr (reference system/__init__.py:197-209 hands it to lmfit): expressions affine in one other
    parameter lower to (source slot, scale, offset); everything else raises NotImplementedError."""
    from xrsdkit_b200 import lowering
    names = ["noise__I0", "particle__r", "shell__r", "shell__a"]
    ok = {"particle__r": (1, 1.0, 0.0), " particle__r ": (1, 1.0, 0.0), "2*particle__r": (1, 2.0, 0.0),
          "particle__r*2 + 5": (1, 2.0, 5.0), "(shell__a - 3) / 2": (3, 0.5, -1.5), "-particle__r + 1e1": (1, -1.0, 10.0),
          "particle__r + particle__r": (1, 2.0, 0.0)}
    for expr, want in ok.items():
        got = lowering.parse_affine_constraint(expr, names)
        assert got[0] == want[0] and got[1] == pytest.approx(want[1]) and got[2] == pytest.approx(want[2]), expr
    for expr in ("particle__r**2", "particle__r*shell__r", "1/particle__r", "sqrt(particle__r)", "unknown__r", "3.0",
                 "particle__r +", "particle__r if 1 else 2"):
        with pytest.raises(NotImplementedError):
            lowering.parse_affine_constraint(expr, names)


def test_linear_combination_constraints_lower_to_ordered_terms():
    """constraint_expr linear in SEVERAL parameters: parse_linear_constraint gives exact coefficients, and
    SystemLayout.ties() lowers them to the ordered terms of xrsd_jacobian_batch (first term sets the slot, further
    terms carry dst = -(slot + 1) and accumulate); _apply_ties reproduces the expression."""
    from xrsdkit_b200 import fitting, lowering
    from xrsdkit_b200.system import System
    names = ["noise__I0", "core__r", "shell__t", "shell__a"]
    terms, off = lowering.parse_linear_constraint("0.5*(core__r + shell__t) - shell__a/4 + 2", names)
    assert dict(terms) == {1: 0.5, 2: 0.5, 3: -0.25} and off == 2.0
    assert lowering.parse_linear_constraint("3.0", names) == ([], 3.0)
    assert lowering.parse_linear_constraint("core__r - core__r + 1", names) == ([], 1.0)
    for expr in ("core__r*shell__t", "core__r/shell__t", "core__r**2", "abs(core__r)", "core__r/0"):
        with pytest.raises(NotImplementedError):
            lowering.parse_linear_constraint(expr, names)
    s = System(a=dict(structure="diffuse", form="spherical", parameters={"r": {"value": 10.0}}),
               b=dict(structure="diffuse", form="spherical", parameters={"r": {"value": 20.0}}),
               c=dict(structure="diffuse", form="spherical",
                      parameters={"r": {"value": 1.0, "constraint_expr": "0.5*a__r + 0.25*b__r + 3"},
                                  "I0": {"value": 1.0, "constraint_expr": "7.5"}}))
    lay = lowering.lower_system(s, np.linspace(0.01, 0.5, 16))
    ties = lay.ties()
    cr, ci = lay.slot("c__r"), lay.slot("c__I0")
    assert [lowering.tie_slot(t[0]) for t in ties] == [ci, cr, cr] or [lowering.tie_slot(t[0]) for t in ties] == [cr, cr, ci]
    assert sum(1 for t in ties if t[0] < 0) == 1
    P = np.tile(lay.values, (3, 1))
    P[:, lay.slot("a__r")] = [10.0, 12.0, 14.0]
    fitting._apply_ties(P, ties)
    np.testing.assert_allclose(P[:, cr], 0.5 * np.array([10.0, 12.0, 14.0]) + 0.25 * 20.0 + 3.0, rtol=1e-15)
    np.testing.assert_array_equal(P[:, ci], 7.5)
    assert cr not in lay.free_slots() and ci not in lay.free_slots()
    s.populations["a"].parameters["r"]["constraint_expr"] = "c__r"         # chained: a constrained parameter as a source
    with pytest.raises(NotImplementedError):
        lowering.lower_system(s, np.linspace(0.01, 0.5, 16)).ties()


def test_nonlinear_constraints_compile_to_expression_programs():
    """Any arithmetic constraint_expr (lmfit hands the string to asteval, reference system/__init__.py:197-209):
    compile_constraint gives a postfix program whose host interpreter (_apply_ties, the twin of the device one)
    reproduces Python's own evaluation of the expression, on numpy arrays and on torch tensors."""
    import math
    import torch
    from xrsdkit_b200 import abi, fitting, lowering
    from xrsdkit_b200.system import System
    names = ["noise__I0", "a__r", "b__r", "c__r"]
    rs = np.random.RandomState(3)
    vals = rs.uniform(0.5, 3.0, (5, 4))
    env = dict(sqrt=np.sqrt, exp=np.exp, log=np.log, log10=np.log10, sin=np.sin, cos=np.cos, tan=np.tan, arcsin=np.arcsin,
               arccos=np.arccos, arctan=np.arctan, abs=np.abs, min=np.minimum, max=np.maximum, pi=np.pi, e=np.e)
    for expr in ("a__r*b__r", "a__r/b__r + 1", "a__r**2 - 3*b__r", "sqrt(a__r*b__r)/2", "exp(-a__r) + log(b__r)*log10(b__r)",
                 "sin(a__r)*cos(b__r) - tan(a__r/4)", "arcsin(a__r/4) + arccos(b__r/4) + arctan(a__r)", "abs(a__r - b__r)",
                 "min(a__r, b__r)*max(a__r, 2)", "2*pi*a__r/(b__r + e)", "-(a__r + b__r)**0.5", "(a__r + 1)*(b__r + 2)*(a__r + 3)*(b__r + 4)",
                 "sqrt(2)*a__r*b__r"):
        words = lowering.compile_constraint(expr, names)
        ties = words + [(3, -1, 1.0, 0.0)]
        want = eval(expr, dict(env, **{n: vals[:, i] for i, n in enumerate(names)}))
        P = vals.copy()
        fitting._apply_ties(P, ties)
        np.testing.assert_allclose(P[:, 3], want, rtol=1e-14, err_msg=expr)
        T = torch.tensor(vals)
        fitting._apply_ties(T, ties)
        np.testing.assert_allclose(T[:, 3].numpy(), want, rtol=1e-14, err_msg=expr)
    assert lowering.compile_constraint("sqrt(2)*3", names) == [(-abi.TIE_WORD, 0, math.sqrt(2) * 3, 0.0)]       # constants fold
    for expr in ("a__r if a__r > 1 else 2", "a__r.real", "unknown__r*a__r", "a__r +", "a__r % 2", "erf(a__r)"):
        with pytest.raises(NotImplementedError):
            lowering.compile_constraint(expr, names)
    deep = "a__r*(b__r + a__r*(b__r + a__r*(b__r + a__r*(b__r + a__r*(b__r + a__r*(b__r + a__r*(b__r + a__r*(b__r + a__r))))))))"
    with pytest.raises(ValueError):
        lowering.compile_constraint(deep, names)
    # through the layout: a non-linear expression becomes program words closed by a store term; linear ones stay linear
    s = System(a=dict(structure="diffuse", form="spherical", parameters={"r": {"value": 10.0}}),
               b=dict(structure="diffuse", form="spherical", parameters={"r": {"value": 20.0}}),
               c=dict(structure="diffuse", form="spherical",
                      parameters={"r": {"value": 1.0, "constraint_expr": "sqrt(a__r*b__r)"},
                                  "I0": {"value": 1.0, "constraint_expr": "2*a__r"}}))
    lay = lowering.lower_system(s, np.linspace(0.01, 0.5, 16))
    ties = lay.ties()
    assert lowering.tied_slots(ties) == {lay.slot("c__r"), lay.slot("c__I0")}
    assert sum(1 for t in ties if t[0] <= -abi.TIE_WORD) == 4 and sum(1 for t in ties if t[1] == -1) == 1
    P = np.tile(lay.values, (2, 1))
    fitting._apply_ties(P, ties)
    np.testing.assert_allclose(P[:, lay.slot("c__r")], np.sqrt(200.0), rtol=1e-15)
    np.testing.assert_allclose(P[:, lay.slot("c__I0")], 20.0, rtol=1e-15)
    s.populations["c"].parameters["r"]["constraint_expr"] = "c__r*a__r"
    with pytest.raises(NotImplementedError):
        lowering.lower_system(s, np.linspace(0.01, 0.5, 16)).ties()
