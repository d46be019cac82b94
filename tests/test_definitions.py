"""xrsdkit_b200.definitions against the tables dumped from the unmodified reference."""
import numpy as np

from xrsdkit_b200 import definitions as d


def _norm(o):
    if isinstance(o, dict):
        return {str(k): _norm(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [_norm(v) for v in o]
    if isinstance(o, (np.floating, np.integer)):
        return o.item()
    return o


def test_tables_equal_reference(golden_definitions):
    g = golden_definitions
    for name in ("structure_names", "form_factor_names", "noise_model_names", "all_lattices", "bravais_lattices",
                 "crystal_point_groups", "all_point_groups"):
        assert _norm(getattr(d, name)) == g[name], name
    assert _norm(d.lattice_space_groups) == g["lattice_space_groups"]
    assert _norm(d.sg_point_groups) == g["sg_point_groups"]
    assert _norm(d.atomic_params) == g["atomic_params"]
    assert _norm(d.structure_settings) == g["structure_settings"]
    # (definitions.json is written with sorted keys; key ORDER is checked through System.to_dict() in test_api_cpu)
    assert _norm(d.form_settings) == g["form_settings"]
    assert _norm(d.form_factor_params) == g["form_factor_params"]
    assert _norm(d.noise_params) == g["noise_params"]


def test_parameter_generators_equal_reference(golden_definitions):
    g = golden_definitions
    for key, ref in g["structure_params_crystalline"].items():
        lat, prof = key.split("|")
        mine = d.structure_params("crystalline", {"lattice": lat, "profile": prof})
        assert _norm(mine) == ref, key
    assert _norm(d.structure_params("disordered", {"interaction": "hard_spheres"})) == g["structure_params_disordered"]
    assert _norm(d.additional_form_factor_params("polyatomic", {"n_atoms": 3})) == g["additional_polyatomic_3"]
    assert _norm(d.additional_form_factor_params("spherical", {"distribution": "r_normal"})) == g["additional_r_normal"]


def test_lattice_geometry_equal_reference(golden_definitions):
    probe = golden_definitions["lattice_geometry_probe"]
    lp = probe["latparams"]
    for lat, ref in probe["vectors"].items():
        a1, a2, a3 = d.lattice_vectors(lat, **lp)
        np.testing.assert_allclose(np.array([a1, a2, a3], dtype=float), np.array(ref["a"]), rtol=1e-15, atol=1e-15)
        b = d.reciprocal_lattice_vectors(a1, a2, a3)
        np.testing.assert_allclose(np.array(b), np.array(ref["b"]), rtol=1e-14, atol=1e-18)
        np.testing.assert_array_equal(d.lattice_coords(lat), np.array(ref["sites"]))


def test_validate_rules():
    import pytest
    for st in ("diffuse", "disordered"):
        with pytest.raises(ValueError):
            d.validate(st, "polyatomic", {})
    with pytest.raises(ValueError):
        d.validate("crystalline", "guinier_porod", {})
    with pytest.raises(ValueError):
        d.validate("crystalline", "spherical", {"distribution": "r_normal"})
    d.validate("crystalline", "spherical", {"distribution": "single"})
