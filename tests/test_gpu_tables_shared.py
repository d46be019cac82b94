"""GPU tests of the shared reflection tables (xrsd_tables_t.table_of) and of candidate boxes beyond shared memory.

For a cubic cell the reduced hkl list, multiplicities, shells and geometric weights depend on the cell edge only
through WHICH integer shells pass the reference's G_min < |G| <= G_max test (scattering/__init__.py:276,
symmetries.py:44-92), so the builder makes each distinct table once.  These tests pin that claim bit for bit
against the per-system build, and the large-cell path against the shared-memory path and the oracle."""
import numpy as np
import pytest

from helpers import TOL_F64, rel_err, cfg3_system, cfg3_params, params_to_sysdict

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from xrsdkit_b200 import engine
    return engine.get_engine()


def _same_tables(A, B):
    """Per-table contents of two TableSets, compared bit for bit on the used part."""
    import torch
    assert torch.equal(A.n_all, B.n_all) and torch.equal(A.n_refl, B.n_refl) and torch.equal(A.n_shell, B.n_shell)
    assert torch.equal(A.status, B.status)
    ns = A.n_shell.long()
    cap = int(ns.max().item())
    used = torch.arange(cap, device=ns.device)[None, :] < ns[:, None]
    ha, hb = A.shell_hkl[:, :cap], B.shell_hkl[:, :cap]
    ga, gb = A.shell_geo[:, :cap], B.shell_geo[:, :cap]
    assert torch.equal(ha[used], hb[used])
    assert torch.equal(ga[used].view(torch.int64), gb[used].view(torch.int64))      # bit patterns, not values


def test_shared_tables_equal_own_tables_bit_for_bit(eng):
    """10^4 random cell edges (BASELINE cfg3: a ~ U(40, 80)): the shared build gives every system exactly the table
    its own build gives it, with ~10^2 distinct tables behind them, and I(q) is bit-identical."""
    import torch
    from xrsdkit_b200 import batch
    from xrsdkit_b200.system import System
    s = cfg3_system(System)
    q = np.linspace(0.01, 1.0, 1024)
    lay = batch.layout_of(s, q)
    B = 10000
    P = cfg3_params(lay, B, seed=77)
    d_p = eng.to_device(P)
    own = eng.build_tables(lay, d_p, B, params_host=P, share=False)
    shared = eng.build_tables(lay, d_p, B, params_host=P, share=True)
    assert own.store.n_shared == 0 and own.shared_fraction() == 0.0
    assert shared.shared_fraction() > 0.999
    rows = torch.unique(shared.table_of)
    assert 50 <= rows.numel() <= 200, rows.numel()
    _same_tables(own, shared)
    Ia = eng.intensity(lay, q, d_p, tables=own)
    Ib = eng.intensity(lay, q, d_p, tables=shared)
    assert torch.equal(Ia.view(torch.int64), Ib.view(torch.int64))
    # a second build into the same store finds every row in place (nothing is flagged for building) ...
    again = eng.build_tables(lay, d_p, B, max_cells=None, params_host=P, reuse=shared)
    assert again is shared and int(shared.build_flag.sum().item()) == 0
    # ... and after reset_shared() rebuilds them to the same bits
    before = [t.clone() for t in (shared.n_shell, shared.shell_hkl, shared.shell_geo)]
    shared.store.reset_shared()
    shared.store.shell_geo.zero_()
    eng.build_tables(lay, d_p, B, params_host=P, reuse=shared)
    assert int(shared.build_flag.sum().item()) == rows.numel()
    for a, b in zip(before, (shared.n_shell, shared.shell_hkl, shared.shell_geo)):
        assert torch.equal(a, b)


def test_shared_tables_with_q_min_and_masks(eng):
    """q_min > 0 keys the tables by both limits; systems outside `active` keep their row index."""
    import torch
    from xrsdkit_b200 import batch
    from xrsdkit_b200.system import System
    s = cfg3_system(System)
    s.populations["x"].update_settings({"q_min": 0.35, "q_max": 0.9})
    q = np.linspace(0.05, 1.0, 512)
    lay = batch.layout_of(s, q)
    B = 600
    P = cfg3_params(lay, B, seed=78)
    d_p = eng.to_device(P)
    own = eng.build_tables(lay, d_p, B, params_host=P, share=False)
    shared = eng.build_tables(lay, d_p, B, params_host=P)
    _same_tables(own, shared)
    assert shared.shared_fraction() > 0.5            # same upper key but another lower key keeps its own row
    # masked rebuild with other parameters: inactive systems keep pointing at their old rows
    P2 = P.copy()
    P2[:, lay.slot("x__a")] *= 1.07
    active = torch.zeros(B, dtype=torch.uint8, device=eng.device)
    active[::2] = 1
    before = shared.table_of.clone()
    eng.build_tables(lay, eng.to_device(P2), B, params_host=P2, reuse=shared, active=active)
    own2 = eng.build_tables(lay, eng.to_device(P2), B, params_host=P2, share=False)
    assert torch.equal(shared.table_of[1::2], before[1::2])
    assert torch.equal(shared.n_shell[::2], own2.n_shell[::2]) and torch.equal(shared.n_shell[1::2], own.n_shell[1::2])


@pytest.mark.parametrize("lattice,sg,share", [("F_cubic", "Fm-3m", True), ("P_cubic", "Pm-3m", False),
                                              ("hexagonal", "P6/mmm", False), ("P_orthorhombic", "Pmmm", False)])
def test_candidate_box_in_global_memory_equals_shared_memory(eng, monkeypatch, lattice, sg, share):
    """The builder with its candidate grid in a global-memory region (boxes beyond the shared-memory budget) against
    the same build in shared memory: identical reduced lists, shells and weights."""
    import torch
    from xrsdkit_b200 import engine, lowering
    from xrsdkit_b200.system import System
    from oracle import xrsd_oracle as orc
    rs = np.random.RandomState(11)
    B = 40
    lp0 = dict(a=30.0, b=33.0, c=37.0)
    s = System(x=dict(structure="crystalline", form="spherical",
                      settings={"lattice": lattice, "space_group": sg, "q_max": 0.9},
                      parameters=dict((k, {"value": lp0[k]}) for k in orc.lattice_param_names(lattice) if k in lp0)),
               source_wavelength=0.8265617)
    lay = lowering.lower_system(s, np.array([0.1, 1.0]))
    P = np.tile(lay.values, (B, 1))
    for k in orc.lattice_param_names(lattice):
        if k in lp0:
            P[:, lay.slot("x__" + k)] = lp0[k] * rs.uniform(0.8, 1.25, B)
    d_p = eng.to_device(P)
    ref = eng.build_tables(lay, d_p, B, params_host=P, share=share, export_reflections=True)
    monkeypatch.setattr(engine, "SMEM_CELLS", 500)            # every box of this batch is "too large" now
    monkeypatch.setattr(engine, "GRID_SCRATCH_BYTES", 7 * 2 * lowering.max_cells(lay, P))     # 7 regions: several rounds
    big = eng.build_tables(lay, d_p, B, params_host=P, share=share, export_reflections=True)
    assert big.store.grid is not None and big.store.grid_slots == 7
    _same_tables(ref, big)
    n = ref.n_refl.long()
    used = torch.arange(int(n.max().item()), device=n.device)[None, :] < n[:, None]
    assert torch.equal(ref.refl_hkl[:, :used.shape[1]][used], big.refl_hkl[:, :used.shape[1]][used])


def test_large_cells_at_default_settings_against_golden(eng):
    """F_cubic with the reference's default q_max = 1.0 (definitions.py:210-211) and cell edges of 150 and 206 A (the
    largest in its fixture corpus): candidate boxes of 47^3 and 65^3 cells -- beyond shared memory -- and 56 k / 140 k
    reflections.  Golden reduced lists, multiplicities and I(q): tests/golden/large_cell.npz, from the oracle's chunked
    restatement of symmetrize_points (oracle/make_golden_large.py; the reference itself needs an N x 3 x N tensor,
    78 / 470 GB here).  Reflection lists bit-exact, I(q) <= 1e-10."""
    import os
    from conftest import GOLDEN
    from xrsdkit_b200 import lowering
    from xrsdkit_b200.system import System
    path = os.path.join(GOLDEN, "large_cell.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/large_cell.npz not generated")
    g = np.load(path)
    q = g["q"]
    for a in (150, 206):
        s = System(x=dict(structure="crystalline", form="spherical",
                          settings={"lattice": "F_cubic", "space_group": "Fm-3m", "profile": "voigt"},
                          parameters={"a": {"value": float(a)}, "r": {"value": 0.3 * a}, "hwhm_g": {"value": 2e-3},
                                      "hwhm_l": {"value": 2e-3}}), source_wavelength=0.8265617)
        assert s.populations["x"].settings["q_max"] == 1.0
        lay = lowering.lower_system(s, q)
        T = eng.build_tables(lay, eng.to_device(lay.values[None, :]), 1, params_host=lay.values[None, :], export_reflections=8192)
        assert T.store.grid is not None                                    # the box did not fit in shared memory
        n = int(T.n_refl[0].item())
        got = T.refl_hkl[0, :n].cpu().numpy()
        assert int(T.n_all[0].item()) == int(g["mult_%d" % a].sum())
        assert got[:, :3].tolist() == g["hkl_%d" % a].tolist()
        assert got[:, 3].tolist() == g["mult_%d" % a].tolist()
        err, same = rel_err(s.compute_intensity(q), g["I_%d" % a])
        assert same and err <= TOL_F64, (a, err)
