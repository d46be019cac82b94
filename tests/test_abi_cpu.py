"""The C-ABI library loads on a CPU-only box and exports every symbol include/xrsd_b200.h declares."""
import ctypes
import os
import re

import numpy as np

from xrsdkit_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "xrsd_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(xrsd_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = abi.load()
    names = _declared()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(abi.EXPORTED_SYMBOLS)


def test_struct_mirror_matches_library():
    lib = abi.load()
    sizes = [ctypes.c_int32() for _ in range(4)]
    lib.xrsd_struct_sizes(*[ctypes.byref(s) for s in sizes])
    assert [s.value for s in sizes] == [ctypes.sizeof(abi.Pop), ctypes.sizeof(abi.Layout), ctypes.sizeof(abi.Tables),
                                        ctypes.sizeof(abi.Objective)]
    assert lib.xrsd_abi_version() == abi.ABI_VERSION == 5
    assert lib.xrsd_table_smem_bytes(10, 10, 10) > 2 * 19 ** 3


def test_argument_errors_without_touching_the_gpu():
    lib = abi.load()
    L = abi.Layout()
    L.n_pops = 0
    rc = lib.xrsd_intensity_batch(ctypes.byref(L), None, 10, None, 1, 1, None, None, 10, None)
    assert rc == abi.EINVAL and b"n_pops" in lib.xrsd_last_error()
    rc = lib.xrsd_peak_profile(7, 1.0, 1.0, None, 0, None, None)
    assert rc == abi.EINVAL
    # entry points of the fit loop: every one of these is refused on the host, before any CUDA call
    assert lib.xrsd_copy_rows_masked(None, 8, None, 8, 8, 4, None, None) == abi.EINVAL          # NULL pointers
    assert lib.xrsd_copy_rows_masked(1 << 20, 4, 1 << 21, 8, 8, 4, 1 << 22, None) == abi.EINVAL    # rows would overlap
    assert b"dst_stride" in lib.xrsd_last_error()
    assert lib.xrsd_copy_rows_masked(None, 0, None, 0, 0, 0, None, None) == abi.OK                 # nothing to do
    assert lib.xrsd_copy_rows_masked(None, 0, None, 0, 8, -1, None, None) == abi.EINVAL
    O = abi.Objective()
    assert lib.xrsd_pattern_chi2_batch(ctypes.byref(O), None, 16, None, 16, 2, None, None, 16, None, None, None) == abi.EINVAL
    assert lib.xrsd_pattern_chi2_batch(ctypes.byref(O), 1 << 20, 16, 1 << 21, 8, 2, 1 << 22, None, 16, 1 << 23, None, None) == abi.EINVAL
    assert b"ld_I" in lib.xrsd_last_error()
    assert lib.xrsd_pattern_chi2_batch(ctypes.byref(O), None, 16, None, 16, 0, None, None, 16, None, None, None) == abi.OK
    from helpers import cfg2_system
    from xrsdkit_b200 import lowering
    from xrsdkit_b200.system import System
    q = np.linspace(0.01, 0.5, 64)
    lay = lowering.lower_system(cfg2_system(System), q)
    free = (ctypes.c_int32 * 2)(lay.slot("p__I0"), lay.slot("p__r"))
    assert lib.xrsd_jacobian_n_variants(ctypes.byref(lay.c), free, 2, None, 0) == 2                # base + r; I0 is linear
    assert lib.xrsd_jacobian_n_variants(ctypes.byref(lay.c), free, 0, None, 0) == -1
    rc = lib.xrsd_jacobian_batch(ctypes.byref(lay.c), ctypes.byref(O), 1 << 20, 64, 1 << 21, lay.c.n_params, 3, None,
                                 None, None, 0, free, 2, None, None, 0, 1e-6, 7, None, None, None, 1 << 22, None, None)
    assert rc == abi.EINVAL and b"mode" in lib.xrsd_last_error()
    rc = lib.xrsd_jacobian_batch(ctypes.byref(lay.c), ctypes.byref(O), 1 << 20, 64, 1 << 21, lay.c.n_params, 3, None,
                                 None, None, 0, free, 2, None, None, 0, 1e-6, abi.JAC_FULL, None, None, None, 1 << 22, None, None)
    assert rc == abi.EINVAL and b"NULL" in lib.xrsd_last_error()                                  # full mode needs I_obs and outputs
    # n_free == 0 ("pattern and running totals") is only legal without the reduction
    rc = lib.xrsd_jacobian_batch(ctypes.byref(lay.c), ctypes.byref(O), 1 << 20, 64, 1 << 21, lay.c.n_params, 3, None,
                                 1 << 23, None, 64, None, 0, None, None, 0, 1e-6, abi.JAC_FULL, 1 << 24, 1 << 25, 1 << 26, 1 << 22, None, None)
    assert rc == abi.EINVAL and b"n_free" in lib.xrsd_last_error()
    # a forced finite-difference variant: -(slot + 1)
    forced = (ctypes.c_int32 * 2)(-(lay.slot("p__I0") + 1), lay.slot("p__r"))
    assert lib.xrsd_jacobian_n_variants(ctypes.byref(lay.c), forced, 2, None, 0) == 3
    # the fused loop kernels
    assert lib.xrsd_objective_prepare(ctypes.byref(O), None, 16, None, None, 16, 2, None, None, 16, None) == abi.EINVAL
    assert lib.xrsd_objective_prepare(ctypes.byref(O), 1 << 20, 16, 1 << 21, None, 8, 2, 1 << 22, 1 << 23, 16, None) == abi.EINVAL
    assert b"leading dimension" in lib.xrsd_last_error()
    assert lib.xrsd_objective_prepare(ctypes.byref(O), None, 16, None, None, 16, 0, None, None, 16, None) == abi.OK
    Op = abi.Objective()
    Op.prepared = 1
    assert lib.xrsd_objective_prepare(ctypes.byref(Op), 1 << 20, 16, 1 << 21, None, 16, 2, 1 << 22, 1 << 23, 16, None) == abi.EINVAL
    assert lib.xrsd_pattern_chi2_batch(ctypes.byref(Op), 1 << 20, 16, 1 << 21, 16, 2, 1 << 22, None, 16, 1 << 23, None, None) == abi.EINVAL
    assert b"prepared" in lib.xrsd_last_error()                                                    # prepared needs the weight stream
    args = [ctypes.byref(O), 1 << 20, 16, None, 0, 4, None, None, 0, 1 << 21, 6, 1 << 22, 1 << 23, 1 << 24, 1 << 25, 1 << 26,
            1 << 27, 1 << 28, 1 << 29, 1 << 30, None, 0, None, 1, 8.0, 4.0, 1e-8, 1e-10, 1e10, None]
    bad = list(args)
    bad[15] = None                                                                                 # d_active
    assert lib.xrsd_lm_trial_update(*bad) == abi.EINVAL
    bad = list(args)
    bad[24] = 1.0                                                                                  # up must be > 1
    assert lib.xrsd_lm_trial_update(*bad) == abi.EINVAL and b"up" in lib.xrsd_last_error()
    bad = list(args)
    bad[3], bad[4] = 1 << 19, 8                                                                    # a pattern, but ld_I < n_q and no I_obs
    assert lib.xrsd_lm_trial_update(*bad) == abi.EINVAL
    bad = list(args)
    bad[5] = 0                                                                                     # nothing to do
    assert lib.xrsd_lm_trial_update(*bad) == abi.OK
    assert lib.xrsd_lm_accept(None, 4, None, 0, None, 0, 16, 2, 2, None, None, 0, None) == abi.EINVAL
    assert lib.xrsd_lm_accept(1 << 20, 4, 1 << 21, 16, 1 << 22, 64, 16, 2, 2, None, None, 0, None) == abi.EINVAL
    assert b"stride" in lib.xrsd_last_error()
    assert lib.xrsd_lm_accept(1 << 20, 4, 1 << 21, 64, None, 64, 16, 2, 2, None, None, 0, None) == abi.EINVAL
    assert lib.xrsd_lm_accept(1 << 20, 0, None, 0, None, 0, 16, 2, 2, None, None, 0, None) == abi.OK
    rc = lib.xrsd_lm_propose(1 << 20, 6, 2, free, 2, None, None, 0, 1 << 21, 1 << 22, 1 << 23, 1 << 24, 1 << 25, None, 1 << 26, -1.0, None, None)
    assert rc == abi.EINVAL and b"max_rel_step" in lib.xrsd_last_error()
    # the whole-iteration entry point checks its state before it launches anything
    assert lib.xrsd_sizeof_lm_state() == ctypes.sizeof(abi.LmState)
    assert lib.xrsd_lm_iteration(None, 0, None) == abi.EINVAL
    st = abi.LmState()
    L1, O1 = abi.Layout(), abi.Objective()
    st.layout, st.obj = ctypes.pointer(L1), ctypes.pointer(O1)
    assert lib.xrsd_lm_iteration(ctypes.byref(st), 0, None) == abi.EINVAL and b"prepared" in lib.xrsd_last_error()
    O1.prepared = 1
    assert lib.xrsd_lm_iteration(ctypes.byref(st), 0, None) == abi.EINVAL and b"workspace" in lib.xrsd_last_error()
    # constraint terms: a program word must be known and the expression stack must balance
    tw = abi.TIE_WORD
    for words, msg in (([-(tw + 99), 0], b"unknown expression word"), ([-(tw + 2), 0], b"underrun"),
                       ([-(tw + 1), 0, -(tw + 1), 1], b"closing store"), ([0, -1], b"exactly one value")):
        arr = (ctypes.c_int32 * len(words))(*words)
        coef = (ctypes.c_double * len(words))(*([0.0] * len(words)))
        rc = lib.xrsd_lm_propose(1 << 20, 6, 2, free, 2, arr, coef, len(words) // 2, 1 << 21, 1 << 22, 1 << 23, 1 << 24, 1 << 25, None,
                                 1 << 26, 0.25, None, None)
        assert rc == abi.EINVAL and msg in lib.xrsd_last_error(), (words, lib.xrsd_last_error())


def test_uniform_step_host_helper_matches_its_numpy_statement():
    """xrsd_uniform_step is plain host code (no device call): the C pass and the numpy statement in
    lowering.uniform_step agree on linspace grids and on everything that is not one."""
    import ctypes as C
    from xrsdkit_b200 import abi, lowering
    lib = abi.load()

    def numpy_statement(q):
        n = len(q)
        if n < 3:
            return 0.0
        step = (q[-1] - q[0]) / (n - 1)
        if not (step > 0):
            return 0.0
        tol = 4 * np.finfo(float).eps * max(abs(q[0]), abs(q[-1]))
        ok = np.max(np.abs(q - (q[0] + np.arange(n) * step))) <= tol and np.all(np.diff(q) > 0)
        return float(step) if ok else 0.0

    rs = np.random.RandomState(0)
    grids = [np.linspace(0.001, 0.5, 2000), np.linspace(0.0, 1.0, 8192), np.linspace(0.01, 1.0, 3), np.linspace(0, 1, 2),
             np.logspace(-3, 0, 500), np.linspace(1.0, 0.0, 100), np.arange(0.0, 0.5, 0.001), np.sort(rs.rand(300)),
             np.full(10, 0.3)]
    bumped = np.linspace(0.01, 0.5, 1001)
    bumped[777] += 1e-9
    with_nan = np.linspace(0.01, 0.5, 1001)
    with_nan[3] = np.nan
    grids += [bumped, with_nan]
    for q in grids:
        q = np.ascontiguousarray(q, dtype=np.float64)
        want = numpy_statement(q)
        got_c = float(lib.xrsd_uniform_step(q.ctypes.data_as(C.c_void_p), len(q)))
        assert got_c == want, (len(q), got_c, want)
        assert lowering.uniform_step(q) == want
        ro = q.copy()
        ro.flags.writeable = False                 # abi.ptr falls back to ndarray.ctypes for read-only buffers
        assert lowering.uniform_step(ro) == want
    assert numpy_statement(grids[0]) > 0 and numpy_statement(grids[4]) == 0.0
