"""ctypes mirror of include/xrsd_b200.h and loader of libxrsd_b200.so.

The library is the product's only compute path.  Loading fails loudly (ImportError with the
build hint) when the .so is missing; there is no Python/CPU fallback."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("XRSD_B200_LIB") or os.path.join(HERE, "libxrsd_b200.so")   # override: A/B timing of two builds

MAX_POPS, MAX_SPECIES, MAX_PAIRS, MAX_OPS, MAX_PARAMS, MAX_FREE, MAX_TIES = 6, 4, 10, 9, 64, 16, 48
TIE_WORD, TIE_STACK = 1000, 8      # expression-program words of the constraint terms (include/xrsd_b200.h)
TIE_OPS = ['pushc', 'pushp', 'add', 'sub', 'mul', 'div', 'pow', 'min', 'max', 'neg', 'sqrt', 'exp', 'log', 'log10', 'sin', 'cos',
           'tan', 'abs', 'asin', 'acos', 'atan']

POP_NONE, POP_NOISE_FLAT, POP_NOISE_LOW_Q, POP_NONCRYSTALLINE, POP_CRYSTALLINE = range(5)
FORM_ATOMIC, FORM_SPHERE, FORM_SPHERE_R_NORMAL, FORM_GUINIER_POROD = range(4)
PROFILE = dict(voigt=0, gaussian=1, lorentzian=2)
LAT_CUBIC, LAT_HCP, LAT_HEXAGONAL, LAT_ORTHORHOMBIC, LAT_TETRAGONAL, LAT_MONOCLINIC, LAT_RHOMBOHEDRAL, LAT_TRICLINIC = range(8)
OPS = dict(inversion=0, mirror_x=1, mirror_y=2, mirror_z=3, mirror_x_y=4, mirror_y_z=5, mirror_z_x=6,
           mirror_nx_y=7, mirror_ny_z=8, mirror_nz_x=9)
JAC_FULL, JAC_BASE_READY, JAC_PATTERNS_ONLY = 0, 1, 2      # mode of xrsd_jacobian_batch
TABLE_GRID_OVERFLOW, TABLE_LIST_OVERFLOW, TABLE_SHELL_OVERFLOW, TABLE_BAD_LATTICE = 1, 2, 4, 8
OK, EINVAL, ECUDA, ECAPACITY, ENOMEM = 0, -1, -2, -3, -4

i32, f64 = C.c_int32, C.c_double
N_FEATURES, FEATURE_MAX_Q = 21, 8192
ABI_VERSION = 5


class Pop(C.Structure):
    _fields_ = [
        ("kind", i32), ("disordered", i32), ("form", i32), ("profile", i32), ("sf_radial", i32),
        ("polz", i32), ("lorentz", i32), ("use_symmetry", i32), ("lattice", i32), ("n_sites", i32),
        ("n_ops", i32), ("ops", i32 * MAX_OPS), ("n_species", i32), ("species_form", i32 * MAX_SPECIES),
        ("p_I0", i32), ("p_r", i32), ("p_sigma", i32), ("p_rg", i32), ("p_D", i32), ("p_r_hard", i32),
        ("p_v_fraction", i32), ("p_flat_fraction", i32),
        ("p_lat", i32 * 6), ("p_hwhm", i32), ("p_hwhm_g", i32), ("p_hwhm_l", i32),
        ("p_uvw", (i32 * 3) * MAX_SPECIES), ("table_slot", i32), ("prune_extinct", i32), ("reserved0", i32),
        ("sampling_width", f64), ("sampling_step", f64), ("q_min", f64), ("q_max", f64),
        ("sites", (f64 * 3) * 8), ("atom_Z", f64 * MAX_SPECIES),
        ("atom_a", (f64 * 4) * MAX_SPECIES), ("atom_b", (f64 * 4) * MAX_SPECIES),
    ]


class Layout(C.Structure):
    _fields_ = [("n_pops", i32), ("n_params", i32), ("n_xtal", i32), ("q0_is_zero", i32),
                ("wavelength", f64), ("q_step", f64), ("pops", Pop * MAX_POPS)]


class Tables(C.Structure):
    _fields_ = [("cap_shell", i32), ("cap_refl", i32), ("n_pairs_max", i32), ("cap_list_hint", i32),
                ("n_shell", C.c_void_p), ("n_refl", C.c_void_p), ("n_all", C.c_void_p), ("status", C.c_void_p),
                ("shell_hkl", C.c_void_p), ("shell_geo", C.c_void_p), ("refl_hkl", C.c_void_p),
                ("big_scratch", C.c_void_p), ("cap_big", C.c_int64),
                ("table_of", C.c_void_p), ("build_flag", C.c_void_p), ("own_base", i32), ("shared_base", i32),
                ("n_shared", i32), ("n_grid_slots", i32), ("shared_tag", C.c_void_p), ("grid_scratch", C.c_void_p),
                ("cap_grid_cells", C.c_int64), ("grid_slot_counter", C.c_void_p)]


class Objective(C.Structure):
    _fields_ = [("error_weighted", i32), ("logI_weighted", i32), ("has_dI", i32), ("prepared", i32),
                ("q_lo", f64), ("q_hi", f64)]


vp = C.c_void_p
i64 = C.c_int64


class LmState(C.Structure):
    """xrsd_lm_state_t: everything one Levenberg-Marquardt iteration of a resident fit touches (xrsd_lm_iteration)."""
    _fields_ = [("layout", C.POINTER(Layout)), ("obj", C.POINTER(Objective)), ("d_q", vp), ("n_q", i32), ("ld_params", i32),
                ("d_params", vp), ("d_trial", vp), ("batch", i32), ("n_free", i32), ("d_fo", vp), ("d_w", vp), ("ld_obs", i64),
                ("free_idx", C.POINTER(i32)), ("jac_idx", C.POINTER(i32)), ("ties", C.POINTER(i32)), ("tie_coef", C.POINTER(f64)),
                ("n_tie", i32), ("n_var", i32), ("n_pops", i32), ("loop_cells", i32),
                ("rel_step", f64), ("max_rel_step", f64), ("up", f64), ("down", f64), ("ftol", f64), ("xtol", f64), ("lambda_max", f64),
                ("d_JTJ", vp), ("d_JTr", vp), ("d_chi2", vp), ("d_chi2_trial", vp), ("d_chi2_init", vp), ("d_lambda", vp),
                ("d_lo", vp), ("d_hi", vp), ("d_active", vp), ("d_accepted", vp), ("d_needjac", vp), ("d_reason", vp),
                ("d_n_accept", vp), ("d_flags", vp), ("tables", C.POINTER(Tables)), ("trial_tables", C.POINTER(Tables)),
                ("d_ws_base", vp), ("d_ws_trial", vp)]


_SIGNATURES = {
    "xrsd_last_error": (C.c_char_p, []),
    "xrsd_abi_version": (C.c_int, []),
    "xrsd_struct_sizes": (C.c_int, [C.POINTER(i32)] * 4),
    "xrsd_table_smem_bytes": (C.c_int64, [i32, i32, i32]),
    "xrsd_table_scratch_bytes": (C.c_int64, [C.c_int64]),
    "xrsd_reflection_tables": (C.c_int, [C.POINTER(Layout), vp, i32, i32, C.POINTER(Tables), i32, vp, vp]),
    "xrsd_intensity_batch": (C.c_int, [C.POINTER(Layout), vp, i32, vp, i32, i32, C.POINTER(Tables), vp, C.c_int64, vp]),
    "xrsd_intensity_batch_f32": (C.c_int, [C.POINTER(Layout), vp, i32, vp, i32, i32, vp, C.c_int64, vp]),
    "xrsd_intensity_components_batch": (C.c_int, [C.POINTER(Layout), vp, i32, vp, i32, i32, C.POINTER(Tables), vp, C.c_int64,
                                                  vp, vp]),
    "xrsd_residual_workspace_bytes": (C.c_int64, [i32, i32]),
    "xrsd_residual_batch": (C.c_int, [C.POINTER(Layout), C.POINTER(Objective), vp, i32, vp, i32, i32,
                                      C.POINTER(Tables), vp, vp, C.c_int64, vp, vp, vp, vp]),
    "xrsd_jacobian_workspace_bytes": (C.c_int64, [i32, i32, i32]),
    "xrsd_jacobian_batch": (C.c_int, [C.POINTER(Layout), C.POINTER(Objective), vp, i32, vp, i32, i32,
                                      C.POINTER(Tables), vp, vp, C.c_int64, C.POINTER(i32), i32, C.POINTER(i32), C.POINTER(f64), i32, f64, i32,
                                      vp, vp, vp, vp, vp, vp]),
    "xrsd_jacobian_n_variants": (i32, [C.POINTER(Layout), C.POINTER(i32), i32, C.POINTER(i32), i32]),
    "xrsd_pattern_chi2_batch": (C.c_int, [C.POINTER(Objective), vp, i32, vp, C.c_int64, i32, vp, vp, C.c_int64, vp, vp, vp]),
    "xrsd_objective_prepare": (C.c_int, [C.POINTER(Objective), vp, i32, vp, vp, i64, i32, vp, vp, i64, vp]),
    "xrsd_lm_trial_update": (C.c_int, [C.POINTER(Objective), vp, i32, vp, i64, i32, vp, vp, i64, vp, i32, vp, vp, vp, vp, vp, vp, vp,
                                       vp, vp, C.POINTER(Tables), i32, vp, i32, f64, f64, f64, f64, f64, vp]),
    "xrsd_lm_iteration": (C.c_int, [C.POINTER(LmState), i32, vp]),
    "xrsd_sizeof_lm_state": (i32, []),
    "xrsd_lm_accept": (C.c_int, [vp, i32, vp, i64, vp, i64, i32, i32, i32, C.POINTER(Tables), C.POINTER(Tables), i32, vp]),
    "xrsd_copy_rows_masked": (C.c_int, [vp, C.c_int64, vp, C.c_int64, C.c_int64, i32, vp, vp]),
    "xrsd_lm_propose": (C.c_int, [vp, i32, i32, C.POINTER(i32), i32, C.POINTER(i32), C.POINTER(f64), i32, vp, vp, vp, vp, vp, vp, vp, f64, vp, vp]),
    "xrsd_lm_update": (C.c_int, [vp, i32, i32, vp, vp, vp, vp, vp, vp, f64, f64, f64, f64, vp]),
    "xrsd_intensity_batch_host": (C.c_int, [C.POINTER(Layout), vp, i32, vp, i32, i32, vp, C.c_int64]),
    "xrsd_release_workspace": (None, []),
    "xrsd_peak_profile": (C.c_int, [i32, f64, f64, vp, i32, vp, vp]),
    "xrsd_uniform_step": (f64, [vp, i32]),
    "xrsd_feature_batch": (C.c_int, [vp, i32, vp, i64, i32, vp, i64, vp]),
    "xrsd_probe_fp64_peak": (C.c_int, [C.POINTER(f64), C.POINTER(f64), vp]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def load():
    """Load (once) and return the C-ABI library with argtypes set."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "xrsdkit_b200: %s is missing. Build it with `python -m xrsdkit_b200.build` "
            "(needs nvcc; compiles for sm_100a). There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError here = the .so does not match the header
        fn.restype = res
        fn.argtypes = args
    sizes = [i32() for _ in range(4)]
    lib.xrsd_struct_sizes(*[C.byref(s) for s in sizes])
    mine = [C.sizeof(Pop), C.sizeof(Layout), C.sizeof(Tables), C.sizeof(Objective)]
    if [s.value for s in sizes] != mine or lib.xrsd_sizeof_lm_state() != C.sizeof(LmState):
        raise ImportError("xrsdkit_b200: struct layout mismatch between abi.py %r and the library %r"
                          % (mine, [s.value for s in sizes]))
    if lib.xrsd_abi_version() != ABI_VERSION:
        raise ImportError("xrsdkit_b200: ABI version mismatch")
    _lib = lib
    return lib


def ptr(a):
    """Address of the data of a C-contiguous numpy array -- a quarter of the cost of ndarray.ctypes.data, which
    matters on the single-pattern path (four arrays per call)."""
    try:
        return C.addressof(C.c_char.from_buffer(a))
    except (TypeError, ValueError, BufferError):      # read-only or empty buffers
        return a.ctypes.data


class XrsdError(RuntimeError):
    pass


def check(rc):
    if rc == OK:
        return
    msg = load().xrsd_last_error().decode("utf-8", "replace")
    if rc == EINVAL:
        raise ValueError(msg)
    if rc == ECAPACITY:
        raise XrsdError("capacity: " + msg)
    raise XrsdError(msg)
