"""The C-ABI library loads on a CPU-only box and exports every symbol include/xrsd_b200.h declares."""
import ctypes
import os
import re

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
    assert lib.xrsd_abi_version() == 3
    assert lib.xrsd_table_smem_bytes(10, 10, 10) > 2 * 19 ** 3


def test_argument_errors_without_touching_the_gpu():
    lib = abi.load()
    L = abi.Layout()
    L.n_pops = 0
    rc = lib.xrsd_intensity_batch(ctypes.byref(L), None, 10, None, 1, 1, None, None, 10, None)
    assert rc == abi.EINVAL and b"n_pops" in lib.xrsd_last_error()
    rc = lib.xrsd_peak_profile(7, 1.0, 1.0, None, 0, None, None)
    assert rc == abi.EINVAL
