"""TEST INFRASTRUCTURE ONLY -- import shim for the *unmodified* reference.

Loads xrsdkit from /root/reference (read-only, exists only in the build container,
never on the GPU box) or from another root given by XRSD_REFERENCE_ROOT / load(root=...):
__graft_entry__.build() copies the reference package, unmodified, to oracle/_ref/ (git-ignored,
travels to the GPU box with the snapshot) so that bench.py's CPU arm can time the reference
itself there.  Used by oracle/make_golden.py and oracle/make_workloads.py to generate the
committed fixtures under tests/golden/, by tests/test_oracle_golden.py::test_oracle_equals_reference_*
(skipped when no tree is present) and by oracle/cpu_arm.py.  Nothing in the product package imports this.

Three blockers are patched, none of them on the numerical path:
  * xrsdkit/__init__.py:4 imports matplotlib        -> stub module
  * xrsdkit/system/__init__.py:5 imports lmfit      -> stub module (fit() unusable)
  * xrsdkit/definitions.py:61 calls yaml.load(f) without a Loader (PyYAML>=6 TypeError)
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("XRSD_REFERENCE_ROOT", "/root/reference")


def available(root=None):
    return os.path.isdir(os.path.join(root or REFERENCE_ROOT, "xrsdkit"))


def load(root=None):
    """Return the imported reference package (raises ImportError when absent)."""
    global REFERENCE_ROOT
    if root is not None:
        REFERENCE_ROOT = root
    if not available():
        raise ImportError("reference tree not present at %s" % REFERENCE_ROOT)
    if "xrsdkit" in sys.modules:
        return sys.modules["xrsdkit"]
    sys.dont_write_bytecode = True
    import yaml

    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        mpl.use = lambda *a, **k: None
        sys.modules["matplotlib"] = mpl
    if "lmfit" not in sys.modules:
        lm = types.ModuleType("lmfit")
        lm.Parameters = type("Parameters", (dict,), {})

        def _no_lmfit(*a, **k):
            raise NotImplementedError("lmfit is not installed in this image")

        lm.minimize = _no_lmfit
        sys.modules["lmfit"] = lm
    if not getattr(yaml.load, "_xrsd_shim", False):
        _orig = yaml.load

        def _load(stream, Loader=None, **kw):
            return _orig(stream, Loader=Loader or yaml.SafeLoader, **kw)

        _load._xrsd_shim = True
        yaml.load = _load
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import xrsdkit  # noqa: F401
    import xrsdkit.system  # noqa: F401
    import xrsdkit.scattering  # noqa: F401
    return sys.modules["xrsdkit"]
