"""Peak profiles with the signatures of the reference's tools/peak_math.py:6-47, evaluated by
the engine's profile kernel (csrc/intensity_kernel.cu: profile_kernel; Voigt through the
Faddeeva implementation in csrc/faddeeva.cuh instead of scipy.special.wofz)."""
import numpy as np

from .. import engine


def _run(kind, x, p0, p1=0.0):
    x = np.asarray(x, dtype=float)
    return engine.get_engine().peak_profile(kind, x.ravel(), p0, p1).cpu().numpy().reshape(x.shape)


def gaussian(x, hwhm_g):
    return _run('gaussian', x, hwhm_g)


def lorentzian(x, hwhm_l):
    return _run('lorentzian', x, hwhm_l)


def voigt(x, hwhm_g, hwhm_l):
    return _run('voigt', x, hwhm_g, hwhm_l)


def gaussian_profile(q, q_pk, hwhm_g):
    return gaussian(np.asarray(q, dtype=float) - q_pk, hwhm_g)


def lorentzian_profile(q, q_pk, hwhm_l):
    return lorentzian(np.asarray(q, dtype=float) - q_pk, hwhm_l)


def voigt_profile(q, q_pk, hwhm_g, hwhm_l):
    return voigt(np.asarray(q, dtype=float) - q_pk, hwhm_g, hwhm_l)


def gaussian_function(hwhm):
    return lambda q, q_pk: gaussian_profile(q, q_pk, hwhm)


def lorentzian_function(hwhm):
    return lambda q, q_pk: lorentzian_profile(q, q_pk, hwhm)


def voigt_function(hwhm_g, hwhm_l):
    return lambda q, q_pk: voigt_profile(q, q_pk, hwhm_g, hwhm_l)
