"""xrsdkit.scattering-compatible entry points, computed by the CUDA engine.

compute_intensity() keeps the reference signature (scattering/__init__.py:12) and is what
Population.compute_intensity calls.  The elementary functions the reference exposes next to it
(guinier_porod_intensity :69, spherical_normal_intensity :111) are provided with the same
signatures by evaluating a one-population system on the device."""
import numpy as np

from . import engine, lowering
from .system.population import Population


def compute_intensity(q, source_wavelength, structure, form, settings, parameters):
    q = np.asarray(q, dtype=float)
    lay = lowering.lower_population_only(structure, form, settings, parameters, q, source_wavelength)
    return engine.get_engine().intensity_single(lay, q)


def _one_population(q, structure, form, settings, values):
    pop = Population(structure, form, settings, dict((k, {'value': v}) for k, v in values.items()))
    return compute_intensity(q, 1.0, pop.structure, pop.form, pop.settings, pop.parameters)


def guinier_porod_intensity(q, rg, porod_exponent):
    return _one_population(q, 'diffuse', 'guinier_porod', {}, dict(rg=rg, D=porod_exponent, I0=1.))


def spherical_normal_intensity(q, r0, sigma, sampling_width=3.5, sampling_step=0.05):
    stg = dict(distribution='r_normal', sampling_width=sampling_width, sampling_step=sampling_step)
    return _one_population(q, 'diffuse', 'spherical', stg, dict(r=r0, sigma=sigma, I0=1.))


def spherical_ff_squared(q, r):
    return _one_population(q, 'diffuse', 'spherical', {}, dict(r=r, I0=1.))


def hard_sphere_sf(q, r_sphere, volume_fraction):
    """S(q) of structure_factors.py:3, obtained as (disordered / diffuse) of a Guinier-Porod D=0 probe
    whose form factor is identically 1."""
    return _one_population(q, 'disordered', 'guinier_porod', {},
                           dict(rg=1., D=0., I0=1., r_hard=r_sphere, v_fraction=volume_fraction))
