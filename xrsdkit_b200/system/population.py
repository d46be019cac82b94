"""Population: one scattering species of a System (API of reference system/population.py:9-116).

Settings and parameters are plain dicts with the reference's schema
({'value','fixed','bounds','constraint_expr'} per parameter); defaults, validation errors and
key order follow xrsdkit so that dicts/YAML written by either package load in the other.
compute_intensity() runs on the GPU engine; there is no CPU path."""
import copy
from collections import OrderedDict

from .. import definitions as xrsdefs
from ._records import reconcile, snapshot


class Population(object):

    def __init__(self, structure, form, settings={}, parameters={}):
        self.structure, self.form = None, None
        self.settings, self.parameters = OrderedDict(), OrderedDict()
        self._reconfigure(structure=structure)
        self._reconfigure(form=form)
        self.update_settings(settings)
        self.update_parameters(parameters)

    # -- structure / form ---------------------------------------------------------------
    def _reconfigure(self, **change):
        """Switch structure or form: validate against the other one and the current settings first
        (an illegal combination must leave the object untouched), then rebuild settings and parameters."""
        target = dict(structure=self.structure, form=self.form)
        target.update(change)
        xrsdefs.validate(target['structure'], target['form'], self.settings)
        self.structure, self.form = target['structure'], target['form']
        self.update_settings()
        self.update_parameters()

    def set_structure(self, structure):
        self._reconfigure(structure=structure)

    def set_form(self, form):
        self._reconfigure(form=form)

    # -- settings -----------------------------------------------------------------------
    def update_settings(self, new_settings={}):
        """Rebuild the settings dict: primary defaults for (structure, form), then the
        secondary settings those imply; existing values win over defaults, `new_settings`
        win over both (reference population.py:33-61)."""
        old = self.settings

        def pick(name, default):
            if name in new_settings:
                return new_settings[name]
            if name in old:
                return copy.deepcopy(old[name])
            return default

        primary = OrderedDict()
        defaults = copy.deepcopy(xrsdefs.structure_settings[self.structure])
        if self.form:   # None while __init__ is still running
            defaults.update(copy.deepcopy(xrsdefs.form_settings[self.form]))
        for name, dflt in defaults.items():
            primary[name] = pick(name, dflt)
        merged = copy.deepcopy(primary)
        for name, dflt in xrsdefs.secondary_settings(self.structure, self.form, primary).items():
            merged[name] = pick(name, dflt)
        xrsdefs.validate(self.structure, self.form, merged)
        self.settings = merged
        self.update_parameters()

    # -- parameters ---------------------------------------------------------------------
    def update_parameters(self, new_params={}):
        """Keep exactly the parameters valid for the current settings (reference
        population.py:63-82); unknown names raise ValueError."""
        reconcile(self.parameters, xrsdefs.all_params(self.structure, self.form, self.settings), new_params,
                  lambda nm: 'Parameter {} is not valid for structure: {}, form: {}, settings: {}'.format(
                      nm, self.structure, self.form, self.settings))

    # -- (de)serialisation --------------------------------------------------------------
    _APPLY = (('structure', 'set_structure'), ('form', 'set_form'),
              ('settings', 'update_settings'), ('parameters', 'update_parameters'))

    def to_dict(self):
        return {'structure': self.structure, 'form': self.form,
                'settings': {k: copy.copy(v) for k, v in self.settings.items()},
                'parameters': snapshot(self.parameters)}

    def update_from_dict(self, d):
        for key, method in self._APPLY:     # in this order: each step narrows what the next may contain
            if key in d:
                getattr(self, method)(d[key])

    @classmethod
    def from_dict(cls, d):
        return cls(d['structure'], d['form'], d.get('settings', {}), d.get('parameters', {}))

    # -- intensity ----------------------------------------------------------------------
    def compute_intensity(self, q, source_wavelength):
        from .. import scattering
        return scattering.compute_intensity(q, source_wavelength, self.structure, self.form,
                                            self.settings, self.parameters)
