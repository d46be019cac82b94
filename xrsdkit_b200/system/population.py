"""Population: one scattering species of a System (API of reference system/population.py:9-116).

Settings and parameters are plain dicts with the reference's schema
({'value','fixed','bounds','constraint_expr'} per parameter); defaults, validation errors and
key order follow xrsdkit so that dicts/YAML written by either package load in the other.
compute_intensity() runs on the GPU engine; there is no CPU path."""
import copy
from collections import OrderedDict

from .. import definitions as xrsdefs


class Population(object):

    def __init__(self, structure, form, settings={}, parameters={}):
        self.structure = None
        self.form = None
        self.settings = OrderedDict()
        self.parameters = OrderedDict()
        self.set_structure(structure)
        self.set_form(form)
        self.update_settings(settings)
        self.update_parameters(parameters)

    # -- structure / form ---------------------------------------------------------------
    def set_structure(self, structure):
        xrsdefs.validate(structure, self.form, self.settings)
        self.structure = structure
        self.update_settings()
        self.update_parameters()

    def set_form(self, form):
        xrsdefs.validate(self.structure, form, self.settings)
        self.form = form
        self.update_settings()
        self.update_parameters()

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
        valid = xrsdefs.all_params(self.structure, self.form, self.settings)
        for name in new_params:
            if name not in valid:
                raise ValueError('Parameter {} is not valid for structure: {}, form: {}, settings: {}'.format(
                    name, self.structure, self.form, self.settings))
        for name in list(self.parameters.keys()):
            if name in valid:
                valid[name].update(self.parameters[name])
            else:
                self.parameters.pop(name)
        for name, rec in new_params.items():
            valid[name].update(rec)
        self.parameters.update(valid)

    # -- (de)serialisation --------------------------------------------------------------
    def to_dict(self):
        return dict(
            structure=copy.copy(self.structure),
            form=copy.copy(self.form),
            settings=dict((k, copy.copy(v)) for k, v in self.settings.items()),
            parameters=dict((k, copy.deepcopy(v)) for k, v in self.parameters.items()))

    def update_from_dict(self, d):
        if 'structure' in d:
            self.set_structure(d['structure'])
        if 'form' in d:
            self.set_form(d['form'])
        if 'settings' in d:
            self.update_settings(d['settings'])
        if 'parameters' in d:
            self.update_parameters(d['parameters'])

    @classmethod
    def from_dict(cls, d):
        return cls(d['structure'], d['form'], d.get('settings', {}), d.get('parameters', {}))

    # -- intensity ----------------------------------------------------------------------
    def compute_intensity(self, q, source_wavelength):
        from .. import scattering
        return scattering.compute_intensity(q, source_wavelength, self.structure, self.form,
                                            self.settings, self.parameters)
