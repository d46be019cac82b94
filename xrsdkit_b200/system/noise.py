"""NoiseModel: the additive background of a System (API of reference system/noise.py:8-63)."""
import copy

from .. import definitions as xrsdefs
from ._records import reconcile, snapshot


class NoiseModel(object):

    def __init__(self, model=None, parameters={}):
        self.model = model or 'flat'
        self.parameters = {}
        self.update_parameters(parameters)

    def to_dict(self):
        return dict(model=copy.copy(self.model), parameters=snapshot(self.parameters))

    def set_model(self, new_model):
        self.model = new_model
        self.update_parameters()

    def update_parameters(self, new_params={}):
        reconcile(self.parameters, copy.deepcopy(xrsdefs.noise_params[self.model]), new_params,
                  lambda nm: 'Parameter {} is not valid for noise model {}'.format(nm, self.model))

    def compute_intensity(self, q):
        if self.model not in xrsdefs.noise_model_names:
            raise ValueError('unsupported noise specification: {}'.format(self.model))
        from .. import lowering, engine
        return engine.get_engine().intensity_single(lowering.lower_noise_only(self, q), q)
