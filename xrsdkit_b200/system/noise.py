"""NoiseModel: the additive background of a System (API of reference system/noise.py:8-63)."""
import copy

from .. import definitions as xrsdefs


class NoiseModel(object):

    def __init__(self, model=None, parameters={}):
        self.model = model or 'flat'
        self.parameters = {}
        self.update_parameters(parameters)

    def to_dict(self):
        return dict(model=copy.copy(self.model),
                    parameters=dict((k, copy.deepcopy(v)) for k, v in self.parameters.items()))

    def set_model(self, new_model):
        self.model = new_model
        self.update_parameters()

    def update_parameters(self, new_params={}):
        valid = copy.deepcopy(xrsdefs.noise_params[self.model])
        for name in new_params:
            if name not in valid:
                raise ValueError('Parameter {} is not valid for noise model {}'.format(name, self.model))
        for name in list(self.parameters.keys()):
            if name in valid:
                valid[name].update(self.parameters[name])
            else:
                self.parameters.pop(name)
        for name, rec in new_params.items():
            valid[name].update(rec)
        self.parameters.update(valid)

    def compute_intensity(self, q):
        if self.model not in xrsdefs.noise_model_names:
            raise ValueError('unsupported noise specification: {}'.format(self.model))
        from .. import lowering, engine
        lay = lowering.lower_noise_only(self, q)
        return engine.get_engine().intensity_single(lay, q)
