"""System: a set of Populations plus a NoiseModel (API of reference system/__init__.py).

Same constructor, dict schema, parameter plumbing and fit-objective definition as xrsdkit;
compute_intensity / evaluate_residual / fit run on the GPU engine (engine.py), and the
batched entry points in xrsdkit_b200.batch evaluate many parameter sets of one layout in a
single launch."""
import copy

import numpy as np

from .noise import NoiseModel
from .population import Population
from .. import definitions as xrsdefs


class System(object):

    def __init__(self, **kwargs):
        self.populations = {}
        self.fit_report = dict(error_weighted=True, logI_weighted=True, good_fit=False,
                               q_range=[0., float('inf')])
        self.features = dict.fromkeys(xrsdefs.profile_keys)
        kwargs = copy.deepcopy(kwargs)
        src_wl = kwargs.pop('source_wavelength') if 'source_wavelength' in kwargs else 1.
        self.sample_metadata = dict(experiment_id='', sample_id='', data_file='',
                                    source_wavelength=src_wl, time=0., notes='')
        self.noise_model = NoiseModel('flat')
        self.update_from_dict(kwargs)

    # -- dict round trip (reference :45-74) ---------------------------------------------
    def to_dict(self):
        sd = {}
        for name, pop in self.populations.items():
            sd[name] = pop.to_dict()
        sd['noise'] = self.noise_model.to_dict()
        sd['fit_report'] = copy.deepcopy(self.fit_report)
        sd['sample_metadata'] = self.sample_metadata
        sd['features'] = self.features
        return sd

    def clone(self):
        return System(**self.to_dict())

    @classmethod
    def from_dict(cls, d):
        return cls(**d)

    def update_from_dict(self, d):
        for key, val in d.items():
            if key == 'features':
                self.features.update(val)
            elif key == 'fit_report':
                self.fit_report.update(val)
            elif key == 'sample_metadata':
                self.sample_metadata.update(val)
            else:
                if not isinstance(val, dict):
                    val = val.to_dict()
                if key == 'noise':
                    self.update_noise_model(val)
                elif key in self.populations:
                    self.populations[key].update_from_dict(val)
                else:
                    self.populations[key] = Population.from_dict(val)

    def update_noise_model(self, noise_dict):
        if 'model' in noise_dict:
            self.noise_model.set_model(noise_dict['model'])
        if 'parameters' in noise_dict:
            self.noise_model.update_parameters(noise_dict['parameters'])

    def update_params_from_dict(self, pd):
        for name, popd in pd.items():
            if name == 'noise':
                self.update_noise_model(popd)
            elif 'parameters' in popd:
                for pname, rec in popd['parameters'].items():
                    self.populations[name].parameters[pname].update(rec)

    # -- objective switches (reference :91-98) ------------------------------------------
    def set_q_range(self, q_min, q_max):
        self.fit_report['q_range'] = [float(q_min), float(q_max)]

    def set_error_weighted(self, err_wtd):
        self.fit_report['error_weighted'] = bool(err_wtd)

    def set_logI_weighted(self, logI_wtd):
        self.fit_report['logI_weighted'] = bool(logI_wtd)

    def remove_population(self, pop_nm):
        self.populations.pop(pop_nm)

    def add_population(self, pop_nm, structure, form, settings={}, parameters={}):
        self.populations[pop_nm] = Population(structure, form, settings, parameters)

    def flatten_params(self):
        """'<population>__<parameter>' -> the (shared, mutable) parameter dict (reference :211-218)."""
        out = {}
        for pname, rec in self.noise_model.parameters.items():
            out['noise__' + pname] = rec
        for name, pop in self.populations.items():
            for pname, rec in pop.parameters.items():
                out[name + '__' + pname] = rec
        return out

    # -- numerics: GPU ------------------------------------------------------------------
    def compute_intensity(self, q):
        """I(q) = noise + sum of populations (reference :112-130), one kernel launch."""
        from .. import engine, lowering
        q = np.ascontiguousarray(q, dtype=np.float64)
        return engine.get_engine().intensity_single(lowering.lower_system(self, q), q)

    def compute_intensity_components(self, q):
        """{'total': I, 'noise': ..., <population name>: ...} from ONE launch -- what the reference's
        plotting helper computes with 2 + n_populations separate evaluations (visualization/__init__.py:23-31)."""
        from .. import engine, lowering
        q = np.asarray(q, dtype=float)
        lay = lowering.lower_system(self, q)
        total, parts = engine.get_engine().intensity_components(lay, q, lay.values[None, :])
        out = {'total': total[0].cpu().numpy(), 'noise': parts[0, 0].cpu().numpy()}
        for k, name in enumerate(self.populations.keys()):
            out[name] = parts[0, k + 1].cpu().numpy()
        return out

    def rescale_I0_to(self, q, I):
        """Scale every I0 (noise model and populations) by sum(I) / sum(I_computed), in place: the last step of the
        reference's system_from_prediction (models/predict.py:372-378), which brings a predicted System onto the
        scale of the measured pattern before it is fitted.  Returns the factor."""
        from .. import batch, lowering
        q = np.ascontiguousarray(q, dtype=np.float64)
        lay = lowering.lower_system(self, q)
        _, factor = batch.rescale_I0_batch(lay, q, np.asarray(I, dtype=float)[None, :], lay.values[None, :])
        f = float(factor[0])
        self.noise_model.parameters['I0']['value'] *= f
        for pop in self.populations.values():
            pop.parameters['I0']['value'] *= f
        return f

    def evaluate_residual(self, q, I, dI=None, I_comp=None):
        """Scalar weighted chi^2 between the model and (q, I) (reference :134-187)."""
        from .. import engine, lowering
        q = np.asarray(q, dtype=float)
        I = np.asarray(I, dtype=float)
        eng = engine.get_engine()
        rep = self.fit_report
        if I_comp is not None:
            return eng_chi2_given_intensity(eng, q, I, dI, np.asarray(I_comp, dtype=float), rep)
        lay = lowering.lower_system(self, q)
        obj = eng.objective(rep['error_weighted'], rep['logI_weighted'], rep['q_range'])
        chi2 = eng.residual(lay, obj, q, lay.values[None, :], I, dI)
        return float(chi2[0].item())

    def lmf_evaluate(self, lmf_params, q, I, dI=None):
        new_params = unpack_lmfit_params(lmf_params)
        old_params = self.flatten_params()
        old_params.update(new_params)
        self.update_params_from_dict(unflatten_params(old_params))
        return self.evaluate_residual(q, I, dI)

    def pack_lmfit_params(self):
        import lmfit   # optional third-party dependency of the reference (system/__init__.py:5)
        lmfp = lmfit.Parameters()
        for key, rec in self.flatten_params().items():
            lmfp.add(key, value=rec['value'], vary=bool(not rec['fixed']),
                     min=rec['bounds'][0], max=rec['bounds'][1])
            if rec['constraint_expr']:
                lmfp[key].set(vary=False)
                lmfp[key].set(expr=rec['constraint_expr'])
        return lmfp


def eng_chi2_given_intensity(eng, q, I, dI, I_comp, rep):
    """evaluate_residual(..., I_comp=given): the reduction alone, on the device (torch ops)."""
    torch = eng.torch
    tq, tI, tc = eng.to_device(q), eng.to_device(I), eng.to_device(I_comp)
    fit = (tI > 0) & (tq >= rep['q_range'][0]) & (tq <= rep['q_range'][1])
    w = torch.ones_like(tq)
    if rep['error_weighted']:
        if dI is None:
            d = torch.where(fit, torch.sqrt(torch.clamp(tI, min=0)), torch.full_like(tI, float('nan')))
        else:
            d = eng.to_device(np.broadcast_to(np.asarray(dI, dtype=float), q.shape))
        w = w * d ** 2
    if rep['logI_weighted']:
        fit = fit & (tc > 0)
        y1, y2 = torch.log(tc[fit]), torch.log(tI[fit])
    else:
        y1, y2 = tc[fit], tI[fit]
    w = w[fit]
    w = w / w.sum()
    return float(((y1 - y2) ** 2 * w).sum().item())


def fit(sys, q, I, dI=None, error_weighted=None, logI_weighted=None, q_range=None, method='lm'):
    """Fit I(q) and return a System with optimised parameters (reference :220-278).

    Same contract as xrsdkit.system.fit -- bounds, `fixed`, and the fit_report keys
    converged / initial_objective / final_objective / fit_snr -- but the optimiser is the
    batched GPU Levenberg-Marquardt of xrsdkit_b200.fitting (batch of one) instead of
    lmfit's Nelder-Mead; method='nelder-mead' runs the reference's algorithm (scipy's Nelder-Mead
    coefficients, lmfit's bound transform) on the same GPU objective.  A constraint_expr ('particle__r',
    '2*a__r + b__r - 3', 'sqrt(a__r*b__r)/2': any arithmetic expression of other parameters) is honoured exactly as
    lmfit does -- re-evaluated whenever the free parameters change (lowering.SystemLayout.ties); conditionals,
    comparisons and constraints on constrained parameters raise NotImplementedError.  The reference
    stores `error_weighted` into logI_weighted here (a bug, :255-256); this keeps the
    documented meaning."""
    from .. import fitting
    sys_opt = System.from_dict(sys.to_dict())
    if error_weighted is not None:
        sys_opt.fit_report.update(error_weighted=error_weighted)
    if logI_weighted is not None:
        sys_opt.fit_report.update(logI_weighted=logI_weighted)
    if q_range is not None:
        sys_opt.fit_report.update(q_range=q_range)
    q, I = np.asarray(q, dtype=float), np.asarray(I, dtype=float)
    fitting.fit_system_inplace(sys_opt, q, I, dI, method=method)
    if q.size >= 8:                              # the measured pattern's numerical profile (reference :276); fewer than 8
        from ..tools import profiler             # points leave the reference's windows empty (it raises there)
        sys_opt.features = profiler.profile_pattern(q, I)
    return sys_opt


def unpack_lmfit_params(lmfit_params):
    pd = {}
    for name, par in lmfit_params.items():
        pd[name] = dict(value=par.value, bounds=[par.min, par.max], fixed=not par.vary,
                        constraint_expr=par._expr)
        if par._expr:
            pd[name]['fixed'] = False
    return pd


def unflatten_params(flat_params):
    pd = {}
    for key, rec in flat_params.items():
        pop_name, param_name = key.split('__')[:2]
        pd.setdefault(pop_name, {}).setdefault('parameters', {})[param_name] = rec
    return pd
