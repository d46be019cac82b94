"""Pattern featuriser with the interface of the reference's tools/profiler.py, evaluated by the
engine's feature kernel (csrc/feature_kernel.cu): one CTA per pattern, the pattern held in shared
memory.  `profile_pattern(q, I)` is the reference call (tools/profiler.py:64); `profile_batch`
evaluates many patterns on a shared q grid in one launch."""
from collections import OrderedDict

import numpy as np

from .. import engine
from ..definitions import profile_keys

profile_defs = OrderedDict((k, v) for k, v in zip(profile_keys, (
    'maximum over mean intensity on the full q-range',
    'mean intensity on the lower 10% of the q-range, divided by the mean intensity on the full q-range',
    'maximum over mean intensity for q-values from 0.9*q(Imax) to 1.1*q(Imax)',
    'sum of |I(i+1)-I(i)| over the turning points of I, divided by the intensity range (Imax minus Imin)',
    'same as I_fluctuation but for log(I)',
    'max(log(I)) divided by std(log(I))',
    'centroid (in length units) of the magnitude of the fourier transform of the scattering spectrum',
    'q-space centroid of the scattering intensity',
    'q-space centroid of log(I)',
    'Pearson correlation between q and I(q)',
    'Pearson correlation between q squared and I(q)',
    'Pearson correlation between exp(q) and I(q)',
    'Pearson correlation between exp(-q) and I(q)',
    'q-vertex of parabola fit to intensity values near the best hump',
    'q-vertex of parabola fit to intensity values near the best trough',
    'Focal width of same parabola used for q_best_hump',
    'Focal width of same parabola used for q_best_trough',
    'like q_best_hump, but fit to standardized log(I)',
    'like q_best_trough, but fit to standardized log(I)',
    'like best_hump_qwidth, but fit to standardized log(I)',
    'like best_trough_qwidth, but fit to standardized log(I)')))


def profile_batch(q, I, eng=None):
    """Features of many patterns: q [n_q], I [B, n_q] (numpy or device tensors) -> device tensor [B, 21]
    in `profile_keys` order."""
    eng = eng or engine.get_engine()
    d_q = eng.to_device(q)
    d_I = eng.to_device(I)
    if d_I.dim() == 1:
        d_I = d_I[None, :]
    if d_I.shape[1] != d_q.numel():
        raise ValueError('I has %d points per pattern, q has %d' % (d_I.shape[1], d_q.numel()))
    return eng.features(d_q, d_I.contiguous())


def profile_pattern(q, I):
    """Numerical profile of one pattern; same keys, order and meaning as the reference's
    profile_pattern (tools/profiler.py:64-221)."""
    f = profile_batch(np.asarray(q, dtype=float), np.asarray(I, dtype=float)).cpu().numpy()[0]
    return OrderedDict((k, float(v)) for k, v in zip(profile_keys, f))
