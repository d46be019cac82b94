"""xrsdkit.tools-compatible helpers that sit on the intensity path."""
from . import peak_math  # noqa: F401


def compute_chi2(y1, y2, weights=None):
    """Reference tools/__init__.py:104-125, for callers that already hold both arrays on the host.
    (The fit objective itself is evaluated on the device: xrsd_residual_batch.)"""
    import numpy as np
    y1, y2 = np.asarray(y1, dtype=float), np.asarray(y2, dtype=float)
    if weights is None:
        return np.sum((y1 - y2) ** 2)
    weights = np.asarray(weights, dtype=float)
    return np.sum((y1 - y2) ** 2 * (weights / np.sum(weights)))
