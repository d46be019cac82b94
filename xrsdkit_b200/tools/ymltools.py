"""On-disk System format (YAML) and a loader that feeds many files to the batched engine.

save_sys_to_yaml / load_sys_from_yaml keep the reference's file format
(tools/ymltools.py:18-26: a plain `yaml.dump(System.to_dict())`), so files written by xrsdkit load here
unchanged and vice versa.  load_batches() groups systems that share a layout (same populations,
settings and noise model) so that each group is evaluated in one launch."""
from collections import OrderedDict

import numpy as np
import yaml

from .. import lowering
from ..system import System


def _plain(o):
    """numpy scalars / OrderedDicts -> builtins, so the dump is readable by yaml.safe_load."""
    if isinstance(o, dict):
        return dict((str(k), _plain(v)) for k, v in o.items())
    if isinstance(o, (list, tuple)):
        return [_plain(v) for v in o]
    if isinstance(o, np.generic):
        return o.item()
    return o


def save_sys_to_yaml(file_path, sys):
    with open(file_path, 'w') as yaml_file:
        yaml.dump(_plain(sys.to_dict()), yaml_file)


def load_sys_from_yaml(file_path):
    with open(file_path, 'r') as yaml_file:
        sd = yaml.safe_load(yaml_file)
    return System(**sd)


def layout_signature(system):
    """Hashable description of everything but the parameter values and the metadata."""
    pops = []
    for name, pop in system.populations.items():
        pops.append((name, pop.structure, pop.form, tuple(sorted((k, repr(v)) for k, v in pop.settings.items()))))
    return (system.noise_model.model, tuple(pops), repr(system.sample_metadata.get('source_wavelength')))


def group_by_layout(systems, q):
    """[(SystemLayout, params[n, n_params], indices)] for an iterable of Systems on a common q grid."""
    groups = OrderedDict()
    for i, s in enumerate(systems):
        groups.setdefault(layout_signature(s), []).append((i, s))
    out = []
    for members in groups.values():
        lays = [lowering.lower_system(s, q) for _, s in members]
        params = np.stack([l.values for l in lays])
        out.append((lays[0], params, [i for i, _ in members]))
    return out


def load_batches(paths, q):
    """Load YAML files and group them for batched evaluation; returns (systems, groups)."""
    systems = [load_sys_from_yaml(p) for p in paths]
    return systems, group_by_layout(systems, q)


def compute_intensity_many(systems, q, eng=None):
    """I[i, :] for a list of Systems of mixed layouts: one launch per distinct layout."""
    from .. import batch
    q = np.asarray(q, dtype=float)
    out = np.empty((len(systems), q.size))
    for lay, params, idx in group_by_layout(systems, q):
        out[idx] = batch.compute_intensity_batch(lay, q, params, eng=eng)
    return out
