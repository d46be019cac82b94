import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _unjson(o):
    if isinstance(o, dict):
        if set(o.keys()) == {"__float__"}:
            return float(o["__float__"])
        return {k: _unjson(v) for k, v in o.items()}
    if isinstance(o, list):
        return [_unjson(v) for v in o]
    return o


@pytest.fixture(scope="session")
def golden_definitions():
    with open(os.path.join(GOLDEN, "definitions.json")) as f:
        return _unjson(json.load(f))


@pytest.fixture(scope="session")
def golden_kernels():
    return np.load(os.path.join(GOLDEN, "kernels.npz"))


@pytest.fixture(scope="session")
def golden_tables():
    with open(os.path.join(GOLDEN, "tables.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_systems():
    with open(os.path.join(GOLDEN, "systems.json")) as f:
        meta = _unjson(json.load(f))
    arrays = np.load(os.path.join(GOLDEN, "systems.npz"))
    return meta, arrays


@pytest.fixture(scope="session")
def ref_tables(golden_definitions):
    """The reference's data tables the oracle needs (from the golden dump, not from the product)."""
    return dict(atomic_table=golden_definitions["atomic_params"],
                sg_point_groups=golden_definitions["sg_point_groups"],
                lattice_space_groups=golden_definitions["lattice_space_groups"])
