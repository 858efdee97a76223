import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def load_golden(name):
    """Load tests/golden/<name>.npz into a list of per-case dicts (+ top-level extras)."""
    z = np.load(os.path.join(GOLDEN, name + '.npz'))
    n_cases = int(z['n_cases'])
    cases = [dict() for _ in range(n_cases)]
    extras = {}
    for key in z.files:
        if key == 'n_cases':
            continue
        if key.startswith('case'):
            head, _, field = key.partition('_')
            cases[int(head[4:])][field] = z[key]
        else:
            extras[key] = z[key]
    return cases, extras


@pytest.fixture(scope='session')
def golden_villain_neighborhood():
    return load_golden('villain_neighborhood')[0]


@pytest.fixture(scope='session')
def golden_villain_observables():
    return load_golden('villain_observables')[0]


@pytest.fixture(scope='session')
def golden_lattice_forms():
    return load_golden('lattice_forms')


@pytest.fixture(scope='session')
def golden_worldline_checkerboard():
    return load_golden('worldline_checkerboard')[0]


@pytest.fixture(scope='session')
def golden_worldline_plaquette():
    return load_golden('worldline_plaquette')[0]


@pytest.fixture(scope='session')
def golden_worldline_observables():
    return load_golden('worldline_observables')[0]


@pytest.fixture(scope='session')
def golden_worldline_wrapping():
    return load_golden('worldline_wrapping')[0]


@pytest.fixture(scope='session')
def golden_villain_decoupled():
    return load_golden('villain_decoupled')[0]


@pytest.fixture(scope='session')
def golden_villain_cohomology():
    return load_golden('villain_cohomology')[0]


@pytest.fixture(scope='session')
def golden_autocorrelation():
    return load_golden('autocorrelation')[0]


@pytest.fixture(scope='session')
def golden_resampling():
    return load_golden('resampling')[0]


@pytest.fixture(scope='session')
def golden_taxicab():
    return load_golden('taxicab')[0]
