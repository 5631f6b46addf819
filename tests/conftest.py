import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200); run with -m gpu')
    config.addinivalue_line('markers', 'needs_reference: needs the reference checkout at /root/reference')


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def sub(d, prefix):
    return {k[len(prefix):]: v for k, v in d.items() if k.startswith(prefix)}


@pytest.fixture(scope='session')
def g_mist():
    return load_golden('mist.npz')


@pytest.fixture(scope='session')
def g_smooth():
    return load_golden('smoothing.npz')


@pytest.fixture(scope='session')
def g_like_phot():
    return load_golden('like_phot.npz')


@pytest.fixture(scope='session')
def g_like_spec():
    return load_golden('like_spec.npz')


def fit_setup(g):
    """(fitargs, fitpars, runbools) of a golden likelihood file, with the model
    containers taken from the file itself."""
    fitargs = dict(ageweight=True, mistdif=True,
                   fixedpars={str(k): float(v) for k, v in zip(g['fixed_keys'], g['fixed_vals'])},
                   MISTpath=sub(g, 'mist_'), photANNpath=sub(g, 'phot_'),
                   obs_phot={str(b): [float(m), float(e)] for b, (m, e) in zip(g['obs_phot_bands'], g['obs_phot'])})
    runbools = [bool(b) for b in g['runbools']]
    if runbools[0]:
        fitargs['specANNpath'] = sub(g, 'spec_')
        fitargs['NNtype'] = 'YST1'
        fitargs['obs_wave_fit'] = g['obs_wave_fit']
        fitargs['obs_flux_fit'] = g['obs_obs_flux']
        fitargs['obs_eflux_fit'] = g['obs_obs_eflux']
    names = [str(n) for n in g['fitpars_all']]
    fitpars = [names, {n: bool(b) for n, b in zip(names, g['fitpars_bool'])}]
    return fitargs, fitpars, runbools
