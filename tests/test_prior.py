"""Batched prior transform / ln-prior against golden vectors written by the reference's own
``prior`` class (tests/golden/prior.npz, make_golden.gen_prior)."""
import json

import numpy as np
import pytest

from conftest import load_golden, sub


@pytest.fixture(scope='module')
def g_prior():
    return load_golden('prior.npz')


def _setup(g):
    fitargs = dict(ageweight=True, mistdif=True, fixedpars={}, MISTpath=sub(g, 'mist_'), photANNpath=sub(g, 'phot_'),
                   obs_phot={str(b): [float(m), float(e)] for b, (m, e) in zip(g['obs_phot_bands'], g['obs_phot'])})
    names = [str(n) for n in g['fitpars_all']]
    fitpars = [names, {n: bool(b) for n, b in zip(names, g['fitpars_bool'])}]
    runbools = [bool(b) for b in g['runbools']]
    return fitargs, fitpars, runbools, json.loads(str(g['priordict']))


@pytest.mark.gpu
def test_prior_transform_and_lnprob_golden(g_prior):
    from minesweeper_b200.likelihood import likelihood
    from minesweeper_b200.prior import prior
    from minesweeper_b200 import fitstar
    g = g_prior
    fitargs, fitpars, runbools, priordict = _setup(g)
    L = likelihood(fitargs, fitpars, runbools, precision='f64', verbose=False)
    P = prior(fitargs, priordict, fitpars, runbools, likeobj=L)
    assert P.fitpars_i == [str(s) for s in g['fitpars_i']]
    # prior transform: uniform (reversed bounds), gaussian, truncated gaussian, exponential, log-uniform
    theta = P.priortrans_batch(g['u'])
    np.testing.assert_allclose(theta.cpu().numpy(), g['theta'], rtol=1e-10, atol=1e-12)
    # lnprob = lnlike + lnprior with the -inf short circuits (fitstar.py:715-727)
    lnl, der, _ = L.lnlike_batch(g['theta'])
    lnprob = P.lnprob_batch(g['theta'], lnl, der).cpu().numpy()
    assert np.array_equal(np.isneginf(lnprob), np.isneginf(g['lnprob']))
    fin = np.isfinite(g['lnprob'])
    assert fin.sum() > 300 and (~fin).sum() > 30
    np.testing.assert_allclose(lnprob[fin], g['lnprob'][fin], rtol=1e-12, atol=1e-6)
    # ln-prior alone
    lnp = P.lnprob_batch(g['theta'], None, der).cpu().numpy()
    ok = np.isfinite(g['lnl'])
    ref = g['lnprior'][ok]
    np.testing.assert_allclose(np.where(np.isfinite(lnp[ok]), lnp[ok], -np.inf),
                               np.where(np.isfinite(ref), ref, -np.inf), rtol=1e-12, atol=1e-9)
    # scalar API used by the unmodified lnprobfn
    for i in (3, 50, 200):
        th = P.priortrans(list(g['u'][i]))
        np.testing.assert_allclose(th, g['theta'][i], rtol=1e-10, atol=1e-12)
        v = fitstar.lnprobfn(list(g['theta'][i]), L, P)
        if np.isfinite(g['lnprob'][i]):
            assert abs(v - g['lnprob'][i]) <= 1e-6 + 1e-12 * abs(g['lnprob'][i])
        else:
            assert v == g['lnprob'][i]


def test_priordict_translation_rules():
    """CPU: column / term tables follow the reference's precedence (no GPU call)."""
    from minesweeper_b200 import prior as P
    from minesweeper_b200 import _lib

    class FakeEngine(object):
        pass
    names = ['Teff', 'log(g)', '[Fe/H]', '[a/Fe]', 'Vrad', 'Vrot', 'Vmic', 'Inst_R', 'log(R)', 'Dist', 'Av',
             'log(A)', 'EEP', 'initial_[Fe/H]', 'initial_[a/Fe]', 'initial_Mass', 'pc_0', 'pc_1']
    on = {'Vrad', 'Dist', 'Av', 'EEP', 'initial_[Fe/H]', 'initial_[a/Fe]', 'initial_Mass', 'pc_0', 'pc_1'}
    fitpars = [names, {k: (k in on) for k in names}]
    pd = {'EEP': {'pv_uniform': [300, 200], 'pv_gaussian': [250, 5]},      # uniform wins, bounds sorted
          'Dist': {'pv_gaussian': [10.0, 1.0], 'pv_loguniform': [1, 100]},
          'Age': {'uniform': [1, 14]}, 'Parallax': {'gaussian': [70.0, 0.1]}, 'Vrad': {'tgaussian': [-5, 5, 0, 1]},
          'blaze_coeff': [[1.0, 0.5], [0.0, 0.1]], 'IMF': {'IMF_type': 'Kroupa'}}
    p = P.prior(dict(fixedpars={'Inst_R': 35000.0}, ageweight=True, mistdif=True), pd, fitpars,
                [True, True, True, False, False], engine=FakeEngine())
    cols = {n: p._cols[i] for i, n in enumerate(p.fitpars_i)}
    assert cols['EEP'].kind == _lib.PRIOR_KINDS['uniform'] and list(cols['EEP'].p)[:2] == [200.0, 300.0]
    assert cols['Dist'].kind == _lib.PRIOR_KINDS['gaussian']
    assert cols['Av'].kind == _lib.PRIOR_KINDS['uniform'] and list(cols['Av'].p)[:2] == [0.0, 5.0]      # default range
    assert cols['pc_1'].kind == _lib.PRIOR_KINDS['tgaussian'] and list(cols['pc_1'].p) == [-0.5, 0.5, 0.0, 0.1]
    kinds = sorted((t.value_id, t.kind) for t in p._terms_c[:p._nterms])
    assert (_lib.MS_V_AGE, _lib.TERM_KINDS['uniform']) in kinds
    assert (_lib.MS_V_PARALLAX, _lib.TERM_KINDS['gaussian']) in kinds
    assert (_lib.param_id('Vrad'), _lib.TERM_KINDS['tgaussian_bounds']) in kinds       # spec family: bounds only
    assert p.imf_bool
    # advanced priors: GAL turns the Dist column into the tabulated distance quantile; VTOT / AngDia are accepted and
    # inert (dead code in the reference's lnpriorfn); pv_texp is refused (NameError in the reference)
    pd2 = dict(pd, GAL={'lb_coords': [30.0, 45.0]}, VTOT={'pmra': 1.0, 'pmdec': 2.0}, AngDia={'gaussian': [1.0, 0.1]},
               ALPHA={'min': 0.0}, Av={'pv_beta': [2.0, 5.0, 0.0, 0.3]}, Dist={'pv_uniform': [10.0, 3000.0]})
    p2 = P.prior(dict(fixedpars={'Inst_R': 35000.0}, ageweight=True, mistdif=True), pd2, fitpars,
                 [True, True, True, False, False], engine=FakeEngine())
    cols2 = {n: p2._cols[i] for i, n in enumerate(p2.fitpars_i)}
    assert cols2['Dist'].kind == _lib.PRIOR_KINDS['table'] and cols2['Dist'].p[0] == 1000.0
    assert cols2['Av'].kind == _lib.PRIOR_KINDS['beta'] and list(cols2['Av'].p) == [2.0, 5.0, 0.0, 0.3]
    assert p2._adv.gal == 1 and p2._adv.galage == 0 and p2._adv.alpha == 1 and p2._adv.vrot == 0
    assert abs(p2._adv.Zp - np.sin(np.deg2rad(45.0))) < 1e-15
    with pytest.raises(NotImplementedError):
        P.prior(dict(fixedpars={'Inst_R': 35000.0}), dict(pd, Dist={'pv_texp': [1, 100, 0, 10]}), fitpars,
                [True, True, True, False, False], engine=FakeEngine())


@pytest.fixture(scope='module')
def g_prior_adv():
    return load_golden('prior_adv.npz')


def _setup_adv(g):
    fitargs, fitpars, runbools, priordict = _setup(g)
    fitargs['fixedpars'] = {str(k): float(v) for k, v in zip(g['fixed_keys'], g['fixed_vals'])}
    fitargs.update(specANNpath=sub(g, 'spec_'), NNtype='YST1', obs_wave_fit=g['obs_wave_fit'],
                   obs_flux_fit=g['obs_flux_fit'], obs_eflux_fit=g['obs_eflux_fit'])
    return fitargs, fitpars, runbools, priordict


def test_galactic_distance_table_matches_reference_gal_ppf(g_prior_adv):
    """CPU: the host-built quantile table (minesweeper_b200/advancedpriors.py) reproduces the reference's
    ``1000 * AP.gal_ppf(u)`` distance transform (prior.py:333-336, advancedpriors.py:43-63, 665-670)."""
    from minesweeper_b200 import advancedpriors as AP
    g = g_prior_adv
    pd = json.loads(str(g['priordict']))
    l, b = pd['GALAGE']['lb_coords']
    x, cdf = AP.distance_table(l, b, pd['Dist']['pv_uniform'][0] / 1000.0, pd['Dist']['pv_uniform'][1] / 1000.0)
    i = [str(s) for s in g['fitpars_i']].index('Dist')
    np.testing.assert_allclose(1000.0 * np.interp(g['u'][:, i], cdf, x), g['theta'][:, i], rtol=1e-12)


@pytest.mark.gpu
def test_advanced_priors_golden(g_prior_adv):
    """GAL / GALAGE / VROT / ALPHA additive priors, the Galactic distance transform and pv_beta on the GPU against
    the reference's own prior class (tests/golden/prior_adv.npz); VTOT / AngDia are in the priordict and, as in the
    reference, change nothing."""
    from minesweeper_b200.likelihood import likelihood
    from minesweeper_b200.prior import prior
    g = g_prior_adv
    fitargs, fitpars, runbools, priordict = _setup_adv(g)
    L = likelihood(fitargs, fitpars, runbools, precision='f64', verbose=False)
    P = prior(fitargs, priordict, fitpars, runbools, likeobj=L)
    assert P.fitpars_i == [str(s) for s in g['fitpars_i']]
    theta = P.priortrans_batch(g['u']).cpu().numpy()
    ib = P.fitpars_i.index('Av')                                                  # the pv_beta column: scipy's own ppf is
    np.testing.assert_allclose(theta[:, ib], g['theta'][:, ib], rtol=1e-9, atol=1e-12)   # good to ~1e-10 in the far tails
    rest = [i for i in range(P.ndim) if i != ib]
    np.testing.assert_allclose(theta[:, rest], g['theta'][:, rest], rtol=1e-10, atol=1e-12)   # incl. the gal_ppf column
    lnl, der, _ = L.lnlike_batch(g['theta'])
    lnp = P.lnprob_batch(g['theta'], None, der).cpu().numpy()
    ok = np.isfinite(g['lnl'])
    ref = g['lnprior'][ok]
    assert np.array_equal(np.isneginf(lnp[ok]), np.isneginf(ref)) and np.isneginf(ref).sum() > 20
    fin = np.isfinite(ref)
    np.testing.assert_allclose(lnp[ok][fin], ref[fin], rtol=1e-10, atol=1e-7)
    # scalar API on a parsdict as lnprobfn passes it
    keys = [str(k) for k in g['parsdict_keys']]
    for i in np.flatnonzero(ok)[:5]:
        pdict = {k: float(v) for k, v in zip(keys, g['parsdict'][i])}
        v = P.lnpriorfn(pdict)
        assert (v == g['lnprior'][i]) or abs(v - g['lnprior'][i]) <= 1e-7 + 1e-10 * abs(g['lnprior'][i])


def test_golden_prior_vectors_are_what_the_reference_class_returns(g_prior):
    """The committed prior vectors are outputs of the reference: its own ``prior.prior`` class (checkout or the verbatim
    copy in oracle/_ref) maps the stored unit-cube points to the stored theta again, and its ``lnpriorfn`` on the
    parsdict the reference likelihood leaves behind gives the stored ln prior."""
    import contextlib
    import io
    from oracle import refshim
    from oracle.cpu_baseline import make_likeobj
    if not refshim.available():
        pytest.skip('no reference checkout and no oracle/_ref copy')
    g = g_prior
    fitargs, fitpars, runbools, priordict = _setup(g)
    R, kind = make_likeobj(fitargs, fitpars, runbools)
    assert kind == 'reference'
    from minesweeper.prior import prior as RefPrior
    with contextlib.redirect_stdout(io.StringIO()):
        P = RefPrior(fitargs, priordict, fitpars, runbools)
    assert type(P).__module__ == 'minesweeper.prior' and P.fitpars_i == [str(s) for s in g['fitpars_i']]
    for i in range(80):
        th = np.array(P.priortrans(list(g['u'][i])), dtype=np.float64)
        np.testing.assert_allclose(th, g['theta'][i], rtol=1e-14, atol=0)
        with contextlib.redirect_stdout(io.StringIO()), np.errstate(all='ignore'):
            lnl = R.lnlikefn(list(g['theta'][i]))
        assert (lnl == g['lnl'][i]) or abs(lnl - g['lnl'][i]) <= 1e-12 * max(1.0, abs(g['lnl'][i]))
        if np.isfinite(lnl):
            with np.errstate(all='ignore'):
                lnp = P.lnpriorfn(dict(R.parsdict))
            want = g['lnprior'][i]
            assert (lnp == want) or abs(lnp - want) <= 1e-12 * max(1.0, abs(want)), (i, lnp, want)
