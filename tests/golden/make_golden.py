"""Writes tests/golden/*.npz by running the REFERENCE's own modules
(/root/reference, imported in place through oracle/refshim.py) on small seeded
synthetic models.  Run in the build container only:

    python tests/golden/make_golden.py

The un-vendored Payne package is replaced by oracle/payne_port.py, so the
vectors pin  fastMISTmod / likelihood / genmod / smoothing / fitutils.polycalc /
fitstar.lnprobfn / prior  exactly, and the ANN arithmetic only as restated
(PARITY UNPINNED, see oracle/payne_port.py).  Every file stores the model arrays
it was generated from, so the tests never depend on RNG reproducibility.
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from minesweeper_b200 import synth                     # noqa: E402
from oracle import refshim                             # noqa: E402

warnings.simplefilter('ignore')


def pack(prefix, d):
    return {prefix + k: np.asarray(v) for k, v in d.items()}


def unpack(z, prefix):
    return {k[len(prefix):]: z[k] for k in z.files if k.startswith(prefix)}


def demo_priordict(spec):
    pd = {'EEP': {'pv_uniform': [202, 261]}, 'initial_Mass': {'pv_uniform': [0.5, 1.25]},
          'initial_[Fe/H]': {'pv_uniform': [-4.0, 0.0]}, 'initial_[a/Fe]': {'pv_uniform': [-0.2, 0.6]},
          'Dist': {'pv_uniform': [1.0, 100.0]}, 'Av': {'pv_uniform': [0.0, 0.1]},
          'Age': {'uniform': [1.0, 14.0]}, 'Parallax': {'gaussian': [70.7371, 0.0631]},
          'IMF': {'IMF_type': 'Kroupa'}}
    if spec:
        pd.update({'Vrad': {'pv_uniform': [-5.0, 5.0]}, 'Vrot': {'pv_uniform': [0.0, 10.0]},
                   'Inst_R': {'fixed': 35000.0},
                   'blaze_coeff': [[1.0, 0.5], [0.0, 0.075], [0.0, 0.025], [0.0, 0.01]]})
    return pd


def gen_mist():
    mist = synth.make_mist(ne=60, nm=24, nf=7, na=5, seed=10, frac_missing=0.04, frac_trunc=0.15)
    refshim.register_h5('golden://mist', synth.mist_rows(mist))
    refshim.install()
    from minesweeper.fastMISTmod import GenMIST
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        G = GenMIST(MISTpath='golden://mist', ageweight=True, verbose=False)
    rng = np.random.default_rng(100)
    n = 3000
    ax = [mist['eep'], mist['mass'], mist['feh'], mist['afe']]
    q = np.stack([rng.uniform(a[0] - 0.02 * (a[-1] - a[0]), a[-1] + 0.02 * (a[-1] - a[0]), n) for a in ax], 1)
    # edge cases
    edge = []
    mid = [a[len(a) // 2] for a in ax]
    edge.append(mid)                                            # exact 4-axis grid hit
    for d in range(4):
        e = list(np.array(mid) + 1e-3)
        e[d] = ax[d][0]; edge.append(list(e))                   # exactly the bottom value
        e[d] = ax[d][-1]; edge.append(list(e))                  # exactly the top value -> out
        e[d] = np.nextafter(ax[d][-1], -np.inf); edge.append(list(e))
        e[d] = np.nextafter(ax[d][0], -np.inf); edge.append(list(e))
        e[d] = ax[d][3]; edge.append(list(e))                   # exact hit on one axis
        e[d] = np.nan; edge.append(list(e))
    # queries next to missing / truncated tracks
    miss = np.argwhere(~mist['valid'])
    for (ia, jf, im, ie) in miss[rng.choice(len(miss), 300, replace=False)]:
        pt = []
        for d, i in zip(range(4), (ie, im, jf, ia)):
            lo = ax[d][max(i - 1, 0)]; hi = ax[d][min(i + 1, len(ax[d]) - 1)]
            pt.append(rng.uniform(lo, hi))
        edge.append(pt)
    q = np.concatenate([q, np.array(edge)], 0)
    pred = np.full((len(q), 13), np.nan)
    ok = np.zeros(len(q), bool)
    br = np.zeros((len(q), 4), np.int32)
    for i, (e, m, f, a) in enumerate(q):
        out = G.getMIST(eep=e, mass=m, feh=f, afe=a)
        br[i] = [np.digitize([v], bins=G.gridpoints[p], right=False)[0] - 1
                 for v, p in zip((e, m, f, a), G.labels)]
        if out is not None:
            ok[i] = True
            pred[i] = out
    print('mist: %d queries, %d ok' % (len(q), ok.sum()))
    np.savez_compressed(os.path.join(HERE, 'mist.npz'), theta4=q, ok=ok, pred=pred, brackets=br,
                        **pack('mist_', mist))


def gen_smoothing():
    refshim.install()
    from minesweeper import smoothing as S
    spec = synth.make_spec_net(npix=4096, h=16, seed=21, nlines=350)
    wave = spec['wavelength']
    rng = np.random.default_rng(5)
    hvec = np.abs(rng.standard_normal(16))
    flux = spec['w2'].astype(np.float64) @ hvec + spec['b2']
    outwave = np.linspace(5152.0, 5185.0, 700)
    vs = np.array([0.2, 1.0, 3.7, 9.5, 25.0])
    vs_out = np.stack([S.smoothspec(wave, flux, v, outwave=None, smoothtype='vsini', fftsmooth=True,
                                    inres=0.0) for v in vs])
    rs = np.array([20000.0 * 2.355, 35000.0 * 2.355, 60000.0])
    r_out = np.stack([S.smoothspec(wave * (1 + 3.3 / 299792.458), flux, r, outwave=outwave,
                                   smoothtype='R', fftsmooth=True, inres=100000.0) for r in rs])
    # wavelength-dependent LSF (array sigma in wavelength units, smoothtype='lsf'): sigma = a + b (lambda - lambda_0)
    lsf_ab = np.array([[0.05, 0.0], [0.08, 4e-5], [0.15, -1e-5]])
    lsf_out = np.stack([S.smoothspec(wave, flux, a + b * (wave - wave[0]), outwave=outwave, smoothtype='lsf',
                                     fftsmooth=True) for a, b in lsf_ab])
    np.savez_compressed(os.path.join(HERE, 'smoothing.npz'), wave=wave, flux=flux, outwave=outwave,
                        vsini=vs, vsini_out=vs_out, rsigma=rs, r_out=r_out, r_inres=100000.0,
                        r_vshift=3.3, lsf_ab=lsf_ab, lsf_out=lsf_out)
    print('smoothing: done')


def gen_like(spec_on, name, n):
    mist = synth.make_mist(ne=60, nm=24, nf=7, na=5, seed=10, frac_missing=0.04, frac_trunc=0.15)
    refshim.register_h5('golden://mist', synth.mist_rows(mist))
    phot = synth.make_phot_net(h=32, seed=11)
    inputdict = dict(phot=synth.demo_obs_phot(), photANNpath=phot, MISTpath='golden://mist',
                     isochrone_prior=True, ageweight=True, priordict=demo_priordict(spec_on))
    extra = {}
    if spec_on:
        specnet = synth.make_spec_net(npix=4096, h=48, seed=12, nlines=300)
        from oracle.payne_port import PayneSpecPredict
        from oracle.like_port import blaze
        PP = PayneSpecPredict(nnpath=specnet)
        ow = np.linspace(5152.0, 5185.0, 300)
        _, truth = PP.getspec(Teff=5600.0, logg=4.4, feh=-0.2, afe=0.1, rad_vel=1.2, rot_vel=3.0,
                              inst_R=35000.0 * 2.355, outwave=ow)
        truth = truth * blaze([1.02, 0.03, -0.01, 0.004], ow)
        obs = synth.make_obs_spectrum(specnet, n_obs=300, wave_lo=5152.0, wave_hi=5185.0, flux=truth)
        obs['obs_flux'][17] = np.nan                  # a masked pixel -> dropped by nansum
        inputdict['spec'] = dict(obs_wave=obs['obs_wave'], obs_flux=obs['obs_flux'],
                                 obs_eflux=obs['obs_eflux'], modpoly=True, convertair=False)
        inputdict['specANNpath'] = specnet
        inputdict['NNtype'] = 'YST1'
        extra.update(pack('spec_', specnet))
        extra.update(pack('obs_', obs))
    fs = refshim.build_fit(inputdict)
    L, P = fs.likeobj, fs.priorobj
    from minesweeper import fitstar
    rng = np.random.default_rng(200 + int(spec_on))
    theta = np.array([P.priortrans(rng.random(L.ndim)) for _ in range(n)], dtype=np.float64)
    ie = L.fitpars_i.index('EEP')
    theta[:max(n // 25, 2), ie] = 900.0                           # out of grid -> -inf
    lnl = np.empty(n); lnp = np.empty(n)
    keys = None
    rows = []
    mags = np.full((n, len(L.GM.filterarray)), np.nan)
    nflux = 8 if spec_on else 0
    fluxes = []
    for i, th in enumerate(theta):
        lnp[i] = fitstar.lnprobfn(list(th), L, P)
        lnl[i] = L.lnlikefn(list(th))
        pd = dict(L.parsdict)
        if keys is None and np.isfinite(lnl[i]):
            keys = list(pd.keys())
        rows.append(pd)
        if np.isfinite(lnl[i]):
            photpars = [pd[k] for k in ('Teff', 'log(g)', '[Fe/H]', '[a/Fe]', 'log(R)', 'Dist', 'Av')]
            sed = L.GM.genphot(photpars)
            mags[i] = [sed[b] for b in L.GM.filterarray]
            if spec_on and len(fluxes) < nflux:
                sp = [pd.get(k, np.nan) for k in ('Teff', 'log(g)', '[Fe/H]', '[a/Fe]', 'Vrad', 'Vrot', 'Vmic', 'Inst_R')]
                sp += [pd[k] for k in L.fitpars_i if 'pc' in k]
                _, fl = L.GM.genspec(sp, outwave=fs.fitargs['obs_wave_fit'], modpoly=True)
                fluxes.append((i, np.array(sp, dtype=np.float64), fl))
    pars = np.array([[r.get(k, np.nan) for k in keys] for r in rows], dtype=np.float64)
    out = dict(theta=theta, lnl=lnl, lnprob=lnp, parsdict_keys=np.array(keys), parsdict=pars,
               mags=mags, bands=np.array(L.GM.filterarray), fitpars_i=np.array(L.fitpars_i),
               obs_phot_bands=np.array(list(fs.fitargs['obs_phot'].keys())),
               obs_phot=np.array([fs.fitargs['obs_phot'][b] for b in fs.fitargs['obs_phot'].keys()]),
               fixed_keys=np.array(list(fs.fitargs['fixedpars'].keys())),
               fixed_vals=np.array([fs.fitargs['fixedpars'][k] for k in fs.fitargs['fixedpars']], dtype=np.float64),
               runbools=np.array([fs.spec_bool, fs.phot_bool, fs.modpoly_bool, fs.photscale_bool, fs.flux_bool]),
               fitpars_all=np.array(fs.fitpars),
               fitpars_bool=np.array([fs.fitpars_bool[k] for k in fs.fitpars]))
    if spec_on:
        out['flux_rows'] = np.array([f[0] for f in fluxes])
        out['flux_specpars'] = np.stack([f[1] for f in fluxes])
        out['flux'] = np.stack([f[2] for f in fluxes])
        out['obs_wave_fit'] = fs.fitargs['obs_wave_fit']
    out.update(pack('mist_', mist)); out.update(pack('phot_', phot)); out.update(extra)
    np.savez_compressed(os.path.join(HERE, name), **out)
    print('%s: %d points, %d finite, lnl range %.3g..%.3g' % (
        name, n, np.isfinite(lnl).sum(), np.nanmin(lnl[np.isfinite(lnl)]), np.nanmax(lnl[np.isfinite(lnl)])))


def gen_prior():
    """u -> theta through the reference's prior.priortrans, and lnpriorfn on the parsdict the
    reference's likelihood leaves behind, for a priordict that exercises every supported kind."""
    import json
    mist = synth.make_mist(ne=60, nm=24, nf=7, na=5, seed=10, frac_missing=0.04, frac_trunc=0.15)
    refshim.register_h5('golden://mist', synth.mist_rows(mist))
    phot = synth.make_phot_net(h=32, seed=11)
    priordict = {'EEP': {'pv_uniform': [202, 261]},
                 'initial_Mass': {'pv_tgaussian': [0.5, 1.25, 1.0, 0.2]},
                 'initial_[Fe/H]': {'pv_gaussian': [-0.3, 0.4]},
                 'initial_[a/Fe]': {'pv_uniform': [0.6, -0.2], 'gaussian': [0.1, 0.2]},
                 'Dist': {'pv_loguniform': [1.0, 100.0]},
                 'Av': {'pv_exp': [0.0, 0.03], 'uniform': [0.0, 0.15]},
                 'Age': {'tgaussian': [0.5, 14.0, 5.0, 3.0]},
                 'log(Age)': {'uniform': [8.0, 10.15]},
                 'Parallax': {'gaussian': [70.7371, 30.0]},
                 'IMF': {'IMF_type': 'Kroupa'}}
    inputdict = dict(phot=synth.demo_obs_phot(), photANNpath=phot, MISTpath='golden://mist',
                     isochrone_prior=True, ageweight=True, priordict=priordict)
    fs = refshim.build_fit(inputdict)
    L, P = fs.likeobj, fs.priorobj
    rng = np.random.default_rng(321)
    n = 500
    u = rng.random((n, L.ndim))
    u[:5] = np.clip(u[:5], 1e-9, 1 - 1e-9)
    u[0, :] = 1e-12; u[1, :] = 1 - 1e-12
    theta = np.array([P.priortrans(list(r)) for r in u], dtype=np.float64)
    lnl = np.empty(n); lnp = np.empty(n); lnprob = np.empty(n)
    from minesweeper import fitstar
    for i, th in enumerate(theta):
        lnl[i] = L.lnlikefn(list(th))
        lnp[i] = P.lnpriorfn(dict(L.parsdict)) if np.isfinite(lnl[i]) else np.nan
        lnprob[i] = fitstar.lnprobfn(list(th), L, P)
    out = dict(u=u, theta=theta, lnl=lnl, lnprior=lnp, lnprob=lnprob, priordict=np.array(json.dumps(priordict)),
               fitpars_i=np.array(L.fitpars_i), fitpars_all=np.array(fs.fitpars),
               fitpars_bool=np.array([fs.fitpars_bool[k] for k in fs.fitpars]),
               runbools=np.array([fs.spec_bool, fs.phot_bool, fs.modpoly_bool, fs.photscale_bool, fs.flux_bool]),
               obs_phot_bands=np.array(list(fs.fitargs['obs_phot'].keys())),
               obs_phot=np.array([fs.fitargs['obs_phot'][b] for b in fs.fitargs['obs_phot'].keys()]),
               fixed_keys=np.array(list(fs.fitargs['fixedpars'].keys()), dtype='<U16'),
               fixed_vals=np.array([fs.fitargs['fixedpars'][k] for k in fs.fitargs['fixedpars']], dtype=np.float64))
    out.update(pack('mist_', mist)); out.update(pack('phot_', phot))
    np.savez_compressed(os.path.join(HERE, 'prior.npz'), **out)
    print('prior.npz: %d points, %d finite lnl, %d finite lnprob' % (n, np.isfinite(lnl).sum(), np.isfinite(lnprob).sum()))


def gen_prior_adv():
    """The advanced priors (prior.py:413-501 -> advancedpriors.py:410-833): Galactic density + age (GALAGE, which also
    replaces the Dist transform by AP.gal_ppf), rotation (VROT), alpha (ALPHA), a pv_beta transform, and the two keys the
    reference accepts but never applies (VTOT, AngDia); spectrum + photometry fit so that Vrot is a parameter."""
    import json
    mist = synth.make_mist(ne=60, nm=24, nf=7, na=5, seed=10, frac_missing=0.04, frac_trunc=0.15)
    refshim.register_h5('golden://mist', synth.mist_rows(mist))
    phot = synth.make_phot_net(h=32, seed=11)
    specnet = synth.make_spec_net(npix=2048, h=16, seed=12, nlines=150)
    ow = np.linspace(5146.0, 5170.0, 64)
    obs = synth.make_obs_spectrum(specnet, n_obs=64, wave_lo=5146.0, wave_hi=5170.0)
    priordict = {'EEP': {'pv_uniform': [202, 261]}, 'initial_Mass': {'pv_uniform': [0.5, 1.6]},
                 'initial_[Fe/H]': {'pv_uniform': [-2.0, 0.3]}, 'initial_[a/Fe]': {'pv_uniform': [-0.2, 0.6]},
                 'Dist': {'pv_uniform': [50.0, 8000.0]}, 'Av': {'pv_beta': [2.0, 5.0, 0.0, 0.4]},
                 'Vrad': {'pv_uniform': [-5.0, 5.0]}, 'Vrot': {'pv_uniform': [0.0, 20.0]}, 'Inst_R': {'fixed': 35000.0},
                 'GALAGE': {'lb_coords': [35.0, 52.0],
                            'pars': {'thin': {'min': 0.5, 'max': 14.0}, 'thick': {'min': 6.0, 'max': 14.0, 'mean': 10.0, 'sigma': 2.0},
                                     'halo': {'min': 8.0, 'max': 14.0, 'mean': 12.0, 'sigma': 2.0}}},
                 'VROT': {'dwarf': {'a': -10.0, 'c': 10.0, 'n': 0.4}, 'giant': {'a': -10.0, 'c': 7.0, 'n': 1.0}},
                 'ALPHA': {'min': 0.1}, 'VTOT': {'pmra': 3.0, 'pmdec': -4.0}, 'AngDia': {'gaussian': [0.7, 0.05]},
                 'IMF': {'IMF_type': 'Kroupa'}}
    inputdict = dict(phot=synth.demo_obs_phot(), photANNpath=phot, MISTpath='golden://mist', specANNpath=specnet,
                     NNtype='YST1', isochrone_prior=True, ageweight=True, priordict=priordict,
                     spec=dict(obs_wave=obs['obs_wave'], obs_flux=obs['obs_flux'], obs_eflux=obs['obs_eflux'],
                               modpoly=False, convertair=False))
    fs = refshim.build_fit(inputdict)
    L, P = fs.likeobj, fs.priorobj
    rng = np.random.default_rng(654)
    n = 400
    u = rng.random((n, L.ndim))
    u[0, :] = 1e-9; u[1, :] = 1 - 1e-9
    theta = np.array([P.priortrans(list(r)) for r in u], dtype=np.float64)
    lnl = np.empty(n); lnp = np.full(n, np.nan)
    keys, rows = None, []
    for i, th in enumerate(theta):
        lnl[i] = L.lnlikefn(list(th))
        pd = dict(L.parsdict)
        if np.isfinite(lnl[i]):
            lnp[i] = P.lnpriorfn(dict(pd))
            if keys is None:
                keys = list(pd.keys())
        rows.append(pd)
    pars = np.array([[r.get(k, np.nan) if np.ndim(r.get(k, np.nan)) == 0 else np.nan for k in keys] for r in rows], dtype=np.float64)
    out = dict(u=u, theta=theta, lnl=lnl, lnprior=lnp, parsdict_keys=np.array(keys), parsdict=pars,
               priordict=np.array(json.dumps(priordict)), fitpars_i=np.array(L.fitpars_i), fitpars_all=np.array(fs.fitpars),
               fitpars_bool=np.array([fs.fitpars_bool[k] for k in fs.fitpars]),
               runbools=np.array([fs.spec_bool, fs.phot_bool, fs.modpoly_bool, fs.photscale_bool, fs.flux_bool]),
               obs_phot_bands=np.array(list(fs.fitargs['obs_phot'].keys())),
               obs_phot=np.array([fs.fitargs['obs_phot'][b] for b in fs.fitargs['obs_phot'].keys()]),
               fixed_keys=np.array(list(fs.fitargs['fixedpars'].keys()), dtype='<U16'),
               fixed_vals=np.array([fs.fitargs['fixedpars'][k] for k in fs.fitargs['fixedpars']], dtype=np.float64),
               obs_wave_fit=fs.fitargs['obs_wave_fit'], obs_flux_fit=fs.fitargs['obs_flux_fit'],
               obs_eflux_fit=fs.fitargs['obs_eflux_fit'])
    out.update(pack('mist_', mist)); out.update(pack('phot_', phot)); out.update(pack('spec_', specnet))
    np.savez_compressed(os.path.join(HERE, 'prior_adv.npz'), **out)
    print('prior_adv.npz: %d points, %d finite lnl, %d finite lnprior, fitpars %s' % (
        n, np.isfinite(lnl).sum(), np.isfinite(lnp).sum(), L.fitpars_i))


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'prior_adv':
        gen_prior_adv()
        sys.exit(0)
    gen_mist()
    gen_smoothing()
    gen_like(False, 'like_phot.npz', 400)
    gen_like(True, 'like_spec.npz', 120)
    gen_prior()
    gen_prior_adv()
