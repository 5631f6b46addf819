#!/usr/bin/env python
"""Config 4 (BASELINE.json configs[3]): one full nested-sampling fit of one synthetic star
(10 bands + Gaussian parallax prior, demo sampler settings: static, rwalk, 150 live points,
walks 5, dlogz 0.01) with the GPU proposal-queue sampler, next to the cost of the same number
of sequential likelihood calls on one CPU core with the reference's own likelihood class from
oracle/_ref (the oracle port where that copy is absent; dynesty itself is not installed, so the
CPU figure is calls x measured per-call time, stated as such).  bench.py reports the
photometric form of this configuration as ``catalog.single_fit``.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tools'))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--queue', type=int, default=64)
    ap.add_argument('--walks', type=int, default=5, help='demo/etc/samplerinfo.dat: 5; the accuracy gate needs 25')
    ap.add_argument('--nlive', type=int, default=150)
    ap.add_argument('--dlogz', type=float, default=0.01)
    ap.add_argument('--repeat', type=int, default=5)
    ap.add_argument('--cpu-calls', type=int, default=2000)
    ap.add_argument('--spec', action='store_true', help='spectrum + photometry fit (the demo/runMS.py configuration)')
    ap.add_argument('--nobs', type=int, default=1536)
    args = ap.parse_args()
    import torch
    from minesweeper_b200 import synth
    from minesweeper_b200.likelihood import likelihood
    from minesweeper_b200.prior import prior
    from minesweeper_b200.sampler import QueuedNestedSampler
    from bench_catalog import NAMES, ON, catalog_priordict
    mist = synth.make_mist(seed=0)
    phot = synth.make_phot_net(h=128, seed=1)
    runbools = [bool(args.spec), True, bool(args.spec), False, False]
    names, on = list(NAMES), set(ON)
    priordict = catalog_priordict(808.0)
    fitargs = dict(ageweight=True, mistdif=True, fixedpars={}, MISTpath=mist, photANNpath=phot,
                   obs_phot=synth.demo_obs_phot())
    if args.spec:                                     # demo/runMS.py:139-166
        spec = synth.make_spec_net(npix=10240, h=300, seed=2)
        obs = synth.make_obs_spectrum(spec, n_obs=args.nobs, seed=4)
        names += ['pc_0', 'pc_1', 'pc_2', 'pc_3']
        on |= {'Vrad', 'Vrot', 'pc_0', 'pc_1', 'pc_2', 'pc_3'}
        priordict.update({'Vrad': {'pv_uniform': [-5.0, 5.0]}, 'Vrot': {'pv_uniform': [0.0, 10.0]},
                          'Inst_R': {'fixed': 35000.0},
                          'blaze_coeff': [[1.0, 0.5], [0.0, 0.075], [0.0, 0.025], [0.0, 0.01]]})
        fitargs.update(specANNpath=spec, NNtype='YST1', fixedpars={'Inst_R': 35000.0}, obs_wave_fit=obs['obs_wave'],
                       obs_flux_fit=obs['obs_flux'], obs_eflux_fit=obs['obs_eflux'])
    fitpars = [names, {k: (k in on) for k in names}]
    L = likelihood(fitargs, fitpars, runbools, precision='tc', verbose=False)
    P = prior(fitargs, priordict, fitpars, runbools, likeobj=L)
    dev = L.engine.device
    # truth from the prior with a valid MIST prediction; observations from the model itself
    gen = torch.Generator(device=dev).manual_seed(123)
    while True:
        u = torch.rand(64, L.ndim, generator=gen, device=dev, dtype=torch.float64)
        th = P.priortrans_batch(u)
        lnl, der, _ = L.lnlike_batch(th)
        good = torch.isfinite(lnl).nonzero(as_tuple=True)[0]
        if len(good):
            break
    truth, dtr = th[good[:1]], der[good[:1]]
    cols = {n: i for i, n in enumerate(L.fitpars_i)}
    pp = torch.stack([10 ** dtr[:, 8], dtr[:, 11], dtr[:, 9], dtr[:, 10], dtr[:, 6], truth[:, cols['Dist']],
                      truth[:, cols['Av']]], 1)
    mags = L.engine.phot_sed(pp)[0].cpu().numpy()
    rng = np.random.default_rng(5)
    obs_phot = {b: [float(m + 0.02 * rng.standard_normal()), 0.02] for b, m in zip(L.engine.bands, mags)}
    kw = {}
    if args.spec:
        sp = [float(10 ** dtr[0, 8]), float(dtr[0, 11]), float(dtr[0, 9]), float(dtr[0, 10]),
              float(truth[0, cols['Vrad']]), float(truth[0, cols['Vrot']]), float('nan'), 35000.0] + \
             [float(truth[0, cols['pc_%d' % i]]) for i in range(4)]
        flux = L.GM.genspec_batch([sp], modpoly=True)[0]
        kw = dict(obs_wave=obs['obs_wave'], obs_flux=flux + obs['obs_eflux'] * rng.standard_normal(len(flux)),
                  obs_eflux=obs['obs_eflux'])
        fitargs['obs_flux_fit'] = kw['obs_flux']
    L.engine.set_star(0, obs_phot=obs_phot, **kw)
    fitargs['obs_phot'] = obs_phot
    plx = torch.tensor([[1000.0 / float(truth[0, cols['Dist']]), 0.1]], dtype=torch.float64, device=dev)
    times, res = [], None
    for r in range(args.repeat + 1):
        ns = QueuedNestedSampler(L, P, nstars=1, nlive=args.nlive, walks=args.walks, queue=args.queue,
                                 dlogz=args.dlogz, seed=10 + r, star_plx=plx)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = ns.run()
        torch.cuda.synchronize()
        if r:
            times.append(time.perf_counter() - t0)
    # CPU: per-call cost of lnprob(priortrans(u)) with the per-point oracle, one core
    from oracle.cpu_baseline import make_likeobj      # the reference's own likelihood class where its copy exists
    ref, cpu_kind = make_likeobj(fitargs, fitpars, runbools)
    u = np.random.default_rng(0).random((args.cpu_calls, L.ndim))
    th = P.priortrans_batch(u).cpu().numpy()
    t0 = time.perf_counter()
    for row in th:
        ref.lnlikefn(list(row))
    cpu_call = (time.perf_counter() - t0) / args.cpu_calls
    seq_calls = int(res['niter'][0]) * args.walks + args.nlive
    print(json.dumps(dict(
        metric='single_star_fit_seconds', value=float(np.median(times)), unit='s', higher_is_better=False,
        queue=args.queue, nlive=args.nlive, walks=args.walks, dlogz=args.dlogz, times=times,
        iterations=int(res['niter'][0]), rounds=int(res['rounds']), gpu_loglike_calls=int(res['ncall'][0]),
        logz=float(res['logz'][0]), logzerr=float(res['logzerr'][0]), spec=bool(args.spec), ndim=L.ndim,
        truth=[float(v) for v in truth[0].cpu().numpy()], post_mean=[float(v) for v in res['mean'][0, :L.ndim]],
        post_std=[float(v) for v in res['std'][0, :L.ndim]],
        cpu_us_per_call_1core=cpu_call * 1e6, cpu_kind=cpu_kind, sequential_calls_equivalent=seq_calls,
        cpu_seconds_estimate_1core=seq_calls * cpu_call,
        note='CPU figure = (iterations x walks + nlive) sequential calls x measured per-call time of the '
             'per-point likelihood (cpu_kind) on one host core; the prior and sampler overhead of dynesty are not included')))


if __name__ == '__main__':
    main()
