#!/usr/bin/env python
"""Sampler settings of the catalog benchmark side by side in ONE process (one model build, one box): every setting fits
the same synthetic catalog and is checked against brute-force quadrature (tools/bench_catalog.py:accuracy_block).

    python tools/catalog_sweep.py --stars 20000 --settings rwalk:25:4 unif:1:16 unif:1:32
Each setting is ``sample:walks:queue[:stars]``; one JSON line per setting."""
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
    ap.add_argument('--stars', type=int, default=20000)
    ap.add_argument('--nlive', type=int, default=150)
    ap.add_argument('--dlogz', type=float, default=0.01)
    ap.add_argument('--maxiter', type=int, default=20000)
    ap.add_argument('--bruteforce', type=int, default=64)
    ap.add_argument('--check-every', type=int, default=16)
    ap.add_argument('--settings', nargs='+', default=['rwalk:25:4', 'unif:1:16', 'unif:1:32'])
    args = ap.parse_args()
    import torch
    import bench_catalog as bc
    from minesweeper_b200 import synth
    from minesweeper_b200.likelihood import likelihood
    from minesweeper_b200.prior import prior
    torch.cuda.set_device(0)
    dev = torch.device('cuda', 0)
    fitpars = [bc.NAMES, {k: (k in bc.ON) for k in bc.NAMES}]
    fitargs = dict(ageweight=True, mistdif=True, fixedpars={}, MISTpath=synth.make_mist(seed=0),
                   photANNpath=synth.make_phot_net(h=128, seed=1), obs_phot=synth.demo_obs_phot())
    runbools = [False, True, False, False, False]
    priordict = bc.catalog_priordict(808.0)
    L = likelihood(fitargs, fitpars, runbools, precision='tc', device=0, verbose=False)
    P = prior(fitargs, priordict, fitpars, runbools, likeobj=L)
    for setting in args.settings:
        parts = setting.split(':')
        sample, walks, queue = parts[:3]
        nstars = int(parts[3]) if len(parts) > 3 else args.stars
        ns, th = bc.make_catalog_sampler(L, P, nstars, dev, args.nlive, int(walks), args.dlogz, args.maxiter, 0,
                                         queued=int(queue), sample=sample)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = ns.run(check_every=args.check_every)
        torch.cuda.synchronize()
        secs = time.perf_counter() - t0
        acc = bc.accuracy_block(L, P, ns, res, th, priordict, args.bruteforce) if args.bruteforce else None
        calls = float(res['ncall'].sum())
        print(json.dumps(dict(setting=setting, stars=nstars, seconds=secs, stars_per_hour=nstars / secs * 3600.0,
                              calls_per_star=calls / nstars, calls_per_iteration=calls / float(res['niter'].sum()),
                              loglike_evals_per_sec=calls / secs, rounds=int(res['rounds']),
                              iterations_median=float(np.median(res['niter'])),
                              converged_frac=float(res['converged'].mean()), vs_bruteforce=acc)), flush=True)


if __name__ == '__main__':
    main()
