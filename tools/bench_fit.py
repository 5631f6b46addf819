#!/usr/bin/env python
"""Config 4 (BASELINE.json configs[3]): one full nested-sampling fit of one synthetic star
(10 bands + Gaussian parallax prior, demo sampler settings: static, rwalk, 150 live points,
walks 5, dlogz 0.01) with the GPU proposal-queue sampler, next to the cost of the same number
of sequential likelihood calls on one CPU core with the per-point oracle (dynesty itself is
not installed, so the CPU figure is calls x measured per-call time, stated as such).
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
    ap.add_argument('--walks', type=int, default=5)
    ap.add_argument('--nlive', type=int, default=150)
    ap.add_argument('--dlogz', type=float, default=0.01)
    ap.add_argument('--repeat', type=int, default=5)
    ap.add_argument('--cpu-calls', type=int, default=2000)
    args = ap.parse_args()
    import torch
    from minesweeper_b200 import synth
    from minesweeper_b200.likelihood import likelihood
    from minesweeper_b200.prior import prior
    from minesweeper_b200.sampler import QueuedNestedSampler
    from bench_catalog import NAMES, ON, catalog_priordict, make_catalog_sampler
    mist = synth.make_mist(seed=0)
    phot = synth.make_phot_net(h=128, seed=1)
    fitpars = [NAMES, {k: (k in ON) for k in NAMES}]
    fitargs = dict(ageweight=True, mistdif=True, fixedpars={}, MISTpath=mist, photANNpath=phot,
                   obs_phot=synth.demo_obs_phot())
    runbools = [False, True, False, False, False]
    L = likelihood(fitargs, fitpars, runbools, precision='tc', verbose=False)
    P = prior(fitargs, catalog_priordict(808.0), fitpars, runbools, likeobj=L)
    dev = L.engine.device
    _, truth = make_catalog_sampler(L, P, 1, dev, args.nlive, args.walks, args.dlogz, 100000, 0)   # registers star 0
    plx = torch.tensor([[1000.0 / float(truth[0, L.fitpars_i.index('Dist')]), 0.1]], dtype=torch.float64, device=dev)
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
    from oracle.like_port import LikelihoodPort
    ref = LikelihoodPort(fitargs, fitpars, runbools)
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
        logz=float(res['logz'][0]), logzerr=float(res['logzerr'][0]),
        cpu_us_per_call_1core=cpu_call * 1e6, sequential_calls_equivalent=seq_calls,
        cpu_seconds_estimate_1core=seq_calls * cpu_call,
        note='CPU figure = (iterations x walks + nlive) sequential calls x measured per-call time of the '
             'per-point oracle likelihood on one host core; the prior and sampler overhead of dynesty are not included')))


if __name__ == '__main__':
    main()
