#!/usr/bin/env python
"""Config 5 (BASELINE.json configs[4]): a synthetic catalog of stars (10 magnitudes + a Gaia-like
parallax prior each) fitted with the lock-step batched nested sampler; reports stars fitted per
hour.  Under torchrun the catalog is sharded over the ranks (contiguous star blocks, no data-path
collective) and the per-star summaries are gathered once at the end.

    python tools/bench_catalog.py --stars 20000 [--nlive 150 --walks 5 --dlogz 0.01]
    torchrun --nproc-per-node N tools/bench_catalog.py --stars 100000
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

NAMES = ['Teff', 'log(g)', '[Fe/H]', '[a/Fe]', 'Vrad', 'Vrot', 'Vmic', 'Inst_R', 'log(R)', 'Dist', 'Av',
         'log(A)', 'EEP', 'initial_[Fe/H]', 'initial_[a/Fe]', 'initial_Mass']
ON = {'Dist', 'Av', 'EEP', 'initial_[Fe/H]', 'initial_[a/Fe]', 'initial_Mass'}


def make_catalog_sampler(L, P, S, dev, nlive, walks, dlogz, maxiter, seed, queued='auto', sample='rwalk'):
    """Synthetic catalog of S stars on this device (truths from the prior, observations from the
    model itself, Gaia-like parallaxes) registered with the engine, and a sampler over it.
    Returns (sampler, truth_theta)."""
    import torch
    from minesweeper_b200.sampler import BatchedNestedSampler, QueuedNestedSampler
    gen = torch.Generator(device=dev).manual_seed(1000 + seed)
    truth = None
    for _ in range(20):                                            # truths with a valid MIST prediction
        u = torch.rand(2 * S, L.ndim, generator=gen, device=dev, dtype=torch.float64)
        th = P.priortrans_batch(u)
        lnl, der, _ = L.lnlike_batch(th)
        good = torch.isfinite(lnl)
        th, der = th[good], der[good]
        truth = (th, der) if truth is None else (torch.cat([truth[0], th]), torch.cat([truth[1], der]))
        if len(truth[0]) >= S:
            break
    th, der = truth[0][:S], truth[1][:S]
    cols = {n: i for i, n in enumerate(L.fitpars_i)}
    pp = torch.stack([10 ** der[:, 8], der[:, 11], der[:, 9], der[:, 10], der[:, 6], th[:, cols['Dist']],
                      th[:, cols['Av']]], 1)
    mags = L.engine.phot_sed(pp)
    err = torch.full_like(mags, 0.02)
    obs = mags + err * torch.randn(mags.shape, generator=gen, device=dev, dtype=torch.float64)
    plx = 1000.0 / th[:, cols['Dist']]
    plx_err = 0.05 + 0.02 * plx
    plx_obs = plx + plx_err * torch.randn(S, generator=gen, device=dev, dtype=torch.float64)
    L.engine.set_stars_phot(obs.cpu().numpy(), err.cpu().numpy())
    star_plx = torch.stack([plx_obs, plx_err], 1)
    if queued:       # bookkeeping in CUDA kernels (csrc/sampler.cu), proposal queue of `queued` walkers per star
        ns = QueuedNestedSampler(L, P, nstars=S, nlive=nlive, walks=walks, dlogz=dlogz,
                                 queue='auto' if (queued == 'auto' or queued < 0) else queued,
                                 maxiter=maxiter, seed=5 + seed, star_plx=star_plx, sample=sample)
    else:            # torch-tensor bookkeeping (reference implementation of the same sampler)
        ns = BatchedNestedSampler(L, P, nstars=S, nlive=nlive, walks=walks, dlogz=dlogz, maxiter=maxiter,
                                  seed=5 + seed, star_plx=star_plx)
    ns.catalog_star_plx = star_plx
    return ns, th


def bruteforce_check(L, P, res, star_plx, stars, priordict, nsamp=1 << 22, nsig=6.0, seed=99):
    """Ground truth for a subsample of fitted stars, independent of the sampler: evidence and posterior moments
    by brute-force quadrature through the SAME kernels -- uniform draws in a box of +-nsig posterior standard
    deviations around the posterior (clipped to the prior box, re-centred once on its own estimate), weights
    exp(lnprob).  All volume priors of the catalog are uniform, so
        Z = (V_box / V_prior) <exp(lnprob)>_box      (ln-prior terms and the per-star parallax prior are in lnprob).
    Returns per-star arrays: lnZ_bf, mean_bf[nd], std_bf[nd], ess."""
    import torch
    eng = L.engine
    dev = eng.device
    nd = L.ndim
    lo_p = np.array([min(priordict[n]['pv_uniform']) for n in L.fitpars_i], dtype=np.float64)
    hi_p = np.array([max(priordict[n]['pv_uniform']) for n in L.fitpars_i], dtype=np.float64)
    gen = torch.Generator(device=dev).manual_seed(seed)
    out = dict(lnz=[], mean=[], std=[], ess=[])
    for s_ in stars:
        m, sd = res['mean'][s_, :nd].copy(), np.maximum(res['std'][s_, :nd], 1e-9)
        for _ in range(2):
            lo = np.maximum(lo_p, m - nsig * sd)
            hi = np.minimum(hi_p, m + nsig * sd)
            u = torch.rand(nsamp, nd, generator=gen, device=dev, dtype=torch.float64)
            th = torch.as_tensor(lo, device=dev) + u * torch.as_tensor(hi - lo, device=dev)
            slot = torch.full((nsamp,), int(s_), dtype=torch.int32, device=dev)
            lnl, der, _ = eng.lnlike_batch(th, L._colmap, L._fixed, L._flags, star_slot=slot, want_derived=True)
            lp = P.lnprob_batch(th, lnl, der, star_slot=slot, star_plx=star_plx)
            lp = torch.where(torch.isfinite(lp), lp, torch.full_like(lp, -float('inf')))
            mx = lp.max()
            w = torch.exp(lp - mx)
            wsum = w.sum()
            lnz = float(mx + torch.log(wsum) - np.log(nsamp) + np.sum(np.log(hi - lo)) - np.sum(np.log(hi_p - lo_p)))
            mean = ((w[:, None] * th).sum(0) / wsum)
            var = (w[:, None] * (th - mean) ** 2).sum(0) / wsum
            ess = float(wsum ** 2 / (w ** 2).sum())
            m, sd = mean.cpu().numpy(), np.maximum(np.sqrt(var.cpu().numpy()), 1e-12)
        out['lnz'].append(lnz); out['mean'].append(m); out['std'].append(sd); out['ess'].append(ess)
    return {k: np.array(v) for k, v in out.items()}


def sampler_accuracy(L, res, bf, stars):
    """Nested-sampling results against the brute-force ground truth of ``bruteforce_check``."""
    nd = L.ndim
    dz = res['logz'][stars] - bf['lnz']
    zerr = np.maximum(res['logzerr'][stars], 1e-3)
    bias = (res['mean'][stars, :nd] - bf['mean']) / bf['std']
    ratio = res['std'][stars, :nd] / bf['std']
    return dict(stars=int(len(stars)), bruteforce_ess_median=float(np.median(bf['ess'])),
                dlnz_mean=float(dz.mean()), dlnz_rms=float(np.sqrt((dz ** 2).mean())),
                dlnz_over_err_rms=float(np.sqrt(((dz / zerr) ** 2).mean())),
                dlnz_median=float(np.median(dz)),
                mean_bias_rms_in_sigma={n: float(np.sqrt((bias[:, i] ** 2).mean())) for i, n in enumerate(L.fitpars_i)},
                mean_bias_median_abs_in_sigma={n: float(np.median(np.abs(bias[:, i]))) for i, n in enumerate(L.fitpars_i)},
                frac_stars_bias_above_1sigma=float(np.mean(np.any(np.abs(bias) > 1.0, axis=1))),
                std_ratio_median={n: float(np.median(ratio[:, i])) for i, n in enumerate(L.fitpars_i)})


# Acceptance gate of a sampler setting, against the brute-force ground truth of a subsample.  Medians over the stars
# (robust against the few stars per hundred whose minor posterior mode dies out early, which dominate any RMS):
# posterior widths within 20 %, posterior means within half a posterior sigma, evidence within 0.5.
# What the measurements behind these numbers show (B200, nlive = 150, 4 queued walkers; tests/test_sampler.py, bench.py):
#  * walks = 5 (demo/etc/samplerinfo.dat): widths 0.70-0.93 of the exact ones, RMS mean bias up to 1.2 sigma -> FAILS;
#  * walks = 25 (dynesty's default): widths 0.86-0.98, RMS mean bias 0.3-1.3 sigma (median 0.2-0.4) -> passes; walks = 50
#    is no better and nlive = 1000 brings the RMS bias to 0.1-0.3 sigma: beyond 25 steps the error is the statistical
#    noise of 150 live points on broad, multimodal EEP / mass posteriors (it scales as nlive^-1/2), not a walk bias;
#  * (mean - truth) / std has RMS 1.0 (1.5 for EEP) for the EXACT posteriors (`truth_z_rms_bruteforce`) but 1.2-2.2 for
#    the sampler's (`posterior_z_rms`): stars whose minor mode died out early get too narrow a posterior
#    (`frac_stars_bias_above_1sigma`).  All of these are reported; the gate uses the medians of the direct comparison.
GATE = dict(std_ratio_min=0.80, std_ratio_max=1.25, mean_bias_median_abs_max=0.50, dlnz_median_abs_max=0.50)
GATE_WALKS = 25          # shortest random walk (dynesty's default) that passes the gate; the demo's walks = 5 does not


def gate(acc):
    ok = all(GATE['std_ratio_min'] <= v <= GATE['std_ratio_max'] for v in acc['std_ratio_median'].values())
    ok = ok and all(v <= GATE['mean_bias_median_abs_max'] for v in acc['mean_bias_median_abs_in_sigma'].values())
    ok = ok and abs(acc['dlnz_median']) <= GATE['dlnz_median_abs_max']
    return bool(ok)


def accuracy_block(L, P, ns, res, th_truth, priordict, nstars=64):
    """vs_bruteforce + gate + the truth z-scores of the sampler's and of the exact posteriors (subsample)."""
    S = len(res['logz'])
    nd = L.ndim
    pick = np.linspace(0, S - 1, min(nstars, S)).astype(int)
    bf = bruteforce_check(L, P, res, ns.catalog_star_plx, pick, priordict)
    acc = sampler_accuracy(L, res, bf, pick)
    tr = th_truth.cpu().numpy()
    z_all = (res['mean'][:, :nd] - tr) / np.maximum(res['std'][:, :nd], 1e-12)
    z_bf = (bf['mean'] - tr[pick]) / bf['std']
    acc['posterior_z_rms'] = {n: float(np.sqrt(np.mean(z_all[:, i] ** 2))) for i, n in enumerate(L.fitpars_i)}
    acc['truth_z_rms_bruteforce'] = {n: float(np.sqrt(np.mean(z_bf[:, i] ** 2))) for i, n in enumerate(L.fitpars_i)}
    acc['gate'] = dict(GATE, passed=gate(acc))
    return acc


def catalog_priordict(eep_hi=808.0):
    return {'EEP': {'pv_uniform': [202, eep_hi]}, 'initial_Mass': {'pv_uniform': [0.5, 1.25]},
            'initial_[Fe/H]': {'pv_uniform': [-2.0, 0.4]}, 'initial_[a/Fe]': {'pv_uniform': [-0.2, 0.6]},
            'Dist': {'pv_uniform': [5.0, 2000.0]}, 'Av': {'pv_uniform': [0.0, 0.5]},
            'IMF': {'IMF_type': 'Kroupa'}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--stars', type=int, default=20000)
    ap.add_argument('--nlive', type=int, default=150)
    ap.add_argument('--walks', type=int, default=GATE_WALKS,
                    help='random-walk steps per proposal (demo/etc/samplerinfo.dat: 5, which fails the accuracy gate)')
    ap.add_argument('--dlogz', type=float, default=0.01)
    ap.add_argument('--maxiter', type=int, default=20000)
    ap.add_argument('--small-grid', action='store_true')
    ap.add_argument('--eager', action='store_true', help='use the eager sampler loop instead of the CUDA-graph one')
    ap.add_argument('--compact-frac', type=float, default=0.85,
                    help='shrink the working set (and re-capture the round graph) once this fraction of it is still running')
    ap.add_argument('--check-every', type=int, default=16)
    ap.add_argument('--bruteforce', type=int, default=0, help='check this many stars against brute-force quadrature')
    ap.add_argument('--queued', type=int, default=-1,
                    help='walkers per star of the kernel-bookkeeping sampler; -1 = sized from the number of stars '
                         '(sampler.auto_queue); 0 = torch-tensor sampler')
    ap.add_argument('--sample', default='rwalk', choices=['rwalk', 'unif'],
                    help="proposal of the queued sampler: random walk of --walks steps, or dynesty's uniform draw from "
                         "the bounding ellipsoid of the live points (one likelihood call per proposal)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    from minesweeper_b200 import synth
    from minesweeper_b200.likelihood import likelihood
    from minesweeper_b200.prior import prior
    from minesweeper_b200.distributed import shard_bounds, gather_rows
    mist = synth.make_mist(ne=60, nm=24, nf=7, na=5, seed=0) if args.small_grid else synth.make_mist(seed=0)
    phot = synth.make_phot_net(h=128, seed=1)
    eep_hi = 261.0 if args.small_grid else 808.0
    fitpars = [NAMES, {k: (k in ON) for k in NAMES}]
    fitargs = dict(ageweight=True, mistdif=True, fixedpars={}, MISTpath=mist, photANNpath=phot,
                   obs_phot=synth.demo_obs_phot())
    runbools = [False, True, False, False, False]
    priordict = catalog_priordict(eep_hi)
    L = likelihood(fitargs, fitpars, runbools, precision='tc', device=local, verbose=False)
    P = prior(fitargs, priordict, fitpars, runbools, likeobj=L)
    # this rank's block of the catalog
    lo, hi = shard_bounds(args.stars, rank, world)
    S = hi - lo
    ns, th = make_catalog_sampler(L, P, S, dev, args.nlive, args.walks, args.dlogz, args.maxiter, rank,
                                  queued=args.queued, sample=args.sample)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = L.engine.launch_count()
    t0 = time.perf_counter()
    if args.queued:
        res = ns.run(check_every=args.check_every, compact_frac=args.compact_frac)
    else:
        res = ns.run() if args.eager else ns.run_graphed()
    torch.cuda.synchronize()
    el = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    summary = torch.from_numpy(np.concatenate([res['logz'][:, None], res['logzerr'][:, None], res['mean'],
                                               res['std']], 1)).to(dev)
    if world > 1:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
        summary = gather_rows(summary, args.stars)                 # the only collective: per-star summaries
    nd = L.ndim
    zs = (res['mean'][:, :nd] - th.cpu().numpy()) / np.maximum(res['std'][:, :nd], 1e-12)
    calls = torch.tensor([float(res['ncall'].sum())], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(calls)
    acc = None
    if args.bruteforce and rank == 0:
        acc = accuracy_block(L, P, ns, res, th, priordict, args.bruteforce)
    if rank == 0:
        secs = float(el.item())
        print(json.dumps(dict(
            metric='stars_fitted_per_hour', value=args.stars / secs * 3600.0, unit='stars/hour', n_gpus=world,
            stars=args.stars, seconds=secs, nlive=args.nlive, walks=args.walks, dlogz=args.dlogz, sample=args.sample,
            iterations_median=float(np.median(res['niter'])), iterations_max=int(res['niter'].max()),
            converged_frac=float(res['converged'].mean()), mode=('queued-%d' % ns.Q) if args.queued else ('eager' if args.eager else 'cuda-graph'),
            loglike_calls=float(calls.item()), loglike_evals_per_sec=float(calls.item()) / secs,
            calls_per_star=float(calls.item()) / args.stars,
            rank0_posterior_z_rms={n: float(np.sqrt(np.mean(zs[:, i] ** 2))) for i, n in enumerate(L.fitpars_i)},
            vs_bruteforce=acc,
            rank0_kernel_launches=int(L.engine.launch_count() - l0),
            summary_shape=list(summary.shape))))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
