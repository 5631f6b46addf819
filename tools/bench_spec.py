#!/usr/bin/env python
"""Config 3 (BASELINE.json configs[2]): spectrum + photometry log-likelihood, spectral ANN
4-300-300-10240 px with vsini + Doppler + R smoothing, batch 65 536, on one B200.
Prints a JSON line with evals/s and the per-kernel device times (ms_profile_* events).

    python tools/bench_spec.py [--batch 65536] [--nobs 1536] [--precision f32] [--steps 3]
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
         'log(A)', 'EEP', 'initial_[Fe/H]', 'initial_[a/Fe]', 'initial_Mass', 'pc_0', 'pc_1', 'pc_2', 'pc_3']
ON = {'Vrad', 'Vrot', 'Dist', 'Av', 'EEP', 'initial_[Fe/H]', 'initial_[a/Fe]', 'initial_Mass',
      'pc_0', 'pc_1', 'pc_2', 'pc_3'}


def make_spec_workload(mist, phot, batch, nobs=1536, npix=10240, seed=0):
    """configs[2] of BASELINE.json (SURVEY.md section 8d "Config 3"): (fitargs, fitpars, runbools, theta[batch, 12])."""
    from minesweeper_b200 import synth
    spec = synth.make_spec_net(npix=npix, h=300, seed=2)
    fitpars = [NAMES, {n: (n in ON) for n in NAMES}]
    obs = synth.make_obs_spectrum(spec, n_obs=nobs, seed=4)
    fitargs = dict(ageweight=True, mistdif=True, fixedpars={'Inst_R': 35000.0}, MISTpath=mist,
                   photANNpath=phot, specANNpath=spec, NNtype='YST1', obs_phot=synth.demo_obs_phot(),
                   obs_wave_fit=obs['obs_wave'], obs_flux_fit=obs['obs_flux'], obs_eflux_fit=obs['obs_eflux'])
    runbools = [True, True, True, False, False]
    rng = np.random.default_rng(4 + seed)
    B = batch
    small = len(mist['eep']) < 100
    box = dict(Dist=(1.0, 100.0), Av=(0.0, 0.1), EEP=(202.0, 261.0), feh=(-4.0, 0.0), afe=(-0.2, 0.6),
               mass=(0.5, 1.25)) if small else None
    base = synth.draw_theta_phot(B, seed=3 + seed, box=box)            # Dist, Av, EEP, feh, afe, mass
    th = np.stack([rng.uniform(-5, 5, B), rng.uniform(0, 10, B), base[:, 0], base[:, 1], base[:, 2], base[:, 3],
                   base[:, 4], base[:, 5], rng.normal(1.0, 0.05, B), rng.normal(0, 0.02, B),
                   rng.normal(0, 0.01, B), rng.normal(0, 0.005, B)], 1)
    return fitargs, fitpars, runbools, th


def run_spec_bench(mist, phot, batch=65536, nobs=1536, npix=10240, precision='tc', steps=3, device=None,
                   cpu_points=0, seed=0, spec_fast=1, parity_ref=None):
    """Times spectrum + photometry lnL over `batch` points resident on the GPU; returns a dict.  parity_ref:
    (rows, lnl_ref, tol) from the CPU leg -- the first rows of this batch evaluated by the per-point reference."""
    import torch
    from minesweeper_b200.likelihood import likelihood
    fitargs, fitpars, runbools, th = make_spec_workload(mist, phot, batch, nobs, npix, seed)
    names, on = NAMES, ON
    B = batch
    cpu = None
    if cpu_points:
        from oracle.like_port import LikelihoodPort
        ref = LikelihoodPort(fitargs, fitpars, runbools)
        t0 = time.perf_counter()
        for i in range(cpu_points):
            ref.lnlikefn(list(th[i]))
        cpu = cpu_points / (time.perf_counter() - t0)
    L = likelihood(fitargs, fitpars, runbools, precision=precision, device=device, verbose=False)
    assert L.fitpars_i == [n for n in names if n in on]
    eng = L.engine
    eng.set_spec_fast(spec_fast)
    theta = torch.from_numpy(th).to(eng.device)
    out = torch.empty(B, dtype=torch.float64, device=eng.device)
    for _ in range(2):
        eng.lnlike_batch(theta, L._colmap, L._fixed, L._flags, want_derived=False, out=out)
    torch.cuda.synchronize(eng.device)
    eng.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        eng.lnlike_batch(theta, L._colmap, L._fixed, L._flags, want_derived=False, out=out)
    e1.record()
    torch.cuda.synchronize(eng.device)
    ms = e0.elapsed_time(e1) / steps
    prof = eng.profile_read()
    eng.profile(False)
    parity = None
    if parity_ref is not None:
        nrow, lref, tol = parity_ref
        got = out[:nrow].cpu().numpy()
        fin = np.isfinite(lref)
        assert np.array_equal(np.isneginf(got), np.isneginf(lref)), 'spec parity: -inf rows differ'
        err = np.abs(got[fin] - lref[fin])
        parity = dict(rows=int(nrow), max_dlnL=float(err.max()), max_dlnL_over_bound=float((err / tol[fin]).max()))
        assert np.all(err <= tol[fin]), 'spec parity: lnL outside the bound (%g x)' % (err / tol[fin]).max()
    flop = 2.0 * (4 * 300 + 300 * 300 + 300 * npix)
    res = dict(metric='spec_phot_loglike_evals_per_sec', value=B / (ms * 1e-3), unit='evals/s', ms_per_step=ms,
               batch=B, n_obs=nobs, npix=npix, precision=precision,
               kernel_ms={k: v[0] / steps for k, v in prof.items() if v[1]},
               spec_mlp_tflops=flop * B / (prof['spec_mlp'][0] / steps * 1e-3) / 1e12,
               finite=int(torch.isfinite(out).sum().item()), cpu_evals_per_sec_1core=cpu, parity_check=parity,
               flop_per_eval=flop,
               workload='configs[2]: spectrum + photometry loglike, spectral ANN 4-300-300-%d px, vsini + Doppler + '
                        'R smoothing, %d observed pixels, batch %d' % (npix, nobs, B))
    eng.close()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=65536)
    ap.add_argument('--nobs', type=int, default=1536)
    ap.add_argument('--npix', type=int, default=10240)
    ap.add_argument('--precision', default='f32')
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--spec-fast', type=int, default=1, help='0 generic FFT kernel for every row, 1 fast broadening kernel')
    ap.add_argument('--cpu-points', type=int, default=0, help='also time the per-point oracle on N points')
    args = ap.parse_args()
    from minesweeper_b200 import synth
    mist = synth.make_mist(seed=0)
    phot = synth.make_phot_net(h=128, seed=1)
    print(json.dumps(run_spec_bench(mist, phot, args.batch, args.nobs, args.npix, args.precision, args.steps,
                                    cpu_points=args.cpu_points, spec_fast=args.spec_fast)))


if __name__ == '__main__':
    main()
