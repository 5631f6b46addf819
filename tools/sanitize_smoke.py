"""Small end-to-end pass over every kernel family and precision mode (seconds on a B200), sized so it can also
run under `compute-sanitizer --tool memcheck` where that tool is available.  The fused photometry kernel needs
>= 4 tile pairs per SM, so that part uses 148 * 256 * 4 points on a small grid."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from minesweeper_b200 import synth                              # noqa: E402
from minesweeper_b200.likelihood import likelihood              # noqa: E402


def main():
    mist = synth.make_mist(ne=40, nm=12, nf=5, na=3, seed=0)
    phot = synth.make_phot_net(h=128, seed=1)
    spec = synth.make_spec_net(npix=2048, h=300, seed=2)
    names = ['Teff', 'log(g)', '[Fe/H]', '[a/Fe]', 'Vrad', 'Vrot', 'Vmic', 'Inst_R', 'log(R)', 'Dist', 'Av',
             'log(A)', 'EEP', 'initial_[Fe/H]', 'initial_[a/Fe]', 'initial_Mass', 'pc_0', 'pc_1', 'pc_2', 'pc_3']
    box = dict(Dist=(1.0, 100.0), Av=(0.0, 0.1), EEP=(202.0, 240.0), feh=(-4.0, -3.0), afe=(-0.2, 0.2),
               mass=(0.25, 30.0))
    on_p = {'Dist', 'Av', 'EEP', 'initial_[Fe/H]', 'initial_[a/Fe]', 'initial_Mass'}
    fitargs = dict(ageweight=True, mistdif=True, fixedpars={}, MISTpath=mist, photANNpath=phot,
                   obs_phot=synth.demo_obs_phot())
    for prec in ('f64', 'f32', 'tc', 'tc16'):
        L = likelihood(fitargs, [names[:16], {k: (k in on_p) for k in names[:16]}], [False, True, False, False, False],
                       precision=prec, verbose=False)
        for n in (1, 300, 148 * 256 * 4 + 3):
            if prec in ('f64', 'f32') and n > 300:
                continue
            th = synth.draw_theta_phot(n, seed=3, frac_out=0.05, box=box)
            lnl, der, _ = L.lnlike_batch(th)
            assert lnl.shape[0] == n and der.shape == (n, 13)
        L.engine.close()
    obs = synth.make_obs_spectrum(spec, n_obs=256, seed=4)
    on_s = on_p | {'Vrad', 'Vrot', 'pc_0', 'pc_1', 'pc_2', 'pc_3'}
    fa = dict(fitargs, fixedpars={'Inst_R': 35000.0}, specANNpath=spec, NNtype='YST1', obs_wave_fit=obs['obs_wave'],
              obs_flux_fit=obs['obs_flux'], obs_eflux_fit=obs['obs_eflux'])
    rng = np.random.default_rng(5)
    for prec in ('f32', 'tc'):
        L = likelihood(fa, [names, {k: (k in on_s) for k in names}], [True, True, True, False, False], precision=prec,
                       verbose=False)
        base = synth.draw_theta_phot(64, seed=6, box=box)
        th = np.stack([rng.uniform(-5, 5, 64), rng.uniform(0, 10, 64), base[:, 0], base[:, 1], base[:, 2], base[:, 3],
                       base[:, 4], base[:, 5], rng.normal(1.0, 0.05, 64), rng.normal(0, 0.02, 64),
                       rng.normal(0, 0.01, 64), rng.normal(0, 0.005, 64)], 1)
        lnl = L.lnlike_batch(th)[0]
        assert lnl.shape[0] == 64
        L.engine.close()
    torch.cuda.synchronize()
    print('sanitize_smoke: all kernel families ran')


if __name__ == '__main__':
    main()
