"""Lock-step batched nested sampler: evidence and posterior moments against brute-force prior
Monte-Carlo integration through the same kernels, and agreement between independent replicas."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

NAMES = ['Teff', 'log(g)', '[Fe/H]', '[a/Fe]', 'Vrad', 'Vrot', 'Vmic', 'Inst_R', 'log(R)', 'Dist', 'Av',
         'log(A)', 'EEP', 'initial_[Fe/H]', 'initial_[a/Fe]', 'initial_Mass']
ON = {'Dist', 'Av', 'EEP', 'initial_[Fe/H]', 'initial_[a/Fe]', 'initial_Mass'}


def _problem():
    from minesweeper_b200 import synth
    from oracle.payne_port import FastPayneSEDPredict
    from oracle import mist_port
    mist = synth.make_mist(ne=60, nm=24, nf=7, na=5, seed=10, frac_missing=0.0, frac_trunc=0.0)
    phot = synth.make_phot_net(h=128, seed=35)
    truth = dict(EEP=230.3, mass=0.93, feh=-0.4, afe=0.15, dist=14.0, av=0.03)
    _, pred, ok = mist_port.interp_dense(mist, [truth['EEP']], [truth['mass']], [truth['feh']], [truth['afe']])
    assert ok[0]
    p = pred[0]
    pp = [[10 ** p[4], p[7], p[5], p[6], p[2], truth['dist'], truth['av']]]
    mags = FastPayneSEDPredict(nnpath=phot).sed_batch(pp)[0]
    rng = np.random.default_rng(2)
    err = 0.08
    obs = {b: [float(m + err * rng.standard_normal()), err] for b, m in zip(synth.DEMO_BANDS, mags)}
    fitpars = [NAMES, {k: (k in ON) for k in NAMES}]
    fitargs = dict(ageweight=True, mistdif=True, fixedpars={}, MISTpath=mist, photANNpath=phot, obs_phot=obs)
    runbools = [False, True, False, False, False]
    priordict = {'EEP': {'pv_uniform': [202, 261]}, 'initial_Mass': {'pv_uniform': [0.5, 1.25]},
                 'initial_[Fe/H]': {'pv_uniform': [-1.0, 0.2]}, 'initial_[a/Fe]': {'pv_uniform': [-0.2, 0.6]},
                 'Dist': {'pv_uniform': [8.0, 20.0]}, 'Av': {'pv_uniform': [0.0, 0.1]},
                 'Parallax': {'gaussian': [1000.0 / 14.0, 3.0]}, 'IMF': {'IMF_type': 'Kroupa'}}
    return fitargs, fitpars, runbools, priordict


def test_nested_sampler_matches_prior_monte_carlo():
    from minesweeper_b200.likelihood import likelihood
    from minesweeper_b200.prior import prior
    from minesweeper_b200.sampler import BatchedNestedSampler
    fitargs, fitpars, runbools, priordict = _problem()
    L = likelihood(fitargs, fitpars, runbools, precision='tc', verbose=False)
    P = prior(fitargs, priordict, fitpars, runbools, likeobj=L)
    # brute force: Z = E_prior[exp(lnprob)], posterior mean = weighted mean
    gen = torch.Generator(device=L.engine.device).manual_seed(7)
    lw, ths = [], []
    for _ in range(8):
        u = torch.rand(1 << 18, L.ndim, generator=gen, device=L.engine.device, dtype=torch.float64)
        th = P.priortrans_batch(u)
        lnl, der, _ = L.lnlike_batch(th)
        lw.append(P.lnprob_batch(th, lnl, der)); ths.append(th)
    lw, ths = torch.cat(lw), torch.cat(ths)
    logz_mc = float(torch.logsumexp(lw, 0) - np.log(len(lw)))
    w = torch.exp(lw - lw.max()); w = w / w.sum()
    mean_mc = (w[:, None] * ths).sum(0).cpu().numpy()
    std_mc = torch.sqrt((w[:, None] * (ths - torch.as_tensor(mean_mc, device=ths.device)) ** 2).sum(0)).cpu().numpy()
    ess = float(1.0 / (w ** 2).sum())
    assert ess > 2000, ess
    # nested sampling: the same star in 6 replicas run in lock step
    S = 6
    L.engine.set_stars_phot(np.tile([[fitargs['obs_phot'][b][0] for b in L.engine.bands]], (S, 1)),
                            np.tile([[fitargs['obs_phot'][b][1] for b in L.engine.bands]], (S, 1)))
    ns = BatchedNestedSampler(L, P, nstars=S, nlive=150, walks=10, dlogz=0.05, seed=11, store_samples=True)
    res = ns.run()
    assert res['converged'].all()
    assert (res['niter'] > 300).all() and (res['ncall'] > res['niter'] * 10).all()
    # evidence: replicas agree with brute force within a few sigma of the NS error estimate
    err = np.maximum(res['logzerr'], 0.05)
    assert np.all(np.abs(res['logz'] - logz_mc) < 5 * err + 0.15), (res['logz'], logz_mc, err)
    assert np.std(res['logz']) < 4 * err.mean() + 0.1
    # posterior means of the sampled parameters
    nd = L.ndim
    m_ns = res['mean'][:, :nd].mean(0)
    assert np.all(np.abs(m_ns - mean_mc) < 0.35 * std_mc + 1e-3), (m_ns, mean_mc, std_mc)
    s_ns = res['std'][:, :nd].mean(0)
    assert np.all(np.abs(s_ns / std_mc - 1.0) < 0.35), (s_ns, std_mc)
    # dead points were recorded in the samples-file layout
    from minesweeper_b200.sampler import write_samples_file
    import tempfile, os
    with tempfile.TemporaryDirectory() as d:
        f = os.path.join(d, 's.dat')
        write_samples_file(f, res, star=2)
        lines = open(f).read().splitlines()
        assert lines[0].startswith('Iter Dist Av EEP') and len(lines) == res['niter'][2] + 1
        assert np.isfinite(float(lines[-1].split()[-2]))            # running log(z), recomputed for this record layout


def test_graphed_sampler_agrees_with_eager():
    """The CUDA-graph replay form gives the same evidence and posterior (statistically) as the
    eager loop."""
    from minesweeper_b200.likelihood import likelihood
    from minesweeper_b200.prior import prior
    from minesweeper_b200.sampler import BatchedNestedSampler
    fitargs, fitpars, runbools, priordict = _problem()
    L = likelihood(fitargs, fitpars, runbools, precision='tc', verbose=False)
    P = prior(fitargs, priordict, fitpars, runbools, likeobj=L)
    S = 8
    L.engine.set_stars_phot(np.tile([[fitargs['obs_phot'][b][0] for b in L.engine.bands]], (S, 1)),
                            np.tile([[fitargs['obs_phot'][b][1] for b in L.engine.bands]], (S, 1)))
    torch.manual_seed(3)
    eager = BatchedNestedSampler(L, P, nstars=S, nlive=150, walks=10, dlogz=0.05, seed=21).run()
    graphed = BatchedNestedSampler(L, P, nstars=S, nlive=150, walks=10, dlogz=0.05, seed=22).run_graphed(check_every=32)
    assert graphed['converged'].all() and eager['converged'].all()
    err = max(eager['logzerr'].mean(), 0.05)
    assert abs(graphed['logz'].mean() - eager['logz'].mean()) < 3 * err / np.sqrt(S) + 0.15
    nd = L.ndim
    sd = eager['std'][:, :nd].mean(0)
    assert np.all(np.abs(graphed['mean'][:, :nd].mean(0) - eager['mean'][:, :nd].mean(0)) < 0.3 * sd + 1e-3)
    assert np.all(np.abs(graphed['std'][:, :nd].mean(0) / sd - 1.0) < 0.3)


def test_queued_sampler_single_star_matches_monte_carlo():
    """Proposal-queue sampler (bookkeeping in CUDA kernels, one CTA per star) on a single star
    and on replicas: evidence vs brute-force prior Monte-Carlo, posterior moments, dead points."""
    from minesweeper_b200.likelihood import likelihood
    from minesweeper_b200.prior import prior
    from minesweeper_b200.sampler import QueuedNestedSampler
    fitargs, fitpars, runbools, priordict = _problem()
    L = likelihood(fitargs, fitpars, runbools, precision='tc', verbose=False)
    P = prior(fitargs, priordict, fitpars, runbools, likeobj=L)
    gen = torch.Generator(device=L.engine.device).manual_seed(7)
    lw, ths = [], []
    for _ in range(8):
        u = torch.rand(1 << 18, L.ndim, generator=gen, device=L.engine.device, dtype=torch.float64)
        th = P.priortrans_batch(u)
        lnl, der, _ = L.lnlike_batch(th)
        lw.append(P.lnprob_batch(th, lnl, der)); ths.append(th)
    lw, ths = torch.cat(lw), torch.cat(ths)
    logz_mc = float(torch.logsumexp(lw, 0) - np.log(len(lw)))
    w = torch.exp(lw - lw.max()); w = w / w.sum()
    mean_mc = (w[:, None] * ths).sum(0).cpu().numpy()
    std_mc = torch.sqrt((w[:, None] * (ths - torch.as_tensor(mean_mc, device=ths.device)) ** 2).sum(0)).cpu().numpy()
    nd = L.ndim
    # one star, 32 queued walkers
    one = QueuedNestedSampler(L, P, nstars=1, nlive=150, walks=10, queue=32, dlogz=0.05, seed=3, store_samples=True).run()
    assert one['converged'].all() and one['niter'][0] > 300
    assert abs(one['logz'][0] - logz_mc) < 5 * max(one['logzerr'][0], 0.05) + 0.2, (one['logz'], logz_mc)
    assert one['rounds'] < one['niter'][0] / 4                      # the queue cut the sequential depth
    dead = one['dead'][0]
    assert dead.shape == (one['niter'][0], nd + 13 + 7)
    assert np.all(np.diff(dead[:, nd + 13]) >= 0)                  # dead-point lnprob is non-decreasing
    # ---- samples file in the reference's layout (fitstar.py:235-242, 392-425, 466-501): header = parsdict keys,
    # real running h / nc / log(z) / delta(log(z)), final live points appended; readable the way demo/postproc.py
    # reads it: Prob = exp(log(wt) - log(z)[-1]) sums to one
    import os, tempfile
    from minesweeper_b200.likelihood import parsdict_keys
    from minesweeper_b200.sampler import write_samples_file
    from minesweeper_b200 import postproc
    with tempfile.TemporaryDirectory() as d:
        f = os.path.join(d, 'star_samp.dat')
        write_samples_file(f, one, star=0, likeobj=L)
        header = open(f).readline().split()
        keys = parsdict_keys(L.fitpars_i, L.fixedpars, iso=True, ageweight=True)
        assert header == ['Iter'] + keys + ['log(lk)', 'log(vol)', 'log(wt)', 'h', 'nc', 'log(z)', 'delta(log(z))']
        assert header[1:7] == L.fitpars_i and header[7:12] == ['Teff', 'log(g)', 'log(R)', '[Fe/H]', '[a/Fe]']
        tab = postproc.read_samples_file(f)
        n = one['niter'][0]
        assert len(tab['Iter']) == n + 150 and np.array_equal(tab['Iter'], np.arange(n + 150))
        prob = np.exp(tab['log(wt)'] - tab['log(z)'][-1])
        assert abs(prob.sum() - 1.0) < 1e-9
        assert abs(tab['log(z)'][-1] - one['logz'][0]) < 1e-9       # the file's final evidence is the sampler's
        assert np.all(np.diff(tab['log(z)']) >= 0) and np.all(np.isfinite(tab['h'])) and np.all(tab['nc'][:n] == 10)
        assert np.all(tab['delta(log(z))'][:n][1:] > 0) and tab['delta(log(z))'][n - 1] < 0.05
        assert abs(tab['h'][-1] - one['information'][0]) < 1e-6
        np.testing.assert_allclose(tab['Teff'][:n], 10.0 ** dead[:, nd + 8], rtol=1e-12)
        st = postproc.summarize_samples_file(f)
        i_d = L.fitpars_i.index('Dist')
        assert abs(st['Dist'][1] - one['mean'][0, i_d]) < 0.5 * one['std'][0, i_d]
    # replicas in lock step
    S = 8
    L.engine.set_stars_phot(np.tile([[fitargs['obs_phot'][b][0] for b in L.engine.bands]], (S, 1)),
                            np.tile([[fitargs['obs_phot'][b][1] for b in L.engine.bands]], (S, 1)))
    res = QueuedNestedSampler(L, P, nstars=S, nlive=150, walks=10, queue=8, dlogz=0.05, seed=4).run()
    assert res['converged'].all()
    err = np.maximum(res['logzerr'], 0.05)
    assert np.all(np.abs(res['logz'] - logz_mc) < 5 * err + 0.2), (res['logz'], logz_mc)
    m_ns = res['mean'][:, :nd].mean(0)
    assert np.all(np.abs(m_ns - mean_mc) < 0.35 * std_mc + 1e-3), (m_ns, mean_mc, std_mc)
    assert np.all(np.abs(res['std'][:, :nd].mean(0) / std_mc - 1.0) < 0.35)


def test_queued_sampler_unif_proposal_matches_monte_carlo():
    """dynesty's 'unif' proposal ("unif / rwalk"): independent draws from the bounding
    ellipsoid of the live points, rejected outside the unit cube, one likelihood call per proposal.  Evidence and
    posterior moments against brute-force prior Monte-Carlo and against the rwalk run of the same problem."""
    from minesweeper_b200.likelihood import likelihood
    from minesweeper_b200.prior import prior
    from minesweeper_b200.sampler import QueuedNestedSampler
    fitargs, fitpars, runbools, priordict = _problem()
    L = likelihood(fitargs, fitpars, runbools, precision='tc', verbose=False)
    P = prior(fitargs, priordict, fitpars, runbools, likeobj=L)
    gen = torch.Generator(device=L.engine.device).manual_seed(7)
    lw, ths = [], []
    for _ in range(8):
        u = torch.rand(1 << 18, L.ndim, generator=gen, device=L.engine.device, dtype=torch.float64)
        th = P.priortrans_batch(u)
        lnl, der, _ = L.lnlike_batch(th)
        lw.append(P.lnprob_batch(th, lnl, der)); ths.append(th)
    lw, ths = torch.cat(lw), torch.cat(ths)
    logz_mc = float(torch.logsumexp(lw, 0) - np.log(len(lw)))
    w = torch.exp(lw - lw.max()); w = w / w.sum()
    mean_mc = (w[:, None] * ths).sum(0).cpu().numpy()
    std_mc = torch.sqrt((w[:, None] * (ths - torch.as_tensor(mean_mc, device=ths.device)) ** 2).sum(0)).cpu().numpy()
    nd = L.ndim
    S = 8
    L.engine.set_stars_phot(np.tile([[fitargs['obs_phot'][b][0] for b in L.engine.bands]], (S, 1)),
                            np.tile([[fitargs['obs_phot'][b][1] for b in L.engine.bands]], (S, 1)))
    with pytest.raises(ValueError):
        QueuedNestedSampler(L, P, nstars=S, sample='slice')
    res = QueuedNestedSampler(L, P, nstars=S, nlive=150, queue=32, dlogz=0.05, seed=11, sample='unif').run()
    assert res['converged'].all()
    err = np.maximum(res['logzerr'], 0.05)
    assert np.all(np.abs(res['logz'] - logz_mc) < 5 * err + 0.2), (res['logz'], logz_mc)
    m_ns = res['mean'][:, :nd].mean(0)
    assert np.all(np.abs(m_ns - mean_mc) < 0.35 * std_mc + 1e-3), (m_ns, mean_mc, std_mc)
    assert np.all(np.abs(res['std'][:, :nd].mean(0) / std_mc - 1.0) < 0.35)
    # every proposal is one likelihood call; the sampling efficiency of the ellipsoid stays sane
    eff = res['niter'] / np.maximum(res['ncall'], 1)
    assert np.all(eff > 0.002) and np.all(eff <= 1.0), eff


def test_catalog_sampler_accuracy_gate_at_benched_settings():
    """The configuration bench.py times (catalog of stars with per-star parallax priors, nlive 150, 4 queued
    walkers, dlogz 0.01) against brute-force quadrature of a subsample through the same kernels
    (tools/bench_catalog.py).  The headline setting (walks = GATE_WALKS) must pass the accuracy gate; the demo's
    walks = 5 (demo/etc/samplerinfo.dat) is measured too: its posteriors are narrower, which is why bench.py does
    not quote stars/hour at that setting.  The truth z-scores of the exact posteriors are ~1 (the ground truth is
    sound); the sampler's are larger (1.2-2.2 at nlive = 150: a few stars per hundred lose a minor posterior mode),
    which bench.py reports next to the gate."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tools'))
    from bench_catalog import make_catalog_sampler, catalog_priordict, accuracy_block, GATE_WALKS
    from minesweeper_b200 import synth
    from minesweeper_b200.likelihood import likelihood
    from minesweeper_b200.prior import prior
    mist = synth.make_mist(ne=200, nm=40, nf=9, na=5, seed=0)
    phot = synth.make_phot_net(h=128, seed=1)
    fitpars = [NAMES, {k: (k in ON) for k in NAMES}]
    fitargs = dict(ageweight=True, mistdif=True, fixedpars={}, MISTpath=mist, photANNpath=phot,
                   obs_phot=synth.demo_obs_phot())
    rb = [False, True, False, False, False]
    pd = catalog_priordict(401.0)
    L = likelihood(fitargs, fitpars, rb, precision='tc', verbose=False)
    P = prior(fitargs, pd, fitpars, rb, likeobj=L)
    out = {}
    for walks in (GATE_WALKS, 5):
        ns, th = make_catalog_sampler(L, P, 1536, L.engine.device, 150, walks, 0.01, 20000, 0, queued=4)
        res = ns.run(check_every=64)
        assert res['converged'].all()
        out[walks] = accuracy_block(L, P, ns, res, th, pd, 48)
    good, demo = out[GATE_WALKS], out[5]
    assert good['bruteforce_ess_median'] > 200
    assert good['gate']['passed'], good
    assert not demo['gate']['passed'], demo
    # shorter walks give narrower posteriors (insufficient decorrelation inside the likelihood constraint)
    assert np.median(list(demo['std_ratio_median'].values())) < np.median(list(good['std_ratio_median'].values())) + 0.02
    # the exact (brute-force) posteriors are calibrated against the truths -- the ground truth itself is sound
    assert all(0.6 < v < 1.8 for v in good['truth_z_rms_bruteforce'].values()), good['truth_z_rms_bruteforce']


def test_fused_walk_step_matches_the_separate_kernels():
    """ms_ns_propose_transform / ms_ns_lnprob_accept (three launches per walk step) against ms_ns_propose +
    ms_prior_transform + ms_lnprob_batch_adv + ms_ns_accept (five): one step on identical state gives identical
    proposals, parameters, ln probabilities and walker records; whole runs from the same seed agree."""
    import ctypes as C
    from minesweeper_b200 import _lib
    from minesweeper_b200.likelihood import likelihood
    from minesweeper_b200.prior import prior
    from minesweeper_b200.sampler import QueuedNestedSampler
    fitargs, fitpars, runbools, priordict = _problem()
    L = likelihood(fitargs, fitpars, runbools, precision='tc', verbose=False)
    P = prior(fitargs, priordict, fitpars, runbools, likeobj=L)
    eng, lib, dev = L.engine, L.engine.lib, L.engine.device
    S, K, Q, nd, nder = 37, 20, 5, L.ndim, _lib.MS_NDERIVED
    ncol = nd + nder
    eng.set_stars_phot(np.tile([[fitargs['obs_phot'][b][0] for b in eng.bands]], (S, 1)),
                       np.tile([[fitargs['obs_phot'][b][1] for b in eng.bands]], (S, 1)))
    gen = torch.Generator(device=dev).manual_seed(5)
    f64 = torch.float64
    r = lambda *sh: torch.rand(*sh, generator=gen, device=dev, dtype=f64)
    z = lambda *sh, dt=f64: torch.zeros(*sh, dtype=dt, device=dev)
    plx = torch.stack([1000.0 / 14.0 + r(S), 3.0 + r(S)], 1).contiguous()
    slot = torch.arange(S, device=dev, dtype=torch.int32).repeat_interleave(Q).contiguous()

    def state():
        g2 = torch.Generator(device=dev).manual_seed(9)
        rr = lambda *sh: torch.rand(*sh, generator=g2, device=dev, dtype=f64)
        sig = torch.zeros(S, nd, nd, dtype=f64, device=dev)
        sig[:, torch.arange(nd), torch.arange(nd)] = 0.02
        sig[:, 1, 0] = 0.01
        t = dict(live_u=rr(S * K, nd), live_lp=rr(S * K), live_row=rr(S * K, ncol), cur_u=0.3 + 0.4 * rr(S * Q, nd),
                 cur_lp=z(S * Q), cur_row=z(S * Q, ncol), prop_u=z(S * Q, nd), prop_lp=z(S * Q), prop_theta=z(S * Q, nd),
                 prop_der=z(S * Q, nder), lnl=z(S * Q), nacc=z(S * Q, dt=torch.int32), lmin=z(S), sigma=sig.reshape(S, nd * nd).contiguous(),
                 scale=torch.ones(S, dtype=f64, device=dev), logz=z(S), hacc=z(S), wmax=z(S), wtot=z(S), wsum=z(S, ncol),
                 wsq=z(S, ncol), niter=z(S, dt=torch.int64), ncall=z(S, dt=torch.int64), done=z(S, dt=torch.int32),
                 rng_round=z(1, dt=torch.int64) + 3, center=z(S, nd), ell_r=z(S))
        st = _lib.NsState()
        st.S, st.K, st.Q, st.nd, st.nder, st.walks = S, K, Q, nd, nder, 4
        st.maxiter, st.dead_cap, st.dlogz, st.unif = 1000, 0, 0.01, 0
        for name, typ in _lib.NsState._fields_[9:]:
            if typ is C.c_void_p:
                setattr(st, name, t[name].data_ptr() if name in t else None)
        return t, st

    seed = C.c_uint64(1234)
    stream = eng._stream()

    def like(t):
        eng.lnlike_batch(t['prop_theta'], L._colmap, L._fixed, L._flags, star_slot=slot, want_derived=True,
                         out=t['lnl'], derived_out=t['prop_der'])

    with torch.cuda.device(dev):
        ta, sa = state()
        tb, sb = state()
        for step in range(2):
            # separate kernels
            _lib.check(lib.ms_ns_propose(eng._ctx, C.byref(sa), seed, C.c_uint64(step), stream))
            P.priortrans_batch(ta['prop_u'], out=ta['prop_theta'])
            like(ta)
            if step == 0:                                      # threshold: about half of the proposals pass
                fin = ta['lnl'][torch.isfinite(ta['lnl'])]
                assert len(fin) > 10
                pre = P.lnprob_batch(ta['prop_theta'], ta['lnl'], ta['prop_der'], star_slot=slot, star_plx=plx)
                thr = pre[torch.isfinite(pre)].median()
                ta['lmin'].fill_(float(thr)); tb['lmin'].fill_(float(thr))
            P.lnprob_batch(ta['prop_theta'], ta['lnl'], ta['prop_der'], out=ta['prop_lp'], star_slot=slot, star_plx=plx)
            _lib.check(lib.ms_ns_accept(eng._ctx, C.byref(sa), stream))
            # fused
            _lib.check(lib.ms_ns_propose_transform(eng._ctx, C.byref(sb), seed, C.c_uint64(step), P._cols, stream))
            like(tb)
            _lib.check(lib.ms_ns_lnprob_accept(
                eng._ctx, C.byref(sb), C.byref(P._flags), P._colmap.ctypes.data_as(_lib.c_ip),
                P._fixed.ctypes.data_as(_lib.c_dp), P._terms_c, P._nterms, int(P.imf_bool), tb['lnl'].data_ptr(),
                slot.data_ptr(), plx.data_ptr(), C.byref(P._adv) if P._adv is not None else None, stream))
            torch.cuda.synchronize()
            assert torch.equal(ta['prop_u'], tb['prop_u'])
            assert torch.allclose(ta['prop_theta'], tb['prop_theta'], rtol=1e-15, atol=0, equal_nan=True)
            assert torch.allclose(ta['prop_lp'], tb['prop_lp'], rtol=1e-13, atol=1e-12, equal_nan=True)
            nacc = int(ta['nacc'].sum())
            assert 0 < nacc < (step + 1) * S * Q
            assert torch.equal(ta['nacc'], tb['nacc'])
            assert torch.equal(ta['cur_u'], tb['cur_u'])
            assert torch.allclose(ta['cur_row'], tb['cur_row'], rtol=1e-15, atol=0)
            assert torch.allclose(ta['cur_lp'], tb['cur_lp'], rtol=1e-13, atol=1e-12)
    # whole runs from the same seed
    kw = dict(nstars=S, nlive=60, walks=6, queue=4, dlogz=0.1, seed=2, star_plx=plx)
    ra = QueuedNestedSampler(L, P, fused=False, **kw).run()
    rb = QueuedNestedSampler(L, P, fused=True, **kw).run()
    assert ra['converged'].all() and rb['converged'].all()
    if np.array_equal(ra['niter'], rb['niter']):               # same trajectories (the usual case)
        np.testing.assert_allclose(ra['logz'], rb['logz'], rtol=0, atol=1e-8)
    else:                                                      # a last-bit difference flipped an acceptance somewhere
        err = np.maximum(np.hypot(ra['logzerr'], rb['logzerr']), 0.1)
        assert np.all(np.abs(ra['logz'] - rb['logz']) < 5 * err)


def test_spread_kernel_matches_a_library_cholesky():
    """ms_ns_spread (covariance + Cholesky of the live points, one warp per star; also the code inside ms_ns_consume)
    against torch.linalg as the checker, including the degenerate fallback."""
    from minesweeper_b200.likelihood import likelihood
    from minesweeper_b200.sampler import _chol_cov
    fitargs, fitpars, runbools, _ = _problem()
    L = likelihood(fitargs, fitpars, runbools, precision='f32', verbose=False)
    dev = L.engine.device
    gen = torch.Generator(device=dev).manual_seed(1)
    S, K, nd = 133, 150, 6
    mix = torch.randn(S, nd, nd, generator=gen, device=dev, dtype=torch.float64)
    U = torch.bmm(torch.randn(S, K, nd, generator=gen, device=dev, dtype=torch.float64), mix) * 0.05 + 0.5
    U[7] = U[7, :1]                                            # all live points equal: zero covariance
    U[11, :, 2] = 0.25                                         # one coordinate constant: singular covariance
    sig = _chol_cov(L.engine, U)
    d = U - U.mean(dim=1, keepdim=True)
    cov = torch.matmul(d.transpose(1, 2), d) / (K - 1)
    good = torch.ones(S, dtype=torch.bool, device=dev)
    good[7] = good[11] = False
    ref = torch.linalg.cholesky(cov[good] + 1e-14 * torch.eye(nd, dtype=torch.float64, device=dev))
    assert torch.allclose(sig[good], ref, rtol=1e-8, atol=1e-11)
    assert torch.equal(sig[good].triu(1), torch.zeros_like(sig[good]))
    # star 7: zero covariance + 1e-14 jitter is still positive definite -> a tiny factor; star 11: the singular
    # coordinate gets sqrt(jitter) on the diagonal, or the diagonal fallback; either way finite and lower triangular
    assert torch.isfinite(sig).all() and torch.equal(sig.triu(1), torch.zeros_like(sig))
    assert float(sig[7].abs().max()) <= 1.1e-7 and float(sig[11, 2, 2]) <= 1.1e-7
    rec = torch.matmul(sig[11], sig[11].T)
    keep = [0, 1, 3, 4, 5]
    assert torch.allclose(rec[keep][:, keep], cov[11][keep][:, keep], rtol=1e-8, atol=1e-12)
