"""Lock-step batched static nested sampling.

The reference hands ``lnprobfn`` / ``priortrans`` to ``dynesty.NestedSampler`` with the ``rwalk``
proposal and evaluates one point per call (reference fitstar.py:337-369; settings in
demo/etc/samplerinfo.dat: static, rwalk, 150 live points, walks 5, dlogz 0.01).  dynesty is not
available here and cannot feed a GPU one point at a time, so this module runs the same kind of
sampler for MANY stars at once: every star has its own live set, all stars advance one
nested-sampling iteration together, and each random-walk step is one batched call of the CUDA
prior-transform / likelihood / ln-prior kernels over all still-active stars.

Algorithm per star (standard static nested sampling, Skilling 2006, with dynesty-style rwalk):
  * live points uniform in the unit cube, lnprob = lnL + lnprior from the kernels;
  * each iteration removes the worst live point (ln X_i = -i / nlive), accumulates ln Z and the
    information H, and replaces it by the end point of a constrained random walk started from
    another live point: ``walks`` Gaussian steps in the unit cube, scaled by the live-point
    spread, reflected at the walls, accepted iff lnprob > lnprob_worst; the step scale adapts to
    a 0.5 acceptance rate (dynesty's update rule); walks that never moved are extended;
  * stop when the remaining evidence ln(1 + L_max X / Z) < dlogz; then add the live points.
The bookkeeping is torch tensor code on the GPU (no host round trip per iteration); all model
arithmetic is in libmsb200's kernels.  Output: per-star ln Z, its error sqrt(H / nlive), call
counts, posterior mean / std of theta and of the derived MIST block, and (optionally) the dead
points with ln weights in the column layout of the reference's samples file
(fitstar.py:235-242, 392-425).
"""
import time

import numpy as np
import torch

from . import _lib


def _chol_cov(eng, U):
    """Cholesky factors [S, nd, nd] of the covariance of the live points U[S, K, nd] (the shape of the random-walk
    proposal, as dynesty's rwalk uses the bounding ellipsoid): ``ms_ns_spread``, the kernel code that also maintains
    the queued sampler's proposal shape (no cuBLAS / cuSOLVER call on the sampler's path)."""
    U = U.contiguous()
    S, K, nd = U.shape
    sig = torch.empty((S, nd, nd), dtype=torch.float64, device=U.device)
    with torch.cuda.device(eng.device):
        _lib.check(eng.lib.ms_ns_spread(eng._ctx, U.data_ptr(), S, K, nd, sig.data_ptr(), eng._stream()))
    return sig


class BatchedNestedSampler(object):
    def __init__(self, likeobj, priorobj, nstars=1, nlive=150, walks=5, dlogz=0.01, maxiter=100000,
                 seed=0, star_plx=None, store_samples=False, max_extend=20):
        self.L, self.P = likeobj, priorobj
        self.eng = likeobj.engine
        self.dev = self.eng.device
        self.S, self.K, self.walks = int(nstars), int(nlive), int(walks)
        self.ndim = likeobj.ndim
        self.dlogz, self.maxiter, self.max_extend = float(dlogz), int(maxiter), int(max_extend)
        self.seed = int(seed)
        self.gen = torch.Generator(device=self.dev)
        self.gen.manual_seed(self.seed)
        self.per_star = self.S > 1 or star_plx is not None
        self.star_plx = None if star_plx is None else torch.as_tensor(star_plx, dtype=torch.float64).to(self.dev).contiguous()
        self.store = bool(store_samples)
        self.ncall = torch.zeros(self.S, dtype=torch.int64, device=self.dev)
        self.eng.check_star_slots(self.S, self.star_plx)

    # one batched evaluation of lnprob for u[N, ndim] belonging to stars slot[N]
    def _lnprob(self, u, slot):
        theta = self.P.priortrans_batch(u)
        slots = slot.to(torch.int32).contiguous() if self.per_star else None
        lnl, der, _ = self.eng.lnlike_batch(theta, self.L._colmap, self.L._fixed, self.L._flags,
                                            star_slot=slots if self.S > 1 else None,
                                            want_derived=self.L.iso_bool)
        out = self.P.lnprob_batch(theta, lnl, der, star_slot=slots if self.star_plx is not None else None,
                                  star_plx=self.star_plx)
        self.ncall.index_add_(0, slot, torch.ones_like(slot))
        return out, theta, der

    def _rand(self, *shape):
        return torch.rand(*shape, generator=self.gen, device=self.dev, dtype=torch.float64)

    def _randn(self, *shape):
        return torch.randn(*shape, generator=self.gen, device=self.dev, dtype=torch.float64)

    def run(self, verbose=False):
        S, K, nd, dev = self.S, self.K, self.ndim, self.dev
        f64 = torch.float64
        star = torch.arange(S, device=dev)
        # ---- initial live points (redrawn until lnprob is finite) ----
        U = self._rand(S, K, nd)
        slot = star.repeat_interleave(K)
        lp, th, der = self._lnprob(U.reshape(-1, nd), slot)
        nder = der.shape[1] if der is not None else 0
        for _ in range(50):
            bad = ~torch.isfinite(lp)
            if not bool(bad.any()):
                break
            idx = bad.nonzero(as_tuple=True)[0]
            un = self._rand(len(idx), nd)
            lpn, thn, dern = self._lnprob(un, slot[idx])
            U.reshape(-1, nd)[idx] = un
            lp[idx], th[idx] = lpn, thn
            if der is not None:
                der[idx] = dern
        if bool((~torch.isfinite(lp)).any()):
            raise RuntimeError('could not draw live points with finite lnprob from the prior')
        LP = lp.reshape(S, K)
        TH = th.reshape(S, K, nd)
        DER = der.reshape(S, K, nder) if der is not None else None
        ncol = nd + nder
        logz = torch.full((S,), -float('inf'), dtype=f64, device=dev)
        hacc = torch.zeros(S, dtype=f64, device=dev)            # sum exp(logwt - logz) * lnL, for H
        wsum = torch.zeros(S, ncol, dtype=f64, device=dev)      # posterior moments with weights exp(logwt - wmax)
        wsq = torch.zeros(S, ncol, dtype=f64, device=dev)
        wmax = torch.full((S,), -float('inf'), dtype=f64, device=dev)
        wtot = torch.zeros(S, dtype=f64, device=dev)
        scale = torch.ones(S, dtype=f64, device=dev)
        active = torch.ones(S, dtype=torch.bool, device=dev)
        niter = torch.zeros(S, dtype=torch.int64, device=dev)
        dead = [] if self.store else None
        lnK = float(np.log(K))
        log_shrink = float(np.log1p(-np.exp(-1.0 / K)))        # ln(1 - e^{-1/K}): shell width factor

        def accumulate(mask, lnl, logwt, row):
            """ln Z, H and posterior-moment updates for the stars in ``mask``."""
            nonlocal logz, hacc, wsum, wsq, wmax, wtot
            lz_new = torch.logaddexp(logz, logwt)
            # H bookkeeping: keep S1 = sum w_i lnL_i / Z in rescaled form
            hacc = torch.where(mask, hacc * torch.exp(logz - lz_new) + torch.exp(logwt - lz_new) * lnl, hacc)
            logz = torch.where(mask, lz_new, logz)
            m_new = torch.where(mask, torch.maximum(wmax, logwt), wmax)
            r_old = torch.exp(wmax - m_new)
            r_old = torch.where(torch.isfinite(r_old), r_old, torch.zeros_like(r_old))
            w = torch.where(mask, torch.exp(logwt - m_new), torch.zeros_like(logwt))
            wsum = wsum * r_old[:, None] + w[:, None] * row
            wsq = wsq * r_old[:, None] + w[:, None] * row * row
            wtot = wtot * r_old + w
            wmax = m_new

        it = 0
        while bool(active.any()) and it < self.maxiter:
            it += 1
            act = active.nonzero(as_tuple=True)[0]
            A = len(act)
            # ---- worst live point of every active star ----
            lmin, w = LP[act].min(dim=1)
            logvol = -(niter[act].to(f64) + 1.0) / K
            logwt_full = torch.full((S,), -float('inf'), dtype=f64, device=dev)
            logwt_full[act] = lmin + (logvol + 1.0 / K + log_shrink)       # lnL_i + ln(X_{i-1} - X_i)
            row = torch.zeros(S, ncol, dtype=f64, device=dev)
            rowA = TH[act, w] if DER is None else torch.cat([TH[act, w], DER[act, w]], dim=1)
            row[act] = torch.where(torch.isfinite(rowA), rowA, torch.zeros_like(rowA))
            lfull = torch.zeros(S, dtype=f64, device=dev)
            lfull[act] = lmin
            accumulate(active, lfull, logwt_full, row)
            if self.store:
                dead.append((act.clone(), rowA.clone(), lmin.clone(), logvol.clone(), logwt_full[act].clone()))
            # ---- constrained random walk from another live point ----
            j = torch.randint(0, K - 1, (A,), generator=self.gen, device=dev)
            j = j + (j >= w).to(j.dtype)                                   # any live point but the worst
            ucur = U[act, j].clone()
            lcur = LP[act, j].clone()
            thcur = TH[act, j].clone()
            dcur = DER[act, j].clone() if DER is not None else None
            sig = _chol_cov(self.eng, U[act]) * scale[act, None, None]
            nacc = torch.zeros(A, dtype=torch.int64, device=dev)
            steps = 0
            while True:
                for _ in range(self.walks):
                    prop = ucur + (sig * self._randn(A, 1, nd)).sum(2)       # L z, element-wise (6 x 6: no library GEMM)
                    prop = torch.remainder(prop, 2.0)                      # reflect into [0, 1]
                    prop = torch.where(prop > 1.0, 2.0 - prop, prop)
                    lpn, thn, dern = self._lnprob(prop, act)
                    ok = lpn > lmin
                    ucur = torch.where(ok[:, None], prop, ucur)
                    lcur = torch.where(ok, lpn, lcur)
                    thcur = torch.where(ok[:, None], thn, thcur)
                    if dcur is not None:
                        dcur = torch.where(ok[:, None], dern, dcur)
                    nacc += ok
                steps += self.walks
                stuck = nacc == 0
                if not bool(stuck.any()) or steps >= self.walks * self.max_extend:
                    break
                sig = torch.where(stuck[:, None, None], sig * 0.5, sig)    # shrink and keep walking
            # dynesty's rwalk scale update towards 50 % acceptance
            facc = nacc.to(f64) / steps
            scale[act] = (scale[act] * torch.exp((facc - 0.5) / nd / 0.5)).clamp(1e-4, 10.0)
            U[act, w], LP[act, w], TH[act, w] = ucur, lcur, thcur
            if DER is not None:
                DER[act, w] = dcur
            niter[act] += 1
            # ---- stopping rule: remaining evidence ----
            lmax = LP[act].max(dim=1).values
            dl = torch.logaddexp(logz[act], lmax + logvol) - logz[act]
            done = dl < self.dlogz
            active[act[done]] = False
            if verbose and it % 200 == 0:
                print('iter %d: %d / %d stars active, median dlogz %.3g' % (it, int(active.sum()), S, float(dl.median())))

        # ---- final live points: each carries X_final / K ----
        logvol_f = -(niter.to(f64)) / K - lnK
        allmask = torch.ones(S, dtype=torch.bool, device=dev)
        for k in range(K):
            rowk = TH[:, k] if DER is None else torch.cat([TH[:, k], DER[:, k]], dim=1)
            rowk = torch.where(torch.isfinite(rowk), rowk, torch.zeros_like(rowk))
            accumulate(allmask, LP[:, k], LP[:, k] + logvol_f, rowk)
        mean = wsum / wtot[:, None]
        var = (wsq / wtot[:, None] - mean * mean).clamp_min(0.0)
        H = hacc - logz                                            # H = sum(w lnL)/Z - ln Z
        res = dict(logz=logz.cpu().numpy(), logzerr=torch.sqrt(H.clamp_min(0.0) / K).cpu().numpy(),
                   information=H.cpu().numpy(), niter=niter.cpu().numpy(), ncall=self.ncall.cpu().numpy(),
                   names=list(self.L.fitpars_i) + (list(_lib.DERIVED_NAMES) if nder else []),
                   mean=mean.cpu().numpy(), std=torch.sqrt(var).cpu().numpy(),
                   live_theta=TH.cpu().numpy(), live_lnprob=LP.cpu().numpy(), converged=(~active).cpu().numpy())
        if self.store:
            res['dead'] = [tuple(t.cpu().numpy() for t in d) for d in dead]
        return res


    # ------------------------------------------------------------------------------------------
    def run_graphed(self, check_every=64, verbose=False, min_compact=512):
        self.eng.reserve(self.S * self.K, lock=True)
        try:
            return self._run_graphed(check_every, verbose, min_compact)
        finally:
            self.eng.reserve(0, lock=False)

    def _run_graphed(self, check_every=64, verbose=False, min_compact=512):
        """Same sampler with one nested-sampling iteration captured in a CUDA graph and replayed.

        Every tensor has a static shape (all stars of the current working set take part in every
        iteration; finished stars are masked), there is no host synchronisation inside an
        iteration, and convergence is polled from the host only every ``check_every`` replays.
        When fewer than half of the working set is still active, the finished stars are retired
        (their live points are added to the evidence and moments), the state is compacted to the
        active ones and the graph is captured again, so stragglers do not keep the whole catalog
        busy.  This removes the launch and Python overhead that bounds ``run`` (about 4 ms per
        iteration regardless of catalog size).  Differences from ``run``: a walk that never moved
        is not extended (the start point, a copy of another live point, is kept -- with the
        adaptive scale this happens in a few per cent of the iterations); dead points are not
        stored.
        """
        S, K, nd, dev = self.S, self.K, self.ndim, self.dev
        f64 = torch.float64
        ninf = -float('inf')
        # the captured iteration draws from the device's default generator (graph-safe): seed it so runs reproduce
        with torch.cuda.device(dev):
            torch.cuda.manual_seed(self.seed)
        star0 = torch.arange(S, device=dev)
        # ---- initial live points, eagerly (redrawn until finite) ----
        U = self._rand(S, K, nd)
        slot = star0.repeat_interleave(K)
        lp, th, der = self._lnprob(U.reshape(-1, nd), slot)
        nder = der.shape[1] if der is not None else 0
        for _ in range(50):
            bad = ~torch.isfinite(lp)
            if not bool(bad.any()):
                break
            idx = bad.nonzero(as_tuple=True)[0]
            un = self._rand(len(idx), nd)
            lpn, thn, dern = self._lnprob(un, slot[idx])
            U.reshape(-1, nd)[idx] = un
            lp[idx], th[idx] = lpn, thn
            if der is not None:
                der[idx] = dern
        if bool((~torch.isfinite(lp)).any()):
            raise RuntimeError('could not draw live points with finite lnprob from the prior')
        ncol = nd + nder
        st = dict(U=U, LP=lp.reshape(S, K).clone(), TH=th.reshape(S, K, nd).clone(),
                  logz=torch.full((S,), ninf, dtype=f64, device=dev), hacc=torch.zeros(S, dtype=f64, device=dev),
                  wsum=torch.zeros(S, ncol, dtype=f64, device=dev), wsq=torch.zeros(S, ncol, dtype=f64, device=dev),
                  wmax=torch.full((S,), ninf, dtype=f64, device=dev), wtot=torch.zeros(S, dtype=f64, device=dev),
                  scale=torch.ones(S, dtype=f64, device=dev), active=torch.ones(S, dtype=torch.bool, device=dev),
                  niter=torch.zeros(S, dtype=torch.int64, device=dev), ncall=self.ncall.clone(),
                  orig=star0.clone())
        if der is not None:
            st['DER'] = der.reshape(S, K, nder).clone()
        log_shrink = float(np.log1p(-np.exp(-1.0 / K)))
        per_star = self.S > 1
        plx = self.star_plx
        # results for the whole catalog, filled as stars retire
        out = dict(logz=torch.empty(S, dtype=f64, device=dev), H=torch.empty(S, dtype=f64, device=dev),
                   niter=torch.zeros(S, dtype=torch.int64, device=dev), ncall=torch.zeros(S, dtype=torch.int64, device=dev),
                   mean=torch.empty(S, ncol, dtype=f64, device=dev), std=torch.empty(S, ncol, dtype=f64, device=dev),
                   converged=torch.zeros(S, dtype=torch.bool, device=dev))

        def accumulate(sx, mask, lnl, logwt, row):
            lz_new = torch.logaddexp(sx['logz'], logwt)
            h_new = sx['hacc'] * torch.exp(sx['logz'] - lz_new) + torch.exp(logwt - lz_new) * lnl
            sx['hacc'].copy_(torch.where(mask & torch.isfinite(h_new), h_new, sx['hacc']))
            sx['logz'].copy_(torch.where(mask, lz_new, sx['logz']))
            m_new = torch.where(mask, torch.maximum(sx['wmax'], logwt), sx['wmax'])
            r_old = torch.exp(sx['wmax'] - m_new)
            r_old = torch.where(torch.isfinite(r_old), r_old, torch.zeros_like(r_old))
            w = torch.where(mask, torch.exp(logwt - m_new), torch.zeros_like(logwt))
            w = torch.where(torch.isfinite(w), w, torch.zeros_like(w))
            sx['wsum'].copy_(sx['wsum'] * r_old[:, None] + w[:, None] * row)
            sx['wsq'].copy_(sx['wsq'] * r_old[:, None] + w[:, None] * row * row)
            sx['wtot'].copy_(sx['wtot'] * r_old + w)
            sx['wmax'].copy_(m_new)

        def retire(idx):
            """Final live-point contribution and summaries of the stars st[...][idx]."""
            sx = {k: v[idx].clone() for k, v in st.items()}
            n = len(idx)
            logvol_f = -(sx['niter'].to(f64)) / K - float(np.log(K))
            allm = torch.ones(n, dtype=torch.bool, device=dev)
            for k in range(K):
                rowk = sx['TH'][:, k] if 'DER' not in sx else torch.cat([sx['TH'][:, k], sx['DER'][:, k]], dim=1)
                rowk = torch.where(torch.isfinite(rowk), rowk, torch.zeros_like(rowk))
                accumulate(sx, allm, sx['LP'][:, k], sx['LP'][:, k] + logvol_f, rowk)
            mean = sx['wsum'] / sx['wtot'][:, None]
            var = (sx['wsq'] / sx['wtot'][:, None] - mean * mean).clamp_min(0.0)
            o = sx['orig']
            out['logz'][o] = sx['logz']; out['H'][o] = sx['hacc'] - sx['logz']
            out['niter'][o] = sx['niter']; out['ncall'][o] = sx['ncall']
            out['mean'][o] = mean; out['std'][o] = torch.sqrt(var)
            out['converged'][o] = ~sx['active']

        total = 0
        ncapt = 0
        while True:
            Sc = len(st['orig'])
            star = torch.arange(Sc, device=dev)
            slot32 = st['orig'].to(torch.int32).contiguous()

            def lnprob_static(u):
                theta = self.P.priortrans_batch(u)
                lnl, dr, _ = self.eng.lnlike_batch(theta, self.L._colmap, self.L._fixed, self.L._flags,
                                                   star_slot=slot32 if per_star else None,
                                                   want_derived=self.L.iso_bool)
                o = self.P.lnprob_batch(theta, lnl, dr, star_slot=slot32 if plx is not None else None, star_plx=plx)
                return o, theta, dr

            def iteration():
                act = st['active']
                LP, U_, TH, DER = st['LP'], st['U'], st['TH'], st.get('DER')
                lmin, w = LP.min(dim=1)
                logvol = -(st['niter'].to(f64) + 1.0) / K
                logwt = torch.where(act, lmin + (logvol + 1.0 / K + log_shrink), torch.full_like(lmin, ninf))
                row = TH[star, w] if DER is None else torch.cat([TH[star, w], DER[star, w]], dim=1)
                row = torch.where(torch.isfinite(row), row, torch.zeros_like(row))
                accumulate(st, act, lmin, logwt, row)
                j = torch.randint(0, K - 1, (Sc,), device=dev)
                j = j + (j >= w).to(j.dtype)
                ucur, lcur, thcur = U_[star, j], LP[star, j], TH[star, j]
                dcur = DER[star, j] if DER is not None else None
                sig = _chol_cov(self.eng, U_) * st['scale'][:, None, None]
                nacc = torch.zeros(Sc, dtype=f64, device=dev)
                for _ in range(self.walks):
                    prop = ucur + (sig * torch.randn(Sc, 1, nd, device=dev, dtype=f64)).sum(2)
                    prop = torch.remainder(prop, 2.0)
                    prop = torch.where(prop > 1.0, 2.0 - prop, prop)
                    lpn, thn, dern = lnprob_static(prop)
                    ok = lpn > lmin
                    ucur = torch.where(ok[:, None], prop, ucur)
                    lcur = torch.where(ok, lpn, lcur)
                    thcur = torch.where(ok[:, None], thn, thcur)
                    if dcur is not None:
                        dcur = torch.where(ok[:, None], dern, dcur)
                    nacc = nacc + ok.to(f64)
                facc = nacc / self.walks
                st['scale'].copy_(torch.where(act, (st['scale'] * torch.exp((facc - 0.5) / nd / 0.5)).clamp(1e-4, 10.0),
                                              st['scale']))
                a2 = act[:, None]
                U_[star, w] = torch.where(a2, ucur, U_[star, w])
                LP[star, w] = torch.where(act, lcur, LP[star, w])
                TH[star, w] = torch.where(a2, thcur, TH[star, w])
                if DER is not None:
                    DER[star, w] = torch.where(a2, dcur, DER[star, w])
                st['niter'] += act.to(torch.int64)
                st['ncall'] += act.to(torch.int64) * self.walks
                lmax = LP.max(dim=1).values
                dl = torch.logaddexp(st['logz'], lmax + logvol) - st['logz']
                st['active'].copy_(act & ~(dl < self.dlogz) & (st['niter'] < self.maxiter))

            # warm up on a side stream (scratch allocations, lazy initialisation), then capture
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(2):
                    iteration()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                iteration()
            ncapt += 1
            total += 3
            while True:
                for _ in range(check_every):
                    graph.replay()
                total += check_every
                n_act = int(st['active'].sum().item())
                if verbose:
                    print('replays %d: %d / %d stars active (working set %d)' % (total, n_act, S, Sc))
                if n_act == 0 or (Sc > min_compact and n_act * 2 <= Sc) or total >= self.maxiter + 3 * ncapt:
                    break
            del graph
            done = (~st['active']).nonzero(as_tuple=True)[0]
            if n_act == 0 or total >= self.maxiter + 3 * ncapt:
                retire(torch.arange(Sc, device=dev))
                break
            retire(done)
            keep = st['active'].nonzero(as_tuple=True)[0]
            st = {k: v[keep].clone() for k, v in st.items()}
        return dict(logz=out['logz'].cpu().numpy(), logzerr=torch.sqrt(out['H'].clamp_min(0.0) / K).cpu().numpy(),
                    information=out['H'].cpu().numpy(), niter=out['niter'].cpu().numpy(),
                    ncall=out['ncall'].cpu().numpy(),
                    names=list(self.L.fitpars_i) + (list(_lib.DERIVED_NAMES) if nder else []),
                    mean=out['mean'].cpu().numpy(), std=out['std'].cpu().numpy(),
                    converged=out['converged'].cpu().numpy(), graph_replays=total, graph_captures=ncapt)


def _capture_graph(fn, dev):
    """One call of ``fn`` captured into a CUDA graph on a side stream.  ``torch.cuda.graph`` does the same after a
    ``gc.collect()`` and ``torch.cuda.empty_cache()``; with the arrays of the previous, larger working set just released
    to the caching allocator that ``cudaFree`` of a few GB cost 0.5 s at the first compaction of a 100 000-star run
    (12 % of its tail), and nothing inside a round allocates."""
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        g.capture_begin()
        try:
            fn()
        finally:
            g.capture_end()
    torch.cuda.current_stream(dev).wait_stream(side)
    return g


def auto_queue(nstars):
    """Walkers per star for a catalog of ``nstars`` stars on one GPU (``queue='auto'``).  A larger queue gives the
    fused photometry kernel bigger batches and fewer bookkeeping rounds; its price is the fraction of queued candidates
    overtaken by the rising threshold (about q / 2K of them for q walkers and K live points).  Measured on a B200 at
    nlive 150, walks 25 (profiles/r02_catalog_sweep.jsonl): 20 000 stars: 5.3 / 3.2 / 2.8 / 3.1 s with 4 / 8 / 16 / 32
    walkers; 100 000 stars: 13.4 / 12.9 s with 4 / 8.  So: about 3e5-8e5 walkers per step, and never more than 64."""
    n = int(nstars)
    if n > 150000:
        return 4
    if n > 40000:
        return 8
    if n > 5000:
        return 16
    return 32 if n > 500 else 64


class QueuedNestedSampler(object):
    """Static nested sampling with a proposal queue, bookkeeping in CUDA kernels (csrc/sampler.cu).

    Each star owns ``queue`` random-walk walkers; a round advances all walkers of all stars by
    ``walks`` steps (one batched prior-transform / likelihood / ln-prior call per step) and then
    ``ns_consume_kernel`` (one warp per star) uses the walkers' end points in order, each one
    replacing the then-worst live point if it is still above the threshold -- dynesty's
    ``queue_size`` scheme.  ``queue > 1`` is what makes a *single* star fast on a GPU (the
    sequential depth drops from ~niter to ~niter / queue rounds); for catalogs a small queue
    (4: 9.0e7 stars/hour on a B200 at 100k stars, against 5.6e7 with 1) amortises the per-round
    bookkeeping and makes the likelihood batches large enough for the fused photometry kernel.  One round is captured in a CUDA graph and replayed; the RNG stream is
    driven by a device-side round counter so replays draw fresh numbers.
    """

    def __init__(self, likeobj, priorobj, nstars=1, nlive=150, walks=5, queue=32, dlogz=0.01, maxiter=100000,
                 seed=0, star_plx=None, store_samples=False, dead_cap=20000, max_queue=64, sample='rwalk', fused=True):
        """sample: 'rwalk' (random walk of ``walks`` steps from a live point; the reference's setting,
        demo/etc/samplerinfo.dat) or 'unif' (dynesty's uniform proposal: one independent draw from the bounding ellipsoid
        of the live points per walker and round, rejected outside the unit cube; ``walks`` is ignored).
        queue: walkers per star, or 'auto' (``auto_queue``: sized from the number of stars).
        fused: a walk step is three launches (propose + prior transform, likelihood, ln prior + accept) instead of five;
        same arithmetic, same random numbers (``fused=False`` keeps the separate kernels; the tests compare the two)."""
        self.fused = bool(fused)
        if not likeobj.iso_bool:
            raise NotImplementedError('the queued sampler tracks the derived MIST block (isochrone fits)')
        if sample not in ('rwalk', 'unif'):
            raise ValueError("sample must be 'rwalk' or 'unif' (dynesty's other proposals are not implemented)")
        self.sample = sample
        if sample == 'unif':
            walks = 1
        self.L, self.P = likeobj, priorobj
        self.eng = likeobj.engine
        self.dev = self.eng.device
        if queue is None or queue == 'auto':
            queue = auto_queue(nstars)
        self.S, self.K, self.Q, self.walks = int(nstars), int(nlive), int(queue), int(walks)
        self.nd, self.nder = likeobj.ndim, _lib.MS_NDERIVED
        self.dlogz, self.maxiter, self.seed = float(dlogz), int(maxiter), int(seed)
        self.star_plx = None if star_plx is None else torch.as_tensor(star_plx, dtype=torch.float64).to(self.dev).contiguous()
        self.store, self.dead_cap = bool(store_samples), int(dead_cap)
        self.max_queue = max(int(max_queue), self.Q)
        self.eng.check_star_slots(self.S, self.star_plx)

    def run(self, check_every=4, verbose=False, min_compact=512, compact_frac=0.85):
        # the rounds are captured into CUDA graphs that bake the ctx's scratch pointers in: size the scratch for the
        # largest batch of the run and pin it for its duration (another caller growing it would free it under us)
        self.eng.reserve(self.S * max(self.K, self.Q), lock=True)
        try:
            return self._run(check_every, verbose, min_compact, compact_frac)
        finally:
            self.eng.reserve(0, lock=False)

    def _run(self, check_every=4, verbose=False, min_compact=512, compact_frac=0.85):
        import ctypes as C
        if self.store and min_compact < self.S:
            raise ValueError('store_samples needs min_compact >= nstars (the dead-point buffer is not compacted)')
        S, K, Q, nd, nder, dev = self.S, self.K, self.Q, self.nd, self.nder, self.dev
        ncol = nd + nder
        f64 = torch.float64
        eng, lib = self.eng, self.eng.lib
        gen = torch.Generator(device=dev).manual_seed(self.seed)
        z = lambda *shape, dt=f64: torch.zeros(*shape, dtype=dt, device=dev)
        # ---- initial live points (eager; redrawn until lnprob is finite) ----
        star_k = torch.arange(S, device=dev, dtype=torch.int32).repeat_interleave(K)
        per_star = S > 1

        def lnprob_eager(u, slots):
            th = self.P.priortrans_batch(u)
            lnl, der, _ = eng.lnlike_batch(th, self.L._colmap, self.L._fixed, self.L._flags,
                                           star_slot=slots if per_star else None, want_derived=True)
            lp = self.P.lnprob_batch(th, lnl, der, star_slot=slots if self.star_plx is not None else None,
                                     star_plx=self.star_plx)
            return lp, th, der
        U = torch.rand(S * K, nd, generator=gen, device=dev, dtype=f64)
        lp, th, der = lnprob_eager(U, star_k)
        for _ in range(50):
            bad = ~torch.isfinite(lp)
            if not bool(bad.any()):
                break
            idx = bad.nonzero(as_tuple=True)[0]
            un = torch.rand(len(idx), nd, generator=gen, device=dev, dtype=f64)
            lpn, thn, dern = lnprob_eager(un, star_k[idx])
            U[idx], lp[idx], th[idx], der[idx] = un, lpn, thn, dern
        if bool((~torch.isfinite(lp)).any()):
            raise RuntimeError('could not draw live points with finite lnprob from the prior')
        t = dict(live_u=U.contiguous(), live_lp=lp.contiguous(), live_row=torch.cat([th, der], 1).contiguous(),
                 cur_u=z(S * Q, nd), cur_lp=z(S * Q), cur_row=z(S * Q, ncol), prop_u=z(S * Q, nd),
                 prop_lp=z(S * Q), prop_theta=z(S * Q, nd), prop_der=z(S * Q, nder), lnl=z(S * Q),
                 nacc=z(S * Q, dt=torch.int32), lmin=z(S), sigma=z(S, nd * nd), scale=torch.ones(S, dtype=f64, device=dev),
                 logz=torch.full((S,), -float('inf'), dtype=f64, device=dev), hacc=z(S),
                 wmax=torch.full((S,), -float('inf'), dtype=f64, device=dev), wtot=z(S), wsum=z(S, ncol), wsq=z(S, ncol),
                 niter=z(S, dt=torch.int64), ncall=z(S, dt=torch.int64), done=z(S, dt=torch.int32),
                 rng_round=z(1, dt=torch.int64), center=z(S, nd), ell_r=z(S))
        if self.store:
            t['dead'] = z(S, self.dead_cap, ncol + _lib.MS_NS_TRAIL)
        orig = torch.arange(S, device=dev)
        out = dict(logz=z(S), H=z(S), niter=z(S, dt=torch.int64), ncall=z(S, dt=torch.int64), mean=z(S, ncol),
                   std=z(S, ncol), converged=z(S, dt=torch.bool))
        seed = C.c_uint64(self.seed * 7919 + 17)
        stream = lambda: eng._stream()
        per_star_keys = ('live_u', 'live_lp', 'live_row', 'lmin', 'sigma', 'scale', 'logz', 'hacc', 'wmax', 'wtot',
                         'wsum', 'wsq', 'niter', 'ncall', 'done', 'dead', 'center', 'ell_r')
        walker_keys = ('cur_u', 'cur_lp', 'cur_row', 'prop_u', 'prop_lp', 'prop_theta', 'prop_der', 'lnl', 'nacc')

        def harvest(mask):
            """Copy the summaries of the (finalised) stars in ``mask`` to the catalog-wide outputs."""
            o = orig[mask]
            mean = t['wsum'][mask] / t['wtot'][mask][:, None]
            var = (t['wsq'][mask] / t['wtot'][mask][:, None] - mean * mean).clamp_min(0.0)
            out['logz'][o] = t['logz'][mask]; out['H'][o] = t['hacc'][mask] - t['logz'][mask]
            out['niter'][o] = t['niter'][mask]; out['ncall'][o] = t['ncall'][mask]
            out['mean'][o] = mean; out['std'][o] = torch.sqrt(var)

        rounds, captures = 0, 0
        bootstrap, reseed = True, False
        t_start = time.perf_counter()

        def lap(what):
            if verbose:
                torch.cuda.synchronize(dev)
                print('    %-28s %.3f s' % (what, time.perf_counter() - t_start))
        with torch.cuda.device(dev):
            while True:
                Sc = len(orig)
                st = _lib.NsState()
                st.S, st.K, st.Q, st.nd, st.nder, st.walks = Sc, K, Q, nd, nder, self.walks
                st.maxiter, st.dead_cap, st.dlogz = self.maxiter, self.dead_cap if self.store else 0, self.dlogz
                st.unif = 1 if self.sample == 'unif' else 0
                for name, typ in _lib.NsState._fields_[9:]:
                    if typ is C.c_void_p:
                        setattr(st, name, t[name].data_ptr() if name in t else None)
                walker_slot = orig.to(torch.int32).repeat_interleave(Q).contiguous()

                def consume():
                    _lib.check(lib.ms_ns_consume(eng._ctx, C.byref(st), seed, stream()))

                P = self.P
                slot_plx = walker_slot if self.star_plx is not None else None

                def round_fused():
                    for step in range(self.walks):
                        _lib.check(lib.ms_ns_propose_transform(eng._ctx, C.byref(st), seed, C.c_uint64(step), P._cols,
                                                               stream()))
                        eng.lnlike_batch(t['prop_theta'], self.L._colmap, self.L._fixed, self.L._flags,
                                         star_slot=walker_slot if per_star else None, want_derived=True,
                                         out=t['lnl'], derived_out=t['prop_der'])
                        _lib.check(lib.ms_ns_lnprob_accept(
                            eng._ctx, C.byref(st), C.byref(P._flags), P._colmap.ctypes.data_as(_lib.c_ip),
                            P._fixed.ctypes.data_as(_lib.c_dp), P._terms_c, P._nterms, int(P.imf_bool), t['lnl'].data_ptr(),
                            slot_plx.data_ptr() if slot_plx is not None else None,
                            self.star_plx.data_ptr() if self.star_plx is not None else None,
                            C.byref(P._adv) if P._adv is not None else None, stream()))
                    consume()

                def round_():
                    if self.fused:
                        return round_fused()
                    for step in range(self.walks):
                        _lib.check(lib.ms_ns_propose(eng._ctx, C.byref(st), seed, C.c_uint64(step), stream()))
                        self.P.priortrans_batch(t['prop_u'], out=t['prop_theta'])
                        eng.lnlike_batch(t['prop_theta'], self.L._colmap, self.L._fixed, self.L._flags,
                                         star_slot=walker_slot if per_star else None, want_derived=True,
                                         out=t['lnl'], derived_out=t['prop_der'])
                        self.P.lnprob_batch(t['prop_theta'], t['lnl'], t['prop_der'], out=t['prop_lp'],
                                            star_slot=walker_slot if self.star_plx is not None else None,
                                            star_plx=self.star_plx)
                        _lib.check(lib.ms_ns_accept(eng._ctx, C.byref(st), stream()))
                    consume()

                if bootstrap:                             # spread, threshold, first restart points
                    consume()
                    t['scale'].fill_(1.0)
                    bootstrap = False
                elif reseed:                              # new queue after a compaction: restart points only (an empty
                    keep_scale, keep_calls = t['scale'].clone(), t['ncall'].clone()      # queue consumes nothing)
                    consume()
                    t['scale'].copy_(keep_scale); t['ncall'].copy_(keep_calls)
                reseed = False
                lap('state ready')
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):
                    round_()                              # warm-up (scratch allocation), a real round
                torch.cuda.current_stream(dev).wait_stream(side)
                torch.cuda.synchronize(dev)
                lap('warm-up round')
                graph = _capture_graph(round_, dev)
                lap('graph captured')
                rounds += 2
                captures += 1
                while True:
                    for _ in range(check_every):
                        graph.replay()
                    rounds += check_every
                    nrun = int((t['done'] == 0).sum().item())
                    if verbose:
                        print('round %d: %d / %d stars running (working set %d, %d walkers per star, %.2f s)'
                              % (rounds, nrun, S, Sc, Q, time.perf_counter() - t_start))
                    if nrun == 0 or (Sc > min_compact and nrun <= compact_frac * Sc) or rounds >= self.maxiter + 16:
                        break
                del graph
                if rounds >= self.maxiter + 16:
                    t['done'][t['done'] == 0] = 3         # give up on the rest: finalise as they are, NOT converged
                    nrun = 0
                lap('replays done')
                _lib.check(lib.ms_ns_finalize(eng._ctx, C.byref(st), stream()))
                lap('finalize')
                fin = (t['done'] == 2) | (t['done'] == 4)
                harvest(fin)
                out['converged'][orig[fin]] = t['done'][fin] == 2        # 2: stopped by the dlogz test; 4: at maxiter
                if self.store and 'dead' in t:
                    pass                                  # dead points are read back below (single working set)
                if nrun == 0:
                    break
                keep = (t['done'] == 0).nonzero(as_tuple=True)[0]
                nk = len(keep)
                for kname in ('live_u', 'live_lp', 'live_row'):          # [S*K, ...] blocks of K rows per star
                    v = t[kname]
                    t[kname] = v.reshape((Sc, K) + tuple(v.shape[1:]))[keep].reshape((nk * K,) + tuple(v.shape[1:])).contiguous()
                # the walkers restart from live points every round, so the queue can be re-dimensioned at a compaction:
                # as the working set shrinks to the stragglers, every star gets more queued walkers (fewer sequential
                # rounds for the tail, batches that still fill the GPU); the walker budget of the first round is kept
                q_new = int(min(self.max_queue, max(Q, (self.S * self.Q) // max(nk, 1))))
                for kname in walker_keys:                                 # [S*Q, ...] blocks of Q rows per star
                    v = t[kname]
                    t[kname] = torch.zeros((nk * q_new,) + tuple(v.shape[1:]), dtype=v.dtype, device=dev)
                for kname in per_star_keys:                               # [S, ...]
                    if kname in t and not kname.startswith('live'):
                        t[kname] = t[kname][keep].contiguous()
                orig = orig[keep]
                Q = q_new
                reseed = True
            torch.cuda.synchronize(dev)
        res = dict(logz=out['logz'].cpu().numpy(), logzerr=torch.sqrt(out['H'].clamp_min(0.0) / K).cpu().numpy(),
                   information=out['H'].cpu().numpy(), niter=out['niter'].cpu().numpy(), ncall=out['ncall'].cpu().numpy(),
                   names=list(self.L.fitpars_i) + list(_lib.DERIVED_NAMES), mean=out['mean'].cpu().numpy(),
                   std=out['std'].cpu().numpy(), converged=out['converged'].cpu().numpy(), rounds=rounds,
                   graph_captures=captures)
        if self.store:
            d = t['dead'].cpu().numpy()
            res['dead'] = [d[s_, :min(int(res['niter'][s_]), self.dead_cap)] for s_ in range(S)]
            # the final live points (the reference appends them to the samples file, fitstar.py:466-501)
            res['live_rows'] = t['live_row'].reshape(S, K, ncol).cpu().numpy()
            res['live_lnprob'] = t['live_lp'].reshape(S, K).cpu().numpy()
            res['nlive'] = K
        return res


def _running_evidence(lnl, logwt, logz0=-np.inf, hacc0=0.0):
    """Running ln Z and information H after each point (same recursion as ns_accumulate in csrc/sampler.cu)."""
    logz, h = np.empty(len(lnl)), np.empty(len(lnl))
    lz, ha = logz0, hacc0
    for i, (l, w) in enumerate(zip(lnl, logwt)):
        lz_new = np.logaddexp(lz, w)
        ha = (ha * np.exp(lz - lz_new) if np.isfinite(lz) else 0.0) + np.exp(w - lz_new) * l
        lz = lz_new
        logz[i], h[i] = lz, ha - lz
    return logz, h


def samples_table(res, star=0, likeobj=None):
    """(column names, rows[n, ncols]) of one star's samples file: every dead point in order, then the final live
    points sorted by ln L (the reference appends them, fitstar.py:466-501), with the seven trailing columns
    ``log(lk) log(vol) log(wt) h nc log(z) delta(log(z))`` (fitstar.py:235-242, 424-425).

    With ``likeobj`` the parameter columns are the keys of the reference's ``parsdict`` in its insertion order
    (likelihood.py:95-141: sampled parameters, fixed parameters, then Teff, log(g), log(R), [Fe/H], [a/Fe],
    log(Age), Mass, log(L), Agewgt for isochrone fits) -- the header ``fitstar._initoutput`` writes and
    ``demo/postproc.py:142-160`` reads; without it the raw sampler columns are kept.
    Accepts the records of both samplers: one ``[niter, ncol + 7]`` array per star (QueuedNestedSampler, running
    values from ``ns_consume``) or per-iteration tuples (BatchedNestedSampler; running values recomputed here)."""
    names = list(res['names'])
    ncol = len(names)
    dead = res['dead']
    if len(dead) and isinstance(dead[0], np.ndarray) and dead[0].ndim == 2:
        d = np.asarray(dead[star])
        rows, lnl, logvol, logwt = d[:, :ncol], d[:, ncol], d[:, ncol + 1], d[:, ncol + 2]
        if d.shape[1] >= ncol + 7:
            h, nc, logz, dlz = d[:, ncol + 3], d[:, ncol + 4], d[:, ncol + 5], d[:, ncol + 6]
        else:
            logz, h = _running_evidence(lnl, logwt)
            nc, dlz = np.full(len(d), np.nan), np.full(len(d), np.nan)
    else:
        pick = [(row[np.flatnonzero(act == star)[0]], l[np.flatnonzero(act == star)[0]],
                 v[np.flatnonzero(act == star)[0]], w[np.flatnonzero(act == star)[0]])
                for act, row, l, v, w in dead if np.any(act == star)]
        rows = np.array([p[0] for p in pick]).reshape(len(pick), ncol)
        lnl, logvol, logwt = (np.array([p[i] for p in pick]) for i in (1, 2, 3))
        logz, h = _running_evidence(lnl, logwt)
        nc = np.full(len(pick), float(res.get('walks', np.nan)))
        dlz = np.full(len(pick), np.nan)
    if 'live_rows' in res:                                   # final live points, ascending ln L
        K = int(res['nlive'])
        lp = np.asarray(res['live_lnprob'][star])
        order = np.argsort(lp, kind='stable')
        lrows, llnl = np.asarray(res['live_rows'][star])[order], lp[order]
        lv_f = (logvol[-1] if len(logvol) else 0.0)
        llogvol = lv_f + np.log1p(-(np.arange(K) + 1.0) / (K + 1.0))
        llogwt = llnl + lv_f - np.log(K)                   # every live point carries X_final / K (ns_finalize)
        lz0 = logz[-1] if len(logz) else -np.inf
        llogz, lh = _running_evidence(llnl, llogwt, lz0, (h[-1] + lz0) if len(h) else 0.0)
        rest = np.append(np.logaddexp.accumulate(llogwt[::-1])[::-1][1:], -np.inf)
        ldlz = np.logaddexp(llogz, rest) - llogz
        rows = np.concatenate([rows, lrows]); lnl = np.concatenate([lnl, llnl])
        logvol = np.concatenate([logvol, llogvol]); logwt = np.concatenate([logwt, llogwt])
        h = np.concatenate([h, lh]); nc = np.concatenate([nc, np.ones(K)])
        logz = np.concatenate([logz, llogz]); dlz = np.concatenate([dlz, ldlz])
    if likeobj is not None:
        nd = likeobj.ndim
        dicts = [likeobj.parsdict_from(r[:nd], r[nd:] if likeobj.iso_bool else None) for r in rows]
        names = list(dicts[0].keys()) if dicts else list(likeobj.parsdict_from(np.zeros(nd), np.ones(ncol - nd)
                                                                               if likeobj.iso_bool else None).keys())
        rows = np.array([[float(dd[k]) if isinstance(dd[k], (int, float, np.floating)) else np.nan for k in names]
                         for dd in dicts]).reshape(len(dicts), len(names))
    trail = np.column_stack([lnl, logvol, logwt, h, nc, logz, dlz])
    cols = ['Iter'] + names + ['log(lk)', 'log(vol)', 'log(wt)', 'h', 'nc', 'log(z)', 'delta(log(z))']
    return cols, np.column_stack([np.arange(len(rows)), rows, trail])


def write_samples_file(path, res, star=0, likeobj=None):
    """One star's samples file in the reference's layout (see ``samples_table``); needs store_samples."""
    cols, tab = samples_table(res, star=star, likeobj=likeobj)
    with open(path, 'w') as f:
        f.write(' '.join(cols) + '\n')                      # header of fitstar._initoutput (fitstar.py:235-242)
        for r in tab:
            f.write('%d ' % int(r[0]) + ' '.join(repr(float(v)) for v in r[1:-7]) +
                    ' {0} {1} {2} {3} {4} {5} {6} \n'.format(*[repr(float(v)) for v in r[-7:-3]] +
                                                              [repr(int(r[-3])) if np.isfinite(r[-3]) else 'nan'] +
                                                              [repr(float(v)) for v in r[-2:]]))
