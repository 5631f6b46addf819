// Nested-sampling bookkeeping kernels: the sequential part of a
// static nested-sampling run with a proposal queue, one CTA per star.
//
// The reference drives dynesty's static sampler with the `rwalk` proposal, one likelihood call at
// a time (reference fitstar.py:337-369; demo/etc/samplerinfo.dat).  Here every star owns Q
// random-walk walkers that are advanced together (one batched prior-transform / likelihood /
// ln-prior call per walk step over all stars x walkers); their end points form a queue that
// `ns_consume_kernel` uses up in order: each accepted candidate replaces the current worst live
// point (ln X_i = -i / K), provided it still lies above the rising likelihood threshold --
// exactly dynesty's `queue_size` logic: a point uniform in {L > L_a} that also satisfies
// L > L_b >= L_a is uniform in {L > L_b}.
//   ns_propose_kernel   Gaussian step in the unit cube, reflected at the walls (counter-based RNG)
//   ns_accept_kernel    accept iff lnprob > threshold of the round; carries theta / derived along
//   ns_consume_kernel   worst point, ln Z / H / posterior moments, replacement, stopping rule,
//                       live-point spread, step-scale adaptation, restart points of the next round
//   ns_finalize_kernel  contribution of the remaining live points
// Fused forms used by the catalog sampler (one launch each instead of two; same arithmetic, bit-identical results):
//   ns_propose_kernel<true>      propose + prior transform (theta leaves the kernel that drew the step)
//   ns_lnprob_accept_kernel      lnL + ln prior (lnprobfn) + accept
#include "prior_dev.cuh"


namespace {

__device__ __forceinline__ double logaddexp_d(double a, double b) {
    if (a == -ms_inf()) return b;
    if (b == -ms_inf()) return a;
    const double m = fmax(a, b);
    return m + log1p(exp(-fabs(a - b)));
}

// counter-based generator: splitmix64 finaliser of a linear mix of (seed, stream, counter)
__device__ __forceinline__ uint64_t ns_mix(uint64_t seed, uint64_t stream, uint64_t counter) {
    uint64_t z = seed + stream * 0xbf58476d1ce4e5b9ull + counter * 0x94d049bb133111ebull + 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

// FUSE: also apply the prior transform to every coordinate and write theta (prop_theta) -- what
// prior_transform_kernel would do in a second pass over prop_u.
template <bool FUSE>
__global__ void ns_propose_kernel(ms_ns_state s, TransformArgs ta, uint64_t seed, uint64_t step) {
    double *theta_out = const_cast<double *>(s.prop_theta);
    const int64_t n = (int64_t)s.S * s.Q;
    for (int64_t wk = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; wk < n; wk += (int64_t)gridDim.x * blockDim.x) {
        const int star = (int)(wk / s.Q);
        const uint64_t ctr = (*s.rng_round * (uint64_t)s.walks + step) * 16ull;
        const double sc = s.scale[star];
        // step = scale * L z with L the Cholesky factor of the live-point covariance (the analogue of
        // dynesty's rwalk proposing along the axes of the bounding ellipsoid)
        double zv[MS_NPARAM];
        for (int d = 0; d < s.nd; d += 2) {
            // Box-Muller on two 32-bit uniforms in (0, 1)
            const uint64_t z = ns_mix(seed, (uint64_t)wk, ctr + (uint64_t)(d >> 1));
            const double u1 = ((double)(uint32_t)(z >> 32) + 0.5) * 2.3283064365386963e-10;
            const double u2 = ((double)(uint32_t)z + 0.5) * 2.3283064365386963e-10;
            const double rad = sqrt(-2.0 * log(u1));
            double sn, cs;
            sincospi(2.0 * u2, &sn, &cs);
            zv[d] = rad * cs;
            if (d + 1 < s.nd) zv[d + 1] = rad * sn;
        }
        const double *Lc = s.sigma + (int64_t)star * s.nd * s.nd;
        if (s.unif) {
            // dynesty's 'unif': a point uniform in the bounding ellipsoid center + r L B (B = unit ball): direction
            // from the Gaussian vector, radius r U^(1/nd); a point outside the unit cube is no proposal at all (NaN
            // coordinates: the prior transform, the likelihood and the acceptance test all pass NaN through as a
            // rejection) -- the rejection of the volume outside the cube that dynesty does by redrawing
            double n2 = 0.0;
            for (int d = 0; d < s.nd; ++d) n2 += zv[d] * zv[d];
            const uint64_t zz = ns_mix(seed ^ 0x2545f4914f6cdd1dull, (uint64_t)wk, ctr + 15ull);
            const double uu = ((double)(uint32_t)(zz >> 32) + 0.5) * 2.3283064365386963e-10;
            const double rad = s.ell_r[star] * pow(uu, 1.0 / s.nd) / sqrt(n2);
            const double *mu = s.center + (int64_t)star * s.nd;
            bool inside = true;
            double uv[MS_NPARAM];
            for (int d = 0; d < s.nd; ++d) {
                double step = 0.0;
                for (int e = 0; e <= d; ++e) step = fma(Lc[d * s.nd + e], zv[e], step);
                uv[d] = mu[d] + rad * step;
                inside = inside && uv[d] > 0.0 && uv[d] < 1.0;
            }
            for (int d = 0; d < s.nd; ++d) {
                const double v = inside ? uv[d] : ms_nan();
                s.prop_u[wk * s.nd + d] = v;
                if (FUSE) theta_out[wk * s.nd + d] = prior_transform_value(ta, d, v);
            }
            continue;
        }
        for (int d = 0; d < s.nd; ++d) {
            double step = 0.0;
            for (int e = 0; e <= d; ++e) step = fma(Lc[d * s.nd + e], zv[e], step);
            double v = s.cur_u[wk * s.nd + d] + sc * step;
            v = fmod(v, 2.0);                       // reflect into [0, 1]
            if (v < 0.0) v += 2.0;
            if (v > 1.0) v = 2.0 - v;
            s.prop_u[wk * s.nd + d] = v;
            if (FUSE) theta_out[wk * s.nd + d] = prior_transform_value(ta, d, v);
        }
    }
}

// Half a WARP per walker, a lane per element of its record (nd unit-cube coordinates, nd parameters, nder derived values,
// the ln prob, the acceptance counter): consecutive lanes touch consecutive addresses.  The thread-per-walker form copied
// 31 doubles per thread with a 248-byte stride between neighbours (107 us per 400 k walkers in the launch list of the
// catalog run, 13 % of a sampler step; a warp per walker: 82 us).  Rejected walkers load nothing beyond their ln prob.
__global__ void ns_accept_kernel(ms_ns_state s) {
    // half a warp per walker (two independent walkers per warp iteration), a lane per element of the record
    const int nd = s.nd, ncol = s.nd + s.nder, w = nd + ncol + 1;        // elements per walker
    const int hl = threadIdx.x & 15;
    const int64_t hpb = blockDim.x >> 4, n = (int64_t)s.S * s.Q;
    for (int64_t wk = blockIdx.x * hpb + (threadIdx.x >> 4); wk < n; wk += (int64_t)gridDim.x * hpb) {
        const double lp = s.prop_lp[wk];
        if (!(lp > s.lmin[wk / s.Q])) continue;
        for (int c = hl; c < w; c += 16) {
            if (c < ncol) {
                s.cur_row[wk * ncol + c] = c < nd ? s.prop_theta[wk * nd + c] : s.prop_der[wk * s.nder + (c - nd)];
            } else if (c < ncol + nd) {
                s.cur_u[wk * nd + (c - ncol)] = s.prop_u[wk * nd + (c - ncol)];
            } else {
                s.cur_lp[wk] = lp;
                s.nacc[wk] += 1;
            }
        }
    }
}

// lnprob + accept in one pass: a lane per walker evaluates lnprobfn (lnL + ln prior with its -inf short circuits, the
// arithmetic of lnprior_kernel), then the warp copies the records of its accepted walkers (ballot) cooperatively, a lane
// per element; same stores as ns_accept_kernel.  a.out (= prop_lp) is still written: ns_consume never reads it, the tests do.
__global__ void __launch_bounds__(256) ns_lnprob_accept_kernel(ms_ns_state s, LnPriorArgs a) {
    const int nd = s.nd, ncol = s.nd + s.nder, w = nd + ncol + 1;
    const int lane = threadIdx.x & 31;
    const int64_t n = (int64_t)s.S * s.Q;
    const int64_t wpb = blockDim.x >> 5;
    for (int64_t base = (blockIdx.x * wpb + (threadIdx.x >> 5)) * 32; base < n; base += (int64_t)gridDim.x * wpb * 32) {
        const int64_t mine = base + lane;
        double lp = -ms_inf();
        bool ok = false;
        if (mine < n) {
            lp = lnprob_value(a, mine);
            a.out[mine] = lp;
            ok = lp > s.lmin[mine / s.Q];
        }
        // the accepted walkers of this warp, one after the other: the whole warp copies a record (26 elements for a
        // 6-parameter isochrone fit: one element per lane, consecutive addresses)
        unsigned todo = __ballot_sync(0xffffffffu, ok);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const double lpj = __shfl_sync(0xffffffffu, lp, src);
            const int64_t wk = base + src;
            for (int c = lane; c < w; c += 32) {
                if (c < ncol) {
                    s.cur_row[wk * ncol + c] = c < nd ? s.prop_theta[wk * nd + c] : s.prop_der[wk * s.nder + (c - nd)];
                } else if (c < ncol + nd) {
                    s.cur_u[wk * nd + (c - ncol)] = s.prop_u[wk * nd + (c - ncol)];
                } else {
                    s.cur_lp[wk] = lpj;
                    s.nacc[wk] += 1;
                }
            }
        }
    }
}

// ---- one warp per star ---------------------------------------------------------------------------
// The per-star bookkeeping is a chain of dependent global round trips and f64 logarithms with a few
// 150-wide or 19-wide loops in between, so a star gets one warp (no block barriers, results of the
// reductions in every lane) and a CTA carries NS_WARPS independent stars; what bounds the kernel is
// how many stars are in flight per SM.
constexpr int NS_WARPS = 4;

// Evidence state of one star, held in registers by every lane of its warp (loaded with one round of
// independent loads, written back once) so the update chain never waits on global memory.
struct NsEv {
    double logz, hacc, wmax, wtot;
    __device__ void load(const ms_ns_state &s, int star) {
        logz = s.logz[star]; hacc = s.hacc[star]; wmax = s.wmax[star]; wtot = s.wtot[star];
    }
    __device__ void store(const ms_ns_state &s, int star) const {
        s.logz[star] = logz; s.hacc[star] = hacc; s.wmax[star] = wmax; s.wtot[star] = wtot;
    }
};

// update of ln Z, H and the weighted posterior moments of one star with a point of ln weight logwt
// (all lanes compute the same scalars; the lanes share the columns of the moment sums)
__device__ __forceinline__ void ns_accumulate(const ms_ns_state &s, NsEv &ev, int star, double lnl, double logwt,
                                              const double *row, int ncol, int lane) {
    const double lz = ev.logz, lz_new = logaddexp_d(lz, logwt);
    double h = ev.hacc * exp(lz - lz_new) + exp(logwt - lz_new) * lnl;
    if (lz == -ms_inf()) h = exp(logwt - lz_new) * lnl;
    ev.hacc = h;
    ev.logz = lz_new;
    const double m_old = ev.wmax, m_new = fmax(m_old, logwt);
    const double r_old = (m_old == -ms_inf()) ? 0.0 : exp(m_old - m_new);
    const double w = exp(logwt - m_new);
    ev.wmax = m_new;
    ev.wtot = ev.wtot * r_old + w;
    for (int c = lane; c < ncol; c += 32) {
        double v = row[c];
        if (!(fabs(v) < 1e300)) v = 0.0;
        const int64_t at = (int64_t)star * ncol + c;
        s.wsum[at] = s.wsum[at] * r_old + w * v;
        s.wsq[at] = s.wsq[at] * r_old + w * v * v;
    }
}

// warp-wide arg-min (lowest index on ties) and max of lp[0..K); the result is in every lane
__device__ __forceinline__ void ns_minmax(const double *lp, int K, int lane, double &vmin, int &imin, double &vmax) {
    double mn = ms_inf(), mx = -ms_inf();
    int im = 0;
    for (int k = lane; k < K; k += 32) {
        const double v = lp[k];
        if (v < mn) { mn = v; im = k; }
        mx = fmax(mx, v);
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double omn = __shfl_xor_sync(0xffffffffu, mn, o);
        const int oim = __shfl_xor_sync(0xffffffffu, im, o);
        const double omx = __shfl_xor_sync(0xffffffffu, mx, o);
        if (omn < mn || (omn == mn && oim < im)) { mn = omn; im = oim; }
        mx = fmax(mx, omx);
    }
    vmin = mn; imin = im; vmax = mx;
}

// Mean, covariance and Cholesky factor (lower triangle, row-major [nd][nd]) of the K points lu[K][nd], by one warp:
// lane pe owns covariance entry pe (nd * nd <= 32 in practice, otherwise the lanes stride); a covariance that is not
// positive definite falls back to the diagonal of standard deviations.  mean[nd], diag[nd], cov[nd*nd] are the warp's
// scratch (shared memory); on return cov holds the factor that was written to `out`.
__device__ __forceinline__ void ns_spread(const double *lu, int K, int nd, int lane, double *mean, double *diag,
                                          double *cov, double *out) {
    for (int d = lane; d < nd; d += 32) {
        double a = 0.0;
        for (int k = 0; k < K; ++k) a += lu[k * nd + d];
        mean[d] = a / K;
    }
    __syncwarp();
    for (int pe = lane; pe < nd * nd; pe += 32) {
        const int d = pe / nd, e = pe - d * nd;
        if (e > d) { cov[pe] = 0.0; continue; }
        double a = 0.0;
        const double md = mean[d], me = mean[e];
        for (int k = 0; k < K; ++k) a += (lu[k * nd + d] - md) * (lu[k * nd + e] - me);
        cov[pe] = a / (K - 1);
    }
    __syncwarp();
    for (int d = lane; d < nd; d += 32) diag[d] = cov[d * nd + d];      // kept for the degenerate fallback
    __syncwarp();
    if (lane == 0) {                                     // in-place Cholesky, row by row, lower triangle
        bool ok = true;
        for (int i = 0; i < nd && ok; ++i)
            for (int j = 0; j <= i; ++j) {
                double a = cov[i * nd + j] + (i == j ? 1e-14 : 0.0);
                for (int k = 0; k < j; ++k) a -= cov[i * nd + k] * cov[j * nd + k];
                if (i == j) { if (!(a > 0.0)) { ok = false; break; } cov[i * nd + i] = sqrt(a); }
                else cov[i * nd + j] = a / cov[j * nd + j];
            }
        for (int i = 0; i < nd; ++i)
            for (int j = 0; j < nd; ++j) {
                const double v = ok ? (j <= i ? cov[i * nd + j] : 0.0)
                                    : (i == j ? sqrt(fmax(diag[i], 1e-24)) : 0.0);    // degenerate: diagonal
                out[i * nd + j] = v;
                cov[i * nd + j] = v;                     // the factor actually used, for the radius below
            }
    }
}

__global__ void __launch_bounds__(32 * NS_WARPS, 8)
ns_consume_kernel(ms_ns_state s, uint64_t seed, double log_shrink) {
    extern __shared__ double sm_all[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int star = blockIdx.x * NS_WARPS + wib;
    if (star >= s.S) return;
    const uint64_t round = *s.rng_round;
    const int K = s.K, Q = s.Q, nd = s.nd, ncol = s.nd + s.nder;
    const int per_warp = K + 2 * nd + nd * nd;             // lp, mean, variances, covariance
    double *lp = sm_all + wib * per_warp;                  // K live ln-probabilities
    double *mean = lp + K, *diag = mean + nd, *cov = diag + nd;
    const int64_t lbase = (int64_t)star * K;
    for (int k = lane; k < K; k += 32) lp[k] = s.live_lp[lbase + k];
    __syncwarp();
    bool done = s.done[star] != 0, hit_max = false;
    int nacc_tot = 0;
    if (!done) {
        // everything the update chain needs, requested together
        NsEv ev; ev.load(s, star);
        int64_t it = s.niter[star];
        const int64_t it_in = it;
        const double scale_in = s.scale[star];
        if (lane == 0 && round > 0) s.ncall[star] += (int64_t)Q * s.walks;           // the walk steps of this round
        for (int q = 0; q < Q; ++q) {
            const int64_t wk = (int64_t)star * Q + q;
            nacc_tot += s.nacc[wk];
            double lmin, lmax; int w;
            ns_minmax(lp, K, lane, lmin, w, lmax);
            const double lpq = s.cur_lp[wk];
            if (!(s.nacc[wk] > 0 && lpq > lmin)) continue;          // never moved, or overtaken by the threshold
            const double logvol = -((double)it + 1.0) / K;
            const double logwt = lmin + (logvol + 1.0 / K + log_shrink);
            const double *row_w = s.live_row + (lbase + w) * ncol;
            // optional dead-point record: the row, then the seven trailing columns of the reference's samples file
            // (fitstar.py:424-425): log(lk) log(vol) log(wt) h nc log(z) delta(log(z))
            double *o = (s.dead && it < s.dead_cap) ? s.dead + ((int64_t)star * s.dead_cap + it) * (ncol + MS_NS_TRAIL)
                                                    : nullptr;
            if (o) {
                for (int c = lane; c < ncol; c += 32) o[c] = row_w[c];
                if (lane == 0) { o[ncol] = lmin; o[ncol + 1] = logvol; o[ncol + 2] = logwt; }
            }
            ns_accumulate(s, ev, star, lmin, logwt, row_w, ncol, lane);
            __syncwarp();                                       // row_w read by every lane before it is replaced
            // the candidate replaces the worst live point
            for (int c = lane; c < ncol; c += 32) s.live_row[(lbase + w) * ncol + c] = s.cur_row[wk * ncol + c];
            for (int d = lane; d < nd; d += 32) s.live_u[(lbase + w) * nd + d] = s.cur_u[wk * nd + d];
            if (lane == 0) { lp[w] = lpq; s.live_lp[lbase + w] = lpq; }
            ++it;
            __syncwarp();
            lmax = fmax(lmax, lpq);
            const double dl = logaddexp_d(ev.logz, lmax + logvol) - ev.logz;
            if (o && lane == 0) {                                   // running information, calls, evidence, remaining evidence
                o[ncol + 3] = ev.hacc - ev.logz; o[ncol + 4] = (double)s.walks; o[ncol + 5] = ev.logz; o[ncol + 6] = dl;
            }
            if (dl < s.dlogz) { done = true; break; }
            if (it >= s.maxiter) { done = true; hit_max = true; break; }
        }
        if (lane == 0) {
            if (it != it_in) { ev.store(s, star); s.niter[star] = it; }
            s.done[star] = done ? (hit_max ? 3 : 1) : 0;     // 1 converged (dlogz), 3 stopped at maxiter
            // dynesty's rwalk scale update towards 50 % acceptance
            const double facc = (double)nacc_tot / ((double)Q * s.walks);
            const double sc = scale_in * exp((facc - 0.5) / nd / 0.5);
            s.scale[star] = fmin(fmax(sc, 1e-4), 10.0);
        }
    }
    __syncwarp();
    // covariance of the live points and its Cholesky factor (proposal shape of the next rounds);
    // at most Q of the K live points change per round, so it is refreshed every ~16 replacements.
    // Lane pe owns covariance entry pe (nd * nd <= 32 in practice, otherwise the lanes stride); the live
    // points stream through once per entry pair from L2, 7 KB per star.
    const uint64_t cov_every = (uint64_t)((Q >= 16 || s.unif) ? 1 : 16 / Q);
    if (round % cov_every == 0) {
        const double *lu = s.live_u + lbase * nd;
        ns_spread(lu, K, nd, lane, mean, diag, cov, s.sigma + (int64_t)star * nd * nd);
        __syncwarp();
        if (s.unif) {
            // bounding ellipsoid of the live points in the metric of the factor: r^2 = max_k |L^-1 (u_k - mean)|^2,
            // enlarged by 25 % in volume as dynesty does (its default `enlarge`)
            double r2 = 0.0;
            for (int k = lane; k < K; k += 32) {
                double y[MS_NPARAM], a2 = 0.0;
                for (int i = 0; i < nd; ++i) {
                    double a = lu[k * nd + i] - mean[i];
                    for (int j = 0; j < i; ++j) a -= cov[i * nd + j] * y[j];
                    y[i] = a / cov[i * nd + i];
                    a2 += y[i] * y[i];
                }
                r2 = fmax(r2, a2);
            }
            for (int o = 16; o > 0; o >>= 1) r2 = fmax(r2, __shfl_xor_sync(0xffffffffu, r2, o));
            if (lane == 0) s.ell_r[star] = sqrt(r2) * pow(1.25, 1.0 / nd);
            for (int d = lane; d < nd; d += 32) s.center[(int64_t)star * nd + d] = mean[d];
        }
        __syncwarp();
    }
    // threshold and restart points of the next round: random live points other than the worst
    double lmin, lmax; int w;
    ns_minmax(lp, K, lane, lmin, w, lmax);
    if (lane == 0) s.lmin[star] = lmin;
    for (int q = 0; q < Q; ++q) {                          // the warp copies one restart row at a time
        const int64_t wk = (int64_t)star * Q + q;
        // one draw per walker and round: a counter-based mix (splitmix64 finaliser) of (seed, walker, round)
        const uint64_t z = ns_mix(seed ^ 0x5851f42d4c957f2dull, (uint64_t)wk, round);
        int j = (int)(((z >> 32) * (uint64_t)(K - 1)) >> 32);       // uniform in [0, K - 1)
        if (j >= w) ++j;
        if (lane == 0) { s.cur_lp[wk] = lp[j]; s.nacc[wk] = 0; }
        for (int d = lane; d < nd; d += 32) s.cur_u[wk * nd + d] = s.live_u[(lbase + j) * nd + d];
        for (int c = lane; c < ncol; c += 32) s.cur_row[wk * ncol + c] = s.live_row[(lbase + j) * ncol + c];
    }
}

// Cholesky factors of the covariances of S point sets U[S][K][nd] (the proposal shape of the torch-tensor sampler)
__global__ void __launch_bounds__(32 * NS_WARPS)
ns_spread_kernel(const double *__restrict__ U, int S, int K, int nd, double *__restrict__ sigma) {
    extern __shared__ double sm_all[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int star = blockIdx.x * NS_WARPS + wib;
    if (star >= S) return;
    double *mean = sm_all + wib * (2 * nd + nd * nd), *diag = mean + nd, *cov = diag + nd;
    ns_spread(U + (int64_t)star * K * nd, K, nd, lane, mean, diag, cov, sigma + (int64_t)star * nd * nd);
}

__global__ void ns_tick_kernel(uint64_t *counter) { *counter += 1; }

__global__ void __launch_bounds__(32 * NS_WARPS)
ns_finalize_kernel(ms_ns_state s) {
    const int lane = threadIdx.x & 31;
    const int star = blockIdx.x * NS_WARPS + (threadIdx.x >> 5);
    if (star >= s.S || (s.done[star] != 1 && s.done[star] != 3)) return;   // only stars that stopped and are not yet finalised
    const int K = s.K, ncol = s.nd + s.nder;
    const double logvol_f = -((double)s.niter[star]) / K - log((double)K);
    NsEv ev; ev.load(s, star);
    for (int k = 0; k < K; ++k) {
        const double l = s.live_lp[(int64_t)star * K + k];
        ns_accumulate(s, ev, star, l, l + logvol_f, s.live_row + ((int64_t)star * K + k) * ncol, ncol, lane);
    }
    __syncwarp();
    if (lane == 0) { ev.store(s, star); s.done[star] = s.done[star] == 3 ? 4 : 2; }    // 2 finalised, 4 finalised after maxiter
}

int check_state(const ms_ns_state *s) {
    if (!s || s->S < 1 || s->K < 2 || s->Q < 1 || s->nd < 1 || s->nd > MS_NPARAM || s->nder < 0 || s->walks < 1 ||
        !s->live_u || !s->live_lp || !s->live_row || !s->cur_u || !s->cur_lp || !s->cur_row || !s->prop_u ||
        !s->prop_lp || !s->prop_theta || !s->nacc || !s->lmin || !s->sigma || !s->scale || !s->logz || !s->hacc ||
        !s->wmax || !s->wtot || !s->wsum || !s->wsq || !s->niter || !s->ncall || !s->done || !s->rng_round ||
        (s->nder && !s->prop_der) || (s->unif && (!s->center || !s->ell_r || s->walks != 1))) {
        ms_set_error("ms_ns_state is incomplete");
        return MS_EINVAL;
    }
    return MS_OK;
}

}  // namespace

extern "C" int ms_ns_propose(ms_ctx *ctx, const ms_ns_state *s, uint64_t seed, uint64_t step, void *stream) {
    int rc = check_state(s);
    if (rc || !ctx) return rc ? rc : MS_EINVAL;
    MS_CUDA(cudaSetDevice(ctx->device));
    const int64_t n = (int64_t)s->S * s->Q;
    int64_t blocks = (n + 127) / 128, cap = (int64_t)ctx->sm_count * 16;
    TransformArgs none;
    memset(&none, 0, sizeof(none));
    ns_propose_kernel<false><<<(unsigned)(blocks < cap ? blocks : cap), 128, 0, (cudaStream_t)stream>>>(*s, none, seed, step);
    ctx->launches++;
    MS_CUDA(cudaGetLastError());
    return MS_OK;
}

extern "C" int ms_ns_propose_transform(ms_ctx *ctx, const ms_ns_state *s, uint64_t seed, uint64_t step,
                                       const ms_prior_col *cols, void *stream) {
    int rc = check_state(s);
    if (rc || !ctx) return rc ? rc : MS_EINVAL;
    MS_CUDA(cudaSetDevice(ctx->device));
    TransformArgs ta;
    rc = fill_transform_args(ctx, s->nd, cols, ta);
    if (rc) return rc;
    const int64_t n = (int64_t)s->S * s->Q;
    int64_t blocks = (n + 127) / 128, cap = (int64_t)ctx->sm_count * 16;
    ns_propose_kernel<true><<<(unsigned)(blocks < cap ? blocks : cap), 128, 0, (cudaStream_t)stream>>>(*s, ta, seed, step);
    ctx->launches++;
    MS_CUDA(cudaGetLastError());
    return MS_OK;
}

extern "C" int ms_ns_lnprob_accept(ms_ctx *ctx, const ms_ns_state *s, const ms_flags *flags, const int32_t *colmap,
                                   const double *fixed, const ms_prior_term *terms, int nterms, int imf,
                                   const double *lnL, const int32_t *star_slot, const double *star_plx,
                                   const ms_adv_prior *adv, void *stream) {
    int rc = check_state(s);
    if (rc || !ctx) return rc ? rc : MS_EINVAL;
    if (!lnL) { ms_set_error("ms_ns_lnprob_accept needs the ln likelihood of the proposals"); return MS_EINVAL; }
    MS_CUDA(cudaSetDevice(ctx->device));
    LnPriorArgs a;
    rc = fill_lnprior_args(flags, s->prop_theta, s->nd, colmap, fixed, s->nder ? s->prop_der : nullptr, terms, nterms, imf,
                           lnL, const_cast<double *>(s->prop_lp), star_slot, star_plx, adv, a);
    if (rc) return rc;
    const int64_t n = (int64_t)s->S * s->Q;                              // a lane per walker, 256 walkers per CTA
    int64_t blocks = (n + 255) / 256, cap = (int64_t)ctx->sm_count * 8;
    ns_lnprob_accept_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(*s, a);
    ctx->launches++;
    MS_CUDA(cudaGetLastError());
    return MS_OK;
}

extern "C" int ms_ns_accept(ms_ctx *ctx, const ms_ns_state *s, void *stream) {
    int rc = check_state(s);
    if (rc || !ctx) return rc ? rc : MS_EINVAL;
    MS_CUDA(cudaSetDevice(ctx->device));
    const int64_t n = (int64_t)s->S * s->Q;                              // half a warp per walker, 16 walkers per CTA
    int64_t blocks = (n + 15) / 16, cap = (int64_t)ctx->sm_count * 8;
    ns_accept_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(*s);
    ctx->launches++;
    MS_CUDA(cudaGetLastError());
    return MS_OK;
}

extern "C" int ms_ns_consume(ms_ctx *ctx, const ms_ns_state *s, uint64_t seed, void *stream) {
    int rc = check_state(s);
    if (rc || !ctx) return rc ? rc : MS_EINVAL;
    MS_CUDA(cudaSetDevice(ctx->device));
    const size_t smem = (size_t)NS_WARPS * (s->K + 2 * s->nd + s->nd * s->nd) * sizeof(double);   // per star: lp, mean, variances, covariance
    const double log_shrink = std::log1p(-std::exp(-1.0 / s->K));             // ln(1 - e^{-1/K}): the shell width
    ns_consume_kernel<<<(unsigned)((s->S + NS_WARPS - 1) / NS_WARPS), 32 * NS_WARPS, smem, (cudaStream_t)stream>>>(*s, seed, log_shrink);
    ns_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(s->rng_round);
    ctx->launches += 2;
    MS_CUDA(cudaGetLastError());
    return MS_OK;
}

extern "C" int ms_ns_finalize(ms_ctx *ctx, const ms_ns_state *s, void *stream) {
    int rc = check_state(s);
    if (rc || !ctx) return rc ? rc : MS_EINVAL;
    MS_CUDA(cudaSetDevice(ctx->device));
    ns_finalize_kernel<<<(unsigned)((s->S + NS_WARPS - 1) / NS_WARPS), 32 * NS_WARPS, 0, (cudaStream_t)stream>>>(*s);
    ctx->launches++;
    MS_CUDA(cudaGetLastError());
    return MS_OK;
}

extern "C" int ms_ns_spread(ms_ctx *ctx, const double *U, int S, int K, int nd, double *sigma, void *stream) {
    if (ctx && S == 0) return MS_OK;
    if (!ctx || !U || !sigma || S < 0 || K < 2 || nd < 1 || nd > MS_NPARAM) { ms_set_error("bad arguments"); return MS_EINVAL; }
    MS_CUDA(cudaSetDevice(ctx->device));
    const size_t smem = (size_t)NS_WARPS * (2 * nd + nd * nd) * sizeof(double);
    ns_spread_kernel<<<(unsigned)((S + NS_WARPS - 1) / NS_WARPS), 32 * NS_WARPS, smem, (cudaStream_t)stream>>>(U, S, K, nd, sigma);
    ctx->launches++;
    MS_CUDA(cudaGetLastError());
    return MS_OK;
}
