// Batched prior transform and ln-prior: the prior side of
// fitstar.lnprobfn, so that a batched sampler never leaves the GPU.
//   prior.priortrans / priortrans_mist / _spec / _phot     reference minesweeper/prior.py:150-385
//   prior.lnpriorfn / lnprior_mist / _spec / _phot          prior.py:387-639
//   AdvancedPriors.imf_lnprior (Kroupa-like broken power law)   advancedpriors.py:93-137
//   fitstar.lnprobfn (lnL = -inf short circuit, lnprior = -inf short circuit, sum)   fitstar.py:715-727
// Element-wise f64 kernels, one thread per point; the host (minesweeper_b200/prior.py) turns the
// reference's priordict into the per-column / per-term tables below with the reference's
// precedence rules.  Not supported on the batched path (host raises): beta, texp (broken in the
// reference, prior.py:367), GAL / GALAGE / VROT / VTOT / ALPHA / AngDia.
#include "common.cuh"

namespace {

struct TransformArgs { int ndim; ms_prior_col col[MS_NPARAM]; };

__device__ __forceinline__ double norm_cdf(double x) { return 0.5 * erfc(-x * 0.70710678118654752440); }

__global__ void prior_transform_kernel(TransformArgs a, const double *__restrict__ u, int64_t B,
                                       double *__restrict__ theta) {
    const int64_t n = B * a.ndim;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(t % a.ndim);
        const ms_prior_col &pc = a.col[c];
        const double x = u[t];
        double v;
        switch (pc.kind) {
            case MS_PRIOR_UNIFORM:                                   // (hi - lo) u + lo
                v = (pc.p[1] - pc.p[0]) * x + pc.p[0];
                break;
            case MS_PRIOR_GAUSSIAN:                                  // norm.ppf(u, loc, scale)
                v = pc.p[0] + pc.p[1] * normcdfinv(x);
                break;
            case MS_PRIOR_TGAUSSIAN: {                               // truncnorm.ppf(u, a, b, loc, scale); p = lo, hi, loc, scale
                const double al = (pc.p[0] - pc.p[2]) / pc.p[3], be = (pc.p[1] - pc.p[2]) / pc.p[3];
                double z;
                if (al > 0.0) {                                      // right tail: work with survival functions
                    const double sa = norm_cdf(-al), sb = norm_cdf(-be);
                    z = -normcdfinv(sa - x * (sa - sb));
                } else {
                    const double ca = norm_cdf(al), cb = norm_cdf(be);
                    z = normcdfinv(ca + x * (cb - ca));
                }
                v = pc.p[2] + pc.p[3] * z;
                if (v == ms_inf()) v = pc.p[1];                      // prior.py:193-194
                break;
            }
            case MS_PRIOR_EXP:                                       // expon.ppf(u, loc, scale)
                v = pc.p[0] - pc.p[1] * log1p(-x);
                break;
            case MS_PRIOR_LOGUNIFORM:                                // reciprocal.ppf(u, a, b)
                v = pc.p[0] * exp(x * log(pc.p[1] / pc.p[0]));
                break;
            default:
                v = ms_nan();
        }
        theta[t] = v;
    }
}

struct LnPriorArgs {
    ThetaView tv;
    const double *derived;          // [B][13] or nullptr
    const double *lnL;              // [B] or nullptr (then out = lnprior)
    double *out;
    int mistdif, nterms, imf;
    const int32_t *star_slot; const double *star_plx;   // per-star Gaussian parallax prior or nullptr
    ms_prior_term term[MS_MAX_PRIOR_TERMS];
};

__device__ __forceinline__ double prior_value(const LnPriorArgs &a, int64_t p, int id) {
    if (id < MS_NPARAM) {
        // sampled / fixed parameter, except the stellar labels an isochrone fit predicts
        if (a.derived && (id == MS_P_TEFF || id == MS_P_LOGG || id == MS_P_FEH || id == MS_P_AFE || id == MS_P_LOGR)) {
            const double *d = a.derived + p * MS_NDERIVED;
            switch (id) {
                case MS_P_TEFF: return pow(10.0, d[8]);
                case MS_P_LOGG: return d[11];
                case MS_P_FEH: return a.mistdif ? d[9] : d[2];
                case MS_P_AFE: return a.mistdif ? d[10] : d[3];
                default: return d[6];
            }
        }
        return theta_get(a.tv, p, id);
    }
    switch (id) {
        case MS_V_LOGAGE: return a.derived ? a.derived[p * MS_NDERIVED + 4] : ms_nan();
        case MS_V_AGE: return a.derived ? pow(10.0, a.derived[p * MS_NDERIVED + 4] - 9.0) : ms_nan();   // prior.py:526
        case MS_V_PARALLAX: return 1000.0 / theta_get(a.tv, p, MS_P_DIST);                           // prior.py:612
        default: return ms_nan();
    }
}

__global__ void lnprior_kernel(LnPriorArgs a, int64_t B) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < B; p += (int64_t)gridDim.x * blockDim.x) {
        const double neg_inf = -ms_inf();
        double l = a.lnL ? a.lnL[p] : 0.0;
        if (a.lnL && l == neg_inf) { a.out[p] = neg_inf; continue; }         // fitstar.py:719-720
        double lp = 0.0;
        if (a.imf) {                                                          // advancedpriors.py:93-137
            const double m = theta_get(a.tv, p, MS_P_INIT_MASS);
            const double al = 1.3, ah = 2.3, mb = 0.5;
            double v = neg_inf;
            if (m <= mb && m > 0.08) v = -al * log(m);
            else if (m > mb) v = -ah * log(m) + (ah - al) * log(mb);
            double nlo = pow(mb, 1.0 - al) / (ah - 1.0);
            double nhi = pow(0.08, 1.0 - al) / (al - 1.0) - pow(mb, 1.0 - al) / (al - 1.0);
            lp += v - log(nlo + nhi);
        }
        if (a.star_plx) {
            const double *pe = a.star_plx + 2 * (int64_t)(a.star_slot ? a.star_slot[p] : 0);
            const double d = 1000.0 / theta_get(a.tv, p, MS_P_DIST) - pe[0];
            lp += -0.5 * ((d * d) / (pe[1] * pe[1]));
        }
        bool dead = false;
        for (int i = 0; i < a.nterms && !dead; ++i) {
            const ms_prior_term &t = a.term[i];
            const double v = prior_value(a, p, t.value_id);
            switch (t.kind) {
                case MS_TERM_UNIFORM:                                         // bounds only
                    if (v < t.p[0] || v > t.p[1]) dead = true;
                    break;
                case MS_TERM_GAUSSIAN:
                    lp += -0.5 * (((v - t.p[0]) * (v - t.p[0])) / (t.p[1] * t.p[1]));
                    break;
                case MS_TERM_TGAUSSIAN:                                       // bounds + gaussian (prior.py:537-542)
                    if (v < t.p[0] || v > t.p[1]) dead = true;
                    else lp += -0.5 * (((v - t.p[2]) * (v - t.p[2])) / (t.p[3] * t.p[3]));
                    break;
                case MS_TERM_TGAUSSIAN_BOUNDS:                                // spec / phot families: bounds only (:575-578, 629-633)
                    if (v < t.p[0] || v > t.p[1]) dead = true;
                    break;
            }
        }
        if (dead || lp == neg_inf) { a.out[p] = neg_inf; continue; }          // fitstar.py:724-725
        a.out[p] = lp + l;
    }
}

}  // namespace

extern "C" int ms_prior_transform(ms_ctx *ctx, const double *u, int64_t B, int ndim, const ms_prior_col *cols,
                                  double *theta, void *stream) {
    if (ctx && B == 0) return MS_OK;
    if (!ctx || !u || !theta || !cols || B < 0 || ndim < 1 || ndim > MS_NPARAM) { ms_set_error("bad arguments"); return MS_EINVAL; }
    MS_CUDA(cudaSetDevice(ctx->device));
    TransformArgs a;
    a.ndim = ndim;
    for (int i = 0; i < ndim; ++i) {
        if (cols[i].kind < 0 || cols[i].kind > MS_PRIOR_LOGUNIFORM) { ms_set_error("column %d: unsupported prior kind %d", i, cols[i].kind); return MS_EINVAL; }
        a.col[i] = cols[i];
    }
    cudaStream_t st = (cudaStream_t)stream;
    int64_t blocks = (B * ndim + 255) / 256, cap = (int64_t)ctx->sm_count * 8;
    KernelTimer timer_(ctx, MS_K_PARS, st);
    prior_transform_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, st>>>(a, u, B, theta);
    ctx->launches++;
    MS_CUDA(cudaGetLastError());
    return MS_OK;
}

extern "C" int ms_lnprob_batch(ms_ctx *ctx, const ms_flags *flags, const double *theta, int64_t B, int ndim,
                               const int32_t *colmap, const double *fixed, const double *derived,
                               const ms_prior_term *terms, int nterms, int imf, const double *lnL, double *out,
                               void *stream) {
    return ms_lnprob_batch_stars(ctx, flags, theta, B, ndim, colmap, fixed, derived, terms, nterms, imf, lnL, out,
                                 nullptr, nullptr, stream);
}

extern "C" int ms_lnprob_batch_stars(ms_ctx *ctx, const ms_flags *flags, const double *theta, int64_t B, int ndim,
                                     const int32_t *colmap, const double *fixed, const double *derived,
                                     const ms_prior_term *terms, int nterms, int imf, const double *lnL,
                                     double *out, const int32_t *star_slot, const double *star_plx, void *stream) {
    if (ctx && B == 0) return MS_OK;
    if (!ctx || !flags || !theta || !colmap || !out || B < 0 || ndim < 1 || ndim > MS_NPARAM || nterms < 0 ||
        nterms > MS_MAX_PRIOR_TERMS || (nterms && !terms)) { ms_set_error("bad arguments"); return MS_EINVAL; }
    if (B == 0) return MS_OK;
    MS_CUDA(cudaSetDevice(ctx->device));
    LnPriorArgs a;
    a.tv.theta = theta; a.tv.ld = ndim;
    for (int i = 0; i < MS_NPARAM; ++i) { a.tv.col[i] = -1; a.tv.fixed[i] = fixed ? fixed[i] : nan(""); }
    for (int c = 0; c < ndim; ++c) {
        if (colmap[c] < 0 || colmap[c] >= MS_NPARAM) { ms_set_error("colmap[%d] out of range", c); return MS_EINVAL; }
        a.tv.col[colmap[c]] = c;
    }
    a.derived = derived; a.lnL = lnL; a.out = out; a.mistdif = flags->mistdif; a.nterms = nterms; a.imf = imf;
    a.star_slot = star_slot; a.star_plx = star_plx;
    for (int i = 0; i < nterms; ++i) {
        if (terms[i].value_id < 0 || terms[i].value_id >= MS_V_END || terms[i].kind < 0 || terms[i].kind > MS_TERM_TGAUSSIAN_BOUNDS) {
            ms_set_error("prior term %d is malformed", i);
            return MS_EINVAL;
        }
        a.term[i] = terms[i];
    }
    cudaStream_t st = (cudaStream_t)stream;
    int64_t blocks = (B + 255) / 256, cap = (int64_t)ctx->sm_count * 8;
    KernelTimer timer_(ctx, MS_K_FINALIZE, st);
    lnprior_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, st>>>(a, B);
    ctx->launches++;
    MS_CUDA(cudaGetLastError());
    return MS_OK;
}
