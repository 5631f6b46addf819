// Batched prior transform and ln-prior: the prior side of
// fitstar.lnprobfn, so that a batched sampler never leaves the GPU.
//   prior.priortrans / priortrans_mist / _spec / _phot     reference minesweeper/prior.py:150-385
//   prior.lnpriorfn / lnprior_mist / _spec / _phot          prior.py:387-639
//   AdvancedPriors.imf_lnprior (Kroupa-like broken power law)   advancedpriors.py:93-137
//   fitstar.lnprobfn (lnL = -inf short circuit, lnprior = -inf short circuit, sum)   fitstar.py:715-727
// Element-wise f64 kernels, one thread per point; the host (minesweeper_b200/prior.py) turns the
// reference's priordict into the per-column / per-term tables below with the reference's
// precedence rules.  Also here: scipy.stats.beta.ppf (pv_beta), the tabulated Galactic distance quantile
// (AdvancedPriors.gal_ppf) and the advanced additive priors GAL / GALAGE / VROT / ALPHA
// (advancedpriors.py:410-833).  Not supported (host raises): texp (NameError in the reference itself, prior.py:367).

#include "prior_dev.cuh"

namespace {

__global__ void prior_transform_kernel(TransformArgs a, const double *__restrict__ u, int64_t B,
                                       double *__restrict__ theta) {
    const int64_t n = B * a.ndim;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
        theta[t] = prior_transform_value(a, (int)(t % a.ndim), u[t]);
}

__global__ void lnprior_kernel(LnPriorArgs a, int64_t B) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < B; p += (int64_t)gridDim.x * blockDim.x)
        a.out[p] = lnprob_value(a, p);
}

}  // namespace

extern "C" int ms_prior_transform(ms_ctx *ctx, const double *u, int64_t B, int ndim, const ms_prior_col *cols,
                                  double *theta, void *stream) {
    if (ctx && B == 0) return MS_OK;
    if (!ctx || !u || !theta || !cols || B < 0 || ndim < 1 || ndim > MS_NPARAM) { ms_set_error("bad arguments"); return MS_EINVAL; }
    MS_CUDA(cudaSetDevice(ctx->device));
    TransformArgs a;
    { const int rc = fill_transform_args(ctx, ndim, cols, a); if (rc) return rc; }
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

extern "C" int ms_set_prior_table(ms_ctx *ctx, const double *x, const double *cdf, int n) {
    if (!ctx || n < 0 || (n > 0 && (n < 2 || !x || !cdf))) { ms_set_error("bad arguments"); return MS_EINVAL; }
    MS_CUDA(cudaSetDevice(ctx->device));
    cudaFree(ctx->ptab_x); cudaFree(ctx->ptab_cdf);
    ctx->ptab_x = ctx->ptab_cdf = nullptr; ctx->ptab_n = 0;
    if (n > 0) {
        MS_CUDA(cudaMalloc((void **)&ctx->ptab_x, (size_t)n * sizeof(double)));
        MS_CUDA(cudaMalloc((void **)&ctx->ptab_cdf, (size_t)n * sizeof(double)));
        MS_CUDA(cudaMemcpy(ctx->ptab_x, x, (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
        MS_CUDA(cudaMemcpy(ctx->ptab_cdf, cdf, (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
        ctx->ptab_n = n;
    }
    return MS_OK;
}

extern "C" int ms_lnprob_batch_stars(ms_ctx *ctx, const ms_flags *flags, const double *theta, int64_t B, int ndim,
                                     const int32_t *colmap, const double *fixed, const double *derived,
                                     const ms_prior_term *terms, int nterms, int imf, const double *lnL,
                                     double *out, const int32_t *star_slot, const double *star_plx, void *stream) {
    return ms_lnprob_batch_adv(ctx, flags, theta, B, ndim, colmap, fixed, derived, terms, nterms, imf, lnL, out,
                               star_slot, star_plx, nullptr, stream);
}

extern "C" int ms_lnprob_batch_adv(ms_ctx *ctx, const ms_flags *flags, const double *theta, int64_t B, int ndim,
                                   const int32_t *colmap, const double *fixed, const double *derived,
                                   const ms_prior_term *terms, int nterms, int imf, const double *lnL,
                                   double *out, const int32_t *star_slot, const double *star_plx,
                                   const ms_adv_prior *adv, void *stream) {
    if (ctx && B == 0) return MS_OK;
    if (!ctx || B < 0) { ms_set_error("bad arguments"); return MS_EINVAL; }
    MS_CUDA(cudaSetDevice(ctx->device));
    LnPriorArgs a;
    { const int rc = fill_lnprior_args(flags, theta, ndim, colmap, fixed, derived, terms, nterms, imf, lnL, out, star_slot,
                                       star_plx, adv, a); if (rc) return rc; }
    cudaStream_t st = (cudaStream_t)stream;
    int64_t blocks = (B + 255) / 256, cap = (int64_t)ctx->sm_count * 8;
    KernelTimer timer_(ctx, MS_K_FINALIZE, st);
    lnprior_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, st>>>(a, B);
    ctx->launches++;
    MS_CUDA(cudaGetLastError());
    return MS_OK;
}
