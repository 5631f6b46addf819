// Small element-wise kernels around the forward model: assembling the per-point
// parameter block and turning chi-square into lnL.  Batched form of the
// bookkeeping in likelihood.lnlikefn (reference minesweeper/likelihood.py:95-176).
#include "common.cuh"

namespace {

__global__ void build_pars_kernel(ThetaView tv, const double *__restrict__ derived, int mistdif,
                                  int64_t B, double *__restrict__ pars, double *__restrict__ chi2) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < B;
         p += (int64_t)gridDim.x * blockDim.x) {
        double *o = pars + p * MS_NPARS;
        double teff, logg, feh, afe, logr, agew = 1.0;
        if (derived) {
            const double *d = derived + p * MS_NDERIVED;
            const double logt = d[8];
            if (fabs(logt) < 1e300) {
                teff = pow(10.0, logt);                       // likelihood.py:121
                logg = d[11]; logr = d[6];
                feh = mistdif ? d[9] : d[2];                   // :127-132
                afe = mistdif ? d[10] : d[3];
                agew = d[12];
            } else {                                          // getMIST -> None (:107-119)
                teff = logg = feh = afe = logr = agew = ms_inf();
            }
        } else {
            teff = theta_get(tv, p, MS_P_TEFF); logg = theta_get(tv, p, MS_P_LOGG);
            feh = theta_get(tv, p, MS_P_FEH);   afe = theta_get(tv, p, MS_P_AFE);
            logr = theta_get(tv, p, MS_P_LOGR);
        }
        o[PV_TEFF] = teff; o[PV_LOGG] = logg; o[PV_FEH] = feh; o[PV_AFE] = afe;
        o[PV_VRAD] = theta_get(tv, p, MS_P_VRAD);
        o[PV_VROT] = theta_get(tv, p, MS_P_VROT);
        o[PV_VMIC] = theta_get(tv, p, MS_P_VMIC);
        o[PV_INSTR] = theta_get(tv, p, MS_P_INST_R);
        o[PV_LOGR] = logr;
        o[PV_DIST] = theta_get(tv, p, MS_P_DIST);
        o[PV_AV] = theta_get(tv, p, MS_P_AV);
        o[PV_AGEWGT] = agew;
#pragma unroll
        for (int i = 0; i < MS_MAX_PC; ++i) o[PV_PC0 + i] = theta_get(tv, p, MS_P_PC0 + i);
        if (chi2) chi2[p] = 0.0;
    }
}

// kind 0: photpars (Teff, logg, feh, afe, logR, Dist, Av); kind 1: specpars
// (Teff, logg, feh, afe, Vrad, Vrot, Vmic, Inst_R, pc...)
__global__ void repack_pars_kernel(const double *__restrict__ in, int ncol, int kind, int64_t B,
                                   double *__restrict__ pars) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < B;
         p += (int64_t)gridDim.x * blockDim.x) {
        const double *r = in + p * ncol;
        double *o = pars + p * MS_NPARS;
        for (int i = 0; i < MS_NPARS; ++i) o[i] = ms_nan();
        o[PV_TEFF] = r[0]; o[PV_LOGG] = r[1]; o[PV_FEH] = r[2]; o[PV_AFE] = r[3];
        o[PV_AGEWGT] = 1.0;
        if (kind == 0) {
            o[PV_LOGR] = r[4]; o[PV_DIST] = r[5]; o[PV_AV] = r[6];
        } else {
            o[PV_VRAD] = r[4]; o[PV_VROT] = r[5]; o[PV_VMIC] = r[6]; o[PV_INSTR] = r[7];
            for (int i = 8; i < ncol && i - 8 < MS_MAX_PC; ++i) o[PV_PC0 + i - 8] = r[i];
        }
    }
}

__global__ void finalize_kernel(const double *__restrict__ pars, const double *__restrict__ chi2,
                                int ageweight, int64_t B, double *__restrict__ lnL) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < B;
         p += (int64_t)gridDim.x * blockDim.x) {
        const double teff = pars[p * MS_NPARS + PV_TEFF];
        double l;
        if (teff == ms_inf()) {
            l = -ms_inf();
        } else {
            l = -0.5 * chi2[p];
            if (ageweight) l += log(pars[p * MS_NPARS + PV_AGEWGT]);   // NaN for Agewgt < 0 (:168-176)
        }
        lnL[p] = l;
    }
}

inline unsigned grid_for(const ms_ctx *ctx, int64_t B, int threads) {
    int64_t blocks = (B + threads - 1) / threads;
    int64_t cap = (int64_t)ctx->sm_count * 8;
    return (unsigned)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace

int launch_build_pars(ms_ctx *ctx, const ThetaView &tv, const double *derived, int mistdif,
                      int64_t B, double *pars, double *chi2, cudaStream_t st) {
    if (B <= 0) return MS_OK;
    KernelTimer timer_(ctx, MS_K_PARS, st);
    build_pars_kernel<<<grid_for(ctx, B, 256), 256, 0, st>>>(tv, derived, mistdif, B, pars, chi2);
    ctx->launches++;
    MS_CUDA(cudaGetLastError());
    return MS_OK;
}

int launch_repack_pars(ms_ctx *ctx, const double *in, int ncol, int kind, int64_t B, double *pars,
                       cudaStream_t st) {
    if (B <= 0) return MS_OK;
    KernelTimer timer_(ctx, MS_K_PARS, st);
    repack_pars_kernel<<<grid_for(ctx, B, 256), 256, 0, st>>>(in, ncol, kind, B, pars);
    ctx->launches++;
    MS_CUDA(cudaGetLastError());
    return MS_OK;
}

int launch_finalize(ms_ctx *ctx, const double *pars, const double *chi2, int ageweight, int64_t B,
                    double *lnL, cudaStream_t st) {
    if (B <= 0) return MS_OK;
    KernelTimer timer_(ctx, MS_K_FINALIZE, st);
    finalize_kernel<<<grid_for(ctx, B, 256), 256, 0, st>>>(pars, chi2, ageweight, B, lnL);
    ctx->launches++;
    MS_CUDA(cudaGetLastError());
    return MS_OK;
}
