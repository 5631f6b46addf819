// Kernel (2), CUDA-core form (MS_PREC_F64 / MS_PREC_F32): per-band photometric
// MLP d_in -> H -> H -> 1 with sigmoid activations, fused with the magnitude
// formula and the photometric chi-square.  Batched form of
//   GenMod.genphot            reference minesweeper/genmod.py:97-142
//   FastPayneSEDPredict.sed   external Payne (PARITY UNPINNED)
//   likelihood.lnlike         reference minesweeper/likelihood.py:204-210
//
// One thread owns one point; a CTA sweeps the bands.  Layer weights are staged
// in shared memory and read as warp-wide broadcasts; the hidden vector of each
// point lives in a shared-memory column private to its thread.  The magnitude /
// residual / chi-square epilogue is f64 in every mode.
// The tensor-core form of the same computation is phot_mlp_tc.cu.
#include "common.cuh"

namespace {

template <typename T> __device__ __forceinline__ T sigmoid_t(T z);
template <> __device__ __forceinline__ float sigmoid_t<float>(float z) { return 1.0f / (1.0f + expf(-z)); }
template <> __device__ __forceinline__ double sigmoid_t<double>(double z) { return 1.0 / (1.0 + exp(-z)); }

struct PhotKArgs {
    const double *pars; const int32_t *star_slot;
    double *mags; double *chi2; double *lnL; int ageweight;
    int slot0, nslots;
    const double *obs_mag, *obs_err2;
    const float *w1, *b1, *w2, *b2, *w3, *b3;
    double xmin[8], xscale[8];
    int nb, d_in, h, jc;
};

// The seven photometric parameters of point p (likelihood.py:154-160):
// (Teff, logg, feh, afe, logR, Dist, Av); false for a failed interpolation.
__device__ __forceinline__ bool phot_inputs(const PhotKArgs &a, int64_t p, double *pp) {
    const double *r = a.pars + p * MS_NPARS;
    pp[0] = r[PV_TEFF]; pp[1] = r[PV_LOGG]; pp[2] = r[PV_FEH]; pp[3] = r[PV_AFE];
    pp[4] = r[PV_LOGR]; pp[5] = r[PV_DIST]; pp[6] = r[PV_AV];
    return pp[0] != ms_inf();
}

template <typename T, int TP>
__global__ void __launch_bounds__(TP)
phot_mlp_kernel(PhotKArgs a, int64_t B) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int H = a.h, D = a.d_in, JC = a.jc;
    T *w1s = reinterpret_cast<T *>(smem_raw);            // [H][D]
    T *b1s = w1s + H * D;
    T *b2s = b1s + H;
    T *w3s = b2s + H;
    T *w2s = w3s + H;                                    // [H (k)][JC]
    T *a1s = w2s + (size_t)H * JC;                       // [H][TP]
    const int tid = threadIdx.x;

    for (int64_t tile = blockIdx.x; tile * TP < B; tile += gridDim.x) {
        const int64_t p = tile * TP + tid;
        const bool live = p < B;
        double pp[7];
        bool ok = live && phot_inputs(a, p, pp);
        T xs[8];
        double mag0 = 0.0;
        if (ok) {
            // genmod.py:111-113: logL from log R and log10(Teff)
            const double logt = log10(pp[0]);
            const double logl = 2.0 * pp[4] + 4.0 * (logt - log10(5770.0));
            mag0 = -2.5 * logl + 4.74 + 5.0 * log10(pp[5]) - 5.0;
            // the predictor receives logt and raises 10 to it again (genmod.py:116, Appendix B)
            const double xin[6] = {pow(10.0, logt), pp[1], pp[2], pp[3], pp[6], 3.1};
#pragma unroll
            for (int d = 0; d < 6; ++d)
                if (d < D) xs[d] = (T)((xin[d] - a.xmin[d]) * a.xscale[d] - 0.5);
        }
        // star slot of this row; a slot outside the loaded table makes the row NaN instead of reading out of bounds
        int slot = (live && a.star_slot) ? a.star_slot[p] : a.slot0;
        const bool slot_ok = (unsigned)slot < (unsigned)a.nslots;
        if (!slot_ok) slot = 0;
        double chi2 = 0.0;
        for (int b = 0; b < a.nb; ++b) {
            __syncthreads();
            for (int i = tid; i < H * D; i += TP) w1s[i] = (T)a.w1[(size_t)b * H * D + i];
            for (int i = tid; i < H; i += TP) {
                b1s[i] = (T)a.b1[b * H + i];
                b2s[i] = (T)a.b2[b * H + i];
                w3s[i] = (T)a.w3[b * H + i];
            }
            __syncthreads();
            if (ok) {
                for (int k = 0; k < H; ++k) {
                    T z = b1s[k];
                    for (int d = 0; d < D; ++d) z = fma(w1s[k * D + d], xs[d], z);
                    a1s[k * TP + tid] = sigmoid_t<T>(z);
                }
            }
            T bc = (T)a.b3[b];
            for (int j0 = 0; j0 < H; j0 += JC) {
                __syncthreads();
                const float *w2g = a.w2 + (size_t)b * H * H + (size_t)j0 * H;
                for (int i = tid; i < H * JC; i += TP) {
                    int jj = i / H, k = i - jj * H;       // coalesced along k
                    w2s[k * JC + jj] = (T)w2g[(size_t)jj * H + k];
                }
                __syncthreads();
                if (ok) {
                    for (int jj0 = 0; jj0 < JC; jj0 += 8) {
                        T acc[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) acc[u] = b2s[j0 + jj0 + u];
                        for (int k = 0; k < H; ++k) {
                            const T av = a1s[k * TP + tid];
                            const T *wr = w2s + k * JC + jj0;
#pragma unroll
                            for (int u = 0; u < 8; ++u) acc[u] = fma(av, wr[u], acc[u]);
                        }
#pragma unroll
                        for (int u = 0; u < 8; ++u) bc = fma(w3s[j0 + jj0 + u], sigmoid_t<T>(acc[u]), bc);
                    }
                }
            }
            if (ok) {
                const double mag = mag0 - (double)bc;
                if (a.mags) a.mags[p * a.nb + b] = mag;
                if (a.chi2 || a.lnL) {
                    const double o = a.obs_mag[(size_t)slot * a.nb + b];
                    const double r = mag - o;
                    const double t = (r * r) / a.obs_err2[(size_t)slot * a.nb + b];
                    if (t == t) chi2 += t;                // nansum (likelihood.py:207)
                }
            } else if (live && a.mags) {
                a.mags[p * a.nb + b] = ms_nan();
            }
        }
        if (live && a.lnL) {
            double l = -ms_inf();
            if (ok) {
                l = -0.5 * ((a.chi2 ? a.chi2[p] : 0.0) + chi2);
                if (a.ageweight) l += log(a.pars[p * MS_NPARS + PV_AGEWGT]);
                if (!slot_ok) l = ms_nan();
            }
            a.lnL[p] = l;
        } else if (ok && a.chi2) {
            a.chi2[p] += slot_ok ? chi2 : ms_nan();
        }
    }
}

template <typename T, int TP>
int run(ms_ctx *ctx, PhotKArgs &ka, int64_t B, cudaStream_t st) {
    const int H = ka.h, D = ka.d_in;
    int jc = H;
    while ((size_t)H * jc * sizeof(T) > 64 * 1024 && (jc % 16) == 0) jc /= 2;
    ka.jc = jc;
    size_t smem = ((size_t)H * D + 3 * H + (size_t)H * jc + (size_t)H * TP) * sizeof(T);
    if (smem > 220 * 1024 || (H % 8) != 0 || (H % jc) != 0 || (jc % 8) != 0) {
        ms_set_error("phot net hidden width %d unsupported by the CUDA-core kernel", H);
        return MS_EINVAL;
    }
    auto kern = phot_mlp_kernel<T, TP>;
    MS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t tiles = (B + TP - 1) / TP;
    int per_sm = (int)((220 * 1024) / smem); if (per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)ctx->sm_count * per_sm;
    if (grid > tiles) grid = tiles;
    KernelTimer timer_(ctx, MS_K_PHOT, st);
    kern<<<(unsigned)grid, TP, smem, st>>>(ka, B);
    ctx->launches++;
    MS_CUDA(cudaGetLastError());
    return MS_OK;
}

}  // namespace

int launch_phot(ms_ctx *ctx, const PhotArgs &a, int64_t B, cudaStream_t st) {
    const PhotNetDev &n = ctx->phot;
    if (!n.w1) { ms_set_error("photometric net not configured"); return MS_ESTATE; }
    if (B <= 0) return MS_OK;
    if ((a.chi2 || a.lnL) && !ctx->obs_mag) { ms_set_error("no star observations set"); return MS_ESTATE; }
    PhotKArgs ka;
    ka.pars = a.pars; ka.star_slot = a.star_slot; ka.mags = a.mags; ka.chi2 = a.chi2; ka.lnL = a.lnL; ka.ageweight = a.ageweight;
    ka.slot0 = a.slot0; ka.nslots = ctx->nslots;
    ka.obs_mag = ctx->obs_mag; ka.obs_err2 = ctx->obs_err2;
    ka.w1 = n.w1; ka.b1 = n.b1; ka.w2 = n.w2; ka.b2 = n.b2; ka.w3 = n.w3; ka.b3 = n.b3;
    for (int i = 0; i < 8; ++i) { ka.xmin[i] = n.xmin[i]; ka.xscale[i] = n.xscale[i]; }
    ka.nb = n.nb; ka.d_in = n.d_in; ka.h = n.h; ka.jc = n.h;
    if (ctx->precision == MS_PREC_F64) return run<double, 64>(ctx, ka, B, st);
    return run<float, 128>(ctx, ka, B, st);
}
