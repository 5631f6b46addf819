// Kernel (1): bracket search + 16-corner gather + N-linear interpolation on the
// dense MIST grid.  Batched form of GenMIST.getMIST (reference
// minesweeper/fastMISTmod.py:100-108) with the semantics of params_to_grid
// (:127-150), linear_weights (:194-213) and weights (:152-170); the KD-tree
// ball query (:172-192) is replaced by direct addressing of the <= 16 corners
// that can carry non-zero weight.
//
// Layout: table[na][nf][nm][ne][CP] (EEP fastest, so the two EEP corners of a
// cell are adjacent in memory); a vertex that is absent from the track tables
// has NaN in column 0.  Axes live in shared memory; theta stays f64 so the
// brackets equal np.digitize(..., right=False) - 1 bit for bit.
//
// HBM-bound: algorithmic bytes per point = 4*8 (theta) + 16*9*sizeof(T) (corners)
// + 13*8 (derived) -> 712 B (f32 table) / 1288 B (f64 table).
#include "common.cuh"
#include "mist_interp.cuh"

namespace {

template <typename T>
__global__ void __launch_bounds__(256)
mist_interp_kernel(ThetaView tv, int64_t B, const double *__restrict__ axes, int ne, int nm, int nf,
                   int na, const T *__restrict__ table, int32_t *__restrict__ brackets,
                   double *__restrict__ derived, uint8_t *__restrict__ okflag,
                   double *__restrict__ pars, double *__restrict__ chi2, int mistdif) {
    extern __shared__ double s_axes[];
    const int ntot = ne + nm + nf + na;
    for (int i = threadIdx.x; i < ntot; i += blockDim.x) s_axes[i] = axes[i];
    __syncthreads();
    const double *g[4] = {s_axes, s_axes + ne, s_axes + ne + nm, s_axes + ne + nm + nf};
    const int n[4] = {ne, nm, nf, na};

    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < B;
         p += (int64_t)gridDim.x * blockDim.x) {
        double x[4];
        x[0] = theta_get(tv, p, MS_P_EEP);
        x[1] = theta_get(tv, p, MS_P_INIT_MASS);
        x[2] = theta_get(tv, p, MS_P_INIT_FEH);
        x[3] = theta_get(tv, p, MS_P_INIT_AFE);
        int idx[4];
        double acc[MS_NCOL], wsum;
        const bool inside = mist::interp_point<T, 1>(g, n, table, x, idx, acc, wsum);
        if (brackets) {
            int4 b4 = make_int4(idx[0], idx[1], idx[2], idx[3]);
            *reinterpret_cast<int4 *>(brackets + 4 * p) = b4;
        }
        const bool ok = inside && (wsum > 0.0);
        if (okflag) okflag[p] = ok ? 1 : 0;
        if (pars) {
            // the per-point parameter block of likelihood.lnlikefn (likelihood.py:121-160),
            // assembled here so the prediction never makes a round trip through HBM
            const double inv = ok ? 1.0 / wsum : 0.0;
            const double inf = ms_inf();
            double o[MS_NPARS];
            o[PV_TEFF] = ok ? pow(10.0, acc[4] * inv) : inf;             // likelihood.py:121
            o[PV_LOGG] = ok ? acc[7] * inv : inf;
            o[PV_FEH] = ok ? (mistdif ? acc[5] * inv : x[2]) : inf;      // :127-132
            o[PV_AFE] = ok ? (mistdif ? acc[6] * inv : x[3]) : inf;
            o[PV_VRAD] = theta_get(tv, p, MS_P_VRAD);
            o[PV_VROT] = theta_get(tv, p, MS_P_VROT);
            o[PV_VMIC] = theta_get(tv, p, MS_P_VMIC);
            o[PV_INSTR] = theta_get(tv, p, MS_P_INST_R);
            o[PV_LOGR] = ok ? acc[2] * inv : inf;
            o[PV_DIST] = theta_get(tv, p, MS_P_DIST);
            o[PV_AV] = theta_get(tv, p, MS_P_AV);
            o[PV_AGEWGT] = ok ? acc[8] * inv : inf;
#pragma unroll
            for (int i = 0; i < MS_MAX_PC; ++i) o[PV_PC0 + i] = theta_get(tv, p, MS_P_PC0 + i);
            // 160-byte row, 16-byte vector stores (each 32-byte sector is touched twice, not four times)
            double2 *dst = reinterpret_cast<double2 *>(pars + p * MS_NPARS);
#pragma unroll
            for (int i = 0; i < MS_NPARS / 2; ++i) dst[i] = make_double2(o[2 * i], o[2 * i + 1]);
            if (chi2) chi2[p] = 0.0;
        }
        if (derived) {
            double *o = derived + p * MS_NDERIVED;
            o[0] = x[0]; o[1] = x[1]; o[2] = x[2]; o[3] = x[3];
            const double inv = ok ? 1.0 / wsum : 0.0;
#pragma unroll
            for (int c = 0; c < MS_NCOL; ++c) o[4 + c] = ok ? acc[c] * inv : ms_inf();
        }
    }
}

}  // namespace

int launch_mist_interp(ms_ctx *ctx, const ThetaView &tv, int64_t B, int32_t *brackets,
                       double *derived, uint8_t *ok, cudaStream_t st, double *pars, double *chi2,
                       int mistdif) {
    const GridDev &g = ctx->grid;
    if (!g.axes) { ms_set_error("MIST grid not configured"); return MS_ESTATE; }
    if (B <= 0) return MS_OK;
    KernelTimer timer_(ctx, MS_K_INTERP, st);
    const int threads = 256;
    int64_t blocks = (B + threads - 1) / threads;
    const int64_t cap = (int64_t)ctx->sm_count * 8;
    if (blocks > cap) blocks = cap;
    size_t smem = (size_t)(g.ne + g.nm + g.nf + g.na) * sizeof(double);
    if (ctx->precision == MS_PREC_F64)
        mist_interp_kernel<double><<<(unsigned)blocks, threads, smem, st>>>(
            tv, B, g.axes, g.ne, g.nm, g.nf, g.na, g.tab64, brackets, derived, ok, pars, chi2, mistdif);
    else
        mist_interp_kernel<float><<<(unsigned)blocks, threads, smem, st>>>(
            tv, B, g.axes, g.ne, g.nm, g.nf, g.na, g.tab32, brackets, derived, ok, pars, chi2, mistdif);
    ctx->launches++;
    MS_CUDA(cudaGetLastError());
    return MS_OK;
}
