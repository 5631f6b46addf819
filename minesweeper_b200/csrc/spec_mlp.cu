// Kernel (2b), CUDA-core form (MS_PREC_F64 / MS_PREC_F32): spectral MLP
// d_in -> H -> H -> npix with leaky-ReLU, producing the native-grid flux of a
// chunk of points.  Batched form of PayneSpecPredict.predictspec (external Payne,
// SURVEY.md Appendix B, PARITY UNPINNED; call site reference genmod.py:80-84).
//
// Three launches of one tiled GEMM (out = act(A W^T + b)), A[M][K] and W[N][K]
// both K-contiguous.  The last layer (H x npix) is the only large contraction:
// 2*H*npix flop per point.  The tensor-core form is spec_mlp_tc.cu.
#include "common.cuh"
#include <type_traits>

namespace {

template <typename T>
__global__ void spec_labels_kernel(const double *__restrict__ pars, int64_t B, int d_in, int logt,
                                   const double *xmin_scale /*[16]*/, T *__restrict__ xs) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < B;
         p += (int64_t)gridDim.x * blockDim.x) {
        const double *r = pars + p * MS_NPARS;
        double lab[5] = {logt ? log10(r[PV_TEFF]) : r[PV_TEFF], r[PV_LOGG], r[PV_FEH], r[PV_AFE],
                         r[PV_VMIC]};
        const bool ok = r[PV_TEFF] != ms_inf();
        for (int d = 0; d < d_in; ++d)
            xs[p * d_in + d] = ok ? (T)((lab[d] - xmin_scale[d]) * xmin_scale[8 + d] - 0.5) : (T)0;
    }
}

// out[M][N] = act(A[M][K] . W[N][K]^T + bias[N]); leak < 0 means no activation.
template <typename T, int BM, int BN, int BK>
__global__ void __launch_bounds__(256)
gemm_bias_act_kernel(const T *__restrict__ A, const float *__restrict__ W,
                     const float *__restrict__ bias, T *__restrict__ out, int M, int N, int K,
                     T leak) {
    constexpr int TM = BM / 16, TN = BN / 16;
    __shared__ T As[BK][BM + 4];
    __shared__ T Ws[BK][BN + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    T acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = (T)0;
    for (int k0 = 0; k0 < K; k0 += BK) {
        for (int i = tid; i < BM * BK; i += 256) {
            int r = i / BK, c = i - r * BK;
            int m = m0 + r, k = k0 + c;
            As[c][r] = (m < M && k < K) ? A[(size_t)m * K + k] : (T)0;
        }
        for (int i = tid; i < BN * BK; i += 256) {
            int r = i / BK, c = i - r * BK;
            int n = n0 + r, k = k0 + c;
            Ws[c][r] = (n < N && k < K) ? (T)W[(size_t)n * K + k] : (T)0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            T a[TM], w[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[kk][ty + 16 * i];
#pragma unroll
            for (int j = 0; j < TN; ++j) w[j] = Ws[kk][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fma(a[i], w[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int m = m0 + ty + 16 * i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int n = n0 + tx + 16 * j;
            if (n >= N) continue;
            T v = acc[i][j] + (T)bias[n];
            if (leak >= (T)0) v = v > (T)0 ? v : leak * v;
            out[(size_t)m * N + n] = v;
        }
    }
}

template <typename T>
int gemm(ms_ctx *ctx, const T *A, const float *W, const float *bias, T *out, int M, int N, int K,
         double leak, cudaStream_t st) {
    dim3 grid((N + 63) / 64, (M + 63) / 64);
    gemm_bias_act_kernel<T, 64, 64, 16><<<grid, 256, 0, st>>>(A, W, bias, out, M, N, K, (T)leak);
    ctx->launches++;
    MS_CUDA(cudaGetLastError());
    return MS_OK;
}

}  // namespace

// Computes native flux for `B` points of `pars` into flux[B][npix] (type by precision).
// hid: scratch of >= B * (d_in + 2 h) elements.
template <typename T>
int spec_mlp_chunk_t(ms_ctx *ctx, const double *pars, int64_t B, T *hid, T *flux,
                     const double *d_xmin_scale, cudaStream_t st) {
    const SpecNetDev &n = ctx->spec;
    if constexpr (std::is_same<T, float>::value) {
        if (ctx->precision >= MS_PREC_TC)          // all three layers on the tensor-core path (spec_mlp_tc.cu)
            return launch_spec_mlp_tc(ctx, pars, B, flux, st);
    }
    KernelTimer timer_(ctx, MS_K_SPEC_MLP, st);
    T *xs = hid, *h1 = hid + (size_t)B * n.d_in, *h2 = h1 + (size_t)B * n.h;
    int blocks = (int)((B + 255) / 256);
    spec_labels_kernel<T><<<blocks, 256, 0, st>>>(pars, B, n.d_in, n.logt_input, d_xmin_scale, xs);
    ctx->launches++;
    MS_CUDA(cudaGetLastError());
    int rc;
    if ((rc = gemm<T>(ctx, xs, n.w0, n.b0, h1, (int)B, n.h, n.d_in, n.leak, st))) return rc;
    if ((rc = gemm<T>(ctx, h1, n.w1, n.b1, h2, (int)B, n.h, n.h, n.leak, st))) return rc;
    if ((rc = gemm<T>(ctx, h2, n.w2, n.b2, flux, (int)B, n.npix, n.h, -1.0, st))) return rc;
    return MS_OK;
}

template int spec_mlp_chunk_t<float>(ms_ctx *, const double *, int64_t, float *, float *, const double *, cudaStream_t);
template int spec_mlp_chunk_t<double>(ms_ctx *, const double *, int64_t, double *, double *, const double *, cudaStream_t);
