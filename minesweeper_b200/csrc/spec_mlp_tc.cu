// Kernel (2b), tensor-core form (MS_PREC_TC): the output layer of the spectral MLP,
// flux[B][P] = h2[B][H] . W2[P][H]^T + b2 -- 97 % of the spectral network's flops -- as a
// persistent tcgen05 GEMM.  Same contract as the last gemm of spec_mlp.cu
// (PayneSpecPredict.predictspec, external Payne, PARITY UNPINNED; call site genmod.py:80-84).
//
// Numerics as in phot_mlp_tc.cu: both operands are split into fp16 hi + lo and every product
// is three kind::f16 MMAs (hi.hi + lo.hi + hi.lo) accumulated in fp32.  W2 is pre-scaled by a
// power of two so that small weights stay in the fp16 normal range; the epilogue undoes it.
//
// Tiling: 128 points x 128 pixels per tile, K streamed in chunks of 64 through a 3-stage
// shared-memory ring filled by bulk async copies (A and B images are stored tile-major in
// global memory in the UMMA K-major core-matrix layout, so a stage is two plain 32 KB
// copies); two 128-column TMEM accumulators so the epilogue of tile i overlaps the MMAs of
// tile i+1.  Warp 0 = producer, warp 1 = MMA issuer, warps 2-5 = epilogue (TMEM -> registers
// -> + bias -> global).  Tiles are ordered m-fastest so concurrently running CTAs share the
// same W2 tile through L2.  L2 -> shared operand traffic (64 KB per 64-wide K chunk) is what
// bounds this kernel, not the tensor pipe.
#include "common.cuh"
#include "tc_ptx.cuh"

#include <cuda_fp16.h>
#include <cmath>

namespace {

constexpr int BM = 128, BN = 128, BK = 64, STAGES = 3;
constexpr int OPER_BYTES = (BK / 8) * 128 * 16;          // one operand half (hi or lo) of a stage: 16 KB
constexpr int STAGE_BYTES = 4 * OPER_BYTES;              // A hi, A lo, B hi, B lo
constexpr int SM_BAR = STAGES * STAGE_BYTES;
constexpr int SM_TOTAL = SM_BAR + 128;
constexpr int NTHREADS = 192;
constexpr uint32_t IDESC = make_idesc_f16(BM, BN);

enum { BAR_FULL = 0, BAR_EMPTY = STAGES, BAR_ACC_FULL = 2 * STAGES, BAR_ACC_EMPTY = 2 * STAGES + 2 };

// h2[B][H] f32 -> A image [m_tile][k_chunk][hi|lo][k8][row][8] fp16 (zero padded)
__global__ void pack_a_kernel(const float *__restrict__ h2, int64_t B, int H, int nkc, __half *__restrict__ img) {
    const int64_t ntask = (int64_t)((B + BM - 1) / BM) * BM * nkc * (BK / 8);
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < ntask; t += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(t % BM);
        const int k8 = (int)((t / BM) % (BK / 8));
        const int kc = (int)((t / (BM * (BK / 8))) % nkc);
        const int64_t mt = t / ((int64_t)BM * (BK / 8) * nkc);
        const int64_t row = mt * BM + r;
        const int k0 = kc * BK + k8 * 8;
        __half hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float v = (row < B && k0 + i < H) ? h2[row * H + k0 + i] : 0.f;
            hi[i] = __float2half_rn(v);
            lo[i] = __float2half_rn(v - __half2float(hi[i]));
        }
        __half *base = img + ((mt * nkc + kc) * 2) * (int64_t)(OPER_BYTES / 2) + (int64_t)k8 * 1024 + r * 8;
        *reinterpret_cast<uint4 *>(base) = *reinterpret_cast<const uint4 *>(hi);
        *reinterpret_cast<uint4 *>(base + OPER_BYTES / 2) = *reinterpret_cast<const uint4 *>(lo);
    }
}

struct GemmArgs {
    const unsigned char *a_img;      // [m_tiles][nkc][2][OPER_BYTES]
    const unsigned char *b_img;      // [n_tiles][nkc][2][OPER_BYTES]
    const float *bias;
    float *out;                      // [B][P]
    int64_t B;
    int P, nkc, m_tiles, n_tiles;
    float inv_scale;
};

__global__ void __launch_bounds__(NTHREADS, 1)
spec_out_gemm_tc_kernel(GemmArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t s_tmem_base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sbase = smem_u32(smem);
    auto bar = [&](int i) { return sbase + SM_BAR + 8u * (uint32_t)i; };
    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(bar(BAR_FULL + i), 1); mbar_init(bar(BAR_EMPTY + i), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(bar(BAR_ACC_FULL + i), 1); mbar_init(bar(BAR_ACC_EMPTY + i), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&s_tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem_base;
    const int ntiles = a.m_tiles * a.n_tiles;
    const int nkc = a.nkc;

    if (warp == 0) {
        if (lane == 0) {
            int g = 0;                                            // global stage counter
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                const int mt = t % a.m_tiles, nt = t / a.m_tiles;
                for (int kc = 0; kc < nkc; ++kc, ++g) {
                    const int st = g % STAGES;
                    mbar_wait(bar(BAR_EMPTY + st), ((g / STAGES) & 1) ^ 1);
                    mbar_expect_tx(bar(BAR_FULL + st), STAGE_BYTES);
                    const uint32_t dst = sbase + st * STAGE_BYTES;
                    bulk_g2s(dst, a.a_img + ((size_t)mt * nkc + kc) * 2 * OPER_BYTES, 2 * OPER_BYTES, bar(BAR_FULL + st));
                    bulk_g2s(dst + 2 * OPER_BYTES, a.b_img + ((size_t)nt * nkc + kc) * 2 * OPER_BYTES, 2 * OPER_BYTES,
                             bar(BAR_FULL + st));
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int g = 0, it = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
                const int buf = it & 1;
                mbar_wait(bar(BAR_ACC_EMPTY + buf), ((it >> 1) & 1) ^ 1);     // epilogue drained this accumulator
                tc_fence_after();
                const uint32_t d = tmem + buf * BN;
                for (int kc = 0; kc < nkc; ++kc, ++g) {
                    const int st = g % STAGES;
                    mbar_wait(bar(BAR_FULL + st), (g / STAGES) & 1);
                    tc_fence_after();
                    const uint32_t sa = sbase + st * STAGE_BYTES;
#pragma unroll
                    for (int ks = 0; ks < BK / 16; ++ks) {
                        const uint64_t ah = make_desc(sa + ks * 4096, 2048, 128);
                        const uint64_t al = make_desc(sa + OPER_BYTES + ks * 4096, 2048, 128);
                        const uint64_t bh = make_desc(sa + 2 * OPER_BYTES + ks * 4096, 2048, 128);
                        const uint64_t bl = make_desc(sa + 3 * OPER_BYTES + ks * 4096, 2048, 128);
                        mma_ss(d, ah, bh, (kc | ks) != 0, IDESC);
                        mma_ss(d, al, bh, 1, IDESC);
                        mma_ss(d, ah, bl, 1, IDESC);
                    }
                    tc_commit(bar(BAR_EMPTY + st));                            // stage reusable once these MMAs retire
                }
                tc_commit(bar(BAR_ACC_FULL + buf));
            }
        }
    } else {
        // epilogue: warp w owns TMEM lanes 32 (w % 4) .. +31, a thread owns one point (row)
        const int q = warp & 3;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        int it = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int mt = t % a.m_tiles, nt = t / a.m_tiles;
            const int buf = it & 1;
            const int64_t row = (int64_t)mt * BM + q * 32 + lane;
            mbar_wait(bar(BAR_ACC_FULL + buf), (it >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int j = 0; j < BN / 32; ++j) {
                uint32_t v[32];
                TMEM_LD32(tmem + lane_base + buf * BN + 32 * j, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const int col0 = nt * BN + 32 * j;
                if (row < a.B) {
                    float *o = a.out + row * a.P + col0;
                    if (col0 + 32 <= a.P) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float4 bb = __ldg(reinterpret_cast<const float4 *>(a.bias + col0) + i);
                            float4 r4;
                            r4.x = fmaf(__uint_as_float(v[4 * i]), a.inv_scale, bb.x);
                            r4.y = fmaf(__uint_as_float(v[4 * i + 1]), a.inv_scale, bb.y);
                            r4.z = fmaf(__uint_as_float(v[4 * i + 2]), a.inv_scale, bb.z);
                            r4.w = fmaf(__uint_as_float(v[4 * i + 3]), a.inv_scale, bb.w);
                            reinterpret_cast<float4 *>(o)[i] = r4;
                        }
                    } else {
                        for (int i = 0; i < 32 && col0 + i < a.P; ++i)
                            o[i] = fmaf(__uint_as_float(v[i]), a.inv_scale, a.bias[col0 + i]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(BAR_ACC_EMPTY + buf));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

}  // namespace

struct SpecTcPack {
    unsigned char *b_img = nullptr;
    int nkc = 0, n_tiles = 0;
    float inv_scale = 1.f;
    __half *a_img = nullptr; size_t a_img_bytes = 0;
};

int spec_tc_prepare(ms_ctx *ctx) {
    SpecNetDev &S = ctx->spec;
    const int H = S.h, P = S.npix;
    if (P % 4 != 0) { ms_set_error("MS_PREC_TC needs a pixel count that is a multiple of 4"); return MS_EINVAL; }
    std::vector<float> w2((size_t)P * H);
    MS_CUDA(cudaMemcpy(w2.data(), S.w2, w2.size() * 4, cudaMemcpyDeviceToHost));
    float wmax = 0.f;
    for (float v : w2) wmax = std::fabs(v) > wmax ? std::fabs(v) : wmax;
    int e = 0;
    if (wmax > 0.f) { std::frexp(wmax, &e); }
    const int shift = 14 - e;                               // max |w| * 2^shift in [2^13, 2^14)
    const double scale = std::ldexp(1.0, shift);
    SpecTcPack *pk = new SpecTcPack();
    pk->nkc = (H + BK - 1) / BK;
    pk->n_tiles = (P + BN - 1) / BN;
    pk->inv_scale = (float)std::ldexp(1.0, -shift);
    const size_t bytes = (size_t)pk->n_tiles * pk->nkc * 2 * OPER_BYTES;
    std::vector<__half> img(bytes / 2, __float2half_rn(0.f));
    for (int n = 0; n < P; ++n)
        for (int k = 0; k < H; ++k) {
            const double v = scale * (double)w2[(size_t)n * H + k];
            const __half hi = __float2half_rn((float)v);
            const __half lo = __float2half_rn((float)(v - (double)__half2float(hi)));
            const size_t at = (((size_t)(n / BN) * pk->nkc + k / BK) * 2) * (OPER_BYTES / 2) +
                              (size_t)((k % BK) / 8) * 1024 + (size_t)(n % BN) * 8 + (k % 8);
            img[at] = hi;
            img[at + OPER_BYTES / 2] = lo;
        }
    S.tc_pack = pk;                                        // owned by the ctx from here: ms_destroy releases it on any error
    MS_CUDA(cudaMalloc((void **)&pk->b_img, bytes));
    MS_CUDA(cudaMemcpy(pk->b_img, img.data(), bytes, cudaMemcpyHostToDevice));
    MS_CUDA(cudaFuncSetAttribute(spec_out_gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
    return MS_OK;
}

void spec_tc_release(ms_ctx *ctx) {
    SpecTcPack *pk = static_cast<SpecTcPack *>(ctx->spec.tc_pack);
    if (!pk) return;
    cudaFree(pk->b_img); cudaFree(pk->a_img);
    delete pk;
    ctx->spec.tc_pack = nullptr;
}

// flux[B][P] (f32) from h2[B][H] (f32) on the tensor cores
int launch_spec_out_tc(ms_ctx *ctx, const float *h2, int64_t B, float *flux, cudaStream_t st) {
    SpecNetDev &S = ctx->spec;
    SpecTcPack *pk = static_cast<SpecTcPack *>(S.tc_pack);
    if (!pk) { ms_set_error("spectral net not packed for MS_PREC_TC"); return MS_ESTATE; }
    const int m_tiles = (int)((B + BM - 1) / BM);
    const size_t need = (size_t)m_tiles * pk->nkc * 2 * OPER_BYTES;
    if (pk->a_img_bytes < need) {
        if (pk->a_img) cudaFree(pk->a_img);
        pk->a_img = nullptr; pk->a_img_bytes = 0;
        MS_CUDA(cudaMalloc((void **)&pk->a_img, need));
        pk->a_img_bytes = need;
    }
    KernelTimer timer_(ctx, MS_K_SPEC_MLP, st);
    {
        const int64_t ntask = (int64_t)m_tiles * BM * pk->nkc * (BK / 8);
        int blocks = (int)((ntask + 255) / 256);
        if (blocks > ctx->sm_count * 16) blocks = ctx->sm_count * 16;
        pack_a_kernel<<<blocks, 256, 0, st>>>(h2, B, S.h, pk->nkc, pk->a_img);
        ctx->launches++;
        MS_CUDA(cudaGetLastError());
    }
    GemmArgs ga;
    ga.a_img = reinterpret_cast<const unsigned char *>(pk->a_img);
    ga.b_img = pk->b_img; ga.bias = S.b2; ga.out = flux; ga.B = B; ga.P = S.npix;
    ga.nkc = pk->nkc; ga.m_tiles = m_tiles; ga.n_tiles = pk->n_tiles; ga.inv_scale = pk->inv_scale;
    const int ntiles = m_tiles * pk->n_tiles;
    const int grid = ntiles < ctx->sm_count ? ntiles : ctx->sm_count;
    spec_out_gemm_tc_kernel<<<grid, NTHREADS, SM_TOTAL, st>>>(ga);
    ctx->launches++;
    MS_CUDA(cudaGetLastError());
    return MS_OK;
}
