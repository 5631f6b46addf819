// Kernel (2b), tensor-core form (MS_PREC_TC): the spectral MLP d_in -> H -> H -> P of
// PayneSpecPredict.predictspec (external Payne, PARITY UNPINNED; call site genmod.py:80-84) as
//   (i)   layer 0 (K = d_in <= 5) on CUDA cores, written straight into the fp16 hi/lo operand image of
//   (ii)  layer 1 (H x H) as a tcgen05 GEMM whose epilogue (bias, leaky ReLU, fp16 hi/lo split) writes
//         the operand image of
//   (iii) the output layer flux[B][P] = h2[B][H] . W2[P][H]^T + b2 -- 97 % of the network's flops -- as the
//         same persistent tcgen05 GEMM with an f32 epilogue.
// No f32 hidden activations and no separate packing pass exist between the layers.
//
// Numerics as in phot_mlp_tc.cu: both operands are split into fp16 hi + lo and every product
// is three kind::f16 MMAs (hi.hi + lo.hi + hi.lo) accumulated in fp32.  Weights are pre-scaled by a
// power of two so that small weights stay in the fp16 normal range; the epilogue undoes it.
//
// Tiling: 128 points x 128 columns per tile, K streamed in chunks of 64 through a 3-stage
// shared-memory ring filled by bulk async copies (A and B images are stored tile-major in
// global memory in the UMMA K-major core-matrix layout, so a stage is two plain 32 KB
// copies); two 128-column TMEM accumulators so the epilogue of tile i overlaps the MMAs of
// tile i+1.  Warp 0 = producer, warp 1 = MMA issuer, warps 2-5 = epilogue (TMEM -> registers
// -> + bias -> global).  Tiles are ordered m-fastest so concurrently running CTAs share the
// same weight tile through L2.  L2 -> shared operand traffic (64 KB per 64-wide K chunk) is what
// bounds the output-layer GEMM, not the tensor pipe: measured 21 B/clk inbound per SM (5.2 TB/s on the chip) whether
// 48 or 148 CTAs run, against the 85 B/clk a tensor-bound 128 x 128 split-precision tile would need.
#include "common.cuh"
#include "tc_ptx.cuh"

#include <cuda_fp16.h>
#include <cmath>
#include <cstdlib>

namespace {

constexpr int BM = 128, BN = 128, BK = 64, STAGES = 3;
constexpr int OPER_BYTES = (BK / 8) * 128 * 16;          // one operand half (hi or lo) of a stage: 16 KB
constexpr int STAGE_BYTES = 4 * OPER_BYTES;              // A hi, A lo, B hi, B lo
constexpr int SM_BAR = STAGES * STAGE_BYTES;
constexpr int SM_TOTAL = SM_BAR + 128;
constexpr int NTHREADS = 192;
constexpr uint32_t IDESC = make_idesc_f16(BM, BN);

enum { BAR_FULL = 0, BAR_EMPTY = STAGES, BAR_ACC_FULL = 2 * STAGES, BAR_ACC_EMPTY = 2 * STAGES + 2 };

// Layer 0 + operand packing: h1 = lrelu(W0 x + b0) for the scaled labels x of every point, written as the
// A image [m_tile][k_chunk][hi|lo][k8][row][8] (fp16, zero padded in rows >= B and k >= H) of the layer-1 GEMM.
// One thread per (row, group of 8 hidden units).
// chunk_pad > 0: the image holds the rows of a super-batch chunk by chunk, every chunk padded to chunk_pad (a multiple of
// BM) rows, so that each chunk's m-tiles start on an m-tile of the image: image row c * chunk_pad + r <- input row
// c * chunk + r.  nrows_img = rows of the image (a multiple of BM).
__global__ void spec_l0_pack_kernel(const double *__restrict__ pars, int64_t B, int d_in, int logt,
                                    const double *__restrict__ xms /*[16]: xmin | xscale*/,
                                    const float *__restrict__ w0, const float *__restrict__ b0, int H, float leak,
                                    int nkc, __half *__restrict__ img, int64_t nrows_img, int chunk, int chunk_pad) {
    const int64_t ntask = nrows_img * nkc * (BK / 8);
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < ntask; t += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(t % BM);
        const int k8 = (int)((t / BM) % (BK / 8));
        const int kc = (int)((t / (BM * (BK / 8))) % nkc);
        const int64_t mt = t / ((int64_t)BM * (BK / 8) * nkc);
        int64_t row = mt * BM + r;                                   // row of the image -> row of the input
        bool inb = row < B;
        if (chunk_pad > 0) {
            const int64_t c = row / chunk_pad, rr = row - c * chunk_pad;
            row = c * chunk + rr;
            inb = rr < chunk && row < B;
        }
        const int k0 = kc * BK + k8 * 8;
        __half hi[8], lo[8];
        float xs[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        bool live = inb;
        if (live) {
            const double *q = pars + row * MS_NPARS;
            live = q[PV_TEFF] != ms_inf();
            const double lab[5] = {logt ? log10(q[PV_TEFF]) : q[PV_TEFF], q[PV_LOGG], q[PV_FEH], q[PV_AFE], q[PV_VMIC]};
            for (int d = 0; d < d_in; ++d) xs[d] = live ? (float)((lab[d] - xms[d]) * xms[8 + d] - 0.5) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float v = 0.f;
            if (inb && k0 + i < H) {
                v = b0[k0 + i];
                for (int d = 0; d < d_in; ++d) v = fmaf(xs[d], w0[(k0 + i) * d_in + d], v);
                v = v > 0.f ? v : leak * v;
            }
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
    float *out;                      // PACK = false: [B][P] f32
    unsigned char *out_img;          // PACK = true: operand image of the next layer, [m_tiles][nkc_out][2][OPER_BYTES]
    int64_t B;
    int P, nkc, m_tiles, n_tiles;    // P = valid output columns
    int nkc_out;
    float inv_scale, leak;
    unsigned char *out_ts;           // PACK = true: when set, the next layer's A operand as a TMEM image (see spec_out_gemm_ts_kernel)
};

// PACK = false: out = acc * inv_scale + bias (f32 rows).  PACK = true: hidden layer: leaky ReLU, then the fp16
// hi / lo operand image of the next GEMM (a thread owns one point = one row of a core-matrix column block).
template <bool PACK>
__global__ void __launch_bounds__(NTHREADS, 1)
spec_gemm_tc_kernel(GemmArgs a) {
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
        // MMA issuer: the whole warp runs the loop converged and one elected lane issues.  Every operand is
        // warp-uniform, so descriptors live in uniform registers and the MMAs of a K chunk are back-to-back
        // instructions (a divergent `if (lane == 0)` makes the compiler wrap each one in an elect / branch loop,
        // ~130 clk per MMA: measured on the photometric kernel, and the reason this GEMM sat at 23 % tensor use).
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
                const uint64_t ah0 = make_desc(sa, 2048, 128), al0 = make_desc(sa + OPER_BYTES, 2048, 128);
                const uint64_t bh0 = make_desc(sa + 2 * OPER_BYTES, 2048, 128), bl0 = make_desc(sa + 3 * OPER_BYTES, 2048, 128);
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < BK / 16; ++ks) {
                        const uint64_t o = (uint64_t)(ks * (4096 / 16));          // start-address field advances by 4096 B
                        mma_ss(d, ah0 + o, bh0 + o, (kc | ks) != 0, IDESC);
                        mma_ss(d, al0 + o, bh0 + o, 1, IDESC);
                        mma_ss(d, ah0 + o, bl0 + o, 1, IDESC);
                    }
                    tc_commit(bar(BAR_EMPTY + st));                            // stage reusable once these MMAs retire
                    if (kc == nkc - 1) tc_commit(bar(BAR_ACC_FULL + buf));
                }
                __syncwarp();
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
                if constexpr (PACK) {
                    // columns of this layer = K of the next: 8 consecutive ones are one 16-byte core-matrix row
#pragma unroll
                    for (int g8 = 0; g8 < 4; ++g8) {
                        const int n0 = col0 + 8 * g8;
                        const int kc_o = n0 / BK;
                        if (kc_o >= a.nkc_out) continue;
                        __half hi[8], lo[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float x = 0.f;
                            if (n0 + i < a.P && row < a.B) {
                                x = fmaf(__uint_as_float(v[8 * g8 + i]), a.inv_scale, __ldg(a.bias + n0 + i));
                                x = x > 0.f ? x : a.leak * x;
                            }
                            hi[i] = __float2half_rn(x);
                            lo[i] = __float2half_rn(x - __half2float(hi[i]));
                        }
                        if (a.out_ts) {
                            // TMEM image: [m_tile][k-step c][hi words 0-3 | hi 4-7 | lo 0-3 | lo 4-7][row][16 B]
                            const int c = n0 / 16, t = (n0 / 8) & 1;
                            unsigned char *dst = a.out_ts + (((size_t)mt * (a.nkc_out * (BK / 16)) + c) * 4 + t) * (BM * 16) +
                                                 (size_t)(q * 32 + lane) * 16;
                            *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(hi);
                            *reinterpret_cast<uint4 *>(dst + 2 * BM * 16) = *reinterpret_cast<const uint4 *>(lo);
                            continue;
                        }
                        unsigned char *dst = a.out_img + (((size_t)mt * a.nkc_out + kc_o) * 2) * OPER_BYTES +
                                             (size_t)((n0 % BK) / 8) * 2048 + (size_t)(q * 32 + lane) * 16;
                        *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(hi);
                        *reinterpret_cast<uint4 *>(dst + OPER_BYTES) = *reinterpret_cast<const uint4 *>(lo);
                    }
                } else if (row < a.B) {
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

// ---- output layer, A-stationary form -------------------------------------------------------------------------
// The output GEMM above is bound by operand bytes into the SM (21 B/clk against the 85 a tensor-bound split-precision
// 128 x 128 tile needs), and half of those bytes are the activations -- the same 128 x K block for every pixel tile of
// an m-tile.  Here a CTA keeps the activations of ITS m-tile in tensor memory for the whole sweep (fp16 hi / lo pairs:
// 16 columns per k-step, 320 columns for K <= 320) and issues the MMAs in the TS form (A from TMEM); only the weight
// tiles stream through shared memory: 80 KB per 128 x 64 output tile instead of 160 KB, through a 12-deep ring of 16 KB
// stages.  Two 64-column accumulators behind the activations (320 + 128 <= 512 columns) keep the epilogue of a tile under
// the MMAs of the next.  Work item = (m-tile, every G-th pixel tile), G = CTAs per m-tile; m-fastest across CTAs so the
// CTAs that run together read the same weight tiles through L2.
constexpr int TBN = 64;                                   // pixel columns per tile
constexpr int TS_OPER = (BK / 8) * TBN * 16;              // one half (hi or lo) of a weight stage: 8 KB
constexpr int TS_STAGE = 2 * TS_OPER;                     // 16 KB
constexpr int TS_NST = 11;
constexpr int TS_ROWF = TBN + 4;                          // floats per staged row: 16-byte rows 17 slots apart -> no bank conflicts
constexpr int TS_SM_STAGE_OUT = TS_NST * TS_STAGE;        // [4 epilogue warps][32 rows][TS_ROWF] f32: transposition buffer
constexpr int TS_SM_BAR = TS_SM_STAGE_OUT + 4 * 32 * TS_ROWF * 4;
constexpr int TS_SM_TOTAL = TS_SM_BAR + 256;
constexpr uint32_t TS_IDESC = make_idesc_f16(BM, TBN);
enum { TB_FULL = 0, TB_EMPTY = TS_NST, TB_ACC_FULL = 2 * TS_NST, TB_ACC_EMPTY = 2 * TS_NST + 2, TB_A_READY = 2 * TS_NST + 4 };

struct TsArgs {
    const unsigned char *a_ts;       // [m_tiles][nks][4][128 rows][16 B]
    const unsigned char *b_img;      // [n_tiles (64 wide)][nkc][hi|lo][64 rows x 128 B, 128-byte swizzle]
    const float *bias;
    float *out;                      // [B][P] f32
    int64_t B;
    int P, nkc, m_tiles, n_tiles, G; // G: work items per m-tile
    float inv_scale;
};

__global__ void __launch_bounds__(NTHREADS, 1)
spec_out_gemm_ts_kernel(TsArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t s_tmem_base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sbase = smem_u32(smem);
    auto bar = [&](int i) { return sbase + TS_SM_BAR + 8u * (uint32_t)i; };
    if (threadIdx.x == 0) {
        for (int i = 0; i < TS_NST; ++i) { mbar_init(bar(TB_FULL + i), 1); mbar_init(bar(TB_EMPTY + i), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(bar(TB_ACC_FULL + i), 1); mbar_init(bar(TB_ACC_EMPTY + i), 4); }
        mbar_init(bar(TB_A_READY), 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem_base;
    const int nkc = a.nkc, nks = nkc * (BK / 16);
    const uint32_t dcol = (uint32_t)(16 * nks);              // accumulators start behind the activations
    const int nitems = a.m_tiles * a.G;

    if (warp == 0) {
        if (lane == 0) {
            int g = 0;
            for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
                for (int nt = item / a.m_tiles; nt < a.n_tiles; nt += a.G) {
                    for (int kc = 0; kc < nkc; ++kc, ++g) {
                        const int st = g % TS_NST;
                        mbar_wait(bar(TB_EMPTY + st), ((g / TS_NST) & 1) ^ 1);
                        mbar_expect_tx(bar(TB_FULL + st), TS_STAGE);
                        bulk_g2s(sbase + st * TS_STAGE, a.b_img + ((size_t)nt * nkc + kc) * TS_STAGE, TS_STAGE, bar(TB_FULL + st));
                    }
                }
            }
        }
    } else if (warp == 1) {
        int g = 0, it = 0, ni = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++ni) {
            mbar_wait(bar(TB_A_READY), ni & 1);                  // this m-tile's activations are in tensor memory
            tc_fence_after();
            for (int nt = item / a.m_tiles; nt < a.n_tiles; nt += a.G, ++it) {
                const int buf = it & 1;
                mbar_wait(bar(TB_ACC_EMPTY + buf), ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d = tmem + dcol + buf * TBN;
                for (int kc = 0; kc < nkc; ++kc, ++g) {
                    const int st = g % TS_NST;
                    mbar_wait(bar(TB_FULL + st), (g / TS_NST) & 1);
                    tc_fence_after();
                    const uint32_t sa = sbase + st * TS_STAGE;
                    const uint64_t bh0 = make_desc_sw128(sa), bl0 = make_desc_sw128(sa + TS_OPER);
                    const uint32_t a0 = tmem + 16 * (kc * (BK / 16));
                    if (elect_one()) {
#pragma unroll
                        for (int ks = 0; ks < BK / 16; ++ks) {
                            const uint64_t o = (uint64_t)(ks * 2);                        // 32 B further along the 128-byte row
                            const uint32_t ah = a0 + 16 * ks;
                            mma_ts(d, ah, bh0 + o, (kc | ks) != 0, TS_IDESC);
                            mma_ts(d, ah + 8, bh0 + o, 1, TS_IDESC);
                            mma_ts(d, ah, bl0 + o, 1, TS_IDESC);
                        }
                        tc_commit(bar(TB_EMPTY + st));
                        if (kc == nkc - 1) tc_commit(bar(TB_ACC_FULL + buf));
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        const int q = warp & 3;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        int it = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const int mt = item % a.m_tiles;
            // activations of the m-tile -> tensor memory (every MMA of the previous item has retired: its last
            // accumulator was drained below)
            {
                const unsigned char *src = a.a_ts + ((size_t)mt * nks * 4) * (BM * 16) + (size_t)(q * 32 + lane) * 16;
                for (int c = 0; c < nks; ++c) {
                    uint32_t v[16];
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const uint4 w = __ldg(reinterpret_cast<const uint4 *>(src + ((size_t)c * 4 + t) * (BM * 16)));
                        v[4 * t] = w.x; v[4 * t + 1] = w.y; v[4 * t + 2] = w.z; v[4 * t + 3] = w.w;
                    }
                    TMEM_ST16(tmem + lane_base + 16 * c, v);
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(TB_A_READY));
            }
            const int64_t row = (int64_t)mt * BM + q * 32 + lane;
            for (int nt = item / a.m_tiles; nt < a.n_tiles; nt += a.G, ++it) {
                const int buf = it & 1;
                mbar_wait(bar(TB_ACC_FULL + buf), (it >> 1) & 1);
                tc_fence_after();
                // A thread owns a ROW of the accumulator; storing it directly would put 32 rows, 40 KB apart, into every
                // store instruction (half-filled sectors: 23 of the 55 us of this kernel).  The warp's 32 x 64 block goes
                // through shared memory instead and leaves as whole 256-byte row segments, two rows per instruction.
                float *stg = reinterpret_cast<float *>(smem + TS_SM_STAGE_OUT) + (size_t)q * 32 * TS_ROWF;
#pragma unroll 1
                for (int j = 0; j < TBN / 32; ++j) {
                    uint32_t v[32];
                    TMEM_LD32(tmem + lane_base + dcol + buf * TBN + 32 * j, v);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        *reinterpret_cast<uint4 *>(stg + lane * TS_ROWF + 32 * j + 4 * i) = make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(TB_ACC_EMPTY + buf));      // the accumulator is free: the MMAs of tile it + 2 may start
                {
                    const int col0 = nt * TBN, hl = lane & 15, rsel = lane >> 4;
                    const int64_t row0 = (int64_t)mt * BM + q * 32;
                    if (col0 + TBN <= a.P) {
                        const float4 bb = __ldg(reinterpret_cast<const float4 *>(a.bias + col0) + hl);
#pragma unroll 4
                        for (int rr = 0; rr < 32; rr += 2) {
                            const int r = rr + rsel;
                            const float4 x = *reinterpret_cast<const float4 *>(stg + r * TS_ROWF + 4 * hl);
                            float4 r4;
                            r4.x = fmaf(x.x, a.inv_scale, bb.x); r4.y = fmaf(x.y, a.inv_scale, bb.y);
                            r4.z = fmaf(x.z, a.inv_scale, bb.z); r4.w = fmaf(x.w, a.inv_scale, bb.w);
                            if (row0 + r < a.B) *reinterpret_cast<float4 *>(a.out + (row0 + r) * a.P + col0 + 4 * hl) = r4;
                        }
                    } else {
                        for (int rr = 0; rr < 32; rr += 2) {
                            const int r = rr + rsel;
                            for (int i = 0; i < 4; ++i) {
                                const int c = col0 + 4 * hl + i;
                                if (c < a.P && row0 + r < a.B)
                                    a.out[(row0 + r) * a.P + c] = fmaf(stg[r * TS_ROWF + 4 * hl + i], a.inv_scale, a.bias[c]);
                            }
                        }
                    }
                    __syncwarp();                                       // the buffer is reused by the next tile
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

}  // namespace

struct TcLayer {                      // packed weights of one GEMM layer
    unsigned char *b_img = nullptr;
    int nkc = 0, n_tiles = 0, N = 0;
    float inv_scale = 1.f;
};
struct SpecTcPack {
    TcLayer l1, l2;
    TcLayer l2ts;                                            // output layer in 64-wide tiles (A-stationary kernel); empty if K is too long
    unsigned char *a2_ts = nullptr;                          // its A operand as a TMEM image
    unsigned char *a1_img = nullptr, *a2_img = nullptr;     // operand images of layer 1 and of the output layer
    size_t a_img_rows = 0;                                   // rows (multiple of BM) both images hold
};

// W[N][K] f32 (device) -> B image [n_tile][k_chunk][hi|lo][k8][row][8], scaled by a power of two
static int pack_weights(const float *w_dev, int N, int K, TcLayer *L, int bn = BN, bool sw128 = false) {
    std::vector<float> w((size_t)N * K);
    MS_CUDA(cudaMemcpy(w.data(), w_dev, w.size() * 4, cudaMemcpyDeviceToHost));
    float wmax = 0.f;
    for (float v : w) wmax = std::fabs(v) > wmax ? std::fabs(v) : wmax;
    int e = 0;
    if (wmax > 0.f) { std::frexp(wmax, &e); }
    const int shift = 14 - e;                               // max |w| * 2^shift in [2^13, 2^14)
    const double scale = std::ldexp(1.0, shift);
    L->N = N;
    L->nkc = (K + BK - 1) / BK;
    L->n_tiles = (N + bn - 1) / bn;
    L->inv_scale = (float)std::ldexp(1.0, -shift);
    const size_t oper = (size_t)(BK / 8) * bn * 16;          // one half (hi or lo) of a (tile, K chunk) block
    const size_t bytes = (size_t)L->n_tiles * L->nkc * 2 * oper;
    std::vector<__half> img(bytes / 2, __float2half_rn(0.f));
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) {
            const double v = scale * (double)w[(size_t)n * K + k];
            const __half hi = __float2half_rn((float)v);
            const __half lo = __float2half_rn((float)(v - (double)__half2float(hi)));
            const int r = n % bn, kk = k % BK;
            // no swizzle: [k / 8][row][k % 8]; 128-byte swizzle: [row / 8][row % 8][(k / 8) ^ (row % 8)][k % 8]
            const size_t in_tile = sw128 ? (size_t)(r / 8) * 512 + (size_t)(r % 8) * 64 + (size_t)((kk / 8) ^ (r % 8)) * 8 + (kk % 8)
                                         : (size_t)(kk / 8) * (bn * 8) + (size_t)r * 8 + (kk % 8);
            const size_t at = (((size_t)(n / bn) * L->nkc + k / BK) * 2) * (oper / 2) + in_tile;
            img[at] = hi;
            img[at + oper / 2] = lo;
        }
    MS_CUDA(cudaMalloc((void **)&L->b_img, bytes));
    MS_CUDA(cudaMemcpy(L->b_img, img.data(), bytes, cudaMemcpyHostToDevice));
    return MS_OK;
}

int spec_tc_prepare(ms_ctx *ctx) {
    SpecNetDev &S = ctx->spec;
    if (S.npix % 4 != 0) { ms_set_error("MS_PREC_TC needs a pixel count that is a multiple of 4"); return MS_EINVAL; }
    SpecTcPack *pk = new SpecTcPack();
    S.tc_pack = pk;                                        // owned by the ctx from here: ms_destroy releases it on any error
    int rc;
    if ((rc = pack_weights(S.w1, S.h, S.h, &pk->l1))) return rc;
    if ((rc = pack_weights(S.w2, S.npix, S.h, &pk->l2))) return rc;
    if (16 * pk->l2.nkc * (BK / 16) + 2 * TBN <= 512) {     // activations + two accumulators fit tensor memory
        if ((rc = pack_weights(S.w2, S.npix, S.h, &pk->l2ts, TBN, true))) return rc;
        MS_CUDA(cudaFuncSetAttribute(spec_out_gemm_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SM_TOTAL));
    }
    MS_CUDA(cudaFuncSetAttribute(spec_gemm_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
    MS_CUDA(cudaFuncSetAttribute(spec_gemm_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
    return MS_OK;
}

void spec_tc_release(ms_ctx *ctx) {
    SpecTcPack *pk = static_cast<SpecTcPack *>(ctx->spec.tc_pack);
    if (!pk) return;
    cudaFree(pk->l1.b_img); cudaFree(pk->l2.b_img); cudaFree(pk->l2ts.b_img); cudaFree(pk->a1_img); cudaFree(pk->a2_img);
    cudaFree(pk->a2_ts);
    delete pk;
    ctx->spec.tc_pack = nullptr;
}

// operand-image scratch for chunks of up to `chunk` points
// whether spec_tc_reserve(ctx, chunk, nchunks) would have to allocate (run_spec refuses that during stream capture)
bool spec_tc_would_grow(const ms_ctx *ctx, int64_t chunk, int nchunks) {
    const SpecTcPack *pk = static_cast<const SpecTcPack *>(ctx->spec.tc_pack);
    if (!pk) return false;
    return pk->a_img_rows < (size_t)((chunk + BM - 1) / BM) * BM * (size_t)(nchunks < 1 ? 1 : nchunks);
}

// chunk: points per chunk; nchunks: chunks whose hidden layers are computed in one go (spec_tc_hidden)
int spec_tc_reserve(ms_ctx *ctx, int64_t chunk, int nchunks) {
    SpecTcPack *pk = static_cast<SpecTcPack *>(ctx->spec.tc_pack);
    if (!pk) { ms_set_error("spectral net not packed for MS_PREC_TC"); return MS_ESTATE; }
    const size_t rows = (size_t)((chunk + BM - 1) / BM) * BM * (size_t)(nchunks < 1 ? 1 : nchunks);
    if (pk->a_img_rows >= rows) return MS_OK;
    if (ctx->scratch_locked) { ms_set_error("spectrum scratch reserved for a smaller batch (ms_reserve)"); return MS_ESTATE; }
    cudaFree(pk->a1_img); cudaFree(pk->a2_img); cudaFree(pk->a2_ts);
    pk->a1_img = pk->a2_img = pk->a2_ts = nullptr; pk->a_img_rows = 0;
    MS_CUDA(cudaMalloc((void **)&pk->a1_img, rows / BM * pk->l1.nkc * 2 * OPER_BYTES));
    MS_CUDA(cudaMalloc((void **)&pk->a2_img, rows / BM * pk->l2.nkc * 2 * OPER_BYTES));
    if (pk->l2ts.b_img) MS_CUDA(cudaMalloc((void **)&pk->a2_ts, rows / BM * pk->l2.nkc * (BK / 16) * 4 * BM * 16));
    pk->a_img_rows = rows;
    return MS_OK;
}

// native flux[B][P] (f32) of a chunk of points from their parameter block: the three layers on one stream
int launch_spec_mlp_tc(ms_ctx *ctx, const double *pars, int64_t B, float *flux, cudaStream_t st) {
    SpecNetDev &S = ctx->spec;
    SpecTcPack *pk = static_cast<SpecTcPack *>(S.tc_pack);
    if (!pk) { ms_set_error("spectral net not packed for MS_PREC_TC"); return MS_ESTATE; }
    int rc;
    if ((rc = spec_tc_reserve(ctx, B, 1))) return rc;
    const int m_tiles = (int)((B + BM - 1) / BM);
    KernelTimer timer_(ctx, MS_K_SPEC_MLP, st);
    {
        const int64_t ntask = (int64_t)m_tiles * BM * pk->l1.nkc * (BK / 8);
        int blocks = (int)((ntask + 255) / 256);
        if (blocks > ctx->sm_count * 16) blocks = ctx->sm_count * 16;
        spec_l0_pack_kernel<<<blocks, 256, 0, st>>>(pars, B, S.d_in, S.logt_input, S.xms, S.w0, S.b0, S.h, (float)S.leak,
                                                    pk->l1.nkc, reinterpret_cast<__half *>(pk->a1_img), (int64_t)m_tiles * BM, 0, 0);
        ctx->launches++;
        MS_CUDA(cudaGetLastError());
    }
    GemmArgs ga;
    ga.B = B; ga.m_tiles = m_tiles; ga.leak = (float)S.leak;
    // layer 1: H -> H, epilogue packs the operand image of the output layer
    const bool ts = pk->l2ts.b_img != nullptr && ctx->spec_gemm_ts;
    ga.a_img = pk->a1_img; ga.b_img = pk->l1.b_img; ga.bias = S.b1; ga.out = nullptr; ga.out_img = pk->a2_img;
    ga.out_ts = ts ? pk->a2_ts : nullptr;
    ga.P = S.h; ga.nkc = pk->l1.nkc; ga.n_tiles = pk->l1.n_tiles; ga.nkc_out = pk->l2.nkc; ga.inv_scale = pk->l1.inv_scale;
    {
        const int ntiles = m_tiles * ga.n_tiles;
        spec_gemm_tc_kernel<true><<<ntiles < ctx->sm_count ? ntiles : ctx->sm_count, NTHREADS, SM_TOTAL, st>>>(ga);
        ctx->launches++;
        MS_CUDA(cudaGetLastError());
    }
    // output layer: A-stationary form when the activations fit tensor memory
    if (ts) {
        TsArgs ta;
        ta.a_ts = pk->a2_ts; ta.b_img = pk->l2ts.b_img; ta.bias = S.b2; ta.out = flux; ta.B = B; ta.P = S.npix;
        ta.nkc = pk->l2ts.nkc; ta.m_tiles = m_tiles; ta.n_tiles = pk->l2ts.n_tiles; ta.inv_scale = pk->l2ts.inv_scale;
        ta.G = ctx->sm_count / m_tiles;
        if (ta.G < 1) ta.G = 1;
        if (ta.G > ta.n_tiles) ta.G = ta.n_tiles;
        const int nitems = m_tiles * ta.G;
        spec_out_gemm_ts_kernel<<<nitems < ctx->sm_count ? nitems : ctx->sm_count, NTHREADS, TS_SM_TOTAL, st>>>(ta);
        ctx->launches++;
        MS_CUDA(cudaGetLastError());
        return MS_OK;
    }
    ga.out_ts = nullptr;
    ga.a_img = pk->a2_img; ga.b_img = pk->l2.b_img; ga.bias = S.b2; ga.out = flux; ga.out_img = nullptr;
    ga.P = S.npix; ga.nkc = pk->l2.nkc; ga.n_tiles = pk->l2.n_tiles; ga.nkc_out = 0; ga.inv_scale = pk->l2.inv_scale;
    {
        const int ntiles = m_tiles * ga.n_tiles;
        spec_gemm_tc_kernel<false><<<ntiles < ctx->sm_count ? ntiles : ctx->sm_count, NTHREADS, SM_TOTAL, st>>>(ga);
        ctx->launches++;
        MS_CUDA(cudaGetLastError());
    }
    return MS_OK;
}

// Hidden layers of a whole super-batch in one go (two launches instead of two per chunk: the 300 x 300 layer of one
// 1480-row chunk is 36 tiles on 148 SMs, 14 us of latency): rows [0, B) of `pars`, laid out chunk by chunk with every chunk
// padded to whole m-tiles so that spec_tc_out can sweep a chunk's own m-tiles.  Returns 1 when the A-stationary output
// kernel is not available (the caller then uses launch_spec_mlp_tc per chunk), 0 on success, < 0 on error.
int spec_tc_hidden(ms_ctx *ctx, const double *pars, int64_t B, int64_t chunk, cudaStream_t st) {
    SpecNetDev &S = ctx->spec;
    SpecTcPack *pk = static_cast<SpecTcPack *>(S.tc_pack);
    if (!pk) { ms_set_error("spectral net not packed for MS_PREC_TC"); return MS_ESTATE; }
    if (!pk->l2ts.b_img || !ctx->spec_gemm_ts) return 1;
    const int64_t nchunks = (B + chunk - 1) / chunk;
    const int64_t chunk_pad = (chunk + BM - 1) / BM * BM;
    int rc;
    if ((rc = spec_tc_reserve(ctx, chunk, (int)nchunks))) return rc;
    const int64_t rows_img = nchunks * chunk_pad;
    const int m_tiles = (int)(rows_img / BM);
    KernelTimer timer_(ctx, MS_K_SPEC_MLP, st);
    {
        const int64_t ntask = rows_img * pk->l1.nkc * (BK / 8);
        int64_t blocks = (ntask + 255) / 256;
        if (blocks > (int64_t)ctx->sm_count * 16) blocks = (int64_t)ctx->sm_count * 16;
        spec_l0_pack_kernel<<<(unsigned)blocks, 256, 0, st>>>(pars, B, S.d_in, S.logt_input, S.xms, S.w0, S.b0, S.h, (float)S.leak,
                                                             pk->l1.nkc, reinterpret_cast<__half *>(pk->a1_img), rows_img,
                                                             (int)chunk, (int)chunk_pad);
        ctx->launches++;
        MS_CUDA(cudaGetLastError());
    }
    GemmArgs ga;
    ga.B = rows_img; ga.m_tiles = m_tiles; ga.leak = (float)S.leak;
    ga.a_img = pk->a1_img; ga.b_img = pk->l1.b_img; ga.bias = S.b1; ga.out = nullptr; ga.out_img = pk->a2_img;
    ga.out_ts = pk->a2_ts;
    ga.P = S.h; ga.nkc = pk->l1.nkc; ga.n_tiles = pk->l1.n_tiles; ga.nkc_out = pk->l2.nkc; ga.inv_scale = pk->l1.inv_scale;
    const int ntiles = m_tiles * ga.n_tiles;
    spec_gemm_tc_kernel<true><<<ntiles < ctx->sm_count ? ntiles : ctx->sm_count, NTHREADS, SM_TOTAL, st>>>(ga);
    ctx->launches++;
    MS_CUDA(cudaGetLastError());
    return MS_OK;
}

// Output layer of chunk `ci` (nb rows) of the super-batch prepared by spec_tc_hidden
int spec_tc_out(ms_ctx *ctx, int64_t ci, int64_t chunk, int64_t nb, float *flux, cudaStream_t st) {
    SpecNetDev &S = ctx->spec;
    SpecTcPack *pk = static_cast<SpecTcPack *>(S.tc_pack);
    const int64_t mt_per_chunk = (chunk + BM - 1) / BM;
    const int m_tiles = (int)((nb + BM - 1) / BM);
    KernelTimer timer_(ctx, MS_K_SPEC_MLP, st);
    TsArgs ta;
    ta.a_ts = pk->a2_ts + (size_t)ci * mt_per_chunk * pk->l2ts.nkc * (BK / 16) * 4 * BM * 16;
    ta.b_img = pk->l2ts.b_img; ta.bias = S.b2; ta.out = flux; ta.B = nb; ta.P = S.npix;
    ta.nkc = pk->l2ts.nkc; ta.m_tiles = m_tiles; ta.n_tiles = pk->l2ts.n_tiles; ta.inv_scale = pk->l2ts.inv_scale;
    ta.G = ctx->sm_count / m_tiles;
    if (ta.G < 1) ta.G = 1;
    if (ta.G > ta.n_tiles) ta.G = ta.n_tiles;
    const int nitems = m_tiles * ta.G;
    spec_out_gemm_ts_kernel<<<nitems < ctx->sm_count ? nitems : ctx->sm_count, NTHREADS, TS_SM_TOTAL, st>>>(ta);
    ctx->launches++;
    MS_CUDA(cudaGetLastError());
    return MS_OK;
}
