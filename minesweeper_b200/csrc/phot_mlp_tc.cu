// Kernel (2), tensor-core form (MS_PREC_TC): the photometric MLP stack
// 6 -> 128 -> 128 -> 1 per band on tcgen05 with TMEM accumulators, fused with
// the magnitude formula and the photometric chi-square.  Same contract as
// phot_mlp.cu (GenMod.genphot, reference genmod.py:97-142; FastPayneSEDPredict.sed,
// external Payne, PARITY UNPINNED; likelihood.lnlike, likelihood.py:204-210).
//
// Numerics: every operand is split into two fp16 halves (hi + lo, 22 significant
// bits) and each GEMM is three kind::f16 MMAs (hi.hi + lo.hi + hi.lo) accumulated
// in fp32, i.e. fp32-class products at one third of the fp16 tensor rate.  The
// factor -log2(e) and the layer-1 bias are folded into the packed weights, so the
// accumulators hold the exponent argument of the sigmoid directly.
//
// Structure (one persistent CTA per SM, 736 threads):
//   warps 0-15       four epilogue warpgroups.  Tile T (128 points = 128 TMEM lanes) is shared by
//                    warpgroup A_T (warps 4T..4T+3) and B_T (warps 8+4T..): a thread owns one point,
//                    A takes the even 16-column chunks of each accumulator, B the odd ones, so four
//                    warps per scheduler are available to hide the MUFU / TMEM latencies.  A also does
//                    the chi-square.
//   warps 16, 17     MMA issuers, one per tile (one elected lane each, blocking on that tile's barriers)
//   warp 18          weight producer: bulk async copies of the per-band weight image
//                    (global/L2 -> shared) into a two-deep ring
//   warps 19-22      prologue warpgroup, one pair ahead: the next pair's network inputs (X tiles) and distance
//                    moduli and, in the fused form, kernel (1) itself: the MIST interpolation of those points
//                    (mist_interp.cuh), so a photometry-only batch is one launch and the parameter block never
//                    exists in HBM
// Per tile and band:
//   D1 = X W1^T            SS MMA, K = 16 (6 inputs + bias column), X tile in shared memory
//   A1 = act(D1)           TMEM -> registers -> activation -> fp16 (hi/lo) -> TMEM, written in place
//                          over D1 so the activations never touch shared memory
//   D2 = A1 W2^T           TS MMA (A operand read from TMEM), K = 128, one k-step per 16-column
//                          chunk, issued as soon as that chunk has landed (K-pipelined)
//   BC = w3 . act(D2 + b2) + b3 ; magnitude ; chi-square term     (registers; B's half of the dot
//                          product goes to A through shared memory)
// TMEM: 2 tiles x (128 columns D1/A1 + 128 columns D2) = 512 columns.
#include "common.cuh"
#include "mist_interp.cuh"
#include "tc_ptx.cuh"

#include <cuda_fp16.h>

namespace {

constexpr int KX = 16;             // padded layer-1 K (d_in <= 6, bias column at 6)
constexpr int TILE = 128;          // points per tile = TMEM lanes
constexpr int NTHREADS = 736;
constexpr int X_BYTES = 2 * 128 * 16;                  // one X tile image (hi or lo)
constexpr int SMEM_LIMIT = 227 * 1024 - 4096;          // opt-in maximum minus the static part and the reserved KB

// what the prologue warpgroup hands to the chi-square threads for one point
struct PointState { double mag0, agewgt; int32_t slot; int32_t ok; };
static_assert(sizeof(PointState) == 24, "PointState layout");

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// Sigmoid of four accumulators that already hold t = -log2(e) z: 1 / (1 + 2^t) with one rcp
// shared by the four (1.25 MUFU per activation).  The clamp keeps the product of the four
// denominators inside the fp32 range (sigmoid error <= 2^-30).  On B200 this code is bound by
// instruction issue, not by the MUFU pipe alone (tools/mufu_bench.cu: 16 ex2/clk/SM alone, but
// ~12.9 elements/clk/SM once 7 FP32 instructions accompany each MUFU), so every instruction
// per activation counts.
__device__ __forceinline__ void sigmoid_quad(float t0, float t1, float t2, float t3, float &a0, float &a1,
                                             float &a2, float &a3) {
    const float d0 = 1.0f + ex2_approx(fminf(t0, 30.0f));
    const float d1 = 1.0f + ex2_approx(fminf(t1, 30.0f));
    const float d2 = 1.0f + ex2_approx(fminf(t2, 30.0f));
    const float d3 = 1.0f + ex2_approx(fminf(t3, 30.0f));
    const float p01 = d0 * d1, p23 = d2 * d3;
    const float r = rcp_approx(p01 * p23);
    const float r01 = r * p23, r23 = r * p01;
    a0 = r01 * d1; a1 = r01 * d0; a2 = r23 * d3; a3 = r23 * d2;
}

// w . sigmoid of four accumulators for the last layer, where only the weighted sum of the activations is needed:
// sum_i w_i / d_i = [ (w0 d1 + w1 d0) d2 d3 + (w2 d3 + w3 d2) d0 d1 ] / (d0 d1 d2 d3), d_i = 1 + 2^t_i -- one rcp per four
// as in sigmoid_quad, but 23 instead of 26 instructions per quad (the four activations are never formed).
__device__ __forceinline__ float sigmoid_dot_quad(float t0, float t1, float t2, float t3, float4 w) {
    const float d0 = 1.0f + ex2_approx(fminf(t0, 30.0f));
    const float d1 = 1.0f + ex2_approx(fminf(t1, 30.0f));
    const float d2 = 1.0f + ex2_approx(fminf(t2, 30.0f));
    const float d3 = 1.0f + ex2_approx(fminf(t3, 30.0f));
    const float p01 = d0 * d1, p23 = d2 * d3;
    const float n01 = fmaf(w.y, d0, w.x * d1), n23 = fmaf(w.w, d2, w.z * d3);
    const float num = fmaf(n23, p01, n01 * p23);
    return num * rcp_approx(p01 * p23);
}

// a - (float)h for both halves of a packed fp16 pair, one mixed-precision FMA each (FHFMA)
__device__ __forceinline__ void sub_half2(float a0, float a1, __half2 h, float &l0, float &l1) {
    const uint32_t hb = *reinterpret_cast<const uint32_t *>(&h);
    unsigned short lo, hi;
    asm("mov.b32 {%0,%1}, %2;" : "=h"(lo), "=h"(hi) : "r"(hb));
    const unsigned short m1 = 0xbc00;                  // -1.0h
    asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(l0) : "h"(lo), "h"(m1), "f"(a0));
    asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(l1) : "h"(hi), "h"(m1), "f"(a1));
}

__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
    uint32_t p;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(hi), "f"(lo));
    return p;
}

struct TcArgs {
    const double *pars; const int32_t *star_slot;
    double *mags; double *chi2; double *lnL; int ageweight;
    int slot0, nslots;                 // constant slot when star_slot is nullptr; loaded slots (bounds check)
    // multi-GPU gather fused into the epilogue: lnL[p] is also stored at peers[r][peer_off + p] for every rank r of the
    // box (peer memory mapped over NVLink / NVSwitch; ms_set_lnl_peers).  No collective kernel, no copy.
    double *peers[MS_MAX_PEERS]; int npeers; int64_t peer_off;
    const double *obs_mag, *obs_err2;
    const unsigned char *wimg;         // [nb][BAND_BYTES]
    int stage;                         // fused form: rows of theta staged through shared memory (ld <= 8, 16-byte aligned)
    const int *rows_ready;             // streaming host entry: rows of theta on the device so far (nullptr: all)
    int *sched;                        // [2]: next pair beyond the first wave, CTAs that have left (both 0 between launches)
    double xmin[8], xscale[8];
    int nb, d_in;                      // bands of this launch (a group when the net has more than MAX_NB)
    int nb_total, b0;                  // row length of obs / mags and the first band of the group
    // fused form: the prologue warpgroup interpolates the MIST grid itself (kernel (1) folded in), so the
    // parameter block never exists in HBM; `pars` is then unused
    int fused;
    ThetaView tv;
    const double *axes; const float *table;
    int ne, nm, nf, na, mistdif;
    double *derived;                   // [B][MS_NDERIVED] or nullptr
};

#ifdef MS_PHOT_STAMPS
// tuning builds only (MS_BUILD_VARIANT): clock64 stamps of CTA 0, one fixed slot per (pair, event), read back once
__device__ long long g_stamps[8192];
__device__ long long g_cta[148 * 8];
#define STAMP(pair, ev) do { if (blockIdx.x == 0 && (pair) < 1000) g_stamps[(pair) * 8 + (ev)] = clock64(); } while (0)
#else
#define STAMP(pair, ev) do { } while (0)
#endif

// one instantiation of the kernel family per supported hidden width
struct TcVariant {
    int h, band_bytes, const_bytes, sm_fixed, max_nb, pair_rows;     // pair_rows: points a CTA takes at a time
    int (*pack)(ms_ctx *, std::vector<unsigned char> &);
    void (*kern)(TcArgs, int64_t, int);
    void (*kern_fast)(TcArgs, int64_t, int);
};

namespace h128 {
#define PH_H 128
#include "phot_mlp_tc_body.cuh"
#undef PH_H
}  // namespace h128
namespace h64 {
#define PH_H 64
#include "phot_mlp_tc_body.cuh"
#undef PH_H
}  // namespace h64
namespace h256 {
#define PH_H 256
#include "phot_mlp_tc_body.cuh"
#undef PH_H
}  // namespace h256

const TcVariant *variant_for(int h) {
    if (h == 128) return &h128::variant;
    if (h == 64) return &h64::variant;
    if (h == 256) return &h256::variant;
    return nullptr;
}

}  // namespace

int phot_tc_prepare(ms_ctx *ctx) {
    PhotNetDev &P = ctx->phot;
    const TcVariant *V = variant_for(P.h);
    if (!V || P.d_in > 6) {
        ms_set_error("MS_PREC_TC supports photometric nets with hidden width 64, 128 or 256 and <= 6 inputs "
                     "(got h=%d, d_in=%d); use MS_PREC_F32", P.h, P.d_in);
        return MS_EINVAL;
    }
    std::vector<unsigned char> img;
    int rc = V->pack(ctx, img);
    if (rc) return rc;
    MS_CUDA(cudaMalloc(&P.tc_pack, img.size() + 64));            // + the pair scheduler's two counters
    MS_CUDA(cudaMemcpy(P.tc_pack, img.data(), img.size(), cudaMemcpyHostToDevice));
    MS_CUDA(cudaMemset(static_cast<unsigned char *>(P.tc_pack) + img.size(), 0, 64));
    P.tc_pack_bytes = img.size();
    MS_CUDA(cudaFuncSetAttribute(V->kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    MS_CUDA(cudaFuncSetAttribute(V->kern_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    return MS_OK;
}

namespace {

inline int ngroups_of(int nb, int max_nb) { return (nb + max_nb - 1) / max_nb; }

int launch_tc(ms_ctx *ctx, const PhotArgs &a, int64_t B, cudaStream_t st, const ThetaView *tv, double *derived,
              int mistdif) {
    const PhotNetDev &n = ctx->phot;
    const TcVariant *V = variant_for(n.h);
    if (!n.tc_pack || !V) { ms_set_error("photometric net not packed for MS_PREC_TC"); return MS_ESTATE; }
    if (B <= 0) return MS_OK;
    if ((a.chi2 || a.lnL) && !ctx->obs_mag) { ms_set_error("no star observations set"); return MS_ESTATE; }
    const int MAX_NB = V->max_nb, BAND_BYTES = V->band_bytes, CONST_BYTES = V->const_bytes, SM_FIXED = V->sm_fixed;
    TcArgs ka;
    ka.pars = a.pars; ka.star_slot = a.star_slot; ka.mags = a.mags; ka.ageweight = a.ageweight;
    ka.slot0 = a.slot0; ka.nslots = ctx->nslots;
    ka.npeers = a.lnL ? ctx->npeers : 0; ka.peer_off = ctx->peer_off;
    for (int i = 0; i < MS_MAX_PEERS; ++i) ka.peers[i] = i < ka.npeers ? ctx->peers[i] : nullptr;
    ka.obs_mag = ctx->obs_mag; ka.obs_err2 = ctx->obs_err2;
    for (int i = 0; i < 8; ++i) { ka.xmin[i] = n.xmin[i]; ka.xscale[i] = n.xscale[i]; }
    ka.d_in = n.d_in; ka.nb_total = n.nb;
    ka.rows_ready = tv ? a.rows_ready : nullptr;
    ka.sched = reinterpret_cast<int *>(static_cast<unsigned char *>(n.tc_pack) + n.tc_pack_bytes);
    ka.fused = tv != nullptr;
    int smem_extra = 0;
    if (tv) {
        const GridDev &g = ctx->grid;
        ka.tv = *tv; ka.axes = g.axes; ka.table = g.tab32;
        ka.ne = g.ne; ka.nm = g.nm; ka.nf = g.nf; ka.na = g.na; ka.mistdif = mistdif; ka.derived = derived;
        smem_extra = (g.ne + g.nm + g.nf + g.na) * (int)sizeof(double);
        // staging buffer for a tile's rows of theta, when it fits next to everything else
        const int stage_bytes = 16 + TILE * 8 * (int)sizeof(double);
        ka.stage = a.stage_theta && tv->ld <= 8 && (reinterpret_cast<uintptr_t>(tv->theta) & 15) == 0 &&
                   SM_FIXED + n.nb * CONST_BYTES + smem_extra + stage_bytes <= SMEM_LIMIT && ngroups_of(n.nb, MAX_NB) == 1;
        if (ka.stage) smem_extra += stage_bytes;
    } else {
        ka.stage = 0;
        ka.axes = nullptr; ka.table = nullptr; ka.ne = ka.nm = ka.nf = ka.na = 0; ka.mistdif = 0; ka.derived = nullptr;
    }
    // A net with more bands than the resident constants allow runs as consecutive band groups: every group but
    // the last adds its chi-square to a carry buffer, the last one turns the sum into lnL (two-kernel form only).
    const int ngroups = (n.nb + MAX_NB - 1) / MAX_NB;
    double *carry = a.chi2;
    if (ngroups > 1) {
        if (tv) { ms_set_error("fused interpolation + photometry needs <= %d bands", MAX_NB); return MS_ESTATE; }
        if (!carry && a.lnL) {
            int rc = ensure_scratch(ctx, &ctx->scr_chi2g, &ctx->scr_chi2g_n, (size_t)B, st);
            if (rc) return rc;
            carry = ctx->scr_chi2g;
            MS_CUDA(cudaMemsetAsync(carry, 0, (size_t)B * sizeof(double), st));
        }
    }
    const int64_t npairs = (B + V->pair_rows - 1) / V->pair_rows;
    const int grid = (int)(npairs < ctx->sm_count ? npairs : ctx->sm_count);
    KernelTimer timer_(ctx, MS_K_PHOT, st);
    for (int gi = 0; gi < ngroups; ++gi) {
        const bool last = gi == ngroups - 1;
        ka.b0 = gi * MAX_NB;
        ka.nb = last ? n.nb - ka.b0 : MAX_NB;
        ka.wimg = static_cast<const unsigned char *>(n.tc_pack) + (size_t)ka.b0 * BAND_BYTES;
        ka.chi2 = carry;
        ka.lnL = last ? a.lnL : nullptr;
        const int smem_total = SM_FIXED + ka.nb * CONST_BYTES + smem_extra;
        if (ctx->precision == MS_PREC_TC16) V->kern_fast<<<grid, NTHREADS, smem_total, st>>>(ka, B, (int)npairs);
        else V->kern<<<grid, NTHREADS, smem_total, st>>>(ka, B, (int)npairs);
        ctx->launches++;
        MS_CUDA(cudaGetLastError());
    }
    return MS_OK;
}

}  // namespace

int launch_phot_tc(ms_ctx *ctx, const PhotArgs &a, int64_t B, cudaStream_t st) {
    return launch_tc(ctx, a, B, st, nullptr, nullptr, 0);
}

// shared memory the fused form needs on top of the kernel's own, and whether it fits
bool phot_tc_can_fuse(const ms_ctx *ctx) {
    const GridDev &g = ctx->grid;
    const TcVariant *V = variant_for(ctx->phot.h);
    if (!ctx->phot.tc_pack || !g.tab32 || !g.axes || !V) return false;
    return ctx->phot.nb <= V->max_nb &&
           V->sm_fixed + ctx->phot.nb * V->const_bytes + (g.ne + g.nm + g.nf + g.na) * (int)sizeof(double) <= SMEM_LIMIT;
}

// whether the fused form would stage theta through shared memory for rows of `ld` doubles (the host entry reads pinned
// host memory in place only then)
bool phot_tc_can_stage(const ms_ctx *ctx, int ld) {
    const GridDev &g = ctx->grid;
    const TcVariant *V = variant_for(ctx->phot.h);
    if (!phot_tc_can_fuse(ctx) || ld > 8) return false;
    return V->sm_fixed + ctx->phot.nb * V->const_bytes + (g.ne + g.nm + g.nf + g.na) * (int)sizeof(double) + 16 +
           TILE * 8 * (int)sizeof(double) <= SMEM_LIMIT;
}

int launch_phot_tc_fused(ms_ctx *ctx, const ThetaView &tv, const PhotArgs &a, double *derived, int mistdif,
                         int64_t B, cudaStream_t st) {
    if (!phot_tc_can_fuse(ctx)) { ms_set_error("fused interpolation + photometry does not fit this model"); return MS_ESTATE; }
    return launch_tc(ctx, a, B, st, &tv, derived, mistdif);
}

#ifdef MS_PHOT_STAMPS
extern "C" int ms_debug_read_cta(long long *out) { return (int)cudaMemcpyFromSymbol(out, g_cta, sizeof(long long) * 148 * 8); }
extern "C" int ms_debug_read_stamps(long long *out, int n) {
    return (int)cudaMemcpyFromSymbol(out, g_stamps, sizeof(long long) * (size_t)(n < 8192 ? n : 8192));
}
#endif
