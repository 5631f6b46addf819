// H-dependent part of phot_mlp_tc.cu: included once per supported hidden width with PH_H defined, inside a
// namespace of its own.  TMEM: NT tiles x (H columns D1/A1 + H columns D2): two tiles for H <= 128, ONE for H = 256
// (512 columns), whose layer-2 weights (256 KB per band as hi + lo) do not fit shared memory either and are streamed
// through the ring in K slices of 32 instead of as one image per band.
constexpr int H = PH_H;              // hidden width of this instantiation
// chunks of an epilogue loop in flight per thread (measured at H = 128, ms per 2^20 points: E1 / E2 = 1 / 1: 1.367,
// 2 / 1: 1.349, 4 / 1: 1.328, 4 / 2: 1.308, 4 / 4: 1.383, 1 / 2: 1.368; H = 256: 3.67 -> 3.59; H = 64 is faster without)
#ifndef MS_PH_E1_UNROLL
#define MS_PH_E1_UNROLL (PH_H >= 128 ? 4 : 1)
#endif
#ifndef MS_PH_E2_UNROLL
#define MS_PH_E2_UNROLL (PH_H >= 128 ? 2 : 1)
#endif
constexpr int PH_E1_UNROLL = MS_PH_E1_UNROLL, PH_E2_UNROLL = MS_PH_E2_UNROLL;
#ifndef MS_PH_E2_DOT
#define MS_PH_E2_DOT 1
#endif
constexpr bool PH_E2_DOT = MS_PH_E2_DOT != 0;     // last layer: weighted sum of the sigmoids without forming them
#ifndef MS_PH_ROWS
#define MS_PH_ROWS 1
#endif
constexpr int PH_ROWS = MS_PH_ROWS;         // corner rows of the MIST gather per memory round trip in the prologue (1: 1.302 ms, 2: 1.308, 4: 1.459, 8: 1.573 -- spills)
constexpr int NT = H <= 128 ? 2 : 1; // tiles (of 128 points) a CTA works on at a time
constexpr bool STREAM = H > 128;     // layer-2 weights arrive in K slices
// Per-band weight image in global memory (and in each shared-memory ring slot).
// Operand images use the K-major no-swizzle core-matrix layout: [k/8][row][k%8] fp16,
// i.e. 8 rows x 16 B core matrices, SBO = 128 B between 8-row groups, LBO = 2048 B
// between 8-element K chunks (128 rows).
constexpr int W1_BYTES = 2 * H * 16;                   // K = 16
constexpr int W2_BYTES = (H / 8) * H * 16;             // K = H
constexpr int OFF_W1H = 0;
constexpr int OFF_W1L = OFF_W1H + W1_BYTES;
constexpr int OFF_W2H = OFF_W1L + W1_BYTES;
constexpr int OFF_W2L = OFF_W2H + W2_BYTES;
constexpr int OFF_CB2 = OFF_W2L + W2_BYTES;            // float[128]: -log2e * b2
constexpr int OFF_W3 = OFF_CB2 + 4 * H;                // float[H]
constexpr int OFF_B3 = OFF_W3 + 4 * H;                 // float[4] (b3 in [0])
constexpr int BAND_BYTES = OFF_B3 + 128;               // 74,880 B per band in global memory (multiple of 128)
constexpr int KSL = 32;                                // STREAM: K of one slice (two k-steps)
constexpr int NSL = H / KSL;                           //         slices per band
constexpr int SLICE_BYTES = (KSL / 8) * H * 16;        //         one slice of W2 hi (or lo): contiguous in the image
constexpr int RING_SLOTS = STREAM ? 3 : 2;
constexpr int RING_BYTES = STREAM ? 2 * SLICE_BYTES : OFF_CB2;   // what a ring slot holds (operand images only)
constexpr int CONST_BYTES = BAND_BYTES - OFF_CB2;      // cb2, w3, b3: resident in shared memory for all bands

constexpr int SM_W = 0;                                // ring slots
constexpr int SM_W1 = SM_W + RING_SLOTS * RING_BYTES;  // STREAM: [2][W1 hi, W1 lo] of the current / next band
constexpr int SM_X = SM_W1 + (STREAM ? 2 * 2 * W1_BYTES : 0);   // [2 pair buffers][2 tiles][hi, lo] X tile images
constexpr int SM_PT = SM_X + 8 * X_BYTES;              // [2 pair buffers][256 points] PointState
constexpr int SM_PART = SM_PT + 2 * 256 * 24;          // float[2 band parities][2 tiles][128]: B's partial dot products
constexpr int SM_BAR = SM_PART + 2048;                 // mbarriers
constexpr int SM_CONST = SM_BAR + 512;                 // [nb][CONST_BYTES] epilogue constants
constexpr int SM_FIXED = SM_CONST;
constexpr int MAX_NB = (SMEM_LIMIT - SM_FIXED) / CONST_BYTES;

// instruction descriptor: D = f32, A = B = f16, both K-major, N = 128, M = 128
constexpr uint32_t IDESC = make_idesc_f16(128, H);
constexpr int NCH = H / 16;            // 16-column chunks of an accumulator = layer-2 k-steps
constexpr int WLBO = H * 16;           // bytes between 8-element K chunks of a weight image (H rows of 16 B)

enum { BAR_WFULL = 0, BAR_WEMPTY = 2, BAR_XFULL = 4, BAR_D1 = 6, BAR_D2 = 8, BAR_A1 = 10 /* [tile][8 chunks] or [16 chunks] */,
       BAR_XEMPTY = 26, BAR_SFULL = 28 /* STREAM: [3] */, BAR_SEMPTY = 31, BAR_W1FULL = 34 /* [2] */, BAR_W1EMPTY = 36,
       BAR_N = 38 };
static_assert(BAR_N * 8 <= 512, "mbarrier area");

// FAST = MS_PREC_TC16: sigmoid(z) = 1/2 + tanh(z/2)/2 with the affine part folded into the next
// layer's packed weights, tanh.approx.f32 (one MUFU per activation), activations handed to layer 2 as
// single fp16 values (weights stay hi + lo).  Stated tolerance instead of fp32-class results.
template <bool FAST>
__global__ void __launch_bounds__(NTHREADS, 1)
phot_mlp_tc_kernel(TcArgs a, int64_t B, int npairs) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t s_tmem_base;
    // Pairs are handed out dynamically (a.sched[0] = next pair beyond the first wave): SMs differ by a few per cent in
    // speed and 4096 pairs are not a multiple of 148, so a static round-robin leaves the slowest CTA 5 % behind the
    // median.  The prologue warpgroup, one pair ahead, draws the index and publishes it here before it arrives on the
    // pair's XFULL barrier; every other role reads it after its own XFULL wait.  An index >= npairs ends the CTA.
    __shared__ int s_pair[4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef MS_PHOT_STAMPS
    if (threadIdx.x == 0) {
        unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
        g_cta[blockIdx.x * 8 + 0] = clock64(); g_cta[blockIdx.x * 8 + 2] = (long long)gt; g_cta[blockIdx.x * 8 + 4] = sm;
    }
#endif
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar0 = sbase + SM_BAR;
    auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar(BAR_WFULL + i), 1);
            mbar_init(bar(BAR_WEMPTY + i), NT);
            mbar_init(bar(BAR_XFULL + i), 4);                   // [pair buffer]: the 4 prologue warps
            mbar_init(bar(BAR_XEMPTY + i), NT * (1 + 4));       // per tile: its issuer (last layer 1 done) + its 4 A warps
            mbar_init(bar(BAR_W1FULL + i), 1);
            mbar_init(bar(BAR_W1EMPTY + i), 1);
            mbar_init(bar(BAR_D1 + i), 1);
            mbar_init(bar(BAR_D2 + i), 1);
            for (int j = 0; j < 8; ++j) mbar_init(bar(BAR_A1 + 8 * i + j), 4);
        }
        for (int i = 0; i < 3; ++i) { mbar_init(bar(BAR_SFULL + i), 1); mbar_init(bar(BAR_SEMPTY + i), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 16) {                                  // TMEM: all 512 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "n"(2 * NT * H) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    const int nb = a.nb;
    for (int i = threadIdx.x; i < nb * (CONST_BYTES / 16); i += NTHREADS) {      // epilogue constants of every band
        const int b = i / (CONST_BYTES / 16), o = i % (CONST_BYTES / 16);
        reinterpret_cast<uint4 *>(smem + SM_CONST + b * CONST_BYTES)[o] =
            reinterpret_cast<const uint4 *>(a.wimg + (size_t)b * BAND_BYTES + OFF_CB2)[o];
    }
    double *s_axes = reinterpret_cast<double *>(smem + SM_CONST + nb * CONST_BYTES);     // grid axes (fused form)
    double *s_stage = s_axes + ((a.ne + a.nm + a.nf + a.na + 1) & ~1);                   // [128][ld <= 8] rows of theta
    if (a.fused)
        for (int i = threadIdx.x; i < a.ne + a.nm + a.nf + a.na; i += NTHREADS) s_axes[i] = a.axes[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem_base;

    if (warp == 18) {
        // ===== weight producer =====
        if (lane == 0) {
            int g = 0, gs = 0;
            for (int it = 0;; ++it) {
                mbar_wait(bar(BAR_XFULL + (it & 1)), (it >> 1) & 1);
                if (*(volatile int *)&s_pair[it & 3] >= npairs) break;
                if constexpr (!STREAM) {
                    for (int b = 0; b < nb; ++b, ++g) {
                        const int buf = g & 1;
                        mbar_wait(bar(BAR_WEMPTY + buf), ((g >> 1) & 1) ^ 1);
                        mbar_expect_tx(bar(BAR_WFULL + buf), RING_BYTES);
                        bulk_g2s(sbase + SM_W + buf * RING_BYTES, a.wimg + (size_t)b * BAND_BYTES, RING_BYTES,
                                 bar(BAR_WFULL + buf));
                    }
                } else {
                    // per band: the layer-1 image (hi + lo, contiguous) into its own double buffer, then the NSL K slices
                    // of the layer-2 image (hi and lo halves are NSL * SLICE_BYTES apart) through the ring
                    for (int b = 0; b < nb; ++b, ++g) {
                        const unsigned char *img = a.wimg + (size_t)b * BAND_BYTES;
                        const int b1 = g & 1;
                        mbar_wait(bar(BAR_W1EMPTY + b1), ((g >> 1) & 1) ^ 1);
                        mbar_expect_tx(bar(BAR_W1FULL + b1), 2 * W1_BYTES);
                        bulk_g2s(sbase + SM_W1 + b1 * 2 * W1_BYTES, img + OFF_W1H, 2 * W1_BYTES, bar(BAR_W1FULL + b1));
                        for (int sl = 0; sl < NSL; ++sl, ++gs) {
                            const int slot = gs % RING_SLOTS;
                            mbar_wait(bar(BAR_SEMPTY + slot), ((gs / RING_SLOTS) & 1) ^ 1);
                            mbar_expect_tx(bar(BAR_SFULL + slot), 2 * SLICE_BYTES);
                            const uint32_t dst = sbase + SM_W + slot * RING_BYTES;
                            bulk_g2s(dst, img + OFF_W2H + (size_t)sl * SLICE_BYTES, SLICE_BYTES, bar(BAR_SFULL + slot));
                            bulk_g2s(dst + SLICE_BYTES, img + OFF_W2L + (size_t)sl * SLICE_BYTES, SLICE_BYTES, bar(BAR_SFULL + slot));
                        }
                    }
                }
            }
        }
    } else if (warp == 16 || warp == 17) {
        // ===== MMA issuers: warp 16 + T drives tile T, blocking on that tile's barriers in program order =====
        // per band: layer 1 (needs the band's weights, and the X tile on band 0), then one layer-2 k-step
        // per A1 chunk as soon as the epilogue has written that chunk back to TMEM.  All descriptors are
        // built once; a k-step advances the start-address field.
        // The whole warp runs the loop converged and one elected lane issues: every operand is
        // warp-uniform, so the descriptors live in uniform registers and an MMA is a single instruction.
        if constexpr (!STREAM) {
            const int T = __shfl_sync(0xffffffffu, warp, 0) - 16;
            const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
            const uint32_t xb = sbase + SM_X + T * 2 * X_BYTES;          // pair buffer 1 is 4 * X_BYTES further
            const uint64_t xh0 = make_desc(xb, 2048, 128), xl0 = make_desc(xb + X_BYTES, 2048, 128);
            const uint32_t d1 = tm + T * 2 * H, d2 = d1 + H;
            uint64_t w1h[2], w1l[2], w2h[2], w2l[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const uint32_t wb = sbase + SM_W + i * RING_BYTES;
                w1h[i] = make_desc(wb + OFF_W1H, WLBO, 128); w1l[i] = make_desc(wb + OFF_W1L, WLBO, 128);
                w2h[i] = make_desc(wb + OFF_W2H, WLBO, 128); w2l[i] = make_desc(wb + OFF_W2L, WLBO, 128);
            }
            int g = 0;
            for (int it = 0;; ++it) {
                mbar_wait(bar(BAR_XFULL + (it & 1)), (it >> 1) & 1);
                if (*(volatile int *)&s_pair[it & 3] >= npairs) break;
                for (int b = 0; b < nb; ++b, ++g) {
                    const int buf = g & 1;
                    mbar_wait(bar(BAR_WFULL + buf), (g >> 1) & 1);
                    const uint64_t xh = xh0 + (uint64_t)((it & 1) * (4 * X_BYTES / 16));
                    const uint64_t xl = xl0 + (uint64_t)((it & 1) * (4 * X_BYTES / 16));
                    tc_fence_after();
                    const uint64_t bh1 = buf ? w1h[1] : w1h[0], bl1 = buf ? w1l[1] : w1l[0];
                    const uint64_t bh2 = buf ? w2h[1] : w2h[0], bl2 = buf ? w2l[1] : w2l[0];
                    if (elect_one()) {
                        mma_ss(d1, xh, bh1, 0, IDESC);
                        mma_ss(d1, xl, bh1, 1, IDESC);
                        mma_ss(d1, xh, bl1, 1, IDESC);
                        tc_commit(bar(BAR_D1 + T));
                        if (b == nb - 1) tc_commit(bar(BAR_XEMPTY + (it & 1)));  // last read of this pair's X tile
                    }
                    __syncwarp();
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {                      // k-step c: two K chunks further
                        mbar_wait(bar(BAR_A1 + 8 * T + c), g & 1);
                        tc_fence_after();
                        const uint64_t bh = bh2 + (uint64_t)(c * 2 * H), bl = bl2 + (uint64_t)(c * 2 * H);
                        const uint32_t ah = d1 + 16 * c;
                        if (elect_one()) {
                            mma_ts(d2, ah, bh, c > 0, IDESC);
                            if constexpr (!FAST) mma_ts(d2, ah + 8, bh, 1, IDESC);
                            mma_ts(d2, ah, bl, 1, IDESC);
                            if (c == NCH - 1) {
                                tc_commit(bar(BAR_D2 + T));
                                tc_commit(bar(BAR_WEMPTY + buf));        // this tile is done with the ring slot
                            }
                        }
                        __syncwarp();
                    }
                }
            }
        } else if (warp == 16) {
            // one tile: layer 1 from the band's own W1 buffer, layer 2 k-step by k-step as the A1 chunks land and the
            // K slices of W2 arrive (a slice = two k-steps; its ring slot is released by the commit behind the second)
            const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
            const uint32_t xb = sbase + SM_X;
            const uint64_t xh0 = make_desc(xb, 2048, 128), xl0 = make_desc(xb + X_BYTES, 2048, 128);
            const uint32_t d1 = tm, d2 = tm + H;
            int g = 0, gs = 0;
            for (int it = 0;; ++it) {
                mbar_wait(bar(BAR_XFULL + (it & 1)), (it >> 1) & 1);
                if (*(volatile int *)&s_pair[it & 3] >= npairs) break;
                const uint64_t xh = xh0 + (uint64_t)((it & 1) * (4 * X_BYTES / 16));
                const uint64_t xl = xl0 + (uint64_t)((it & 1) * (4 * X_BYTES / 16));
                for (int b = 0; b < nb; ++b, ++g) {
                    const int b1 = g & 1;
                    mbar_wait(bar(BAR_W1FULL + b1), (g >> 1) & 1);
                    tc_fence_after();
                    const uint32_t w1b = sbase + SM_W1 + b1 * 2 * W1_BYTES;
                    const uint64_t bh1 = make_desc(w1b, WLBO, 128), bl1 = make_desc(w1b + W1_BYTES, WLBO, 128);
                    if (elect_one()) {
                        mma_ss(d1, xh, bh1, 0, IDESC);
                        mma_ss(d1, xl, bh1, 1, IDESC);
                        mma_ss(d1, xh, bl1, 1, IDESC);
                        tc_commit(bar(BAR_D1));
                        tc_commit(bar(BAR_W1EMPTY + b1));
                        if (b == nb - 1) tc_commit(bar(BAR_XEMPTY + (it & 1)));  // last read of this tile's X image
                    }
                    __syncwarp();
#pragma unroll 1
                    for (int sl = 0; sl < NSL; ++sl, ++gs) {
                        const int slot = gs % RING_SLOTS;
                        mbar_wait(bar(BAR_SFULL + slot), (gs / RING_SLOTS) & 1);
                        const uint32_t wb = sbase + SM_W + slot * RING_BYTES;
                        const uint64_t bh0 = make_desc(wb, WLBO, 128), bl0 = make_desc(wb + SLICE_BYTES, WLBO, 128);
#pragma unroll
                        for (int h = 0; h < KSL / 16; ++h) {
                            const int c = sl * (KSL / 16) + h;                   // k-step = A1 chunk
                            mbar_wait(bar(BAR_A1 + c), g & 1);
                            tc_fence_after();
                            const uint64_t bh = bh0 + (uint64_t)(h * 2 * H), bl = bl0 + (uint64_t)(h * 2 * H);
                            const uint32_t ah = d1 + 16 * c;
                            if (elect_one()) {
                                mma_ts(d2, ah, bh, c > 0, IDESC);
                                if constexpr (!FAST) mma_ts(d2, ah + 8, bh, 1, IDESC);
                                mma_ts(d2, ah, bl, 1, IDESC);
                                if (h == KSL / 16 - 1) tc_commit(bar(BAR_SEMPTY + slot));
                                if (c == NCH - 1) tc_commit(bar(BAR_D2));
                            }
                            __syncwarp();
                        }
                    }
                }
            }
        }
    } else if (warp >= 19) {
        // ===== prologue warpgroup: network inputs and distance modulus of the next pair's 256 points =====
        // (GenMod.genphot, genmod.py:97-142), one pair ahead of the epilogue so the f64 logarithms and the
        // strided loads of the parameter block never sit on the critical path.
        const int tid = threadIdx.x - 19 * 32;
        for (int it = 0;; ++it) {
            const int xbuf = it & 1;
            if (tid == 0) STAMP(it, 0);
            if (tid == 0) {
                const int mine = it == 0 ? (int)blockIdx.x : (int)gridDim.x + atomicAdd(a.sched, 1);
                s_pair[it & 3] = mine;
                if (a.rows_ready && mine < npairs) {
                    // streaming host entry: this pair's rows of theta are still on their way over PCIe (copy engine
                    // writes, then the counter: both through L2; the rows are first touched after this load)
                    int64_t need = ((int64_t)mine + 1) * (NT * TILE);
                    if (need > B) need = B;
                    while ((int64_t)*reinterpret_cast<const volatile int *>(a.rows_ready) < need) __nanosleep(256);
                    __threadfence();
                }
            }
            mbar_wait(bar(BAR_XEMPTY + xbuf), ((it >> 1) & 1) ^ 1);
            asm volatile("bar.sync 3, 128;" ::: "memory");              // the four prologue warps: index visible
            if (tid == 0) STAMP(it, 1);
            const int pr = s_pair[it & 3];
            if (pr >= npairs) {                                           // publish the end marker and stop
                if (lane == 0) mbar_arrive(bar(BAR_XFULL + xbuf));
                break;
            }
#pragma unroll 1
            for (int T = 0; T < NT; ++T) {
                const int64_t r0 = (int64_t)pr * (NT * TILE) + T * TILE, p = r0 + tid;
                if (a.stage) {
                    // the tile's rows of theta through shared memory, fetched with coalesced 16-byte loads: theta may
                    // be pinned HOST memory (ms_lnlike_batch_host reads it in place: no upload, every byte crosses PCIe
                    // once), where the per-thread 8-byte loads of theta_get would fetch every line six times
                    asm volatile("bar.sync 3, 128;" ::: "memory");      // the previous tile's rows are no longer read
                    int64_t nrow = B - r0;
                    nrow = nrow > TILE ? TILE : (nrow < 0 ? 0 : nrow);
                    const int nd8 = (int)nrow * a.tv.ld;
                    const double *src = a.tv.theta + r0 * a.tv.ld;
                    for (int i = tid; i < (nd8 >> 1); i += 128)
                        reinterpret_cast<double2 *>(s_stage)[i] = reinterpret_cast<const double2 *>(src)[i];
                    if ((nd8 & 1) && tid == 0) s_stage[nd8 - 1] = src[nd8 - 1];
                    asm volatile("bar.sync 3, 128;" ::: "memory");
                }
                auto tget = [&](int pid) {
                    const int c = a.tv.col[pid];
                    if (c < 0) return a.tv.fixed[pid];
                    return a.stage ? s_stage[tid * a.tv.ld + c] : a.tv.theta[p * (int64_t)a.tv.ld + c];
                };
                PointState ps{0.0, 1.0, 0, 0};
                float xs[8];
#pragma unroll
                for (int d = 0; d < 8; ++d) xs[d] = 0.f;
                if (p < B) {
                    double teff, logt, logg, feh, afe, logr, dist, av;
                    if (a.fused) {
                        // kernel (1) for this point: brackets, corner gather, renormalised sum, then the
                        // parameter block of likelihood.lnlikefn (likelihood.py:121-141) in registers
                        const double *g[4] = {s_axes, s_axes + a.ne, s_axes + a.ne + a.nm, s_axes + a.ne + a.nm + a.nf};
                        const int n[4] = {a.ne, a.nm, a.nf, a.na};
                        double x[4];
                        x[0] = tget(MS_P_EEP);
                        x[1] = tget(MS_P_INIT_MASS);
                        x[2] = tget(MS_P_INIT_FEH);
                        x[3] = tget(MS_P_INIT_AFE);
                        int idx[4];
                        float acc[MS_NCOL], wsum;
                        double accd[2], wsumd;
                        const bool inside = mist::interp_point_f32<FAST ? 1 : PH_ROWS>(g, n, a.table, x, idx, acc, wsum, accd, wsumd);   // rows per round trip: measured
                        const bool okp = inside && (wsum > 0.f);
                        const float inv = okp ? __fdividef(1.f, wsum) : 0.f;
                        const double inf = ms_inf();
                        // Teff = 10**log(Teff) (likelihood.py:121) is only needed as a network input here;
                        // genphot's log10(Teff) (genmod.py:111) is the interpolated value itself
                        const double invd = okp ? 1.0 / wsumd : 0.0;
                        logt = accd[1] * invd;
                        // 10**logt from the f64 sum: f32 exponential of the rounded value times the first-order term of
                        // the remainder (an f32 logt alone would cost 9e-7 of Teff: ln 10 times one ulp of 3.7)
                        const float lt_hi = (float)logt, lt_lo = (float)(logt - (double)lt_hi);
                        teff = okp ? (double)(exp10f(lt_hi) * fmaf(2.302585093f, lt_lo, 1.f)) : inf;
                        logg = okp ? (double)(acc[7] * inv) : inf;
                        feh = okp ? (a.mistdif ? (double)(acc[5] * inv) : x[2]) : inf;     // :127-132
                        afe = okp ? (a.mistdif ? (double)(acc[6] * inv) : x[3]) : inf;
                        logr = okp ? accd[0] * invd : inf;
                        ps.agewgt = okp ? (double)(acc[8] * inv) : inf;
                        dist = tget(MS_P_DIST);
                        av = tget(MS_P_AV);
                        if (a.derived) {
                            double *o = a.derived + p * MS_NDERIVED;
                            o[0] = x[0]; o[1] = x[1]; o[2] = x[2]; o[3] = x[3];
#pragma unroll
                            for (int c = 0; c < MS_NCOL; ++c)
                                o[4 + c] = okp ? (c == 2 ? logr : (c == 4 ? logt : (double)(acc[c] * inv))) : inf;
                        }
                    } else {
                        const double *q = a.pars + p * MS_NPARS;
                        teff = q[PV_TEFF]; logg = q[PV_LOGG]; feh = q[PV_FEH]; afe = q[PV_AFE];
                        logr = q[PV_LOGR]; dist = q[PV_DIST]; av = q[PV_AV];
                        ps.agewgt = q[PV_AGEWGT];
                        logt = log10(teff);                                  // genmod.py:111-113
                    }
                    if (teff != ms_inf()) {
                        ps.ok = 1;
                        const double logl = 2.0 * logr + 4.0 * (logt - log10(5770.0));
                        ps.mag0 = -2.5 * logl + 4.74 + 5.0 * log10(dist) - 5.0;
                        const double xin[6] = {teff, logg, feh, afe, av, 3.1};    // 10**logt of the reference's sed()
#pragma unroll
                        for (int d = 0; d < 6; ++d)
                            if (d < a.d_in) xs[d] = (float)((xin[d] - a.xmin[d]) * a.xscale[d] - 0.5);
                    }
                    // a slot outside the loaded table is marked -1: the row comes out NaN, nothing is read out of bounds
                    const int sl = a.star_slot ? a.star_slot[p] : a.slot0;
                    ps.slot = (unsigned)sl < (unsigned)a.nslots ? sl : -1;
                }
                xs[6] = 1.0f;                                                // bias column
                // X tile row in the core-matrix layout: chunk c (8 k) at c*2048 + r*16
                unsigned char *xrow_h = smem + SM_X + (xbuf * 2 + T) * 2 * X_BYTES, *xrow_l = xrow_h + X_BYTES;
                __half2 hh[8], ll[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const __half2 h2 = __floats2half2_rn(xs[2 * i], xs[2 * i + 1]);
                    const float2 back = __half22float2(h2);
                    hh[i] = h2;
                    ll[i] = __floats2half2_rn(xs[2 * i] - back.x, xs[2 * i + 1] - back.y);
                    hh[4 + i] = __floats2half2_rn(0.f, 0.f);
                    ll[4 + i] = hh[4 + i];
                }
                *reinterpret_cast<uint4 *>(xrow_h + tid * 16) = *reinterpret_cast<uint4 *>(&hh[0]);
                *reinterpret_cast<uint4 *>(xrow_h + 2048 + tid * 16) = *reinterpret_cast<uint4 *>(&hh[4]);
                *reinterpret_cast<uint4 *>(xrow_l + tid * 16) = *reinterpret_cast<uint4 *>(&ll[0]);
                *reinterpret_cast<uint4 *>(xrow_l + 2048 + tid * 16) = *reinterpret_cast<uint4 *>(&ll[4]);
                reinterpret_cast<PointState *>(smem + SM_PT)[(xbuf * 2 + T) * TILE + tid] = ps;
            }
            fence_proxy_async();
            __syncwarp();
            if (tid == 0) STAMP(it, 2);
            if (lane == 0) mbar_arrive(bar(BAR_XFULL + xbuf));
        }
    } else {
        // ===== epilogue warpgroups: tile T, row r, chunk parity `half` =====
        const int half = warp >> 3;                    // 0: warpgroup A (even chunks, prologue, chi-square), 1: B
        const int T = (warp >> 2) & 1;
        const int r = threadIdx.x & 127;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t t_reg1 = tmem + lane_base + T * 2 * H, t_d2 = t_reg1 + H;
        float *part = reinterpret_cast<float *>(smem + SM_PART);
        int g = 0;
        // skew: tile 1 starts when tile 0 has finished its first layer-1 epilogue, so that one tile's waits for
        // the tensor pipe fall into the other tile's activation phases instead of coinciding with its own
        if (NT == 2 && T == 1) mbar_wait(bar(BAR_A1 + NCH - 1), 0);
        for (int it = 0; T < NT; ++it) {               // one-tile form (NT = 1): the warpgroups of tile 1 have nothing to do
            mbar_wait(bar(BAR_XFULL + (it & 1)), (it >> 1) & 1);
            const int pr = *(volatile int *)&s_pair[it & 3];
            if (pr >= npairs) break;
            const int64_t p = (int64_t)pr * (NT * TILE) + T * TILE + r;
            const bool live = half == 0 && p < B;
            bool ok = false;
            double mag0 = 0.0, agewgt = 1.0;
            int slot = 0;
            if (threadIdx.x == 0) STAMP(it, 3);
            if (half == 0) {                           // point state from the prologue warpgroup
                if (threadIdx.x == 0) STAMP(it, 4);
                const PointState ps = reinterpret_cast<const PointState *>(smem + SM_PT)[((it & 1) * 2 + T) * TILE + r];
                ok = ps.ok != 0; mag0 = ps.mag0; agewgt = ps.agewgt; slot = ps.slot;
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(BAR_XEMPTY + (it & 1)));
            }
            double chi2 = 0.0;

            for (int b = 0; b < nb; ++b, ++g) {
                const unsigned char *cb = smem + SM_CONST + b * CONST_BYTES;
                // this band's observation, requested now and used after the two epilogues
                double obs_o = 0.0, obs_e2 = 1.0;
                if (half == 0 && (a.chi2 || a.lnL) && slot >= 0) {
                    const size_t at = (size_t)slot * a.nb_total + a.b0 + b;
                    obs_o = a.obs_mag[at]; obs_e2 = a.obs_err2[at];
                }
                // ---- layer-1 epilogue: activation, fp16 (split), written back in place as the TS A operand
                mbar_wait(bar(BAR_D1 + T), g & 1);
                tc_fence_after();
#pragma unroll PH_E1_UNROLL
                for (int i4 = 0; i4 < NCH / 2; ++i4) {
                    const int c = 2 * i4 + half;
                    uint32_t v[16], o[16];
                    TMEM_LD16(t_reg1 + 16 * c, v);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if constexpr (FAST) {
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            o[i] = pack_half2(tanh_approx(__uint_as_float(v[2 * i])), tanh_approx(__uint_as_float(v[2 * i + 1])));
                        TMEM_ST8(t_reg1 + 16 * c, o);
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            float a0, a1, a2, a3;
                            sigmoid_quad(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                         __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]), a0, a1, a2, a3);
                            const __half2 h01 = __floats2half2_rn(a0, a1), h23 = __floats2half2_rn(a2, a3);
                            float e0, e1, e2, e3;
                            sub_half2(a0, a1, h01, e0, e1);
                            sub_half2(a2, a3, h23, e2, e3);
                            const __half2 l01 = __floats2half2_rn(e0, e1), l23 = __floats2half2_rn(e2, e3);
                            o[2 * i] = *reinterpret_cast<const uint32_t *>(&h01);
                            o[2 * i + 1] = *reinterpret_cast<const uint32_t *>(&h23);
                            o[8 + 2 * i] = *reinterpret_cast<const uint32_t *>(&l01);
                            o[8 + 2 * i + 1] = *reinterpret_cast<const uint32_t *>(&l23);
                        }
                        TMEM_ST16(t_reg1 + 16 * c, o);
                    }
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar(BAR_A1 + 8 * T + c));   // chunk c usable as TS A operand
                }

                // ---- layer-2 epilogue: activation, dot with w3
                const float4 *cb2 = reinterpret_cast<const float4 *>(cb);
                const float4 *w3 = reinterpret_cast<const float4 *>(cb + (OFF_W3 - OFF_CB2));
                float bc0 = 0.f, bc1 = 0.f;
                mbar_wait(bar(BAR_D2 + T), g & 1);
                tc_fence_after();
#pragma unroll PH_E2_UNROLL
                for (int i4 = 0; i4 < NCH / 2; ++i4) {
                    const int c = 2 * i4 + half;
                    uint32_t v[16];
                    TMEM_LD16(t_d2 + 16 * c, v);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 cc = cb2[4 * c + i], w = w3[4 * c + i];
                        if constexpr (!FAST && PH_E2_DOT) {
                            const float sdot = sigmoid_dot_quad(__uint_as_float(v[4 * i]) + cc.x, __uint_as_float(v[4 * i + 1]) + cc.y,
                                                               __uint_as_float(v[4 * i + 2]) + cc.z, __uint_as_float(v[4 * i + 3]) + cc.w, w);
                            if (i & 1) bc1 += sdot; else bc0 += sdot;
                            continue;
                        }
                        float a0, a1, a2, a3;
                        if constexpr (FAST) {
                            a0 = tanh_approx(__uint_as_float(v[4 * i]) + cc.x);
                            a1 = tanh_approx(__uint_as_float(v[4 * i + 1]) + cc.y);
                            a2 = tanh_approx(__uint_as_float(v[4 * i + 2]) + cc.z);
                            a3 = tanh_approx(__uint_as_float(v[4 * i + 3]) + cc.w);
                        } else {
                            sigmoid_quad(__uint_as_float(v[4 * i]) + cc.x, __uint_as_float(v[4 * i + 1]) + cc.y,
                                         __uint_as_float(v[4 * i + 2]) + cc.z, __uint_as_float(v[4 * i + 3]) + cc.w,
                                         a0, a1, a2, a3);
                        }
                        bc0 = fmaf(w.x, a0, bc0); bc1 = fmaf(w.y, a1, bc1);
                        bc0 = fmaf(w.z, a2, bc0); bc1 = fmaf(w.w, a3, bc1);
                    }
                }
                const float b3 = *reinterpret_cast<const float *>(cb + (OFF_B3 - OFF_CB2));
                float *pslot = part + ((g & 1) * 2 + T) * TILE + r;
                if (half == 1) {                                         // B: hand the odd-chunk half to A
                    *pslot = bc0 + bc1;
                    __threadfence_block();
                    asm volatile("bar.arrive %0, 256;" ::"r"(1 + T) : "memory");
                    continue;
                }
                asm volatile("bar.sync %0, 256;" ::"r"(1 + T) : "memory");
                const float bc = (bc0 + bc1) + *pslot + b3;
                if (ok) {
                    const double mag = mag0 - (double)bc;
                    if (a.mags) a.mags[p * a.nb_total + a.b0 + b] = mag;
                    if (a.chi2 || a.lnL) {
                        const double rr = mag - obs_o;
                        const double t = (rr * rr) / obs_e2;
                        if (t == t) chi2 += t;                           // nansum (likelihood.py:207)
                    }
                } else if (live && a.mags) {
                    a.mags[p * a.nb_total + a.b0 + b] = ms_nan();
                }
            }
            if (threadIdx.x == 0) STAMP(it, 5);
            if (half == 1) continue;
            if (live && a.lnL) {
                double l = -ms_inf();
                if (ok) {
                    l = -0.5 * ((a.chi2 ? a.chi2[p] : 0.0) + chi2);
                    if (a.ageweight) l += log(agewgt);
                    if (slot < 0) l = ms_nan();
                }
                a.lnL[p] = l;
                for (int rk = 0; rk < a.npeers; ++rk) a.peers[rk][a.peer_off + p] = l;      // posted writes to every rank
            } else if (ok && a.chi2) {
                a.chi2[p] += slot < 0 ? ms_nan() : chi2;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    // the last CTA to leave re-arms the pair counter for the next launch (every CTA has drawn its end marker by then)
    if (threadIdx.x == 0 && atomicAdd(a.sched + 1, 1) == (int)gridDim.x - 1) { a.sched[0] = 0; a.sched[1] = 0; __threadfence(); }
#ifdef MS_PHOT_STAMPS
    if (threadIdx.x == 0) {
        unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        g_cta[blockIdx.x * 8 + 1] = clock64(); g_cta[blockIdx.x * 8 + 3] = (long long)gt;
    }
#endif
    if (warp == 16) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(2 * NT * H) : "memory");
    }
}

// fp16 hi / lo split of a double, returned as raw bits
void split_f16(double v, uint16_t &hi, uint16_t &lo) {
    const __half h = __float2half_rn((float)v);
    const double back = (double)__half2float(h);
    const __half l = __float2half_rn((float)(v - back));
    hi = *reinterpret_cast<const uint16_t *>(&h);
    lo = *reinterpret_cast<const uint16_t *>(&l);
}


// packs the weights of every band into the per-band images the kernel streams (see the layout above)
int pack_images(ms_ctx *ctx, std::vector<unsigned char> &img) {
    PhotNetDev &P = ctx->phot;
    const int nb = P.nb, D = P.d_in;
    std::vector<float> w1((size_t)nb * H * D), b1((size_t)nb * H), w2((size_t)nb * H * H), b2((size_t)nb * H),
        w3((size_t)nb * H), b3(nb);
    MS_CUDA(cudaMemcpy(w1.data(), P.w1, w1.size() * 4, cudaMemcpyDeviceToHost));
    MS_CUDA(cudaMemcpy(b1.data(), P.b1, b1.size() * 4, cudaMemcpyDeviceToHost));
    MS_CUDA(cudaMemcpy(w2.data(), P.w2, w2.size() * 4, cudaMemcpyDeviceToHost));
    MS_CUDA(cudaMemcpy(b2.data(), P.b2, b2.size() * 4, cudaMemcpyDeviceToHost));
    MS_CUDA(cudaMemcpy(w3.data(), P.w3, w3.size() * 4, cudaMemcpyDeviceToHost));
    MS_CUDA(cudaMemcpy(b3.data(), P.b3, b3.size() * 4, cudaMemcpyDeviceToHost));
    const bool fast = ctx->precision == MS_PREC_TC16;
    // precise: accumulators hold -log2(e) z.  fast: accumulators hold z / 2 with sigmoid = (1 + tanh) / 2
    // folded forward: W2 t-form = W2 / 4, bias2 = b2 / 2 + rowsum(W2) / 4, w3 t-form = w3 / 2, b3 += sum(w3) / 2.
    const double NL2E = fast ? 0.5 : -1.4426950408889634;         // -log2(e), folded into the weights
    img.assign((size_t)nb * BAND_BYTES, 0);
    for (int b = 0; b < nb; ++b) {
        unsigned char *base = img.data() + (size_t)b * BAND_BYTES;
        uint16_t *w1h = reinterpret_cast<uint16_t *>(base + OFF_W1H), *w1l = reinterpret_cast<uint16_t *>(base + OFF_W1L);
        uint16_t *w2h = reinterpret_cast<uint16_t *>(base + OFF_W2H), *w2l = reinterpret_cast<uint16_t *>(base + OFF_W2L);
        float *cb2 = reinterpret_cast<float *>(base + OFF_CB2), *w3o = reinterpret_cast<float *>(base + OFF_W3);
        float *b3o = reinterpret_cast<float *>(base + OFF_B3);
        for (int n = 0; n < H; ++n) {
            for (int k = 0; k < KX; ++k) {
                double v = 0.0;
                if (k < D) v = NL2E * (double)w1[((size_t)b * H + n) * D + k];
                else if (k == 6) v = NL2E * (double)b1[(size_t)b * H + n];
                if (std::fabs(v) > 6.0e4) { ms_set_error("layer-1 weight exceeds the fp16 range"); return MS_EINVAL; }
                const size_t at = (size_t)(k / 8) * (H * 8) + (size_t)n * 8 + (k % 8);   // in fp16 elements
                split_f16(v, w1h[at], w1l[at]);
            }
            for (int k = 0; k < H; ++k) {
                const double v = (fast ? 0.25 : NL2E) * (double)w2[((size_t)b * H + n) * H + k];
                if (std::fabs(v) > 6.0e4) { ms_set_error("layer-2 weight exceeds the fp16 range"); return MS_EINVAL; }
                const size_t at = (size_t)(k / 8) * (H * 8) + (size_t)n * 8 + (k % 8);
                split_f16(v, w2h[at], w2l[at]);
            }
            if (fast) {
                double rs = 0.0;
                for (int k = 0; k < H; ++k) rs += (double)w2[((size_t)b * H + n) * H + k];
                cb2[n] = (float)(0.5 * (double)b2[(size_t)b * H + n] + 0.25 * rs);
                w3o[n] = 0.5f * w3[(size_t)b * H + n];
            } else {
                cb2[n] = (float)(NL2E * (double)b2[(size_t)b * H + n]);
                w3o[n] = w3[(size_t)b * H + n];
            }
        }
        double s3 = 0.0;
        for (int n = 0; n < H; ++n) s3 += (double)w3[(size_t)b * H + n];
        b3o[0] = fast ? (float)((double)b3[b] + 0.5 * s3) : b3[b];
    }
    return MS_OK;
}

const TcVariant variant = {H, BAND_BYTES, CONST_BYTES, SM_FIXED, MAX_NB, NT * TILE, pack_images,
                           phot_mlp_tc_kernel<false>, phot_mlp_tc_kernel<true>};
