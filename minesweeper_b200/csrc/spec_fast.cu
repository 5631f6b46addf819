// Kernel (3), f32 fast path: rotational + instrumental broadening, resampling onto the observed
// pixels, blaze polynomial and the spectral chi-square for the fp32-class modes (MS_PREC_F32 / TC).
// Same contract and reference lines as spec_smooth.cu (smoothing.py:92-114, 131-134, 202-264, 517-598;
// fitutils.py:11-20; likelihood.py:188-195); that file keeps the generic kernel (f64, and the f32 rows
// this kernel declines).  What is different here, and why it still reproduces the reference's numbers:
//
//  * Rotational stage (smooth_vsini_fft): the reference's kernel is DEFINED by its FFT construction (analytic
//    transfer function sampled on the rfft frequencies of the power-of-two log grid: band-limited, 1/j^2 tails),
//    so it stays an FFT -- but the real transform of n1 samples is one in-place complex transform of M = n1/2
//    points in (padded) shared memory: radix-8 decimation-in-frequency passes forward (digit-reversed spectrum),
//    the mirrored decimation-in-time passes back, and between them ONE pass that does the last forward radix-2
//    stage, the real-spectrum split, the transfer function, the merge and the first inverse radix-2 stage on the
//    index quadruples {k, M/2-k, M/2+k, M-k} at their permuted positions.  The first pass reads the native
//    spectrum through the linear resampling (smoothing.py:587-597) instead of from a resampled copy.
//  * Instrumental stage (smooth_vel_fft, a Gaussian of sigma = sqrt(sigma_out^2 - sigma_in^2)): the reference's
//    transfer function exp(-2 pi^2 sigma^2 f^2) is the exact DFT of the periodised sampled Gaussian whenever it has
//    decayed at the Nyquist frequency, i.e. the circular FFT convolution EQUALS the direct sum with taps
//    exp(-t^2 / 2 s^2) / (s sqrt(2 pi)), s = sigma / dv in pixels of the resampled grid (Poisson summation;
//    error exp(-pi^2 s^2 / 2) < 3e-9 for s >= 2; SURVEY.md Appendix A.11 measured 4e-16).  The taps are cut at
//    5.5 s (tail 4e-8) and the sum is evaluated only at the two grid points that bracket each observed pixel.
//    Rows with s < 2 or more than MAXTAP taps are not evaluated here: they are appended to a fallback list that
//    the generic FFT kernel then processes.
//  * Every resampling position is affine in the pixel index (log-uniform grids, Doppler shift = translation in
//    ln lambda) and is computed in f64 from the index: no map tables are read.
#include "spec_args.cuh"

namespace {

#ifndef MS_FAST_FT
#define MS_FAST_FT 512
#endif
#ifndef MS_IP_UNROLL
#define MS_IP_UNROLL 2        // butterflies in flight per thread in a transform stage (2 = all of them at 512 threads: 11.34 -> 10.67 ms)
#endif
constexpr int IP_UNROLL = MS_IP_UNROLL;
#ifndef MS_FIR_UNROLL
#define MS_FIR_UNROLL 2
#endif
#ifndef MS_BACK_UNROLL
#define MS_BACK_UNROLL 4           // resampling back onto the native pixels: 10.89 -> 10.72 ms
#endif
constexpr int FIR_UNROLL = MS_FIR_UNROLL, BACK_UNROLL = MS_BACK_UNROLL;
#ifndef MS_QUAD_UNROLL
#define MS_QUAD_UNROLL 1
#endif
constexpr int QUAD_UNROLL = MS_QUAD_UNROLL;
constexpr int FT = MS_FAST_FT;          // threads per CTA (two CTAs per SM at 64 registers)
constexpr int MAXM = 8192;              // complex transform length limit (n1 <= 16384)
constexpr int MAXTAP = 64;              // half-width limit of the Gaussian FIR (5.5 sigma: sigma <= 11.6 pixels)
constexpr int WU = 2 * MAXTAP + 8;      // longest aligned FIR window (multiple of 4)
constexpr int WSTRIDE = WU + 4;         // row stride of the weight table: 4 mod 32 floats -> the 4 rows hit distinct banks

__device__ __forceinline__ int padi(int i) { return i + (i >> 4); }    // bank-conflict padding of the complex buffer

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return {a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return {a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) { return {a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y}; }   // a conj(b)
template <bool INV> __device__ __forceinline__ float2 mul_mi(float2 a) {       // a * (-i) forward, a * (+i) inverse
    return INV ? float2{-a.y, a.x} : float2{a.y, -a.x};
}

// ---- small DFTs in registers, natural order in and out ---------------------------------------
template <bool INV> __device__ __forceinline__ void dft2(float2 &a, float2 &b) {
    const float2 s = cadd(a, b), d = csub(a, b);
    a = s; b = d;
}
template <bool INV> __device__ __forceinline__ void dft4(float2 &x0, float2 &x1, float2 &x2, float2 &x3) {
    const float2 a = cadd(x0, x2), b = csub(x0, x2), c = cadd(x1, x3), d = mul_mi<INV>(csub(x1, x3));
    x0 = cadd(a, c); x1 = cadd(b, d); x2 = csub(a, c); x3 = csub(b, d);
}
// multiply by W8^m (forward: exp(-2 pi i m / 8)), m = 1, 2, 3
template <bool INV, int m> __device__ __forceinline__ float2 mul_w8(float2 a) {
    const float r = 0.70710678118654752440f;
    if (m == 1) return INV ? float2{r * (a.x - a.y), r * (a.x + a.y)} : float2{r * (a.x + a.y), r * (a.y - a.x)};
    if (m == 2) return mul_mi<INV>(a);
    return INV ? float2{-r * (a.x + a.y), r * (a.x - a.y)} : float2{r * (a.y - a.x), -r * (a.x + a.y)};
}
template <int R, bool INV> struct Dft;
template <bool INV> struct Dft<2, INV> {
    static __device__ __forceinline__ void run(float2 (&v)[2]) { dft2<INV>(v[0], v[1]); }
};
template <bool INV> struct Dft<4, INV> {
    static __device__ __forceinline__ void run(float2 (&v)[4]) { dft4<INV>(v[0], v[1], v[2], v[3]); }
};
template <bool INV> struct Dft<8, INV> {
    // n = c + 2 n1 (c < 2, n1 < 4), k = q + 4 p:  X[q + 4p] = sum_c W2^(cp) W8^(cq) sum_n1 x[c + 2 n1] W4^(n1 q)
    static __device__ __forceinline__ void run(float2 (&v)[8]) {
        dft4<INV>(v[0], v[2], v[4], v[6]);                 // c = 0: q in v[0], v[2], v[4], v[6]
        dft4<INV>(v[1], v[3], v[5], v[7]);                 // c = 1
        v[3] = mul_w8<INV, 1>(v[3]); v[5] = mul_w8<INV, 2>(v[5]); v[7] = mul_w8<INV, 3>(v[7]);
        // radix-2 over c for every q: outputs X[q] and X[q + 4]
        float2 o[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) { o[q] = cadd(v[2 * q], v[2 * q + 1]); o[q + 4] = csub(v[2 * q], v[2 * q + 1]); }
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = o[i];
    }
};
// ---- sources of a stage's inputs: the shared buffer itself, or the resampled native spectrum (first stage) ----
struct LoadZ {
    const float2 *z;
    __device__ __forceinline__ float2 get(int, int pm) const { return z[pm]; }          // pm = padi(m), supplied by the caller
};
// Affine resampling position pos(k) = k * ratio (0 < ratio < 4) in 32.32 fixed point: the integer part is the
// high word of one 64-bit multiply, the fraction its low word (rounding of the ratio: < 2^-33 per step, < 2e-6
// pixel at k = 16384, i.e. < 1e-6 of the local flux slope -- far below the fp32 tolerance; the f64 mode keeps exact
// positions).  The interpolation is linear in ln(lambda) here; the reference's np.interp is linear in lambda, which
// differs by t (1 - t) d ln(lambda) / 2 <= 4e-7 of the pixel-to-pixel flux difference (also kept in the f64 mode).
struct AffinePos {
    unsigned long long step;          // round(ratio * 2^32)
    __device__ __forceinline__ void set(double ratio) { step = (unsigned long long)(ratio * 4294967296.0 + 0.5); }
    __device__ __forceinline__ int split(int k, float &t) const {
        const unsigned long long fx = (unsigned long long)(unsigned)k * step;
        t = (float)(unsigned)fx * (1.f / 4294967296.f);
        return (int)(fx >> 32);
    }
};

struct LoadResampled {            // z[m] = x[2m] + i x[2m+1], x[k] = native spectrum interpolated at k * ratio
    const float *fr; AffinePos ap; // fr[npix] duplicates fr[npix - 1]: the last sample needs no clamp
    __device__ __forceinline__ float at(int k) const {
        float t;
        const int i = ap.split(k, t);
        const float a = fr[i], b = fr[i + 1];
        return fmaf(b - a, t, a);
    }
    __device__ __forceinline__ float2 get(int m, int) const { return {at(2 * m), at(2 * m + 1)}; }
};

// split / transfer / merge of one conjugate pair of the packed real transform (k, M - k):
// in: A = Z[k], B = Z[M-k]; u = exp(-2 pi i k / n); tk, tm = transfer(k), transfer(M-k) (1/M folded in)
__device__ __forceinline__ void real_pair(float2 &A, float2 &B, const float2 u, const float tk, const float tm) {
    const float2 Bc = {B.x, -B.y};
    const float2 E = {0.5f * (A.x + Bc.x), 0.5f * (A.y + Bc.y)};
    const float2 dif = csub(A, Bc);
    const float2 O = {0.5f * dif.y, -0.5f * dif.x};                   // -i/2 (A - conj B)
    const float2 wO = cmul(u, O);
    const float2 Yk = {tk * (E.x + wO.x), tk * (E.y + wO.y)};          // X'[k]
    const float2 Ym = {tm * (E.x - wO.x), tm * (E.y - wO.y)};          // conj X'[M-k]
    const float2 Ep = {0.5f * (Yk.x + Ym.x), 0.5f * (Yk.y + Ym.y)};
    const float2 D = {0.5f * (Yk.x - Ym.x), 0.5f * (Yk.y - Ym.y)};
    const float2 Op = cmulc(D, u);
    A = {Ep.x - Op.y, Ep.y + Op.x};                                    // Ep + i Op
    B = {Ep.x + Op.y, Op.x - Ep.y};                                    // conj(Ep) + i conj(Op)
}

struct RotTransfer {
    // u = coef k in f32: relative error 6e-8 of u, i.e. < 2e-5 at the end of the table where the transfer function
    // (~ u^-3/2) and its slope are below 3e-4 -- far inside the fp32 budget of the spectrum itself
    float coef; const float *tab; float inv;
    __device__ __forceinline__ float operator()(int k) const {
        const float u = coef * (float)k;
        if (!(u == u)) return __int_as_float(0x7fc00000);
        return inv * rot_transfer_f(u, tab);
    }
};


// ---- the transform: decimation-in-frequency forward, decimation-in-time inverse, in place ---------------------
// A Stockham (out-of-place order, in-place storage) pass has to hold all of a thread's outputs across a CTA barrier (16 complex numbers at 512 threads) and the
// 64-register budget of two CTAs per SM then spills ~32 registers per pass to local memory (ncu r02: 12 M local requests
// per 1536 spectra, L1 hit rate 23 %).  A convolution does not care about the order of the spectrum, so the forward
// transform here is the Gentleman-Sande in-place form (natural order in, digit-reversed out), the transfer function is
// applied at the permuted positions and the inverse is the mirrored Cooley-Tukey form (digit-reversed in, natural out):
// every butterfly writes the slots it read, one butterfly is live per thread, one barrier per pass.
// Stage list for M = 2^lM: [2^lL leftover] 8 ... 8, then a final radix 2 that is fused -- together with the first inverse
// radix 2 -- into the split / transfer / merge pass (ip_quad_pass).  X[k] ends at the position whose digits are those of
// k in reversed stage order.
struct IpPlan {
    static constexpr int LRM = 3;  // main radix 8 (radix 16 was measured: same time, more spills)
    int lM, lL, nm;               // log2 M; log2 of the leftover radix (0 .. LRM-1); number of main-radix stages
    __device__ __forceinline__ void set(int lM_) { lM = lM_; lL = (lM_ - 1) % LRM; nm = (lM_ - 1) / LRM; }
    __device__ __forceinline__ int freq_at(int i) const {           // frequency index stored at position i
        int rem = lM, shift = 0, k = 0;
        if (lL) { rem -= lL; k = (i >> rem) & ((1 << lL) - 1); shift = lL; }
        for (int s = 0; s < nm; ++s) { rem -= LRM; k |= ((i >> rem) & ((1 << LRM) - 1)) << shift; shift += LRM; }
        return k | ((i & 1) << shift);
    }
    __device__ __forceinline__ int pos_of(int k) const {            // position of frequency index k
        int rem = lM, shift = 0, i = 0;
        if (lL) { rem -= lL; i = (k & ((1 << lL) - 1)) << rem; shift = lL; }
        for (int s = 0; s < nm; ++s) { rem -= LRM; i |= ((k >> shift) & ((1 << LRM) - 1)) << rem; shift += LRM; }
        return i | ((k >> shift) & 1);
    }
};

// One in-place stage over blocks of length L = 2^lL_, radix R = 2^LR, stride S = L / R.  Forward: DFT over the R inputs,
// then output p times W_L^(j p).  Inverse: input p times conj(W_L^(j p)), then the inverse DFT.
template <int LR, bool INV, typename LOADER>
__device__ __forceinline__ void ip_stage(float2 *z, int lM, int lL_, const float2 *__restrict__ twlev, const LOADER &ld) {
    constexpr int R = 1 << LR;
    const int lS = lL_ - LR, S = 1 << lS;
    const int nbf = 1 << (lM - LR), lnblk = lM - lL_;
    const float2 *tw = twlev + (1 << (lL_ - 1));                    // exp(-2 pi i j / L), j < L / 2
    const bool lin = lS >= 4 || lL_ <= 4;
    const int pstride = lS >= 4 ? S + (S >> 4) : S;
#pragma unroll IP_UNROLL
    for (int t = threadIdx.x; t < nbf; t += FT) {
        int blk, j;
        if (lS >= 4) { blk = t >> lS; j = t & (S - 1); }            // consecutive threads: consecutive j
        else { blk = t & ((1 << lnblk) - 1); j = t >> lnblk; }      // short strides: consecutive blocks (padded stride L + L/16)
        const int base = (blk << lL_) + j;
        // padded slot of element r: linear in r when the stride is a multiple of 16 or the whole block sits in one
        // 16-element padding group (all stages of the usual sizes); otherwise element by element
        const int pb = padi(base);
        auto slot = [&](int r) { return lin ? pb + r * pstride : padi(base + (r << lS)); };
        float2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = ld.get(base + (r << lS), slot(r));
        if (!INV) Dft<R, false>::run(v);
        if (lS > 0) {
            float2 w1 = tw[j];
            if (INV) w1.y = -w1.y;
            v[1] = cmul(v[1], w1);
            if (R >= 4) {
                const float2 w2 = cmul(w1, w1), w3 = cmul(w2, w1);
                v[2] = cmul(v[2], w2); v[3] = cmul(v[3], w3);
                if (R == 8) {
                    const float2 w4 = cmul(w2, w2);
                    v[4] = cmul(v[4], w4); v[5] = cmul(v[5], cmul(w4, w1));
                    v[6] = cmul(v[6], cmul(w3, w3)); v[7] = cmul(v[7], cmul(w4, w3));
                }
            }
        }
        if (INV) Dft<R, true>::run(v);
#pragma unroll
        for (int r = 0; r < R; ++r) z[slot(r)] = v[r];
    }
    __syncthreads();
}

// Final forward radix 2 (adjacent positions, no twiddle), real-spectrum split, transfer function, merge and first
// inverse radix 2, in place.  Positions (2a, 2a+1) hold the radix-2 inputs of frequencies (k, k + M/2) with
// k = freq_at(2a) < M/2; a quadruple {k, M/2-k, M/2+k, M-k} is two such position pairs, owned by the thread of k < M/4.
__device__ __forceinline__ void ip_quad_pass(float2 *z, const IpPlan &pl, const RotTransfer &tf,
                                             const float2 *__restrict__ twlev) {
    const int M = 1 << pl.lM, H = M >> 1, Q = M >> 2;
    const float2 *twN = twlev + M;                 // exp(-2 pi i k / (2M)), k < M
    const int lrp = pl.nm ? IpPlan::LRM : pl.lL;           // log2 radix of the last stage before the fused radix 2
    const int lh = lrp - 1;
#pragma unroll QUAD_UNROLL
    for (int w = threadIdx.x; w < Q; w += FT) {
        const int a = ((w >> lh) << lrp) + (w & ((1 << lh) - 1));
        const int k = pl.freq_at(2 * a);
        if (k == 0) continue;
        const int ap = pl.pos_of(H - k);
        const float2 x0 = z[padi(2 * a)], x1 = z[padi(2 * a + 1)], y0 = z[padi(ap)], y1 = z[padi(ap + 1)];
        float2 Za = cadd(x0, x1), Zc = csub(x0, x1);                   // Z[k], Z[k + H]
        float2 Zb = cadd(y0, y1), Zd = csub(y0, y1);                   // Z[H - k], Z[M - k]
        const float2 u = twN[k];
        real_pair(Za, Zd, u, tf(k), tf(M - k));
        const float2 ub = {-u.y, -u.x};                                // exp(-2 pi i (H - k) / (2M)) = -i conj(u)
        real_pair(Zb, Zc, ub, tf(H - k), tf(H + k));
        z[padi(2 * a)] = cadd(Za, Zc); z[padi(2 * a + 1)] = csub(Za, Zc);
        z[padi(ap)] = cadd(Zb, Zd); z[padi(ap + 1)] = csub(Zb, Zd);
    }
    if (threadIdx.x == 0) {
        {   // k = 0 (pairs with itself) and k = H (with itself): positions 0, 1
            const float2 x0 = z[0], x1 = z[padi(1)];
            float2 Z0 = cadd(x0, x1), ZH = csub(x0, x1);
            const float y0 = tf(0) * (Z0.x + Z0.y), ym = tf(M) * (Z0.x - Z0.y);
            Z0 = {0.5f * (y0 + ym), 0.5f * (y0 - ym)};
            float2 dup = ZH;
            const float th = tf(H);
            real_pair(ZH, dup, twN[H], th, th);
            z[0] = cadd(Z0, ZH); z[padi(1)] = csub(Z0, ZH);
        }
        if (Q > 0) {   // k = M/4: the pair (Q, 3Q)
            const int pq = pl.pos_of(Q);
            const float2 x0 = z[padi(pq)], x1 = z[padi(pq + 1)];
            float2 Za = cadd(x0, x1), Zc = csub(x0, x1);
            real_pair(Za, Zc, twN[Q], tf(Q), tf(M - Q));
            z[padi(pq)] = cadd(Za, Zc); z[padi(pq + 1)] = csub(Za, Zc);
        }
    }
    __syncthreads();
}

// y = irfft(rfft(x) * transfer) of the n1 = 2M samples obtained by resampling the native spectrum in `fr`;
// result left in z as M complex numbers z[m] = y[2m] + i y[2m+1] (padded layout, natural order)
__device__ __noinline__ void rotational_filter_ip(float2 *z, const float *fr, int npix, int n1, double dlnw, double vrot,
                                                  const float *rot_tab, const float2 *__restrict__ twlev) {
    const int M = n1 >> 1;
    constexpr int LRM = IpPlan::LRM;
    IpPlan pl;
    pl.set(31 - __clz(M));
    LoadResampled src;
    src.fr = fr; src.ap.set((double)(npix - 1) / (double)(n1 - 1));
    const LoadZ own{z};
    const double step1 = (double)(npix - 1) * dlnw / (double)(n1 - 1);
    RotTransfer tf;
    tf.tab = rot_tab; tf.inv = 1.f / (float)M;
    tf.coef = (float)(2.0 * M_PI * sqrt(vrot * vrot) / ((double)n1 * MS_CKMS * step1));
    if (pl.lM == 13) {
        // n1 = 16384 (nets of 8193 .. 16384 pixels): the same stage list with every length a literal, so strides, masks
        // and the digit reversal fold into immediates
        IpPlan pc;
        pc.set(13);
        ip_stage<3, false>(z, 13, 13, twlev, src);
        ip_stage<3, false>(z, 13, 10, twlev, own);
        ip_stage<3, false>(z, 13, 7, twlev, own);
        ip_stage<3, false>(z, 13, 4, twlev, own);
        ip_quad_pass(z, pc, tf, twlev);
        ip_stage<3, true>(z, 13, 4, twlev, own);
        ip_stage<3, true>(z, 13, 7, twlev, own);
        ip_stage<3, true>(z, 13, 10, twlev, own);
        ip_stage<3, true>(z, 13, 13, twlev, own);
        return;
    }
    int l = pl.lM;
    bool first = true;
    if (pl.lL == 1) { ip_stage<1, false>(z, pl.lM, l, twlev, src); l -= 1; first = false; }
    else if (pl.lL == 2) { ip_stage<2, false>(z, pl.lM, l, twlev, src); l -= 2; first = false; }
    for (int s = 0; s < pl.nm; ++s) {
        if (first) ip_stage<LRM, false>(z, pl.lM, l, twlev, src);
        else ip_stage<LRM, false>(z, pl.lM, l, twlev, own);
        first = false; l -= LRM;
    }
    ip_quad_pass(z, pl, tf, twlev);
    l = 1;
    for (int s = 0; s < pl.nm; ++s) { l += LRM; ip_stage<LRM, true>(z, pl.lM, l, twlev, own); }
    if (pl.lL == 1) ip_stage<1, true>(z, pl.lM, l + 1, twlev, own);
    else if (pl.lL == 2) ip_stage<2, true>(z, pl.lM, l + 2, twlev, own);
}

__device__ __forceinline__ float conv_at(const float2 *z, int k) {          // real sample k of the filtered sequence
    const float2 c = z[padi(k >> 1)];
    return (k & 1) ? c.y : c.x;
}

// Per-point decisions (window, grid of the instrumental stage, kernel width, which path), one thread per point, ahead
// of the broadening kernel: its CTAs fetch the 96-byte record of their NEXT point with an async copy while they work on
// the current one, instead of one thread computing it (three f64 logarithms behind two dependent global loads) with
// 511 threads waiting.  Rows the fast kernel declines are appended to the fallback list here.
__global__ void spec_setup_kernel(SmoothArgs a, int nB, PointSetup *__restrict__ out, int *__restrict__ fb_rows,
                                  int *__restrict__ fb_count) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nB) return;
    const int npix = a.npix;
    const double *r = a.pars + (size_t)p * MS_NPARS;
    const int slot = a.star_slot ? a.star_slot[p] : a.slot0;
    const bool slot_ok = (unsigned)slot < (unsigned)a.nstars;
    const StarView sv = a.stars[slot_ok ? slot : 0];
    PointSetup q;
    q.nobs = slot_ok ? sv.n_obs : 0;
    q.i_s = 0; q.i_e = npix - 1; q.n2 = 0; q.ntap = 0; q.d2 = 0.0; q.ratio2 = 0.0; q.spx = 0.0;
    const double vrad = r[PV_VRAD];
    q.vrot = r[PV_VROT]; q.rot = q.vrot != 0.0;
    q.inst = 2.355 * r[PV_INSTR];                                    // genmod.py:74-75
    const double scale = (vrad != 0.0) ? 1.0 + vrad / MS_C_KMS : 1.0;
    q.lnscale = log(scale);
    if (r[PV_TEFF] == ms_inf() || q.nobs <= 0) q.mode = 4;
    else if (q.inst != 0.0) {
        if (!(q.inst == q.inst) || !(scale == scale)) q.mode = 2;
        else {
            mask_window(a.wave, npix, a.lnw0, a.dlnw, scale, sv.wmin * (1.0 - 20.0 / q.inst),
                        sv.wmax * (1.0 + 20.0 / q.inst), q.i_s, q.i_e);
            const int nmask = q.i_e - q.i_s + 1;
            const double so = MS_CKMS / q.inst, si = MS_CKMS / a.resolution;
            const double sigma = sqrt(so * so - si * si);            // NaN when sharper than the net
            if (nmask < 4) q.mode = 2;
            else if (sigma <= 0.0) q.mode = 1;
            else if (!(sigma == sigma)) q.mode = 3;                  // NaN kernel: let the generic kernel propagate it
            else {
                int n2 = 1; while (n2 < nmask) n2 <<= 1;
                q.n2 = n2;
                q.ratio2 = (double)(q.i_e - q.i_s) / (double)(n2 - 1);
                q.d2 = q.ratio2 * a.dlnw;
                q.spx = sigma / (MS_CKMS * q.d2);                    // kernel width in pixels of the n2 grid
                q.ntap = (int)ceil(5.5 * q.spx);
                q.mode = (q.spx >= 2.0 && q.ntap <= MAXTAP && 2 * q.ntap + 12 < n2) ? 0 : 3;
            }
        }
    } else {
        q.mode = 1;
    }
    q.lnw_s = a.lnw0 + (double)q.i_s * a.dlnw;
    q.inv_d2 = q.d2 != 0.0 ? 1.0 / q.d2 : 0.0;
    q.slot_ok = slot_ok; q.teff_inf = r[PV_TEFF] == ms_inf();
    if (q.mode == 3) fb_rows[atomicAdd(fb_count, 1)] = p;
    out[p] = q;
}

__device__ __forceinline__ void fetch_setup(PointSetup *dst, const PointSetup *src) {       // 6 lanes x 16 bytes
    if (threadIdx.x < sizeof(PointSetup) / 16) {
        const uint32_t d = (uint32_t)__cvta_generic_to_shared(reinterpret_cast<unsigned char *>(dst) + 16 * threadIdx.x);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(reinterpret_cast<const unsigned char *>(src) + 16 * threadIdx.x) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

__global__ void __launch_bounds__(FT, 2)
spec_fast_kernel(SmoothArgs a, int nB, const PointSetup *__restrict__ setup) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int npix = a.npix, n1 = a.n1, M = n1 >> 1;
    float2 *z = reinterpret_cast<float2 *>(smem_raw);                        // padded complex buffer
    float *s2 = reinterpret_cast<float *>(smem_raw);                         // later: the n2-point instrumental grid
    const size_t fr_off = (((size_t)padi(M) * sizeof(float2) + 15) & ~(size_t)15) + 16;
    float *fr = reinterpret_cast<float *>(smem_raw + fr_off);                // native / rotated spectrum
    __shared__ __align__(16) float wtab[4 * WSTRIDE];                        // FIR weights for the 4 window alignments
    __shared__ float gtap[MAXTAP + 2];
    __shared__ double red[FT / 32];
    __shared__ __align__(16) PointSetup ps2[2];
    const float2 *twlev = static_cast<const float2 *>(a.twlev);

    if ((int)blockIdx.x < nB) fetch_setup(&ps2[0], setup + blockIdx.x);
    int it = 0;
    bool fr_ready = false;                 // this point's native spectrum was fetched during the previous point's last phase
    for (int p = blockIdx.x; p < nB; p += gridDim.x, ++it) {
        const double *r = a.pars + (size_t)p * MS_NPARS;
        asm volatile("cp.async.wait_all;" ::: "memory");                      // this point's record (issued one point ago)
        __syncthreads();
        const PointSetup &ps = ps2[it & 1];
        if (p + (int)gridDim.x < nB) fetch_setup(&ps2[(it & 1) ^ 1], setup + p + gridDim.x);
        const bool slot_ok = ps.slot_ok != 0;
        const int slot = a.star_slot ? a.star_slot[p] : a.slot0;
        const StarView sv = a.stars[slot_ok ? slot : 0];
        const int mode = ps.mode, nobs = ps.nobs;
        if (mode >= 3) {                                                      // uniform: skipped or handed to the generic kernel
            if (mode == 4) {
                if (a.flux_out && slot_ok)
                    for (int j = threadIdx.x; j < nobs; j += FT) a.flux_out[(size_t)p * nobs + j] = ms_nan();
                if (a.chi2 && !slot_ok && threadIdx.x == 0 && !ps.teff_inf) a.chi2[p] = ms_nan();
            }
            fr_ready = false;
            __syncthreads();
            continue;
        }
        const int i_s = ps.i_s, i_e = ps.i_e, n2 = ps.n2, ntap = ps.ntap;
        const double d2 = ps.d2, inst = ps.inst, inv_d2 = ps.inv_d2;
        // ---- native spectrum -> shared memory (unless the previous point's last phase already fetched it)
        {
            const float *frow = static_cast<const float *>(a.flux_native) + (size_t)p * npix;
            if (fr_ready) {
                if (threadIdx.x == FT - 1) fr[npix] = fr[npix - 1];
            } else {
                if ((npix & 3) == 0) {
                    const float4 *src4 = reinterpret_cast<const float4 *>(frow);
                    float4 *dst4 = reinterpret_cast<float4 *>(fr);
                    for (int i = threadIdx.x; i < (npix >> 2); i += FT) dst4[i] = __ldcs(src4 + i);
                } else {
                    for (int i = threadIdx.x; i < npix; i += FT) fr[i] = __ldcs(frow + i);
                }
                if (threadIdx.x == FT - 1) fr[npix] = __ldcs(frow + npix - 1);   // one-sample pad: interpolation at the last pixel
            }
            fr_ready = false;
        }
        if (mode == 0 && threadIdx.x <= ntap) {                               // Gaussian taps, t = 0 .. ntap
            const double t = (double)threadIdx.x / ps.spx;
            gtap[threadIdx.x] = (float)(exp(-0.5 * t * t) / (ps.spx * 2.5066282746310002));
        }
        __syncthreads();
        // FIR weight rows: W[o][u] = g[u - o], g[s] = tap (ntap - s) for s = 0 .. 2 ntap, zero elsewhere.  An observed
        // pixel whose window starts o samples after a 16-byte boundary uses rows o (sample i) and o + 1 (sample i + 1).
        const int U = (2 * ntap + 2 + 3 + 3) & ~3;
        if (mode == 0) {
            for (int idx = threadIdx.x; idx < 4 * U; idx += FT) {
                const int o = idx / U, u = idx - o * U;
                const int sidx = u - o;
                float w = 0.f;
                if (sidx >= 0 && sidx <= 2 * ntap) { const int t = ntap - sidx; w = gtap[t < 0 ? -t : t]; }
                wtab[o * WSTRIDE + u] = w;
            }
        }

        // ---- rotational broadening on the full native grid (outwave = None), back onto the native pixels
        if (ps.rot) {
            rotational_filter_ip(z, fr, npix, n1, a.dlnw, ps.vrot, a.rot_tab, twlev);
            // smoothing.py:262: interpolation of the filtered n1 grid at the native wavelengths
            AffinePos ap;
            ap.set((double)(n1 - 1) / (double)(npix - 1));
#pragma unroll BACK_UNROLL
            for (int i = threadIdx.x; i < npix; i += FT) {
                float t;
                int k = ap.split(i, t);
                if (k > n1 - 2) { k = n1 - 2; t = 1.f; }
                const float v0 = conv_at(z, k), v1 = conv_at(z, k + 1);
                fr[i] = fmaf(v1 - v0, t, v0);
            }
            __syncthreads();
        }

        // ---- Doppler shift + instrumental broadening + resampling to the observed pixels
        const double lnscale = ps.lnscale;
        if (mode == 0) {
            AffinePos ap;
            ap.set(ps.ratio2);
            const float *fs = fr + i_s;                                       // fs[last + 1] exists (fr is padded by one sample)
#pragma unroll 4
            for (int k = threadIdx.x; k < n2; k += FT) {
                float t;
                const int i = ap.split(k, t);
                const float v0 = fs[i], v1 = fs[i + 1];
                s2[k] = fmaf(v1 - v0, t, v0);
            }
        }
        __syncthreads();
        // the native spectrum is dead from here on in this mode: fetch the next point's row into its place while the
        // Gaussian sums run (async copies, waited for at the top of the next iteration)
        if (mode == 0 && (npix & 3) == 0 && p + (int)gridDim.x < nB) {
            const float *nrow = static_cast<const float *>(a.flux_native) + (size_t)(p + gridDim.x) * npix;
            const uint32_t d0 = (uint32_t)__cvta_generic_to_shared(fr);
            for (int i = threadIdx.x; i < (npix >> 2); i += FT)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0 + 16u * (uint32_t)i), "l"(nrow + 4 * i) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
            fr_ready = true;
        }
        double csum = 0.0;
        const double lnw_s = ps.lnw_s;
        const int mask2 = n2 - 1;
        // the model at one observed pixel from the (smoothed) grid values; the Gaussian sums of the instrumental stage
        // are evaluated by fir_pair below for two pixels at a time (independent chains hide the shared-memory latency)
        auto finish = [&](int j, double m) {
            if (a.modpoly) m *= cheb_eval(sv.cheb_x[j], r + PV_PC0, a.n_pc);  // genmod.py:90-93
            if (a.flux_out) a.flux_out[(size_t)p * nobs + j] = m;
            if (a.chi2) {
                const double d = m - sv.obs_flux[j];
                const double t = (d * d) / sv.obs_err2[j];
                if (t == t) csum += t;                                        // nansum (likelihood.py:193)
            }
        };
        if (mode == 0) {
            auto locate = [&](int j, int &i, double &f) {
                const double q = (sv.obs_lnwave[j] - lnscale - lnw_s) * inv_d2;
                if (q <= 0.0) { i = 0; f = 0.0; }
                else if (q >= (double)(n2 - 1)) { i = n2 - 2; f = 1.0; }
                else {
                    i = (int)q;
                    const double t = q - (double)i;
                    f = t * (1.0 + (t - 1.0) * 0.5 * d2);
                }
            };
            // conv[i] and conv[i + 1] of the circular Gaussian convolution, sharing the samples
            auto fir_slow = [&](int i, float &a0, float &a1) {                // window wraps around the circular grid
                const int base = i - ntap;
                float x = s2[base & mask2];
                a0 = 0.f; a1 = 0.f;
                for (int s = 0; s <= 2 * ntap; ++s) {
                    const int t = ntap - s;
                    const float g = gtap[t < 0 ? -t : t];
                    a0 = fmaf(g, x, a0);                                      // sample i - t
                    x = s2[(base + s + 1) & mask2];
                    a1 = fmaf(g, x, a1);                                      // sample i + 1 - t
                }
            };
            // conv[i] = sum_u W[o][u] x[b + u] and conv[i + 1] = sum_u W[o][u] x[b + u + 1]: one weight row, the sample
            // window shifted by one in registers (the first sample of the next 16-byte chunk is carried along), so a
            // chunk costs two 16-byte shared loads for eight multiply-adds.  W[o][U - 1] = 0: nothing past the window counts.
            const int nch = U >> 2;
            for (int j = threadIdx.x; j < nobs; j += FT) {
                int i; double f;
                locate(j, i, f);
                const int b = (i - ntap) & ~3, o = (i - ntap) - b;
                float a0 = 0.f, a1 = 0.f;
                if (b >= 0 && b + U <= n2) {
                    const float4 *xa = reinterpret_cast<const float4 *>(s2 + b);
                    const float4 *wa = reinterpret_cast<const float4 *>(wtab + o * WSTRIDE);
                    float a0b = 0.f, a1b = 0.f;
                    float4 x = xa[0];
#pragma unroll FIR_UNROLL
                    for (int c = 0; c < nch; ++c) {
                        const float4 g = wa[c];
                        float4 xn = {0.f, 0.f, 0.f, 0.f};
                        if (c + 1 < nch) xn = xa[c + 1];
                        a0 = fmaf(g.x, x.x, a0); a0b = fmaf(g.y, x.y, a0b); a0 = fmaf(g.z, x.z, a0); a0b = fmaf(g.w, x.w, a0b);
                        a1 = fmaf(g.x, x.y, a1); a1b = fmaf(g.y, x.z, a1b); a1 = fmaf(g.z, x.w, a1); a1b = fmaf(g.w, xn.x, a1b);
                        x = xn;
                    }
                    a0 += a0b; a1 += a1b;
                } else {
                    fir_slow(i, a0, a1);
                }
                finish(j, (double)a0 + ((double)a1 - (double)a0) * f);
            }
        } else {
            for (int j = threadIdx.x; j < nobs; j += FT) {
                double m;
                if (mode == 1) {
                    const double q = (sv.obs_lnwave[j] - lnscale - a.lnw0) / a.dlnw;
                    const double qlo = (double)i_s, qhi = (double)i_e;
                    if (q < qlo || q > qhi) {
                        // smoothing branch clamps (np.interp default); plain branch returns NaN outside
                        m = (inst != 0.0) ? (double)fr[q < qlo ? i_s : i_e] : ms_nan();
                    } else {
                        int i = (int)q;
                        if (i > i_e - 1) i = i_e - 1;
                        const double t = q - (double)i;
                        const double f = t * (1.0 + (t - 1.0) * 0.5 * a.dlnw);
                        const double v0 = (double)fr[i], v1 = (double)fr[i + 1];
                        m = v0 + (v1 - v0) * f;
                    }
                } else {
                    m = ms_nan();
                }
                finish(j, m);
            }
        }
        if (a.chi2) {
            for (int o = 16; o > 0; o >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, o);
            if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = csum;
            __syncthreads();
            if (threadIdx.x == 0) {
                double tot = 0.0;
                for (int w = 0; w < FT / 32; ++w) tot += red[w];
                a.chi2[p] += tot;
            }
        }
        __syncthreads();
    }
}

size_t fast_smem_bytes(int npix, int n1) {
    const int M = n1 >> 1;
    const size_t fr_bytes = ((size_t)(npix + 1) * sizeof(float) + 15) & ~(size_t)15;
    return (((size_t)(M + (M >> 4)) * sizeof(float2) + 15) & ~(size_t)15) + 16 + fr_bytes + 16;
}

}  // namespace

bool spec_fast_usable(const ms_ctx *ctx) {
    const SpecNetDev &n = ctx->spec;
    return ctx->precision != MS_PREC_F64 && n.n1 >= 8 && (n.n1 >> 1) <= MAXM &&
           fast_smem_bytes(n.npix, n.n1) <= 226 * 1024;
}

// Two CTAs per SM when two spectra fit the shared memory (nets up to ~11 000 pixels), else one.
bool spec_fast_two_per_sm(const ms_ctx *ctx) {
    return 2 * (fast_smem_bytes(ctx->spec.npix, ctx->spec.n1) + 1024 + 4096) <= 227 * 1024;
}

// Rows the kernel declines are appended to fb_rows / *fb_count (device; the caller zeroes the counter).  setup: scratch
// for nB PointSetup records.
int launch_spec_fast(ms_ctx *ctx, const SmoothArgs &sa, int nB, void *setup, int *fb_rows, int *fb_count, cudaStream_t st) {
    const size_t smem = fast_smem_bytes(sa.npix, sa.n1);
    static size_t cfg = 0;
    if (cfg < smem) { MS_CUDA(cudaFuncSetAttribute(spec_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); cfg = smem; }
    int grid = ctx->sm_count * (spec_fast_two_per_sm(ctx) ? 2 : 1);
    if (grid > nB) grid = nB;
    PointSetup *su = static_cast<PointSetup *>(setup);
    spec_setup_kernel<<<(nB + 127) / 128, 128, 0, st>>>(sa, nB, su, fb_rows, fb_count);
    ctx->launches++;
    MS_CUDA(cudaGetLastError());
    spec_fast_kernel<<<grid, FT, smem, st>>>(sa, nB, su);
    ctx->launches++;
    MS_CUDA(cudaGetLastError());
    return MS_OK;
}
