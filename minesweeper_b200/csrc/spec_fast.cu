// Kernel (3), f32 fast path: rotational + instrumental broadening, resampling onto the observed
// pixels, blaze polynomial and the spectral chi-square for the fp32-class modes (MS_PREC_F32 / TC).
// Same contract and reference lines as spec_smooth.cu (smoothing.py:92-114, 131-134, 202-264, 517-598;
// fitutils.py:11-20; likelihood.py:188-195); that file keeps the generic kernel (f64, and the f32 rows
// this kernel declines).  What is different here, and why it still reproduces the reference's numbers:
//
//  * Rotational stage (smooth_vsini_fft): the reference's kernel is DEFINED by its FFT construction (analytic
//    transfer function sampled on the rfft frequencies of the power-of-two log grid: band-limited, 1/j^2 tails),
//    so it stays an FFT -- but the real transform of n1 samples is one in-place complex transform of M = n1/2
//    points in (padded) shared memory with radix-16 Stockham passes staged through registers (3 passes for
//    M/2 = 4096 instead of 5 radix-8 ones, one buffer instead of two), and the last forward radix-2 stage, the
//    real-spectrum split, the transfer function, the merge and the first inverse radix-2 stage are ONE pass over
//    index quadruples {k, M/2-k, M/2+k, M-k}.  The first pass reads the native spectrum through the linear
//    resampling (smoothing.py:587-597) instead of from a resampled copy.
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
constexpr int FT = MS_FAST_FT;          // threads per CTA (256: experiment, two CTAs per SM at 128 registers)
constexpr int MAXM = 8192;              // complex transform length limit (n1 <= 16384)
constexpr int EPT = MAXM / FT;          // complex elements a thread stages per pass: 16
constexpr int MAXTAP = 64;              // half-width limit of the Gaussian FIR (5.5 sigma: sigma <= 11.6 pixels)
constexpr int WU = 2 * MAXTAP + 8;      // longest aligned FIR window (multiple of 4)
constexpr int WSTRIDE = WU + 4;         // row stride of the weight table: 4 mod 32 floats -> the 5 rows hit distinct banks

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
// multiply by W16^m for the odd m that are not multiples of W8: m = 1, 3 (and 9 = -W16^1, handled by the caller)
template <bool INV, int m> __device__ __forceinline__ float2 mul_w16(float2 a) {
    const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;   // cos, sin of pi / 8
    const float c = (m == 1) ? c1 : s1, s = (m == 1) ? s1 : c1;                 // W16^m = c - i s (forward)
    return INV ? float2{a.x * c - a.y * s, a.x * s + a.y * c} : float2{a.x * c + a.y * s, a.y * c - a.x * s};
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
template <bool INV> struct Dft<16, INV> {
    // n = c + 4 n1, k = q + 4 p:  X[q + 4p] = sum_c W4^(cp) [ W16^(cq) sum_n1 x[c + 4 n1] W4^(n1 q) ]
    static __device__ __forceinline__ void run(float2 (&v)[16]) {
#pragma unroll
        for (int c = 0; c < 4; ++c) dft4<INV>(v[c], v[c + 4], v[c + 8], v[c + 12]);      // result q at v[c + 4q]
        // twiddles W16^(cq), c, q in 1..3
        v[5] = mul_w16<INV, 1>(v[5]);                              // c1 q1: W16^1
        v[9] = mul_w8<INV, 1>(v[9]);                               // c1 q2: W16^2 = W8^1
        v[13] = mul_w16<INV, 3>(v[13]);                            // c1 q3: W16^3
        v[6] = mul_w8<INV, 1>(v[6]);                               // c2 q1: W16^2
        v[10] = mul_mi<INV>(v[10]);                                // c2 q2: W16^4 = -i
        v[14] = mul_w8<INV, 3>(v[14]);                             // c2 q3: W16^6 = W8^3
        v[7] = mul_w16<INV, 3>(v[7]);                              // c3 q1: W16^3
        v[11] = mul_w8<INV, 3>(v[11]);                             // c3 q2: W16^6
        { const float2 t = mul_w16<INV, 1>(v[15]); v[15] = {-t.x, -t.y}; }   // c3 q3: W16^9 = -W16^1
        float2 o[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float2 a = v[4 * q], b = v[4 * q + 1], c = v[4 * q + 2], d = v[4 * q + 3];
            dft4<INV>(a, b, c, d);                                 // over c -> p
            o[q] = a; o[q + 4] = b; o[q + 8] = c; o[q + 12] = d;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = o[i];
    }
};

// ---- one Stockham pass, in place: every thread reads all its inputs, the CTA synchronises, then writes -------
// Natural order in, natural order out after the last pass.  Butterfly j (< M / R) reads z[j + r M/R], multiplies
// by exp(-+2 pi i k r / (Ns R)), k = j mod Ns, transforms, writes z[(j - k) R + k + r Ns].  `twlev` holds the
// per-level tables: level l at offset 2^(l-1), entry k = exp(-2 pi i k / 2^l), k < 2^(l-1).
// LOADER: source of the inputs (the shared buffer itself, or the resampled native spectrum for the first pass).
struct LoadZ {
    const float2 *z;
    __device__ __forceinline__ float2 operator()(int m) const { return z[padi(m)]; }
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
    __device__ __forceinline__ float2 operator()(int m) const { return {at(2 * m), at(2 * m + 1)}; }
};

template <int R, bool INV, typename LOADER>
__device__ __noinline__ void fft_pass(float2 *z, int M, int lNs, const float2 *__restrict__ twlev, const LOADER &ld) {
    constexpr int LR = (R == 16) ? 4 : (R == 8 ? 3 : (R == 4 ? 2 : 1));
    constexpr int NBT = EPT / R;                          // butterflies per thread when M = MAXM
    const int nb = M >> LR, Ns = 1 << lNs;
    float2 v[NBT][R];
#pragma unroll
    for (int b = 0; b < NBT; ++b) {
        const int j = threadIdx.x + b * FT;
        if (j < nb) {
#pragma unroll
            for (int r = 0; r < R; ++r) v[b][r] = ld(j + r * nb);
            if (lNs > 0) {
                float2 w1 = twlev[(1 << (lNs + LR - 1)) + (j & (Ns - 1))];
                if (INV) w1.y = -w1.y;
                if (R == 2) {
                    v[b][1] = cmul(v[b][1], w1);
                } else {
                    const float2 w2 = cmul(w1, w1), w3 = cmul(w2, w1);
                    v[b][1] = cmul(v[b][1], w1); v[b][2] = cmul(v[b][2], w2); v[b][3] = cmul(v[b][3], w3);
                    if (R >= 8) {
                        const float2 w4 = cmul(w2, w2), w5 = cmul(w4, w1), w6 = cmul(w3, w3), w7 = cmul(w4, w3);
                        v[b][4] = cmul(v[b][4], w4); v[b][5] = cmul(v[b][5], w5);
                        v[b][6] = cmul(v[b][6], w6); v[b][7] = cmul(v[b][7], w7);
                        if (R == 16) {
                            const float2 w8 = cmul(w4, w4);
                            v[b][8] = cmul(v[b][8], w8);
                            v[b][9] = cmul(v[b][9], cmul(w8, w1));
                            v[b][10] = cmul(v[b][10], cmul(w5, w5));
                            v[b][11] = cmul(v[b][11], cmul(w8, w3));
                            v[b][12] = cmul(v[b][12], cmul(w6, w6));
                            v[b][13] = cmul(v[b][13], cmul(w8, w5));
                            v[b][14] = cmul(v[b][14], cmul(w7, w7));
                            v[b][15] = cmul(v[b][15], cmul(w8, w7));
                        }
                    }
                }
            }
            Dft<R, INV>::run(v[b]);
        }
    }
    __syncthreads();
#pragma unroll
    for (int b = 0; b < NBT; ++b) {
        const int j = threadIdx.x + b * FT;
        if (j < nb) {
            const int k = j & (Ns - 1);
            const int j0 = ((j >> lNs) << (lNs + LR)) + k;
#pragma unroll
            for (int r = 0; r < R; ++r) z[padi(j0 + r * Ns)] = v[b][r];
        }
    }
    __syncthreads();
}

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
    double coef; const float *tab; float inv;
    __device__ __forceinline__ float operator()(int k) const {
        const double u = coef * (double)k;
        if (!(u == u)) return __int_as_float(0x7fc00000);
        return inv * rot_transfer_f((float)u, tab);
    }
};

// Last forward radix-2 stage (Ns = M/2), real-spectrum split, transfer function, merge, first inverse radix-2
// stage (Ns = 1) in one in-place pass over the quadruples {k, H-k, H+k, M-k}, H = M/2, 0 <= k <= M/4.
__device__ __noinline__ void quad_pass(float2 *z, int M, const RotTransfer &tf, const float2 *__restrict__ twlev) {
    const int H = M >> 1, Q = M >> 2;
    const float2 *twM = twlev + (M >> 1);          // exp(-2 pi i k / M), k < M/2
    const float2 *twN = twlev + M;                 // exp(-2 pi i k / (2M)), k < M
    constexpr int NQ = MAXM / 4 / FT;              // quadruples per thread at M = MAXM: 4
    float2 o[NQ][4];
    float2 e0 = {0.f, 0.f}, e1 = {0.f, 0.f}, e2 = {0.f, 0.f}, e3 = {0.f, 0.f};      // thread 0: k = 0 and k = M/4
#pragma unroll
    for (int b = 0; b < NQ; ++b) {
        const int k = threadIdx.x + b * FT;
        if (k > 0 && k < Q) {
            const float2 Ya = z[padi(k)], Yc = z[padi(k + H)], Yb = z[padi(H - k)], Yd = z[padi(M - k)];
            const float2 w = twM[k];
            const float2 t = cmul(w, Yc);
            float2 Za = cadd(Ya, t), Zc = csub(Ya, t);                 // Z[k], Z[k + H]
            const float2 wb = {-w.x, w.y};                             // exp(-2 pi i (H - k) / M) = -conj(w)
            const float2 t2 = cmul(wb, Yd);
            float2 Zb = cadd(Yb, t2), Zd = csub(Yb, t2);               // Z[H - k], Z[M - k]
            const float2 u = twN[k];
            real_pair(Za, Zd, u, tf(k), tf(M - k));
            const float2 ub = {-u.y, -u.x};                            // exp(-2 pi i (H - k) / (2M)) = -i conj(u)
            real_pair(Zb, Zc, ub, tf(H - k), tf(H + k));
            o[b][0] = cadd(Za, Zc); o[b][1] = csub(Za, Zc);            // out[2k], out[2k + 1]
            o[b][2] = cadd(Zb, Zd); o[b][3] = csub(Zb, Zd);            // out[M - 2k], out[M - 2k + 1]
        }
    }
    if (threadIdx.x == 0) {
        {   // k = 0: Z[0] pairs with itself, Z[H] with itself
            const float2 Y0 = z[0], YH = z[padi(H)];
            float2 Z0 = cadd(Y0, YH), ZH = csub(Y0, YH);
            const float y0 = tf(0) * (Z0.x + Z0.y), ym = tf(M) * (Z0.x - Z0.y);
            Z0 = {0.5f * (y0 + ym), 0.5f * (y0 - ym)};
            float2 dup = ZH;
            const float th = tf(H);
            real_pair(ZH, dup, twN[H], th, th);
            e0 = cadd(Z0, ZH); e1 = csub(Z0, ZH);                      // out[0], out[1]
        }
        if (Q > 0) {   // k = M/4: the pair (Q, 3Q)
            const float2 Ya = z[padi(Q)], Yc = z[padi(Q + H)];
            const float2 t = cmul(twM[Q], Yc);
            float2 Za = cadd(Ya, t), Zc = csub(Ya, t);
            real_pair(Za, Zc, twN[Q], tf(Q), tf(M - Q));
            e2 = cadd(Za, Zc); e3 = csub(Za, Zc);                      // out[2Q], out[2Q + 1]
        }
    }
    __syncthreads();
#pragma unroll
    for (int b = 0; b < NQ; ++b) {
        const int k = threadIdx.x + b * FT;
        if (k > 0 && k < Q) {
            z[padi(2 * k)] = o[b][0]; z[padi(2 * k + 1)] = o[b][1];
            z[padi(M - 2 * k)] = o[b][2]; z[padi(M - 2 * k + 1)] = o[b][3];
        }
    }
    if (threadIdx.x == 0) {
        z[0] = e0; z[padi(1)] = e1;
        if (Q > 0) { z[padi(2 * Q)] = e2; z[padi(2 * Q + 1)] = e3; }
    }
    __syncthreads();
}

// y = irfft(rfft(x) * transfer) of the n1 = 2M samples obtained by resampling the native spectrum in `fr`;
// result left in z as M complex numbers z[m] = y[2m] + i y[2m+1] (padded layout).  RM = main radix (16: three
// passes per direction at M = 8192, needs 128 registers -> one CTA per SM; 8: four passes, 64 registers -> two).
template <int RM>
__device__ __noinline__ void rotational_filter(float2 *z, const float *fr, int npix, int n1, double dlnw, double vrot,
                                               const float *rot_tab, const float2 *__restrict__ twlev) {
    constexpr int LRM = (RM == 16) ? 4 : 3;
    const int M = n1 >> 1;
    const int lM = 31 - __clz(M);
    const int lH = lM - 1;                        // log2(M/2) = LRM a + lL
    const int lL = lH % LRM;                      // leftover radix 2^lL done first (Ns = 1: no twiddles)
    LoadResampled src;
    src.fr = fr; src.ap.set((double)(npix - 1) / (double)(n1 - 1));
    const LoadZ own{z};
    int lNs = 0;
    // ---- forward: [leftover] RM ... RM, then radix 2 inside quad_pass
    if (lL == 1) { fft_pass<2, false>(z, M, 0, twlev, src); lNs = 1; }
    else if (lL == 2) { fft_pass<4, false>(z, M, 0, twlev, src); lNs = 2; }
    else if (RM == 16 && lL == 3) { fft_pass<8, false>(z, M, 0, twlev, src); lNs = 3; }
    else if (lH >= LRM) { fft_pass<RM, false>(z, M, 0, twlev, src); lNs = LRM; }
    else {                                        // M == 2: the quad pass alone is the transform
        if (threadIdx.x < M) z[padi(threadIdx.x)] = src(threadIdx.x);
        __syncthreads();
    }
    while (lNs < lH) { fft_pass<RM, false>(z, M, lNs, twlev, own); lNs += LRM; }
    // ---- last forward stage + split + transfer + merge + first inverse stage
    const double step1 = (double)(npix - 1) * dlnw / (double)(n1 - 1);
    RotTransfer tf;
    tf.tab = rot_tab; tf.inv = 1.f / (float)M;
    tf.coef = 2.0 * M_PI * sqrt(vrot * vrot) / ((double)n1 * MS_CKMS * step1);
    quad_pass(z, M, tf, twlev);
    // ---- inverse: Ns = 2 after the quad pass: [leftover] RM ... RM
    lNs = 1;
    if (lL == 1) { fft_pass<2, true>(z, M, lNs, twlev, own); lNs += 1; }
    else if (lL == 2) { fft_pass<4, true>(z, M, lNs, twlev, own); lNs += 2; }
    else if (RM == 16 && lL == 3) { fft_pass<8, true>(z, M, lNs, twlev, own); lNs += 3; }
    while (lNs < lM) { fft_pass<RM, true>(z, M, lNs, twlev, own); lNs += LRM; }
}

// ---- in-place variant (RM = 0): decimation-in-frequency forward, decimation-in-time inverse ------------------
// A Stockham pass has to hold all of a thread's outputs across a CTA barrier (16 complex numbers at 512 threads) and the
// 64-register budget of two CTAs per SM then spills ~32 registers per pass to local memory (ncu r02: 12 M local requests
// per 1536 spectra, L1 hit rate 23 %).  A convolution does not care about the order of the spectrum, so the forward
// transform here is the Gentleman-Sande in-place form (natural order in, digit-reversed out), the transfer function is
// applied at the permuted positions and the inverse is the mirrored Cooley-Tukey form (digit-reversed in, natural out):
// every butterfly writes the slots it read, one butterfly is live per thread, one barrier per pass.
// Stage list for M = 2^lM: [2^lL leftover] 8 ... 8, then a final radix 2 that is fused -- together with the first inverse
// radix 2 -- into the split / transfer / merge pass (ip_quad_pass).  X[k] ends at the position whose digits are those of
// k in reversed stage order.
struct IpPlan {
    int lM, lL, n8;               // log2 M; log2 of the leftover radix (0, 1, 2); number of radix-8 stages
    __device__ __forceinline__ void set(int lM_) { lM = lM_; lL = (lM_ - 1) % 3; n8 = (lM_ - 1) / 3; }
    __device__ __forceinline__ int freq_at(int i) const {           // frequency index stored at position i
        int rem = lM, shift = 0, k = 0;
        if (lL) { rem -= lL; k = (i >> rem) & ((1 << lL) - 1); shift = lL; }
        for (int s = 0; s < n8; ++s) { rem -= 3; k |= ((i >> rem) & 7) << shift; shift += 3; }
        return k | ((i & 1) << shift);
    }
    __device__ __forceinline__ int pos_of(int k) const {            // position of frequency index k
        int rem = lM, shift = 0, i = 0;
        if (lL) { rem -= lL; i = (k & ((1 << lL) - 1)) << rem; shift = lL; }
        for (int s = 0; s < n8; ++s) { rem -= 3; i |= ((k >> shift) & 7) << rem; shift += 3; }
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
#pragma unroll 1
    for (int t = threadIdx.x; t < nbf; t += FT) {
        int blk, j;
        if (lS >= 4) { blk = t >> lS; j = t & (S - 1); }            // consecutive threads: consecutive j
        else { blk = t & ((1 << lnblk) - 1); j = t >> lnblk; }      // short strides: consecutive blocks (padded stride L + L/16)
        const int base = (blk << lL_) + j;
        float2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = ld(base + (r << lS));
        if (!INV) Dft<R, false>::run(v);
        if (lS > 0) {
            float2 w1 = tw[j];
            if (INV) w1.y = -w1.y;
            v[1] = cmul(v[1], w1);
            if (R >= 4) {
                const float2 w2 = cmul(w1, w1), w3 = cmul(w2, w1);
                v[2] = cmul(v[2], w2); v[3] = cmul(v[3], w3);
                if (R >= 8) {
                    const float2 w4 = cmul(w2, w2);
                    v[4] = cmul(v[4], w4); v[5] = cmul(v[5], cmul(w4, w1));
                    v[6] = cmul(v[6], cmul(w3, w3)); v[7] = cmul(v[7], cmul(w4, w3));
                }
            }
        }
        if (INV) Dft<R, true>::run(v);
#pragma unroll
        for (int r = 0; r < R; ++r) z[padi(base + (r << lS))] = v[r];
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
    const int lrp = pl.n8 ? 3 : pl.lL;             // log2 radix of the last stage before the fused radix 2
    const int lh = lrp - 1;
#pragma unroll 1
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

// same contract as rotational_filter<RM>: z[m] = y[2m] + i y[2m+1] in natural order
__device__ __noinline__ void rotational_filter_ip(float2 *z, const float *fr, int npix, int n1, double dlnw, double vrot,
                                                  const float *rot_tab, const float2 *__restrict__ twlev) {
    const int M = n1 >> 1;
    IpPlan pl;
    pl.set(31 - __clz(M));
    LoadResampled src;
    src.fr = fr; src.ap.set((double)(npix - 1) / (double)(n1 - 1));
    const LoadZ own{z};
    int l = pl.lM;
    bool first = true;
    if (pl.lL == 1) { ip_stage<1, false>(z, pl.lM, l, twlev, src); l -= 1; first = false; }
    else if (pl.lL == 2) { ip_stage<2, false>(z, pl.lM, l, twlev, src); l -= 2; first = false; }
    for (int s = 0; s < pl.n8; ++s) {
        if (first) ip_stage<3, false>(z, pl.lM, l, twlev, src);
        else ip_stage<3, false>(z, pl.lM, l, twlev, own);
        first = false; l -= 3;
    }
    const double step1 = (double)(npix - 1) * dlnw / (double)(n1 - 1);
    RotTransfer tf;
    tf.tab = rot_tab; tf.inv = 1.f / (float)M;
    tf.coef = 2.0 * M_PI * sqrt(vrot * vrot) / ((double)n1 * MS_CKMS * step1);
    ip_quad_pass(z, pl, tf, twlev);
    l = 1;
    for (int s = 0; s < pl.n8; ++s) { l += 3; ip_stage<3, true>(z, pl.lM, l, twlev, own); }
    if (pl.lL == 1) ip_stage<1, true>(z, pl.lM, l + 1, twlev, own);
    else if (pl.lL == 2) ip_stage<2, true>(z, pl.lM, l + 2, twlev, own);
}

__device__ __forceinline__ float conv_at(const float2 *z, int k) {          // real sample k of the filtered sequence
    const float2 c = z[padi(k >> 1)];
    return (k & 1) ? c.y : c.x;
}

// per-point decisions, made once by thread 0 and broadcast through shared memory
struct PointSetup {
    int mode;             // 0: Gaussian FIR on the n2 grid; 1: plain interpolation; 2: all NaN; 3: fallback; 4: skip row
    int i_s, i_e, n2, ntap, nobs, rot;
    double d2, ratio2, spx, lnscale, lnw_s, inst, vrot;
};

__device__ __forceinline__ void mbar_init_(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx_(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s_(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// RM: main FFT radix; PF: the native spectrum of the CTA's next point is fetched by a bulk async copy (TMA) into a
// second buffer while the current point is processed (one CTA per SM: nothing else hides that latency).
template <int RM, bool PF>
__global__ void __launch_bounds__(FT, (RM == 16 && FT == 512) ? 1 : 2)
spec_fast_kernel(SmoothArgs a, int nB, int *__restrict__ fb_rows, int *__restrict__ fb_count) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int npix = a.npix, n1 = a.n1, M = n1 >> 1;
    float2 *z = reinterpret_cast<float2 *>(smem_raw);                        // padded complex buffer
    float *s2 = reinterpret_cast<float *>(smem_raw);                         // later: the n2-point instrumental grid
    const size_t fr_off = (size_t)padi(M) * sizeof(float2) + 16;
    const size_t fr_bytes = ((size_t)(npix + 1) * sizeof(float) + 15) & ~(size_t)15;
    __shared__ __align__(16) float wtab[5 * WSTRIDE];                        // FIR weights for the 4 window alignments (+1)
    __shared__ float gtap[MAXTAP + 2];
    __shared__ double red[FT / 32];
    __shared__ PointSetup ps;
    __shared__ __align__(8) unsigned long long bar_store[2];
    const float2 *twlev = static_cast<const float2 *>(a.twlev);
    const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(&bar_store[0]);

    if (PF) {
        if (threadIdx.x == 0) {
            mbar_init_(bar0, 1); mbar_init_(bar0 + 8, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            if ((int)blockIdx.x < nB) {                                      // first point's spectrum
                mbar_expect_tx_(bar0, (uint32_t)(npix * sizeof(float)));
                bulk_g2s_((uint32_t)__cvta_generic_to_shared(smem_raw + fr_off),
                          static_cast<const float *>(a.flux_native) + (size_t)blockIdx.x * npix,
                          (uint32_t)(npix * sizeof(float)), bar0);
            }
        }
        __syncthreads();
    }

    int it = 0;
    for (int p = blockIdx.x; p < nB; p += gridDim.x, ++it) {
        const int buf = PF ? (it & 1) : 0;
        float *fr = reinterpret_cast<float *>(smem_raw + fr_off + (size_t)buf * fr_bytes);   // native / rotated spectrum
        const double *r = a.pars + (size_t)p * MS_NPARS;
        const int slot = a.star_slot ? a.star_slot[p] : a.slot0;
        const bool slot_ok = (unsigned)slot < (unsigned)a.nstars;
        const StarView sv = a.stars[slot_ok ? slot : 0];
        if (PF && threadIdx.x == 0) {
            // all threads finished with buffer buf ^ 1 at the end of the previous iteration (block barrier there)
            const int pn = p + gridDim.x;
            if (pn < nB) {
                const uint32_t b = bar0 + 8 * (buf ^ 1);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic-proxy accesses of that buffer are done
                mbar_expect_tx_(b, (uint32_t)(npix * sizeof(float)));
                bulk_g2s_((uint32_t)__cvta_generic_to_shared(smem_raw + fr_off + (size_t)(buf ^ 1) * fr_bytes),
                          static_cast<const float *>(a.flux_native) + (size_t)pn * npix,
                          (uint32_t)(npix * sizeof(float)), b);
            }
        }
        if (threadIdx.x == 0) {
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
            if (q.mode == 3) fb_rows[atomicAdd(fb_count, 1)] = p;
            ps = q;
        }
        __syncthreads();
        const int mode = ps.mode, nobs = ps.nobs;
        if (mode >= 3) {                                                      // uniform: skipped or handed to the generic kernel
            if (mode == 4) {
                if (a.flux_out && slot_ok)
                    for (int j = threadIdx.x; j < nobs; j += FT) a.flux_out[(size_t)p * nobs + j] = ms_nan();
                if (a.chi2 && !slot_ok && threadIdx.x == 0 && r[PV_TEFF] != ms_inf()) a.chi2[p] = ms_nan();
            }
            if (PF) mbar_wait_(bar0 + 8 * buf, (it >> 1) & 1);               // consume the phase of this buffer
            __syncthreads();
            continue;
        }
        const int i_s = ps.i_s, i_e = ps.i_e, n2 = ps.n2, ntap = ps.ntap;
        const double d2 = ps.d2, inst = ps.inst;
        // ---- native spectrum -> shared memory
        if (PF) {
            mbar_wait_(bar0 + 8 * buf, (it >> 1) & 1);
        } else {
            const float *frow = static_cast<const float *>(a.flux_native) + (size_t)p * npix;
            if ((npix & 3) == 0) {
                const float4 *src4 = reinterpret_cast<const float4 *>(frow);
                float4 *dst4 = reinterpret_cast<float4 *>(fr);
                for (int i = threadIdx.x; i < (npix >> 2); i += FT) dst4[i] = __ldcs(src4 + i);
            } else {
                for (int i = threadIdx.x; i < npix; i += FT) fr[i] = __ldcs(frow + i);
            }
        }
        if (threadIdx.x == FT - 1)                                            // one-sample pad: interpolation at the last pixel
            fr[npix] = PF ? fr[npix - 1] : __ldcs(static_cast<const float *>(a.flux_native) + (size_t)p * npix + npix - 1);
        if (mode == 0 && threadIdx.x <= ntap) {                               // Gaussian taps, t = 0 .. ntap
            const double t = (double)threadIdx.x / ps.spx;
            gtap[threadIdx.x] = (float)(exp(-0.5 * t * t) / (ps.spx * 2.5066282746310002));
        }
        __syncthreads();
        // FIR weight rows: W[o][u] = g[u - o], g[s] = tap (ntap - s) for s = 0 .. 2 ntap, zero elsewhere.  An observed
        // pixel whose window starts o samples after a 16-byte boundary uses rows o (sample i) and o + 1 (sample i + 1).
        const int U = (2 * ntap + 2 + 3 + 3) & ~3;
        if (mode == 0) {
            for (int idx = threadIdx.x; idx < 5 * U; idx += FT) {
                const int o = idx / U, u = idx - o * U;
                const int sidx = u - o;
                float w = 0.f;
                if (sidx >= 0 && sidx <= 2 * ntap) { const int t = ntap - sidx; w = gtap[t < 0 ? -t : t]; }
                wtab[o * WSTRIDE + u] = w;
            }
        }

        // ---- rotational broadening on the full native grid (outwave = None), back onto the native pixels
        if (ps.rot) {
            if constexpr (RM == 0) rotational_filter_ip(z, fr, npix, n1, a.dlnw, ps.vrot, a.rot_tab, twlev);
            else rotational_filter<RM>(z, fr, npix, n1, a.dlnw, ps.vrot, a.rot_tab, twlev);
            // smoothing.py:262: interpolation of the filtered n1 grid at the native wavelengths
            AffinePos ap;
            ap.set((double)(n1 - 1) / (double)(npix - 1));
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
                const double q = (sv.obs_lnwave[j] - lnscale - lnw_s) / d2;
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
            const int nch = U >> 2;
            for (int j = threadIdx.x; j < nobs; j += 2 * FT) {
                const int j2 = j + FT;
                const bool two = j2 < nobs;
                int iA, iB = 0; double fA, fB = 0.0;
                locate(j, iA, fA);
                if (two) locate(j2, iB, fB);
                const int bA = (iA - ntap) & ~3, oA = (iA - ntap) - bA;
                const int bB = (iB - ntap) & ~3, oB = (iB - ntap) - bB;
                const bool fastA = bA >= 0 && bA + U <= n2, fastB = two && bB >= 0 && bB + U <= n2;
                float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
                if (fastA && fastB) {
                    const float4 *xa = reinterpret_cast<const float4 *>(s2 + bA), *xb = reinterpret_cast<const float4 *>(s2 + bB);
                    const float4 *wa0 = reinterpret_cast<const float4 *>(wtab + oA * WSTRIDE), *wa1 = wa0 + WSTRIDE / 4;
                    const float4 *wb0 = reinterpret_cast<const float4 *>(wtab + oB * WSTRIDE), *wb1 = wb0 + WSTRIDE / 4;
                    float a0b = 0.f, a1b = 0.f, b0b = 0.f, b1b = 0.f;
#pragma unroll 2
                    for (int c = 0; c < nch; ++c) {
                        const float4 x = xa[c], g0 = wa0[c], g1 = wa1[c], y = xb[c], h0 = wb0[c], h1 = wb1[c];
                        a0 = fmaf(g0.x, x.x, a0); a0b = fmaf(g0.y, x.y, a0b); a0 = fmaf(g0.z, x.z, a0); a0b = fmaf(g0.w, x.w, a0b);
                        a1 = fmaf(g1.x, x.x, a1); a1b = fmaf(g1.y, x.y, a1b); a1 = fmaf(g1.z, x.z, a1); a1b = fmaf(g1.w, x.w, a1b);
                        b0 = fmaf(h0.x, y.x, b0); b0b = fmaf(h0.y, y.y, b0b); b0 = fmaf(h0.z, y.z, b0); b0b = fmaf(h0.w, y.w, b0b);
                        b1 = fmaf(h1.x, y.x, b1); b1b = fmaf(h1.y, y.y, b1b); b1 = fmaf(h1.z, y.z, b1); b1b = fmaf(h1.w, y.w, b1b);
                    }
                    a0 += a0b; a1 += a1b; b0 += b0b; b1 += b1b;
                } else {
                    if (fastA) {
                        const float4 *xa = reinterpret_cast<const float4 *>(s2 + bA);
                        const float4 *wa0 = reinterpret_cast<const float4 *>(wtab + oA * WSTRIDE), *wa1 = wa0 + WSTRIDE / 4;
                        for (int c = 0; c < nch; ++c) {
                            const float4 x = xa[c], g0 = wa0[c], g1 = wa1[c];
                            a0 = fmaf(g0.x, x.x, a0); a0 = fmaf(g0.y, x.y, a0); a0 = fmaf(g0.z, x.z, a0); a0 = fmaf(g0.w, x.w, a0);
                            a1 = fmaf(g1.x, x.x, a1); a1 = fmaf(g1.y, x.y, a1); a1 = fmaf(g1.z, x.z, a1); a1 = fmaf(g1.w, x.w, a1);
                        }
                    } else fir_slow(iA, a0, a1);
                    if (two) {
                        if (fastB) {
                            const float4 *xb = reinterpret_cast<const float4 *>(s2 + bB);
                            const float4 *wb0 = reinterpret_cast<const float4 *>(wtab + oB * WSTRIDE), *wb1 = wb0 + WSTRIDE / 4;
                            for (int c = 0; c < nch; ++c) {
                                const float4 y = xb[c], h0 = wb0[c], h1 = wb1[c];
                                b0 = fmaf(h0.x, y.x, b0); b0 = fmaf(h0.y, y.y, b0); b0 = fmaf(h0.z, y.z, b0); b0 = fmaf(h0.w, y.w, b0);
                                b1 = fmaf(h1.x, y.x, b1); b1 = fmaf(h1.y, y.y, b1); b1 = fmaf(h1.z, y.z, b1); b1 = fmaf(h1.w, y.w, b1);
                            }
                        } else fir_slow(iB, b0, b1);
                    }
                }
                finish(j, (double)a0 + ((double)a1 - (double)a0) * fA);
                if (two) finish(j2, (double)b0 + ((double)b1 - (double)b0) * fB);
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

size_t fast_smem_bytes(int npix, int n1, bool pf) {
    const int M = n1 >> 1;
    const size_t fr_bytes = ((size_t)(npix + 1) * sizeof(float) + 15) & ~(size_t)15;
    return (size_t)(M + (M >> 4)) * sizeof(float2) + 16 + (pf ? 2 : 1) * fr_bytes + 16;
}

}  // namespace

bool spec_fast_usable(const ms_ctx *ctx) {
    const SpecNetDev &n = ctx->spec;
    return ctx->precision != MS_PREC_F64 && n.n1 >= 8 && (n.n1 >> 1) <= MAXM &&
           fast_smem_bytes(n.npix, n.n1, false) <= 226 * 1024;
}

// Rows the kernel declines are appended to fb_rows / *fb_count (device; the caller zeroes the counter).
// ctx->spec_fast selects the variant: 1 = radix 16, one CTA per SM, bulk-copy prefetch of the next spectrum;
// 2 = radix 8 Stockham, two CTAs per SM (when two fit); 3 = in-place radix 8 (DIF / DIT), two CTAs per SM.
int launch_spec_fast(ms_ctx *ctx, const SmoothArgs &sa, int nB, int *fb_rows, int *fb_count, cudaStream_t st) {
    const bool aligned = (sa.npix & 3) == 0;
    const size_t smem_pf = fast_smem_bytes(sa.npix, sa.n1, true), smem_1 = fast_smem_bytes(sa.npix, sa.n1, false);
    const bool fits2 = 2 * (smem_1 + 1024 + 4096) <= 227 * 1024;
    const bool inplace = ctx->spec_fast >= 3;
    const bool two = (inplace || ctx->spec_fast == 2 || FT == 256) && fits2;
    const bool pf = !inplace && !two && aligned && smem_pf <= 226 * 1024;
    const size_t smem = pf ? smem_pf : smem_1;
    auto k16 = spec_fast_kernel<16, false>;
    auto k16p = spec_fast_kernel<16, true>;
    auto k8 = spec_fast_kernel<8, false>;
    auto kip = spec_fast_kernel<0, false>;
    void (*kern)(SmoothArgs, int, int *, int *);
    size_t *cfg;
    static size_t cfg16 = 0, cfg16p = 0, cfg8 = 0, cfgip = 0;
    if (inplace) { kern = kip; cfg = &cfgip; }
    else if (two && FT == 256 && ctx->spec_fast != 2) { kern = k16; cfg = &cfg16; }
    else if (two) { kern = k8; cfg = &cfg8; }
    else if (pf) { kern = k16p; cfg = &cfg16p; }
    else { kern = k16; cfg = &cfg16; }
    if (*cfg < smem) { MS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); *cfg = smem; }
    int grid = ctx->sm_count * (two ? 2 : 1);
    if (grid > nB) grid = nB;
    kern<<<grid, FT, smem, st>>>(sa, nB, fb_rows, fb_count);
    ctx->launches++;
    MS_CUDA(cudaGetLastError());
    return MS_OK;
}
