// Kernel (3): rotational + instrumental broadening, resampling onto the observed
// pixels, blaze polynomial and the spectral chi-square, one CTA per point with
// the whole spectrum resident in shared memory.  Batched form of
//   PayneSpecPredict.getspec broadening chain   external Payne (SURVEY.md Appendix B,
//                                               PARITY UNPINNED; call site genmod.py:80-84)
//   smoothspec 'vsini' / 'R'    reference minesweeper/smoothing.py:92-114, 131-134
//   smooth_vsini_fft / smooth_vel_fft            smoothing.py:242-264, 202-240
//   smooth_fft / smooth_fft_vsini                smoothing.py:517-558
//   mask_wave / resample_wave                    smoothing.py:560-598
//   polycalc                                     reference minesweeper/fitutils.py:11-20
//   likelihood.lnlike spectral chi-square        reference minesweeper/likelihood.py:188-195
//
// The reference convolves by FFT on a power-of-two log-uniform resampling of the
// spectrum; its result is defined by that (circular, band-limited) construction,
// so this kernel performs the same steps: linear resampling, a real FFT of the
// resampled spectrum (as a half-length complex FFT held in shared memory,
// decimation-in-frequency forward / decimation-in-time inverse so that no
// bit-reversal pass is needed), multiplication by the analytic transfer function,
// inverse FFT, linear interpolation back.  Because the native wavelength grid is
// log-uniform and a Doppler shift is a translation in ln(lambda), every
// resampling position is an affine function of the pixel index; no per-pixel
// logarithm or exponential is evaluated.
#include "spec_args.cuh"

// one CTA per SM; 32 warps in f32 (64 registers), 16 in f64 (the f64 kernel needs > 64 registers)
template <typename T> struct SmoothCfg { static constexpr int THREADS = sizeof(T) == 4 ? 1024 : 512; };

namespace {

template <typename T> struct Cx;
template <> struct Cx<float> { typedef float2 type; };
template <> struct Cx<double> { typedef double2 type; };

template <typename T2> __device__ __forceinline__ T2 cadd(T2 a, T2 b) { return {a.x + b.x, a.y + b.y}; }
template <typename T2> __device__ __forceinline__ T2 csub(T2 a, T2 b) { return {a.x - b.x, a.y - b.y}; }
template <typename T2> __device__ __forceinline__ T2 cmul(T2 a, T2 b) {
    return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}
template <typename T2> __device__ __forceinline__ T2 cmulc(T2 a, T2 b) {   // a * conj(b)
    return {a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y};
}

// one-pass data (flux row, resampling maps) is streamed past L1 so that the tables every
// point re-reads (FFT twiddles, rotational transfer) stay L1-resident
__device__ __forceinline__ float ld_stream(const float *p) {
    float v; asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v;
}
__device__ __forceinline__ double ld_stream(const double *p) {
    double v; asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v;
}
__device__ __forceinline__ int ld_stream(const int *p) {
    int v; asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p)); return v;
}

// Transfer function of the rotational profile, smoothing.py:541-548.
__device__ __forceinline__ double rot_transfer(double u) {
    if (u == 0.0) return 1.0;
    double s, c;
    sincos(u, &s, &c);
    const double u2 = u * u;
    return j1(u) / u - 3.0 * c / (2.0 * u2) + 3.0 * s / (2.0 * u2 * u);
}
template <typename T> struct Transfer {
    int kind;            // 0 gaussian, 1 rotation
    double coef;         // gaussian: 2 pi^2 sigma^2 / (n dv)^2 ; rotation: 2 pi vsini / (n dv)
    const float *tab;
    __device__ __forceinline__ T operator()(int k) const;
};
template <> __device__ __forceinline__ double Transfer<double>::operator()(int k) const {
    if (kind == 0) return exp(-coef * (double)k * (double)k);
    return rot_transfer(coef * (double)k);
}
template <> __device__ __forceinline__ float Transfer<float>::operator()(int k) const {
    if (kind == 0) return expf(-(float)coef * (float)k * (float)k);
    const double u = coef * (double)k;
    if (!(u == u)) return __int_as_float(0x7fc00000);
    return rot_transfer_f((float)u, tab);
}

// ---- in-place FFT passes in shared memory ------------------------------------------------
// The transform is the textbook radix-2 decimation-in-frequency flow (natural order in,
// bit-reversed order out) and its decimation-in-time mirror for the inverse, but three
// consecutive radix-2 stages are fused into one pass: a thread loads the eight elements that
// interact in stages with half-lengths 4q, 2q, q, does the three stages in registers and
// stores them back, so a 8192-point transform makes 5 shared-memory passes instead of 13.
// Twiddles come from the table tw[k] = exp(-2 pi i k / tw_n).
template <typename T2> __device__ __forceinline__ void bfly_dif(T2 &a, T2 &b, const T2 w) {
    const T2 d = csub(a, b);
    a = cadd(a, b);
    b = cmul(d, w);
}
template <typename T2> __device__ __forceinline__ void bfly_dit(T2 &a, T2 &b, const T2 w) {   // w conjugated here
    const T2 t = cmulc(b, w);
    b = csub(a, t);
    a = cadd(a, t);
}
template <typename T2> __device__ __forceinline__ T2 mul_w8(const T2 t, int m) {   // t * exp(-2 pi i m / 8), m = 1..3
    typedef decltype(t.x) T;
    const T r = (T)0.70710678118654752440;
    if (m == 1) return {r * (t.x + t.y), r * (t.y - t.x)};
    if (m == 2) return {t.y, -t.x};
    return {r * (t.y - t.x), -r * (t.x + t.y)};
}

template <typename T, bool INVERSE>
__device__ __forceinline__ void fft_pass8(typename Cx<T>::type *z, int M, int q, int lq,
                                          const typename Cx<T>::type *__restrict__ tw, int tw_n) {
    typedef typename Cx<T>::type T2;
    for (int t = threadIdx.x; t < (M >> 3); t += blockDim.x) {
        const int posq = t & (q - 1);
        const int base = ((t >> lq) << (lq + 3)) + posq;
        const T2 t1 = tw[posq * (tw_n >> (lq + 3))];      // W_{8q}^{posq}
        const T2 t2 = tw[posq * (tw_n >> (lq + 2))];      // W_{4q}^{posq}
        const T2 t3 = tw[posq * (tw_n >> (lq + 1))];      // W_{2q}^{posq}
        T2 x[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) x[m] = z[base + m * q];
        const T2 wa1 = mul_w8(t1, 1), wa2 = mul_w8(t1, 2), wa3 = mul_w8(t1, 3);
        const T2 wb1 = mul_w8(t2, 2);
        if (!INVERSE) {
            bfly_dif(x[0], x[4], t1); bfly_dif(x[1], x[5], wa1); bfly_dif(x[2], x[6], wa2); bfly_dif(x[3], x[7], wa3);
            bfly_dif(x[0], x[2], t2); bfly_dif(x[1], x[3], wb1); bfly_dif(x[4], x[6], t2); bfly_dif(x[5], x[7], wb1);
            bfly_dif(x[0], x[1], t3); bfly_dif(x[2], x[3], t3); bfly_dif(x[4], x[5], t3); bfly_dif(x[6], x[7], t3);
        } else {
            bfly_dit(x[0], x[1], t3); bfly_dit(x[2], x[3], t3); bfly_dit(x[4], x[5], t3); bfly_dit(x[6], x[7], t3);
            bfly_dit(x[0], x[2], t2); bfly_dit(x[1], x[3], wb1); bfly_dit(x[4], x[6], t2); bfly_dit(x[5], x[7], wb1);
            bfly_dit(x[0], x[4], t1); bfly_dit(x[1], x[5], wa1); bfly_dit(x[2], x[6], wa2); bfly_dit(x[3], x[7], wa3);
        }
#pragma unroll
        for (int m = 0; m < 8; ++m) z[base + m * q] = x[m];
    }
    __syncthreads();
}

// one plain radix-2 stage (used for the 1 or 2 stages left over when log2 M is not a multiple of 3)
template <typename T, bool INVERSE>
__device__ __forceinline__ void fft_pass2(typename Cx<T>::type *z, int M, int s,
                                          const typename Cx<T>::type *__restrict__ tw, int tw_n) {
    typedef typename Cx<T>::type T2;
    const int half = 1 << s, tstep = tw_n >> (s + 1);
    for (int t = threadIdx.x; t < (M >> 1); t += blockDim.x) {
        const int pos = t & (half - 1);
        const int i0 = ((t >> s) << (s + 1)) + pos;
        T2 a = z[i0], b = z[i0 + half];
        if (!INVERSE) bfly_dif(a, b, tw[pos * tstep]); else bfly_dit(a, b, tw[pos * tstep]);
        z[i0] = a; z[i0 + half] = b;
    }
    __syncthreads();
}

// y = irfft(rfft(x) * transfer) for the n real samples stored in z (z[m] = x[2m] + i x[2m+1]).
template <typename T>
__device__ void fft_filter(typename Cx<T>::type *z, int n, const Transfer<T> &tf,
                           const typename Cx<T>::type *__restrict__ tw, int tw_n) {
    typedef typename Cx<T>::type T2;
    const int M = n >> 1;
    const int L = 31 - __clz(M);
    const int rem = L % 3;
    // forward: stages L-1 ... 0
    for (int s = L - 1; s >= rem + 2; s -= 3) fft_pass8<T, false>(z, M, 1 << (s - 2), s - 2, tw, tw_n);
    for (int s = rem - 1; s >= 0; --s) fft_pass2<T, false>(z, M, s, tw, tw_n);
    // split into the spectrum of the real sequence, apply the transfer (with the 1/M of the
    // inverse transform folded in), merge back
    const int nbf = M >> 1;
    const int tstepn = tw_n / n;
    const T inv = (T)1 / (T)M;
    for (int k = threadIdx.x; k <= nbf; k += blockDim.x) {
        if (k == 0) {
            const T2 z0 = z[0];
            const T y0 = inv * tf(0) * (z0.x + z0.y), ym = inv * tf(M) * (z0.x - z0.y);
            z[0] = {(T)0.5 * (y0 + ym), (T)0.5 * (y0 - ym)};
        } else {
            const int pk = __brev((unsigned)k) >> (32 - L);
            const int pm = __brev((unsigned)(M - k)) >> (32 - L);
            const T2 A = z[pk];
            T2 Bc = z[pm]; Bc.y = -Bc.y;
            const T2 E = {(T)0.5 * (A.x + Bc.x), (T)0.5 * (A.y + Bc.y)};
            const T2 dif = csub(A, Bc);
            const T2 O = {(T)0.5 * dif.y, (T)-0.5 * dif.x};          // -i/2 (A - Bc)
            const T2 w = tw[k * tstepn];
            const T2 wO = cmul(w, O);
            const T tk = inv * tf(k), tm = inv * tf(M - k);
            const T2 Yk = {tk * (E.x + wO.x), tk * (E.y + wO.y)};
            const T2 Ymc = {tm * (E.x - wO.x), tm * (E.y - wO.y)};   // conj(Y[M-k])
            const T2 Ep = {(T)0.5 * (Yk.x + Ymc.x), (T)0.5 * (Yk.y + Ymc.y)};
            const T2 D = {(T)0.5 * (Yk.x - Ymc.x), (T)0.5 * (Yk.y - Ymc.y)};
            const T2 Op = cmulc(D, w);
            z[pk] = {Ep.x - Op.y, Ep.y + Op.x};                      // Ep + i Op
            if (pm != pk) z[pm] = {Ep.x + Op.y, Op.x - Ep.y};        // conj(Ep) + i conj(Op)
        }
    }
    __syncthreads();
    // inverse: stages 0 ... L-1
    for (int s = 0; s < rem; ++s) fft_pass2<T, true>(z, M, s, tw, tw_n);
    for (int s = rem + 2; s <= L - 1; s += 3) fft_pass8<T, true>(z, M, 1 << (s - 2), s - 2, tw, tw_n);
}

// ---- Stockham autosort FFT (f32 path) ---------------------------------------------------
// Out-of-place passes between two shared-memory buffers: a thread reads R inputs at stride
// M/R (consecutive threads -> consecutive addresses, conflict-free), applies the pass
// twiddles, does an R-point transform in registers and writes at stride Ns.  Output is in
// natural order, so the split/merge pass walks k and M-k linearly (conflict-free, and its
// twiddle / transfer-table reads are coalesced).  Pass schedule: radix 8 while >= 3 bits
// remain, then radix 4 or 2 (8192 = 8.8.8.8.2).
template <typename T2, bool INV> __device__ __forceinline__ T2 tw_get(const T2 *__restrict__ tw, int idx) {
    T2 w = tw[idx];
    if (INV) w.y = -w.y;
    return w;
}
template <typename T2, bool INV> __device__ __forceinline__ void fft8_reg(T2 (&v)[8]) {
    typedef decltype(v[0].x) T;
    const T r = (T)0.70710678118654752440;
    // stage 1: (m, m+4), twiddle W8^m on the difference
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const T2 a = v[m], b = v[m + 4];
        v[m] = cadd(a, b);
        const T2 d = csub(a, b);
        T2 o;
        if (m == 0) o = d;
        else if (m == 1) o = INV ? T2{r * (d.x - d.y), r * (d.x + d.y)} : T2{r * (d.x + d.y), r * (d.y - d.x)};
        else if (m == 2) o = INV ? T2{-d.y, d.x} : T2{d.y, -d.x};
        else o = INV ? T2{-r * (d.x + d.y), r * (d.x - d.y)} : T2{r * (d.y - d.x), -r * (d.x + d.y)};
        v[m + 4] = o;
    }
    // stage 2: (m, m+2) inside each half, twiddle W4^m
#pragma unroll
    for (int h = 0; h < 8; h += 4) {
        const T2 a0 = v[h], b0 = v[h + 2], a1 = v[h + 1], b1 = v[h + 3];
        v[h] = cadd(a0, b0); v[h + 2] = csub(a0, b0);
        v[h + 1] = cadd(a1, b1);
        const T2 d = csub(a1, b1);
        v[h + 3] = INV ? T2{-d.y, d.x} : T2{d.y, -d.x};
    }
    // stage 3: (m, m+1)
#pragma unroll
    for (int m = 0; m < 8; m += 2) {
        const T2 a = v[m], b = v[m + 1];
        v[m] = cadd(a, b); v[m + 1] = csub(a, b);
    }
    // v[i] now holds X[bitrev3(i)]
}

template <typename T, int R, bool INV>
__device__ __forceinline__ void stockham_pass(const typename Cx<T>::type *__restrict__ in,
                                              typename Cx<T>::type *__restrict__ out, int M, int lNs,
                                              const typename Cx<T>::type *__restrict__ tw, int tw_n) {
    typedef typename Cx<T>::type T2;
    constexpr int LR = (R == 8) ? 3 : (R == 4 ? 2 : 1);
    const int nb = M >> LR, Ns = 1 << lNs;
    for (int j = threadIdx.x; j < nb; j += blockDim.x) {
        const int k = j & (Ns - 1);
        T2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = in[j + r * nb];
        if (lNs > 0) {
            const T2 w1 = tw_get<T2, INV>(tw, (1 << (lNs + LR - 1)) + k);      // level table: exp(-+2 pi i k / (Ns R))
            if (R == 2) {
                v[1] = cmul(v[1], w1);
            } else {
                const T2 w2 = cmul(w1, w1), w3 = cmul(w2, w1);
                v[1] = cmul(v[1], w1); v[2] = cmul(v[2], w2); v[3] = cmul(v[3], w3);
                if (R == 8) {
                    const T2 w4 = cmul(w2, w2), w5 = cmul(w4, w1), w6 = cmul(w3, w3), w7 = cmul(w6, w1);
                    v[4] = cmul(v[4], w4); v[5] = cmul(v[5], w5); v[6] = cmul(v[6], w6); v[7] = cmul(v[7], w7);
                }
            }
        }
        const int j0 = ((j >> lNs) << (lNs + LR)) + k;
        if (R == 8) {
            T2 (&v8)[8] = reinterpret_cast<T2 (&)[8]>(v);
            fft8_reg<T2, INV>(v8);
            out[j0] = v[0]; out[j0 + Ns] = v[4]; out[j0 + 2 * Ns] = v[2]; out[j0 + 3 * Ns] = v[6];
            out[j0 + 4 * Ns] = v[1]; out[j0 + 5 * Ns] = v[5]; out[j0 + 6 * Ns] = v[3]; out[j0 + 7 * Ns] = v[7];
        } else if (R == 4) {
            const T2 a0 = cadd(v[0], v[2]), a1 = csub(v[0], v[2]), b0 = cadd(v[1], v[3]), d = csub(v[1], v[3]);
            const T2 b1 = INV ? T2{-d.y, d.x} : T2{d.y, -d.x};
            out[j0] = cadd(a0, b0); out[j0 + Ns] = cadd(a1, b1); out[j0 + 2 * Ns] = csub(a0, b0); out[j0 + 3 * Ns] = csub(a1, b1);
        } else {
            out[j0] = cadd(v[0], v[1]); out[j0 + Ns] = csub(v[0], v[1]);
        }
    }
    __syncthreads();
}

// runs all passes of an M-point transform from buffer a to ... ; returns which buffer holds the result
template <typename T, bool INV>
__device__ __forceinline__ int stockham_fft(typename Cx<T>::type *b0, typename Cx<T>::type *b1, int src, int M, int L,
                                            const typename Cx<T>::type *__restrict__ tw, int tw_n) {
    typedef typename Cx<T>::type T2;
    int lNs = 0, rem = L;
    // leftover bits first (Ns = 1: no twiddles), so the last and largest radix-8 pass needs only
    // M / 8 distinct twiddles
    if (rem % 3 == 2) {
        stockham_pass<T, 4, INV>(src ? b1 : b0, src ? b0 : b1, M, lNs, tw, tw_n); lNs += 2; rem -= 2; src ^= 1;
    } else if (rem % 3 == 1) {
        stockham_pass<T, 2, INV>(src ? b1 : b0, src ? b0 : b1, M, lNs, tw, tw_n); lNs += 1; rem -= 1; src ^= 1;
    }
    while (rem > 0) {
        stockham_pass<T, 8, INV>(src ? b1 : b0, src ? b0 : b1, M, lNs, tw, tw_n); lNs += 3; rem -= 3; src ^= 1;
    }
    return src;
}

// y = irfft(rfft(x) * transfer): input reals in z0 (z0[m] = x[2m] + i x[2m+1]), result back in z0;
// z1 is the ping-pong buffer.  Natural-order data throughout.
template <typename T>
__device__ void fft_filter_stockham(typename Cx<T>::type *z0, typename Cx<T>::type *z1, int n, const Transfer<T> &tf,
                                    const typename Cx<T>::type *__restrict__ tw, int tw_n) {
    typedef typename Cx<T>::type T2;
    const int M = n >> 1;
    const int L = 31 - __clz(M);
    const int where = stockham_fft<T, false>(z0, z1, 0, M, L, tw, tw_n);
    T2 *z = where ? z1 : z0;
    const int nbf = M >> 1;
    const T2 *twn = tw + (n >> 1);                            // level log2(n): exp(-2 pi i k / n), k < n / 2
    const T inv = (T)1 / (T)M;
    for (int k = threadIdx.x; k <= nbf; k += blockDim.x) {
        if (k == 0) {
            const T2 zz = z[0];
            const T y0 = inv * tf(0) * (zz.x + zz.y), ym = inv * tf(M) * (zz.x - zz.y);
            z[0] = {(T)0.5 * (y0 + ym), (T)0.5 * (y0 - ym)};
        } else {
            const int pm = M - k;
            const T2 A = z[k];
            T2 Bc = z[pm]; Bc.y = -Bc.y;
            const T2 E = {(T)0.5 * (A.x + Bc.x), (T)0.5 * (A.y + Bc.y)};
            const T2 dif = csub(A, Bc);
            const T2 O = {(T)0.5 * dif.y, (T)-0.5 * dif.x};
            const T2 w = twn[k];
            const T2 wO = cmul(w, O);
            const T tk = inv * tf(k), tm = inv * tf(pm);
            const T2 Yk = {tk * (E.x + wO.x), tk * (E.y + wO.y)};
            const T2 Ymc = {tm * (E.x - wO.x), tm * (E.y - wO.y)};
            const T2 Ep = {(T)0.5 * (Yk.x + Ymc.x), (T)0.5 * (Yk.y + Ymc.y)};
            const T2 D = {(T)0.5 * (Yk.x - Ymc.x), (T)0.5 * (Yk.y - Ymc.y)};
            const T2 Op = cmulc(D, w);
            z[k] = {Ep.x - Op.y, Ep.y + Op.x};
            if (pm != k) z[pm] = {Ep.x + Op.y, Op.x - Ep.y};
        }
    }
    __syncthreads();
    stockham_fft<T, true>(z0, z1, where, M, L, tw, tw_n);     // same number of passes -> ends in z0
}

template <typename T>
__global__ void __launch_bounds__(SmoothCfg<T>::THREADS, 1)
spec_smooth_kernel(SmoothArgs a, int nB) {
    typedef typename Cx<T>::type T2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr bool kStockham = sizeof(T) == 4;                // f32: ping-pong Stockham FFT; f64: in-place
    T2 *z = reinterpret_cast<T2 *>(smem_raw);                 // n1/2 complex = n1 reals
    T *zr = reinterpret_cast<T *>(smem_raw);
    T2 *z1 = z + (a.n1 >> 1);                                 // second FFT buffer (f32 only)
    T *fr = zr + (kStockham ? 2 * a.n1 : a.n1);               // npix
    __shared__ double red[32];
    const T *a1f = static_cast<const T *>(a.a1_frac), *a2f = static_cast<const T *>(a.a2_frac);
    const T2 *tw = static_cast<const T2 *>(a.twiddle);
    const int npix = a.npix, n1 = a.n1;
    // all rows of the chunk, or (fallback launch behind the fast kernel) the listed ones
    const int nrows = a.rows ? *a.nrows : nB;

    for (int pi = blockIdx.x; pi < nrows; pi += gridDim.x) {
        const int p = a.rows ? a.rows[pi] : pi;
        const double *r = a.pars + (size_t)p * MS_NPARS;
        const int slot = a.star_slot ? a.star_slot[p] : a.slot0;
        const bool slot_ok = (unsigned)slot < (unsigned)a.nstars;
        const StarView sv = a.stars[slot_ok ? slot : 0];
        const int nobs = slot_ok ? sv.n_obs : 0;
        const bool ok = r[PV_TEFF] != ms_inf();
        if (!ok || nobs <= 0) {                               // uniform across the CTA
            if (a.flux_out && slot_ok)
                for (int j = threadIdx.x; j < nobs; j += blockDim.x)
                    a.flux_out[(size_t)p * nobs + j] = ms_nan();
            if (a.chi2 && ok && !slot_ok && threadIdx.x == 0) a.chi2[p] = ms_nan();   // slot outside the table
            continue;
        }
        const double vrad = r[PV_VRAD], vrot = r[PV_VROT];
        const double inst = 2.355 * r[PV_INSTR];              // genmod.py:74-75
        const T *frow = static_cast<const T *>(a.flux_native) + (size_t)p * npix;
        for (int i0 = threadIdx.x; i0 < npix; i0 += 4 * blockDim.x) {          // 4 loads in flight per thread
            T v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int i = i0 + u * blockDim.x; v[u] = i < npix ? ld_stream(frow + i) : (T)0; }
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int i = i0 + u * blockDim.x; if (i < npix) fr[i] = v[u]; }
        }
        __syncthreads();

        // ---- rotational broadening on the full native grid (outwave = None) ----
        if (vrot != 0.0) {
            const double sigma = sqrt(vrot * vrot);
            for (int k0 = threadIdx.x; k0 < n1; k0 += 4 * blockDim.x) {        // n1 is a power of two >= 4 * 512
                int ii[4]; T ff[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int k = k0 + u * blockDim.x;
                    ii[u] = k < n1 ? ld_stream(a.a1_idx + k) : 0;
                    ff[u] = k < n1 ? ld_stream(a1f + k) : (T)0;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int k = k0 + u * blockDim.x;
                    if (k < n1) zr[k] = fr[ii[u]] + (fr[ii[u] + 1] - fr[ii[u]]) * ff[u];
                }
            }
            __syncthreads();
            const double step1 = (double)(npix - 1) * a.dlnw / (double)(n1 - 1);
            Transfer<T> tf;
            tf.kind = 1; tf.tab = a.rot_tab;
            tf.coef = 2.0 * M_PI * sigma / ((double)n1 * MS_CKMS * step1);
            if (kStockham) fft_filter_stockham<T>(z, z1, n1, tf, static_cast<const T2 *>(a.twlev), n1); else fft_filter<T>(z, n1, tf, tw, n1);
            for (int i0 = threadIdx.x; i0 < npix; i0 += 4 * blockDim.x) {
                int kk[4]; T ff[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + u * blockDim.x;
                    kk[u] = i < npix ? ld_stream(a.a2_idx + i) : 0;
                    ff[u] = i < npix ? ld_stream(a2f + i) : (T)0;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + u * blockDim.x;
                    if (i < npix) fr[i] = zr[kk[u]] + (zr[kk[u] + 1] - zr[kk[u]]) * ff[u];
                }
            }
            __syncthreads();
        }

        if (a.native_out) {                                   // LSF mode: the instrumental stage runs in spec_lsf_kernel
            T *orow = static_cast<T *>(a.native_out) + (size_t)p * npix;
            for (int i = threadIdx.x; i < npix; i += blockDim.x) orow[i] = fr[i];
            __syncthreads();
            continue;
        }
        // ---- Doppler shift + instrumental broadening + resampling to the observed pixels ----
        const double scale = (vrad != 0.0) ? 1.0 + vrad / MS_C_KMS : 1.0;
        const double lnscale = log(scale);
        int mode;             // 0: FFT-smoothed grid in zr; 1: direct interp of fr; 2: all NaN
        int i_s = 0, i_e = npix - 1, n2 = 0;
        double d2 = 0.0;
        if (inst != 0.0) {
            const double lo = sv.wmin * (1.0 - 20.0 / inst), hi = sv.wmax * (1.0 + 20.0 / inst);
            int l = 0, h = npix;                              // first i with wave[i]*scale > lo
            while (l < h) { int m = (l + h) >> 1; if (a.wave[m] * scale > lo) h = m; else l = m + 1; }
            i_s = l;
            l = 0; h = npix;                                  // first i with wave[i]*scale >= hi
            while (l < h) { int m = (l + h) >> 1; if (a.wave[m] * scale < hi) l = m + 1; else h = m; }
            i_e = l - 1;
            const int nmask = i_e - i_s + 1;
            const double so = MS_CKMS / inst, si = MS_CKMS / a.resolution;
            const double sigma = sqrt(so * so - si * si);     // NaN when sharper than the net
            if (nmask < 4 || !(inst == inst)) mode = 2;      // degenerate window: all-NaN model
            else if (sigma <= 0.0) mode = 1;
            else {
                mode = 0;
                n2 = 1; while (n2 < nmask) n2 <<= 1;
                const double ratio = (double)(i_e - i_s) / (double)(n2 - 1);
                d2 = ratio * a.dlnw;
                const T half_d = (T)(0.5 * a.dlnw);
                for (int k = threadIdx.x; k < n2; k += blockDim.x) {
                    const double pos = (double)i_s + (double)k * ratio;
                    int i = (int)pos;
                    if (i > i_e - 1) i = i_e - 1;
                    const T t = (T)(pos - (double)i);
                    const T f = t * ((T)1 + (t - (T)1) * half_d);     // linear in lambda, not ln(lambda)
                    zr[k] = fr[i] + (fr[i + 1] - fr[i]) * f;
                }
                __syncthreads();
                Transfer<T> tf;
                tf.kind = 0; tf.tab = nullptr;
                const double ndv = (double)n2 * MS_CKMS * d2;
                tf.coef = 2.0 * M_PI * M_PI * sigma * sigma / (ndv * ndv);
                if (kStockham) fft_filter_stockham<T>(z, z1, n2, tf, static_cast<const T2 *>(a.twlev), n1); else fft_filter<T>(z, n2, tf, tw, n1);
            }
        } else {
            mode = 1;
        }

        double csum = 0.0;
        const double lnw_s = a.lnw0 + (double)i_s * a.dlnw;
        for (int j = threadIdx.x; j < nobs; j += blockDim.x) {
            const double lx = sv.obs_lnwave[j] - lnscale;
            double m;
            if (mode == 0) {
                const double q = (lx - lnw_s) / d2;
                if (q <= 0.0) m = (double)zr[0];
                else if (q >= (double)(n2 - 1)) m = (double)zr[n2 - 1];
                else {
                    const int i = (int)q;
                    const double t = q - (double)i;
                    const double f = t * (1.0 + (t - 1.0) * 0.5 * d2);
                    const double v0 = (double)zr[i], v1 = (double)zr[i + 1];
                    m = v0 + (v1 - v0) * f;
                }
            } else if (mode == 1) {
                const double q = (lx - a.lnw0) / a.dlnw;
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
            if (a.modpoly) m *= cheb_eval(sv.cheb_x[j], r + PV_PC0, a.n_pc);   // genmod.py:90-93
            if (a.flux_out) a.flux_out[(size_t)p * nobs + j] = m;
            if (a.chi2) {
                const double d = m - sv.obs_flux[j];
                const double t = (d * d) / sv.obs_err2[j];
                if (t == t) csum += t;                        // nansum (likelihood.py:193)
            }
        }
        if (a.chi2) {
            for (int o = 16; o > 0; o >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, o);
            if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = csum;
            __syncthreads();
            if (threadIdx.x == 0) {
                double tot = 0.0;
                for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
                a.chi2[p] += tot;
            }
        }
        __syncthreads();
    }
}

// ---- wavelength-dependent line-spread function (array Inst_R): smooth_lsf_fft, smoothing.py:412-514 --------------
// The reference warps the wavelength axis by the cumulative distribution of d(lambda) / sigma(lambda) so that the
// kernel has constant width, resamples the spectrum on nx = 2^k uniform points of the warped coordinate, smooths by
// FFT with a Gaussian of `x_per_sigma` and interpolates onto the observed wavelengths.  Everything up to the resampling
// map depends only on the net's grid and the star's sigma array (a Doppler factor s rescales the warped wavelengths
// lam -> s lam and the kernel width x_per_sigma -> x_per_sigma / s, nothing else), so the host prepares the map once
// (minesweeper_b200/lsf.py -> ms_set_lsf) and this kernel does, per point: gather-resample, Gaussian smoothing, search
// + interpolation at the observed pixels, blaze, chi-square.  nx = 2^ceil(log2(2 / x_per_sigma)) puts the kernel width
// at 2 .. 4 grid pixels, where the reference's circular FFT convolution equals the direct sum with Gaussian taps
// (Poisson summation, error exp(-pi^2 sigma^2 / 2) times the spectrum's Nyquist content); taps are cut at 5.5 sigma in
// f32 and 8.5 sigma in f64.
template <typename T>
__global__ void __launch_bounds__(256)
spec_lsf_kernel(SmoothArgs a, int nB) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sx = reinterpret_cast<T *>(smem_raw);                  // spectrum on the warped grid
    __shared__ double gt[64];
    __shared__ double red[8];
    for (int p = blockIdx.x; p < nB; p += gridDim.x) {
        const double *r = a.pars + (size_t)p * MS_NPARS;
        const int slot = a.star_slot ? a.star_slot[p] : a.slot0;
        const bool slot_ok = (unsigned)slot < (unsigned)a.nstars;
        const StarView sv = a.stars[slot_ok ? slot : 0];
        const int nobs = slot_ok ? sv.n_obs : 0, nx = sv.lsf_nx;
        if (r[PV_TEFF] == ms_inf() || nobs <= 0 || nx <= 0) {
            if (a.flux_out && slot_ok)
                for (int j = threadIdx.x; j < nobs; j += blockDim.x) a.flux_out[(size_t)p * nobs + j] = ms_nan();
            if (a.chi2 && r[PV_TEFF] != ms_inf() && threadIdx.x == 0 && (!slot_ok || nx <= 0)) a.chi2[p] = ms_nan();
            continue;
        }
        const double vrad = r[PV_VRAD];
        const double scale = (vrad != 0.0) ? 1.0 + vrad / MS_C_KMS : 1.0;
        const T *frow = static_cast<const T *>(a.flux_native) + (size_t)p * a.npix;
        for (int k = threadIdx.x; k < nx; k += blockDim.x) {
            const int i = sv.lsf_idx[k];
            const T f = (T)sv.lsf_frac[k];
            const T v0 = frow[i], v1 = frow[i + 1];
            sx[k] = v0 + (v1 - v0) * f;
        }
        const double spx = sv.lsf_xps / scale * (double)nx;  // kernel width in grid pixels
        const int ntap = min(63, (int)ceil((sizeof(T) == 4 ? 5.5 : 8.5) * spx));
        if (threadIdx.x <= ntap) {
            const double t = (double)threadIdx.x / spx;
            gt[threadIdx.x] = exp(-0.5 * t * t) / (spx * 2.5066282746310002);
        }
        __syncthreads();
        double csum = 0.0;
        const int mask = nx - 1;
        const double *lam = sv.lsf_lam;
        for (int j = threadIdx.x; j < nobs; j += blockDim.x) {
            const double x = sv.obs_wave[j] / scale;          // position on the unshifted warped wavelengths
            int lo = 0, hi = nx - 1;                          // largest i with lam[i] <= x (np.interp semantics)
            double m;
            auto conv = [&](int i) {
                double acc = 0.0;
                for (int t = -ntap; t <= ntap; ++t) acc += gt[t < 0 ? -t : t] * (double)sx[(i - t) & mask];
                return acc;
            };
            if (!(x == x)) m = ms_nan();
            else if (x <= lam[0]) m = conv(0);
            else if (x >= lam[nx - 1]) m = conv(nx - 1);
            else {
                while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (lam[mid] <= x) lo = mid; else hi = mid; }
                const double l0 = lam[lo], l1 = lam[lo + 1];
                const double c0 = conv(lo), c1 = conv(lo + 1);
                m = c0 + (c1 - c0) * ((x - l0) / (l1 - l0));
            }
            if (a.modpoly) m *= cheb_eval(sv.cheb_x[j], r + PV_PC0, a.n_pc);
            if (a.flux_out) a.flux_out[(size_t)p * nobs + j] = m;
            if (a.chi2) {
                const double d = m - sv.obs_flux[j];
                const double t = (d * d) / sv.obs_err2[j];
                if (t == t) csum += t;
            }
        }
        if (a.chi2) {
            for (int o = 16; o > 0; o >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, o);
            if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = csum;
            __syncthreads();
            if (threadIdx.x == 0) {
                double tot = 0.0;
                for (int w = 0; w < 8; ++w) tot += red[w];
                a.chi2[p] += tot;
            }
        }
        __syncthreads();
    }
}

}  // namespace

template <typename T>
int spec_mlp_chunk_t(ms_ctx *ctx, const double *pars, int64_t B, T *hid, T *flux,
                     const double *d_xmin_scale, cudaStream_t st);

// Chunk of points whose native flux sits between the MLP and the smoother.  Sized so that the chunk's flux
// (chunk x npix x 4 B) stays resident in the 126 MB L2: the GEMM's stores and the smoother's loads then never
// reach HBM (11 row tiles of 128 points x 80 pixel tiles = 880 tiles = 5.95 waves of 148 CTAs at npix = 10240).
// f64 (validation mode): small chunks, CUDA-core GEMM.
constexpr int64_t SPEC_SUPER = 64;      // tc modes: chunks whose hidden layers are computed in one go (operand images: 320 KB per 128 rows)

static int64_t spec_chunk(const ms_ctx *ctx) {
    if (ctx->precision == MS_PREC_F64) return 1024;
    const int64_t by_l2 = (int64_t)(60u << 20) / ((int64_t)ctx->spec.npix * 4);      // <= 60 MB of flux in flight
    int64_t c = (by_l2 / 128) * 128;
    // the fast broadening kernel runs two CTAs per SM, one spectrum per CTA at a time: a chunk that is a whole number
    // of waves has no partly filled last wave (1536 spectra on 296 CTAs would run 6 rounds for 5.2 rounds of work)
    const int64_t wave = (spec_fast_two_per_sm(ctx) ? 2 : 1) * (int64_t)ctx->sm_count;
    if (ctx->spec_fast && spec_fast_usable(ctx) && wave > 0 && by_l2 >= wave) c = (by_l2 / wave) * wave;
    if (c > 8192) c = 8192;
    if (c < 128) c = 128;
    return c;
}

int spec_reserve(ms_ctx *ctx, int64_t max_batch) {
    const SpecNetDev &n = ctx->spec;
    const size_t esz = ctx->precision == MS_PREC_F64 ? 8 : 4;
    int64_t chunk = spec_chunk(ctx);
    if (max_batch < chunk) chunk = max_batch;
    const size_t flux_bytes = (size_t)chunk * n.npix * esz;
    const size_t hid_bytes = (size_t)chunk * (n.d_in + 2 * n.h) * esz;
    // fallback counter (16 B), PointSetup records and fallback rows of a full chunk
    const size_t fb_bytes = 16 + (size_t)spec_chunk(ctx) * (sizeof(PointSetup) + sizeof(int));
    auto grow = [&](void **p, size_t *have, size_t need) -> int {
        if (*have >= need) return MS_OK;
        if (ctx->scratch_locked) { ms_set_error("spectrum scratch reserved for a smaller batch (ms_reserve)"); return MS_ESTATE; }
        if (*p) cudaFree(*p);
        *p = nullptr; *have = 0;
        MS_CUDA(cudaMalloc(p, need));
        *have = need;
        return MS_OK;
    };
    int rc;
    if ((rc = grow(&ctx->scr_flux, &ctx->scr_flux_bytes, flux_bytes))) return rc;
    if ((rc = grow(&ctx->scr_hid, &ctx->scr_hid_bytes, hid_bytes))) return rc;
    if ((rc = grow((void **)&ctx->scr_fb, &ctx->scr_fb_bytes, fb_bytes))) return rc;
    if (ctx->precision >= MS_PREC_TC) {            // hidden layers of up to SPEC_SUPER chunks at a time
        int64_t nch = (max_batch + chunk - 1) / chunk;
        if (nch > SPEC_SUPER) nch = SPEC_SUPER;
        if ((rc = spec_tc_reserve(ctx, chunk, (int)nch))) return rc;
    }
    return MS_OK;
}

template <typename T>
static int run_spec(ms_ctx *ctx, const SpecArgs &a, int64_t B, cudaStream_t st) {
    const SpecNetDev &n = ctx->spec;
    const int64_t chunk = spec_chunk(ctx);
    {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        const bool capturing = cudaStreamIsCapturing(st, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone;
        const int64_t need = B < chunk ? B : chunk;
        const size_t esz = sizeof(T);
        int64_t nch = (B + chunk - 1) / chunk;
        if (nch > SPEC_SUPER) nch = SPEC_SUPER;
        const bool tc_grow = ctx->precision >= MS_PREC_TC && spec_tc_would_grow(ctx, chunk, (int)nch);
        if (capturing && (ctx->scr_flux_bytes < (size_t)need * n.npix * esz || !ctx->scr_fb || tc_grow)) {
            ms_set_error("spectrum scratch would be allocated during stream capture: call ms_reserve first");
            return MS_ESTATE;
        }
        int rc = spec_reserve(ctx, B);
        if (rc) return rc;
    }
    const size_t smem = (size_t)((sizeof(T) == 4 ? 2 : 1) * n.n1 + n.npix) * sizeof(T);
    const bool generic_fits = smem <= 226 * 1024;
    const bool fast = sizeof(T) == 4 && ctx->spec_fast && spec_fast_usable(ctx);
    if (!generic_fits && !fast) {
        ms_set_error("spectral net with %d pixels does not fit the shared-memory smoother", n.npix);
        return MS_EINVAL;
    }
    auto kern = spec_smooth_kernel<T>;
    if (generic_fits) MS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // LSF mode (array Inst_R registered with ms_set_lsf for the star): all-or-nothing per batch
    bool lsf = false;
    int lsf_nx_max = 0;
    if (!a.star_slots) lsf = ctx->stars[a.star_slot].lsf_nx > 0, lsf_nx_max = ctx->stars[a.star_slot].lsf_nx;
    else {
        int n_lsf = 0, n_spec = 0;
        for (const StarDev &sd : ctx->stars) {
            if (sd.n_obs > 0) { ++n_spec; if (sd.lsf_nx > 0) { ++n_lsf; if (sd.lsf_nx > lsf_nx_max) lsf_nx_max = sd.lsf_nx; } }
        }
        if (n_lsf && n_lsf != n_spec) { ms_set_error("per-row star slots: either every star has an LSF array or none"); return MS_EINVAL; }
        lsf = n_lsf > 0;
    }
    if (lsf && (size_t)lsf_nx_max * sizeof(T) > 200 * 1024) { ms_set_error("LSF resampling length %d too large", lsf_nx_max); return MS_EINVAL; }
    int *fb_count = ctx->scr_fb;
    PointSetup *fb_setup = reinterpret_cast<PointSetup *>(ctx->scr_fb + 4);
    int *fb_rows = reinterpret_cast<int *>(fb_setup + chunk);
    const bool tc = sizeof(T) == 4 && ctx->precision >= MS_PREC_TC;
    bool hoisted = false;                          // hidden layers of the current super-batch are in place
    int64_t super0 = 0;
    for (int64_t b0 = 0; b0 < B; b0 += chunk) {
        const int64_t nb = (B - b0 < chunk) ? B - b0 : chunk;
        int rc;
        if (tc && (b0 - super0) % (SPEC_SUPER * chunk) == 0) {     // a new super-batch starts here
            super0 = b0;
            const int64_t nsup = B - b0 < SPEC_SUPER * chunk ? B - b0 : SPEC_SUPER * chunk;
            rc = spec_tc_hidden(ctx, a.pars + b0 * MS_NPARS, nsup, chunk, st);
            if (rc < 0) return rc;
            hoisted = rc == 0;
        }
        if (tc && hoisted) rc = spec_tc_out(ctx, (b0 - super0) / chunk, chunk, nb, (float *)ctx->scr_flux, st);
        else rc = spec_mlp_chunk_t<T>(ctx, a.pars + b0 * MS_NPARS, nb, (T *)ctx->scr_hid, (T *)ctx->scr_flux, n.xms, st);
        if (rc) return rc;
        SmoothArgs sa;
        sa.pars = a.pars + b0 * MS_NPARS;
        sa.flux_native = ctx->scr_flux;
        sa.npix = n.npix; sa.n1 = n.n1; sa.lnw0 = n.lnw0; sa.dlnw = n.dlnw; sa.resolution = n.resolution;
        sa.wave = n.wave; sa.a1_idx = n.a1_idx; sa.a2_idx = n.a2_idx;
        if (sizeof(T) == 4) { sa.a1_frac = n.a1_frac; sa.a2_frac = n.a2_frac; sa.twiddle = n.twiddle32; }
        else { sa.a1_frac = n.a1_fracd; sa.a2_frac = n.a2_fracd; sa.twiddle = n.twiddle64; }
        sa.rot_tab = n.rot_tab; sa.twlev = n.twlev32;
        sa.stars = ctx->stars_dev; sa.nstars = ctx->stars_dev_n;
        sa.star_slot = a.star_slots ? a.star_slots + b0 : nullptr; sa.slot0 = a.star_slot;
        sa.modpoly = a.modpoly; sa.n_pc = a.n_pc;
        sa.flux_out = a.flux ? a.flux + b0 * ctx->stars[a.star_slot].n_obs : nullptr;
        sa.chi2 = a.chi2 ? a.chi2 + b0 : nullptr;
        sa.rows = nullptr; sa.nrows = nullptr; sa.native_out = nullptr;
        KernelTimer timer_(ctx, MS_K_SPEC_SMOOTH, st);
        if (lsf) {
            // array Inst_R: rotational stage by the generic kernel, written back onto the native grid in place,
            // then the LSF kernel (warped-grid resampling, Gaussian smoothing, observed pixels, chi-square)
            if (!generic_fits) { ms_set_error("LSF mode needs the generic broadening kernel, which does not fit"); return MS_EINVAL; }
            int64_t grid = (int64_t)ctx->sm_count;
            if (grid > nb) grid = nb;
            SmoothArgs rot = sa;
            rot.native_out = ctx->scr_flux; rot.flux_out = nullptr; rot.chi2 = nullptr;
            kern<<<(unsigned)grid, SmoothCfg<T>::THREADS, smem, st>>>(rot, (int)nb);
            ctx->launches++;
            MS_CUDA(cudaGetLastError());
            const size_t lsmem = (size_t)lsf_nx_max * sizeof(T);
            auto lk = spec_lsf_kernel<T>;
            MS_CUDA(cudaFuncSetAttribute(lk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lsmem));
            int64_t lgrid = (int64_t)ctx->sm_count * (lsmem > 100 * 1024 ? 1 : 2);
            if (lgrid > nb) lgrid = nb;
            lk<<<(unsigned)lgrid, 256, lsmem, st>>>(sa, (int)nb);
        } else if (fast) {
            // fast kernel; rows it declines (unusual broadening widths) go through the generic FFT kernel
            MS_CUDA(cudaMemsetAsync(fb_count, 0, sizeof(int), st));
            if ((rc = launch_spec_fast(ctx, sa, (int)nb, fb_setup, fb_rows, fb_count, st))) return rc;
            if (!generic_fits) continue;
            sa.rows = fb_rows; sa.nrows = fb_count;
            const int64_t grid = nb < ctx->sm_count ? nb : ctx->sm_count;
            kern<<<(unsigned)grid, SmoothCfg<T>::THREADS, smem, st>>>(sa, (int)nb);
        } else {
            int64_t grid = (int64_t)ctx->sm_count;
            if (grid > nb) grid = nb;
            kern<<<(unsigned)grid, SmoothCfg<T>::THREADS, smem, st>>>(sa, (int)nb);
        }
        ctx->launches++;
        MS_CUDA(cudaGetLastError());
    }
    return MS_OK;
}

int launch_spec(ms_ctx *ctx, const SpecArgs &a, int64_t B, cudaStream_t st) {
    if (!ctx->spec.w0) { ms_set_error("spectral net not configured"); return MS_ESTATE; }
    if (!a.star_slots || a.flux) {
        if (a.star_slot < 0 || a.star_slot >= (int)ctx->stars.size() || !ctx->stars[a.star_slot].set ||
            ctx->stars[a.star_slot].n_obs <= 0) {
            ms_set_error("star slot %d has no observed spectrum", a.star_slot);
            return MS_ESTATE;
        }
    }
    if (a.star_slots && a.flux) { ms_set_error("model spectra with per-row star slots are ragged: not supported"); return MS_EINVAL; }
    if (!ctx->stars_dev) { ms_set_error("no star observations set"); return MS_ESTATE; }
    if (a.modpoly && (a.n_pc < 1 || a.n_pc > MS_MAX_PC)) { ms_set_error("n_pc out of range"); return MS_EINVAL; }
    if (B <= 0) return MS_OK;
    if (ctx->precision == MS_PREC_F64) return run_spec<double>(ctx, a, B, st);
    return run_spec<float>(ctx, a, B, st);
}
