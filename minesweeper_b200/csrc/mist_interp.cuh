// Device-side pieces of kernel (1), shared by the standalone interpolation kernel (mist_interp.cu)
// and by the prologue warpgroup of the tensor-core photometry kernel (phot_mlp_tc.cu), so both
// produce the same brackets and the same interpolated values bit for bit.
// Batched form of GenMIST.getMIST (reference minesweeper/fastMISTmod.py:100-108) with the semantics of
// params_to_grid (:127-150), linear_weights (:194-213) and weights (:152-170).
#pragma once
#include "common.cuh"

namespace mist {


// number of axis values <= x, minus one  ==  np.digitize([x], g, right=False) - 1.
// NaN compares false everywhere and lands at n - 1, like NumPy's NaN-last ordering.
__device__ __forceinline__ int bracket_of(const double *g, int n, double x) {
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (x < g[mid]) hi = mid; else lo = mid + 1;
    }
    return lo - 1;
}

template <typename T> struct Row;
template <> struct Row<float> {
    static constexpr int CP = MS_CP32;
    struct Raw { float4 a, b, c; };
    __device__ static Raw fetch(const float *p) {
        const float4 *q = reinterpret_cast<const float4 *>(p);
        return {__ldg(q), __ldg(q + 1), __ldg(q + 2)};
    }
    __device__ static void unpack(const Raw &r, double *v) {
        v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w;
        v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w; v[8] = r.c.x;
    }
};
template <> struct Row<double> {
    static constexpr int CP = MS_CP64;
    struct Raw { double2 a, b, c, d, e; };
    __device__ static Raw fetch(const double *p) {
        const double2 *q = reinterpret_cast<const double2 *>(p);
        return {__ldg(q), __ldg(q + 1), __ldg(q + 2), __ldg(q + 3), __ldg(q + 4)};
    }
    __device__ static void unpack(const Raw &r, double *v) {
        v[0] = r.a.x; v[1] = r.a.y; v[2] = r.b.x; v[3] = r.b.y; v[4] = r.c.x;
        v[5] = r.c.y; v[6] = r.d.x; v[7] = r.d.y; v[8] = r.e.x;
    }
};

// Requests the 16 corner rows of the cell that contains x into L2 (no registers held, nothing returned):
// a caller that processes points one after the other issues this for the next points first, so the
// dependent gathers below find their rows on chip.
template <typename T>
__device__ __forceinline__ void prefetch_point(const double *const g[4], const int n[4], const T *__restrict__ table,
                                               const double x[4]) {
    constexpr int CP = Row<T>::CP;
    int idx[4];
    bool inside = true;
#pragma unroll
    for (int d = 0; d < 4; ++d) {
        idx[d] = bracket_of(g[d], n[d], x[d]);
        inside = inside && (idx[d] >= 0) && (idx[d] <= n[d] - 2);
    }
    if (!inside) return;
    const int64_t sM = (int64_t)n[0] * CP, sF = (int64_t)n[1] * sM, sA = (int64_t)n[2] * sF;
    const T *base = table + idx[3] * sA + idx[2] * sF + idx[1] * sM + (int64_t)idx[0] * CP;
#pragma unroll
    for (int k = 0; k < 8; ++k) {                      // (afe, feh, mass) corner; the two EEP rows are adjacent
        const T *r = base + (k >> 2) * sA + ((k >> 1) & 1) * sF + (k & 1) * sM;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(r));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(r + 2 * CP - 1));
    }
}

// Brackets, weights and the renormalised 16-corner sum of one point.  g[d] / n[d]: axis values and
// lengths in the order EEP, mass, [Fe/H], [a/Fe]; table[na][nf][nm][ne][CP].  Returns true when the point
// lies inside the grid; acc / wsum are the un-normalised sums over the present vertices (wsum = 0 when
// none is present).  idx[d] is np.digitize(x_d, g_d, right=False) - 1 in every case.
// ROWS corner rows are requested per memory round trip (4, 8 or 16: a register / latency trade-off of the
// caller); the sums always run over the rows in ascending reference order, so the result does not depend
// on ROWS.
template <typename T, int ROWS = 4>
__device__ __forceinline__ bool interp_point(const double *const g[4], const int n[4], const T *__restrict__ table,
                                             const double x[4], int idx[4], double acc[MS_NCOL], double &wsum) {
    constexpr int CP = Row<T>::CP;
    const int ne = n[0], nm = n[1], nf = n[2];
    double wl[4], wu[4];
    bool inside = true;
#pragma unroll
    for (int d = 0; d < 4; ++d) {
        int i = bracket_of(g[d], n[d], x[d]);
        idx[d] = i;
        inside = inside && (i >= 0) && (i <= n[d] - 2);
        int ic = inside ? i : 0;
        // same operation order as the reference: frac, i + frac, then dx per corner
        double frac = (x[d] - g[d][ic]) / (g[d][ic + (n[d] > 1 ? 1 : 0)] - g[d][ic]);
        double xt = (double)ic + frac;
        double dl = xt - (double)ic;           // >= 0
        double du = xt - (double)(ic + 1);     // in [-1, 0)
        wl[d] = 1.0 - dl;
        wu[d] = (du > -1.0) ? 1.0 + du : 0.0;
    }
#pragma unroll
    for (int c = 0; c < MS_NCOL; ++c) acc[c] = 0.0;
    wsum = 0.0;
    if (inside) {
        const int64_t sE = CP, sM = (int64_t)ne * CP, sF = (int64_t)nm * sM, sA = (int64_t)nf * sF;
        const T *base = table + idx[3] * sA + idx[2] * sF + idx[1] * sM + idx[0] * sE;
        // ascending reference row id = (feh, afe, mass, eep) nesting: corner k has bits (cf, ca, cm, ce).
        // Every address is valid inside the grid, also for a zero-weight upper neighbour.
#pragma unroll
        for (int k0 = 0; k0 < 16; k0 += ROWS) {
            typename Row<T>::Raw raw[ROWS];
#pragma unroll
            for (int j = 0; j < ROWS; ++j) {
                const int k = k0 + j;
                raw[j] = Row<T>::fetch(base + ((k >> 2) & 1) * sA + (k >> 3) * sF + ((k >> 1) & 1) * sM + (k & 1) * sE);
            }
#pragma unroll
            for (int j = 0; j < ROWS; ++j) {
                const int k = k0 + j;
                const int cf = k >> 3, ca = (k >> 2) & 1, cm = (k >> 1) & 1, ce = k & 1;
                double w = (ce ? wu[0] : wl[0]);
                w = w * (cm ? wu[1] : wl[1]);
                w = w * (cf ? wu[2] : wl[2]);
                w = w * (ca ? wu[3] : wl[3]);
                if (w > 0.0) {
                    double v[MS_NCOL];
                    Row<T>::unpack(raw[j], v);
                    if (v[0] == v[0]) {      // vertex present
                        wsum += w;
#pragma unroll
                        for (int c = 0; c < MS_NCOL; ++c) acc[c] = fma(w, v[c], acc[c]);
                    }
                }
            }
        }
    }
    return inside;
}

// fp32-class form for the prologue of the tensor-core photometry kernel (f32 table): the brackets are found exactly as
// above (f64 comparisons: bit-exact cells), but the weights and the 16-corner sums are f32 -- the sums feed f32 network
// inputs and a distance modulus with an fp32-class tolerance, and the kernel is bound by instruction issue, where every
// f64 instruction costs two slots.  Same corner order, same renormalisation; results agree with interp_point to ~2e-7.
// The two columns that enter the distance modulus with a factor of 5 and 10 (log R, log Teff) are ALSO summed in f64
// with the same f32 weights (accd[0], accd[1], wsumd): a rounded weight only moves the ratio by its error times the
// spread of the corner values, so these two come out at ~1e-9 instead of one f32 ulp of 3.7 (2.4e-6 mag after x 10).
template <int ROWS = 2>
__device__ __forceinline__ bool interp_point_f32(const double *const g[4], const int n[4], const float *__restrict__ table,
                                                 const double x[4], int idx[4], float acc[MS_NCOL], float &wsum,
                                                 double accd[2], double &wsumd) {
    constexpr int CP = MS_CP32;
    const int ne = n[0], nm = n[1], nf = n[2];
    float wl[4], wu[4];
    bool inside = true;
#pragma unroll
    for (int d = 0; d < 4; ++d) {
        int i = bracket_of(g[d], n[d], x[d]);
        idx[d] = i;
        inside = inside && (i >= 0) && (i <= n[d] - 2);
        int ic = inside ? i : 0;
        const double g0 = g[d][ic];
        const float frac = __fdividef((float)(x[d] - g0), (float)(g[d][ic + (n[d] > 1 ? 1 : 0)] - g0));
        wl[d] = 1.f - frac;
        wu[d] = frac > 0.f ? frac : 0.f;
    }
#pragma unroll
    for (int c = 0; c < MS_NCOL; ++c) acc[c] = 0.f;
    wsum = 0.f;
    accd[0] = accd[1] = 0.0; wsumd = 0.0;
    if (inside) {
        const int64_t sE = CP, sM = (int64_t)ne * CP, sF = (int64_t)nm * sM, sA = (int64_t)nf * sF;
        const float *base = table + idx[3] * sA + idx[2] * sF + idx[1] * sM + idx[0] * sE;
#pragma unroll
        for (int k0 = 0; k0 < 16; k0 += ROWS) {
            Row<float>::Raw raw[ROWS];
#pragma unroll
            for (int j = 0; j < ROWS; ++j) {
                const int k = k0 + j;
                raw[j] = Row<float>::fetch(base + ((k >> 2) & 1) * sA + (k >> 3) * sF + ((k >> 1) & 1) * sM + (k & 1) * sE);
            }
#pragma unroll
            for (int j = 0; j < ROWS; ++j) {
                const int k = k0 + j;
                const int cf = k >> 3, ca = (k >> 2) & 1, cm = (k >> 1) & 1, ce = k & 1;
                float w = (ce ? wu[0] : wl[0]) * (cm ? wu[1] : wl[1]);
                w = w * ((cf ? wu[2] : wl[2]) * (ca ? wu[3] : wl[3]));
                const Row<float>::Raw &r = raw[j];
                if (w > 0.f && r.a.x == r.a.x) {       // weight present, vertex present
                    wsum += w;
                    const double wd = (double)w;
                    wsumd += wd;
                    accd[0] = fma(wd, (double)r.a.z, accd[0]); accd[1] = fma(wd, (double)r.b.x, accd[1]);
                    acc[0] = fmaf(w, r.a.x, acc[0]); acc[1] = fmaf(w, r.a.y, acc[1]); acc[2] = fmaf(w, r.a.z, acc[2]);
                    acc[3] = fmaf(w, r.a.w, acc[3]); acc[4] = fmaf(w, r.b.x, acc[4]); acc[5] = fmaf(w, r.b.y, acc[5]);
                    acc[6] = fmaf(w, r.b.z, acc[6]); acc[7] = fmaf(w, r.b.w, acc[7]); acc[8] = fmaf(w, r.c.x, acc[8]);
                }
            }
        }
    }
    return inside;
}

}  // namespace mist
