// Argument block and small device helpers shared by the two broadening kernels:
// spec_smooth.cu (generic: f64, and the f32 fallback for unusual rows) and
// spec_fast.cu (f32 fast path).  See spec_smooth.cu for the reference lines.
#pragma once
#include "common.cuh"

#define MS_CKMS 2.998e5            // smoothing.py:14
#define MS_C_KMS 299792.458        // Payne's Doppler shift (scipy.constants.c / 1000)
#define MS_ROT_TAB_H (1.0 / 32.0)  // spacing of the rotational-transfer table
#define MS_ROT_TAB_N 8200

struct SmoothArgs {
    const double *pars;            // [B][MS_NPARS] (already offset to the chunk)
    const void *flux_native;       // [B][npix] T
    int npix, n1;
    double lnw0, dlnw, resolution;
    const double *wave;            // native wavelengths
    const int *a1_idx, *a2_idx;
    const void *a1_frac, *a2_frac; // T
    const void *twiddle;           // T2 [n1/2]: exp(-2 pi i k / n1)
    const void *twlev;             // float2 per-level compact tables (Stockham path)
    const float *rot_tab;          // S(u) on a uniform grid (float mode)
    // observations: per-row star slots (catalog fits with spectra) or one slot for every row
    const StarView *stars;         // [nstars] device table
    const int32_t *star_slot;      // [B] (offset to the chunk) or nullptr -> slot0
    int slot0, nstars;
    int modpoly, n_pc;
    double *flux_out;              // [B][n_obs of slot0] or nullptr
    double *chi2;                  // [B] or nullptr
    // row indirection (fallback launches of the generic kernel): rows[0 .. *nrows) or nullptr = all rows
    const int *rows, *nrows;
    // LSF mode: the generic kernel stops after the rotational stage and writes the native-grid spectrum here ([B][npix] T)
    void *native_out;
};

// Transfer function of the rotational profile, smoothing.py:541-548 (f32: 4-point Lagrange on a table).
__device__ __forceinline__ float rot_transfer_f(float u, const float *__restrict__ tab) {
    const float x = u * (float)(1.0 / MS_ROT_TAB_H);
    int i = (int)x;
    if (i >= 1 && i < MS_ROT_TAB_N - 3) {
        const float t = x - (float)i;
        const float p0 = __ldg(tab + i - 1), p1 = __ldg(tab + i), p2 = __ldg(tab + i + 1), p3 = __ldg(tab + i + 2);
        const float c0 = -t * (t - 1.f) * (t - 2.f) * (1.f / 6.f);
        const float c1 = (t + 1.f) * (t - 1.f) * (t - 2.f) * 0.5f;
        const float c2 = -(t + 1.f) * t * (t - 2.f) * 0.5f;
        const float c3 = (t + 1.f) * t * (t - 1.f) * (1.f / 6.f);
        return c0 * p0 + c1 * p1 + c2 * p2 + c3 * p3;
    }
    if (i < 1) {                                    // series about 0: 1 - u^2/10 + u^4/280
        const float u2 = u * u;
        return 1.f - u2 * 0.1f + u2 * u2 * (1.f / 280.f);
    }
    // beyond the table (u >= 256, i.e. vsini above ~50 km/s at the Nyquist frequency): Hankel asymptotics of J1,
    // J1(u) ~ sqrt(2 / (pi u)) [P cos(chi) - Q sin(chi)], chi = u - 3 pi / 4, P = 1 + 15 / (128 u^2),
    // Q = 3 / (8 u) - 105 / (1024 u^3)  (truncation error ~ u^-4 < 3e-10), with the argument reduced in f64 so
    // that the fast sine / cosine see |x| < 2 pi (no slow-path calls, hence no register spills around them)
    const double ud = (double)u;
    const double red = ud - 6.283185307179586 * floor(ud * 0.15915494309189535);
    const float s = __sinf((float)red), c = __cosf((float)red);
    const float iu = 1.f / u, iu2 = iu * iu;
    const float r = 0.70710678118654752440f;
    const float cchi = r * (s - c), schi = -r * (s + c);            // cos / sin of (u - 3 pi / 4)
    const float P = 1.f + 0.1171875f * iu2, Q = iu * (0.375f - 0.1025390625f * iu2);
    const float j1v = sqrtf(0.63661977236758134f * iu) * (P * cchi - Q * schi);
    return j1v * iu - 1.5f * c * iu2 + 1.5f * s * iu2 * iu;
}

// numpy.polynomial.chebyshev.chebval recurrence (Clenshaw), n coefficients
__device__ __forceinline__ double cheb_eval(double x, const double *c, int n) {
    if (n == 1) return c[0];
    if (n == 2) return c[0] + c[1] * x;
    const double x2 = 2.0 * x;
    double c0 = c[n - 2], c1 = c[n - 1];
    for (int i = 3; i <= n; ++i) {
        const double tmp = c0;
        c0 = c[n - i] - c1;
        c1 = tmp + c1 * x2;
    }
    return c0 + c1 * x;
}

// First native pixel inside the padded window and the last one (mask_wave, smoothing.py:560-576, on the
// Doppler-shifted grid wave * scale): i_s = first i with wave[i] * scale > lo, i_e = last i with
// wave[i] * scale < hi.  The grid is log-uniform, so a candidate comes from one logarithm and is then fixed up
// with the reference's own comparisons (exact at the boundaries).
__device__ __forceinline__ void mask_window(const double *__restrict__ wave, int npix, double lnw0, double dlnw,
                                            double scale, double lo, double hi, int &i_s, int &i_e) {
    int cs = lo > 0.0 ? (int)ceil((log(lo / scale) - lnw0) / dlnw) : 0;
    cs = cs < 0 ? 0 : (cs > npix ? npix : cs);
    int ce = (int)floor((log(hi / scale) - lnw0) / dlnw) + 1;        // candidate for "first i with wave >= hi"
    ce = ce < 0 ? 0 : (ce > npix ? npix : ce);
    // the four neighbours that decide whether the candidates are right, fetched together (one memory round trip)
    const double ws0 = wave[cs > 0 ? cs - 1 : 0], ws1 = wave[cs < npix ? cs : npix - 1];
    const double we0 = wave[ce > 0 ? ce - 1 : 0], we1 = wave[ce < npix ? ce : npix - 1];
    const bool s_ok = (cs == 0 || !(ws0 * scale > lo)) && (cs == npix || ws1 * scale > lo);
    const bool e_ok = (ce == 0 || we0 * scale < hi) && (ce == npix || !(we1 * scale < hi));
    if (!s_ok) {
        while (cs > 0 && wave[cs - 1] * scale > lo) --cs;
        while (cs < npix && !(wave[cs] * scale > lo)) ++cs;
    }
    if (!e_ok) {
        while (ce > 0 && !(wave[ce - 1] * scale < hi)) --ce;
        while (ce < npix && wave[ce] * scale < hi) ++ce;
    }
    i_s = cs;
    i_e = ce - 1;
}

// per-point decisions of the fast broadening kernel (spec_fast.cu), one record per point of a chunk
struct __align__(16) PointSetup {
    int mode;             // 0: Gaussian FIR on the n2 grid; 1: plain interpolation; 2: all NaN; 3: fallback; 4: skip row
    int i_s, i_e, n2, ntap, nobs, rot, slot_ok, teff_inf, pad_[3];
    double d2, inv_d2, ratio2, spx, lnscale, lnw_s, inst, vrot;
};
static_assert(sizeof(PointSetup) == 112, "PointSetup is copied in 16-byte pieces");

// kernel launchers
int launch_spec_fast(ms_ctx *ctx, const SmoothArgs &sa, int nB, void *setup, int *fallback_rows, int *fallback_count,
                     cudaStream_t st);
bool spec_fast_usable(const ms_ctx *ctx);
bool spec_fast_two_per_sm(const ms_ctx *ctx);
