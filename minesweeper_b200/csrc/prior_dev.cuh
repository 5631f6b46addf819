// Device-side arithmetic of the batched prior, shared by the element-wise kernels of
// prior.cu and by the fused sampler kernels of sampler.cu (propose + prior transform, ln prior + accept).
// Everything here has internal linkage (the library is built without relocatable device code).
#pragma once
#include "common.cuh"
#include <cstring>
#include <cmath>

namespace {

struct TransformArgs { int ndim; ms_prior_col col[MS_NPARAM]; const double *tab_x, *tab_cdf; int tab_n; };

__device__ __forceinline__ double norm_cdf(double x) { return 0.5 * erfc(-x * 0.70710678118654752440); }

// Regularised incomplete beta function I_x(a, b) by the continued fraction (modified Lentz), with the usual
// symmetry switch so that the fraction converges quickly; and its inverse (scipy.stats.beta.ppf) by bisection
// polished with Newton steps.
__device__ double beta_cf(double a, double b, double x) {
    const double tiny = 1e-300;
    const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0, d = 1.0 - qab * x / qap;
    if (fabs(d) < tiny) d = tiny;
    d = 1.0 / d;
    double h = d;
    for (int m = 1; m <= 300; ++m) {
        const int m2 = 2 * m;
        double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
        d = 1.0 + aa * d; if (fabs(d) < tiny) d = tiny;
        c = 1.0 + aa / c; if (fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        h *= d * c;
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
        d = 1.0 + aa * d; if (fabs(d) < tiny) d = tiny;
        c = 1.0 + aa / c; if (fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        const double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < 1e-16) break;
    }
    return h;
}
__device__ double beta_inc(double a, double b, double x) {
    if (x <= 0.0) return 0.0;
    if (x >= 1.0) return 1.0;
    const double lbt = lgamma(a + b) - lgamma(a) - lgamma(b) + a * log(x) + b * log1p(-x);
    if (x < (a + 1.0) / (a + b + 2.0)) return exp(lbt) * beta_cf(a, b, x) / a;
    return 1.0 - exp(lbt) * beta_cf(b, a, 1.0 - x) / b;
}
__device__ double beta_ppf(double u, double a, double b) {
    if (!(u > 0.0)) return u == 0.0 ? 0.0 : ms_nan();
    if (!(u < 1.0)) return u == 1.0 ? 1.0 : ms_nan();
    double lo = 0.0, hi = 1.0, x = 0.5;
    const double lnorm = lgamma(a + b) - lgamma(a) - lgamma(b);
    for (int it = 0; it < 200; ++it) {
        const double f = beta_inc(a, b, x) - u;
        if (f > 0.0) hi = x; else lo = x;
        // Newton step from the density; fall back to bisection when it leaves the bracket
        const double pdf = exp(lnorm + (a - 1.0) * log(x) + (b - 1.0) * log1p(-x));
        double xn = (pdf > 0.0 && isfinite(pdf)) ? x - f / pdf : 0.5 * (lo + hi);
        if (!(xn > lo && xn < hi)) xn = 0.5 * (lo + hi);
        if (fabs(xn - x) <= 1e-16 * fmax(x, 1e-300) || hi - lo <= 1e-17) { x = xn; break; }
        x = xn;
    }
    return x;
}

// prior transform of one unit-cube coordinate x of column c (prior.py:150-385)
__device__ __forceinline__ double prior_transform_value(const TransformArgs &a, int c, double x) {
    const ms_prior_col &pc = a.col[c];
    double v;
    switch (pc.kind) {
        case MS_PRIOR_UNIFORM:                                   // (hi - lo) u + lo
            v = (pc.p[1] - pc.p[0]) * x + pc.p[0];
            break;
        case MS_PRIOR_GAUSSIAN:                                  // norm.ppf(u, loc, scale)
            v = pc.p[0] + pc.p[1] * normcdfinv(x);
            break;
        case MS_PRIOR_TGAUSSIAN: {                               // truncnorm.ppf(u, a, b, loc, scale); p = lo, hi, loc, scale
            const double al = (pc.p[0] - pc.p[2]) / pc.p[3], be = (pc.p[1] - pc.p[2]) / pc.p[3];
            double z;
            if (al > 0.0) {                                      // right tail: work with survival functions
                const double sa = norm_cdf(-al), sb = norm_cdf(-be);
                z = -normcdfinv(sa - x * (sa - sb));
            } else {
                const double ca = norm_cdf(al), cb = norm_cdf(be);
                z = normcdfinv(ca + x * (cb - ca));
            }
            v = pc.p[2] + pc.p[3] * z;
            if (v == ms_inf()) v = pc.p[1];                      // prior.py:193-194
            break;
        }
        case MS_PRIOR_EXP:                                       // expon.ppf(u, loc, scale)
            v = pc.p[0] - pc.p[1] * log1p(-x);
            break;
        case MS_PRIOR_LOGUNIFORM:                                // reciprocal.ppf(u, a, b)
            v = pc.p[0] * exp(x * log(pc.p[1] / pc.p[0]));
            break;
        case MS_PRIOR_BETA:                                      // beta.ppf(u, a, b, loc, scale)
            v = pc.p[2] + pc.p[3] * beta_ppf(x, pc.p[0], pc.p[1]);
            break;
        case MS_PRIOR_TABLE: {                                   // scale * np.interp(u, cdf, x)
            const int n = a.tab_n;
            if (!(x == x) || n < 2) { v = ms_nan(); break; }
            if (x <= a.tab_cdf[0]) { v = pc.p[0] * a.tab_x[0]; break; }
            if (x >= a.tab_cdf[n - 1]) { v = pc.p[0] * a.tab_x[n - 1]; break; }
            int lo = 0, hi = n - 1;                              // largest i with cdf[i] <= u
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (a.tab_cdf[mid] <= x) lo = mid; else hi = mid; }
            const double c0 = a.tab_cdf[lo], c1 = a.tab_cdf[lo + 1];
            const double slope = (a.tab_x[lo + 1] - a.tab_x[lo]) / (c1 - c0);
            v = pc.p[0] * (slope * (x - c0) + a.tab_x[lo]);      // numpy's interp arithmetic
            break;
        }
        default:
            v = ms_nan();
    }
    return v;
}

struct LnPriorArgs {
    ThetaView tv;
    const double *derived;          // [B][13] or nullptr
    const double *lnL;              // [B] or nullptr (then out = lnprior)
    double *out;
    int mistdif, nterms, imf;
    double imf_hi_off, imf_lognorm; // constants of the Kroupa-like IMF, computed once on the host (fill_lnprior_args)
    const int32_t *star_slot; const double *star_plx;   // per-star Gaussian parallax prior or nullptr
    ms_prior_term term[MS_MAX_PRIOR_TERMS];
    int has_adv; ms_adv_prior adv;
};

__device__ __forceinline__ double lse3(double a, double b, double c) {
    const double m = fmax(a, fmax(b, c));
    if (m == -ms_inf()) return m;
    return m + log(exp(a - m) + exp(b - m) + exp(c - m));
}
// ln number density of the disk and halo components (advancedpriors.py:241-327), defaults of gal_lnprior
__device__ __forceinline__ double logn_disk(double R, double Z, double R_scale, double Z_scale) {
    return -((R - 8.2) / R_scale + (fabs(Z) - fabs(0.025)) / Z_scale);
}
__device__ __forceinline__ double logn_halo(double R, double Z) {
    const double R_solar = 8.2, Z_solar = 0.025, R_smooth = 0.5, eta = 4.2, q_ctr = 0.2, q_inf = 0.8, r_q = 6.0;
    const double r = sqrt(R * R + Z * Z);
    const double rp = sqrt(r * r + r_q * r_q);
    const double q = q_inf - (q_inf - q_ctr) * exp(1.0 - rp / r_q);
    const double Reff = sqrt(R * R + (Z / q) * (Z / q) + R_smooth * R_smooth);
    const double rp_solar = sqrt(R_solar * R_solar + Z_solar * Z_solar + r_q * r_q);
    const double q_solar = q_inf - (q_inf - q_ctr) * exp(1.0 - rp_solar / r_q);
    const double Reff_solar = sqrt(R_solar * R_solar + (Z_solar / q_solar) + R_smooth * R_smooth);   // as the reference (:322)
    return -eta * log(Reff / Reff_solar);
}
// age prior of one Galactic component (logp_age_normal / logp_age_unif, advancedpriors.py:836-905)
__device__ __forceinline__ double logp_age(double age, const double *p) {
    if (!(age >= p[0] && age <= p[1])) return -ms_inf();
    if (p[3] > 0.0) {
        const double chi2 = (p[2] - age) * (p[2] - age) / (p[3] * p[3]);
        return -0.5 * (chi2 + log(2.0 * M_PI * p[3] * p[3]));
    }
    return 0.0;
}

__device__ __forceinline__ double prior_value(const LnPriorArgs &a, int64_t p, int id) {
    if (id < MS_NPARAM) {
        // sampled / fixed parameter, except the stellar labels an isochrone fit predicts
        if (a.derived && (id == MS_P_TEFF || id == MS_P_LOGG || id == MS_P_FEH || id == MS_P_AFE || id == MS_P_LOGR)) {
            const double *d = a.derived + p * MS_NDERIVED;
            switch (id) {
                case MS_P_TEFF: return pow(10.0, d[8]);
                case MS_P_LOGG: return d[11];
                case MS_P_FEH: return a.mistdif ? d[9] : d[2];
                case MS_P_AFE: return a.mistdif ? d[10] : d[3];
                default: return d[6];
            }
        }
        return theta_get(a.tv, p, id);
    }
    switch (id) {
        case MS_V_LOGAGE: return a.derived ? a.derived[p * MS_NDERIVED + 4] : ms_nan();
        case MS_V_AGE: return a.derived ? pow(10.0, a.derived[p * MS_NDERIVED + 4] - 9.0) : ms_nan();   // prior.py:526
        case MS_V_PARALLAX: return 1000.0 / theta_get(a.tv, p, MS_P_DIST);                           // prior.py:612
        default: return ms_nan();
    }
}

// lnL + ln prior of point p with lnprobfn's two -inf short circuits (fitstar.py:715-727, prior.py:387-639)
__device__ __forceinline__ double lnprob_value(const LnPriorArgs &a, int64_t p) {
    const double neg_inf = -ms_inf();
    double l = a.lnL ? a.lnL[p] : 0.0;
    if (a.lnL && l == neg_inf) return neg_inf;         // fitstar.py:719-720
    double lp = 0.0;
    if (a.imf) {                                                          // advancedpriors.py:93-137
        const double m = theta_get(a.tv, p, MS_P_INIT_MASS);
        const double al = 1.3, ah = 2.3, mb = 0.5;
        double v = neg_inf;
        if (m <= mb && m > 0.08) v = -al * log(m);
        else if (m > mb) v = -ah * log(m) + a.imf_hi_off;                 // + (ah - al) log(mb)
        lp += v - a.imf_lognorm;                                         // - log(norm_low + norm_high)
    }
    if (a.star_plx) {
        const double *pe = a.star_plx + 2 * (int64_t)(a.star_slot ? a.star_slot[p] : 0);
        const double d = 1000.0 / theta_get(a.tv, p, MS_P_DIST) - pe[0];
        lp += -0.5 * ((d * d) / (pe[1] * pe[1]));
    }
    if (a.has_adv) {
        const ms_adv_prior &A = a.adv;
        if (A.gal) {                                                      // prior.py:413-446
            const double logage = prior_value(a, p, MS_V_LOGAGE);
            const double dist = theta_get(a.tv, p, MS_P_DIST);
            if (isfinite(logage) && isfinite(dist)) {
                const double d = dist / 1000.0;
                if (d <= 0.0) lp = neg_inf;                               // advancedpriors.py:513-518
                else {
                    const double vol = 2.0 * log(d + 1e-300);
                    const double X = d * A.Xp - A.sol_X, Y = d * A.Yp, Z = d * A.Zp - A.sol_Z;
                    const double R = hypot(X, Y);
                    const double thin = logn_disk(R, Z, 2.6, 0.3) + vol;
                    const double thick = logn_disk(R, Z, 2.0, 0.9) + vol + log(0.04);
                    const double halo = logn_halo(R, Z) + vol + log(0.005);
                    const double lnp_dist = lse3(thin, thick, halo);
                    double lnp_age = 0.0;
                    if (A.galage) {
                        const double age = pow(10.0, logage - 9.0);
                        lnp_age = lse3(logp_age(age, A.age_thin) + (thin - lnp_dist),
                                       logp_age(age, A.age_thick) + (thick - lnp_dist),
                                       logp_age(age, A.age_halo) + (halo - lnp_dist));
                    }
                    lp += lnp_dist + lnp_age;
                }
            } else lp = ms_nan();        // the reference raises NameError here; it never gets there (lnL = -inf first)
        }
        if (A.vrot) {                                                     // prior.py:453-466, advancedpriors.py:691-733
            const double vrot = theta_get(a.tv, p, MS_P_VROT);
            const double logg = prior_value(a, p, MS_P_LOGG);
            double mass = theta_get(a.tv, p, MS_P_INIT_MASS);
            if (!(mass == mass)) mass = pow(10.0, logg + 2.0 * prior_value(a, p, MS_P_LOGR));
            const double eep = theta_get(a.tv, p, MS_P_EEP);
            double pa, pc_, pn;
            if (mass > 1.25) { pa = -1.0; pc_ = 100.0; pn = 1.0; }        // Kraft break
            else if ((logg < 3.5) || (eep > 450.0)) { pa = A.vrot_giant[0]; pc_ = A.vrot_giant[1]; pn = A.vrot_giant[2]; }
            else { pa = A.vrot_dwarf[0]; pc_ = A.vrot_dwarf[1]; pn = A.vrot_dwarf[2]; }
            lp += pa / (1.0 + pn * exp(-(vrot - pc_)));
        }
        if (A.alpha) {                                                    // prior.py:494-501, advancedpriors.py:672-688
            double afe = theta_get(a.tv, p, MS_P_INIT_AFE);
            if (!(afe == afe)) afe = prior_value(a, p, MS_P_AFE);
            const double logg = prior_value(a, p, MS_P_LOGG), eep = theta_get(a.tv, p, MS_P_EEP);
            if (((logg < 3.5) || (eep > 450.0)) && afe < A.minalpha) lp += -0.5 * (afe / 0.05) * (afe / 0.05);
        }
    }
    bool dead = false;
    for (int i = 0; i < a.nterms && !dead; ++i) {
        const ms_prior_term &t = a.term[i];
        const double v = prior_value(a, p, t.value_id);
        switch (t.kind) {
            case MS_TERM_UNIFORM:                                         // bounds only
                if (v < t.p[0] || v > t.p[1]) dead = true;
                break;
            case MS_TERM_GAUSSIAN:
                lp += -0.5 * (((v - t.p[0]) * (v - t.p[0])) / (t.p[1] * t.p[1]));
                break;
            case MS_TERM_TGAUSSIAN:                                       // bounds + gaussian (prior.py:537-542)
                if (v < t.p[0] || v > t.p[1]) dead = true;
                else lp += -0.5 * (((v - t.p[2]) * (v - t.p[2])) / (t.p[3] * t.p[3]));
                break;
            case MS_TERM_TGAUSSIAN_BOUNDS:                                // spec / phot families: bounds only (:575-578, 629-633)
                if (v < t.p[0] || v > t.p[1]) dead = true;
                break;
        }
    }
    if (dead || lp == neg_inf) return neg_inf;          // fitstar.py:724-725
    return lp + l;
}

// ---- host side: argument blocks from the C-ABI tables (validation included) ------------------------------------------
inline int fill_transform_args(ms_ctx *ctx, int ndim, const ms_prior_col *cols, TransformArgs &a) {
    if (!cols || ndim < 1 || ndim > MS_NPARAM) { ms_set_error("bad arguments"); return MS_EINVAL; }
    a.ndim = ndim;
    a.tab_x = ctx->ptab_x; a.tab_cdf = ctx->ptab_cdf; a.tab_n = ctx->ptab_n;
    for (int i = 0; i < ndim; ++i) {
        if (cols[i].kind == MS_PRIOR_TABLE && ctx->ptab_n < 2) { ms_set_error("column %d uses the quantile table but none is set", i); return MS_ESTATE; }
        if (cols[i].kind < 0 || cols[i].kind > MS_PRIOR_TABLE) { ms_set_error("column %d: unsupported prior kind %d", i, cols[i].kind); return MS_EINVAL; }
        a.col[i] = cols[i];
    }
    return MS_OK;
}

inline int fill_lnprior_args(const ms_flags *flags, const double *theta, int ndim, const int32_t *colmap,
                             const double *fixed, const double *derived, const ms_prior_term *terms, int nterms, int imf,
                             const double *lnL, double *out, const int32_t *star_slot, const double *star_plx,
                             const ms_adv_prior *adv, LnPriorArgs &a) {
    if (!flags || !theta || !colmap || !out || ndim < 1 || ndim > MS_NPARAM || nterms < 0 ||
        nterms > MS_MAX_PRIOR_TERMS || (nterms && !terms)) { ms_set_error("bad arguments"); return MS_EINVAL; }
    a.tv.theta = theta; a.tv.ld = ndim;
    for (int i = 0; i < MS_NPARAM; ++i) { a.tv.col[i] = -1; a.tv.fixed[i] = fixed ? fixed[i] : nan(""); }
    for (int c = 0; c < ndim; ++c) {
        if (colmap[c] < 0 || colmap[c] >= MS_NPARAM) { ms_set_error("colmap[%d] out of range", c); return MS_EINVAL; }
        a.tv.col[colmap[c]] = c;
    }
    a.derived = derived; a.lnL = lnL; a.out = out; a.mistdif = flags->mistdif; a.nterms = nterms; a.imf = imf;
    {   // normalisation of the broken power law (advancedpriors.py:93-137): three pow() and a log() per point otherwise
        const double al = 1.3, ah = 2.3, mb = 0.5;
        const double nlo = std::pow(mb, 1.0 - al) / (ah - 1.0);
        const double nhi = std::pow(0.08, 1.0 - al) / (al - 1.0) - std::pow(mb, 1.0 - al) / (al - 1.0);
        a.imf_hi_off = (ah - al) * std::log(mb);
        a.imf_lognorm = std::log(nlo + nhi);
    }
    a.star_slot = star_slot; a.star_plx = star_plx;
    a.has_adv = adv != nullptr;
    if (adv) a.adv = *adv; else memset(&a.adv, 0, sizeof(a.adv));
    if (adv && (adv->gal || adv->galage) && !derived) { ms_set_error("the Galactic priors need the derived block (log(Age))"); return MS_EINVAL; }
    for (int i = 0; i < nterms; ++i) {
        if (terms[i].value_id < 0 || terms[i].value_id >= MS_V_END || terms[i].kind < 0 || terms[i].kind > MS_TERM_TGAUSSIAN_BOUNDS) {
            ms_set_error("prior term %d is malformed", i);
            return MS_EINVAL;
        }
        a.term[i] = terms[i];
    }
    return MS_OK;
}

}  // namespace
