// extern "C" entry points of libmsb200.so (see include/minesweeper_b200.h for the
// reference interface each one replaces).  Host-side work here is limited to
// one-time model packing, argument checks and launch sequencing; all per-point
// arithmetic runs in the kernels.
#include "common.cuh"

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <limits>
#include <new>

static thread_local char g_err[512] = "";

void ms_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *ms_last_error(void) { return g_err; }
extern "C" const char *ms_version(void) { return "minesweeper_b200 0.1 (sm_100a)"; }

namespace {

template <typename T> int upload(T **dst, const T *src, size_t n) {
    MS_CUDA(cudaMalloc((void **)dst, n * sizeof(T)));
    MS_CUDA(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
    return MS_OK;
}

int setup_grid(ms_ctx *ctx, const ms_grid_desc *g) {
    if (g->ne < 2 || g->nm < 2 || g->nf < 2 || g->na < 2 || !g->eep || !g->mass || !g->feh ||
        !g->afe || !g->values || !g->valid) {
        ms_set_error("grid descriptor incomplete (every axis needs >= 2 values)");
        return MS_EINVAL;
    }
    const int n[4] = {g->ne, g->nm, g->nf, g->na};
    const double *ax[4] = {g->eep, g->mass, g->feh, g->afe};
    std::vector<double> axes;
    for (int d = 0; d < 4; ++d)
        for (int i = 0; i < n[d]; ++i) {
            if (i && !(ax[d][i] > ax[d][i - 1])) {
                ms_set_error("grid axis %d is not strictly increasing at %d", d, i);
                return MS_EINVAL;
            }
            axes.push_back(ax[d][i]);
        }
    GridDev &G = ctx->grid;
    G.ne = g->ne; G.nm = g->nm; G.nf = g->nf; G.na = g->na;
    int rc = upload(&G.axes, axes.data(), axes.size());
    if (rc) return rc;
    const size_t nv = (size_t)g->ne * g->nm * g->nf * g->na;
    const double qnan = std::numeric_limits<double>::quiet_NaN();
    // pack in slabs so the host staging buffer stays small
    const size_t slab = 1u << 20;
    if (ctx->precision == MS_PREC_F64) {
        MS_CUDA(cudaMalloc((void **)&G.tab64, nv * MS_CP64 * sizeof(double)));
        std::vector<double> buf(slab * MS_CP64);
        for (size_t v0 = 0; v0 < nv; v0 += slab) {
            const size_t m = (nv - v0 < slab) ? nv - v0 : slab;
            for (size_t v = 0; v < m; ++v) {
                const double *src = g->values + (v0 + v) * MS_NCOL;
                double *dst = buf.data() + v * MS_CP64;
                const bool ok = g->valid[v0 + v] != 0;
                if (ok && !std::isfinite(src[0])) {
                    ms_set_error("grid vertex %zu is marked valid but has a non-finite value", v0 + v);
                    return MS_EINVAL;
                }
                for (int c = 0; c < MS_NCOL; ++c) dst[c] = ok ? src[c] : qnan;
                for (int c = MS_NCOL; c < MS_CP64; ++c) dst[c] = 0.0;
            }
            MS_CUDA(cudaMemcpy(G.tab64 + v0 * MS_CP64, buf.data(), m * MS_CP64 * sizeof(double),
                               cudaMemcpyHostToDevice));
        }
    } else {
        MS_CUDA(cudaMalloc((void **)&G.tab32, nv * MS_CP32 * sizeof(float)));
        std::vector<float> buf(slab * MS_CP32);
        for (size_t v0 = 0; v0 < nv; v0 += slab) {
            const size_t m = (nv - v0 < slab) ? nv - v0 : slab;
            for (size_t v = 0; v < m; ++v) {
                const double *src = g->values + (v0 + v) * MS_NCOL;
                float *dst = buf.data() + v * MS_CP32;
                const bool ok = g->valid[v0 + v] != 0;
                if (ok && !std::isfinite(src[0])) {
                    ms_set_error("grid vertex %zu is marked valid but has a non-finite value", v0 + v);
                    return MS_EINVAL;
                }
                for (int c = 0; c < MS_NCOL; ++c) dst[c] = ok ? (float)src[c] : (float)qnan;
                for (int c = MS_NCOL; c < MS_CP32; ++c) dst[c] = 0.f;
            }
            MS_CUDA(cudaMemcpy(G.tab32 + v0 * MS_CP32, buf.data(), m * MS_CP32 * sizeof(float),
                               cudaMemcpyHostToDevice));
        }
    }
    return MS_OK;
}

int setup_phot(ms_ctx *ctx, const ms_phot_net_desc *d) {
    if (d->nb < 1 || d->d_in < 1 || d->d_in > 6 || d->h < 8 || !d->w1 || !d->b1 || !d->w2 ||
        !d->b2 || !d->w3 || !d->b3 || !d->xmin || !d->xmax) {
        ms_set_error("phot net descriptor incomplete");
        return MS_EINVAL;
    }
    PhotNetDev &P = ctx->phot;
    P.nb = d->nb; P.d_in = d->d_in; P.h = d->h;
    const size_t nb = d->nb, h = d->h, di = d->d_in;
    int rc;
    if ((rc = upload(&P.w1, d->w1, nb * h * di))) return rc;
    if ((rc = upload(&P.b1, d->b1, nb * h))) return rc;
    if ((rc = upload(&P.w2, d->w2, nb * h * h))) return rc;
    if ((rc = upload(&P.b2, d->b2, nb * h))) return rc;
    if ((rc = upload(&P.w3, d->w3, nb * h))) return rc;
    if ((rc = upload(&P.b3, d->b3, nb))) return rc;
    for (int i = 0; i < 8; ++i) { P.xmin[i] = 0; P.xscale[i] = 0; }
    for (int i = 0; i < d->d_in; ++i) {
        P.xmin[i] = d->xmin[i];
        P.xscale[i] = 1.0 / (d->xmax[i] - d->xmin[i]);
    }
    if (ctx->precision >= MS_PREC_TC) return phot_tc_prepare(ctx);
    return MS_OK;
}

}  // namespace

// Spectral net upload + the theta-independent pieces of the broadening chain:
// the two resampling maps of the vsini stage (native -> power-of-two log grid ->
// native; smoothing.py:587-597 and :262), FFT twiddles, rotational-transfer table.
int spec_prepare(ms_ctx *ctx, const ms_spec_net_desc *d) {
    if (d->d_in < 1 || d->d_in > 5 || d->h < 1 || d->npix < 8 || !d->w0 || !d->b0 || !d->w1 ||
        !d->b1 || !d->w2 || !d->b2 || !d->xmin || !d->xmax || !d->wavelength) {
        ms_set_error("spec net descriptor incomplete");
        return MS_EINVAL;
    }
    SpecNetDev &S = ctx->spec;
    S.d_in = d->d_in; S.h = d->h; S.npix = d->npix;
    S.resolution = d->resolution; S.leak = d->leak; S.logt_input = d->logt_input;
    const size_t h = d->h, di = d->d_in, P = d->npix;
    int rc;
    if ((rc = upload(&S.w0, d->w0, h * di))) return rc;
    if ((rc = upload(&S.b0, d->b0, h))) return rc;
    if ((rc = upload(&S.w1, d->w1, h * h))) return rc;
    if ((rc = upload(&S.b1, d->b1, h))) return rc;
    if ((rc = upload(&S.w2, d->w2, P * h))) return rc;
    if ((rc = upload(&S.b2, d->b2, P))) return rc;
    if ((rc = upload(&S.wave, d->wavelength, P))) return rc;
    double xms[16] = {0};
    for (int i = 0; i < d->d_in; ++i) {
        S.xmin[i] = xms[i] = d->xmin[i];
        S.xscale[i] = xms[8 + i] = 1.0 / (d->xmax[i] - d->xmin[i]);
    }
    if ((rc = upload(&S.xms, xms, 16))) return rc;

    // log-uniformity of the native grid (required: DESIGN.md "spectral resampling")
    const double *w = d->wavelength;
    S.lnw0 = std::log(w[0]);
    S.dlnw = (std::log(w[P - 1]) - S.lnw0) / (double)(P - 1);
    for (size_t i = 0; i < P; ++i) {
        const double dev = std::fabs(std::log(w[i]) - (S.lnw0 + (double)i * S.dlnw)) / S.dlnw;
        if (!(dev < 1e-6)) {
            ms_set_error("spectral wavelength grid is not log-uniform at pixel %zu (%.3g px)", i, dev);
            return MS_EINVAL;
        }
    }
    int n1 = 1;
    while (n1 < d->npix) n1 <<= 1;
    S.n1 = n1;
    // map 1: native -> n1-point log grid between w[0] and w[P-1]
    std::vector<int> i1(n1), i2(P);
    std::vector<double> f1(n1), f2(P), g1(n1);
    const double l0 = std::log(w[0]), l1 = std::log(w[P - 1]);
    for (int k = 0; k < n1; ++k) {
        // np.linspace(l0, l1, n1): start + k*step, last sample exactly l1
        const double step = (l1 - l0) / (double)(n1 - 1);
        g1[k] = std::exp(k == n1 - 1 ? l1 : l0 + (double)k * step);
    }
    {
        size_t i = 0;
        for (int k = 0; k < n1; ++k) {
            const double x = g1[k];
            while (i + 2 < P && w[i + 1] <= x) ++i;
            double f = (x - w[i]) / (w[i + 1] - w[i]);
            if (x <= w[0]) f = 0.0;                              // np.interp clamps outside
            if (x >= w[P - 1]) { i = P - 2; f = 1.0; }
            i1[k] = (int)i; f1[k] = f;
        }
    }
    {
        int k = 0;
        for (size_t i = 0; i < P; ++i) {
            const double x = w[i];
            while (k + 2 < n1 && g1[k + 1] <= x) ++k;
            double f = (x - g1[k]) / (g1[k + 1] - g1[k]);
            if (x <= g1[0]) f = 0.0;
            if (x >= g1[n1 - 1]) { k = n1 - 2; f = 1.0; }
            i2[i] = k; f2[i] = f;
        }
    }
    std::vector<float> f1f(f1.begin(), f1.end()), f2f(f2.begin(), f2.end());
    if ((rc = upload(&S.a1_idx, i1.data(), i1.size()))) return rc;
    if ((rc = upload(&S.a2_idx, i2.data(), i2.size()))) return rc;
    if ((rc = upload(&S.a1_fracd, f1.data(), f1.size()))) return rc;
    if ((rc = upload(&S.a2_fracd, f2.data(), f2.size()))) return rc;
    if ((rc = upload(&S.a1_frac, f1f.data(), f1f.size()))) return rc;
    if ((rc = upload(&S.a2_frac, f2f.data(), f2f.size()))) return rc;
    // twiddles exp(-2 pi i k / n1), k < n1/2
    std::vector<double2> tw(n1 / 2);
    std::vector<float2> twf(n1 / 2);
    for (int k = 0; k < n1 / 2; ++k) {
        const double a = -2.0 * M_PI * (double)k / (double)n1;
        tw[k] = make_double2(std::cos(a), std::sin(a));
        twf[k] = make_float2((float)tw[k].x, (float)tw[k].y);
    }
    S.twiddle_n = n1;
    if ((rc = upload(&S.twiddle64, tw.data(), tw.size()))) return rc;
    if ((rc = upload(&S.twiddle32, twf.data(), twf.size()))) return rc;
    {   // compact per-level tables for the Stockham passes (unit-stride reads, small L1 footprint)
        int lmax = 0;
        while ((1 << lmax) < n1) ++lmax;
        std::vector<float2> lev((size_t)1 << lmax, make_float2(1.f, 0.f));
        for (int l = 1; l <= lmax; ++l)
            for (int k = 0; k < (1 << (l - 1)); ++k) {
                const double a = -2.0 * M_PI * (double)k / (double)(1 << l);
                lev[((size_t)1 << (l - 1)) + k] = make_float2((float)std::cos(a), (float)std::sin(a));
            }
        if ((rc = upload(&S.twlev32, lev.data(), lev.size()))) return rc;
    }
    // rotational transfer S(u) = J1(u)/u - 3 cos u / (2 u^2) + 3 sin u / (2 u^3) on u = i/32
    const int NT = 8200;
    std::vector<float> rt(NT);
    for (int i = 0; i < NT; ++i) {
        const double u = (double)i / 32.0;
        double s;
        if (u < 0.05) {
            const double u2 = u * u;
            s = 1.0 - u2 / 10.0 + u2 * u2 / 280.0 - u2 * u2 * u2 / 15120.0;
        } else {
            s = j1(u) / u - 3.0 * std::cos(u) / (2.0 * u * u) + 3.0 * std::sin(u) / (2.0 * u * u * u);
        }
        rt[i] = (float)s;
    }
    if ((rc = upload(&S.rot_tab, rt.data(), rt.size()))) return rc;
    if (ctx->precision >= MS_PREC_TC) return spec_tc_prepare(ctx);
    return MS_OK;
}

extern "C" int ms_create(ms_ctx **out, int device, int precision, const ms_grid_desc *grid,
                         const ms_phot_net_desc *phot, const ms_spec_net_desc *spec) {
    if (!out) { ms_set_error("out is NULL"); return MS_EINVAL; }
    *out = nullptr;
    if (precision < MS_PREC_F64 || precision > MS_PREC_TC16) {
        ms_set_error("unknown precision %d", precision);
        return MS_EINVAL;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        ms_set_error("no CUDA device available: minesweeper_b200 has no CPU fallback");
        return MS_ENODEV;
    }
    if (device < 0 || device >= ndev) { ms_set_error("device %d out of range", device); return MS_ENODEV; }
    MS_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    MS_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        ms_set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                     prop.minor);
        return MS_ENODEV;
    }
    ms_ctx *ctx = new (std::nothrow) ms_ctx();
    if (!ctx) return MS_ENOMEM;
    ctx->device = device;
    ctx->precision = precision;
    ctx->sm_count = prop.multiProcessorCount;
    int rc = MS_OK;
    if (grid && (rc = setup_grid(ctx, grid))) { ms_destroy(ctx); return rc; }
    if (phot && (rc = setup_phot(ctx, phot))) { ms_destroy(ctx); return rc; }
    if (spec && (rc = spec_prepare(ctx, spec))) { ms_destroy(ctx); return rc; }
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking) != cudaSuccess) {
        ms_destroy(ctx);
        ms_set_error("stream creation failed");
        return MS_ECUDA;
    }
    *out = ctx;
    return MS_OK;
}

extern "C" void ms_destroy(ms_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    GridDev &G = ctx->grid;
    cudaFree(G.axes); cudaFree(G.tab32); cudaFree(G.tab64);
    PhotNetDev &P = ctx->phot;
    cudaFree(P.w1); cudaFree(P.b1); cudaFree(P.w2); cudaFree(P.b2); cudaFree(P.w3); cudaFree(P.b3);
    cudaFree(P.tc_pack);
    SpecNetDev &S = ctx->spec;
    cudaFree(S.w0); cudaFree(S.b0); cudaFree(S.w1); cudaFree(S.b1); cudaFree(S.w2); cudaFree(S.b2);
    cudaFree(S.wave); cudaFree(S.a1_idx); cudaFree(S.a2_idx); cudaFree(S.a1_frac); cudaFree(S.a2_frac);
    cudaFree(S.a1_fracd); cudaFree(S.a2_fracd); cudaFree(S.twiddle32); cudaFree(S.twiddle64);
    cudaFree(S.xms); cudaFree(S.rot_tab); cudaFree(S.twlev32); spec_tc_release(ctx);
    cudaFree(ctx->obs_mag); cudaFree(ctx->obs_err2); cudaFree(ctx->stars_dev); cudaFree(ctx->scr_fb);
    cudaFree(ctx->ptab_x); cudaFree(ctx->ptab_cdf);
    for (StarDev &s : ctx->stars) {
        cudaFree(s.obs_wave); cudaFree(s.obs_lnwave); cudaFree(s.obs_flux); cudaFree(s.obs_err2);
        cudaFree(s.cheb_x); cudaFree(s.lsf_lam); cudaFree(s.lsf_frac); cudaFree(s.lsf_idx);
    }
    cudaFree(ctx->scr_derived); cudaFree(ctx->scr_pars); cudaFree(ctx->scr_chi2); cudaFree(ctx->scr_chi2g);
    cudaFree(ctx->scr_flux); cudaFree(ctx->scr_hid); cudaFree(ctx->scr_theta); cudaFree(ctx->scr_out);
    for (auto &r : ctx->prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
    cudaFree(ctx->rows_ready_dev);
    if (ctx->rows_ready_host) cudaFreeHost(ctx->rows_ready_host);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    if (ctx->h2d_stream) cudaStreamDestroy(ctx->h2d_stream);
    if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
    delete ctx;
}

extern "C" int ms_precision(const ms_ctx *ctx) { return ctx ? ctx->precision : MS_EINVAL; }
extern "C" int64_t ms_launch_count(const ms_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int ms_set_spec_fast(ms_ctx *ctx, int enable) {
    if (!ctx) { ms_set_error("bad arguments"); return MS_EINVAL; }
    ctx->spec_fast = enable ? 1 : 0;
    return MS_OK;
}

extern "C" int ms_set_lnl_peers(ms_ctx *ctx, const uint64_t *peers, int n, int64_t offset) {
    if (!ctx || n < 0 || n > MS_MAX_PEERS || (n > 0 && !peers) || offset < 0) { ms_set_error("bad arguments"); return MS_EINVAL; }
    for (int i = 0; i < MS_MAX_PEERS; ++i) ctx->peers[i] = i < n ? reinterpret_cast<double *>(peers[i]) : nullptr;
    ctx->npeers = n; ctx->peer_off = offset;
    return MS_OK;
}

extern "C" int ms_set_fusion(ms_ctx *ctx, int enable) {
    if (!ctx) { ms_set_error("bad arguments"); return MS_EINVAL; }
    ctx->fuse_interp = enable ? 1 : 0;
    return MS_OK;
}

extern "C" int ms_profile_enable(ms_ctx *ctx, int on) {
    if (!ctx) return MS_EINVAL;
    for (auto &r : ctx->prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    ctx->prof.clear();
    ctx->prof_on = on != 0;
    return MS_OK;
}

extern "C" int ms_profile_read(ms_ctx *ctx, double *ms_by_kind, int64_t *launches_by_kind) {
    if (!ctx || !ms_by_kind || !launches_by_kind) return MS_EINVAL;
    for (int k = 0; k < MS_NKIND; ++k) { ms_by_kind[k] = 0.0; launches_by_kind[k] = 0; }
    for (auto &r : ctx->prof) {
        MS_CUDA(cudaEventSynchronize(r.e1));
        float ms = 0.f;
        MS_CUDA(cudaEventElapsedTime(&ms, r.e0, r.e1));
        ms_by_kind[r.kind] += ms;
        launches_by_kind[r.kind] += 1;
    }
    return MS_OK;
}

// device table of the per-slot spectrum views (what the broadening kernels index by star slot)
static int refresh_star_views(ms_ctx *ctx) {
    std::vector<StarView> v(ctx->stars.size());
    for (size_t i = 0; i < v.size(); ++i) {
        const StarDev &s = ctx->stars[i];
        v[i] = StarView{s.n_obs, 0, s.obs_wave, s.obs_lnwave, s.obs_flux, s.obs_err2, s.cheb_x, s.wmin, s.wmax,
                        s.lsf_nx, 0, s.lsf_lam, s.lsf_frac, s.lsf_idx, s.lsf_xps};
    }
    if ((int)v.size() > ctx->stars_dev_n) {
        cudaFree(ctx->stars_dev);
        ctx->stars_dev = nullptr; ctx->stars_dev_n = 0;
        int cap = 1;
        while (cap < (int)v.size()) cap *= 2;
        MS_CUDA(cudaMalloc((void **)&ctx->stars_dev, (size_t)cap * sizeof(StarView)));
        MS_CUDA(cudaMemset(ctx->stars_dev, 0, (size_t)cap * sizeof(StarView)));
        ctx->stars_dev_n = cap;
    }
    if (!v.empty())
        MS_CUDA(cudaMemcpy(ctx->stars_dev, v.data(), v.size() * sizeof(StarView), cudaMemcpyHostToDevice));
    return MS_OK;
}

extern "C" int ms_set_star(ms_ctx *ctx, int star_slot, const double *obs_mag, const double *obs_err,
                           int nb, const double *obs_wave, const double *obs_flux,
                           const double *obs_eflux, int n_obs) {
    if (!ctx || star_slot < 0 || star_slot > (1 << 24)) { ms_set_error("bad ctx / star slot"); return MS_EINVAL; }
    MS_CUDA(cudaSetDevice(ctx->device));
    if (nb > 0) {
        if (nb != ctx->phot.nb || !obs_mag || !obs_err) {
            ms_set_error("photometry has %d bands, the net has %d", nb, ctx->phot.nb);
            return MS_EINVAL;
        }
        if (star_slot >= ctx->nslots) {                      // grow the slot table, keep old rows
            int ns = ctx->nslots ? ctx->nslots : 1;
            while (ns <= star_slot) ns *= 2;
            double *m = nullptr, *e = nullptr;
            MS_CUDA(cudaMalloc((void **)&m, (size_t)ns * nb * sizeof(double)));
            if (cudaMalloc((void **)&e, (size_t)ns * nb * sizeof(double)) != cudaSuccess) {
                cudaGetLastError(); cudaFree(m);
                ms_set_error("cudaMalloc of the star-slot table failed");
                return MS_ENOMEM;
            }
            // unobserved everywhere by default: NaN magnitude drops the band
            MS_CUDA(cudaMemset(m, 0xff, (size_t)ns * nb * sizeof(double)));
            MS_CUDA(cudaMemset(e, 0xff, (size_t)ns * nb * sizeof(double)));
            if (ctx->nslots) {
                MS_CUDA(cudaMemcpy(m, ctx->obs_mag, (size_t)ctx->nslots * nb * sizeof(double), cudaMemcpyDeviceToDevice));
                MS_CUDA(cudaMemcpy(e, ctx->obs_err2, (size_t)ctx->nslots * nb * sizeof(double), cudaMemcpyDeviceToDevice));
            }
            cudaFree(ctx->obs_mag); cudaFree(ctx->obs_err2);
            ctx->obs_mag = m; ctx->obs_err2 = e; ctx->nslots = ns;
        }
        std::vector<double> e2(nb);
        for (int b = 0; b < nb; ++b) e2[b] = obs_err[b] * obs_err[b];
        MS_CUDA(cudaMemcpy(ctx->obs_mag + (size_t)star_slot * nb, obs_mag, nb * sizeof(double), cudaMemcpyHostToDevice));
        MS_CUDA(cudaMemcpy(ctx->obs_err2 + (size_t)star_slot * nb, e2.data(), nb * sizeof(double), cudaMemcpyHostToDevice));
    }
    if ((int)ctx->stars.size() <= star_slot) ctx->stars.resize(star_slot + 1);
    StarDev &s = ctx->stars[star_slot];
    cudaFree(s.obs_wave); cudaFree(s.obs_lnwave); cudaFree(s.obs_flux); cudaFree(s.obs_err2); cudaFree(s.cheb_x);
    cudaFree(s.lsf_lam); cudaFree(s.lsf_frac); cudaFree(s.lsf_idx);
    s = StarDev();
    s.set = true;
    if (n_obs > 0) {
        if (!obs_wave || !obs_flux || !obs_eflux) { ms_set_error("spectrum arrays missing"); return MS_EINVAL; }
        std::vector<double> lw(n_obs), e2(n_obs), cx(n_obs);
        double wmin = obs_wave[0], wmax = obs_wave[0];
        for (int j = 0; j < n_obs; ++j) {
            wmin = obs_wave[j] < wmin ? obs_wave[j] : wmin;
            wmax = obs_wave[j] > wmax ? obs_wave[j] : wmax;
        }
        for (int j = 0; j < n_obs; ++j) {
            lw[j] = std::log(obs_wave[j]);
            e2[j] = obs_eflux[j] * obs_eflux[j];
            const double x = obs_wave[j] - wmin;               // fitutils.py:13-14
            cx[j] = 2.0 * (x / (wmax - wmin)) - 1.0;
        }
        int rc;
        if ((rc = upload(&s.obs_wave, obs_wave, n_obs))) return rc;
        if ((rc = upload(&s.obs_lnwave, lw.data(), n_obs))) return rc;
        if ((rc = upload(&s.obs_flux, obs_flux, n_obs))) return rc;
        if ((rc = upload(&s.obs_err2, e2.data(), n_obs))) return rc;
        if ((rc = upload(&s.cheb_x, cx.data(), n_obs))) return rc;
        s.n_obs = n_obs; s.wmin = wmin; s.wmax = wmax;
    }
    return refresh_star_views(ctx);
}

// Wavelength-dependent LSF of a star slot (array Inst_R; smooth_lsf_fft, smoothing.py:412-514): the theta-independent
// warped grid prepared on the host (minesweeper_b200/lsf.py).  nx = 0 clears.
extern "C" int ms_set_lsf(ms_ctx *ctx, int star_slot, int nx, const double *lam, const int32_t *idx, const double *frac,
                          double x_per_sigma) {
    if (!ctx || star_slot < 0 || star_slot >= (int)ctx->stars.size() || !ctx->stars[star_slot].set) {
        ms_set_error("ms_set_lsf: star slot %d is not set (call ms_set_star first)", star_slot);
        return MS_ESTATE;
    }
    if (nx < 0 || (nx & (nx - 1)) != 0 || (nx > 0 && (!lam || !idx || !frac || !(x_per_sigma > 0.0)))) {
        ms_set_error("ms_set_lsf: nx must be a power of two with its three arrays");
        return MS_EINVAL;
    }
    MS_CUDA(cudaSetDevice(ctx->device));
    StarDev &s = ctx->stars[star_slot];
    cudaFree(s.lsf_lam); cudaFree(s.lsf_frac); cudaFree(s.lsf_idx);
    s.lsf_lam = s.lsf_frac = nullptr; s.lsf_idx = nullptr; s.lsf_nx = 0; s.lsf_xps = 0;
    if (nx > 0) {
        for (int k = 0; k < nx; ++k)
            if (idx[k] < 0 || idx[k] > ctx->spec.npix - 2) { ms_set_error("ms_set_lsf: idx[%d] out of range", k); return MS_EINVAL; }
        int rc;
        if ((rc = upload(&s.lsf_lam, lam, (size_t)nx))) return rc;
        if ((rc = upload(&s.lsf_frac, frac, (size_t)nx))) return rc;
        if ((rc = upload(&s.lsf_idx, (const int *)idx, (size_t)nx))) return rc;
        s.lsf_nx = nx; s.lsf_xps = x_per_sigma;
    }
    return refresh_star_views(ctx);
}

extern "C" int ms_set_stars_phot(ms_ctx *ctx, int first_slot, int nstars, const double *obs_mag,
                                 const double *obs_err, int nb) {
    if (!ctx || first_slot < 0 || nstars < 1 || !obs_mag || !obs_err || nb != ctx->phot.nb) {
        ms_set_error("bad arguments (nb must equal the net's band count)");
        return MS_EINVAL;
    }
    // grow the table through the single-star path, then overwrite the block in two copies
    int rc = ms_set_star(ctx, first_slot + nstars - 1, obs_mag + (size_t)(nstars - 1) * nb,
                         obs_err + (size_t)(nstars - 1) * nb, nb, nullptr, nullptr, nullptr, 0);
    if (rc) return rc;
    if ((int)ctx->stars.size() < first_slot + nstars) ctx->stars.resize(first_slot + nstars);
    for (int i = 0; i < nstars; ++i) ctx->stars[first_slot + i].set = true;
    std::vector<double> e2((size_t)nstars * nb);
    for (size_t i = 0; i < e2.size(); ++i) e2[i] = obs_err[i] * obs_err[i];
    MS_CUDA(cudaMemcpy(ctx->obs_mag + (size_t)first_slot * nb, obs_mag, e2.size() * sizeof(double), cudaMemcpyHostToDevice));
    MS_CUDA(cudaMemcpy(ctx->obs_err2 + (size_t)first_slot * nb, e2.data(), e2.size() * sizeof(double), cudaMemcpyHostToDevice));
    return refresh_star_views(ctx);
}

static void identity_view(ThetaView &tv, const double *theta, int ld) {
    tv.theta = theta; tv.ld = ld;
    for (int i = 0; i < MS_NPARAM; ++i) { tv.col[i] = -1; tv.fixed[i] = std::numeric_limits<double>::quiet_NaN(); }
}

extern "C" int ms_mist_interp(ms_ctx *ctx, const double *theta4, int64_t B, int32_t *brackets,
                              double *derived, uint8_t *ok, void *stream) {
    if (!ctx || (!theta4 && B > 0) || B < 0) { ms_set_error("bad arguments"); return MS_EINVAL; }
    MS_CUDA(cudaSetDevice(ctx->device));
    ThetaView tv;
    identity_view(tv, theta4, 4);
    tv.col[MS_P_EEP] = 0; tv.col[MS_P_INIT_MASS] = 1; tv.col[MS_P_INIT_FEH] = 2; tv.col[MS_P_INIT_AFE] = 3;
    return launch_mist_interp(ctx, tv, B, brackets, derived, ok, (cudaStream_t)stream);
}

static int run_phot(ms_ctx *ctx, const PhotArgs &a, int64_t B, cudaStream_t st) {
    if (ctx->precision >= MS_PREC_TC) return launch_phot_tc(ctx, a, B, st);
    return launch_phot(ctx, a, B, st);
}

extern "C" int ms_phot_sed(ms_ctx *ctx, const double *photpars, int64_t B, double *mags, void *stream) {
    if (ctx && B == 0) return MS_OK;
    if (!ctx || !photpars || !mags || B < 0) { ms_set_error("bad arguments"); return MS_EINVAL; }
    MS_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if ((rc = ensure_scratch(ctx, &ctx->scr_pars, &ctx->scr_pars_n, (size_t)B * MS_NPARS, st))) return rc;
    if ((rc = launch_repack_pars(ctx, photpars, 7, 0, B, ctx->scr_pars, st))) return rc;
    PhotArgs a{ctx->scr_pars, nullptr, mags, nullptr};
    return run_phot(ctx, a, B, st);
}

extern "C" int ms_spec_model(ms_ctx *ctx, int star_slot, const double *specpars, int n_pc, int64_t B,
                             double *flux, void *stream) {
    if (ctx && B == 0) return MS_OK;
    if (!ctx || !specpars || !flux || B < 0 || n_pc < 0 || n_pc > MS_MAX_PC) { ms_set_error("bad arguments"); return MS_EINVAL; }
    MS_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if ((rc = ensure_scratch(ctx, &ctx->scr_pars, &ctx->scr_pars_n, (size_t)B * MS_NPARS, st))) return rc;
    if ((rc = launch_repack_pars(ctx, specpars, 8 + n_pc, 1, B, ctx->scr_pars, st))) return rc;
    SpecArgs a{ctx->scr_pars, star_slot, n_pc > 0, n_pc, flux, nullptr};
    return launch_spec(ctx, a, B, st);
}

static int check_batch_args(const ms_ctx *ctx, const ms_flags *flags, int64_t B, int ndim, const int32_t *colmap);

extern "C" int ms_lnlike_batch(ms_ctx *ctx, const ms_flags *flags, const double *theta,
                               const int32_t *star_slot, int64_t B, int ndim, const int32_t *colmap,
                               const double *fixed, double *lnL, double *derived, int32_t *brackets,
                               void *stream) {
    int rc;
    if ((rc = check_batch_args(ctx, flags, B, ndim, colmap))) return rc;
    if (B == 0) return MS_OK;                              // empty batch: nothing to do, NULL buffers allowed
    if (!theta || !lnL) { ms_set_error("bad arguments"); return MS_EINVAL; }
    MS_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    ThetaView tv;
    identity_view(tv, theta, ndim);
    if (fixed) for (int i = 0; i < MS_NPARAM; ++i) tv.fixed[i] = fixed[i];
    for (int c = 0; c < ndim; ++c) tv.col[colmap[c]] = c;
    if (flags->phot && !star_slot &&
        (flags->slot >= ctx->nslots || flags->slot >= (int)ctx->stars.size() || !ctx->stars[flags->slot].set)) {
        ms_set_error("star slot %d has no photometry loaded (ms_set_star)", flags->slot);
        return MS_ESTATE;
    }
    const bool iso = tv.col[MS_P_EEP] >= 0 || tv.fixed[MS_P_EEP] == tv.fixed[MS_P_EEP];
    if (!iso && (derived || brackets)) {
        ms_set_error("derived / brackets requested but EEP is not a parameter");
        return MS_EINVAL;
    }
    // photometry-only isochrone batch on the tensor-core path: interpolation folded into the MLP kernel
    // (worth it once every persistent CTA sweeps several tile pairs, so that the interpolation of pair k + 1
    // hides behind the MLP of pair k; below that the two-kernel sequence is faster)
    if (iso && flags->phot && !flags->spec && !brackets && ctx->fuse_interp && ctx->precision >= MS_PREC_TC &&
        B >= (int64_t)ctx->sm_count * 256 * 4 && phot_tc_can_fuse(ctx)) {
        PhotArgs pa{nullptr, star_slot, nullptr, nullptr};
        pa.lnL = lnL; pa.ageweight = flags->ageweight; pa.slot0 = flags->slot;
        return launch_phot_tc_fused(ctx, tv, pa, derived, flags->mistdif, B, st);
    }
    if ((rc = ensure_scratch(ctx, &ctx->scr_pars, &ctx->scr_pars_n, (size_t)B * MS_NPARS, st))) return rc;
    // chi2 scratch is only needed when a spectrum term precedes the photometry
    double *chi2 = nullptr;
    if (flags->spec || !flags->phot) {
        if ((rc = ensure_scratch(ctx, &ctx->scr_chi2, &ctx->scr_chi2_n, (size_t)B, st))) return rc;
        chi2 = ctx->scr_chi2;
    }
    if (iso) {      // interpolation + parameter assembly in one kernel
        if ((rc = launch_mist_interp(ctx, tv, B, brackets, derived, nullptr, st, ctx->scr_pars, chi2,
                                     flags->mistdif))) return rc;
    } else {
        if ((rc = launch_build_pars(ctx, tv, nullptr, flags->mistdif, B, ctx->scr_pars, chi2, st))) return rc;
    }
    if (flags->spec) {
        SpecArgs sa{ctx->scr_pars, flags->slot, flags->modpoly, flags->n_pc, nullptr, chi2};
        sa.star_slots = star_slot;                     // per-row slots: every row against its own star's spectrum
        if ((rc = launch_spec(ctx, sa, B, st))) return rc;
    }
    const int agew = iso && flags->ageweight;
    if (ctx->npeers && !(flags->phot && ctx->precision >= MS_PREC_TC)) {
        ms_set_error("lnL peers are set (ms_set_lnl_peers) but this batch does not end in the tensor-core photometry kernel");
        return MS_ESTATE;
    }
    if (flags->phot) {   // photometry is the last term: it also writes lnL
        PhotArgs pa{ctx->scr_pars, star_slot, nullptr, chi2};
        pa.lnL = lnL; pa.ageweight = agew; pa.slot0 = flags->slot;
        return run_phot(ctx, pa, B, st);
    }
    return launch_finalize(ctx, ctx->scr_pars, chi2, agew, B, lnL, st);
}

static int check_batch_args(const ms_ctx *ctx, const ms_flags *flags, int64_t B, int ndim, const int32_t *colmap) {
    if (!ctx || !flags || !colmap || B < 0 || ndim < 1 || ndim > MS_NPARAM) {
        ms_set_error("bad arguments (ctx / flags / colmap NULL, B < 0 or ndim outside 1..%d)", MS_NPARAM);
        return MS_EINVAL;
    }
    for (int c = 0; c < ndim; ++c)
        if (colmap[c] < 0 || colmap[c] >= MS_NPARAM) {
            ms_set_error("colmap[%d] = %d out of range", c, colmap[c]);
            return MS_EINVAL;
        }
    if (flags->n_pc < 0 || flags->n_pc > MS_MAX_PC) { ms_set_error("n_pc out of range"); return MS_EINVAL; }
    if (flags->slot < 0) { ms_set_error("flags->slot is negative"); return MS_EINVAL; }
    return MS_OK;
}

extern "C" int ms_lnlike_batch_host(ms_ctx *ctx, const ms_flags *flags, const double *theta_host,
                                    int64_t B, int ndim, const int32_t *colmap, const double *fixed,
                                    double *lnL_host, double *derived_host) {
    int rc;
    if ((rc = check_batch_args(ctx, flags, B, ndim, colmap))) return rc;       // before any allocation
    if (B == 0) return MS_OK;
    if (!theta_host || !lnL_host) { ms_set_error("bad arguments"); return MS_EINVAL; }
    MS_CUDA(cudaSetDevice(ctx->device));
    if ((rc = ensure_scratch(ctx, &ctx->scr_theta, &ctx->scr_theta_n, (size_t)B * ndim, ctx->own_stream))) return rc;
    if ((rc = ensure_scratch(ctx, &ctx->scr_out, &ctx->scr_out_n, (size_t)B * (1 + MS_NDERIVED), ctx->own_stream))) return rc;
    double *lnl_d = ctx->scr_out, *der_d = derived_host ? ctx->scr_out + B : nullptr;
    // Streaming form (photometry-only isochrone batches on the tensor-core path, no derived block): ONE launch of the
    // fused kernel starts as soon as the first rows have landed; theta keeps arriving in chunks on the copy stream, each
    // followed by a 4-byte copy that advances a device counter the kernel's prologue waits on before it touches a pair's
    // rows, and lnL is stored straight into the caller's buffer when that is pinned, device-mapped host memory
    // (PCIe posted writes of 256 B per warp).  No per-chunk launches (each costs a non-overlapped first prologue and a
    // tail), no device->host copy behind the last kernel.
    {
        ThetaView tv0;
        identity_view(tv0, nullptr, ndim);
        if (fixed) for (int i = 0; i < MS_NPARAM; ++i) tv0.fixed[i] = fixed[i];
        for (int c = 0; c < ndim; ++c) tv0.col[colmap[c]] = c;
        const bool iso = tv0.col[MS_P_EEP] >= 0 || tv0.fixed[MS_P_EEP] == tv0.fixed[MS_P_EEP];
        const int64_t wave = (int64_t)ctx->sm_count * 256;
        if (iso && flags->phot && !flags->spec && !derived_host && ctx->fuse_interp && ctx->precision >= MS_PREC_TC &&
            B >= 4 * wave && B < (int64_t)1 << 31 && phot_tc_can_fuse(ctx) && !ctx->npeers) {
            const int kChunks = 4;
            if (!ctx->rows_ready_dev) {
                MS_CUDA(cudaMalloc((void **)&ctx->rows_ready_dev, sizeof(int)));
                MS_CUDA(cudaHostAlloc((void **)&ctx->rows_ready_host, (kChunks + 2) * sizeof(int), cudaHostAllocDefault));
            }
            while (ctx->ev_pool.size() < 2) {
                cudaEvent_t e;
                MS_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                ctx->ev_pool.push_back(e);
            }
            auto fail = [&](int code) {
                cudaStreamSynchronize(ctx->h2d_stream); cudaStreamSynchronize(ctx->own_stream);
                return code;
            };
            // lnL straight into the caller's buffer when the device can address it
            double *lnl_target = lnl_d;
            cudaPointerAttributes at;
            if (cudaPointerGetAttributes(&at, lnL_host) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer)
                lnl_target = static_cast<double *>(at.devicePointer);
            else
                cudaGetLastError();
            if (flags->slot >= ctx->nslots || flags->slot >= (int)ctx->stars.size() || !ctx->stars[flags->slot].set) {
                ms_set_error("star slot %d has no photometry loaded (ms_set_star)", flags->slot);
                return MS_ESTATE;
            }
            // In-place form: theta is pinned, device-mapped host memory too -> the kernel reads it where it is (its
            // prologue stages a tile's rows through shared memory with coalesced loads, so every byte crosses PCIe once)
            // and writes lnL where it belongs.  No copies, no counter, one launch.
            if (lnl_target != lnl_d && (reinterpret_cast<uintptr_t>(theta_host) & 15) == 0 && phot_tc_can_stage(ctx, ndim) &&
                cudaPointerGetAttributes(&at, theta_host) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) {
                ThetaView tv = tv0;
                tv.theta = static_cast<const double *>(at.devicePointer);
                PhotArgs pa{nullptr, nullptr, nullptr, nullptr};
                pa.lnL = lnl_target; pa.ageweight = flags->ageweight; pa.slot0 = flags->slot;
                pa.stage_theta = true;
                if ((rc = launch_phot_tc_fused(ctx, tv, pa, nullptr, flags->mistdif, B, ctx->own_stream))) return fail(rc);
                MS_CUDA(cudaStreamSynchronize(ctx->own_stream));
                return MS_OK;
            }
            cudaGetLastError();
            // upload plan: one wave first (the kernel starts after 1.8 MB), then equal chunks
            std::vector<int64_t> cuts{0, wave};
            const int64_t step = ((B - wave + kChunks - 2) / (kChunks - 1) + 255) / 256 * 256;
            while (cuts.back() < B) cuts.push_back(cuts.back() + step < B ? cuts.back() + step : B);
            cudaError_t e = cudaMemsetAsync(ctx->rows_ready_dev, 0, sizeof(int), ctx->h2d_stream);
            auto upload = [&](size_t k) {
                const int64_t b0 = cuts[k], nb = cuts[k + 1] - b0;
                ctx->rows_ready_host[k] = (int)cuts[k + 1];
                if (e == cudaSuccess)
                    e = cudaMemcpyAsync(ctx->scr_theta + b0 * ndim, theta_host + b0 * ndim, (size_t)nb * ndim * sizeof(double),
                                        cudaMemcpyHostToDevice, ctx->h2d_stream);
                if (e == cudaSuccess)
                    e = cudaMemcpyAsync(ctx->rows_ready_dev, ctx->rows_ready_host + k, sizeof(int), cudaMemcpyHostToDevice,
                                        ctx->h2d_stream);
            };
            // EVERY upload is queued before the launch: a tool that makes launches synchronous (ncu, compute-sanitizer,
            // CUDA_LAUNCH_BLOCKING) would otherwise leave the kernel waiting for copies the host has not queued yet.
            // Few chunks, so that queueing them delays the launch by tens of microseconds only.
            for (size_t k = 0; k + 1 < cuts.size(); ++k) {
                upload(k);
                if (k == 0 && e == cudaSuccess) e = cudaEventRecord(ctx->ev_pool[0], ctx->h2d_stream);
            }
            if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->own_stream, ctx->ev_pool[0], 0);
            if (e != cudaSuccess) { ms_set_error("host entry (streaming): %s", cudaGetErrorString(e)); return fail(MS_ECUDA); }
            ThetaView tv = tv0;
            tv.theta = ctx->scr_theta;
            PhotArgs pa{nullptr, nullptr, nullptr, nullptr};
            pa.lnL = lnl_target; pa.ageweight = flags->ageweight; pa.slot0 = flags->slot;
            pa.rows_ready = ctx->rows_ready_dev;
            if ((rc = launch_phot_tc_fused(ctx, tv, pa, nullptr, flags->mistdif, B, ctx->own_stream))) return fail(rc);
            if (lnl_target == lnl_d) {
                e = cudaMemcpyAsync(lnL_host, lnl_d, (size_t)B * sizeof(double), cudaMemcpyDeviceToHost, ctx->own_stream);
                if (e != cudaSuccess) { ms_set_error("host entry (streaming): %s", cudaGetErrorString(e)); return fail(MS_ECUDA); }
            }
            MS_CUDA(cudaStreamSynchronize(ctx->own_stream));
            MS_CUDA(cudaStreamSynchronize(ctx->h2d_stream));
            return MS_OK;
        }
    }
    // Chunked three-stream pipeline: the host->device copy of chunk k+1 and the device->host
    // copy of chunk k-1 overlap the kernels of chunk k (pinned host buffers make the copies
    // truly asynchronous; pageable ones still work, serialised by the driver).
    const int64_t kMaxChunks = 8, kMinChunk = 32768;
    int64_t chunk = (B + kMaxChunks - 1) / kMaxChunks;
    if (chunk < kMinChunk) chunk = kMinChunk;
    {   // whole waves of the persistent phot kernel: one CTA per SM, 256 points per CTA iteration
        const int64_t wave = (int64_t)ctx->sm_count * 256;
        chunk = (chunk + wave - 1) / wave * wave;
    }
    // chunk boundaries: a short first chunk (one wave) so the kernels start after a small copy, a short last chunk
    // (what is left, at most one chunk) so little is left to copy back after the last kernel
    std::vector<int64_t> cuts{0};
    {
        const int64_t wave = (int64_t)ctx->sm_count * 256;
        if (B > chunk + wave) cuts.push_back(wave);
        while (B - cuts.back() > chunk) cuts.push_back(cuts.back() + chunk);
        cuts.push_back(B);
    }
    const int nchunks = (int)cuts.size() - 1;
    while ((int)ctx->ev_pool.size() < 2 * nchunks) {
        cudaEvent_t e;
        MS_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->ev_pool.push_back(e);
    }
    // pre-size the per-batch scratch once so the chunk calls never reallocate mid-pipeline
    if ((rc = ensure_scratch(ctx, &ctx->scr_pars, &ctx->scr_pars_n, (size_t)chunk * MS_NPARS, ctx->own_stream))) return rc;
    if ((rc = ensure_scratch(ctx, &ctx->scr_chi2, &ctx->scr_chi2_n, (size_t)chunk, ctx->own_stream))) return rc;
    // on an error in the middle of the pipeline: let the copies already queued into caller buffers finish
    // before returning, so the caller may free them
    auto fail = [&](int code) {
        cudaStreamSynchronize(ctx->h2d_stream); cudaStreamSynchronize(ctx->own_stream);
        cudaStreamSynchronize(ctx->d2h_stream);
        return code;
    };
#define MS_CUDA_PIPE(call)                                                                       \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess) {                                                                 \
            ms_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));   \
            return fail(MS_ECUDA);                                                               \
        }                                                                                        \
    } while (0)
    for (int k = 0; k < nchunks; ++k) {
        const int64_t b0 = cuts[k], nb = cuts[k + 1] - cuts[k];
        cudaEvent_t up = ctx->ev_pool[2 * k], done = ctx->ev_pool[2 * k + 1];
        MS_CUDA_PIPE(cudaMemcpyAsync(ctx->scr_theta + b0 * ndim, theta_host + b0 * ndim,
                                     (size_t)nb * ndim * sizeof(double), cudaMemcpyHostToDevice, ctx->h2d_stream));
        MS_CUDA_PIPE(cudaEventRecord(up, ctx->h2d_stream));
        MS_CUDA_PIPE(cudaStreamWaitEvent(ctx->own_stream, up, 0));
        if ((rc = ms_lnlike_batch(ctx, flags, ctx->scr_theta + b0 * ndim, nullptr, nb, ndim, colmap, fixed,
                                  lnl_d + b0, der_d ? der_d + b0 * MS_NDERIVED : nullptr, nullptr,
                                  ctx->own_stream))) return fail(rc);
        MS_CUDA_PIPE(cudaEventRecord(done, ctx->own_stream));
        MS_CUDA_PIPE(cudaStreamWaitEvent(ctx->d2h_stream, done, 0));
        MS_CUDA_PIPE(cudaMemcpyAsync(lnL_host + b0, lnl_d + b0, (size_t)nb * sizeof(double), cudaMemcpyDeviceToHost,
                                     ctx->d2h_stream));
        if (derived_host)
            MS_CUDA_PIPE(cudaMemcpyAsync(derived_host + b0 * MS_NDERIVED, der_d + b0 * MS_NDERIVED,
                                         (size_t)nb * MS_NDERIVED * sizeof(double), cudaMemcpyDeviceToHost,
                                         ctx->d2h_stream));
    }
#undef MS_CUDA_PIPE
    MS_CUDA(cudaStreamSynchronize(ctx->d2h_stream));
    return MS_OK;
}

// Pre-size every per-batch scratch buffer for batches of up to max_batch rows (and, with a spectral net, its
// chunk buffers) so that no later call allocates; lock != 0 also pins them: a call that would need more then
// fails with MS_ESTATE instead of freeing memory that a captured CUDA graph still replays into.  lock == 0
// (with any max_batch, 0 included) lifts the pin.
extern "C" int ms_reserve(ms_ctx *ctx, int64_t max_batch, int lock) {
    if (!ctx || max_batch < 0) { ms_set_error("bad arguments"); return MS_EINVAL; }
    MS_CUDA(cudaSetDevice(ctx->device));
    ctx->scratch_locked = false;
    int rc;
    const size_t B = (size_t)max_batch;
    if (B > 0) {
        cudaStream_t st = nullptr;
        if ((rc = ensure_scratch(ctx, &ctx->scr_pars, &ctx->scr_pars_n, B * MS_NPARS, st))) return rc;
        if ((rc = ensure_scratch(ctx, &ctx->scr_chi2, &ctx->scr_chi2_n, B, st))) return rc;
        if ((rc = ensure_scratch(ctx, &ctx->scr_chi2g, &ctx->scr_chi2g_n, B, st))) return rc;
        if (ctx->spec.w0 && (rc = spec_reserve(ctx, max_batch))) return rc;
    }
    ctx->scratch_locked = lock != 0;
    return MS_OK;
}
