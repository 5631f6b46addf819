// Shared declarations for the minesweeper_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <vector>
#include "../../include/minesweeper_b200.h"

#define MS_NCOL 9          // GenMIST.output columns (fastMISTmod.py:84-87)
#define MS_CP32 12         // padded row of the f32 table: 48 B = 3 x 16 B
#define MS_CP64 10         // padded row of the f64 table: 80 B = 5 x 16 B

void ms_set_error(const char *fmt, ...);

#define MS_CUDA(call)                                                                  \
    do {                                                                               \
        cudaError_t e_ = (call);                                                       \
        if (e_ != cudaSuccess) {                                                       \
            ms_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call,                  \
                         cudaGetErrorString(e_));                                      \
            return MS_ECUDA;                                                           \
        }                                                                              \
    } while (0)

// ---- device-resident model ------------------------------------------------
struct GridDev {
    int ne = 0, nm = 0, nf = 0, na = 0;
    double *axes = nullptr;        // eep | mass | feh | afe, concatenated
    float *tab32 = nullptr;        // [na][nf][nm][ne][MS_CP32]; col 0 NaN = vertex absent
    double *tab64 = nullptr;       // [na][nf][nm][ne][MS_CP64]
};

struct PhotNetDev {
    int nb = 0, d_in = 0, h = 0;
    // raw f32 weights (all modes)
    float *w1 = nullptr, *b1 = nullptr, *w2 = nullptr, *b2 = nullptr, *w3 = nullptr, *b3 = nullptr;
    double xmin[8], xscale[8];     // scaled = (x - xmin) * xscale - 0.5
    // tensor-core packing (MS_PREC_TC): see phot_mlp_tc.cu
    void *tc_pack = nullptr;
    size_t tc_pack_bytes = 0;
};

struct SpecNetDev {
    int d_in = 0, h = 0, npix = 0;
    float *w0 = nullptr, *b0 = nullptr, *w1 = nullptr, *b1 = nullptr, *w2 = nullptr, *b2 = nullptr;
    double xmin[8], xscale[8];
    double *wave = nullptr;        // native wavelength [npix]
    double lnw0 = 0, dlnw = 0;     // log-uniform grid: ln w[i] = lnw0 + i * dlnw
    double resolution = 0, leak = 0.01;
    int logt_input = 0;
    // theta-independent resampling maps of the vsini stage (spec_smooth.cu)
    int n1 = 0;                    // pow2ceil(npix)
    int *a1_idx = nullptr; float *a1_frac = nullptr;       // native -> n1 grid
    int *a2_idx = nullptr; float *a2_frac = nullptr;       // n1 grid -> native
    double *a1_fracd = nullptr, *a2_fracd = nullptr;
    double *xms = nullptr;         // device copy of xmin[8] | xscale[8]
    float *rot_tab = nullptr;      // rotational transfer S(u) on a uniform u grid (f32 mode)
    float2 *twiddle32 = nullptr;   // exp(-2 pi i k / n1), k < n1/2
    double2 *twiddle64 = nullptr;
    int twiddle_n = 0;
    float2 *twlev32 = nullptr;     // per-level compact tables: level l at offset 2^(l-1), entry k = exp(-2 pi i k / 2^l), k < 2^(l-1)
    void *tc_pack = nullptr;
};

struct StarDev {
    bool set = false;
    int n_obs = 0;
    double *obs_wave = nullptr, *obs_lnwave = nullptr, *obs_flux = nullptr, *obs_err2 = nullptr;
    double *cheb_x = nullptr;      // wavelength rescaled to [-1, 1] (fitutils.py:13-14)
    double wmin = 0, wmax = 0;
    // LSF mode (ms_set_lsf): lam[nx] warped wavelengths, (idx, frac)[nx] resampling map from the native grid
    int lsf_nx = 0;
    double *lsf_lam = nullptr, *lsf_frac = nullptr; int *lsf_idx = nullptr;
    double lsf_xps = 0;
};

// device-side view of one star slot's observed spectrum (table ctx->stars_dev, indexed by slot)
struct StarView {
    int n_obs, pad_;
    const double *obs_wave, *obs_lnwave, *obs_flux, *obs_err2, *cheb_x;
    double wmin, wmax;
    // wavelength-dependent LSF (array Inst_R, ms_set_lsf): warped-grid map, nx = 0 when unset
    int lsf_nx, pad2_;
    const double *lsf_lam, *lsf_frac; const int *lsf_idx;
    double lsf_xps;
};

struct ms_ctx {
    int device = 0;
    int precision = MS_PREC_F32;
    int fuse_interp = 1;           // ms_set_fusion
    double *peers[MS_MAX_PEERS] = {nullptr}; int npeers = 0; int64_t peer_off = 0;   // ms_set_lnl_peers
    bool scratch_locked = false;   // ms_reserve(..., lock = 1): per-batch scratch may no longer grow
    int sm_count = 148;
    GridDev grid;
    PhotNetDev phot;
    SpecNetDev spec;
    // photometric observations for all star slots: [nslots][nb] (mag, err^2; NaN mag = unobserved)
    int nslots = 0;
    double *obs_mag = nullptr, *obs_err2 = nullptr;
    std::vector<StarDev> stars;
    StarView *stars_dev = nullptr; int stars_dev_n = 0;      // device copy of the spectrum views, refreshed by ms_set_star
    int spec_gemm_ts = 1;          // tc modes: output layer of the spectral net with the activations resident in tensor memory
    int spec_fast = 1;             // f32-class modes: fast broadening kernel (ms_set_spec_fast); 0 = generic kernel only
    int *scr_fb = nullptr; size_t scr_fb_bytes = 0;          // fallback row list of the fast broadening kernel
    // scratch (grown on demand, stream-ordered use on one stream at a time)
    double *scr_derived = nullptr; size_t scr_derived_n = 0;
    double *scr_pars = nullptr;    size_t scr_pars_n = 0;
    double *scr_chi2 = nullptr;    size_t scr_chi2_n = 0;
    double *scr_chi2g = nullptr;   size_t scr_chi2g_n = 0;   // chi-square carried between band groups (phot_mlp_tc.cu)
    void *scr_flux = nullptr;      size_t scr_flux_bytes = 0;
    void *scr_hid = nullptr;       size_t scr_hid_bytes = 0;
    double *scr_theta = nullptr;   size_t scr_theta_n = 0;     // for the host-buffer entry point
    double *scr_out = nullptr;     size_t scr_out_n = 0;
    int32_t *colmap_dev = nullptr; double *fixed_dev = nullptr;
    double *ptab_x = nullptr, *ptab_cdf = nullptr; int ptab_n = 0;      // ms_set_prior_table
    cudaStream_t own_stream = nullptr;           // compute stream of the host-buffer entry point
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
    std::vector<cudaEvent_t> ev_pool;            // chunk pipeline events
    int *rows_ready_dev = nullptr;               // streaming host entry: rows of theta uploaded so far (device int)
    int *rows_ready_host = nullptr;              // pinned staging of the values copied into it, one per upload chunk
    int64_t launches = 0;
    // live timing (ms_profile_*)
    bool prof_on = false;
    struct ProfRec { int kind; cudaEvent_t e0, e1; };
    std::vector<ProfRec> prof;
};

// RAII: brackets the launches issued in its scope with two events on `st`.
struct KernelTimer {
    ms_ctx *ctx; cudaStream_t st; cudaEvent_t e0 = nullptr, e1 = nullptr; int kind;
    KernelTimer(ms_ctx *c, int k, cudaStream_t s) : ctx(c), st(s), kind(k) {
        if (!ctx->prof_on) return;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, st);
    }
    ~KernelTimer() {
        if (!e0) return;
        cudaEventRecord(e1, st);
        ctx->prof.push_back({kind, e0, e1});
    }
};

// Canonical per-point parameter block pars[B][MS_NPARS] (f64) that the MLP /
// broadening kernels read: built from theta, fixed parameters and the MIST
// prediction exactly as likelihood.lnlikefn assembles specpars / photpars
// (likelihood.py:121-160).  Teff = +inf marks a failed interpolation.
#define MS_NPARS 20
enum { PV_TEFF = 0, PV_LOGG, PV_FEH, PV_AFE, PV_VRAD, PV_VROT, PV_VMIC, PV_INSTR,
       PV_LOGR, PV_DIST, PV_AV, PV_AGEWGT, PV_PC0 = 12 };

// ---- kernel launchers (one per .cu) ------------------------------------------
// theta: [B][ld] f64; col4 = column index of (EEP, mass, feh, afe) in theta or -1 -> fixed[]
struct ThetaView {
    const double *theta; int ld;
    int col[MS_NPARAM];            // theta column for each canonical parameter, -1 = not sampled
    double fixed[MS_NPARAM];       // value when not sampled (NaN = absent)
};

// pars / chi2 (optional): also emit the per-point parameter block and zero chi2 (fused
// form of launch_build_pars for isochrone fits)
int launch_mist_interp(ms_ctx *ctx, const ThetaView &tv, int64_t B, int32_t *brackets,
                       double *derived, uint8_t *ok, cudaStream_t st, double *pars = nullptr,
                       double *chi2 = nullptr, int mistdif = 1);

// pars[B][MS_NPARS] from theta / fixed / derived (derived may be nullptr when EEP is
// not sampled); chi2[B] = 0 for live rows.
int launch_build_pars(ms_ctx *ctx, const ThetaView &tv, const double *derived, int mistdif,
                      int64_t B, double *pars, double *chi2, cudaStream_t st);
// repack the standalone-API inputs photpars[B][7] / specpars[B][8+n_pc] into pars
int launch_repack_pars(ms_ctx *ctx, const double *in, int ncol, int kind, int64_t B, double *pars,
                       cudaStream_t st);
// lnL = -0.5 chi2 (+ log Agewgt); failed rows -> -inf  (likelihood.py:107-119,168-176,219)
int launch_finalize(ms_ctx *ctx, const double *pars, const double *chi2, int ageweight, int64_t B,
                    double *lnL, cudaStream_t st);

// photometry: writes mags[B][nb] and / or accumulates chi2[B] += chi2_phot
struct PhotArgs {
    const double *pars;            // [B][MS_NPARS]
    const int32_t *star_slot;      // [B] or nullptr
    double *mags;                  // [B][nb] or nullptr
    double *chi2;                  // [B] or nullptr
    // fused finalisation (likelihood.py:168-176, 219): when lnL is set the kernel writes
    // lnL = -0.5 (chi2[p] + chi2_phot) + (ageweight ? log Agewgt : 0), -inf for failed rows,
    // and chi2 is only read
    double *lnL = nullptr;
    int ageweight = 0;
    int slot0 = 0;                 // star slot of every row when star_slot is nullptr (ms_flags.slot)
    // streaming host entry (fused tensor-core form only): theta rows [0, *rows_ready) have landed on the device; the
    // kernel waits for a pair's rows before touching them.  nullptr: every row is there.
    const int *rows_ready = nullptr;
    // fused form: fetch each tile's rows of theta through shared memory with coalesced loads (theta in pinned HOST
    // memory, read in place by the host entry); device-resident theta is read directly
    bool stage_theta = false;
};
int launch_phot(ms_ctx *ctx, const PhotArgs &a, int64_t B, cudaStream_t st);
int phot_tc_prepare(ms_ctx *ctx);          // pack weights for the tensor-core kernel
int launch_phot_tc(ms_ctx *ctx, const PhotArgs &a, int64_t B, cudaStream_t st);
// fused kernels (1) + (2): the prologue warpgroup of the tensor-core kernel interpolates the grid itself
bool phot_tc_can_fuse(const ms_ctx *ctx);
bool phot_tc_can_stage(const ms_ctx *ctx, int ld);
int launch_phot_tc_fused(ms_ctx *ctx, const ThetaView &tv, const PhotArgs &a, double *derived, int mistdif,
                         int64_t B, cudaStream_t st);

struct SpecArgs {
    const double *pars;            // [B][MS_NPARS]
    int star_slot;                 // slot of every row when star_slots is nullptr
    int modpoly, n_pc;
    double *flux;                  // [B][n_obs] or nullptr
    double *chi2;                  // [B] or nullptr: chi2[i] += chi2_spec
    const int32_t *star_slots = nullptr;   // [B] per-row slots (catalog fits with spectra), device
};
int launch_spec(ms_ctx *ctx, const SpecArgs &a, int64_t B, cudaStream_t st);
int spec_prepare(ms_ctx *ctx, const ms_spec_net_desc *d);
int spec_reserve(ms_ctx *ctx, int64_t max_batch);      // allocate the chunk buffers of the spectrum path now
int spec_tc_prepare(ms_ctx *ctx);           // pack W2 for the tensor-core output GEMM
int spec_tc_reserve(ms_ctx *ctx, int64_t chunk, int nchunks);   // operand-image scratch of the tensor-core GEMMs: nchunks chunks of `chunk` points
bool spec_tc_would_grow(const ms_ctx *ctx, int64_t chunk, int nchunks);
int spec_tc_hidden(ms_ctx *ctx, const double *pars, int64_t B, int64_t chunk, cudaStream_t st);   // hidden layers of a super-batch (1: not available)
int spec_tc_out(ms_ctx *ctx, int64_t ci, int64_t chunk, int64_t nb, float *flux, cudaStream_t st);  // output layer of its chunk ci
void spec_tc_release(ms_ctx *ctx);
int launch_spec_mlp_tc(ms_ctx *ctx, const double *pars, int64_t B, float *flux, cudaStream_t st);


// Per-batch scratch of a ctx: grown on demand (free + malloc), which is only legal while nothing else can
// still hold the old pointer.  Two guards: (a) a stream that is capturing a CUDA graph must not
// see cudaFree / cudaMalloc, and the captured graph would bake the pointer in -- growth during capture is
// refused with MS_ESTATE; (b) ms_reserve(ctx, max_batch, lock = 1) pre-sizes every buffer and pins them:
// later calls that would need more fail loudly instead of freeing memory a captured graph replays into.
template <typename T> int ensure_scratch(ms_ctx *ctx, T **p, size_t *have, size_t need, cudaStream_t st) {
    if (*have >= need) return MS_OK;
    if (ctx->scratch_locked) {
        ms_set_error("batch needs %zu scratch elements but the ctx was reserved for less (ms_reserve): "
                     "reserve a larger max_batch before capturing graphs", need);
        return MS_ESTATE;
    }
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone) {
        ms_set_error("scratch would be reallocated while the stream is capturing a CUDA graph: run one eager "
                     "warm-up batch of the same size or call ms_reserve first");
        return MS_ESTATE;
    }
    if (*p) cudaFree(*p);
    *p = nullptr; *have = 0;
    size_t n = need + need / 4;
    if (cudaMalloc((void **)p, n * sizeof(T)) != cudaSuccess) {
        cudaGetLastError();
        ms_set_error("cudaMalloc of %zu bytes failed", n * sizeof(T));
        return MS_ENOMEM;
    }
    *have = n;
    return MS_OK;
}

// ---- device helpers -------------------------------------------------------------
__device__ __forceinline__ double ms_inf() { return __longlong_as_double(0x7ff0000000000000LL); }
__device__ __forceinline__ double ms_nan() { return __longlong_as_double(0x7ff8000000000000LL); }

__device__ __forceinline__ double theta_get(const ThetaView &tv, int64_t i, int pid) {
    int c = tv.col[pid];
    return c >= 0 ? tv.theta[i * (int64_t)tv.ld + c] : tv.fixed[pid];
}
