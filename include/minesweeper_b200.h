/* minesweeper_b200 -- C-ABI of the B200-native MINESweeper likelihood hot path.
 *
 * The reference (pacargile/MINESweeper) is pure Python and has no FFI; the
 * entry points below are what a binding for this path replaces, one batched
 * call per reference method (file:line relative to the reference checkout):
 *
 *   ms_create / ms_destroy   likelihood.__init__            likelihood.py:9-69
 *                            GenMIST.__init__/make_lib/lib_as_grid   fastMISTmod.py:28-98,110-125
 *                            GenMod._initphotnn/_initspecnn genmod.py:15-44
 *   ms_set_star              fitargs['obs_phot'|'obs_*_fit'] consumed at likelihood.py:188-195,207-210
 *   ms_mist_interp           GenMIST.getMIST                fastMISTmod.py:100-108
 *   ms_phot_sed              GenMod.genphot                 genmod.py:97-142
 *   ms_spec_model            GenMod.genspec                 genmod.py:58-95
 *   ms_lnlike_batch          likelihood.lnlikefn + lnlike   likelihood.py:92-181,183-219
 *   ms_lnlike_batch_host     same, host buffers (what a ctypes/cffi caller with NumPy arrays uses)
 *
 * Conventions
 *   - plain pointers and sizes; every array is caller-owned and row-major.
 *   - functions return 0 on success or a negative MS_E* code; they never
 *     throw.  ms_last_error() returns a message for the calling thread.
 *   - "device" pointers must be device-accessible on the ctx's GPU; "host"
 *     pointers are ordinary (preferably pinned) host memory.
 *   - all *_batch / *_interp / *_sed / *_model calls are stream-ordered on
 *     `stream` (a cudaStream_t passed as void*; NULL = legacy default stream)
 *     and do not synchronise, except ms_lnlike_batch_host which returns with
 *     its outputs complete.
 *   - out-of-grid / no-corner / zero-weight points follow the reference:
 *     lnL = -inf and the nine derived MIST values = +inf (likelihood.py:107-119);
 *     NaN residuals are dropped from the chi-square (nansum, :193,207).
 *   - a ctx owns per-batch scratch buffers: use it from one host thread and one stream at a
 *     time (calls on one stream are ordered; create one ctx per stream / thread for concurrency).
 *   - there is no CPU fallback: without a CUDA device ms_create fails with
 *     MS_ENODEV.
 */
#ifndef MINESWEEPER_B200_H
#define MINESWEEPER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MS_OK        0
#define MS_EINVAL   -1   /* bad argument */
#define MS_ENODEV   -2   /* no usable CUDA device */
#define MS_ECUDA    -3   /* CUDA runtime error (see ms_last_error) */
#define MS_ENOMEM   -4
#define MS_ESTATE   -5   /* model part or star slot not configured */

/* Arithmetic mode of the forward model.  Brackets, weights, the magnitude /
 * chi-square epilogue and lnL are IEEE f64 in every mode. */
#define MS_PREC_F64   0  /* f64 table + f64 MLP/FFT on CUDA cores: validation grade */
#define MS_PREC_F32   1  /* f32 table, f32 MLP/FFT on CUDA cores */
#define MS_PREC_TC    2  /* f32 table, MLP layers on tcgen05 tensor cores (split-fp16 x3, fp32 accumulate: fp32-class) */
#define MS_PREC_TC16  3  /* as MS_PREC_TC, but the photometric MLP uses single fp16 activations and tanh.approx
                            (the "bf16-class" mode: magnitudes to a stated ~1e-3 mag, about twice the throughput) */

/* canonical parameter ids for `colmap` (theta column -> parameter);
 * order = fitstar.py:50-67 */
enum {
    MS_P_TEFF = 0, MS_P_LOGG, MS_P_FEH, MS_P_AFE, MS_P_VRAD, MS_P_VROT, MS_P_VMIC,
    MS_P_INST_R, MS_P_LOGR, MS_P_DIST, MS_P_AV, MS_P_LOGA, MS_P_EEP, MS_P_INIT_FEH,
    MS_P_INIT_AFE, MS_P_INIT_MASS, MS_P_PC0 /* pc_i = MS_P_PC0 + i */
};
#define MS_MAX_PC     8
#define MS_NPARAM     (MS_P_PC0 + MS_MAX_PC)

/* columns of `derived[B, MS_NDERIVED]` = GenMIST.modpararr (fastMISTmod.py:55,60):
 * EEP, initial_Mass, initial_[Fe/H], initial_[a/Fe], log(Age), Mass, log(R),
 * log(L), log(Teff), [Fe/H], [a/Fe], log(g), Agewgt */
#define MS_NDERIVED   13

typedef struct ms_ctx ms_ctx;

/* Dense MIST grid (host pointers, copied to the device by ms_create).
 * values[na][nf][nm][ne][9] in GenMIST.output column order
 * (log_age, star_mass, log_R, log_L, log_Teff, [Fe/H], [a/Fe], log_g, Agewgt);
 * valid[na][nf][nm][ne] != 0 where the track table has that row. */
typedef struct {
    int32_t ne, nm, nf, na;
    const double *eep, *mass, *feh, *afe;     /* strictly increasing axes */
    const double *values;
    const uint8_t *valid;
} ms_grid_desc;

/* Photometric MLP stack, per band d_in -> h -> h -> 1, sigmoid activations.
 * w1[nb][h][d_in], b1[nb][h], w2[nb][h][h], b2[nb][h], w3[nb][h], b3[nb];
 * inputs (Teff, logg, feh, afe, Av, Rv)[:d_in] scaled (x-xmin)/(xmax-xmin)-0.5. */
typedef struct {
    int32_t nb, d_in, h;
    const float *w1, *b1, *w2, *b2, *w3, *b3;
    const double *xmin, *xmax;
} ms_phot_net_desc;

/* Spectral MLP d_in -> h -> h -> npix, leaky-ReLU(leak); w0[h][d_in], w1[h][h],
 * w2[npix][h].  wavelength[npix] must be log-uniform.  resolution is the value
 * handed to the R-smoothing as `inres` (lambda/sigma units). */
typedef struct {
    int32_t d_in, h, npix;
    const float *w0, *b0, *w1, *b1, *w2, *b2;
    const double *xmin, *xmax;
    const double *wavelength;
    double resolution;
    double leak;
    int32_t logt_input;       /* first label is log10(Teff) instead of Teff */
} ms_spec_net_desc;

/* Run flags = runbools of likelihood.py:16-20 plus the two fitargs switches. */
typedef struct {
    int32_t spec, phot, modpoly;     /* runbools[0..2] */
    int32_t ageweight;               /* fitargs['ageweight']  likelihood.py:23,168-176 */
    int32_t mistdif;                 /* fitargs['mistdif']    likelihood.py:27,127-132 */
    int32_t n_pc;                    /* number of blaze coefficients when modpoly */
    int32_t slot;                    /* star slot whose observations (ms_set_star) every row is compared with when
                                        the per-row star_slot array is NULL; 0 for a single-star ctx */
} ms_flags;

const char *ms_version(void);
const char *ms_last_error(void);

/* Any of grid / phot / spec may be NULL when that part is unused. */
int ms_create(ms_ctx **out, int device, int precision, const ms_grid_desc *grid,
              const ms_phot_net_desc *phot, const ms_spec_net_desc *spec);
void ms_destroy(ms_ctx *ctx);
int ms_precision(const ms_ctx *ctx);

/* Observations of one star (host pointers; copied).  obs_mag/obs_err[nb] in the
 * ctx's band order, NaN = band not observed (dropped, as the reference iterates
 * observed bands only).  obs_wave must be increasing; pass n_obs = 0 for none. */
int ms_set_star(ms_ctx *ctx, int star_slot, const double *obs_mag, const double *obs_err,
                int nb, const double *obs_wave, const double *obs_flux,
                const double *obs_eflux, int n_obs);

/* Wavelength-dependent line-spread function of a star (array-valued Inst_R: GenMod.genspec passes it through to the
 * predictor, genmod.py:74-77, which smooths with smoothspec(..., smoothtype='lsf'), smoothing.py:125-128, 412-514).
 * The reference's warped wavelength axis depends only on the net's grid and sigma(lambda), so it is prepared once on the
 * host (minesweeper_b200/lsf.py): lam[nx] = wavelengths of the nx = 2^k uniform points of the warped coordinate,
 * (idx, frac)[nx] = linear-interpolation map from the net's native pixels onto them, x_per_sigma = kernel width in the
 * warped coordinate.  While set, batches against this slot ignore the scalar Inst_R.  nx = 0 clears. */
int ms_set_lsf(ms_ctx *ctx, int star_slot, int nx, const double *lam, const int32_t *idx, const double *frac,
               double x_per_sigma);

/* Photometry of many stars at once (catalog mode): obs_mag / obs_err[nstars][nb] host arrays go to
 * slots first_slot .. first_slot + nstars - 1. */
int ms_set_stars_phot(ms_ctx *ctx, int first_slot, int nstars, const double *obs_mag, const double *obs_err,
                      int nb);

/* theta4[B][4] = (EEP, initial_Mass, initial_[Fe/H], initial_[a/Fe]) device f64.
 * brackets[B][4] int32 = digitize-1 per axis (may be NULL);
 * derived[B][13] f64 (may be NULL); ok[B] uint8 (may be NULL). */
int ms_mist_interp(ms_ctx *ctx, const double *theta4, int64_t B, int32_t *brackets,
                   double *derived, uint8_t *ok, void *stream);

/* photpars[B][7] = (Teff, logg, feh, afe, logR, Dist, Av) device f64 -> mags[B][nb]. */
int ms_phot_sed(ms_ctx *ctx, const double *photpars, int64_t B, double *mags, void *stream);

/* specpars[B][8 + n_pc] = (Teff, logg, feh, afe, Vrad, Vrot, Vmic, Inst_R, pc...) device
 * f64 -> flux[B][n_obs] on star_slot's obs_wave.  Inst_R is the un-multiplied
 * fit parameter (the 2.355 factor of genmod.py:74-75 is applied inside). */
int ms_spec_model(ms_ctx *ctx, int star_slot, const double *specpars, int n_pc,
                  int64_t B, double *flux, void *stream);

/* The batched likelihood.  theta[B][ndim] device f64; colmap[ndim] host int32
 * mapping each column to an MS_P_* id; fixed[MS_NPARAM] host f64 with NaN for
 * "not fixed" (fitargs['fixedpars']); star_slot[B] device int32 or NULL (= slot
 * flags->slot for every row).  A row whose slot lies outside the loaded slots gets
 * lnL = NaN (nothing is read out of bounds).  lnL[B] device f64; derived[B][13] and
 * brackets[B][4] may be NULL. */
int ms_lnlike_batch(ms_ctx *ctx, const ms_flags *flags, const double *theta,
                    const int32_t *star_slot, int64_t B, int ndim, const int32_t *colmap,
                    const double *fixed, double *lnL, double *derived, int32_t *brackets,
                    void *stream);

/* Same with HOST buffers; returns with the results in place (synchronises).  Photometry-only isochrone batches in the
 * tensor-core modes run as ONE launch of the fused kernel: when theta_host and lnL_host are pinned, device-mapped memory
 * (cudaHostAlloc / cudaHostRegister; 16-byte aligned, ndim <= 8) the kernel reads theta and writes lnL in place over PCIe,
 * otherwise theta is uploaded in chunks behind a device counter the kernel waits on and lnL is copied back.  Every other
 * mode copies theta in, runs and copies lnL (and derived when non-NULL) out in a chunked three-stream pipeline. */
int ms_lnlike_batch_host(ms_ctx *ctx, const ms_flags *flags, const double *theta_host,
                         int64_t B, int ndim, const int32_t *colmap, const double *fixed,
                         double *lnL_host, double *derived_host);

/* ---- prior side -------------------------------------------
 * ms_prior_transform   prior.priortrans                 prior.py:150-385
 * ms_lnprob_batch      prior.lnpriorfn + fitstar.lnprobfn   prior.py:387-639, fitstar.py:715-727 */
enum { MS_PRIOR_UNIFORM = 0,     /* p = lo, hi                        pv_uniform / default range */
       MS_PRIOR_GAUSSIAN,        /* p = loc, scale                    pv_gaussian                */
       MS_PRIOR_TGAUSSIAN,       /* p = lo, hi, loc, scale            pv_tgaussian, blaze coeffs */
       MS_PRIOR_EXP,             /* p = loc, scale                    pv_exp                     */
       MS_PRIOR_LOGUNIFORM,      /* p = a, b                          pv_loguniform              */
       MS_PRIOR_BETA,            /* p = a, b, loc, scale              pv_beta (scipy.stats.beta.ppf, prior.py:198-203) */
       MS_PRIOR_TABLE };         /* p = scale: scale * weighted quantile of the table of ms_set_prior_table
                                    (Galactic distance prior: 1000 AP.gal_ppf(u), prior.py:333-336)           */
typedef struct { int32_t kind; int32_t pad_; double p[4]; } ms_prior_col;

/* value a prior term applies to: an MS_P_* id (sampled / fixed / isochrone-predicted label) or */
enum { MS_V_LOGAGE = MS_NPARAM, MS_V_AGE, MS_V_PARALLAX, MS_V_END };
enum { MS_TERM_UNIFORM = 0,      /* p = lo, hi: -inf outside */
       MS_TERM_GAUSSIAN,         /* p = mu, sigma            */
       MS_TERM_TGAUSSIAN,        /* p = lo, hi, mu, sigma (MIST family: bounds and Gaussian) */
       MS_TERM_TGAUSSIAN_BOUNDS  /* p = lo, hi (spec / phot families apply the bounds only)  */ };
typedef struct { int32_t value_id; int32_t kind; double p[4]; } ms_prior_term;
#define MS_MAX_PRIOR_TERMS 32

/* Tabulated inverse CDF for MS_PRIOR_TABLE columns: x[n] ascending and cdf[n] (cdf[0] = 0, non-decreasing), host
 * arrays, copied; the transform is the linear interpolation np.interp(u, cdf, x) of Payne's weighted quantile
 * (advancedpriors.py:665-670).  One table per ctx. */
int ms_set_prior_table(ms_ctx *ctx, const double *x, const double *cdf, int n);

/* Physically motivated additive priors of advancedpriors.py, evaluated per point inside the ln-prior kernel:
 *   gal     AdvancedPriors.gal_lnprior(Dist / 1000) with coords = [] (thin + thick disk + halo number density along the
 *           sight line (l, b) given at construction; advancedpriors.py:410-663, prior.py:413-446)
 *   galage  + age_lnprior on 10^(log(Age) - 9) with the component membership probabilities (advancedpriors.py:776-833)
 *   vrot    vrot_lnprior (Kraft break / giant / dwarf sigmoid; advancedpriors.py:691-733, prior.py:453-466)
 *   alpha   alpha_lnprior (advancedpriors.py:672-688, prior.py:494-501)
 * age_*: min, max, mean, sigma (sigma <= 0: uniform between min and max). */
typedef struct {
    int32_t gal, galage, vrot, alpha;
    double Xp, Yp, Zp, sol_X, sol_Z;
    double age_thin[4], age_thick[4], age_halo[4];
    double vrot_dwarf[3], vrot_giant[3];             /* a, c, n */
    double minalpha;
} ms_adv_prior;

/* u[B][ndim] in the unit cube -> theta[B][ndim] (device pointers); cols[ndim] host. */
int ms_prior_transform(ms_ctx *ctx, const double *u, int64_t B, int ndim, const ms_prior_col *cols,
                       double *theta, void *stream);
/* out[p] = lnL[p] + lnprior(theta_p, derived_p) with the -inf short circuits of lnprobfn; lnL may
 * be NULL (then out = lnprior).  derived[B][13] from ms_lnlike_batch (NULL without isochrones);
 * imf != 0 adds the Kroupa IMF prior on initial_Mass. */
int ms_lnprob_batch(ms_ctx *ctx, const ms_flags *flags, const double *theta, int64_t B, int ndim,
                    const int32_t *colmap, const double *fixed, const double *derived,
                    const ms_prior_term *terms, int nterms, int imf, const double *lnL, double *out,
                    void *stream);
/* Same, with a per-star Gaussian parallax prior (catalog mode): star_slot[B] device int32 and
 * star_plx[nslots][2] = (parallax, error) in mas, device f64; adds -0.5 ((1000/Dist - plx)/err)^2. */
int ms_lnprob_batch_stars(ms_ctx *ctx, const ms_flags *flags, const double *theta, int64_t B, int ndim,
                          const int32_t *colmap, const double *fixed, const double *derived,
                          const ms_prior_term *terms, int nterms, int imf, const double *lnL, double *out,
                          const int32_t *star_slot, const double *star_plx, void *stream);

/* As ms_lnprob_batch_stars, plus the advanced priors (adv may be NULL). */
int ms_lnprob_batch_adv(ms_ctx *ctx, const ms_flags *flags, const double *theta, int64_t B, int ndim,
                        const int32_t *colmap, const double *fixed, const double *derived,
                        const ms_prior_term *terms, int nterms, int imf, const double *lnL, double *out,
                        const int32_t *star_slot, const double *star_plx, const ms_adv_prior *adv, void *stream);

/* ---- nested-sampling bookkeeping ---------------------------
 * Replaces the sequential part of dynesty's static sampler as driven by fitstar._runsampler
 * (fitstar.py:337-369): S stars x Q queued random-walk walkers, K live points per star.  All
 * pointers are device memory owned by the caller (see minesweeper_b200/sampler.py). */
#define MS_NS_TRAIL 7
typedef struct {
    int32_t S, K, Q, nd, nder, walks;
    int64_t maxiter, dead_cap;
    double dlogz;
    double *live_u, *live_lp, *live_row;              /* [S][K][nd], [S][K], [S][K][nd+nder]         */
    double *cur_u, *cur_lp, *cur_row, *prop_u;        /* walkers: [S*Q][nd], [S*Q], [S*Q][nd+nder]   */
    const double *prop_lp, *prop_theta, *prop_der;    /* outputs of the batched lnprob at prop_u     */
    int32_t *nacc;                                    /* [S*Q] accepted steps this round             */
    double *lmin, *sigma, *scale;                     /* [S], [S][nd][nd] Cholesky factor of the live-point covariance, [S] */
    double *logz, *hacc, *wmax, *wtot, *wsum, *wsq;   /* evidence / information / posterior moments  */
    int64_t *niter, *ncall;                           /* [S] iterations, likelihood calls            */
    int32_t *done;                                    /* [S] 0 running, 1 converged (dlogz), 2 finalised, 3 stopped at
                                                         maxiter, 4 finalised after maxiter             */
    double *dead;                                     /* [S][dead_cap][nd+nder+MS_NS_TRAIL] or NULL: the row, then
                                                         log(lk) log(vol) log(wt) h nc log(z) delta(log(z))  */
    uint64_t *rng_round;                              /* device counter of completed rounds (graph-replay safe RNG stream) */
    /* proposal: 0 = random walk from a live point (dynesty 'rwalk'), 1 = uniform in the bounding ellipsoid of the live
     * points (dynesty 'unif': walks must be 1).  center[S][nd] and ell_r[S] (radius of the ellipsoid in units of the
     * Cholesky factor `sigma`, enlargement included) are maintained by ms_ns_consume; they may be NULL when unif == 0. */
    double *center, *ell_r;
    int32_t unif;
} ms_ns_state;
/* step = walk step inside the current round (0 .. walks-1); the round number is read from rng_round */
int ms_ns_propose(ms_ctx *ctx, const ms_ns_state *s, uint64_t seed, uint64_t step, void *stream);
int ms_ns_accept(ms_ctx *ctx, const ms_ns_state *s, void *stream);
/* Fused forms of a walk step (bit-identical to the separate calls, one launch each):
 * ms_ns_propose_transform = ms_ns_propose + ms_prior_transform(prop_u -> prop_theta) with the column table `cols[nd]`;
 * ms_ns_lnprob_accept = ms_lnprob_batch_adv(prop_theta, prop_der, lnL -> prop_lp) + ms_ns_accept, lnL[S*Q] being the
 * ln likelihood of the proposals (ms_lnlike_batch on prop_theta); the other arguments as in ms_lnprob_batch_adv.
 * Both write through the `const` proposal pointers of the state (prop_theta, prop_lp). */
int ms_ns_propose_transform(ms_ctx *ctx, const ms_ns_state *s, uint64_t seed, uint64_t step,
                            const ms_prior_col *cols, void *stream);
int ms_ns_lnprob_accept(ms_ctx *ctx, const ms_ns_state *s, const ms_flags *flags, const int32_t *colmap,
                        const double *fixed, const ms_prior_term *terms, int nterms, int imf, const double *lnL,
                        const int32_t *star_slot, const double *star_plx, const ms_adv_prior *adv, void *stream);
/* consumes the queue, prepares the next round and advances rng_round */
int ms_ns_consume(ms_ctx *ctx, const ms_ns_state *s, uint64_t seed, void *stream);
/* adds the remaining live points of every star with done == 1 and marks it 2 */
int ms_ns_finalize(ms_ctx *ctx, const ms_ns_state *s, void *stream);
/* sigma[S][nd][nd] = lower Cholesky factor of the covariance of the point sets U[S][K][nd] (device f64): the proposal
 * shape dynesty's rwalk takes from the bounding ellipsoid; a covariance that is not positive definite gives the
 * diagonal of standard deviations.  The same code maintains ms_ns_state.sigma inside ms_ns_consume. */
int ms_ns_spread(ms_ctx *ctx, const double *U, int S, int K, int nd, double *sigma, void *stream);

/* In the tensor-core modes a photometry-only isochrone batch (likelihood.lnlikefn with runbools =
 * [spec off, phot on], reference likelihood.py:92-181) runs as ONE kernel: the MIST interpolation of
 * fastMISTmod.GenMIST.getMIST (fastMISTmod.py:100-108) is folded into the prologue of the photometric
 * MLP kernel and the per-point parameter block never exists in HBM.  Results are identical to the
 * two-kernel sequence up to rounding (same interpolation code).  Applied to batches of at least four tile pairs
 * per SM (smaller batches keep the two-kernel sequence, which is faster there).  ms_set_fusion(ctx, 0) selects
 * the two-kernel sequence always (used by the tests and to time kernel (1) on its own); the default is 1. */
int ms_set_fusion(ms_ctx *ctx, int enable);

/* f32-class modes: the broadening chain of genspec (smoothing.py:202-264, 517-598) normally runs in the fast
 * kernel (csrc/spec_fast.cu: in-place radix-8 FFT for the rotational stage, direct Gaussian sum for the
 * instrumental stage, two CTAs per SM) with the generic FFT kernel (csrc/spec_smooth.cu) taking the rows it declines.
 * ms_set_spec_fast(ctx, 0) sends every row through the generic kernel (used by the tests to compare the two); any other
 * value (default 1) enables the fast kernel. */
int ms_set_spec_fast(ms_ctx *ctx, int enable);

/* Multi-GPU gather fused into the likelihood kernel (the path's one collective is the gather of
 * lnL).  peers[0 .. n) are device pointers, valid in THIS process, to every rank's gather buffer (peer memory mapped
 * over NVLink, e.g. torch symmetric memory); while set, the tensor-core photometry kernel stores each lnL[p] also at
 * peers[r][offset + p] for every r, so the gathered vector materialises on all GPUs as the tiles finish -- no NCCL
 * kernel takes SMs from the persistent kernel and no separate copy runs.  The caller orders the ranks afterwards (a
 * barrier on the stream).  n = 0 clears.  Only MS_PREC_TC / TC16 photometry batches honour it; other paths return
 * MS_ESTATE while peers are set. */
#define MS_MAX_PEERS 8
int ms_set_lnl_peers(ms_ctx *ctx, const uint64_t *peers, int n, int64_t offset);

/* Scratch policy.  A ctx grows its per-batch scratch on demand (cudaFree + cudaMalloc inside the call), which
 * is illegal while `stream` captures a CUDA graph and unsafe once a captured graph has baked the pointers in.
 * ms_reserve pre-sizes every scratch buffer for batches of up to max_batch rows; with lock != 0 the buffers are
 * pinned and a later call that needs more returns MS_ESTATE instead of reallocating (callers that capture
 * ms_lnlike_batch / ms_lnprob_batch into graphs -- minesweeper_b200/sampler.py -- reserve first).  lock == 0
 * lifts the pin.  Growth attempted during stream capture is refused with MS_ESTATE either way. */
int ms_reserve(ms_ctx *ctx, int64_t max_batch, int lock);

/* Number of kernel launches issued by this ctx so far (bench bookkeeping). */
int64_t ms_launch_count(const ms_ctx *ctx);

/* Live per-kernel timing with CUDA events recorded on the launching stream.
 * ms_profile_enable(ctx, 1) starts (and clears) collection; ms_profile_read
 * synchronises and returns the summed device milliseconds and launch counts per
 * kernel family since then. */
enum { MS_K_INTERP = 0, MS_K_PARS, MS_K_PHOT, MS_K_SPEC_MLP, MS_K_SPEC_SMOOTH, MS_K_FINALIZE, MS_NKIND };
int ms_profile_enable(ms_ctx *ctx, int on);
int ms_profile_read(ms_ctx *ctx, double *ms_by_kind, int64_t *launches_by_kind);

#ifdef __cplusplus
}
#endif
#endif
