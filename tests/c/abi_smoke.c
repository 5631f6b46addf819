/* Plain-C consumer of include/minesweeper_b200.h: builds a tiny MIST grid and photometric net on
 * the host, runs a batch through ms_lnlike_batch_host and prints lnL.  No Python, no torch.
 * Exit codes: 0 ok, 3 no GPU (ms_create returned MS_ENODEV, the expected outcome on a CPU box). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "../../include/minesweeper_b200.h"

int main(void) {
    enum { NE = 8, NM = 6, NF = 4, NA = 3, NB = 3, H = 16, D = 6, B = 5 };
    double eep[NE], mass[NM], feh[NF], afe[NA];
    for (int i = 0; i < NE; ++i) eep[i] = 202 + i;
    for (int i = 0; i < NM; ++i) mass[i] = 0.5 + 0.15 * i;
    for (int i = 0; i < NF; ++i) feh[i] = -1.0 + 0.5 * i;
    for (int i = 0; i < NA; ++i) afe[i] = -0.2 + 0.4 * i;
    const size_t nv = (size_t)NE * NM * NF * NA;
    double *values = malloc(nv * 9 * sizeof(double));
    uint8_t *valid = malloc(nv);
    for (size_t v = 0; v < nv; ++v) {
        valid[v] = 1;
        const double t = (double)v / nv;
        double row[9] = {9.5 + 0.3 * t, 1.0, 0.1 * t, 0.2 * t, 3.76 + 0.01 * t, -0.2, 0.1, 4.4 - 0.2 * t, 0.5 + 0.4 * t};
        for (int c = 0; c < 9; ++c) values[v * 9 + c] = row[c];
    }
    float *w1 = calloc(NB * H * D, 4), *b1 = calloc(NB * H, 4), *w2 = calloc(NB * H * H, 4), *b2 = calloc(NB * H, 4),
          *w3 = calloc(NB * H, 4), *b3 = calloc(NB, 4);
    for (int i = 0; i < NB * H * D; ++i) w1[i] = 0.1f * (float)sin(0.7 * i);
    for (int i = 0; i < NB * H * H; ++i) w2[i] = 0.1f * (float)cos(0.3 * i);
    for (int i = 0; i < NB * H; ++i) { w3[i] = 0.05f * (float)sin(1.1 * i); b1[i] = 0.01f * i; }
    double xmin[6] = {2500, -1, -4, -0.2, 0, 2}, xmax[6] = {15000, 5.5, 0.5, 0.6, 5, 5};
    ms_grid_desc g = {NE, NM, NF, NA, eep, mass, feh, afe, values, valid};
    ms_phot_net_desc p = {NB, D, H, w1, b1, w2, b2, w3, b3, xmin, xmax};
    ms_ctx *ctx = NULL;
    int rc = ms_create(&ctx, 0, MS_PREC_F32, &g, &p, NULL);
    if (rc == MS_ENODEV) { printf("no GPU: %s\n", ms_last_error()); return 3; }
    if (rc) { printf("ms_create failed %d: %s\n", rc, ms_last_error()); return 1; }
    double mag[NB] = {5.0, 4.8, 4.5}, err[NB] = {0.05, 0.05, 0.05};
    if ((rc = ms_set_star(ctx, 0, mag, err, NB, NULL, NULL, NULL, 0))) { printf("set_star %d\n", rc); return 1; }
    /* theta columns: Dist, Av, EEP, initial_[Fe/H], initial_[a/Fe], initial_Mass (fitstar.py:50-67 order) */
    int32_t colmap[6] = {MS_P_DIST, MS_P_AV, MS_P_EEP, MS_P_INIT_FEH, MS_P_INIT_AFE, MS_P_INIT_MASS};
    double fixed[MS_NPARAM];
    for (int i = 0; i < MS_NPARAM; ++i) fixed[i] = NAN;
    double theta[B][6] = {{10, 0.01, 203.5, -0.4, 0.1, 0.7}, {20, 0.05, 205.2, 0.2, 0.3, 1.0}, {30, 0.0, 300.0, -0.4, 0.1, 0.7},
                          {15, 0.02, 202.0, -1.0, -0.2, 0.5}, {15, 0.02, 208.9, 0.49, 0.59, 1.24}};
    ms_flags fl = {0, 1, 0, 1, 1, 0};
    double lnl[B], der[B][MS_NDERIVED];
    rc = ms_lnlike_batch_host(ctx, &fl, &theta[0][0], B, 6, colmap, fixed, lnl, &der[0][0]);
    if (rc) { printf("lnlike failed %d: %s\n", rc, ms_last_error()); return 1; }
    for (int i = 0; i < B; ++i) printf("lnL[%d] = %.6f  Teff = %.2f\n", i, lnl[i], pow(10.0, der[i][8]));
    if (!(isinf(lnl[2]) && lnl[2] < 0) || !isfinite(lnl[0]) || !isfinite(lnl[3])) { printf("unexpected values\n"); return 1; }
    printf("launches = %lld, %s\n", (long long)ms_launch_count(ctx), ms_version());
    ms_destroy(ctx);
    return 0;
}
