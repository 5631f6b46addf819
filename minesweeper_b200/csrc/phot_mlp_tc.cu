// Kernel (2), tensor-core form (MS_PREC_TC) -- placeholder until the tcgen05 kernel lands.
#include "common.cuh"

int phot_tc_prepare(ms_ctx *ctx) {
    (void)ctx;
    ms_set_error("MS_PREC_TC photometric kernel not available in this build");
    return MS_ESTATE;
}
int launch_phot_tc(ms_ctx *ctx, const PhotArgs &a, int64_t B, cudaStream_t st) {
    (void)ctx; (void)a; (void)B; (void)st;
    ms_set_error("MS_PREC_TC photometric kernel not available in this build");
    return MS_ESTATE;
}
