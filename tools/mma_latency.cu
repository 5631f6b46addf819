// Microbenchmark: time per tcgen05.mma (kind::f16, M = 128, K = 16) in a chain of MMAs that accumulate into ONE tensor-
// memory accumulator against the same number of MMAs alternating between TWO accumulators, for N = 64 / 128 / 256, with the
// A operand in shared memory (SS) or tensor memory (TS).  One CTA, one issuing warp; clock64 around issue + commit + wait.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../minesweeper_b200/csrc -o mma_latency mma_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"

template <int N, bool TS, int NACC>
__global__ void __launch_bounds__(128, 1) chain(long long *out, int nmma) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) unsigned long long bar_store;
    const uint32_t sb = smem_u32(smem), bar = smem_u32(&bar_store);
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_tmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = s_tmem;
    if (threadIdx.x < 32) {
        const uint64_t adesc = make_desc_sw128(sb), bdesc = make_desc_sw128(sb + 16384);
        const uint32_t idesc = make_idesc_f16(128, N);
        for (int rep = 0; rep < 3; ++rep) {
            const long long t0 = clock64();
            if (elect_one()) {
                for (int i = 0; i < nmma; ++i) {
                    const uint32_t d = tm + 256 + (NACC == 2 ? (i & 1) * N : 0);      // accumulators behind column 256
                    if (TS) mma_ts(d, tm + (i & 7) * 8, bdesc, i >= NACC, idesc);
                    else mma_ss(d, adesc, bdesc, i >= NACC, idesc);
                }
                tc_commit(bar);
            }
            __syncwarp();
            mbar_wait(bar, rep & 1);
            const long long t1 = clock64();
            if (threadIdx.x == 0) out[rep] = t1 - t0;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

template <int N, bool TS, int NACC> void run(const char *name, int nmma) {
    long long *d, h[3];
    cudaMalloc(&d, sizeof(h));
    cudaFuncSetAttribute(chain<N, TS, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    chain<N, TS, NACC><<<1, 128, 64 * 1024>>>(d, nmma);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-34s %4d MMAs: %7lld clk  = %6.1f clk per MMA   (%s)\n", name, nmma, h[2], (double)h[2] / nmma, cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    const int n = 240;
    run<64, true, 1>("TS N=64  one accumulator", n);   run<64, true, 2>("TS N=64  two accumulators", n);
    run<128, true, 1>("TS N=128 one accumulator", n);  run<128, true, 2>("TS N=128 two accumulators", n);
    run<64, false, 1>("SS N=64  one accumulator", n);  run<64, false, 2>("SS N=64  two accumulators", n);
    run<128, false, 1>("SS N=128 one accumulator", n); run<128, false, 2>("SS N=128 two accumulators", n);
    run<256, false, 1>("SS N=256 one accumulator", n);
    return 0;
}
