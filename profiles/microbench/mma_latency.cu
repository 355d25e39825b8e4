// Probe: how long does a short chain of tcgen05.mma (M = 128, A in TMEM, B in shared memory, K = 16 per instruction,
// kind::f16 / bf16) take from issue to mbarrier completion, as a function of N and of the number of MMAs in the chain?
// One CTA, thread 0 issues `count` MMAs accumulating into the same D, then one commit; every thread waits.
// Operand contents are irrelevant (zeros).  Prints cycles for N in {16, 32, 64, 128} x count in {1, 2, 6, 12, 24}, and
// the same with the chain split over two accumulators (independent MMAs).
//   nvcc -gencode arch=compute_100a,code=sm_100a -I../../pesto_b200/csrc -o mma_latency mma_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#include "tc_common.cuh"
using namespace pesto;

template <int N, int COUNT, int ND>
__device__ __forceinline__ long long run_cfg(uint32_t tbase, unsigned char *smem, uint64_t *bar, uint32_t &phase, int tid) {
    constexpr uint32_t idesc = tc::idesc_bf16(128, N);
    long long best = 1ll << 60;
    for (int rep = 0; rep < 5; ++rep) {
        __syncthreads();
        const long long t0 = clock64();
        if (tid < 32 && tc::elect_one()) {
            tc::fence_after_sync();
            const uint64_t d = tc::smem_desc(tc::smem_u32(smem), (uint32_t)N * 16u, 128u);
#pragma unroll
            for (int i = 0; i < COUNT; ++i)
                tc::umma_ts(tbase + 128 + (uint32_t)(i % ND) * 128u, tbase + 8u * (i & 7), d, idesc, i >= ND);
            tc::umma_commit(bar);
        }
        tc::mbar_wait(bar, phase);
        phase ^= 1u;
        const long long t1 = clock64();
        tc::fence_after_sync();
        if (t1 - t0 < best) best = t1 - t0;
    }
    return best;
}

template <int N, int ND>
__device__ void run_n(long long *out, uint32_t tbase, unsigned char *smem, uint64_t *bar, uint32_t &phase, int tid) {
    long long v[6];
    v[0] = run_cfg<N, 1, ND>(tbase, smem, bar, phase, tid);
    v[1] = run_cfg<N, 2, ND>(tbase, smem, bar, phase, tid);
    v[2] = run_cfg<N, 6, ND>(tbase, smem, bar, phase, tid);
    v[3] = run_cfg<N, 12, ND>(tbase, smem, bar, phase, tid);
    v[4] = run_cfg<N, 24, ND>(tbase, smem, bar, phase, tid);
    v[5] = run_cfg<N, 48, ND>(tbase, smem, bar, phase, tid);
    if (tid == 0)
        for (int i = 0; i < 6; ++i) out[i] = v[i];
}

__global__ void __launch_bounds__(128) probe(long long *out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint32_t slot;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tc::tmem_alloc(&slot, 512);
    if (tid == 0) {
        tc::mbar_init(&bar, 1);
        tc::fence_mbar_init();
    }
    for (int i = tid; i < 128 * 128 * 2 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0u;
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = slot;
    {   // zero the A operand columns
        uint32_t z[32];
        for (int i = 0; i < 32; ++i) z[i] = 0u;
        for (int c = 0; c < 128; c += 32) tc::tmem_st32(tbase + ((uint32_t)(warp * 32) << 16) + c, z);
        tc::wait_st();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    uint32_t phase = 0;
    run_n<16, 1>(out + 0, tbase, smem, &bar, phase, tid);
    run_n<32, 1>(out + 6, tbase, smem, &bar, phase, tid);
    run_n<64, 1>(out + 12, tbase, smem, &bar, phase, tid);
    run_n<128, 1>(out + 18, tbase, smem, &bar, phase, tid);
    run_n<16, 2>(out + 24, tbase, smem, &bar, phase, tid);
    run_n<32, 2>(out + 30, tbase, smem, &bar, phase, tid);
    run_n<64, 2>(out + 36, tbase, smem, &bar, phase, tid);
    run_n<128, 2>(out + 42, tbase, smem, &bar, phase, tid);
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tbase, 512);
}

int main() {
    const int Ns[4] = {16, 32, 64, 128}, Cs[6] = {1, 2, 6, 12, 24, 48};
    long long *dout, hout[48];
    cudaMalloc(&dout, sizeof hout);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    probe<<<1, 128, 65536>>>(dout);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(hout, dout, sizeof hout, cudaMemcpyDeviceToHost);
    printf("tcgen05.mma M=128 K=16 bf16, A in TMEM, fully unrolled issue by one elected lane: cycles from issue of a chain of\n`count` MMAs + commit to mbarrier completion (min of 5)\n");
    for (int nd = 1; nd <= 2; ++nd) {
        printf("accumulators: %d\n  N \\ count", nd);
        for (int b = 0; b < 6; ++b) printf("%8d", Cs[b]);
        printf("\n");
        for (int a = 0; a < 4; ++a) {
            printf("  N = %3d  ", Ns[a]);
            for (int b = 0; b < 6; ++b) printf("%8lld", hout[(nd - 1) * 24 + a * 6 + b]);
            printf("\n");
        }
    }
    return 0;
}
