// Probe: which (TMEM lane, column) does register i of thread t address in tcgen05.st/ld .16x256b.x2 ?
// Writes tag = (t << 8) | i with st.16x256b.x2 at lane bases 0 and 16 of each warp's quarter, reads back with
// ld.32x32b.x16 and prints the decoded map for warps 0 and 3.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tmem_shape tmem_shape.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__global__ void __launch_bounds__(128) probe(uint32_t *out) {
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(32u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot;
    for (int half = 0; half < 2; ++half) {
        uint32_t r[8];
        for (int i = 0; i < 8; ++i) r[i] = ((uint32_t)lane << 8) | (uint32_t)(half * 8 + i) | 0x10000u;
        const uint32_t taddr = base + ((uint32_t)(warp * 32 + half * 16) << 16);
        asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                     "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[16];
    const uint32_t taddr = base + ((uint32_t)(warp * 32) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int c = 0; c < 16; ++c) out[tid * 16 + c] = v[c];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(32u) : "memory");
}
int main() {
    uint32_t *d, h[128 * 16];
    cudaMalloc(&d, sizeof(h));
    cudaMemset(d, 0, sizeof(h));
    probe<<<1, 128>>>(d);
    cudaDeviceSynchronize();
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%s\nrow(lane) : col -> (thread, half*8+reg)\n", cudaGetErrorString(cudaGetLastError()));
    for (int warp = 0; warp < 4; warp += 3)
        for (int l = 0; l < 32; ++l) {
            printf("w%d lane %2d:", warp, l);
            for (int c = 0; c < 16; ++c) {
                uint32_t x = h[(warp * 32 + l) * 16 + c];
                if (x & 0x10000u) printf(" c%-2d=(t%-2u,r%-2u)", c, (x >> 8) & 0xff, x & 0xff); else printf(" c%-2d=(-------)", c);
            }
            printf("\n");
        }
    return 0;
}
