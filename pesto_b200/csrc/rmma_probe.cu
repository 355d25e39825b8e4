// Building-block probe for the edge kernel's tensor-core reduction (tests/test_gpu_umma.py): everything the fused
// kernel relies on that the first-round kernels did not use yet, in isolation against a CPU matmul:
//   * a TMA row gather (cp.async.bulk.tensor.2d ... tile::gather4, 64-byte swizzle) of neighbour records
//     p16[row] = fp16 hi plane (96) | fp16 lo plane (96) into shared memory, 128 rows = one tile of edge slots;
//   * tcgen05.mma with BOTH operands MN-major in shared memory: A = the gathered records as they landed
//     (M = channel, K = edge slot; canonical SWIZZLE_64B layout), B = per-edge weights written by threads
//     (N = weight kind, K = edge slot; canonical un-swizzled "interleaved" layout), 3-term split product;
//   * thread-side reads of the swizzled records (the p_j . r features are computed from the same staged rows).
#include <cuda.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace pesto {
namespace {

constexpr int PR_EDGES = 128;                 // edge slots of a tile
constexpr int PR_GROUP = 8192;                // bytes of one 32-channel group: [128 edges][64 B]
constexpr int PR_P16 = 6 * PR_GROUP;          // hi c0 c1 c2 | lo c0 c1 c2
constexpr int PR_PAD = 4 * PR_GROUP;          // M = 128 reads four groups from the descriptor's start: keep them inside the allocation
constexpr int PR_B = PR_P16 + PR_PAD;         // B planes: [N/8 = 2][K/8 = 16][8 k][8 n] fp16 = 4096 B each
constexpr int PR_SMEM = PR_B + 2 * 4096 + 1024;

__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap *map, int col, int r0, int r1, int r2, int r3, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst),
        "l"(map), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar)
        : "memory");
}

// descriptor with an explicit layout type (bits [61,64)): 0 none, 2 SWIZZLE_128B, 4 SWIZZLE_64B, 6 SWIZZLE_32B
__device__ __forceinline__ uint64_t smem_desc_sw(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    return tc::smem_desc(saddr, lbo_bytes, sbo_bytes) | ((uint64_t)layout << 61);
}

__global__ void __launch_bounds__(128)
rmma_probe_kernel(const __grid_constant__ CUtensorMap map, const int32_t *__restrict__ ids, const float *__restrict__ W,
                  float *__restrict__ D, float *__restrict__ Prec, uint32_t *__restrict__ raw, int a_lbo, int a_sbo, int b_lbo,
                  int b_sbo, uint32_t idesc, int issue_lanes, int *status) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t bars[2];
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 32);
    if (tid == 0) {
        tc::mbar_init(&bars[0], 1);
        tc::mbar_init(&bars[1], 1);
        tc::fence_mbar_init();
    }
    for (int u = tid; u < PR_SMEM / 4; u += 128) reinterpret_cast<uint32_t *>(smem)[u] = 0u;
    __syncthreads();
    {   // B: weights of edge slot k = tid, all 16 kinds, fp16 hi | lo planes, MN-major: (n/8) sbo + (k/8) lbo + (k%8) 16 + (n%8) 2
        const int k = tid;
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                tc::split_h16x2(W[(nb * 8 + 2 * u) * PR_EDGES + k], W[(nb * 8 + 2 * u + 1) * PR_EDGES + k], hi[u], lo[u]);
            const int off = nb * b_sbo + (k >> 3) * b_lbo + (k & 7) * 16;
            *reinterpret_cast<uint4 *>(smem + PR_B + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4 *>(smem + PR_B + 4096 + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
    }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = tmem_base_s;
    // timing of the gather (clock64 stamps -> status[1..2]): the 192 gather4 instructions of the tile are dealt round-robin
    // to issue_lanes issuing threads: 1 = thread 0; 4 = lane 0 of every warp; 32 = the lanes of warp 0; 128 = every thread
    __shared__ __align__(16) int ids_s[PR_EDGES];
    ids_s[tid] = ids[tid];
    long long t0 = 0, t1 = 0;
    if (tid == 0) tc::mbar_arrive_expect_tx(&bars[0], PR_EDGES * 384);
    __syncthreads();
    const int lane_ = tid & 31;
    const int issuer = issue_lanes == 1 ? (tid == 0 ? 0 : -1) : issue_lanes == 4 ? (lane_ == 0 ? warp : -1)
                       : issue_lanes == 32 ? (warp == 0 ? lane_ : -1) : tid;
    if (issuer >= 0) {
        t0 = clock64();
        const uint32_t bar_a = tc::smem_u32(&bars[0]), dst0 = tc::smem_u32(smem);
#pragma unroll 2
        for (int q = issuer; q < 6 * PR_EDGES / 4; q += issue_lanes) {
            const int e4 = q / 6, g = q % 6;
            const int4 r = *reinterpret_cast<const int4 *>(ids_s + 4 * e4);
            tma_gather4(dst0 + g * PR_GROUP + e4 * 256, &map, (g / 3) * 96 + (g % 3) * 32, r.x, r.y, r.z, r.w, bar_a);
        }
        t1 = clock64();
    }
    bool ok = tc::mbar_wait(&bars[0], 0, status, 1, 1u << 14);
    if (tid == 0) {
        const long long t2 = clock64();
        status[1] = (int)(t1 - t0);
        status[2] = (int)(t2 - t0);
    }
    if (tid == 0 && ok) {
        tc::fence_after_sync();
        const uint32_t p = tc::smem_u32(smem), b = tc::smem_u32(smem + PR_B);
        for (int ks = 0; ks < PR_EDGES / 16; ++ks) {
            const uint64_t ah = smem_desc_sw(p + ks * 1024, a_lbo, a_sbo, 4u);
            const uint64_t al = smem_desc_sw(p + 3 * PR_GROUP + ks * 1024, a_lbo, a_sbo, 4u);
            const uint64_t bh = smem_desc_sw(b + ks * 2 * b_lbo, b_lbo, b_sbo, 0u);
            const uint64_t bl = smem_desc_sw(b + 4096 + ks * 2 * b_lbo, b_lbo, b_sbo, 0u);
            tc::umma_ss(tbase, ah, bh, idesc, ks > 0);
            tc::umma_ss(tbase, al, bh, idesc, 1u);
            tc::umma_ss(tbase, ah, bl, idesc, 1u);
        }
        tc::umma_commit(&bars[1]);
    }
    ok = ok && tc::mbar_wait(&bars[1], 0, status, 2, 1u << 14);
    tc::fence_after_sync();
    {
        uint32_t r[16];
        tc::tmem_ld16(tbase + ((uint32_t)(warp * 32) << 16), r);
        tc::wait_ld();
#pragma unroll
        for (int u = 0; u < 16; ++u) D[tid * 16 + u] = __uint_as_float(r[u]);
    }
    {   // thread-side read of edge slot e = tid: chunk c4 of row e lives at chunk c4 ^ ((e >> 1) & 3) (64-byte swizzle)
        const int e = tid;
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                const int off = e * 64 + ((c4 ^ ((e >> 1) & 3)) << 4);
                const uint4 h = *reinterpret_cast<const uint4 *>(smem + c * PR_GROUP + off);
                const uint4 l = *reinterpret_cast<const uint4 *>(smem + (3 + c) * PR_GROUP + off);
                const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float h0, h1, l0, l1;
                    tc::unpack_h16x2(hw[u], h0, h1);
                    tc::unpack_h16x2(lw[u], l0, l1);
                    Prec[e * 96 + c * 32 + c4 * 8 + 2 * u] = h0 + l0;
                    Prec[e * 96 + c * 32 + c4 * 8 + 2 * u + 1] = h1 + l1;
                }
            }
    }
    if (raw)
        for (int u = tid; u < PR_P16 / 4; u += 128) raw[u] = reinterpret_cast<const uint32_t *>(smem)[u];
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tbase, 32);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

// Tensor map over a row-major fp16 matrix [n_rows][row_elems] (row pitch in bytes), box = {box_elems, 1}: the shape
// tile::gather4 expects (four such rows per instruction).  swizzle: 0 none, 1 32B, 2 64B, 3 128B.
int make_row_gather_map(void *map_out, const void *base, int n_rows, int row_elems, int pitch_bytes, int box_elems, int swizzle) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        cudaDriverEntryPointQueryResult qres;
        void *p = nullptr;
        PESTO_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
        if (!p || qres != cudaDriverEntryPointSuccess) {
            set_error("cuTensorMapEncodeTiled is not available from this driver");
            return PESTO_ECUDA;
        }
        fn = (EncodeTiledFn)p;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)row_elems, (cuuint64_t)n_rows};
    const cuuint64_t strides[1] = {(cuuint64_t)pitch_bytes};
    const cuuint32_t box[2] = {(cuuint32_t)box_elems, 1u};
    const cuuint32_t estr[2] = {1u, 1u};
    const CUtensorMapSwizzle sw = swizzle == 3 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_64B
                                  : swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
    const CUresult r = fn((CUtensorMap *)map_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for [%d][%d] pitch %d box %d", (int)r, n_rows, row_elems, pitch_bytes, box_elems);
        return PESTO_ECUDA;
    }
    return PESTO_OK;
}

}  // namespace pesto

extern "C" int pesto_debug_rmma_probe(const void *p16, int n_rows, const int32_t *ids, const float *W, float *D, float *Prec,
                                      void *raw, int a_lbo, int a_sbo, int b_lbo, int b_sbo, int idesc, int issue_lanes, int *status, void *stream) {
    using namespace pesto;
    CUtensorMap map;
    int rc = make_row_gather_map(&map, p16, n_rows, 192, 384, 32, 2);
    if (rc != PESTO_OK) return rc;
    const uint32_t id = idesc ? (uint32_t)idesc : (tc::idesc_h16(128, 16) | (1u << 15) | (1u << 16));
    PESTO_CUDA(cudaFuncSetAttribute(rmma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PR_SMEM));
    rmma_probe_kernel<<<1, 128, PR_SMEM, (cudaStream_t)stream>>>(map, ids, W, D, Prec, (uint32_t *)raw, a_lbo >= 0 ? a_lbo : PR_GROUP,
                                                               a_sbo >= 0 ? a_sbo : 512, b_lbo >= 0 ? b_lbo : 128,
                                                               b_sbo >= 0 ? b_sbo : 2048, id, issue_lanes, status);
    PESTO_CUDA(cudaGetLastError());
    return PESTO_OK;
}

// ---- timing of smem-operand MMA chains (which operand layouts the tensor core reads at full speed) -----------------------
namespace pesto {
namespace {
__global__ void __launch_bounds__(128)
mma_time_kernel(int n_mma, int N, uint32_t a_layout, uint32_t a_lbo, uint32_t a_sbo, uint32_t a_kstep, uint32_t b_layout, uint32_t b_lbo,
                uint32_t b_sbo, uint32_t b_kstep, uint32_t idesc, int a_tmem, long long *out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 256);
    if (tid == 0) {
        tc::mbar_init(&bar, 1);
        tc::fence_mbar_init();
    }
    for (int u = tid; u < 96 * 1024 / 4; u += 128) reinterpret_cast<uint32_t *>(smem)[u] = 0u;
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
    if (warp_u == 0) {
        // warp-uniform operands (kernel parameters and the broadcast TMEM base): the issue loop runs on the uniform datapath,
        // eight K steps unrolled with descriptors that differ by constants
        const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base_s, 0), a0 = tc::smem_u32(smem), b0 = tc::smem_u32(smem + 64 * 1024);
        const uint64_t ad0 = tc::smem_desc(a0, a_lbo, a_sbo) | ((uint64_t)a_layout << 61);
        const uint64_t bd0 = tc::smem_desc(b0, b_lbo, b_sbo) | ((uint64_t)b_layout << 61);
        const uint64_t astep = a_kstep >> 4, bstep = b_kstep >> 4;
        long long t0 = 0, t1 = 0;
        if (tc::elect_one()) {
            t0 = clock64();
            for (int i = 0; i < n_mma; i += 8) {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    if (a_tmem) tc::umma_ts(tb + 128, tb + 8 * ks, bd0 + ks * bstep, idesc, (i | ks) > 0);
                    else tc::umma_ss(tb + 128, ad0 + ks * astep, bd0 + ks * bstep, idesc, (i | ks) > 0);
                }
            }
            t1 = clock64();
            tc::umma_commit(&bar);
        }
        __syncwarp();
        tc::mbar_wait(&bar, 0, nullptr, 0, 1u << 16);
        const long long t2 = clock64();
        if (t0) {
            out[0] = t1 - t0;
            out[1] = t2 - t0;
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base_s, 256);
}
}  // namespace
}  // namespace pesto

/* Debug: cycles of a chain of n_mma tcgen05.mma (M = 128, K = 16, fp16) with the given operand descriptors (layout type as in
 * the descriptor's bits [61,64); a_tmem != 0: A operand from TMEM); out (device int64[2]) = issue cycles, cycles until done. */
extern "C" int pesto_debug_mma_time(int n_mma, int N, int a_major_mn, int a_layout, int a_lbo, int a_sbo, int a_kstep, int b_major_mn,
                                    int b_layout, int b_lbo, int b_sbo, int b_kstep, int a_tmem, long long *out, void *stream) {
    using namespace pesto;
    const uint32_t id = tc::idesc_h16(128, N) | ((uint32_t)(a_major_mn != 0) << 15) | ((uint32_t)(b_major_mn != 0) << 16);
    PESTO_CUDA(cudaFuncSetAttribute(mma_time_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    mma_time_kernel<<<1, 128, 96 * 1024, (cudaStream_t)stream>>>(n_mma, N, a_layout, a_lbo, a_sbo, a_kstep, b_layout, b_lbo, b_sbo,
                                                               b_kstep, id, a_tmem, out);
    PESTO_CUDA(cudaGetLastError());
    return PESTO_OK;
}
