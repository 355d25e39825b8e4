// Thin inline-PTX wrappers for the sm_100a tensor-core path: TMEM allocation, tcgen05.mma / ld / st / commit,
// mbarrier, UMMA shared-memory and instruction descriptors, bf16 hi/lo splitting.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace pesto {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- TMEM ----------------------------------------------------------------------------------------------------
// one full warp; writes the base address (lane 0, first column) to *dst_smem
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (UMMA operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// thread t of warp w reads / writes TMEM lane 32*(w%4)+t, 32 (16) consecutive 32-bit columns starting at taddr
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}

// one lane of the (converged) warp: the canonical single-thread issue pattern for tcgen05.mma / commit
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}"
        : "=r"(pred));
    return pred != 0;
}

// 16x256b.x2: the warp writes 16 TMEM lanes x 16 columns; thread t supplies, with m = t % 4 and row = t / 4:
//   r0,r1 -> lane base+row,   columns 2m, 2m+1      r2,r3 -> lane base+8+row, columns 2m, 2m+1
//   r4,r5 -> lane base+row,   columns 8+2m, 8+2m+1  r6,r7 -> lane base+8+row, columns 8+2m, 8+2m+1
// (measured with profiles/microbench/tmem_shape.cu); four neighbouring threads share a row, which lets them read
// one contiguous 128-byte piece of a gathered row
__device__ __forceinline__ void tmem_st_16x256b_x2(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}

__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_st_16x256b_x1(uint32_t taddr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
    asm volatile("tcgen05.st.sync.aligned.16x256b.x1.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}

// ---- mbarrier ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// try_wait suspends the warp in hardware until the phase completes or the time hint (ns) expires
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)
        : "memory");
    return ok != 0;
}
// Bounded wait: a tensor-core stage that never completes must not hang the GPU.  On timeout the stage id is
// recorded in *watchdog (global memory) and the caller carries on with garbage; hosts check the flag.
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity, int *watchdog = nullptr, int stage = 0,
                                          uint32_t max_spin = 1u << 20) {
#pragma unroll 1
    for (uint32_t spin = 0; spin < max_spin; ++spin)
        if (mbar_try_wait(bar, parity)) return true;
    if (watchdog) atomicMax(watchdog, stage);
    return false;
}

// plain arrival (release, CTA scope) and a non-blocking completion test
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

// The same bounded wait on a barrier given by its 32-bit shared-memory address (computed once per kernel: the generic ->
// shared conversion costs a handful of instructions per call, and a tile waits a dozen times)
__device__ __forceinline__ bool mbar_wait_a(uint32_t bar_addr, uint32_t parity, int *watchdog, int stage, uint32_t max_spin = 1u << 20) {
#pragma unroll 1
    for (uint32_t spin = 0; spin < max_spin; ++spin) {
        uint32_t ok;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(ok)
            : "r"(bar_addr), "r"(parity), "r"(20000u)
            : "memory");
        if (ok) return true;
    }
    if (watchdog) atomicMax(watchdog, stage);
    return false;
}

// ---- bulk async copy global -> shared (TMA engine, no tensor map): completes `bytes` on the mbarrier ---------------
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// dst (shared), src (global) and bytes must be multiples of 16
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// A contiguous block global -> shared through the TMA engine (SASS UBLKCP), issued by ONE thread: arms `bar` with the byte
// count and splits the block into pieces of at most 32 KB; the block's readers wait for the barrier's phase.
__device__ __forceinline__ void bulk_g2s_block(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    mbar_arrive_expect_tx(bar, bytes);
    for (uint32_t o = 0; o < bytes; o += 32768u)
        bulk_g2s(reinterpret_cast<unsigned char *>(dst_smem) + o, reinterpret_cast<const unsigned char *>(src_gmem) + o,
                 bytes - o < 32768u ? bytes - o : 32768u, bar);
}

// ---- programmatic dependent launch (the launch carries cudaLaunchAttributeProgrammaticStreamSerialization) ---------------
// pdl_wait: everything the previous kernel in the stream wrote is visible afterwards (no-op for a normal launch);
// pdl_launch_dependents: the next kernel's CTAs may start (on SMs with room) and run their prologue up to their own pdl_wait
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- cp.async (LDGSTS): 16 bytes global -> shared per lane, L2 only -------------------------------------------------
__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src_gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
// the mbarrier receives one arrival when all cp.async operations issued so far by this thread have completed
// (.noinc: the arrival counts against the barrier's expected count)
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---- UMMA ----------------------------------------------------------------------------------------------------
// Shared-memory operand descriptor, K-major, no swizzle: element (row n, k) of a bf16 operand lives at
//     base + (k/8)*lbo + (n/8)*sbo + (n%8)*16 + (k%8)*2   bytes
// (8 rows x 16 bytes "core matrices" stored contiguously).  Bits: start>>4 [0,14), lbo>>4 [16,30),
// sbo>>4 [32,46), version=1 [46,48), layout_type=0 (SWIZZLE_NONE) [61,64).
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// The same descriptor from the operand region's base address in 16-byte units (base14 = smem address >> 4, a uniform
// value computed once per kernel) plus a compile-time byte offset: one add on the low word, constant high word.
// (Shared memory is < 256 KB, so base14 + offset never carries into the lbo field.)
__device__ __forceinline__ uint64_t smem_desc14(uint32_t base14, uint32_t off_bytes, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    const uint32_t lo = base14 + ((off_bytes >> 4) + (((lbo_bytes >> 4) & 0x3FFFu) << 16));
    const uint32_t hi = ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14);
    return ((uint64_t)hi << 32) | (uint64_t)lo;
}
// 16-bit format of the split operand planes (hi | lo).  Default fp16: 11 + 11 mantissa bits, so the dropped lo x lo term of
// the 3-term product is ~2^-24 and the product is fp32-accurate (~2^-21); fp16's range (65504, conversions saturate) is
// ample for this model's activations (|state| <= ~50) and weights.  -DPESTO_SPLIT_BF16 selects bf16
// planes (8 + 8 bits, error ~2^-16 per product, unlimited range).
#ifdef PESTO_SPLIT_BF16
constexpr bool H16_IS_FP16 = false;
#else
constexpr bool H16_IS_FP16 = true;
#endif
constexpr uint32_t H16_ONE = H16_IS_FP16 ? 0x3C00u : 0x3F80u;       // 1.0
// Instruction descriptor for kind::f16 (A/B both K-major, both in the configured 16-bit format) and fp32 accumulation:
// c_format=F32 [4,6), a_format [7,10), b_format [10,13) (0 = F16, 1 = BF16), n>>3 [17,23), m>>4 [24,29)
__host__ __device__ constexpr uint32_t idesc_h16(int m, int n) {
    return (1u << 4) | ((H16_IS_FP16 ? 0u : 1u) << 7) | ((H16_IS_FP16 ? 0u : 1u) << 10) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(m >> 4) << 24);
}
// host-side conversions for the weight images
__host__ inline uint16_t h16_from_f32_host(float f) {
    if (H16_IS_FP16) {
        const __half_raw r = __float2half_rn(f);
        return r.x;
    }
    const __nv_bfloat16_raw r = __float2bfloat16_rn(f);
    return r.x;
}
__host__ inline float h16_to_f32_host(uint16_t h) {
    if (H16_IS_FP16) {
        __half_raw r;
        r.x = h;
        return __half2float(__half(r));
    }
    __nv_bfloat16_raw r;
    r.x = h;
    return __bfloat162float(__nv_bfloat16(r));
}
// D[tmem] (+)= A[tmem] * B[smem]^T ; one thread issues for the CTA
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- 256-bit read-only global load (sm_100: LDG.E.256): one full 32-byte sector per lane ------------------------
__device__ __forceinline__ void ldg256(const float *p, float (&v)[8]) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
}

// 64-bit read-only global load with a fixed place in the instruction stream (volatile: the compiler keeps the issue point)
__device__ __forceinline__ float2 ldg64(const float *p) {
    float2 v;
    asm volatile("ld.global.nc.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}

__device__ __forceinline__ unsigned long long ldg64u(const void *p) {
    unsigned long long v;
    asm volatile("ld.global.nc.b64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

// ---- 16-bit splitting ------------------------------------------------------------------------------------------
// pack two floats into the configured 16-bit format (round to nearest even, saturating): low half = a, high half = b
__device__ __forceinline__ uint32_t pack_h16x2(float a, float b) {
    uint32_t r;
    if (H16_IS_FP16) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
// the two floats of a packed word
__device__ __forceinline__ void unpack_h16x2(uint32_t w, float &a, float &b) {
    if (H16_IS_FP16) {
        asm("{\n\t.reg .f16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}" : "=f"(a), "=f"(b) : "r"(w));
    } else {
        a = __uint_as_float(w << 16);
        b = __uint_as_float(w & 0xffff0000u);
    }
}
// remainders a - float(w.lo), b - float(w.hi) of a packed fp16 word: one mixed-precision FMA each (FHFMA: f16 x f16 + f32,
// the product w.x * -1 is exact, one rounding of an exactly representable difference) instead of two conversions and a subtract
__device__ __forceinline__ void residual_f16x2(uint32_t w, float a, float b, float &la, float &lb) {
    asm("{\n\t.reg .b16 l, h, m;\n\tmov.b32 {l, h}, %2;\n\tmov.b16 m, 0xBC00;\n\t"
        "fma.rn.f32.f16 %0, l, m, %3;\n\tfma.rn.f32.f16 %1, h, m, %4;\n\t}"
        : "=f"(la), "=f"(lb) : "r"(w), "f"(a), "f"(b));
}
// one float -> 16-bit pattern and back (U planes)
__device__ __forceinline__ uint16_t h16_from_f32(float f) { return (uint16_t)(pack_h16x2(f, 0.f) & 0xffffu); }
__device__ __forceinline__ float h16_to_f32(uint16_t h) {
    float a, b;
    unpack_h16x2((uint32_t)h, a, b);
    return a;
}
// hi = h16x2(a, b); lo = h16x2(a - float(hi.a), b - float(hi.b))
__device__ __forceinline__ void split_h16x2(float a, float b, uint32_t &hi, uint32_t &lo) {
    hi = pack_h16x2(a, b);
    if (H16_IS_FP16) {
        float ra, rb;
        residual_f16x2(hi, a, b, ra, rb);
        lo = pack_h16x2(ra, rb);
        return;
    }
    float ha, hb;
    unpack_h16x2(hi, ha, hb);
    unsigned long long x, h, l;                       // one packed subtract (FADD2) for the two remainders
    asm("mov.b64 %0, {%1,%2};" : "=l"(x) : "f"(a), "f"(b));
    asm("mov.b64 %0, {%1,%2};" : "=l"(h) : "f"(-ha), "f"(-hb));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(l) : "l"(x), "l"(h));
    float la, lb;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(la), "=f"(lb) : "l"(l));
    lo = pack_h16x2(la, lb);
}

}  // namespace tc
}  // namespace pesto
