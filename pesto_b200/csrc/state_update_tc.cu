// Tensor-core (tcgen05 / TMEM) path of the StateUpdate edge kernel (src/model_operations.py:87-154).
//
// Tile = 128 edge slots = 128/nn atoms; one CTA of 128 threads, thread t <-> edge t <-> TMEM lane t, so every
// MMA result row is read back by the thread that owns the edge.  Two CTAs per SM (256 TMEM columns each) overlap
// one CTA's tensor-core wait with the other's CUDA-core stage.  Per tile:
//
//   S0   gather p_j; A1 = [p_j.r (32) | p_i.r (32) | d, 0.. (16)] as bf16 (hi | lo) -> TMEM            (CUDA cores)
//   M1   D1[128x128] = A1 . B1^T,  B1 = [W1 cols of p_j.r ; p_i.r ; d]                                 (tcgen05.mma)
//   E1   h1 = ELU(D1 + U_i + T_j)  (U_i, T_j: per-atom factors from the node kernel) -> A2 in place    (CUDA cores)
//   M2   D2 = blockdiag(eqkm.2, epkm.2, evm.2) applied to A2's three column groups
//   E2   h2 = ELU(D2 + b2) -> A3 in place
//   M3   D3 = [eqkm.4 | epkm.4 | evm.4] applied to A3's column groups
//   E3   logits, softmax over the atom's nn / 3nn tokens (warp shuffles), attention-weighted sums of V0, V1 (x) r,
//        p_j by a recursive-halving transpose-reduce across the warp; then the per-atom qpm / ppm projections.
//
// The A operand of every MMA lives in TMEM (written by tcgen05.st, thread-per-row, bf16 packed two per column),
// B (weights) in shared memory as K-major un-swizzled UMMA images prepared on the host at model-finalize time.
// SPLIT = true computes hi*hi + lo*hi + hi*lo (3 MMAs per K step, ~2^-17 relative error: parity mode);
// SPLIT = false is a single bf16 pass (speed mode).
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "tc_common.cuh"

namespace pesto {

using L = LayerLayout;

// ------------------------------------------------------------------------------------------------------------
// host: tensor-core weight images of one layer
// ------------------------------------------------------------------------------------------------------------
namespace tcimg {
// byte offsets inside one precision image (hi or lo); every matrix is [K/8][N][8] bf16
constexpr int B1 = 0;                       // N=128, K=80
constexpr int B2Q = B1 + 128 * 80 * 2;      // N=32,  K=32
constexpr int B2P = B2Q + 32 * 32 * 2;
constexpr int B2V = B2P + 32 * 32 * 2;      // N=64,  K=64
constexpr int B3Q = B2V + 64 * 64 * 2;      // N=16,  K=32 (3 rows used)
constexpr int B3P = B3Q + 16 * 32 * 2;      // N=16,  K=32 (9 rows used)
constexpr int B3V = B3P + 16 * 32 * 2;      // N=64,  K=64
constexpr int IMG = B3V + 64 * 64 * 2;      // 43008 bytes
constexpr int BIAS = 2 * IMG;               // fp32: b2[128] | b3[96]
constexpr int TOTAL = BIAS + (128 + 96) * 4;
static_assert(IMG % 16 == 0 && TOTAL % 16 == 0, "images are copied with 16-byte vectors");
}  // namespace tcimg

size_t tc_layer_bytes() { return tcimg::TOTAL; }

static inline uint16_t f32_to_bf16_rne(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);   // NaN
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
static inline float bf16_to_f32(uint16_t h) {
    uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

void pack_tc_layer(const float *blob, void *dst_v) {
    unsigned char *dst = (unsigned char *)dst_v;
    memset(dst, 0, tcimg::TOTAL);
    auto put = [&](int img_off, int N, int n, int k, float w) {
        const size_t e = (size_t)(k / 8) * N * 8 + (size_t)n * 8 + (k % 8);
        const uint16_t hi = f32_to_bf16_rne(w);
        const uint16_t lo = f32_to_bf16_rne(w - bf16_to_f32(hi));
        ((uint16_t *)(dst + img_off))[e] = hi;
        ((uint16_t *)(dst + tcimg::IMG + img_off))[e] = lo;
    };
    for (int o = 0; o < 128; ++o) {
        for (int s = 0; s < 32; ++s) {
            put(tcimg::B1, 128, o, s, blob[L::E_WB + s * 128 + o]);          // p_j . r
            put(tcimg::B1, 128, o, 32 + s, blob[L::N_A + s * 128 + o]);      // p_i . r
        }
        put(tcimg::B1, 128, o, 64, blob[L::E_WD + o]);                       // d
    }
    for (int n = 0; n < 32; ++n)
        for (int k = 0; k < 32; ++k) {
            put(tcimg::B2Q, 32, n, k, blob[L::E_2Q + k * 32 + n]);
            put(tcimg::B2P, 32, n, k, blob[L::E_2P + k * 32 + n]);
            if (n < 3) put(tcimg::B3Q, 16, n, k, blob[L::E_3Q + k * 4 + n]);
            if (n < 9) put(tcimg::B3P, 16, n, k, blob[L::E_3P + k * 12 + n]);
        }
    for (int n = 0; n < 64; ++n)
        for (int k = 0; k < 64; ++k) {
            put(tcimg::B2V, 64, n, k, blob[L::E_2V + k * 64 + n]);
            put(tcimg::B3V, 64, n, k, blob[L::E_3V + k * 64 + n]);
        }
    float *bias = (float *)(dst + tcimg::BIAS);
    for (int i = 0; i < 32; ++i) {
        bias[i] = blob[L::E_2QB + i];
        bias[32 + i] = blob[L::E_2PB + i];
    }
    for (int i = 0; i < 64; ++i) {
        bias[64 + i] = blob[L::E_2VB + i];
        bias[128 + 32 + i] = blob[L::E_3VB + i];
    }
    for (int i = 0; i < 3; ++i) bias[128 + i] = blob[L::E_3QB + i];
    for (int i = 0; i < 9; ++i) bias[128 + 16 + i] = blob[L::E_3PB + i];
}

namespace {

__device__ int g_tc_watchdog = 0;     // != 0: a tensor-core stage timed out (stage id), see mbar_wait

constexpr int TC_THREADS = 256;     // 8 warps per tile: two column groups x four TMEM lane quarters
constexpr unsigned FULLM = 0xffffffffu;
constexpr uint32_t TM_COLS = 256;     // TMEM columns per CTA: X = [0,128), Y = [128,256)
constexpr uint32_t TX = 0, TY = 128;

// exp via one MUFU: ex2.approx.ftz (no denormal fix-up code around it; inputs below -126 flush to 0, which is exact
// enough for ELU's exp(x) - 1 and for softmax weights)
__device__ __forceinline__ float exp_fast(float x) {
    float t;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(x * 1.4426950408889634f));
    return t;
}
__device__ __forceinline__ float elu_fast(float x) { return x > 0.f ? x : (exp_fast(x) - 1.0f); }

template <int SEG>
__device__ __forceinline__ float seg_max_tc(float v) {
#pragma unroll
    for (int o = SEG / 2; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULLM, v, o));
    return v;
}
template <int SEG>
__device__ __forceinline__ float seg_sum_tc(float v) {
#pragma unroll
    for (int o = SEG / 2; o; o >>= 1) v += __shfl_xor_sync(FULLM, v, o);
    return v;
}

// Recursive-halving transpose-reduce: every lane holds v[0..32); afterwards lane l holds, in v[0 .. 32/SEG), the
// sums over its SEG-lane segment of elements (l % SEG) * (32/SEG) + t.
template <int SEG>
__device__ __forceinline__ void transpose_reduce(float (&v)[32], int lane) {
    int len = 32;
#pragma unroll
    for (int off = SEG / 2; off >= 1; off >>= 1) {
        const bool upper = (lane & off) != 0;
        const int half = len / 2;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (i < half) {
                const float send = upper ? v[i] : v[half + i];
                const float keep = upper ? v[half + i] : v[i];
                v[i] = keep + __shfl_xor_sync(FULLM, send, off);
            }
        }
        len = half;
    }
}

// ELU + bf16 (hi|lo) packing of 32 fp32 values -> 32 TMEM columns: [0,16) hi pairs, [16,32) lo pairs
template <bool SPLIT>
__device__ __forceinline__ void store_activation_chunk(uint32_t taddr, const float (&x)[32]) {
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
        if (SPLIT) {
            tc::split_bf16x2(x[2 * u], x[2 * u + 1], hi[u], lo[u]);
        } else {
            hi[u] = tc::pack_bf16x2(x[2 * u], x[2 * u + 1]);
        }
    }
    tc::tmem_st16(taddr, hi);
    if (SPLIT) tc::tmem_st16(taddr + 16, lo);
}

// issue D[d_col .. d_col+N) (+)= A(K columns packed at a_col: hi at +8s, lo at +lo_off+8s per 16-wide K step) . B^T
template <bool SPLIT>
__device__ __forceinline__ void issue_gemm(uint32_t tbase, uint32_t d_col, uint32_t a_col, uint32_t lo_off, int ksteps,
                                           uint32_t b_hi, uint32_t b_lo, int N) {
    const uint32_t idesc = tc::idesc_bf16(128, N);
    const uint32_t lbo = (uint32_t)N * 16u;
    for (int s = 0; s < ksteps; ++s) {
        // K steps inside one 32-wide activation chunk are 8 columns apart; chunks are 32 columns apart
        const uint32_t a = tbase + a_col + (uint32_t)(s >> 1) * 32u + (uint32_t)(s & 1) * 8u;
        const uint32_t koff = (uint32_t)s * 2u * lbo;
        const uint64_t dh = tc::smem_desc(b_hi + koff, lbo, 128u);
        tc::umma_ts(tbase + d_col, a, dh, idesc, s > 0);
        if (SPLIT) {
            tc::umma_ts(tbase + d_col, a + lo_off, dh, idesc, 1u);
            tc::umma_ts(tbase + d_col, a, tc::smem_desc(b_lo + koff, lbo, 128u), idesc, 1u);
        }
    }
}

// Column-split mapping: a tile (128 edges) is worked on by 8 warps.  Warps 0-3 (group 0) and warps 4-7 (group 1)
// both map thread -> edge/TMEM lane 32*(warp%4)+lane, and each group handles half of the columns of every
// stage.  This doubles the warps per SM (2 CTAs x 8 warps) at half the registers per thread, which is what hides
// the gather / TMEM / shuffle latencies.
template <int NN, bool SPLIT>
__global__ void __launch_bounds__(TC_THREADS, 2)
edge_kernel_tc(const float *__restrict__ lw, const unsigned char *__restrict__ tcw, int n_atoms,
               const int32_t *__restrict__ ids32, const float4 *__restrict__ geom, const float *__restrict__ state_in,
               const float *__restrict__ nodeT, const float *__restrict__ nodeC, float *__restrict__ Zout) {
    constexpr int TA = 128 / NN;                    // atoms per tile
    constexpr int SEG = NN < 32 ? NN : 32;          // lanes of one atom inside a warp
    constexpr int APW = 32 / SEG;                   // atoms per warp
    constexpr int EPL = 32 / SEG;                   // reduced elements per lane and 32-vector
    constexpr int WPA = NN / SEG;                   // warps (of one group) per atom: 2 for nn = 64
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char *img = smem_raw;                                      // weight images + biases
    const float *b2 = reinterpret_cast<const float *>(img + tcimg::BIAS);
    const float *b3 = b2 + 128;
    float *Zs = reinterpret_cast<float *>(img + tcimg::TOTAL);          // [4 lane quarters][APW][256]
    float *red = Zs + 4 * APW * 256;                                    // [2 groups][4][8]
    float *Ws = red + 64;                                               // [128 edges][4]: Mp[h, token p_j] (2), row j, pad
    uint64_t *bar = reinterpret_cast<uint64_t *>(Ws + 128 * 4);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar + 1);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = warp >> 2, quarter = warp & 3;       // column group, TMEM lane quarter
    if (warp == 0) tc::tmem_alloc(tmem_slot, TM_COLS);
    if (tid == 0) {
        tc::mbar_init(bar, 1);
        tc::fence_mbar_init();
    }
    for (int u = tid; u < tcimg::TOTAL / 16; u += TC_THREADS)
        reinterpret_cast<uint4 *>(img)[u] = __ldg(reinterpret_cast<const uint4 *>(tcw) + u);
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = *tmem_slot;
    const uint32_t tlane = tbase + ((uint32_t)(quarter * 32) << 16);
    const uint32_t img_hi = tc::smem_u32(img), img_lo = img_hi + tcimg::IMG;
    uint32_t phase = 0;
    bool alive = true;      // false after a tensor-core stage timed out: finish with garbage, but finish
    float *redg = red + grp * 32;

    const int n_tiles = (n_atoms + TA - 1) / TA;
    const int e = tid & 127;                                       // edge slot inside the tile = TMEM lane
    const int a_loc = e / NN, k = e % NN;
    int j_next = 0;
    float4 g_next = make_float4(0.f, 0.f, 0.f, 0.f);
    if ((int)blockIdx.x < n_tiles) {
        const int i0 = min((int)blockIdx.x * TA + a_loc, n_atoms - 1);
        j_next = ids32[(size_t)i0 * KMAX + k];
        g_next = geom[(size_t)i0 * KMAX + k];
    }
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int i = min(tile * TA + a_loc, n_atoms - 1);         // tail tile: clamp (results are not written)
        const int j = j_next;
        const float4 g = g_next;
        const float *sI = state_in + (size_t)(i + 1) * SR;
        const float *sJ = state_in + (size_t)j * SR;
        const float *cI = nodeC + (size_t)(i + 1) * NODE_C_STRIDE;
        const float *tJ = nodeT + (size_t)j * NODE_T_STRIDE;

        // ---------------------------------------------------------------- S0: A1 = [p_j.r | p_i.r | d] -> TMEM (Y)
        // hi: Y + [0,16) p_j.r, [16,32) p_i.r, [32,40) d block;  lo: Y + 40 + same.  group 0: p_j.r, group 1: p_i.r, d
        {
            const float *src = grp == 0 ? sJ : sI;
            float pr[32];
#pragma unroll
            for (int s = 0; s < S; s += 8) {
                float x[8], y[8], z[8];
                tc::ldg256(src + 32 + s, x);
                tc::ldg256(src + 64 + s, y);
                tc::ldg256(src + 96 + s, z);
#pragma unroll
                for (int u = 0; u < 8; ++u) pr[s + u] = fmaf(g.z, z[u], fmaf(g.y, y[u], g.x * x[u]));
            }
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                if (SPLIT) tc::split_bf16x2(pr[2 * u], pr[2 * u + 1], hi[u], lo[u]);
                else hi[u] = tc::pack_bf16x2(pr[2 * u], pr[2 * u + 1]);
            }
            tc::tmem_st16(tlane + TY + 16 * grp, hi);
            if (SPLIT) tc::tmem_st16(tlane + TY + 40 + 16 * grp, lo);
            if (grp == 1) {
                uint32_t hd[8], ld[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) hd[u] = ld[u] = 0u;
                if (SPLIT) tc::split_bf16x2(g.w, 0.f, hd[0], ld[0]); else hd[0] = tc::pack_bf16x2(g.w, 0.f);
                tc::tmem_st8(tlane + TY + 32, hd);
                if (SPLIT) tc::tmem_st8(tlane + TY + 72, ld);
            }
        }
        tc::wait_st();
        tc::fence_before_sync();
        __syncthreads();
        if (tid == 0) {                                                  // M1: D1 (X) = A1 . B1^T, K = 80
            tc::fence_after_sync();
            const uint32_t idesc = tc::idesc_bf16(128, 128);
            const uint32_t lbo = 128u * 16u;
#pragma unroll 1
            for (int s = 0; s < 5; ++s) {
                const uint32_t koff = (uint32_t)s * 2u * lbo;
                const uint64_t dh = tc::smem_desc(img_hi + tcimg::B1 + koff, lbo, 128u);
                tc::umma_ts(tbase + TX, tbase + TY + 8u * s, dh, idesc, s > 0);
                if (SPLIT) {
                    tc::umma_ts(tbase + TX, tbase + TY + 40u + 8u * s, dh, idesc, 1u);
                    tc::umma_ts(tbase + TX, tbase + TY + 8u * s, tc::smem_desc(img_lo + tcimg::B1 + koff, lbo, 128u), idesc, 1u);
                }
            }
            tc::umma_commit(bar);
        }
        // prefetch the per-atom factors of this thread's first E1 chunk while the tensor core works
        float pu[4][8], pt[4][8];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            tc::ldg256(cI + 64 * grp + 8 * u, pu[u]);
            tc::ldg256(tJ + 64 * grp + 8 * u, pt[u]);
        }
        if (alive) alive = tc::mbar_wait(bar, phase, &g_tc_watchdog, 1);
        phase ^= 1u;
        tc::fence_after_sync();

        // ---------------------------------------------------------------- E1: h1 = ELU(D1 + U_i + T_j) -> A2 (X, in place)
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
            const int c = 2 * grp + cc;
            uint32_t r[32];
            tc::tmem_ld32(tlane + TX + 32 * c, r);
            float x[32];
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int w = 0; w < 8; ++w) x[8 * u + w] = pu[u][w] + pt[u][w];
            if (cc == 0) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    tc::ldg256(cI + 32 * (c + 1) + 8 * u, pu[u]);
                    tc::ldg256(tJ + 32 * (c + 1) + 8 * u, pt[u]);
                }
            }
            tc::wait_ld();
#pragma unroll
            for (int u = 0; u < 32; ++u) x[u] = elu_fast(x[u] + __uint_as_float(r[u]));
            store_activation_chunk<SPLIT>(tlane + TX + 32 * c, x);
        }
        tc::wait_st();
        tc::fence_before_sync();
        __syncthreads();
        if (tid == 0) {                                                  // M2: D2 (Y) = blockdiag(eqkm.2, epkm.2, evm.2)
            tc::fence_after_sync();
            issue_gemm<SPLIT>(tbase, TY + 0, TX + 0, 16, 2, img_hi + tcimg::B2Q, img_lo + tcimg::B2Q, 32);
            issue_gemm<SPLIT>(tbase, TY + 32, TX + 32, 16, 2, img_hi + tcimg::B2P, img_lo + tcimg::B2P, 32);
            issue_gemm<SPLIT>(tbase, TY + 64, TX + 64, 16, 4, img_hi + tcimg::B2V, img_lo + tcimg::B2V, 64);
            tc::umma_commit(bar);
        }
        if (alive) alive = tc::mbar_wait(bar, phase, &g_tc_watchdog, 2);
        phase ^= 1u;
        tc::fence_after_sync();

        // ---------------------------------------------------------------- E2: h2 = ELU(D2 + b2) -> A3 (Y, in place)
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
            const int c = 2 * grp + cc;
            uint32_t r[32];
            tc::tmem_ld32(tlane + TY + 32 * c, r);
            tc::wait_ld();
            float x[32];
#pragma unroll
            for (int u = 0; u < 32; u += 4) {
                const float4 bb = *reinterpret_cast<const float4 *>(b2 + 32 * c + u);
                x[u + 0] = elu_fast(__uint_as_float(r[u + 0]) + bb.x);
                x[u + 1] = elu_fast(__uint_as_float(r[u + 1]) + bb.y);
                x[u + 2] = elu_fast(__uint_as_float(r[u + 2]) + bb.z);
                x[u + 3] = elu_fast(__uint_as_float(r[u + 3]) + bb.w);
            }
            store_activation_chunk<SPLIT>(tlane + TY + 32 * c, x);
        }
        tc::wait_st();
        tc::fence_before_sync();
        __syncthreads();
        if (tid == 0) {                                                  // M3: D3 (X) = [eqkm.4 | epkm.4 | evm.4]
            tc::fence_after_sync();
            issue_gemm<SPLIT>(tbase, TX + 0, TY + 0, 16, 2, img_hi + tcimg::B3Q, img_lo + tcimg::B3Q, 16);
            issue_gemm<SPLIT>(tbase, TX + 16, TY + 32, 16, 2, img_hi + tcimg::B3P, img_lo + tcimg::B3P, 16);
            issue_gemm<SPLIT>(tbase, TX + 32, TY + 64, 16, 4, img_hi + tcimg::B3V, img_lo + tcimg::B3V, 64);
            tc::umma_commit(bar);
        }
        // queries of the centre atom while the tensor core works: [t][h][k] = t*6 + h*3 + k, pre-divided by sdk
        float qv[12];
        {
            const float *Qi = cI + NODE_C_Q;
#pragma unroll
            for (int u = 0; u < 12; u += 4) {
                const float4 q4 = __ldg(reinterpret_cast<const float4 *>(Qi + u));
                qv[u] = q4.x; qv[u + 1] = q4.y; qv[u + 2] = q4.z; qv[u + 3] = q4.w;
            }
        }
        if (alive) alive = tc::mbar_wait(bar, phase, &g_tc_watchdog, 3);
        phase ^= 1u;
        tc::fence_after_sync();

        // ---------------------------------------------------------------- E3: attention (both groups compute the weights)
        float wq[NH], wp0[NH], wp1s[NH];
        {
            uint32_t r[32];
            tc::tmem_ld32(tlane + TX, r);                               // [0,3) Kq, [16,25) Kp
            tc::wait_ld();
            float kq[3], kp[9];
#pragma unroll
            for (int u = 0; u < 3; ++u) kq[u] = __uint_as_float(r[u]) + b3[u];
#pragma unroll
            for (int u = 0; u < 9; ++u) kp[u] = __uint_as_float(r[16 + u]) + b3[16 + u];
            float lq[NH], lp[NH][3];
#pragma unroll
            for (int h = 0; h < NH; ++h) {
                lq[h] = fmaf(qv[h * 3 + 2], kq[2], fmaf(qv[h * 3 + 1], kq[1], qv[h * 3] * kq[0]));
#pragma unroll
                for (int gk = 0; gk < 3; ++gk)
                    lp[h][gk] = fmaf(qv[6 + h * 3 + 2], kp[gk * 3 + 2],
                                     fmaf(qv[6 + h * 3 + 1], kp[gk * 3 + 1], qv[6 + h * 3] * kp[gk * 3]));
            }
            float mx[4];
            mx[0] = seg_max_tc<SEG>(lq[0]);
            mx[1] = seg_max_tc<SEG>(lq[1]);
            mx[2] = seg_max_tc<SEG>(fmaxf(lp[0][0], fmaxf(lp[0][1], lp[0][2])));
            mx[3] = seg_max_tc<SEG>(fmaxf(lp[1][0], fmaxf(lp[1][1], lp[1][2])));
            if (WPA == 2) {
                if (lane == 0) { redg[quarter * 8 + 0] = mx[0]; redg[quarter * 8 + 1] = mx[1]; redg[quarter * 8 + 2] = mx[2]; redg[quarter * 8 + 3] = mx[3]; }
                __syncthreads();
#pragma unroll
                for (int u = 0; u < 4; ++u) mx[u] = fmaxf(mx[u], redg[(quarter ^ 1) * 8 + u]);
            }
            float eq[NH], ep[NH][3], sm[4];
#pragma unroll
            for (int h = 0; h < NH; ++h) {
                eq[h] = exp_fast(lq[h] - mx[h]);
#pragma unroll
                for (int gk = 0; gk < 3; ++gk) ep[h][gk] = exp_fast(lp[h][gk] - mx[2 + h]);
            }
            sm[0] = seg_sum_tc<SEG>(eq[0]);
            sm[1] = seg_sum_tc<SEG>(eq[1]);
            sm[2] = seg_sum_tc<SEG>(ep[0][0] + ep[0][1] + ep[0][2]);
            sm[3] = seg_sum_tc<SEG>(ep[1][0] + ep[1][1] + ep[1][2]);
            if (WPA == 2) {
                if (lane == 0) { redg[quarter * 8 + 4] = sm[0]; redg[quarter * 8 + 5] = sm[1]; redg[quarter * 8 + 6] = sm[2]; redg[quarter * 8 + 7] = sm[3]; }
                __syncthreads();
#pragma unroll
                for (int u = 0; u < 4; ++u) sm[u] += redg[(quarter ^ 1) * 8 + 4 + u];
            }
#pragma unroll
            for (int h = 0; h < NH; ++h) {
                const float iq = 1.0f / sm[h], ip = 1.0f / sm[2 + h];
                wq[h] = eq[h] * iq;                    // Mq[h]
                wp0[h] = ep[h][0] * ip;                // Mp[h, token V1 (x) r]
                wp1s[h] = seg_sum_tc<SEG>(ep[h][1] * ip);   // sum over this warp's edges of Mp[h, token p_i]
            }
            if (grp == 0)                              // the p_j token is summed warp-per-atom below (coalesced rows)
                *reinterpret_cast<float4 *>(Ws + 4 * e) = make_float4(ep[0][2] / sm[2], ep[1][2] / sm[3], __int_as_float(j), 0.f);
        }
        // weighted sums, reduced over the atom's edges; lane keeps EPL elements per 32-vector.
        // group 0: Zq (2 vectors) + Zp[c=0] (2 vectors); group 1: Zp[c=1], Zp[c=2] (4 vectors)
        float *zw = Zs + (quarter * APW + (lane / SEG)) * 256 + (lane % SEG) * EPL;
        if (grp == 0) {
            uint32_t r[32];
            tc::tmem_ld32(tlane + TX + 32, r);                          // V0
            tc::wait_ld();
#pragma unroll
            for (int h = 0; h < NH; ++h) {                              // Zq = Mq . V0   (src/model_operations.py:143)
                float v[32];
#pragma unroll
                for (int u = 0; u < 32; ++u) v[u] = wq[h] * (__uint_as_float(r[u]) + b3[32 + u]);
                transpose_reduce<SEG>(v, lane);
#pragma unroll
                for (int t = 0; t < EPL; ++t) zw[h * 32 + t] = v[t];
            }
        }
        {
            uint32_t r[32];
            tc::tmem_ld32(tlane + TX + 64, r);                          // V1
            tc::wait_ld();
            float v1[32];
#pragma unroll
            for (int u = 0; u < 32; ++u) v1[u] = __uint_as_float(r[u]) + b3[64 + u];
            const float gr[3] = {g.x, g.y, g.z};
#pragma unroll
            for (int c = 0; c < 3; ++c) {                               // Zp = Mp . [V1 (x) r ; p_i ; p_j]   (:131-136, :144)
                if ((c == 0) != (grp == 0)) continue;                   // warp-uniform
#pragma unroll
                for (int h = 0; h < NH; ++h) {
                    float v[32];
                    const float a0 = wp0[h] * gr[c];
#pragma unroll
                    for (int u = 0; u < 32; ++u) v[u] = a0 * v1[u];
                    transpose_reduce<SEG>(v, lane);
#pragma unroll
                    for (int t = 0; t < EPL; ++t) {
                        const float pi = __ldg(sI + 32 + 32 * c + (lane % SEG) * EPL + t);
                        zw[64 + c * 64 + h * 32 + t] = fmaf(wp1s[h], pi, v[t]);
                    }
                }
            }
        }
        if (tile + (int)gridDim.x < n_tiles) {       // next tile's edge slot: hide the index -> gather dependency
            const int in = min((tile + (int)gridDim.x) * TA + a_loc, n_atoms - 1);
            j_next = ids32[(size_t)in * KMAX + k];
            g_next = geom[(size_t)in * KMAX + k];
        }
        tc::fence_before_sync();       // all TMEM reads of this tile are done before the next tile's stores
        __syncthreads();

        // ---------------------------------------------------------------- attention sums Z -> global
        // work unit = (atom, part): part 0 = Zq (64 values), parts 1..3 = Zp[c] incl. the p_j token, summed here with
        // lane = channel and one coalesced 128 B row per edge.  The per-atom projections qpm / ppm run in the next
        // node kernel, where their weights are reused across 8 atoms.
        for (int unit = warp; unit < TA * 4; unit += TC_THREADS / 32) {
            const int a = unit >> 2, part = unit & 3;
            const int ia = tile * TA + a;
            if (ia >= n_atoms) continue;
            const float *z0 = Zs + (WPA == 2 ? a * 2 : a) * 256 + part * 64;
            float zq0 = z0[lane], zq1 = z0[32 + lane];
            if (WPA == 2) { zq0 += z0[256 + lane]; zq1 += z0[256 + 32 + lane]; }
            if (part > 0) {
                const float *ws = Ws + 4 * (a * NN);
                constexpr int PB = NN < 16 ? NN : 16;          // independent row loads in flight per batch
#pragma unroll 1
                for (int e0 = 0; e0 < NN; e0 += PB) {
                    float4 w4[PB];
                    float pj[PB];
#pragma unroll
                    for (int u = 0; u < PB; ++u) w4[u] = *reinterpret_cast<const float4 *>(ws + 4 * (e0 + u));
#pragma unroll
                    for (int u = 0; u < PB; ++u)
                        pj[u] = __ldg(state_in + (size_t)__float_as_int(w4[u].z) * SR + part * 32 + lane);
#pragma unroll
                    for (int u = 0; u < PB; ++u) {
                        zq0 = fmaf(w4[u].x, pj[u], zq0);
                        zq1 = fmaf(w4[u].y, pj[u], zq1);
                    }
                }
            }
            float *zo = Zout + (size_t)(ia + 1) * 256 + part * 64;
            zo[lane] = zq0;
            zo[32 + lane] = zq1;
        }
        __syncthreads();
        tc::fence_after_sync();
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tbase, TM_COLS);
}

template <int NN>
constexpr size_t tc_smem_bytes() {
    constexpr int SEG = NN < 32 ? NN : 32;
    return (size_t)tcimg::TOTAL + (size_t)(4 * (32 / SEG) * 256 + 64 + 128 * 4) * sizeof(float) + 32;
}

template <int NN, bool SPLIT>
int launch_edge_tc(const float *lw, const void *tcw, int n_atoms, const int32_t *ids32, const float *geom,
                   const float *state_in, const float *nodeT, const float *nodeC, float *Zout, cudaStream_t st) {
    static int configured = 0, n_sm = 0;
    constexpr size_t smem = tc_smem_bytes<NN>();
    if (!configured) {
        PESTO_CUDA(cudaFuncSetAttribute(edge_kernel_tc<NN, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int dev = 0;
        PESTO_CUDA(cudaGetDevice(&dev));
        PESTO_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        configured = 1;
    }
    constexpr int TA = 128 / NN;
    const int n_tiles = (n_atoms + TA - 1) / TA;
    const int grid = n_tiles < 2 * n_sm ? n_tiles : 2 * n_sm;
    edge_kernel_tc<NN, SPLIT><<<grid, TC_THREADS, smem, st>>>(lw, (const unsigned char *)tcw, n_atoms, ids32,
                                                            (const float4 *)geom, state_in, nodeT, nodeC, Zout);
    PESTO_CUDA(cudaGetLastError());
    if (getenv("PESTO_TC_DEBUG")) {       // debugging aid: synchronise and report a timed-out tensor-core stage
        PESTO_CUDA(cudaStreamSynchronize(st));
        int wd = 0;
        PESTO_CUDA(cudaMemcpyFromSymbol(&wd, g_tc_watchdog, sizeof(int)));
        if (wd) {
            int zero = 0;
            cudaMemcpyToSymbol(g_tc_watchdog, &zero, sizeof(int));
            set_error("edge_kernel_tc<%d>: tensor-core stage M%d never completed (watchdog)", NN, wd);
            return PESTO_ECUDA;
        }
    }
    return PESTO_OK;
}

template <bool SPLIT>
int dispatch_tc(int nn, const float *lw, const void *tcw, int n_atoms, const int32_t *ids32, const float *geom,
                const float *state_in, const float *nodeT, const float *nodeC, float *Zout, cudaStream_t st) {
    switch (nn) {
        case 8:  return launch_edge_tc<8, SPLIT>(lw, tcw, n_atoms, ids32, geom, state_in, nodeT, nodeC, Zout, st);
        case 16: return launch_edge_tc<16, SPLIT>(lw, tcw, n_atoms, ids32, geom, state_in, nodeT, nodeC, Zout, st);
        case 32: return launch_edge_tc<32, SPLIT>(lw, tcw, n_atoms, ids32, geom, state_in, nodeT, nodeC, Zout, st);
        case 64: return launch_edge_tc<64, SPLIT>(lw, tcw, n_atoms, ids32, geom, state_in, nodeT, nodeC, Zout, st);
        default:
            set_error("state_update: unsupported nn=%d (supported: 8, 16, 32, 64)", nn);
            return PESTO_EINVAL;
    }
}

}  // namespace

// Edge kernel only: attention sums of one layer -> Z[n_atoms+1][256] (row 0 unused).  nodeT / nodeC must hold the
// layer's per-atom factors (launch_node_fused).
int launch_edge_tc_layer(const float *lw, const void *tcw, int nn, int n_atoms, const int32_t *ids32, const float *geom,
                         const float *state_in, float *node_scratch, float *Z, int mode, cudaStream_t st) {
    if (!tcw) {
        set_error("state_update: tensor-core weight images are missing");
        return PESTO_ESTATE;
    }
    const int n_rows = n_atoms + 1;
    float *nodeT = node_scratch;
    float *nodeC = node_scratch + (size_t)n_rows * NODE_T_STRIDE;
    return mode == PESTO_MODE_BF16X3 ? dispatch_tc<true>(nn, lw, tcw, n_atoms, ids32, geom, state_in, nodeT, nodeC, Z, st)
                                     : dispatch_tc<false>(nn, lw, tcw, n_atoms, ids32, geom, state_in, nodeT, nodeC, Z, st);
}

// One complete layer state_in -> state_out (staged API, three launches): head factors, edge kernel, per-atom tail.
// `ev` (optional, 3 events): before the head kernel, between head and edge kernel, after the edge kernel.
int launch_state_update_tc(const float *lw, const void *tcw, int nn, int n_atoms, const int32_t *ids32, const float *geom,
                           const float *state_in, float *state_out, float *node_scratch, float *Z, int mode,
                           cudaStream_t st, cudaEvent_t *ev) {
    if (ev) PESTO_CUDA(cudaEventRecord(ev[0], st));
    int rc = launch_node_fused(nullptr, lw, state_in, nullptr, nullptr, n_atoms, node_scratch, st);
    if (rc != PESTO_OK) return rc;
    if (ev) PESTO_CUDA(cudaEventRecord(ev[1], st));
    rc = launch_edge_tc_layer(lw, tcw, nn, n_atoms, ids32, geom, state_in, node_scratch, Z, mode, st);
    if (rc != PESTO_OK) return rc;
    if (ev) PESTO_CUDA(cudaEventRecord(ev[2], st));
    return launch_node_fused(lw, nullptr, state_in, Z, state_out, n_atoms, node_scratch, st);
}

// ------------------------------------------------------------------------------------------------------------
// UMMA probe: D[128,N] = A[128,K] * B[N,K]^T with A staged in TMEM (thread = row, bf16 packed two per 32-bit
// column) and B in shared memory (K-major, no swizzle), optionally as the 3-term split-bf16 product.  Exercises
// exactly the operand layouts / descriptors the fused kernel relies on (tests/test_gpu_umma.py).
// ------------------------------------------------------------------------------------------------------------
namespace {

__global__ void __launch_bounds__(128)
umma_probe_kernel(const float *__restrict__ A, const float *__restrict__ B, float *__restrict__ D, int K, int N, int split,
                  uint32_t lbo, uint32_t sbo, uint32_t idesc) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t bar;
    __nv_bfloat16 *Bhi = reinterpret_cast<__nv_bfloat16 *>(smem_raw);
    __nv_bfloat16 *Blo = Bhi + (size_t)N * K;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 256);
    if (tid == 0) {
        tc::mbar_init(&bar, 1);
        tc::fence_mbar_init();
    }
    for (int e = tid; e < N * K; e += 128) {                  // B image [K/8][N][8]
        int n = e / K, k = e % K;
        float w = B[e];
        __nv_bfloat16 h = __float2bfloat16_rn(w);
        size_t off = (size_t)(k / 8) * N * 8 + (size_t)n * 8 + (k % 8);
        Bhi[off] = h;
        Blo[off] = __float2bfloat16_rn(w - __bfloat162float(h));
    }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = tmem_base_s;
    const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < K; c0 += 32) {                      // A row -> TMEM: hi at [0, K/2), lo at [64, 64 + K/2)
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            float a = A[(size_t)tid * K + c0 + 2 * u], b = A[(size_t)tid * K + c0 + 2 * u + 1];
            tc::split_bf16x2(a, b, hi[u], lo[u]);
        }
        tc::tmem_st16(lane_addr + c0 / 2, hi);
        tc::tmem_st16(lane_addr + 64 + c0 / 2, lo);
    }
    tc::wait_st();
    tc::fence_before_sync();
    __syncthreads();
    if (tid == 0) {
        tc::fence_after_sync();
        const uint32_t bhi = tc::smem_u32(Bhi), blo = tc::smem_u32(Blo);
        const uint32_t d = tbase + 128;
        uint32_t acc = 0;
        for (int s = 0; s < K / 16; ++s) {
            const uint32_t koff = (uint32_t)s * 2 * (uint32_t)N * 16;     // two 8-element K groups per MMA
            tc::umma_ts(d, tbase + 8 * s, tc::smem_desc(bhi + koff, lbo, sbo), idesc, acc);
            acc = 1;
            if (split) {
                tc::umma_ts(d, tbase + 64 + 8 * s, tc::smem_desc(bhi + koff, lbo, sbo), idesc, 1);
                tc::umma_ts(d, tbase + 8 * s, tc::smem_desc(blo + koff, lbo, sbo), idesc, 1);
            }
        }
        tc::umma_commit(&bar);
    }
    tc::mbar_wait(&bar, 0, &g_tc_watchdog, 9);
    tc::fence_after_sync();
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t r[32];
        tc::tmem_ld32(lane_addr + 128 + c0, r);
        tc::wait_ld();
#pragma unroll
        for (int u = 0; u < 32; ++u)
            if (c0 + u < N) D[(size_t)tid * N + c0 + u] = __uint_as_float(r[u]);
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tbase, 256);
}

}  // namespace
}  // namespace pesto

extern "C" int pesto_debug_umma_probe(const float *A, const float *B, float *D, int K, int N, int split, int lbo, int sbo,
                                      int idesc, void *stream) {
    using namespace pesto;
    if (K % 32 || K < 32 || K > 128 || N % 16 || N < 16 || N > 128) {
        set_error("umma_probe: need K in {32..128 step 32}, N in {16..128 step 16}");
        return PESTO_EINVAL;
    }
    uint32_t l = lbo >= 0 ? (uint32_t)lbo : (uint32_t)N * 16, s = sbo >= 0 ? (uint32_t)sbo : 128u;
    uint32_t id = idesc ? (uint32_t)idesc : tc::idesc_bf16(128, N);
    size_t smem = (size_t)N * K * 2 * 2;
    PESTO_CUDA(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, D, K, N, split, l, s, id);
    PESTO_CUDA(cudaGetLastError());
    return PESTO_OK;
}
