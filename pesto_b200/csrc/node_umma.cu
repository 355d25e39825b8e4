// Per-atom kernel of the tensor-core path on tcgen05 / TMEM.  One launch per layer boundary does, for 128 atoms per
// tile (thread-block of 128 threads, TMEM lane = atom, persistent over tiles, two CTAs per SM),
//   (a) the per-atom tail of the PREVIOUS layer: q += qpm(Zq), p += ppm(Zp)   (src/model_operations.py:147-152),
//       from the attention sums Z the fused edge kernel wrote, producing the new 512 B state record, and
//   (b) the per-atom head of the NEXT layer: |p|, the exact first-layer factors T_j, U_i and the
//       queries Q = nqm([q,|p|]) / sdk                                          (src/model_operations.py:103-119)
// as dense [128 atoms x K] x [K x N] contractions: every Linear is a tcgen05.mma with the activations as the TMEM A
// operand (bf16 hi | lo planes) and the weights as K-major shared-memory images (hi | lo), fp32 accumulation in TMEM;
// SPLIT = 3-term split-bf16 product (parity mode), otherwise one bf16 pass.
//
// All register <-> TMEM / global traffic uses the 16x256b shape: four neighbouring lanes share an atom row and read or
// write 128 (loads) / 32 (stores) contiguous bytes of it, so global accesses are sector-efficient although the rows
// are a kilobyte apart.  Lane (rl = lane / 4, m = lane % 4) of warp w owns the rows 32 w + 8 k + rl (k < 4) and, in
// every 16-column block, the columns 2m, 2m+1, 8+2m, 9+2m.  The K order of every weight image is permuted on the
// host to the order in which this mapping produces the operand words (node_k_global / node_k_regs below).
#include <cstring>

#include "common.cuh"
#include "tc_common.cuh"

namespace pesto {

using L = LayerLayout;

namespace nimg {
// tail section (weights of the layer being finished); every matrix is [K/8][N][8] bf16, hi plane then lo plane
constexpr int T_WQ1 = 0;                         // qpm.0: K = 64 (global order), N = 32
constexpr int T_WP = T_WQ1 + 2 * 4096;           // ppm.0: K = 64 (global order), N = 32
constexpr int T_WQ2 = T_WP + 2 * 4096;           // qpm.2: K = 32 (register order), N = 32
constexpr int T_WQ3 = T_WQ2 + 2 * 2048;          // qpm.4
constexpr int T_BIAS = T_WQ3 + 2 * 2048;         // fp32: b_q1[32] | b_q2[32] | b_q3[32]
constexpr int T_BYTES = T_BIAS + 96 * 4;
// head section (weights of the layer being started)
constexpr int H_WTU = T_BYTES;                   // [T | U] first-layer factors: K = 64 (register order), N = 256, x log2(e)
constexpr int H_WN1 = H_WTU + 2 * 32768;         // nqm.0: K = 64 (register order), N = 32
constexpr int H_WN2 = H_WN1 + 2 * 4096;          // nqm.2: K = 32, N = 32
constexpr int H_WN3 = H_WN2 + 2 * 2048;          // nqm.4: K = 32, N = 16 (12 used), pre-divided by sdk
constexpr int H_BIAS = H_WN3 + 2 * 1024;         // fp32: b_u[128] (x log2 e) | b_n1[32] | b_n2[32] | b_n3[16]
constexpr int TOTAL = H_BIAS + 208 * 4;
constexpr int H_BYTES = TOTAL - T_BYTES;
static_assert(T_BYTES % 16 == 0 && TOTAL % 16 == 0, "images are copied with 16-byte vectors");
}  // namespace nimg

size_t node_tc_layer_bytes() { return nimg::TOTAL; }

namespace {

inline uint16_t h_bf16(float f) { return tc::h16_from_f32_host(f); }      // (configured 16-bit plane format)
inline float h_f32(uint16_t h) { return tc::h16_to_f32_host(h); }
// K position (0..31) of an operand block -> channel, for operand words built from a row piece loaded from global
// memory (lane m reads channels 8m .. 8m+7 and owns word columns 2m, 2m+1, 8+2m, 9+2m)
inline int node_k_global(int p) { return p < 16 ? 8 * (p / 4) + p % 4 : 8 * ((p - 16) / 4) + 4 + (p - 16) % 4; }
// ... and for operand words built from fp32 accumulator registers (lane m holds columns 2m, 2m+1, 8+2m, 9+2m of each
// 16-column block and packs them into the word columns 2m, 2m+1 (block 0) and 8+2m, 9+2m (block 1))
inline int node_k_regs(int p) {
    const int w = p / 2, h = p % 2;
    if (w < 8) return ((w % 2) ? 8 : 0) + 2 * (w / 2) + h;
    return ((w % 2) ? 24 : 16) + 2 * ((w - 8) / 2) + h;
}

// image element (n, kpos) of a [K/8][N][8] bf16 matrix pair (hi plane at dst, lo plane at dst + plane_bytes)
inline void put2(unsigned char *dst, int plane_bytes, int N, int n, int kpos, float w) {
    const size_t e = (size_t)(kpos / 8) * N * 8 + (size_t)n * 8 + (kpos % 8);
    const uint16_t hi = h_bf16(w), lo = h_bf16(w - h_f32(hi));
    ((uint16_t *)dst)[e] = hi;
    ((uint16_t *)(dst + plane_bytes))[e] = lo;
}

}  // namespace

// blob: the layer's fp32 weight block (LayerLayout)
void pack_node_tc_layer(const float *blob, void *dst_v) {
    unsigned char *dst = (unsigned char *)dst_v;
    memset(dst, 0, nimg::TOTAL);
    for (int p = 0; p < 64; ++p) {
        const int kg = 32 * (p / 32) + node_k_global(p % 32), kr = 32 * (p / 32) + node_k_regs(p % 32);
        for (int n = 0; n < 32; ++n) {
            put2(dst + nimg::T_WQ1, 4096, 32, n, p, blob[L::O_Q1 + kg * 32 + n]);
            put2(dst + nimg::T_WP, 4096, 32, n, p, blob[L::O_P + kg * 32 + n]);
            put2(dst + nimg::H_WN1, 4096, 32, n, p, blob[L::NQ_W1 + kr * 32 + n]);
        }
        for (int n = 0; n < 256; ++n) put2(dst + nimg::H_WTU, 32768, 256, n, p, LOG2E * blob[L::N_TU + kr * 256 + n]);
    }
    for (int p = 0; p < 32; ++p) {
        const int kr = node_k_regs(p);
        for (int n = 0; n < 32; ++n) {
            put2(dst + nimg::T_WQ2, 2048, 32, n, p, blob[L::O_Q2 + kr * 32 + n]);
            put2(dst + nimg::T_WQ3, 2048, 32, n, p, blob[L::O_Q3 + kr * 32 + n]);
            put2(dst + nimg::H_WN2, 2048, 32, n, p, blob[L::NQ_W2 + kr * 32 + n]);
        }
        // queries in log2(e) units: the edge kernel's softmax takes 2^(logit - max) straight from MUFU.EX2
        for (int n = 0; n < 16; ++n) put2(dst + nimg::H_WN3, 1024, 16, n, p, LOG2E * blob[L::NQ_W3 + kr * 16 + n]);
    }
    float *tb = (float *)(dst + nimg::T_BIAS);
    for (int i = 0; i < 32; ++i) {
        tb[i] = blob[L::O_Q1B + i];
        tb[32 + i] = blob[L::O_Q2B + i];
        tb[64 + i] = blob[L::O_Q3B + i];
    }
    float *hb = (float *)(dst + nimg::H_BIAS);
    for (int i = 0; i < 128; ++i) hb[i] = LOG2E * blob[L::N_BU + i] - LOG2E;      // (-c: the edge kernel's first E-stage takes shifted input)
    for (int i = 0; i < 32; ++i) {
        hb[128 + i] = blob[L::NQ_B1 + i];
        hb[160 + i] = blob[L::NQ_B2 + i];
    }
    for (int i = 0; i < 16; ++i) hb[192 + i] = LOG2E * blob[L::NQ_B3 + i];
}

namespace {

constexpr unsigned FULLM = 0xffffffffu;
constexpr int NODE_THREADS = 128;
constexpr uint32_t NT_COLS = 256;                 // TMEM columns per CTA
constexpr uint32_t A0 = 0, A1 = 64, DC = 128;     // operand regions (64 columns each) and accumulator region (128)
constexpr int NSM_BAR = nimg::TOTAL;              // 2 mbarriers + TMEM slot (+16) + weight-image barrier (+24)
constexpr int NSM_TOTAL = NSM_BAR + 32;

// ELU without a branch: max(x, 0) + (exp(min(x, 0)) - 1), one MUFU.EX2 (the edge kernel's formulation)
__device__ __forceinline__ float n_elu(float x) {
    const float n = fminf(x, 0.f);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(n * LOG2E));
    return (x - n) + (e - 1.0f);
}

// |p| feeds 16-bit operand planes only (the state record keeps p itself): one MUFU.SQRT (<= 1 ulp) instead of the IEEE
// sequence with its slow-path call
__device__ __forceinline__ float n_sqrt(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// (a, b) -> 16-bit x2 hi word and, if SPLIT, the word of the remainders
template <bool SPLIT>
__device__ __forceinline__ void nsplit(float a, float b, uint32_t &hi, uint32_t &lo) {
    if (SPLIT) tc::split_h16x2(a, b, hi, lo);
    else { hi = tc::pack_h16x2(a, b); lo = 0u; }
}

// fp32 accumulator block (16 columns at taddr, the warp's 32 lanes): v[k][0..3] = row 8k + rl, columns 2m, 2m+1, 8+2m, 9+2m
__device__ __forceinline__ void load_d16(uint32_t taddr, float (&v)[4][4]) {
#pragma unroll
    for (int hb = 0; hb < 2; ++hb) {
        uint32_t r[8];
        tc::tmem_ld_16x256b_x2(taddr + ((uint32_t)(16 * hb) << 16), r);
        tc::wait_ld();
        v[2 * hb][0] = __uint_as_float(r[0]); v[2 * hb][1] = __uint_as_float(r[1]);
        v[2 * hb + 1][0] = __uint_as_float(r[2]); v[2 * hb + 1][1] = __uint_as_float(r[3]);
        v[2 * hb][2] = __uint_as_float(r[4]); v[2 * hb][3] = __uint_as_float(r[5]);
        v[2 * hb + 1][2] = __uint_as_float(r[6]); v[2 * hb + 1][3] = __uint_as_float(r[7]);
    }
}
// operand words of an 8-word-column block (16 channels): w[k][0..1] = row 8k + rl, word columns 2m, 2m+1
template <bool SPLIT>
__device__ __forceinline__ void store_a8(uint32_t t_hi, uint32_t t_lo, const uint32_t (&hi)[4][2], const uint32_t (&lo)[4][2]) {
#pragma unroll
    for (int hb = 0; hb < 2; ++hb) {
        const uint32_t lo16 = (uint32_t)(16 * hb) << 16;
        tc::tmem_st_16x256b_x1(t_hi + lo16, hi[2 * hb][0], hi[2 * hb][1], hi[2 * hb + 1][0], hi[2 * hb + 1][1]);
        if (SPLIT) tc::tmem_st_16x256b_x1(t_lo + lo16, lo[2 * hb][0], lo[2 * hb][1], lo[2 * hb + 1][0], lo[2 * hb + 1][1]);
    }
}

// D[d_col .. d_col + N) = A (K = 16 KSTEPS, hi words at a_col, lo words at a_col + lo_off) . B^T; issued by one thread
// (N = columns of this MMA, NIMG = rows of the [K/8][NIMG][8] weight image the N rows at b_hi / b_lo belong to)
template <bool SPLIT, int KSTEPS, int N, int NIMG = N>
__device__ __forceinline__ void node_gemm(uint32_t tbase, uint32_t d_col, uint32_t a_col, uint32_t lo_off, uint32_t b_hi, uint32_t b_lo) {
    constexpr uint32_t idesc = tc::idesc_h16(128, N);
    constexpr uint32_t lbo = (uint32_t)NIMG * 16u;
#pragma unroll
    for (int s = 0; s < KSTEPS; ++s) {
        const uint32_t a = tbase + a_col + 8u * s;
        const uint32_t koff = (uint32_t)s * 2u * lbo;
        const uint64_t dh = tc::smem_desc(b_hi + koff, lbo, 128u);
        tc::umma_ts(tbase + d_col, a, dh, idesc, s > 0);
        if (SPLIT) {
            tc::umma_ts(tbase + d_col, a + lo_off, dh, idesc, 1u);
            tc::umma_ts(tbase + d_col, a, tc::smem_desc(b_lo + koff, lbo, 128u), idesc, 1u);
        }
    }
}


template <bool SPLIT, bool FUSE_PREV, bool NEXT>
__global__ void __launch_bounds__(NODE_THREADS, 2)
node_umma_kernel(const unsigned char *__restrict__ img_tail, const unsigned char *__restrict__ img_head,
                 const float *__restrict__ state_prev, const float *__restrict__ Z, float *__restrict__ state_new,
                 int n_rows, float *__restrict__ nodeT, float *__restrict__ nodeC, int *__restrict__ wd) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const float *tbias = reinterpret_cast<const float *>(smem_raw + nimg::T_BIAS);
    const float *hbias = reinterpret_cast<const float *>(smem_raw + nimg::H_BIAS);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + NSM_BAR);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem_raw + NSM_BAR + 16);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int m = lane & 3, rl = lane >> 2;
    if (warp == 0) tc::tmem_alloc(tmem_slot, NT_COLS);
    // Prologue = model data only (weight images through the TMA engine, barriers, TMEM): under programmatic dependent launch
    // it overlaps the tail of the previous kernel of the forward; pdl_wait below is where that kernel's outputs are needed.
    uint64_t *wbar = bars + 3;
    if (t == 0) {
        tc::mbar_init(bars, 1);
        tc::mbar_init(bars + 1, 1);
        tc::mbar_init(wbar, (FUSE_PREV ? 1 : 0) + (NEXT ? 1 : 0));
        tc::fence_mbar_init();
        if (FUSE_PREV) tc::bulk_g2s_block(smem_raw, img_tail, nimg::T_BYTES, wbar);
        if (NEXT) tc::bulk_g2s_block(smem_raw + nimg::T_BYTES, img_head + nimg::T_BYTES, nimg::H_BYTES, wbar);
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    tc::mbar_wait(wbar, 0, wd, 17);
    tc::pdl_launch_dependents();
    tc::pdl_wait();
    const int warp_u = __shfl_sync(FULLM, warp, 0);
    const uint32_t tbase = __shfl_sync(FULLM, *tmem_slot, 0);
    const uint32_t tq = tbase + ((uint32_t)(warp_u * 32) << 16);     // warp-uniform: TMEM addresses stay in uniform registers
    const uint32_t sb = tc::smem_u32(smem_raw);
    uint32_t pa = 0, pb = 0;
    bool alive = true;
    float big = 0.f;      // max over this thread's state entries of |q| and |p|^2 (fp16 operand range check, see the kernel's end)

    // block barrier between "operand written / accumulator read" and the next MMA issue
    auto sync_tmem = [&]() {
        tc::wait_st();
        tc::fence_before_sync();
        __syncthreads();
    };
    auto wait_a = [&]() {
        if (alive) alive = tc::mbar_wait(bars, pa, wd, 11);
        pa ^= 1u;
        tc::fence_after_sync();
    };
    auto wait_b = [&]() {
        if (alive) alive = tc::mbar_wait(bars + 1, pb, wd, 12);
        pb ^= 1u;
        tc::fence_after_sync();
    };

    const int n_tiles = (n_rows + 127) / 128;
    // the first two 64-column pieces of a tile's Z rows: loaded one tile ahead (under the wait for the U GEMM of the
    // previous tile), so that a tile starts with its operands in registers instead of with an exposed HBM round trip
    float za[2][4][8], zb[2][4][8];
    auto load_z_tile = [&](int tile_, int col0, float (&v)[2][4][8]) {
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int row = min(tile_ * 128 + warp * 32 + rl + 8 * k, n_rows - 1);
                tc::ldg256(Z + (size_t)row * 256 + col0 + 32 * b + 8 * m, v[b][k]);
            }
    };
    if (FUSE_PREV && (int)blockIdx.x < n_tiles) {
        load_z_tile(blockIdx.x, 0, za);
        load_z_tile(blockIdx.x, 64, zb);
    }
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int rbase = tile * 128 + warp * 32 + rl;           // row of k = 0; rows rbase + 8 k
        int rowc[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) rowc[k] = min(rbase + 8 * k, n_rows - 1);

        // operand (K = 64, global order) <- 64 consecutive floats of each row of Z, in two steps so that the loads of the
        // next operand are in flight while the previous one is converted and multiplied
        auto load_z = [&](int col0, float (&v)[2][4][8]) {
#pragma unroll
            for (int b = 0; b < 2; ++b)
#pragma unroll
                for (int k = 0; k < 4; ++k) tc::ldg256(Z + (size_t)rowc[k] * 256 + col0 + 32 * b + 8 * m, v[b][k]);
        };
        auto conv_z = [&](const float (&v)[2][4][8], uint32_t a_col) {
#pragma unroll
            for (int b = 0; b < 2; ++b)
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    uint32_t hi[4][2], lo[4][2];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        nsplit<SPLIT>(v[b][k][4 * g], v[b][k][4 * g + 1], hi[k][0], lo[k][0]);
                        nsplit<SPLIT>(v[b][k][4 * g + 2], v[b][k][4 * g + 3], hi[k][1], lo[k][1]);
                    }
                    store_a8<SPLIT>(tq + a_col + 16 * b + 8 * g, tq + a_col + 32 + 16 * b + 8 * g, hi, lo);
                }
        };
        // previous state of this thread's rows, 16-column half g: [0] = q, [1 + c] = p_c; [row k][columns 2m.. | 8 + 2m..]
        auto load_state = [&](int g, float2 (&so)[4][4][2]) {
#pragma unroll
            for (int sgm = 0; sgm < 4; ++sgm)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float *sp = state_prev + (size_t)rowc[k] * SR + 32 * sgm + 16 * g + 2 * m;
                    so[sgm][k][0] = tc::ldg64(sp);
                    so[sgm][k][1] = tc::ldg64(sp + 8);
                }
        };
        float2 so[2][4][4][2];
        // operand (K = 32, register order) <- ELU(accumulator[32 columns] + bias)
        auto build_a_elu32 = [&](uint32_t d_col, const float *bias, uint32_t a_col) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                float v[4][4];
                load_d16(tq + d_col + 16 * g, v);
                const float2 b0 = *reinterpret_cast<const float2 *>(bias + 16 * g + 2 * m);
                const float2 b1 = *reinterpret_cast<const float2 *>(bias + 16 * g + 8 + 2 * m);
                uint32_t hi[4][2], lo[4][2];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    nsplit<SPLIT>(n_elu(v[k][0] + b0.x), n_elu(v[k][1] + b0.y), hi[k][0], lo[k][0]);
                    nsplit<SPLIT>(n_elu(v[k][2] + b1.x), n_elu(v[k][3] + b1.y), hi[k][1], lo[k][1]);
                }
                store_a8<SPLIT>(tq + a_col + 8 * g, tq + a_col + 16 + 8 * g, hi, lo);
            }
        };

        if (FUSE_PREV) {
            // ---- tail: q1 = Zq . qpm.0, dp[c] = Zp[c] . ppm.0 (operands double-buffered in A0 / A1, their Z columns
            //      double-buffered in registers one step ahead)
            conv_z(za, A0);                                      // za, zb: loaded one tile ahead
            sync_tmem();
            if (warp_u == 0 && tc::elect_one()) {
                tc::fence_after_sync();
                node_gemm<SPLIT, 4, 32>(tbase, DC + 0, A0, 32, sb + nimg::T_WQ1, sb + nimg::T_WQ1 + 4096);
                tc::umma_commit(bars);
            }
            load_z(128, za);
            conv_z(zb, A1);
            sync_tmem();
            if (warp_u == 0 && tc::elect_one()) {
                tc::fence_after_sync();
                node_gemm<SPLIT, 4, 32>(tbase, DC + 32, A1, 32, sb + nimg::T_WP, sb + nimg::T_WP + 4096);
                tc::umma_commit(bars + 1);
            }
            load_z(192, zb);
            wait_a();
            conv_z(za, A0);
            sync_tmem();
            if (warp_u == 0 && tc::elect_one()) {
                tc::fence_after_sync();
                node_gemm<SPLIT, 4, 32>(tbase, DC + 64, A0, 32, sb + nimg::T_WP, sb + nimg::T_WP + 4096);
                tc::umma_commit(bars);
            }
            load_state(0, so[0]);                                // previous state: in flight under the rest of the tail
            wait_b();
            conv_z(zb, A1);
            sync_tmem();
            if (warp_u == 0 && tc::elect_one()) {
                tc::fence_after_sync();
                node_gemm<SPLIT, 4, 32>(tbase, DC + 96, A1, 32, sb + nimg::T_WP, sb + nimg::T_WP + 4096);
                tc::umma_commit(bars + 1);
            }
            load_state(1, so[1]);
            // ---- qpm layers 2 and 3 (src/model_operations.py:71-77)
            wait_a();                                            // dp[1] done: A0 is free; q1 was done before it
            build_a_elu32(DC + 0, tbias, A0);
            sync_tmem();
            if (warp_u == 0 && tc::elect_one()) {
                tc::fence_after_sync();
                node_gemm<SPLIT, 2, 32>(tbase, DC + 0, A0, 16, sb + nimg::T_WQ2, sb + nimg::T_WQ2 + 2048);
                tc::umma_commit(bars);
            }
            wait_a();
            build_a_elu32(DC + 0, tbias + 32, A0);
            sync_tmem();
            if (warp_u == 0 && tc::elect_one()) {
                tc::fence_after_sync();
                node_gemm<SPLIT, 2, 32>(tbase, DC + 0, A0, 16, sb + nimg::T_WQ3, sb + nimg::T_WQ3 + 2048);
                tc::umma_commit(bars);
            }
            wait_a();
            wait_b();                                            // dp[2] done
        }

        // ---- new state record (residuals :151-152) and, for the next layer, x = [q | |p|] as the operand in A0
        if (!FUSE_PREV) {
            load_state(0, so[0]);
            load_state(1, so[1]);
        }
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            float q[4][4], pn2[4][4];
            const float2 bq0 = FUSE_PREV ? *reinterpret_cast<const float2 *>(tbias + 64 + 16 * g + 2 * m) : make_float2(0.f, 0.f);
            const float2 bq1 = FUSE_PREV ? *reinterpret_cast<const float2 *>(tbias + 64 + 16 * g + 8 + 2 * m) : make_float2(0.f, 0.f);
            if (FUSE_PREV) load_d16(tq + DC + 16 * g, q);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 o0 = so[g][0][k][0], o1 = so[g][0][k][1];
                if (FUSE_PREV && rowc[k] > 0) {                  // row 0 = sink: stays zero
                    q[k][0] += bq0.x + o0.x; q[k][1] += bq0.y + o0.y; q[k][2] += bq1.x + o1.x; q[k][3] += bq1.y + o1.y;
                } else {
                    q[k][0] = o0.x; q[k][1] = o0.y; q[k][2] = o1.x; q[k][3] = o1.y;
                }
                if (FUSE_PREV && rbase + 8 * k < n_rows) {
                    float *dq = state_new + (size_t)(rbase + 8 * k) * SR + 16 * g + 2 * m;
                    *reinterpret_cast<float2 *>(dq) = make_float2(q[k][0], q[k][1]);
                    *reinterpret_cast<float2 *>(dq + 8) = make_float2(q[k][2], q[k][3]);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    pn2[k][u] = 0.f;
                    big = fmaxf(big, q[k][u] * q[k][u]);
                }
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float p[4][4];
                if (FUSE_PREV) load_d16(tq + DC + 32 + 32 * c + 16 * g, p);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float2 o0 = so[g][1 + c][k][0], o1 = so[g][1 + c][k][1];
                    if (FUSE_PREV && rowc[k] > 0) {
                        p[k][0] += o0.x; p[k][1] += o0.y; p[k][2] += o1.x; p[k][3] += o1.y;
                    } else {
                        p[k][0] = o0.x; p[k][1] = o0.y; p[k][2] = o1.x; p[k][3] = o1.y;
                    }
                    if (FUSE_PREV && rbase + 8 * k < n_rows) {
                        float *dp = state_new + (size_t)(rbase + 8 * k) * SR + 32 + 32 * c + 16 * g + 2 * m;
                        *reinterpret_cast<float2 *>(dp) = make_float2(p[k][0], p[k][1]);
                        *reinterpret_cast<float2 *>(dp + 8) = make_float2(p[k][2], p[k][3]);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) pn2[k][u] = fmaf(p[k][u], p[k][u], pn2[k][u]);
                }
            }
            if (NEXT) {
                uint32_t hi[4][2], lo[4][2];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    nsplit<SPLIT>(q[k][0], q[k][1], hi[k][0], lo[k][0]);
                    nsplit<SPLIT>(q[k][2], q[k][3], hi[k][1], lo[k][1]);
                }
                store_a8<SPLIT>(tq + A0 + 8 * g, tq + A0 + 32 + 8 * g, hi, lo);                 // q: K positions 0..31
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    big = fmaxf(fmaxf(big, fmaxf(pn2[k][0], pn2[k][1])), fmaxf(pn2[k][2], pn2[k][3]));
                    nsplit<SPLIT>(n_sqrt(pn2[k][0]), n_sqrt(pn2[k][1]), hi[k][0], lo[k][0]);     // |p| (:105)
                    nsplit<SPLIT>(n_sqrt(pn2[k][2]), n_sqrt(pn2[k][3]), hi[k][1], lo[k][1]);
                }
                store_a8<SPLIT>(tq + A0 + 16 + 8 * g, tq + A0 + 48 + 8 * g, hi, lo);           // |p|: K positions 32..63
            }
        }
        if (NEXT) {
            // ---- head: queries nqm (accumulators and operands in the A1 region) and the factors T, U (accumulator DC)
            sync_tmem();
            if (warp_u == 0 && tc::elect_one()) {
                tc::fence_after_sync();
                node_gemm<SPLIT, 4, 32>(tbase, A1, A0, 32, sb + nimg::H_WN1, sb + nimg::H_WN1 + 4096);
                tc::umma_commit(bars);
                node_gemm<SPLIT, 4, 128, 256>(tbase, DC, A0, 32, sb + nimg::H_WTU, sb + nimg::H_WTU + 32768);
                tc::umma_commit(bars + 1);
            }
            wait_a();
            build_a_elu32(A1, hbias + 128, A1 + 32);
            sync_tmem();
            if (warp_u == 0 && tc::elect_one()) {
                tc::fence_after_sync();
                node_gemm<SPLIT, 2, 32>(tbase, A1, A1 + 32, 16, sb + nimg::H_WN2, sb + nimg::H_WN2 + 2048);
                tc::umma_commit(bars);
            }
            wait_a();
            build_a_elu32(A1, hbias + 160, A1 + 32);
            sync_tmem();
            if (warp_u == 0 && tc::elect_one()) {
                tc::fence_after_sync();
                node_gemm<SPLIT, 2, 16>(tbase, A1, A1 + 32, 16, sb + nimg::H_WN3, sb + nimg::H_WN3 + 1024);
                tc::umma_commit(bars);
            }
            wait_b();                                            // T = x . W_T (already in log2(e) units)
#pragma unroll 2
            for (int g = 0; g < 8; ++g) {
                float v[4][4];
                load_d16(tq + DC + 16 * g, v);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (rbase + 8 * k < n_rows) {
                        float *d = nodeT + (size_t)(rbase + 8 * k) * NODE_T_STRIDE + 16 * g + 2 * m;
                        *reinterpret_cast<float2 *>(d) = make_float2(v[k][0], v[k][1]);
                        *reinterpret_cast<float2 *>(d + 8) = make_float2(v[k][2], v[k][3]);
                    }
            }
            tc::fence_before_sync();
            __syncthreads();                                     // every read of T is done: U may overwrite the accumulator
            if (warp_u == 0 && tc::elect_one()) {
                tc::fence_after_sync();
                node_gemm<SPLIT, 4, 128, 256>(tbase, DC, A0, 32, sb + nimg::H_WTU + 128 * 16, sb + nimg::H_WTU + 32768 + 128 * 16);
                tc::umma_commit(bars + 1);
            }
            if (FUSE_PREV) {          // next tile's first operands: in flight under the U GEMM (unconditional, on a clamped
                                      // tile index: a conditional reload would keep the old values alive across the tile)
                const int tn = min(tile + (int)gridDim.x, n_tiles - 1);
                load_z_tile(tn, 0, za);
                load_z_tile(tn, 64, zb);
            }
            wait_a();                                            // Q = nqm(x) / sdk: 12 of 16 columns
            {
                float v[4][4];
                load_d16(tq + A1, v);
                const float2 b0 = *reinterpret_cast<const float2 *>(hbias + 192 + 2 * m);
                const float2 b1 = *reinterpret_cast<const float2 *>(hbias + 192 + 8 + 2 * m);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (rbase + 8 * k < n_rows) {
                        float *d = nodeC + (size_t)(rbase + 8 * k) * NODE_C_STRIDE + NODE_C_Q + 2 * m;
                        *reinterpret_cast<float2 *>(d) = make_float2(v[k][0] + b0.x, v[k][1] + b0.y);
                        *reinterpret_cast<float2 *>(d + 8) = make_float2(v[k][2] + b1.x, v[k][3] + b1.y);
                    }
            }
            wait_b();                                            // U = x . W_U + b
#pragma unroll 2
            for (int g = 0; g < 8; ++g) {
                float v[4][4];
                load_d16(tq + DC + 16 * g, v);
                const float2 b0 = *reinterpret_cast<const float2 *>(hbias + 16 * g + 2 * m);
                const float2 b1 = *reinterpret_cast<const float2 *>(hbias + 16 * g + 8 + 2 * m);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (rbase + 8 * k < n_rows) {
                        float *d = nodeC + (size_t)(rbase + 8 * k) * NODE_C_STRIDE + 16 * g + 2 * m;
                        *reinterpret_cast<float2 *>(d) = make_float2(v[k][0] + b0.x, v[k][1] + b0.y);
                        *reinterpret_cast<float2 *>(d + 8) = make_float2(v[k][2] + b1.x, v[k][3] + b1.y);
                    }
            }
        }
        if (FUSE_PREV && !NEXT) {
            const int tn = min(tile + (int)gridDim.x, n_tiles - 1);
            load_z_tile(tn, 0, za);
            load_z_tile(tn, 64, zb);
        }
        tc::fence_before_sync();
        __syncthreads();                                         // all TMEM reads of this tile precede the next tile's stores
        tc::fence_after_sync();
    }
    // The tensor-core modes convert the state, its per-atom factors and the edge activations to fp16 planes with saturation
    // (65504).  The shipped checkpoints keep |state| below ~50; a state beyond 2^14 is outside the range these planes were
    // validated for, so it is flagged (code 100 in the forward's watchdog word -> NaN logits, PestoError "use mode fp32")
    // instead of silently saturating somewhere downstream.
    if (big > 16384.f * 16384.f || !(big == big)) atomicMax(wd, 100);
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tbase, NT_COLS);
}

template <bool SPLIT, bool FUSE_PREV, bool NEXT>
int launch_node_umma_variant(const void *img_tail, const void *img_head, const float *state_prev, const float *Z,
                             float *state_new, int n_rows, float *nodeT, float *nodeC, cudaStream_t st, int *wd) {
    int n_sm = 0;
    { const int rc_ = device_setup((const void *)node_umma_kernel<SPLIT, FUSE_PREV, NEXT>, NSM_TOTAL, &n_sm); if (rc_ != PESTO_OK) return rc_; }
    const int n_tiles = (n_rows + 127) / 128;
    const int grid = n_tiles < 2 * n_sm ? n_tiles : 2 * n_sm;
    PESTO_CUDA(launch_pdl(node_umma_kernel<SPLIT, FUSE_PREV, NEXT>, dim3(grid), dim3(NODE_THREADS), NSM_TOTAL, st,
                          (const unsigned char *)img_tail, (const unsigned char *)img_head, state_prev, Z, state_new, n_rows, nodeT,
                          nodeC, wd ? wd : device_watchdog_word()));
    return PESTO_OK;
}

template <bool SPLIT>
int launch_node_umma_mode(const void *img_tail, const void *img_head, const float *state_prev, const float *Z, float *state_new,
                          int n_rows, float *nodeT, float *nodeC, cudaStream_t st, int *wd) {
    if (img_tail && img_head)
        return launch_node_umma_variant<SPLIT, true, true>(img_tail, img_head, state_prev, Z, state_new, n_rows, nodeT, nodeC, st, wd);
    if (img_tail)
        return launch_node_umma_variant<SPLIT, true, false>(img_tail, img_head, state_prev, Z, state_new, n_rows, nodeT, nodeC, st, wd);
    return launch_node_umma_variant<SPLIT, false, true>(img_tail, img_head, state_prev, Z, state_new, n_rows, nodeT, nodeC, st, wd);
}

}  // namespace

// img_tail: node image of the layer being finished (NULL: none); img_head: node image of the layer being started (NULL: none)
int launch_node_umma(const void *img_tail, const void *img_head, const float *state_prev, const float *Z, float *state_new,
                     int n_atoms, float *node_scratch, int mode, cudaStream_t st, int *wd) {
    const int n_rows = n_atoms + 1;
    float *nodeT = node_scratch;
    float *nodeC = node_scratch + (size_t)n_rows * NODE_T_STRIDE;
    return mode == PESTO_MODE_BF16X3
               ? launch_node_umma_mode<true>(img_tail, img_head, state_prev, Z, state_new, n_rows, nodeT, nodeC, st, wd)
               : launch_node_umma_mode<false>(img_tail, img_head, state_prev, Z, state_new, n_rows, nodeT, nodeC, st, wd);
}

}  // namespace pesto
