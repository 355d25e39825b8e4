// Tensor-core (tcgen05 / TMEM) path of the StateUpdate edge kernel (src/model_operations.py:87-154).
//
// Tile = 128 edge slots = 128/nn atoms.  One CTA per SM with 512 threads = two independent 256-thread tile
// pipelines ("halves") sharing one copy of the weight images; inside a half, thread t and thread t+128 both map to
// edge t % 128 <-> TMEM lane t % 128 and split the columns of every stage.  Per tile:
//
//   S0   gather p_j; A1 = [p_j.r (32) | p_i.r (32) | d, 1(atom) (16)] as 16-bit (hi | lo) planes -> TMEM   (CUDA cores)
//        U_i (per-atom factor) as three 16-bit planes -> spare K rows of B1 in shared memory (nn >= 32).
//        S0 runs one tile AHEAD: its arithmetic under the third-layer MMA of the previous tile, its stores after that tile's
//        E3 (nn = 64: the p_j gather is split between the two column groups)
//   M1   D1[128x128] = A1 . B1^T,  B1 = [W1 cols of p_j.r ; p_i.r ; d ; U planes]                       (tcgen05.mma)
//        issued by four warps (32 accumulator columns each) after the previous tile's E3: it runs under that tile's R
//   E1   h1 = ELU(D1 + T_j)  (T_j: per-atom factor from the node kernel, gathered from L2; for nn >= 32 the first
//        chunk is loaded one tile ahead) -> A2 in place                                                  (CUDA cores)
//   M2   D2 = blockdiag(eqkm.2, epkm.2, evm.2) applied to A2's three column groups (each group issues what it consumes)
//   E2   h2 = ELU(D2 + b2) -> A3 in place
//   M3   D3 = [eqkm.4 | epkm.4 | evm.4] applied to A3's column groups
//   E3   group 0: logits, softmax over the atom's nn / 3nn tokens (warp shuffles) -> attention weights in smem;
//        group 1: V0 | V1 -> smem
//   R    thread = (8-edge group, channel pair): attention-weighted sums of V0, V1 (x) r, p_i, p_j with packed
//        fp32x2 FMAs, partial sums combined through shared memory -> Z[atom][256]
//
// All pre-activations are carried scaled by log2(e) (folded into B1, b2, T, U on one side and 1/log2(e) into B3
// on the other), so ELU needs a bare ex2 and no multiply.  The A operand of every MMA lives in TMEM (written by
// tcgen05.st, 16-bit values packed two per column), B (weights) in shared memory as K-major un-swizzled UMMA
// images prepared on the host at model-finalize time.  SPLIT = true computes hi*hi + lo*hi + hi*lo (3 MMAs per
// K step) over fp16 planes (tc_common.cuh: ~2^-21 relative error, parity mode); SPLIT = false is a single pass (speed mode).
//
// Two rules this file follows because breaking them cost 10-17 % each time (profiles/README.md):
//   * 512 threads x 128 registers is the whole register file: any value kept alive across a phase boundary spills;
//   * a prefetch into a loop-carried array must be UNCONDITIONAL (load a valid dummy row when there is no next tile) and
//     group-specific definitions / uses must test the SAME predicate: otherwise the old values stay live on the path
//     the compiler cannot rule out, i.e. through the whole tile.
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

size_t tc_edge_bytes() { return tcimg::TOTAL; }
size_t tc_layer_bytes() { return tcimg::TOTAL + node_tc_layer_bytes(); }

// K position (0..31) of the p_j.r block of A1 -> state channel: the thread with m = lane % 4 gathers channels 8m..8m+7
// of its row and owns TMEM columns 2m, 2m+1 (first four channels) and 8+2m, 8+2m+1 (last four) of the 16-column plane
static inline int pjr_channel(int p) { return p < 16 ? 8 * (p / 4) + p % 4 : 8 * ((p - 16) / 4) + 4 + (p - 16) % 4; }

// Column (accumulator / operand position 0..31 inside a 32-column chunk) -> channel, for the E-stages' 16x256b register
// mapping: lane m = lane % 4 owns the columns 2m, 2m+1, 8+2m, 9+2m of each 16-column block and is given the eight
// consecutive channels 8m .. 8m+7 of the chunk (block 0: 8m..8m+3, block 1: 8m+4..8m+7), so that four lanes read 128
// contiguous bytes of a gathered T_j row.  The K order this induces on the next GEMM's operand words is pjr_channel.
__host__ __device__ static inline int chunk_channel(int r) {
    const int blk = r / 16, q = r % 16;
    const int m = q < 8 ? q / 2 : (q - 8) / 2, j = q < 8 ? q % 2 : 2 + (q - 8) % 2;
    return 8 * m + 4 * blk + j;
}
__host__ __device__ static inline int col_channel(int c) { return 32 * (c / 32) + chunk_channel(c % 32); }     // any number of 32-column chunks
static inline int kpos_channel(int p) { return 32 * (p / 32) + pjr_channel(p % 32); }

void pack_tc_layer(const float *blob, void *dst_v) {
    unsigned char *dst = (unsigned char *)dst_v;
    memset(dst, 0, tcimg::TOTAL);
    pack_node_tc_layer(blob, dst + tcimg::TOTAL);
    auto put = [&](int img_off, int N, int n, int k, float w) {
        const size_t e = (size_t)(k / 8) * N * 8 + (size_t)n * 8 + (k % 8);
        const uint16_t hi = tc::h16_from_f32_host(w);
        const uint16_t lo = tc::h16_from_f32_host(w - tc::h16_to_f32_host(hi));
        ((uint16_t *)(dst + img_off))[e] = hi;
        ((uint16_t *)(dst + tcimg::IMG + img_off))[e] = lo;
    };
    for (int n = 0; n < 128; ++n) {                    // first layer: accumulator column n holds channel col_channel(n)
        const int o = col_channel(n);
        for (int s = 0; s < 32; ++s) {
            // p_j . r: K position s holds channel pjr_channel(s) (the order in which the gather threads produce them)
            put(tcimg::B1, 128, n, s, LOG2E * blob[L::E_WB + pjr_channel(s) * 128 + o]);
            put(tcimg::B1, 128, n, 32 + s, LOG2E * blob[L::N_A + s * 128 + o]);      // p_i . r
        }
        put(tcimg::B1, 128, n, 64, LOG2E * blob[L::E_WD + o]);                       // d
    }
    for (int n = 0; n < 32; ++n)
        for (int k = 0; k < 32; ++k) {                 // second / third layer of eqkm, epkm: operand K position k, column n
            put(tcimg::B2Q, 32, n, k, blob[L::E_2Q + kpos_channel(k) * 32 + col_channel(n)]);
            put(tcimg::B2P, 32, n, k, blob[L::E_2P + kpos_channel(k) * 32 + col_channel(n)]);
            if (n < 3) put(tcimg::B3Q, 16, n, k, ILOG2E * blob[L::E_3Q + kpos_channel(k) * 4 + n]);
            if (n < 9) put(tcimg::B3P, 16, n, k, ILOG2E * blob[L::E_3P + kpos_channel(k) * 12 + n]);
        }
    for (int n = 0; n < 64; ++n)
        for (int k = 0; k < 64; ++k) {
            put(tcimg::B2V, 64, n, k, blob[L::E_2V + kpos_channel(k) * 64 + col_channel(n)]);
            put(tcimg::B3V, 64, n, k, ILOG2E * blob[L::E_3V + kpos_channel(k) * 64 + n]);
        }
    // Biases, natural channel order (threads index them by channel).  The E-stages work on shifted activations (elu2_shifted):
    // a stage's input carries -c, its output +c, so b2' = c b2 - c - c rowsum(W2) and b3' = b3 - c rowsum(W3 / c), the row sums
    // taken over the operand planes as stored (hi + lo), i.e. over what the tensor core multiplies the +c by
    auto stored = [&](int img_off, int N, int n, int k) {
        const size_t e = (size_t)(k / 8) * N * 8 + (size_t)n * 8 + (k % 8);
        return (double)tc::h16_to_f32_host(((const uint16_t *)(dst + img_off))[e]) +
               (double)tc::h16_to_f32_host(((const uint16_t *)(dst + tcimg::IMG + img_off))[e]);
    };
    auto rowsum = [&](int img_off, int N, int n, int K) {
        double t = 0.0;
        for (int k = 0; k < K; ++k) t += stored(img_off, N, n, k);
        return t;
    };
    float *bias = (float *)(dst + tcimg::BIAS);
    const double c = (double)LOG2E;
    for (int n = 0; n < 32; ++n) {                     // accumulator column n <-> channel col_channel(n)
        bias[col_channel(n)] = (float)(c * blob[L::E_2QB + col_channel(n)] - c - c * rowsum(tcimg::B2Q, 32, n, 32));
        bias[32 + col_channel(n)] = (float)(c * blob[L::E_2PB + col_channel(n)] - c - c * rowsum(tcimg::B2P, 32, n, 32));
    }
    for (int n = 0; n < 64; ++n) {
        bias[64 + col_channel(n)] = (float)(c * blob[L::E_2VB + col_channel(n)] - c - c * rowsum(tcimg::B2V, 64, n, 64));
        bias[128 + 32 + n] = (float)(blob[L::E_3VB + n] - c * rowsum(tcimg::B3V, 64, n, 64));
    }
    for (int i = 0; i < 3; ++i) bias[128 + i] = (float)(blob[L::E_3QB + i] - c * rowsum(tcimg::B3Q, 16, i, 32));
    for (int i = 0; i < 9; ++i) bias[128 + 16 + i] = (float)(blob[L::E_3PB + i] - c * rowsum(tcimg::B3P, 16, i, 32));
}

namespace {

__device__ int g_tc_watchdog = 0;     // UMMA probe only: != 0 = its tensor-core stage timed out
__device__ int g_tc_debug = 0;        // pesto_debug_force_watchdog: bit 0 = the third-layer GEMMs are never committed (a "hung" stage)

constexpr int HALF_THREADS = 256;     // one tile pipeline: 8 warps = two column groups x four TMEM lane quarters
constexpr int CTA_THREADS = 2 * HALF_THREADS;
constexpr unsigned FULLM = 0xffffffffu;
constexpr uint32_t TM_COLS = 512;     // TMEM columns per CTA; half H owns [256 H, 256 H + 256): X = +0, Y = +128
constexpr uint32_t TX = 0, TY = 128;
constexpr int VS_STRIDE = 68;         // floats per edge row of the V0|V1 staging buffer (272 B: conflict-free STS.128)
constexpr int P_STRIDE = 8 * VS_STRIDE;        // floats between the partial sums (256 per 8-edge group) of two groups: each group's
                                               // sums alias its own eight V rows
constexpr int WS_STRIDE = 12;         // floats per edge row of the attention-weight buffer: 11 single weights (48 B: conflict-free
                                      // STS.128); FFMA2 takes them as scalar multipliers.  Round 1 stored every weight as a pair
                                      // (112 B rows): 5.5 instead of 3 loads per edge and thread -- shared-memory traffic, not
                                      // arithmetic, was the cost (all four nn variants -4 .. -6 % with the single weights)
constexpr int WS_GROUP = 8 * WS_STRIDE + 4;      // floats per 8-edge reduction group: 4 floats of skew, so that the two groups a
                                                 // warp reads in one broadcast load sit in different banks (+0.2 %)
constexpr int PROF_STAMPS = 19;       // clock stamps per tile of the debug timeline (pesto_debug_edge_timeline)

// per-half shared memory (byte offsets)
constexpr int HS_EXT_HI = 0;                                  // B1 rows k = 64..79: W_d (hi), U planes per tile, W_d (lo)
constexpr int HS_VS = HS_EXT_HI + 4096;                       // [128][VS_STRIDE] fp32 (E3 -> R); aliased by the partial sums P
constexpr int HS_WS = HS_VS + 128 * VS_STRIDE * 4;            // [16 groups][WS_GROUP] fp32 attention weights (E3 -> R)
constexpr int HS_RED = HS_WS + 16 * WS_GROUP * 4;             // [4 quarters][8] softmax exchange (nn = 64)
constexpr int HS_P = HS_VS;                                   // partial sums of R alias Vs (a buffer of their own takes
constexpr int HS_BYTES = HS_RED + 4 * 8 * 4;                  // the CTA past the 196 KB carve-out: 28 KB of L1 left for the gathers, nn = 64 +3.8 %)
constexpr int SM_PAT = tcimg::TOTAL;                          // [TA <= 4][8] indicator words of the U columns
constexpr int SM_HALF0 = SM_PAT + 128;
constexpr int SM_BAR = SM_HALF0 + 2 * HS_BYTES;               // per half: 4 MMA chunk mbarriers (+ 1 spare); TMEM slot
constexpr int SM_TOTAL = SM_BAR + 96;
static_assert(SM_TOTAL <= 227 * 1024, "shared memory budget of one CTA per SM");
static_assert(SM_HALF0 % 128 == 0 && HS_BYTES % 128 == 0, "per-half regions stay 128-byte aligned");

typedef unsigned long long u64;

// ---- packed fp32x2 arithmetic (sm_100: FFMA2 / FADD2 / FMUL2 on 64-bit register pairs) -----------------------------
__device__ __forceinline__ u64 pk2(float a, float b) {
    u64 r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ u64 pk2u(uint32_t a, uint32_t b) {
    u64 r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ void up2(u64 v, float &a, float &b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
    u64 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ float ex2_fast(float x) {     // one MUFU; inputs below -126 flush to 0
    float t;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(x));
    return t;
}
__device__ __forceinline__ float rcp_fast(float x) {     // one MUFU (normal, finite x)
    float t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(x));
    return t;
}

struct PairConsts {
    u64 neg1, k;           // (-1,-1), (c 2^c, c 2^c) with c = log2(e)
};

// Activations travel scaled by c = log2(e) and shifted: a stage receives u = c x - c (the -c sits in the bias it adds: U_i
// for the first layer, b2' / b3' below) and hands on c ELU(x) + c (the next GEMM's bias takes c * rowsum(W) back out):
//   n = min(u, -c);   c ELU(x) + c = (u - n) + c 2^c 2^n        (u <= -c: c 2^(c x);  u > -c: c x + c)
// one FMNMX, one MUFU and one packed FMA-pipe instruction per element (the unshifted form needs half an instruction more).
__device__ __forceinline__ u64 elu2_shifted(u64 u, const PairConsts &k) {
    float u0, u1;
    up2(u, u0, u1);
    const float n0 = fminf(u0, -LOG2E), n1 = fminf(u1, -LOG2E);
    return fma2(pk2(ex2_fast(n0), ex2_fast(n1)), k.k, fma2(pk2(n0, n1), k.neg1, u));
}
// hi = h16x2(x), lo = h16x2(x - hi)
template <bool SPLIT>
__device__ __forceinline__ void split2(u64 x, const PairConsts &k, uint32_t &hi, uint32_t &lo) {
    float x0, x1;
    up2(x, x0, x1);
    hi = tc::pack_h16x2(x0, x1);
    if (SPLIT && tc::H16_IS_FP16) {      // remainders by one mixed-precision FMA each (nn = 64 kernel -2.5 % against unpack + packed subtract)
        float l0, l1;
        tc::residual_f16x2(hi, x0, x1, l0, l1);
        lo = tc::pack_h16x2(l0, l1);
        return;
    }
    if (SPLIT) {
        float h0, h1;
        tc::unpack_h16x2(hi, h0, h1);
        const u64 l = fma2(pk2(h0, h1), k.neg1, x);
        float l0, l1;
        up2(l, l0, l1);
        lo = tc::pack_h16x2(l0, l1);
    }
}

__device__ __forceinline__ void bar_named(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

template <int SEG>
__device__ __forceinline__ float seg_max_tc(float v) {
    if (SEG == 32) {      // whole warp: one warp-wide reduction instruction (redux.sync.max.f32 -> CREDUX.MAX.F32, sm_100a)
        float r;          // instead of five shuffle + max steps (nn = 32: -0.7 %)
        asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
        return r;
    }
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

// issue D[d_col .. d_col+N) (+)= A(K columns packed at a_col: hi at +8s, lo at +lo_off+8s per 16-wide K step) . B^T
// N = columns of this MMA, NIMG = rows of the weight image ([K/8][NIMG][8]) the N rows belong to; the image's hi plane
// starts B_OFF bytes after the shared-memory base (base14 = base address >> 4), its lo plane tcimg::IMG bytes later
template <bool SPLIT, int KSTEPS, int N, int NIMG, uint32_t B_OFF>
__device__ __forceinline__ void issue_gemm(uint32_t tbase, uint32_t d_col, uint32_t a_col, uint32_t lo_off, uint32_t base14) {
    constexpr uint32_t idesc = tc::idesc_h16(128, N);
    constexpr uint32_t lbo = (uint32_t)NIMG * 16u;
#pragma unroll
    for (int s = 0; s < KSTEPS; ++s) {
        // K steps inside one 32-wide activation chunk are 8 columns apart; chunks are 32 columns apart
        const uint32_t a = tbase + a_col + (uint32_t)(s >> 1) * 32u + (uint32_t)(s & 1) * 8u;
        const uint32_t koff = (uint32_t)s * 2u * lbo;
        const uint64_t dh = tc::smem_desc14(base14, B_OFF + koff, lbo, 128u);
        tc::umma_ts(tbase + d_col, a, dh, idesc, s > 0);
        if (SPLIT) {
            tc::umma_ts(tbase + d_col, a + lo_off, dh, idesc, 1u);
            tc::umma_ts(tbase + d_col, a, tc::smem_desc14(base14, B_OFF + tcimg::IMG + koff, lbo, 128u), idesc, 1u);
        }
    }
}

// E-stage tail in the 16x256b register mapping: y (the addends of this thread's 4 rows x 8 channels of a 32-column
// chunk) += accumulator; ELU; bf16 hi | lo words stored in place over the chunk: [0,16) hi words, [16,32) lo words.
// Thread (rl = lane / 4, m = lane % 4): rows 8 k + rl; block bk, word w <-> columns 16 bk + {2m, 2m+1} (w = 0) and
// 16 bk + {8+2m, 9+2m} (w = 1); as operand words they go to word columns 8 bk + 2m + w (K order pjr_channel).
template <bool SPLIT, bool LD4>
__device__ __forceinline__ void estage_finish(uint32_t tchunk, u64 (&y)[2][4][2], const PairConsts &k) {
    if (LD4) {   // the chunk's four accumulator loads in flight together, one wait (nn >= 16; spills at nn = 8)
        uint32_t r[2][2][8];
#pragma unroll
        for (int bk = 0; bk < 2; ++bk)
#pragma unroll
            for (int hb = 0; hb < 2; ++hb) tc::tmem_ld_16x256b_x2(tchunk + 16 * bk + ((uint32_t)(16 * hb) << 16), r[bk][hb]);
        tc::wait_ld();
#pragma unroll
        for (int bk = 0; bk < 2; ++bk)
#pragma unroll
            for (int hb = 0; hb < 2; ++hb) {
                y[bk][2 * hb][0] = add2(y[bk][2 * hb][0], pk2u(r[bk][hb][0], r[bk][hb][1]));
                y[bk][2 * hb + 1][0] = add2(y[bk][2 * hb + 1][0], pk2u(r[bk][hb][2], r[bk][hb][3]));
                y[bk][2 * hb][1] = add2(y[bk][2 * hb][1], pk2u(r[bk][hb][4], r[bk][hb][5]));
                y[bk][2 * hb + 1][1] = add2(y[bk][2 * hb + 1][1], pk2u(r[bk][hb][6], r[bk][hb][7]));
            }
    } else {
#pragma unroll
    for (int bk = 0; bk < 2; ++bk)
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {
            uint32_t r[8];
            tc::tmem_ld_16x256b_x2(tchunk + 16 * bk + ((uint32_t)(16 * hb) << 16), r);
            tc::wait_ld();
            y[bk][2 * hb][0] = add2(y[bk][2 * hb][0], pk2u(r[0], r[1]));
            y[bk][2 * hb + 1][0] = add2(y[bk][2 * hb + 1][0], pk2u(r[2], r[3]));
            y[bk][2 * hb][1] = add2(y[bk][2 * hb][1], pk2u(r[4], r[5]));
            y[bk][2 * hb + 1][1] = add2(y[bk][2 * hb + 1][1], pk2u(r[6], r[7]));
        }
    }
    uint32_t hi[2][4][2], lo[2][4][2];
#pragma unroll
    for (int bk = 0; bk < 2; ++bk)
#pragma unroll
        for (int kr = 0; kr < 4; ++kr)
#pragma unroll
            for (int w = 0; w < 2; ++w) split2<SPLIT>(elu2_shifted(y[bk][kr][w], k), k, hi[bk][kr][w], lo[bk][kr][w]);
#pragma unroll
    for (int hb = 0; hb < 2; ++hb) {
        const uint32_t ta = tchunk + ((uint32_t)(16 * hb) << 16);
        const uint32_t h8[8] = {hi[0][2 * hb][0], hi[0][2 * hb][1], hi[0][2 * hb + 1][0], hi[0][2 * hb + 1][1],
                                hi[1][2 * hb][0], hi[1][2 * hb][1], hi[1][2 * hb + 1][0], hi[1][2 * hb + 1][1]};
        tc::tmem_st_16x256b_x2(ta, h8);
        if (SPLIT) {
            const uint32_t l8[8] = {lo[0][2 * hb][0], lo[0][2 * hb][1], lo[0][2 * hb + 1][0], lo[0][2 * hb + 1][1],
                                    lo[1][2 * hb][0], lo[1][2 * hb][1], lo[1][2 * hb + 1][0], lo[1][2 * hb + 1][1]};
            tc::tmem_st_16x256b_x2(ta + 16, l8);
        }
    }
}

// One CTA per SM, 512 threads = two independent 256-thread tile pipelines ("halves") that share the weight images
// in shared memory and interleave on the SM's four schedulers: while one half waits for the tensor core or a
// gather, the other runs its CUDA-core stage.  Inside a half, warps 0-3 (group 0) and 4-7 (group 1) both map
// thread -> edge / TMEM lane 32 * (warp % 4) + lane and split the columns of every stage.
template <int NN, bool SPLIT, bool PROF>
__global__ void __launch_bounds__(CTA_THREADS, 1)
edge_kernel_tc(const unsigned char *__restrict__ tcw, int n_atoms, const int32_t *__restrict__ ids32,
               const float4 *__restrict__ geom, const float *__restrict__ state_in, const float *__restrict__ nodeT,
               const float *__restrict__ nodeC, float *__restrict__ Zout, long long *__restrict__ prof, int prof_tiles,
               int *__restrict__ wd) {
    // wd: status word of the forward (or the device's fallback word): a tensor-core stage that never completes records its
    // id there, and the forward's last kernel turns a non-zero word into NaN logits -- a hung MMA is never silent
    constexpr int TA = 128 / NN;                    // atoms per tile
    constexpr int SEG = NN < 32 ? NN : 32;          // lanes of one atom inside a warp
    constexpr int WPA = NN / SEG;                   // warps (of one group) per atom: 2 for nn = 64
    constexpr int GA = NN / 8;                      // 8-edge reduction groups per atom
    constexpr bool UMMA = NN >= 32;                 // U_i enters through spare K columns of the first MMA
    constexpr bool TPREF = NN >= 32;                // first T_j chunk loaded one tile ahead (measured slower at nn <= 16)
    constexpr bool V0BIAS = NN == 64 || NN == 16;   // the attention weights Mq of an atom sum to one: the bias of V0 is added once per
                                                    // atom to Zq instead of once per edge in E3 (nn = 64: -2.1 %, 16: -0.7 %; 8 / 32: +1 %)
    constexpr bool EPLATE = NN <= 16;               // tile epilogue (8 / 16 atoms) deferred into the next tile, where barrier B orders the
                                                    // partial sums: one half-wide barrier less (nn = 8: -5.7 %, 16: -1.8 %; nn >= 32: +3.5 %,
                                                    // the epilogue delays the short second-layer stage there)
    constexpr bool UEARLY = NN == 8;                // U_i loaded with T_j at the tile's start and added there: the loads' latency
                                                    // is not paid once per chunk inside E1 (nn = 8: -3.4 %; nn = 16: +5 %, spills)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char *img = smem_raw;                                      // weight images + biases (shared by both halves)
    const float *b2 = reinterpret_cast<const float *>(img + tcimg::BIAS);
    const float *b3 = b2 + 128;
    const uint32_t *pat = reinterpret_cast<const uint32_t *>(smem_raw + SM_PAT);

    const int tid = threadIdx.x, lane = tid & 31;
    // nn <= 16: warp-uniform role indices, broadcast from lane 0, so that the compiler keeps them (and what derives from them:
    // the half's shared-memory base, barrier addresses) in uniform registers instead of recomputing them from the thread index
    // wherever registers are short: -4 % kernel time at nn = 16, -1 % at nn = 8, but +1.5 % / +5 % at nn = 32 / 64 (measured)
    constexpr bool UNI = NN <= 16;
    const int ht = tid & 255;
    const int hwarp = UNI ? __shfl_sync(FULLM, ht >> 5, 0) : (ht >> 5), H = UNI ? __shfl_sync(FULLM, tid >> 8, 0) : (tid >> 8);
    const int grp = hwarp >> 2, quarter = hwarp & 3;       // column group, TMEM lane quarter
    unsigned char *hs = smem_raw + SM_HALF0 + H * HS_BYTES;
    unsigned char *ext_hi = hs + HS_EXT_HI;
    float *Vs = reinterpret_cast<float *>(hs + HS_VS);
    float *Ws = reinterpret_cast<float *>(hs + HS_WS);
    float *Ps = reinterpret_cast<float *>(hs + HS_P);
    float *red = reinterpret_cast<float *>(hs + HS_RED);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + SM_BAR) + 5 * H;     // [0..3]: column chunks
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem_raw + SM_BAR + 80);

    if (tid < 32) tc::tmem_alloc(tmem_slot, TM_COLS);
    // Prologue = everything that depends on the model only (weight images, barriers, TMEM): under programmatic dependent
    // launch it runs while the previous kernel of the forward is still finishing (see pdl_wait below).  The images come
    // through the TMA engine (one thread, four bulk copies, completion on wbar).
    uint64_t *wbar = reinterpret_cast<uint64_t *>(smem_raw + SM_BAR + 88);
    if (tid == 0) {
        for (int b = 0; b < 10; ++b)
            tc::mbar_init(reinterpret_cast<uint64_t *>(smem_raw + SM_BAR) + b, b % 5 == 4 ? HALF_THREADS : 1);
        tc::mbar_init(wbar, 1);
        tc::fence_mbar_init();
        tc::bulk_g2s_block(img, tcw, tcimg::TOTAL, wbar);
    }
    {   // rows k = 64..79 of B1 -> the per-half, per-tile mutable copy: row 64 = W_d (hi plane), rows 65..76 = U planes
        // (per tile), row 77 = W_d (lo plane; the distance enters twice, so this K step needs no lo-plane MMA)
        uint4 v = __ldg(reinterpret_cast<const uint4 *>(tcw + tcimg::B1 + 8 * 2048) + ht);
        if (ht >= 128) {
            const uint32_t wl = __ldg(reinterpret_cast<const uint32_t *>(tcw + tcimg::IMG + tcimg::B1 + 8 * 2048 + (ht - 128) * 16));
            v.z = (v.z & 0x0000ffffu) | ((wl & 0xffffu) << 16);                         // k = 77: element 5 of K group 1
        }
        reinterpret_cast<uint4 *>(hs + HS_EXT_HI)[ht] = v;
    }
    if (tid < 32) {      // indicator words: column k = 65 + 3 a + p (p < 3) carries plane p of U of the tile's atom a
        const int a = tid >> 3, u = tid & 7;
        uint32_t w = 0;
        if (UMMA && a < TA) {
            const int k0 = 64 + 2 * u, k1 = k0 + 1;
            if (k0 >= 65 + 3 * a && k0 < 68 + 3 * a) w |= tc::H16_ONE;
            if (k1 >= 65 + 3 * a && k1 < 68 + 3 * a) w |= tc::H16_ONE << 16;
        }
        reinterpret_cast<uint32_t *>(smem_raw + SM_PAT)[tid] = w;
    }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    tc::mbar_wait(wbar, 0, wd, 7);            // weight images have landed
    tc::pdl_launch_dependents();              // the next kernel (per-atom kernel) may start its own prologue on free SMs
    tc::pdl_wait();                           // from here on: data of the previous kernels (per-atom factors, state, topology)
    // warp-uniform copies (shuffles from lane 0 let the compiler keep them in uniform registers: the MMA issue code
    // then needs no per-instruction vote / broadcast)
    const int hwarp_u = __shfl_sync(FULLM, hwarp, 0), H_u = __shfl_sync(FULLM, H, 0);
    const uint32_t tbase = __shfl_sync(FULLM, *tmem_slot, 0) + (uint32_t)H_u * 256u;
    const uint32_t tlane = tbase + ((uint32_t)((hwarp_u & 3) * 32) << 16);        // (warp-uniform: TMEM addresses stay in uniform registers)
    const uint32_t img14 = tc::smem_u32(smem_raw) >> 4;                     // operand descriptors: base in 16-byte units
    const uint32_t ext14 = img14 + ((SM_HALF0 + (uint32_t)H_u * HS_BYTES + HS_EXT_HI) >> 4);
    uint64_t *bars_u = reinterpret_cast<uint64_t *>(smem_raw + SM_BAR) + 5 * H_u;
    const int bar_id = 1 + H, bar_g0 = 3 + H;
    uint32_t ph0 = 0, ph1 = 0;               // parities of this thread's two chunk barriers (2 grp, 2 grp + 1)
    uint64_t *bar0 = bars + 2 * grp, *bar1 = bar0 + 1;
    const uint32_t bars_a = tc::smem_u32(bars), bar0_a = bars_a + 16u * (uint32_t)grp, bar1_a = bar0_a + 8u;      // shared-window addresses
    bool alive = true;      // false after a tensor-core stage timed out: finish with garbage (flagged through wd), but finish
    const int dbg = g_tc_debug;
    const uint32_t max_spin = dbg ? 1u << 6 : 1u << 20;
    PairConsts kc;
    kc.neg1 = pk2(-1.f, -1.f);
    kc.k = pk2(ELU_K, ELU_K);

    const int n_tiles = (n_atoms + TA - 1) / TA;
    const int e = ht & 127;                                        // edge slot inside the tile = TMEM lane
    const int a_loc = e / NN, k = e % NN;
    const int tile0 = (int)blockIdx.x * 2 + H, tstride = (int)gridDim.x * 2;
    // optional phase timeline (debug): CTA 0, first thread of each group of each half, PROF_STAMPS clock stamps per tile
    // (PROF is a compile-time switch: even predicated off, the 19 stamps cost ~110 issue slots per tile and thread)
    const bool profiling = PROF && prof != nullptr && blockIdx.x == 0 && (ht & 127) == 0;
    int prof_seq = 0;
#define PROF_STAMP(kk)                                                                                          \
    do {                                                                                                         \
        if (PROF && profiling && prof_seq < prof_tiles) prof[((size_t)prof_seq * 4 + H * 2 + grp) * PROF_STAMPS + (kk)] = clock64(); \
    } while (0)
    // S0 of one tile in two parts.  s0_compute: gather + A1 = [p_j.r | p_i.r] as packed bf16 words in registers (s0h / s0l)
    // and the U_i values; s0_store: words -> TMEM (Y), [d, 1(a), d] columns, U planes -> B1's spare K rows.  S0 runs one
    // tile AHEAD (after E3 of the previous tile, when Y is free again), so that the first-layer MMA of a tile executes
    // while the reduction phase R of the previous tile occupies the CUDA cores.
    // nn = 64: the p_j . r gather of a tile is shared by the two column groups (one gather round trip each instead of two on
    // the group that also computes the attention weights): -2 % there; at nn <= 32 the extra live words spill and it loses
    constexpr bool S0SPLIT = NN == 64;              // (re-measured at the end of round 2 for nn = 32 and all nn: no gain)
    uint32_t s0h[16], s0l[16];          // group 0 (!S0SPLIT): p_j . r words of all four row groups; group 1: p_i . r words
    uint32_t s0ph[8], s0pl[8];          // S0SPLIT: p_j . r words of this thread's two row groups (group g: 2g, 2g + 1)
    float u0v[UMMA ? TA : 1];
    auto s0_compute = [&](int tile, int j, const float4 &g) {
        if (UMMA && grp == 1) {      // U_i of the tile's atoms (group 1 has the lighter S0)
#pragma unroll
            for (int m = 0; m < TA; ++m) {
                const int v = e + 128 * m;
                const int ia = min(tile * TA + (v >> 7), n_atoms - 1);
                u0v[m] = __ldg(nodeC + (size_t)(ia + 1) * NODE_C_STRIDE + col_channel(v & 127));    // column n holds channel col_channel(n)
            }
        }
        if (S0SPLIT) {
            // both groups: rows 8 (2 grp + q) + lane / 4, q < 2, of the warp's 32 edges (see the 4-row-group version below)
            float x[2][8], y[2][8], z[2][8];
            u64 kx[2], ky[2], kz[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int sl = 8 * (2 * grp + q) + (lane >> 2);
                const int jk = __shfl_sync(FULLM, j, sl);
                const float rx = __shfl_sync(FULLM, g.x, sl), ry = __shfl_sync(FULLM, g.y, sl), rz = __shfl_sync(FULLM, g.z, sl);
                kx[q] = pk2(rx, rx); ky[q] = pk2(ry, ry); kz[q] = pk2(rz, rz);
                const float *sJk = state_in + (size_t)jk * SR + 32 + 8 * (lane & 3);
                tc::ldg256(sJk, x[q]);
                tc::ldg256(sJk + 32, y[q]);
                tc::ldg256(sJk + 64, z[q]);
            }
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const u64 pr = fma2(kz[q], pk2(z[q][2 * u], z[q][2 * u + 1]),
                                        fma2(ky[q], pk2(y[q][2 * u], y[q][2 * u + 1]), mul2(kx[q], pk2(x[q][2 * u], x[q][2 * u + 1]))));
                    split2<SPLIT>(pr, kc, s0ph[4 * q + u], s0pl[4 * q + u]);     // [row group q][word u]
                }
        } else if (grp == 0) {
            // p_j . r through the 16x256b store shape: four neighbouring lanes share an edge row and read one 128-byte
            // line of p_j per component (8 lines per load instruction instead of 32); each thread covers the rows
            // 8k + lane/4 (k < 4) of its warp's 32 edges and the channels 8m .. 8m+7, m = lane % 4
            float x[4][8], y[4][8], z[4][8];
            u64 kx[4], ky[4], kz[4];
            auto gather_rows = [&](int k4) {          // rows 8 k4 + lane / 4 of this warp's 32 edges
                const int sl = 8 * k4 + (lane >> 2);
                const int jk = __shfl_sync(FULLM, j, sl);
                const float rx = __shfl_sync(FULLM, g.x, sl), ry = __shfl_sync(FULLM, g.y, sl), rz = __shfl_sync(FULLM, g.z, sl);
                kx[k4] = pk2(rx, rx); ky[k4] = pk2(ry, ry); kz[k4] = pk2(rz, rz);
                const float *sJk = state_in + (size_t)jk * SR + 32 + 8 * (lane & 3);
                tc::ldg256(sJk, x[k4]);
                tc::ldg256(sJk + 32, y[k4]);
                tc::ldg256(sJk + 64, z[k4]);
            };
            auto dot_rows = [&](int k4) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const u64 pr = fma2(kz[k4], pk2(z[k4][2 * u], z[k4][2 * u + 1]),
                                        fma2(ky[k4], pk2(y[k4][2 * u], y[k4][2 * u + 1]), mul2(kx[k4], pk2(x[k4][2 * u], x[k4][2 * u + 1]))));
                    split2<SPLIT>(pr, kc, s0h[4 * k4 + u], s0l[4 * k4 + u]);     // [row k4][word u]
                }
            };
            gather_rows(0); gather_rows(1);
            dot_rows(0); dot_rows(1);
            gather_rows(2); gather_rows(3);
            dot_rows(2); dot_rows(3);
        } else if (!S0SPLIT) {
            const float *src = state_in + (size_t)(min(tile * TA + a_loc, n_atoms - 1) + 1) * SR;
            const u64 gx = pk2(g.x, g.x), gy = pk2(g.y, g.y), gz = pk2(g.z, g.z);
#pragma unroll
            for (int s = 0; s < S; s += 8) {
                float x[8], y[8], z[8];
                tc::ldg256(src + 32 + s, x);
                tc::ldg256(src + 64 + s, y);
                tc::ldg256(src + 96 + s, z);
#pragma unroll
                for (int u = 0; u < 8; u += 2) {
                    const u64 pr = fma2(gz, pk2(z[u], z[u + 1]), fma2(gy, pk2(y[u], y[u + 1]), mul2(gx, pk2(x[u], x[u + 1]))));
                    split2<SPLIT>(pr, kc, s0h[(s + u) >> 1], s0l[(s + u) >> 1]);
                }
            }
        }
    };
    // S0SPLIT: p_j . r words -> TMEM (Y), 16 rows of the warp's quarter per group.  Group 0 stores with the rest of S0 (after
    // the attention stage), group 1 as soon as every MMA of M3 has read Y (in the middle of its V copy), which frees its registers
    auto s0_store_pj = [&]() {
        const uint32_t ta = tlane + ((uint32_t)(16 * grp) << 16) + TY;
        const uint32_t h8[8] = {s0ph[0], s0ph[1], s0ph[4], s0ph[5], s0ph[2], s0ph[3], s0ph[6], s0ph[7]};
        tc::tmem_st_16x256b_x2(ta, h8);
        if (SPLIT) {
            const uint32_t l8[8] = {s0pl[0], s0pl[1], s0pl[4], s0pl[5], s0pl[2], s0pl[3], s0pl[6], s0pl[7]};
            tc::tmem_st_16x256b_x2(ta + 40, l8);
        }
    };
    auto s0_store = [&](int tile, const float4 &g) {
        // hi: Y + [0,16) p_j.r, [16,32) p_i.r, [32,40) d + U indicator columns;  lo: Y + 40 + same.
        if (grp == 0) {
            if (S0SPLIT) s0_store_pj();
#pragma unroll
            for (int hb = 0; hb < (S0SPLIT ? 0 : 2); ++hb) {
                const uint32_t ta = tlane + ((uint32_t)(16 * hb) << 16) + TY;
                const uint32_t h8[8] = {s0h[8 * hb + 0], s0h[8 * hb + 1], s0h[8 * hb + 4], s0h[8 * hb + 5],
                                        s0h[8 * hb + 2], s0h[8 * hb + 3], s0h[8 * hb + 6], s0h[8 * hb + 7]};
                tc::tmem_st_16x256b_x2(ta, h8);
                if (SPLIT) {
                    const uint32_t l8[8] = {s0l[8 * hb + 0], s0l[8 * hb + 1], s0l[8 * hb + 4], s0l[8 * hb + 5],
                                            s0l[8 * hb + 2], s0l[8 * hb + 3], s0l[8 * hb + 6], s0l[8 * hb + 7]};
                    tc::tmem_st_16x256b_x2(ta + 40, l8);
                }
            }
        } else {
            if (S0SPLIT) {     // p_i . r (rows of the centre atom: L1 hits), computed next to its only use so that the words are
                               // not held across the V copy
                const float *src = state_in + (size_t)(min(tile * TA + a_loc, n_atoms - 1) + 1) * SR;
                const u64 gx = pk2(g.x, g.x), gy = pk2(g.y, g.y), gz = pk2(g.z, g.z);
#pragma unroll
                for (int sc = 0; sc < S; sc += 8) {
                    float x[8], y[8], z[8];
                    tc::ldg256(src + 32 + sc, x);
                    tc::ldg256(src + 64 + sc, y);
                    tc::ldg256(src + 96 + sc, z);
#pragma unroll
                    for (int u = 0; u < 8; u += 2) {
                        const u64 pr = fma2(gz, pk2(z[u], z[u + 1]), fma2(gy, pk2(y[u], y[u + 1]), mul2(gx, pk2(x[u], x[u + 1]))));
                        split2<SPLIT>(pr, kc, s0h[(sc + u) >> 1], s0l[(sc + u) >> 1]);
                    }
                }
            }
            tc::tmem_st16(tlane + TY + 16, s0h);
            if (SPLIT) tc::tmem_st16(tlane + TY + 40 + 16, s0l);
            {
                uint32_t hd[8], ld[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) { hd[u] = UMMA ? pat[a_loc * 8 + u] : 0u; ld[u] = 0u; }     // indicator words of this edge's atom
                uint32_t dh, dl = 0u;
                split2<SPLIT>(pk2(g.w, 0.f), kc, dh, dl);
                hd[0] |= dh;                 // k = 64 (x W_d hi) and k = 77 (x W_d lo)
                hd[6] |= dh << 16;
                ld[0] = dl;
                ld[6] = dl << 16;
                tc::tmem_st8(tlane + TY + 32, hd);
                if (SPLIT) tc::tmem_st8(tlane + TY + 72, ld);
            }
        }
        if (UMMA && grp == 1) {      // U_i (already scaled by log2 e) as three bf16 planes -> rows 65 + 3 a + p of B1
#pragma unroll
            for (int m = 0; m < TA; ++m) {
                const int v = e + 128 * m, a = v >> 7, n = v & 127;
                const float u0 = u0v[m];
                const uint16_t h0 = tc::h16_from_f32(u0);
                const float r1 = u0 - tc::h16_to_f32(h0);
                const uint16_t h1 = tc::h16_from_f32(r1);
                const uint16_t h2 = tc::h16_from_f32(r1 - tc::h16_to_f32(h1));
                const uint16_t hp[3] = {h0, h1, h2};
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    const int kk = 1 + 3 * a + p;        // row 64 + kk; rows 64..71 in K group 0, 72..79 in K group 1
                    *reinterpret_cast<uint16_t *>(ext_hi + (kk >> 3) * 2048 + n * 16 + (kk & 7) * 2) = hp[p];
                }
            }
        }
        if (UMMA) tc::fence_async_smem();
        tc::wait_st();
        tc::fence_before_sync();
    };
    // M1: D1 (X) = A1 . B1^T, K = 80, split by accumulator columns over M1S issuing warps (each a chain over 128 / M1S
    // columns with its own chunk commits).  An issuing warp is held back while the tensor-core queue drains -- and M1 is the
    // longest chain (0.9 k cycles) -- so one issuer arrives that much later at the reduction's barrier than the other seven
    // warps; four issuers (warps 0, 2, 4, 6) are held back a quarter as long (+1.9 %; +2.4 % at nn = 8 since round 2)
    constexpr int M1S = 4;
    auto issue_m1 = [&]() {
        constexpr int NC = 128 / M1S;                                   // columns per issuer
        if ((hwarp_u % (8 / M1S)) == 0 && tc::elect_one()) {
            const uint32_t c = M1S == 1 ? 0u : (uint32_t)hwarp_u / (8 / M1S);      // column block of this issuer
            tc::fence_after_sync();
            constexpr uint32_t idesc = tc::idesc_h16(128, NC);
            constexpr uint32_t lbo = 128u * 16u;
            const uint32_t r14 = c * NC;                                 // rows NC c .. of every K group, in 16-byte units
#pragma unroll
            for (int s = 0; s < 5; ++s) {
                const uint64_t dh = s < 4 ? tc::smem_desc14(img14 + r14, tcimg::B1 + (uint32_t)s * 2u * lbo, lbo, 128u)
                                          : tc::smem_desc14(ext14 + r14, 0u, lbo, 128u);
                tc::umma_ts(tbase + TX + c * NC, tbase + TY + 8u * s, dh, idesc, s > 0);
                if (SPLIT) {
                    tc::umma_ts(tbase + TX + c * NC, tbase + TY + 40u + 8u * s, dh, idesc, 1u);
                    if (s < 4)
                        tc::umma_ts(tbase + TX + c * NC, tbase + TY + 8u * s,
                                    tc::smem_desc14(img14 + r14, tcimg::IMG + tcimg::B1 + (uint32_t)s * 2u * lbo, lbo, 128u), idesc, 1u);
                }
            }
#pragma unroll
            for (int q = 0; q < 4 / M1S; ++q) tc::umma_commit(bars_u + c * (4 / M1S) + q);      // chunk barriers of these columns
        }
    };
    // E-stage register mapping (16x256b): this thread owns the rows 8 k + rl (k < 4) of its warp's 32 edges and, in a
    // 32-column chunk c, the channels 32 c + 8 m4 .. + 7.  T_j of those rows and channels: one 32-byte load per row
    // and chunk, four lanes reading one 128-byte line.
    const int m4 = lane & 3, rl = lane >> 2;
    float tv[2][4][8];
    auto load_T = [&](int jv, int cc) {
        int jrow[4];
#pragma unroll
        for (int kr = 0; kr < 4; ++kr) jrow[kr] = __shfl_sync(FULLM, jv, 8 * kr + rl);
#pragma unroll
        for (int kr = 0; kr < 4; ++kr)
            tc::ldg256(nodeT + (size_t)jrow[kr] * NODE_T_STRIDE + 64 * grp + 32 * cc + 8 * m4, tv[cc][kr]);
    };
    int j_next = 0;
    float4 g_next = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tile0 < n_tiles) {
        const int i0 = min(tile0 * TA + a_loc, n_atoms - 1);
        j_next = ids32[(size_t)i0 * KMAX + k];
        g_next = geom[(size_t)i0 * KMAX + k];
        {
            s0_compute(tile0, j_next, g_next);
            if (S0SPLIT && grp != 0) s0_store_pj();
            s0_store(tile0, g_next);
            if (TPREF) load_T(j_next, 0);
            bar_named(bar_id, HALF_THREADS);
            issue_m1();
        }
    }
    // Tile epilogue: Z record of atom a = [Zq h*32+s | Zp c*64 + h*32 + s], the sum of the reduction groups' partial sums; the
    // per-atom projections qpm / ppm run in the next node kernel, where their weights are reused across 8 atoms
    const float zb = V0BIAS && ht < 64 ? b3[32 + (ht & 31)] : 0.f;      // bias of V0 (evm.4 rows 0..31), both heads of Zq
    auto epilogue = [&](int t) {
#pragma unroll
        for (int a = 0; a < TA; ++a) {
            float z = zb;
#pragma unroll
            for (int gg = 0; gg < GA; ++gg) z += Ps[(a * GA + gg) * P_STRIDE + ht];
            const int io = t * TA + a;
            if (io < n_atoms) Zout[(size_t)(io + 1) * 256 + ht] = z;
        }
    };
    int ep_tile = -1;
    for (int tile = tile0; tile < n_tiles; tile += tstride) {
        PROF_STAMP(0);
        const int i = min(tile * TA + a_loc, n_atoms - 1);         // tail tile: clamp (results are not written)
        const int j = j_next;
        const float4 g = g_next;
        const float *cI = nodeC + (size_t)(i + 1) * NODE_C_STRIDE;
        const bool more = tile + tstride < n_tiles;
        int jn = 0;                                                // the next tile's edge slot (S0 of that tile runs after E3)
        float4 gn = make_float4(0.f, 0.f, 0.f, 0.f);
        if (more) {
            const int in = min((tile + tstride) * TA + a_loc, n_atoms - 1);
            jn = ids32[(size_t)in * KMAX + k];
            gn = geom[(size_t)in * KMAX + k];
        }
        PROF_STAMP(1);
        PROF_STAMP(2);
        PROF_STAMP(3);
        // T_j: for nn >= 32 the first chunk was loaded one tile ahead, after the reduction loop (both chunks ahead: slower)
        if (!TPREF) load_T(j, 0);
        load_T(j, 1);
        if (UEARLY) {       // U_i of the row's atom, added to T_j while both are in flight (not between the chunks' waits)
#pragma unroll
            for (int cc = 0; cc < 2; ++cc)
#pragma unroll
                for (int kr = 0; kr < 4; ++kr) {
                    const int ik = min(tile * TA + (quarter * 32 + 8 * kr + rl) / NN, n_atoms - 1);
                    float pu[8];
                    tc::ldg256(nodeC + (size_t)(ik + 1) * NODE_C_STRIDE + 32 * (2 * grp + cc) + 8 * m4, pu);
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        up2(add2(pk2(tv[cc][kr][2 * e], tv[cc][kr][2 * e + 1]), pk2(pu[2 * e], pu[2 * e + 1])),
                            tv[cc][kr][2 * e], tv[cc][kr][2 * e + 1]);
                }
        }
        PROF_STAMP(4);

        // ---------------------------------------------------------------- E1: h1 = ELU(D1 + T_j [+ U_i]) -> A2 (X, in place)
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
            const int c = 2 * grp + cc;
            u64 y[2][4][2];                                  // [16-column block][row][word = channel pair]
#pragma unroll
            for (int bk = 0; bk < 2; ++bk)
#pragma unroll
                for (int kr = 0; kr < 4; ++kr) {
                    y[bk][kr][0] = pk2(tv[cc][kr][4 * bk], tv[cc][kr][4 * bk + 1]);
                    y[bk][kr][1] = pk2(tv[cc][kr][4 * bk + 2], tv[cc][kr][4 * bk + 3]);
                }
            if (!UMMA && !UEARLY) {                          // U_i of the row's atom (nn = 16: not folded into the MMA)
#pragma unroll
                for (int kr = 0; kr < 4; ++kr) {
                    const int ik = min(tile * TA + (quarter * 32 + 8 * kr + rl) / NN, n_atoms - 1);
                    float pu[8];
                    tc::ldg256(nodeC + (size_t)(ik + 1) * NODE_C_STRIDE + 32 * c + 8 * m4, pu);
#pragma unroll
                    for (int bk = 0; bk < 2; ++bk) {
                        y[bk][kr][0] = add2(y[bk][kr][0], pk2(pu[4 * bk], pu[4 * bk + 1]));
                        y[bk][kr][1] = add2(y[bk][kr][1], pk2(pu[4 * bk + 2], pu[4 * bk + 3]));
                    }
                }
            }
            if (alive) alive = tc::mbar_wait_a(cc ? bar1_a : bar0_a, cc ? ph1 : ph0, wd, 1, max_spin);
            tc::fence_after_sync();
            estage_finish<SPLIT, (NN >= 16)>(tlane + TX + 32 * c, y, kc);
        }
        tc::wait_st();
        tc::fence_before_sync();
        PROF_STAMP(5);
        bar_named(bar_id, HALF_THREADS);   // (per-group barriers here and before M3 -- M2 / M3 are block diagonal by group -- need an
        PROF_STAMP(6);                     //  "all of M1 done" guard before M2 overwrites A1 in Y, and then measure +0.1 %: not kept)
        // M2: D2 (Y) = blockdiag(eqkm.2, epkm.2, evm.2).  Each column group issues the GEMMs it consumes itself (an issuing
        // warp is held back while the tensor-core queue drains, so it should only wait for its own results); separate
        // commits: the ELU stage of the first chunks overlaps the remaining MMAs
        if (hwarp_u == 0 && tc::elect_one()) {
            tc::fence_after_sync();
            issue_gemm<SPLIT, 2, 32, 32, tcimg::B2Q>(tbase, TY + 0, TX + 0, 16, img14);
            tc::umma_commit(bars_u + 0);
            issue_gemm<SPLIT, 2, 32, 32, tcimg::B2P>(tbase, TY + 32, TX + 32, 16, img14);
            tc::umma_commit(bars_u + 1);
        }
        if (hwarp_u == 4 && tc::elect_one()) {
            tc::fence_after_sync();
            issue_gemm<SPLIT, 4, 64, 64, tcimg::B2V>(tbase, TY + 64, TX + 64, 16, img14);
            tc::umma_commit(bars_u + 2);
            tc::umma_commit(bars_u + 3);
        }
        ph0 ^= 1u;
        ph1 ^= 1u;
        PROF_STAMP(7);
        // nn <= 16: the previous tile's epilogue, in the shadow of this tile's second-layer MMAs: every warp has passed barrier B,
        // so all partial sums are written (they sit in the V buffer, which is not written again before barrier C)
        if (EPLATE && ep_tile >= 0) epilogue(ep_tile);

        // ---------------------------------------------------------------- E2: h2 = ELU(D2 + b2) -> A3 (Y, in place)
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
            const int c = 2 * grp + cc;
            u64 y[2][4][2];
#pragma unroll
            for (int bk = 0; bk < 2; ++bk) {
                const ulonglong2 bb = *reinterpret_cast<const ulonglong2 *>(b2 + 32 * c + 8 * m4 + 4 * bk);
#pragma unroll
                for (int kr = 0; kr < 4; ++kr) { y[bk][kr][0] = bb.x; y[bk][kr][1] = bb.y; }
            }
            if (alive) alive = tc::mbar_wait_a(cc ? bar1_a : bar0_a, cc ? ph1 : ph0, wd, 2, max_spin);
            tc::fence_after_sync();
            estage_finish<SPLIT, (NN >= 16)>(tlane + TY + 32 * c, y, kc);
        }
        tc::wait_st();
        tc::fence_before_sync();
        PROF_STAMP(8);
        bar_named(bar_id, HALF_THREADS);
        PROF_STAMP(9);
        // M3: D3 (X) = [eqkm.4 | epkm.4 | evm.4]; Kq | Kp for group 0 (attention weights), V0 | V1 for group 1
        if (hwarp_u == 0 && tc::elect_one()) {
            tc::fence_after_sync();
            issue_gemm<SPLIT, 2, 16, 16, tcimg::B3Q>(tbase, TX + 0, TY + 0, 16, img14);
            issue_gemm<SPLIT, 2, 16, 16, tcimg::B3P>(tbase, TX + 16, TY + 32, 16, img14);
            if (!(dbg & 1)) {
                tc::umma_commit(bars_u + 0);
                tc::umma_commit(bars_u + 1);                             // (keeps three phases per tile on every barrier)
            }
        }
        if (hwarp_u == 4 && tc::elect_one()) {
            tc::fence_after_sync();
            issue_gemm<SPLIT, 4, 64, 64, tcimg::B3V>(tbase, TX + 32, TY + 64, 16, img14);
            tc::umma_commit(bars_u + 2);
            tc::umma_commit(bars_u + 3);
        }
        ph0 ^= 1u;
        ph1 ^= 1u;
        // neighbour ids of this thread's 8-edge reduction group (phase R), while the tensor core works
        const int pair = ht & 15, rg = ht >> 4;
        const int iaR = min(tile * TA + (rg * 8) / NN, n_atoms - 1);
        int4 idr[2];
        {
            const int4 *idp = reinterpret_cast<const int4 *>(ids32 + (size_t)iaR * KMAX + (rg * 8) % NN);
            idr[0] = __ldg(idp);
            idr[1] = __ldg(idp + 1);
        }
        // queries of the centre atom: [t][h][k] = t*6 + h*3 + k, pre-divided by sdk
        float qv[12];
        if (grp == 0) {
            const float *Qi = cI + NODE_C_Q;
#pragma unroll
            for (int u = 0; u < 12; u += 4) {
                const float4 q4 = __ldg(reinterpret_cast<const float4 *>(Qi + u));
                qv[u] = q4.x; qv[u + 1] = q4.y; qv[u + 2] = q4.z; qv[u + 3] = q4.w;
            }
        }
        u64 pjr[8][3];       // p_j of this thread's 8-edge reduction group (phase R)
#define PJR_LOAD()                                                                                                       \
        {                                                                                                                \
            const int jr[8] = {idr[0].x, idr[0].y, idr[0].z, idr[0].w, idr[1].x, idr[1].y, idr[1].z, idr[1].w};          \
            _Pragma("unroll") for (int ee = 0; ee < 8; ++ee) {                                                           \
                const float *pJ = state_in + (size_t)jr[ee] * SR + 32 + 2 * pair;                                        \
                _Pragma("unroll") for (int c = 0; c < 3; ++c) pjr[ee][c] = tc::ldg64u(pJ + 32 * c);                      \
            }                                                                                                            \
        }
        // S0 arithmetic of the next tile in the shadow of this tile's third-layer MMA (unconditional for the same reason as
        // the T_j prefetch: on the last tile it recomputes this tile's words, which are never stored)
        s0_compute(more ? tile + tstride : tile, jn, gn);
        if (alive) alive = tc::mbar_wait_a(bar0_a, ph0, wd, 3, max_spin);     // group 0: Kq | Kp; group 1: V0
        tc::fence_after_sync();
        PROF_STAMP(10);

        // ---------------------------------------------------------------- E3
        PROF_STAMP(17);
        if (grp == 0) {
            // attention weights of this edge (src/model_operations.py:139-140) -> Ws row, every weight duplicated
            // into a pair so that the reduction below can use packed FMAs without register shuffling
            uint32_t rq[4], rp[16];
            tc::tmem_ld4(tlane + TX, rq);                               // [0,3) Kq
            tc::tmem_ld16(tlane + TX + 16, rp);                         // [16,25) Kp
            tc::wait_ld();
            float kq[3], kp[9];
#pragma unroll
            for (int u = 0; u < 3; ++u) kq[u] = __uint_as_float(rq[u]) + b3[u];
#pragma unroll
            for (int u = 0; u < 9; ++u) kp[u] = __uint_as_float(rp[u]) + b3[16 + u];
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
                if (lane == 0) *reinterpret_cast<float4 *>(red + quarter * 8) = make_float4(mx[0], mx[1], mx[2], mx[3]);
                bar_named(bar_g0, 128);
                const float4 o = *reinterpret_cast<const float4 *>(red + (quarter ^ 1) * 8);
                mx[0] = fmaxf(mx[0], o.x); mx[1] = fmaxf(mx[1], o.y); mx[2] = fmaxf(mx[2], o.z); mx[3] = fmaxf(mx[3], o.w);
            }
            float eq[NH], ep[NH][3], sm[4];
#pragma unroll
            for (int h = 0; h < NH; ++h) {
                eq[h] = ex2_fast(lq[h] - mx[h]);                  // (Q arrives scaled by log2(e) from the per-atom kernel)
#pragma unroll
                for (int gk = 0; gk < 3; ++gk) ep[h][gk] = ex2_fast(lp[h][gk] - mx[2 + h]);
            }
            sm[0] = seg_sum_tc<SEG>(eq[0]);
            sm[1] = seg_sum_tc<SEG>(eq[1]);
            sm[2] = seg_sum_tc<SEG>(ep[0][0] + ep[0][1] + ep[0][2]);
            sm[3] = seg_sum_tc<SEG>(ep[1][0] + ep[1][1] + ep[1][2]);
            if (WPA == 2) {
                if (lane == 0) *reinterpret_cast<float4 *>(red + quarter * 8 + 4) = make_float4(sm[0], sm[1], sm[2], sm[3]);
                bar_named(bar_g0, 128);
                const float4 o = *reinterpret_cast<const float4 *>(red + (quarter ^ 1) * 8 + 4);
                sm[0] += o.x; sm[1] += o.y; sm[2] += o.z; sm[3] += o.w;
            }
            // the sums lie in [1, 3 nn] (the largest term is exp(0)): one MUFU.RCP each (<= 1 ulp) instead of the IEEE division
            // with its range fix-up code where that measures faster (nn <= 32: -0.6 .. -1.2 % per launch; nn = 64: +1 %)
            constexpr bool RCP = NN <= 32;
            const float iq0 = RCP ? rcp_fast(sm[0]) : 1.0f / sm[0], iq1 = RCP ? rcp_fast(sm[1]) : 1.0f / sm[1];
            const float ip0 = RCP ? rcp_fast(sm[2]) : 1.0f / sm[2], ip1 = RCP ? rcp_fast(sm[3]) : 1.0f / sm[3];
            const float wq0 = eq[0] * iq0, wq1 = eq[1] * iq1;                     // Mq[h]
            const float wv0 = ep[0][0] * ip0, wv1 = ep[1][0] * ip1;               // Mp[h, token V1 (x) r]
            const float wi0 = ep[0][1] * ip0, wi1 = ep[1][1] * ip1;               // Mp[h, token p_i]
            const float wj0 = ep[0][2] * ip0, wj1 = ep[1][2] * ip1;               // Mp[h, token p_j]
            float4 *row = reinterpret_cast<float4 *>(Ws + (e >> 3) * WS_GROUP + (e & 7) * WS_STRIDE);      // 11 weights, single: FFMA2 takes a scalar multiplier
            row[0] = make_float4(wq0, wq1, wj0, wj1);
            row[1] = make_float4(wv0 * g.x, wv0 * g.y, wv0 * g.z, wv1 * g.x);
            row[2] = make_float4(wv1 * g.y, wv1 * g.z, wi0, wi1);
        } else {
            // V0 | V1 (+ bias) of this edge -> Vs row
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                if (half == 1) {
                    if (alive) alive = tc::mbar_wait_a(bar1_a, ph1, wd, 3, max_spin);
                    if (S0SPLIT && alive) alive = tc::mbar_wait_a(bars_a + 8u, ph1, wd, 3, max_spin);      // ... and Kq | Kp: Y is free
                    tc::fence_after_sync();
                    if (S0SPLIT) s0_store_pj();
                }
#pragma unroll
                for (int q16 = 0; q16 < 2; ++q16) {
                    uint32_t r[16];
                    tc::tmem_ld16(tlane + TX + 32 + 32 * half + 16 * q16, r);
                    tc::wait_ld();
                    float4 *dst = reinterpret_cast<float4 *>(Vs + e * VS_STRIDE + 32 * half + 16 * q16);
                    if (V0BIAS && half == 0) {      // V0 without its bias (added once per atom to Zq in EP below)
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            *reinterpret_cast<uint4 *>(dst + u) = make_uint4(r[4 * u], r[4 * u + 1], r[4 * u + 2], r[4 * u + 3]);
                        continue;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const ulonglong2 bb = *reinterpret_cast<const ulonglong2 *>(b3 + 32 + 32 * half + 16 * q16 + 4 * u);
                        ulonglong2 o;
                        o.x = add2(pk2u(r[4 * u], r[4 * u + 1]), bb.x);
                        o.y = add2(pk2u(r[4 * u + 2], r[4 * u + 3]), bb.y);
                        *reinterpret_cast<ulonglong2 *>(dst + u) = o;
                    }
                }
            }
        }
        ph0 ^= 1u;      // every chunk barrier completes three times per tile (M1, M2, M3)
        ph1 ^= 1u;
        PROF_STAMP(18);
        if (more) {
            // every MMA of M3 has read Y: each group has waited for its own GEMMs, now for the other group's last commit
            if (alive) alive = tc::mbar_wait_a(bars_a + (grp == 0 ? 24u : 8u), ph1 ^ 1u, wd, 3, max_spin);
            tc::fence_after_sync();
            s0_store(tile + tstride, gn);
        }
        // p_j of the reduction group's 8 edges (phase R): issued before the barrier so that part of the gather latency overlaps it
        PJR_LOAD();
        // ---------------------------------------------------------------- R: attention-weighted sums over the edges
        // thread = (8-edge group rg, channel pair): Zq = Mq . V0 (:143), Zp = Mp . [V1 (x) r ; p_i ; p_j] (:131-136, :144)
        {
            tc::fence_before_sync();       // all TMEM reads of this tile are done before the next tile's stores
            PROF_STAMP(11);
            bar_named(bar_id, HALF_THREADS);
            PROF_STAMP(12);
            if (more) issue_m1();          // the next tile's first-layer MMA runs underneath this tile's reduction
            const float *pI = state_in + (size_t)(iaR + 1) * SR + 32 + 2 * pair;
            u64 zq[2], zp[3][2], wi = 0ull;
            zq[0] = zq[1] = 0ull;
#pragma unroll
            for (int c = 0; c < 3; ++c) zp[c][0] = zp[c][1] = 0ull;
            const float *vrow = Vs + (rg * 8) * VS_STRIDE + 2 * pair;
            const float *wrow = Ws + rg * WS_GROUP;
#pragma unroll
            for (int ee = 0; ee < 8; ++ee) {
                const float4 *wr = reinterpret_cast<const float4 *>(wrow + ee * WS_STRIDE);
                const float4 a = wr[0], b = wr[1], c = wr[2];
                const u64 v0 = *reinterpret_cast<const u64 *>(vrow + ee * VS_STRIDE);
                const u64 v1 = *reinterpret_cast<const u64 *>(vrow + ee * VS_STRIDE + 32);
                zq[0] = fma2(pk2(a.x, a.x), v0, zq[0]);
                zq[1] = fma2(pk2(a.y, a.y), v0, zq[1]);
                zp[0][0] = fma2(pk2(b.x, b.x), v1, zp[0][0]);
                zp[1][0] = fma2(pk2(b.y, b.y), v1, zp[1][0]);
                zp[2][0] = fma2(pk2(b.z, b.z), v1, zp[2][0]);
                zp[0][1] = fma2(pk2(b.w, b.w), v1, zp[0][1]);
                zp[1][1] = fma2(pk2(c.x, c.x), v1, zp[1][1]);
                zp[2][1] = fma2(pk2(c.y, c.y), v1, zp[2][1]);
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) {
                    zp[cc][0] = fma2(pk2(a.z, a.z), pjr[ee][cc], zp[cc][0]);
                    zp[cc][1] = fma2(pk2(a.w, a.w), pjr[ee][cc], zp[cc][1]);
                }
                wi = add2(wi, pk2(c.z, c.w));             // (the last two words of the row's third load: an aligned register pair)
            }
            float wi0, wi1;
            up2(wi, wi0, wi1);
            const u64 w0 = pk2(wi0, wi0), w1 = pk2(wi1, wi1);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const u64 pi = __ldg(reinterpret_cast<const u64 *>(pI + 32 * c));
                zp[c][0] = fma2(w0, pi, zp[c][0]);
                zp[c][1] = fma2(w1, pi, zp[c][1]);
            }
            PROF_STAMP(13);
            __syncwarp();                             // the partial sums of a reduction group overwrite that group's OWN eight V rows,
                                                      // which only its 16 threads (half of this warp) read: no half-wide barrier
            u64 *P = reinterpret_cast<u64 *>(Ps + rg * P_STRIDE + 2 * pair);
            P[0] = zq[0];
            P[16] = zq[1];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                P[32 + 32 * c] = zp[c][0];
                P[48 + 32 * c] = zp[c][1];
            }
            if (TPREF) {          // unconditional (row 0 when there is no next tile): a load under `if (more)` would keep tv alive
                load_T(jn, 0);    // -- and 32 registers occupied -- through the whole tile
            }
            if (EPLATE) {
                ep_tile = tile;   // the tile's epilogue runs after barrier B of the half's next tile (or after the loop)
            } else {
                bar_named(bar_id, HALF_THREADS);
                PROF_STAMP(14);
                epilogue(tile);
            }
        }
        PROF_STAMP(15);

        PROF_STAMP(16);
        ++prof_seq;
        j_next = jn;
        g_next = gn;
    }
#undef PROF_STAMP
    if (EPLATE && ep_tile >= 0) {                      // the half's last tile
        bar_named(bar_id, HALF_THREADS);
        epilogue(ep_tile);
    }
    tc::fence_before_sync();
    __syncthreads();
    if (tid < 32) tc::tmem_dealloc(*tmem_slot, TM_COLS);
}

long long *g_prof_buf = nullptr;      // host-side: device buffer for the debug timeline (nullptr = off)
int g_prof_tiles = 0;

template <int NN, bool SPLIT>
int launch_edge_tc(const void *tcw, int n_atoms, const int32_t *ids32, const float *geom, const float *state_in,
                   const float *nodeT, const float *nodeC, float *Zout, cudaStream_t st, int *wd) {
    int n_sm = 0;
    // the debug timeline (pesto_debug_edge_timeline) exists for the parity mode's nn = 64 kernel only
#ifdef PESTO_PROF_ALL_NN                            // measurement builds only (profiles/variants.py): timeline stamps for every nn
    constexpr bool CAN_PROF = SPLIT;
#else
    constexpr bool CAN_PROF = NN == 64 && SPLIT;
#endif
    const bool prof_on = CAN_PROF && g_prof_buf != nullptr;
    auto kernel = prof_on ? edge_kernel_tc<NN, SPLIT, CAN_PROF> : edge_kernel_tc<NN, SPLIT, false>;
    { const int rc_ = device_setup((const void *)kernel, SM_TOTAL, &n_sm); if (rc_ != PESTO_OK) return rc_; }
    constexpr int TA = 128 / NN;
    const int n_tiles = (n_atoms + TA - 1) / TA;
    const int grid = (n_tiles + 1) / 2 < n_sm ? (n_tiles + 1) / 2 : n_sm;
    if (!wd) wd = device_watchdog_word();
    PESTO_CUDA(launch_pdl(kernel, dim3(grid), dim3(CTA_THREADS), SM_TOTAL, st, (const unsigned char *)tcw, n_atoms,
                          ids32, (const float4 *)geom, state_in, nodeT, nodeC, Zout, g_prof_buf, g_prof_tiles, wd));
    return PESTO_OK;
}

template <bool SPLIT>
int dispatch_tc(int nn, const void *tcw, int n_atoms, const int32_t *ids32, const float *geom, const float *state_in,
                const float *nodeT, const float *nodeC, float *Zout, cudaStream_t st, int *wd) {
    switch (nn) {
        case 8:  return launch_edge_tc<8, SPLIT>(tcw, n_atoms, ids32, geom, state_in, nodeT, nodeC, Zout, st, wd);
        case 16: return launch_edge_tc<16, SPLIT>(tcw, n_atoms, ids32, geom, state_in, nodeT, nodeC, Zout, st, wd);
        case 32: return launch_edge_tc<32, SPLIT>(tcw, n_atoms, ids32, geom, state_in, nodeT, nodeC, Zout, st, wd);
        case 64: return launch_edge_tc<64, SPLIT>(tcw, n_atoms, ids32, geom, state_in, nodeT, nodeC, Zout, st, wd);
        default:
            set_error("state_update: unsupported nn=%d (supported: 8, 16, 32, 64)", nn);
            return PESTO_EINVAL;
    }
}

}  // namespace

// Edge kernel only: attention sums of one layer -> Z[n_atoms+1][256] (row 0 unused).  nodeT / nodeC must hold the
// layer's per-atom factors (launch_node_umma).
int launch_edge_tc_layer(const float *lw, const void *tcw, int nn, int n_atoms, const int32_t *ids32, const float *geom,
                         const float *state_in, float *node_scratch, float *Z, int mode, cudaStream_t st, int *wd) {
    if (!tcw) {
        set_error("state_update: tensor-core weight images are missing");
        return PESTO_ESTATE;
    }
    const int n_rows = n_atoms + 1;
    float *nodeT = node_scratch;
    float *nodeC = node_scratch + (size_t)n_rows * NODE_T_STRIDE;
    return mode == PESTO_MODE_BF16X3 ? dispatch_tc<true>(nn, tcw, n_atoms, ids32, geom, state_in, nodeT, nodeC, Z, st, wd)
                                     : dispatch_tc<false>(nn, tcw, n_atoms, ids32, geom, state_in, nodeT, nodeC, Z, st, wd);
}

// One complete layer state_in -> state_out (staged API, three launches): head factors, edge kernel, per-atom tail.
// `ev` (optional, 3 events): before the head kernel, between head and edge kernel, after the edge kernel.
int launch_state_update_tc(const float *lw, const void *tcw, int nn, int n_atoms, const int32_t *ids32, const float *geom,
                           const float *state_in, float *state_out, float *node_scratch, float *Z, int mode,
                           cudaStream_t st, cudaEvent_t *ev, int *wd) {
    if (ev) PESTO_CUDA(cudaEventRecord(ev[0], st));
    const void *nimg = tcw ? (const void *)((const unsigned char *)tcw + tc_edge_bytes()) : nullptr;
    if (!nimg) {
        set_error("state_update: tensor-core weight images are missing");
        return PESTO_ESTATE;
    }
    int rc = launch_node_umma(nullptr, nimg, state_in, nullptr, nullptr, n_atoms, node_scratch, mode, st, wd);
    if (rc != PESTO_OK) return rc;
    if (ev) PESTO_CUDA(cudaEventRecord(ev[1], st));
    rc = launch_edge_tc_layer(lw, tcw, nn, n_atoms, ids32, geom, state_in, node_scratch, Z, mode, st, wd);
    if (rc != PESTO_OK) return rc;
    if (ev) PESTO_CUDA(cudaEventRecord(ev[2], st));
    return launch_node_umma(nimg, nullptr, state_in, Z, state_out, n_atoms, node_scratch, mode, st, wd);
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
    uint16_t *Bhi = reinterpret_cast<uint16_t *>(smem_raw);
    uint16_t *Blo = Bhi + (size_t)N * K;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 256);
    if (tid == 0) {
        tc::mbar_init(&bar, 1);
        tc::fence_mbar_init();
    }
    for (int e = tid; e < N * K; e += 128) {                  // B image [K/8][N][8]
        int n = e / K, k = e % K;
        float w = B[e];
        const uint16_t h = tc::h16_from_f32(w);
        size_t off = (size_t)(k / 8) * N * 8 + (size_t)n * 8 + (k % 8);
        Bhi[off] = h;
        Blo[off] = tc::h16_from_f32(w - tc::h16_to_f32(h));
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
            tc::split_h16x2(a, b, hi[u], lo[u]);
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
    uint32_t id = idesc ? (uint32_t)idesc : tc::idesc_h16(128, N);
    size_t smem = (size_t)N * K * 2 * 2;
    PESTO_CUDA(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, D, K, N, split, l, s, id);
    PESTO_CUDA(cudaGetLastError());
    return PESTO_OK;
}

/* Debug: while buf != NULL every tensor-core edge-kernel launch records, for CTA 0, clock64() stamps at 17 phase
 * boundaries per tile into buf[tile_seq][half][group][17] (device memory, int64) for the first max_tiles tiles of
 * each half.  Stamp order: 0 tile start, 1 S0 done, 2 barrier, 3 (unused), 4 M1 done, 5 E1 done, 6 barrier,
 * 7 M2 done, 8 E2 done, 9 barrier, 10 M3 done, 11 E3 done, 12 barrier, 13 R loop done, 14 partial sums visible,
 * 15 Z written, 16 next T copies issued, 17 p_j prefetch issued (inside E3), 18 E3 arithmetic done. */
/* Debug: on != 0 makes every tensor-core edge kernel skip the completion signal of its third-layer GEMMs (and shortens its
 * waits), i.e. simulates a tensor-core stage that hangs: the kernels must finish, and the forward must report it. */
extern "C" int pesto_debug_force_watchdog(int on) {
    const int v = on ? 1 : 0;
    PESTO_CUDA(cudaMemcpyToSymbol(pesto::g_tc_debug, &v, sizeof(int)));
    return PESTO_OK;
}

extern "C" int pesto_debug_edge_timeline(void *buf, int max_tiles) {
    pesto::g_prof_buf = (long long *)buf;
    pesto::g_prof_tiles = buf ? max_tiles : 0;
    return PESTO_OK;
}
