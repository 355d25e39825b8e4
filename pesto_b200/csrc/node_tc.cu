// Per-atom kernel of the tensor-core path.  One launch per layer boundary does, for 8 atoms per CTA,
//   (a) the per-atom tail of the PREVIOUS layer: q += qpm(Zq), p += ppm(Zp)   (src/model_operations.py:147-152),
//       from the attention sums Z the fused edge kernel wrote, producing the new 512 B state record, and
//   (b) the per-atom head of the NEXT layer: |p|, the exact first-layer factors T_j, U_i and the
//       queries Q = nqm([q,|p|]) / sdk                                          (src/model_operations.py:103-119).
// Weights are reused across the 8 atoms of a CTA from registers (ppm, T/U) -- the per-atom projections used to sit
// at the end of the edge kernel where every weight was loaded for one or two atoms only.
#include "common.cuh"

namespace pesto {

namespace {

using L = LayerLayout;
constexpr int NA = 8;          // atoms per CTA
constexpr unsigned FULLM = 0xffffffffu;

template <bool FUSE_PREV, bool NEXT>
__global__ void __launch_bounds__(128)
node_fused_kernel(const float *__restrict__ lw_prev, const float *__restrict__ lw_next,
                  const float *__restrict__ state_prev, const float *__restrict__ Z, float *__restrict__ state_new,
                  int n_rows, float *__restrict__ nodeT, float *__restrict__ nodeC) {
    __shared__ __align__(16) float xs[NA][160];      // q(32) | p(96) | |p|(32)
    __shared__ __align__(16) float zs[FUSE_PREV ? NA : 1][256];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int r0 = blockIdx.x * NA;
#pragma unroll
    for (int a = 0; a < NA; ++a) {
        const int r = r0 + a;
        xs[a][t] = r < n_rows ? state_prev[(size_t)r * SR + t] : 0.f;
        if (FUSE_PREV) {
            const bool live = r < n_rows && r > 0;           // row 0 = sink: stays zero
            zs[a][t] = live ? Z[(size_t)r * 256 + t] : 0.f;
            zs[a][128 + t] = live ? Z[(size_t)r * 256 + 128 + t] : 0.f;
        }
    }
    __syncthreads();

    if (FUSE_PREV) {
        float res[NA];
        if (warp < 3) {                                      // ppm: component c = warp, output o = lane
            const int c = warp;
#pragma unroll
            for (int a = 0; a < NA; ++a) res[a] = 0.f;
#pragma unroll 4
            for (int k = 0; k < 64; k += 4) {
                float w[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) w[u] = __ldg(lw_prev + L::O_P + (k + u) * 32 + lane);
#pragma unroll
                for (int a = 0; a < NA; ++a) {
                    const float4 z4 = *reinterpret_cast<const float4 *>(&zs[a][64 + c * 64 + k]);
                    res[a] = fmaf(z4.w, w[3], fmaf(z4.z, w[2], fmaf(z4.y, w[1], fmaf(z4.x, w[0], res[a]))));
                }
            }
        } else {                                             // qpm: 64 -> 32 -> 32 -> 32, output o = lane
            float h[NA];
            const float b1 = __ldg(lw_prev + L::O_Q1B + lane);
#pragma unroll
            for (int a = 0; a < NA; ++a) h[a] = b1;
#pragma unroll 4
            for (int k = 0; k < 64; k += 4) {
                float w[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) w[u] = __ldg(lw_prev + L::O_Q1 + (k + u) * 32 + lane);
#pragma unroll
                for (int a = 0; a < NA; ++a) {
                    const float4 z4 = *reinterpret_cast<const float4 *>(&zs[a][k]);
                    h[a] = fmaf(z4.w, w[3], fmaf(z4.z, w[2], fmaf(z4.y, w[1], fmaf(z4.x, w[0], h[a]))));
                }
            }
            float g[NA];
            const float b2 = __ldg(lw_prev + L::O_Q2B + lane), b3 = __ldg(lw_prev + L::O_Q3B + lane);
#pragma unroll
            for (int a = 0; a < NA; ++a) { h[a] = elu(h[a]); g[a] = b2; }
#pragma unroll 8
            for (int k = 0; k < 32; ++k) {
                const float w = __ldg(lw_prev + L::O_Q2 + k * 32 + lane);
#pragma unroll
                for (int a = 0; a < NA; ++a) g[a] = fmaf(__shfl_sync(FULLM, h[a], k), w, g[a]);
            }
#pragma unroll
            for (int a = 0; a < NA; ++a) { g[a] = elu(g[a]); res[a] = b3; }
#pragma unroll 8
            for (int k = 0; k < 32; ++k) {
                const float w = __ldg(lw_prev + L::O_Q3 + k * 32 + lane);
#pragma unroll
                for (int a = 0; a < NA; ++a) res[a] = fmaf(__shfl_sync(FULLM, g[a], k), w, res[a]);
            }
        }
        // residual into the staged record: warp 3 -> q (offset 0), warps 0..2 -> p_c (offset 32 + 32 c)
        const int off = warp < 3 ? 32 + 32 * warp : 0;
#pragma unroll
        for (int a = 0; a < NA; ++a)
            if (r0 + a > 0) xs[a][off + lane] += res[a];
        __syncthreads();
#pragma unroll
        for (int a = 0; a < NA; ++a) {
            const int r = r0 + a;
            if (r < n_rows) state_new[(size_t)r * SR + t] = xs[a][t];
        }
    }
    if (!NEXT) return;

    for (int u = t; u < NA * S; u += 128) {
        const int a = u >> 5, s = u & 31;
        const float x = xs[a][32 + s], y = xs[a][64 + s], z = xs[a][96 + s];
        xs[a][128 + s] = sqrtf(x * x + y * y + z * z);            // |p| (src/model_operations.py:105)
    }
    __syncthreads();

    float accT[NA], accU[NA];
    const float bu = __ldg(lw_next + L::N_BU + t);
#pragma unroll
    for (int a = 0; a < NA; ++a) { accT[a] = 0.f; accU[a] = bu; }
#pragma unroll 1
    for (int k4 = 0; k4 < 8; ++k4) {
        float wTq[4], wUq[4], wTn[4], wUn[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int k = k4 * 4 + u;
            wTq[u] = __ldg(lw_next + L::N_TU + k * 256 + t);
            wUq[u] = __ldg(lw_next + L::N_TU + k * 256 + 128 + t);
            wTn[u] = __ldg(lw_next + L::N_TU + (32 + k) * 256 + t);
            wUn[u] = __ldg(lw_next + L::N_TU + (32 + k) * 256 + 128 + t);
        }
#pragma unroll
        for (int a = 0; a < NA; ++a) {
            const float4 q4 = *reinterpret_cast<const float4 *>(&xs[a][k4 * 4]);
            const float4 n4 = *reinterpret_cast<const float4 *>(&xs[a][128 + k4 * 4]);
            const float qv[4] = {q4.x, q4.y, q4.z, q4.w}, nv[4] = {n4.x, n4.y, n4.z, n4.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                accT[a] = fmaf(wTq[u], qv[u], accT[a]);
                accT[a] = fmaf(wTn[u], nv[u], accT[a]);
                accU[a] = fmaf(wUq[u], qv[u], accU[a]);
                accU[a] = fmaf(wUn[u], nv[u], accU[a]);
            }
        }
    }
#pragma unroll
    for (int a = 0; a < NA; ++a) {
        const int r = r0 + a;
        if (r < n_rows) {
            nodeT[(size_t)r * NODE_T_STRIDE + t] = LOG2E * accT[a];      // the edge kernel works in log2(e)-scaled units
            nodeC[(size_t)r * NODE_C_STRIDE + t] = LOG2E * accU[a];
        }
    }
    // queries nqm([q, |p|]) (src/model_operations.py:119), pre-divided by sdk (:139-140); warp handles 2 atoms
#pragma unroll
    for (int aa = 0; aa < NA / 4; ++aa) {
        const int a = warp * (NA / 4) + aa;
        const int r = r0 + a;
        float h = __ldg(lw_next + L::NQ_B1 + lane);
#pragma unroll 8
        for (int k = 0; k < 32; ++k) {
            h = fmaf(xs[a][k], __ldg(lw_next + L::NQ_W1 + k * 32 + lane), h);
            h = fmaf(xs[a][128 + k], __ldg(lw_next + L::NQ_W1 + (32 + k) * 32 + lane), h);
        }
        h = elu(h);
        float g = __ldg(lw_next + L::NQ_B2 + lane);
#pragma unroll
        for (int k = 0; k < 32; ++k) g = fmaf(__shfl_sync(FULLM, h, k), __ldg(lw_next + L::NQ_W2 + k * 32 + lane), g);
        g = elu(g);
        float o = __ldg(lw_next + L::NQ_B3 + (lane & 15));
#pragma unroll
        for (int k = 0; k < 32; ++k)
            o = fmaf(__shfl_sync(FULLM, g, k), __ldg(lw_next + L::NQ_W3 + k * 16 + (lane & 15)), o);
        if (lane < 16 && r < n_rows) nodeC[(size_t)r * NODE_C_STRIDE + NODE_C_Q + lane] = o;
    }
}

}  // namespace

// lw_prev / Z / state_new may be NULL (no previous layer to finish); lw_next may be NULL (no next layer)
int launch_node_fused(const float *lw_prev, const float *lw_next, const float *state_prev, const float *Z,
                      float *state_new, int n_atoms, float *node_scratch, cudaStream_t st) {
    const int n_rows = n_atoms + 1;
    float *nodeT = node_scratch;
    float *nodeC = node_scratch + (size_t)n_rows * NODE_T_STRIDE;
    const int grid = (n_rows + NA - 1) / NA;
    if (lw_prev && lw_next)
        node_fused_kernel<true, true><<<grid, 128, 0, st>>>(lw_prev, lw_next, state_prev, Z, state_new, n_rows, nodeT, nodeC);
    else if (lw_prev)
        node_fused_kernel<true, false><<<grid, 128, 0, st>>>(lw_prev, lw_next, state_prev, Z, state_new, n_rows, nodeT, nodeC);
    else
        node_fused_kernel<false, true><<<grid, 128, 0, st>>>(lw_prev, lw_next, state_prev, Z, state_new, n_rows, nodeT, nodeC);
    PESTO_CUDA(cudaGetLastError());
    return PESTO_OK;
}

}  // namespace pesto
