// One StateUpdateLayer (src/model_operations.py:225-242 -> StateUpdate.forward :87-154), fp32 FFMA path.
//
// Two launches per layer:
//   node_kernel   per atom:  |p|, the exact per-atom factors of the first edge-MLP layer
//                 (T_j, U_i, A_i -- SURVEY.md A.3) and the queries Q = nqm([q,|p|]) / sdk;
//   edge_kernel   persistent CTAs over tiles of 128 edge slots (128/nn atoms):
//                 phase A  thread = edge: gather p_j and T_j, finish layer 1, layers 2-3 of eqkm / epkm / evm
//                          from shared-memory weights (broadcast float4 loads), attention logits,
//                          softmax over the atom's nn (scalar) and 3*nn (vector) tokens with warp shuffles;
//                 phase B  thread = channel: attention-weighted sums Zq, Zp over the atom's edges;
//                 phase C  warp = atom: qpm / ppm projections, residual, write the new 512 B state record.
// The 193-wide edge feature, the gathered q_nn / p_nn tensors, the MLP hiddens, V and Vp of the reference
// are never materialised in HBM.
#include "common.cuh"

namespace pesto {

namespace {

using L = LayerLayout;

// ------------------------------------------------------------------------------------------------------------
// node kernel
// ------------------------------------------------------------------------------------------------------------
constexpr int NODE_ATOMS = 8;
constexpr int NODE_XS = 160;   // q(32) | p(96) | pn(32)

__global__ void __launch_bounds__(128)
node_kernel(const float *__restrict__ lw, const float *__restrict__ state, int n_rows, float *__restrict__ nodeT,
            float *__restrict__ nodeC) {
    __shared__ __align__(16) float xs[NODE_ATOMS][NODE_XS];
    const int t = threadIdx.x;
    const int r0 = blockIdx.x * NODE_ATOMS;
#pragma unroll
    for (int a = 0; a < NODE_ATOMS; ++a) {
        int r = r0 + a;
        xs[a][t] = r < n_rows ? state[(size_t)r * SR + t] : 0.f;
    }
    __syncthreads();
    for (int u = t; u < NODE_ATOMS * S; u += 128) {
        int a = u >> 5, s = u & 31;
        float x = xs[a][32 + s], y = xs[a][64 + s], z = xs[a][96 + s];
        xs[a][128 + s] = sqrtf(x * x + y * y + z * z);            // |p| (src/model_operations.py:105)
    }
    __syncthreads();

    float accT[NODE_ATOMS], accU[NODE_ATOMS], accA[3][NODE_ATOMS];
    const float bu = lw[L::N_BU + t];
#pragma unroll
    for (int a = 0; a < NODE_ATOMS; ++a) {
        accT[a] = 0.f;
        accU[a] = bu;
        accA[0][a] = accA[1][a] = accA[2][a] = 0.f;
    }
#pragma unroll 1
    for (int k4 = 0; k4 < 8; ++k4) {
        float wTq[4], wUq[4], wTn[4], wUn[4], wA[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            int k = k4 * 4 + u;
            wTq[u] = __ldg(lw + L::N_TU + k * 256 + t);
            wUq[u] = __ldg(lw + L::N_TU + k * 256 + 128 + t);
            wTn[u] = __ldg(lw + L::N_TU + (32 + k) * 256 + t);
            wUn[u] = __ldg(lw + L::N_TU + (32 + k) * 256 + 128 + t);
            wA[u] = __ldg(lw + L::N_A + k * 128 + t);
        }
#pragma unroll
        for (int a = 0; a < NODE_ATOMS; ++a) {
            float4 q4 = *(const float4 *)&xs[a][k4 * 4];
            float4 x4 = *(const float4 *)&xs[a][32 + k4 * 4];
            float4 y4 = *(const float4 *)&xs[a][64 + k4 * 4];
            float4 z4 = *(const float4 *)&xs[a][96 + k4 * 4];
            float4 n4 = *(const float4 *)&xs[a][128 + k4 * 4];
            const float qv[4] = {q4.x, q4.y, q4.z, q4.w}, xv[4] = {x4.x, x4.y, x4.z, x4.w};
            const float yv[4] = {y4.x, y4.y, y4.z, y4.w}, zv[4] = {z4.x, z4.y, z4.z, z4.w};
            const float nv[4] = {n4.x, n4.y, n4.z, n4.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                accT[a] = fmaf(wTq[u], qv[u], accT[a]);
                accT[a] = fmaf(wTn[u], nv[u], accT[a]);
                accU[a] = fmaf(wUq[u], qv[u], accU[a]);
                accU[a] = fmaf(wUn[u], nv[u], accU[a]);
                accA[0][a] = fmaf(wA[u], xv[u], accA[0][a]);
                accA[1][a] = fmaf(wA[u], yv[u], accA[1][a]);
                accA[2][a] = fmaf(wA[u], zv[u], accA[2][a]);
            }
        }
    }
#pragma unroll
    for (int a = 0; a < NODE_ATOMS; ++a) {
        int r = r0 + a;
        if (r < n_rows) {
            nodeT[(size_t)r * NODE_T_STRIDE + t] = accT[a];
            float *c = nodeC + (size_t)r * NODE_C_STRIDE;
            c[t] = accU[a];
            c[128 + t] = accA[0][a];
            c[256 + t] = accA[1][a];
            c[384 + t] = accA[2][a];
        }
    }
    // queries: nqm([q, |p|]) (src/model_operations.py:119), already divided by sdk (:139-140)
    const int lane = t & 31, w = t >> 5;
#pragma unroll
    for (int aa = 0; aa < NODE_ATOMS / 4; ++aa) {
        int a = w * (NODE_ATOMS / 4) + aa;
        int r = r0 + a;
        float h = lw[L::NQ_B1 + lane];
#pragma unroll 8
        for (int k = 0; k < 32; ++k) {
            h = fmaf(xs[a][k], __ldg(lw + L::NQ_W1 + k * 32 + lane), h);
            h = fmaf(xs[a][128 + k], __ldg(lw + L::NQ_W1 + (32 + k) * 32 + lane), h);
        }
        h = elu(h);
        float g = lw[L::NQ_B2 + lane];
#pragma unroll
        for (int k = 0; k < 32; ++k) g = fmaf(__shfl_sync(0xffffffffu, h, k), __ldg(lw + L::NQ_W2 + k * 32 + lane), g);
        g = elu(g);
        float o = lw[L::NQ_B3 + (lane & 15)];
#pragma unroll
        for (int k = 0; k < 32; ++k)
            o = fmaf(__shfl_sync(0xffffffffu, g, k), __ldg(lw + L::NQ_W3 + k * 16 + (lane & 15)), o);
        if (lane < 16 && r < n_rows) nodeC[(size_t)r * NODE_C_STRIDE + NODE_C_Q + lane] = o;
    }
}

// ------------------------------------------------------------------------------------------------------------
// edge kernel
// ------------------------------------------------------------------------------------------------------------
constexpr int EDGE_THREADS = 128;
constexpr int VS_STRIDE = 68;     // floats per edge row of V in shared memory (conflict-free float4 rows)
constexpr int ES_STRIDE = 16;     // per-edge attention scalars
constexpr size_t EDGE_SMEM = (size_t)(L::E_SIZE + EDGE_THREADS * VS_STRIDE + EDGE_THREADS * ES_STRIDE + 64) * sizeof(float);

// y[o] += sum_k x[k] * W[k][o]   (W in shared memory, k-major, row stride OP; all lanes read the same address)
template <int K, int O, int OP>
__device__ __forceinline__ void gemv_acc(const float *__restrict__ W, const float (&x)[K], float (&y)[O]) {
    static_assert(O % 4 == 0, "outputs are processed four at a time");
#pragma unroll
    for (int o = 0; o < O; o += 4) {
        float a0 = y[o], a1 = y[o + 1], a2 = y[o + 2], a3 = y[o + 3];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float4 w = *reinterpret_cast<const float4 *>(W + k * OP + o);
            a0 = fmaf(x[k], w.x, a0);
            a1 = fmaf(x[k], w.y, a1);
            a2 = fmaf(x[k], w.z, a2);
            a3 = fmaf(x[k], w.w, a3);
        }
        y[o] = a0; y[o + 1] = a1; y[o + 2] = a2; y[o + 3] = a3;
    }
}

template <int O>
__device__ __forceinline__ void load_bias(const float *__restrict__ b, float (&y)[O]) {
#pragma unroll
    for (int o = 0; o < O; o += 4) {
        const float4 v = *reinterpret_cast<const float4 *>(b + o);
        y[o] = v.x; y[o + 1] = v.y; y[o + 2] = v.z; y[o + 3] = v.w;
    }
}

template <int O>
__device__ __forceinline__ void elu_inplace(float (&y)[O]) {
#pragma unroll
    for (int o = 0; o < O; ++o) y[o] = elu(y[o]);
}

// first edge-MLP layer for outputs [O0, O0+NO):  U_i + T_j + w_d d + r.A_i + W_B (p_j.r)   
template <int O0, int NO>
__device__ __forceinline__ void first_layer(const float *__restrict__ wE, const float *__restrict__ cI,
                                            const float *__restrict__ tJ, const float4 g, const float (&pr)[S],
                                            float (&h)[NO]) {
#pragma unroll
    for (int o = 0; o < NO; o += 4) {
        const float4 u = __ldg(reinterpret_cast<const float4 *>(cI + O0 + o));
        const float4 ax = __ldg(reinterpret_cast<const float4 *>(cI + 128 + O0 + o));
        const float4 ay = __ldg(reinterpret_cast<const float4 *>(cI + 256 + O0 + o));
        const float4 az = __ldg(reinterpret_cast<const float4 *>(cI + 384 + O0 + o));
        const float4 tj = __ldg(reinterpret_cast<const float4 *>(tJ + O0 + o));
        const float4 wd = *reinterpret_cast<const float4 *>(wE + (L::E_WD - L::E_BEGIN) + O0 + o);
        h[o + 0] = fmaf(g.z, az.x, fmaf(g.y, ay.x, fmaf(g.x, ax.x, fmaf(g.w, wd.x, u.x + tj.x))));
        h[o + 1] = fmaf(g.z, az.y, fmaf(g.y, ay.y, fmaf(g.x, ax.y, fmaf(g.w, wd.y, u.y + tj.y))));
        h[o + 2] = fmaf(g.z, az.z, fmaf(g.y, ay.z, fmaf(g.x, ax.z, fmaf(g.w, wd.z, u.z + tj.z))));
        h[o + 3] = fmaf(g.z, az.w, fmaf(g.y, ay.w, fmaf(g.x, ax.w, fmaf(g.w, wd.w, u.w + tj.w))));
    }
    gemv_acc<S, NO, 128>(wE + (L::E_WB - L::E_BEGIN) + O0, pr, h);
    elu_inplace(h);
}

template <int SEG>
__device__ __forceinline__ float seg_max(float v) {
#pragma unroll
    for (int o = SEG / 2; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
template <int SEG>
__device__ __forceinline__ float seg_sum(float v) {
#pragma unroll
    for (int o = SEG / 2; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int NN>
__global__ void __launch_bounds__(EDGE_THREADS, 2)
edge_kernel_fp32(const float *__restrict__ lw, int n_atoms, const int32_t *__restrict__ ids32,
                 const float4 *__restrict__ geom, const float *__restrict__ state_in,
                 const float *__restrict__ nodeT, const float *__restrict__ nodeC, float *__restrict__ state_out) {
    constexpr int TA = EDGE_THREADS / NN;           // atoms per tile
    constexpr int SEG = NN < 32 ? NN : 32;          // lanes of one atom inside a warp
    extern __shared__ __align__(16) float smem[];
    float *wE = smem;                                // edge weights, resident for the whole launch
    float *Vs = wE + L::E_SIZE;                      // [128][VS_STRIDE] values; reused as Zs[TA][256]
    float *Es = Vs + EDGE_THREADS * VS_STRIDE;       // [128][ES_STRIDE]
    float *red = Es + EDGE_THREADS * ES_STRIDE;      // [4 warps][8] cross-warp softmax partials (NN = 64)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int u = tid; u < L::E_SIZE / 4; u += EDGE_THREADS)
        reinterpret_cast<float4 *>(wE)[u] = __ldg(reinterpret_cast<const float4 *>(lw + L::E_BEGIN) + u);
    if (blockIdx.x == 0) state_out[tid] = 0.f;       // sink row stays zero (src/model_operations.py:239-240)
    __syncthreads();

    const int n_tiles = (n_atoms + TA - 1) / TA;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        // ------------------------------------------------------------------ phase A: thread = edge
        {
            const int a = tid / NN, k = tid % NN;
            const int i = min(tile * TA + a, n_atoms - 1);     // tail tile: clamp (results are not written)
            const int j = ids32[(size_t)i * KMAX + k];
            const float4 g = geom[(size_t)i * KMAX + k];
            const float *cI = nodeC + (size_t)(i + 1) * NODE_C_STRIDE;
            const float *tJ = nodeT + (size_t)j * NODE_T_STRIDE;
            const float *sJ = state_in + (size_t)j * SR;

            float pr[S];                                        // p_j . r   (src/model_operations.py:115)
#pragma unroll
            for (int s = 0; s < S; s += 4) {
                const float4 x = __ldg(reinterpret_cast<const float4 *>(sJ + 32 + s));
                const float4 y = __ldg(reinterpret_cast<const float4 *>(sJ + 64 + s));
                const float4 z = __ldg(reinterpret_cast<const float4 *>(sJ + 96 + s));
                pr[s + 0] = fmaf(g.z, z.x, fmaf(g.y, y.x, g.x * x.x));
                pr[s + 1] = fmaf(g.z, z.y, fmaf(g.y, y.y, g.x * x.y));
                pr[s + 2] = fmaf(g.z, z.z, fmaf(g.y, y.z, g.x * x.z));
                pr[s + 3] = fmaf(g.z, z.w, fmaf(g.y, y.w, g.x * x.w));
            }
            const float *Qi = cI + NODE_C_Q;                    // [t][h][k] = t*6 + h*3 + k, pre-divided by sdk
            float lq[NH], lp[NH][3];
            {   // scalar keys: eqkm (src/model_operations.py:122)
                float h1[32], h2[32], kq[4];
                first_layer<0, 32>(wE, cI, tJ, g, pr, h1);
                load_bias(wE + (L::E_2QB - L::E_BEGIN), h2);
                gemv_acc<32, 32, 32>(wE + (L::E_2Q - L::E_BEGIN), h1, h2);
                elu_inplace(h2);
                load_bias(wE + (L::E_3QB - L::E_BEGIN), kq);
                gemv_acc<32, 4, 4>(wE + (L::E_3Q - L::E_BEGIN), h2, kq);
#pragma unroll
                for (int h = 0; h < NH; ++h)
                    lq[h] = fmaf(__ldg(Qi + h * 3 + 2), kq[2], fmaf(__ldg(Qi + h * 3 + 1), kq[1], __ldg(Qi + h * 3) * kq[0]));
            }
            {   // vector keys: epkm, chunk gk <-> token group gk (src/model_operations.py:125)
                float h1[32], h2[32], kp[12];
                first_layer<32, 32>(wE, cI, tJ, g, pr, h1);
                load_bias(wE + (L::E_2PB - L::E_BEGIN), h2);
                gemv_acc<32, 32, 32>(wE + (L::E_2P - L::E_BEGIN), h1, h2);
                elu_inplace(h2);
                load_bias(wE + (L::E_3PB - L::E_BEGIN), kp);
                gemv_acc<32, 12, 12>(wE + (L::E_3P - L::E_BEGIN), h2, kp);
#pragma unroll
                for (int h = 0; h < NH; ++h)
#pragma unroll
                    for (int gk = 0; gk < 3; ++gk)
                        lp[h][gk] = fmaf(__ldg(Qi + 6 + h * 3 + 2), kp[gk * 3 + 2],
                                         fmaf(__ldg(Qi + 6 + h * 3 + 1), kp[gk * 3 + 1], __ldg(Qi + 6 + h * 3) * kp[gk * 3]));
            }
            {   // values: evm (src/model_operations.py:128)
                float h1[64], h2[64];
                first_layer<64, 64>(wE, cI, tJ, g, pr, h1);
                load_bias(wE + (L::E_2VB - L::E_BEGIN), h2);
                gemv_acc<64, 64, 64>(wE + (L::E_2V - L::E_BEGIN), h1, h2);
                elu_inplace(h2);
                load_bias(wE + (L::E_3VB - L::E_BEGIN), h1);   // h1 is dead: reuse as V
                gemv_acc<64, 64, 64>(wE + (L::E_3V - L::E_BEGIN), h2, h1);
                float4 *vrow = reinterpret_cast<float4 *>(Vs + tid * VS_STRIDE);
#pragma unroll
                for (int o = 0; o < 64; o += 4) vrow[o / 4] = make_float4(h1[o], h1[o + 1], h1[o + 2], h1[o + 3]);
            }
            // softmax over the atom's nn scalar tokens and 3*nn vector tokens (src/model_operations.py:139-140)
            float mx[4];
            mx[0] = seg_max<SEG>(lq[0]);
            mx[1] = seg_max<SEG>(lq[1]);
            mx[2] = seg_max<SEG>(fmaxf(lp[0][0], fmaxf(lp[0][1], lp[0][2])));
            mx[3] = seg_max<SEG>(fmaxf(lp[1][0], fmaxf(lp[1][1], lp[1][2])));
            if (NN == 64) {
                if (lane == 0) { red[warp * 8 + 0] = mx[0]; red[warp * 8 + 1] = mx[1]; red[warp * 8 + 2] = mx[2]; red[warp * 8 + 3] = mx[3]; }
                __syncthreads();
#pragma unroll
                for (int u = 0; u < 4; ++u) mx[u] = fmaxf(mx[u], red[(warp ^ 1) * 8 + u]);
            }
            float eq[NH], ep[NH][3], sm[4];
#pragma unroll
            for (int h = 0; h < NH; ++h) {
                eq[h] = expf(lq[h] - mx[h]);
#pragma unroll
                for (int gk = 0; gk < 3; ++gk) ep[h][gk] = expf(lp[h][gk] - mx[2 + h]);
            }
            sm[0] = seg_sum<SEG>(eq[0]);
            sm[1] = seg_sum<SEG>(eq[1]);
            sm[2] = seg_sum<SEG>(ep[0][0] + ep[0][1] + ep[0][2]);
            sm[3] = seg_sum<SEG>(ep[1][0] + ep[1][1] + ep[1][2]);
            if (NN == 64) {
                if (lane == 0) { red[warp * 8 + 4] = sm[0]; red[warp * 8 + 5] = sm[1]; red[warp * 8 + 6] = sm[2]; red[warp * 8 + 7] = sm[3]; }
                __syncthreads();
#pragma unroll
                for (int u = 0; u < 4; ++u) sm[u] += red[(warp ^ 1) * 8 + 4 + u];
            }
            float *es = Es + tid * ES_STRIDE;
#pragma unroll
            for (int h = 0; h < NH; ++h) {
                const float iq = 1.0f / sm[h], ip = 1.0f / sm[2 + h];
                es[h] = eq[h] * iq;                            // Mq[h]
                const float w0 = ep[h][0] * ip;                // Mp[h, group 0]: token V1 (x) r
                es[2 + h * 3 + 0] = w0 * g.x;
                es[2 + h * 3 + 1] = w0 * g.y;
                es[2 + h * 3 + 2] = w0 * g.z;
                es[8 + h] = ep[h][2] * ip;                     // group 2: token p_j
                es[11 + h] = ep[h][1] * ip;                    // group 1: token p_i
            }
            es[10] = __int_as_float(j);
        }
        __syncthreads();

        // ------------------------------------------------------------------ phase B: thread = channel
        float zr[TA][NH];
        {
            const int s = lane;
#pragma unroll
            for (int a = 0; a < TA; ++a) {
                float z0 = 0.f, z1 = 0.f;
                const float *es = Es + a * NN * ES_STRIDE;
                const float *vs = Vs + a * NN * VS_STRIDE;
                if (warp == 0) {                               // Zq = Mq . V0   (src/model_operations.py:143)
#pragma unroll 8
                    for (int e = 0; e < NN; ++e) {
                        const float v = vs[e * VS_STRIDE + s];
                        z0 = fmaf(es[e * ES_STRIDE + 0], v, z0);
                        z1 = fmaf(es[e * ES_STRIDE + 1], v, z1);
                    }
                } else {                                       // Zp = Mp . [V1 (x) r ; p_i ; p_j]   (:131-136, :144)
                    const int c = warp - 1;
                    float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
                    for (int e = 0; e < NN; ++e) {
                        const float v = vs[e * VS_STRIDE + 32 + s];
                        const int j = __float_as_int(es[e * ES_STRIDE + 10]);
                        const float pj = __ldg(state_in + (size_t)j * SR + 32 + 32 * c + s);
                        z0 = fmaf(es[e * ES_STRIDE + 2 + c], v, z0);
                        z1 = fmaf(es[e * ES_STRIDE + 5 + c], v, z1);
                        z0 = fmaf(es[e * ES_STRIDE + 8], pj, z0);
                        z1 = fmaf(es[e * ES_STRIDE + 9], pj, z1);
                        s0 += es[e * ES_STRIDE + 11];
                        s1 += es[e * ES_STRIDE + 12];
                    }
                    const int i = min(tile * TA + a, n_atoms - 1);
                    const float pi = __ldg(state_in + (size_t)(i + 1) * SR + 32 + 32 * c + s);
                    z0 = fmaf(s0, pi, z0);
                    z1 = fmaf(s1, pi, z1);
                }
                zr[a][0] = z0;
                zr[a][1] = z1;
            }
        }
        __syncthreads();                                       // all reads of Vs done: reuse it as Zs[TA][256]
        float *Zs = Vs;
#pragma unroll
        for (int a = 0; a < TA; ++a) {
            Zs[a * 256 + warp * 64 + lane] = zr[a][0];         // [warp 0: Zq | warp 1+c: Zp[c]] x [h*32 + s]
            Zs[a * 256 + warp * 64 + 32 + lane] = zr[a][1];
        }
        __syncthreads();

        // ------------------------------------------------------------------ phase C: warp = atom
        for (int a = warp; a < TA; a += EDGE_THREADS / 32) {
            const int i = tile * TA + a;
            if (i >= n_atoms) break;
            const float *z = Zs + a * 256;
            const float *si = state_in + (size_t)(i + 1) * SR;
            float h = __ldg(lw + L::O_Q1B + lane);             // qpm (src/model_operations.py:147)
            float p0 = 0.f, p1 = 0.f, p2 = 0.f;                // ppm (:148)
#pragma unroll 8
            for (int k = 0; k < 64; ++k) {
                h = fmaf(z[k], __ldg(lw + L::O_Q1 + k * 32 + lane), h);
                const float wp = __ldg(lw + L::O_P + k * 32 + lane);
                p0 = fmaf(z[64 + k], wp, p0);
                p1 = fmaf(z[128 + k], wp, p1);
                p2 = fmaf(z[192 + k], wp, p2);
            }
            h = elu(h);
            float g2 = __ldg(lw + L::O_Q2B + lane);
#pragma unroll
            for (int k = 0; k < 32; ++k) g2 = fmaf(__shfl_sync(0xffffffffu, h, k), __ldg(lw + L::O_Q2 + k * 32 + lane), g2);
            g2 = elu(g2);
            float o = __ldg(lw + L::O_Q3B + lane);
#pragma unroll
            for (int k = 0; k < 32; ++k) o = fmaf(__shfl_sync(0xffffffffu, g2, k), __ldg(lw + L::O_Q3 + k * 32 + lane), o);
            float *so = state_out + (size_t)(i + 1) * SR;
            so[lane] = __ldg(si + lane) + o;                   // residual (:151-152)
            so[32 + lane] = __ldg(si + 32 + lane) + p0;
            so[64 + lane] = __ldg(si + 64 + lane) + p1;
            so[96 + lane] = __ldg(si + 96 + lane) + p2;
        }
        __syncthreads();
    }
}

template <int NN>
int launch_edge(const float *lw, int n_atoms, const int32_t *ids32, const float *geom, const float *state_in,
                const float *nodeT, const float *nodeC, float *state_out, cudaStream_t st) {
    int n_sm = 0;
    { const int rc_ = device_setup((const void *)edge_kernel_fp32<NN>, (int)EDGE_SMEM, &n_sm); if (rc_ != PESTO_OK) return rc_; }
    constexpr int TA = EDGE_THREADS / NN;
    int n_tiles = (n_atoms + TA - 1) / TA;
    int grid = n_tiles < 2 * n_sm ? n_tiles : 2 * n_sm;
    edge_kernel_fp32<NN><<<grid, EDGE_THREADS, EDGE_SMEM, st>>>(lw, n_atoms, ids32, (const float4 *)geom, state_in, nodeT,
                                                               nodeC, state_out);
    PESTO_CUDA(cudaGetLastError());
    return PESTO_OK;
}

}  // namespace

int launch_node(const float *lw, int n_atoms, const float *state_in, float *node_scratch, cudaStream_t st) {
    const int n_rows = n_atoms + 1;
    float *nodeT = node_scratch;
    float *nodeC = node_scratch + (size_t)n_rows * NODE_T_STRIDE;
    node_kernel<<<(n_rows + NODE_ATOMS - 1) / NODE_ATOMS, 128, 0, st>>>(lw, state_in, n_rows, nodeT, nodeC);
    PESTO_CUDA(cudaGetLastError());
    return PESTO_OK;
}

// `ev` (optional, 3 events): recorded before the node kernel, between node and edge kernel, after the edge kernel
int launch_state_update_fp32(const float *lw, int nn, int n_atoms, const int32_t *ids32, const float *geom,
                             const float *state_in, float *state_out, float *node_scratch, cudaStream_t st,
                             cudaEvent_t *ev) {
    const int n_rows = n_atoms + 1;
    float *nodeT = node_scratch;
    float *nodeC = node_scratch + (size_t)n_rows * NODE_T_STRIDE;
    if (ev) PESTO_CUDA(cudaEventRecord(ev[0], st));
    int rc = launch_node(lw, n_atoms, state_in, node_scratch, st);
    if (rc != PESTO_OK) return rc;
    if (ev) PESTO_CUDA(cudaEventRecord(ev[1], st));
    switch (nn) {
        case 8:  rc = launch_edge<8>(lw, n_atoms, ids32, geom, state_in, nodeT, nodeC, state_out, st); break;
        case 16: rc = launch_edge<16>(lw, n_atoms, ids32, geom, state_in, nodeT, nodeC, state_out, st); break;
        case 32: rc = launch_edge<32>(lw, n_atoms, ids32, geom, state_in, nodeT, nodeC, state_out, st); break;
        case 64: rc = launch_edge<64>(lw, n_atoms, ids32, geom, state_in, nodeT, nodeC, state_out, st); break;
        default:
            set_error("state_update: unsupported nn=%d (supported: 8, 16, 32, 64)", nn);
            return PESTO_EINVAL;
    }
    if (rc == PESTO_OK && ev) PESTO_CUDA(cudaEventRecord(ev[2], st));
    return rc;
}

}  // namespace pesto
