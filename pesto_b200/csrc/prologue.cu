// Forward prologue: q = em(q0), p = 0 (model/model.py:34-37) and the edge geometry D_nn, R_nn of
// unpack_state_features (src/model_operations.py:6-22), written once into the packed layouts that the 32
// StateUpdate launches re-read:  state float[N+1][128],  ids32 int32[N][64],  geom float4[N][64].
#include "common.cuh"

namespace pesto {

namespace {

// one warp per atom; lane = output unit of the embedding MLP (Linear-ELU-Linear-ELU-Linear, or the single Linear of
// model/save/i_v3_1*/model.py:9-11)
__global__ void __launch_bounds__(256)
embed_kernel(const float *__restrict__ hw, int q0_dim, const float *__restrict__ q0, int n_atoms,
             float *__restrict__ state) {
    int lane = threadIdx.x & 31;
    int row = blockIdx.x * 8 + (threadIdx.x >> 5);   // state row; row 0 = sink
    if (row > n_atoms) return;
    float *out = state + (size_t)row * SR;
    if (row == 0) {
        out[lane] = out[32 + lane] = out[64 + lane] = out[96 + lane] = 0.f;
        return;
    }
    const float *x = q0 + (size_t)(row - 1) * q0_dim;
    float h = hw[HeadLayout::EM_B1 + lane];
    for (int k = 0; k < q0_dim; ++k) h = fmaf(__ldg(x + k), hw[HeadLayout::EM_W1 + k * 32 + lane], h);
    if (hw[HeadLayout::META_EM_LAYERS] < 2.f) {                 // one-layer embedding
        out[lane] = h;
        out[32 + lane] = out[64 + lane] = out[96 + lane] = 0.f;
        return;
    }
    h = elu(h);
    float g = hw[HeadLayout::EM_B2 + lane];
#pragma unroll
    for (int k = 0; k < 32; ++k) g = fmaf(__shfl_sync(0xffffffffu, h, k), hw[HeadLayout::EM_W2 + k * 32 + lane], g);
    g = elu(g);
    float o = hw[HeadLayout::EM_B3 + lane];
#pragma unroll
    for (int k = 0; k < 32; ++k) o = fmaf(__shfl_sync(0xffffffffu, g, k), hw[HeadLayout::EM_W3 + k * 32 + lane], o);
    out[lane] = o;
    out[32 + lane] = out[64 + lane] = out[96 + lane] = 0.f;
}

// pass 1: raw displacement + distance per edge slot, int64 -> int32 ids, global max of D
__global__ void __launch_bounds__(256)
geom_pass1_kernel(const float *__restrict__ X, const int64_t *__restrict__ ids1, int ids_cols, int n_atoms,
                  int32_t *__restrict__ ids32, float4 *__restrict__ geom, unsigned *__restrict__ gmax,
                  int32_t *__restrict__ bad) {
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    float d = 0.f;
    if (e < (size_t)n_atoms * KMAX) {
        int i = (int)(e / KMAX), k = (int)(e % KMAX);
        int id = 0;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < ids_cols) {
            long long idl = ids1[(size_t)i * ids_cols + k];
            if (idl < 0 || idl > n_atoms) { *bad = 1; idl = 0; }
            id = (int)idl;
            int j = id == 0 ? n_atoms - 1 : id - 1;              // X[ids_topk-1]: id 0 wraps to the last atom
            float dx = X[3 * (size_t)j + 0] - X[3 * (size_t)i + 0];
            float dy = X[3 * (size_t)j + 1] - X[3 * (size_t)i + 1];
            float dz = X[3 * (size_t)j + 2] - X[3 * (size_t)i + 2];
            d = dist_exact(dx, dy, dz);
            g = make_float4(dx, dy, dz, d);
        }
        ids32[e] = id;
        geom[e] = g;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) d = fmaxf(d, __shfl_xor_sync(0xffffffffu, d, o));
    if ((threadIdx.x & 31) == 0 && d > 0.f) atomicMax(gmax, __float_as_uint(d));
}

// pass 2: D += max(D) * (D < 1e-2);  R /= D        (src/model_operations.py:12-14)
__global__ void __launch_bounds__(256)
geom_pass2_kernel(float4 *__restrict__ geom, size_t n_slots, int ids_cols, const unsigned *__restrict__ gmax) {
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_slots || (int)(e % KMAX) >= ids_cols) return;
    float maxd = __uint_as_float(*gmax);
    float4 g = geom[e];
    float d = g.w < 1e-2f ? __fadd_rn(g.w, maxd) : g.w;
    geom[e] = make_float4(__fdiv_rn(g.x, d), __fdiv_rn(g.y, d), __fdiv_rn(g.z, d), d);
}

__global__ void unpack_state_kernel(const float *__restrict__ state, size_t n_rows, float *__restrict__ q,
                                    float *__restrict__ p) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_rows * SR) return;
    size_t row = t / SR;
    int c = (int)(t % SR);
    if (c < S) q[row * S + c] = state[t];
    else p[row * 3 * S + (c - S)] = state[t];
}

}  // namespace

int launch_prologue(const float *head_w, int q0_dim, const float *X, const int64_t *ids1, int ids_cols, const float *q0,
                    int n_atoms, float *state, int32_t *ids32, float *geom, void *scratch4, cudaStream_t st) {
    // scratch: [0] = max(D) bits, [1] = bad-index flag
    PESTO_CUDA(cudaMemsetAsync(scratch4, 0, 8, st));
    embed_kernel<<<(n_atoms + 1 + 7) / 8, 256, 0, st>>>(head_w, q0_dim, q0, n_atoms, state);
    size_t n_slots = (size_t)n_atoms * KMAX;
    unsigned blocks = (unsigned)((n_slots + 255) / 256);
    geom_pass1_kernel<<<blocks, 256, 0, st>>>(X, ids1, ids_cols, n_atoms, ids32, (float4 *)geom, (unsigned *)scratch4,
                                              (int32_t *)scratch4 + 1);
    geom_pass2_kernel<<<blocks, 256, 0, st>>>((float4 *)geom, n_slots, ids_cols, (const unsigned *)scratch4);
    PESTO_CUDA(cudaGetLastError());
    return PESTO_OK;
}

int launch_unpack_state(const float *state, int n_atoms, float *q, float *p, cudaStream_t st) {
    size_t n_rows = (size_t)n_atoms + 1;
    unpack_state_kernel<<<(unsigned)((n_rows * SR + 255) / 256), 256, 0, st>>>(state, n_rows, q, p);
    PESTO_CUDA(cudaGetLastError());
    return PESTO_OK;
}

}  // namespace pesto
