// Shared definitions of the pesto_b200 CUDA library: packed-weight layout, device helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pesto_b200.h"

namespace pesto {

constexpr int S = PESTO_NS;             // 32 state channels
constexpr int NH = PESTO_NH;            // 2 heads
constexpr int NK = PESTO_NK;            // 3
constexpr int KMAX = PESTO_MAX_NN;      // 64
constexpr int SR = PESTO_STATE_STRIDE;  // 128 floats per atom record
// The tensor-core path carries edge-MLP pre-activations scaled by log2(e) so that ELU is a bare ex2 (state_update_tc.cu)
constexpr float LOG2E = 1.4426950408889634f;
constexpr float ILOG2E = 0.6931471805599453f;
constexpr float ELU_K = 3.9216517136564484f;          // log2(e) * 2^log2(e) = log2(e) * e (shifted ELU of the tensor-core edge kernel)

// ---------------------------------------------------------------------------------------------
// Per-layer packed weights (float offsets inside one layer block).  "T" = stored transposed,
// i.e. [input k][output o], so that consecutive threads / float4 lanes read consecutive outputs.
//
// Stacked first edge-MLP layer W1 = [eqkm.0 ; epkm.0 ; evm.0] (128 x 193), columns:
//   0: d | 1..32: q_i | 33..64: |p_i| | 65..96: q_j | 97..128: |p_j| | 129..160: p_i.r | 161..192: p_j.r
// is split exactly into per-atom parts (node kernel) and a per-edge part.
// ---------------------------------------------------------------------------------------------
struct LayerLayout {
    // ---- node kernel (per atom) ----
    static constexpr int N_TU  = 0;                   // [64][256]: k = q(32)|pn(32); o = T(128) | U(128)
    static constexpr int N_A   = N_TU + 64 * 256;     // [32][128]: W1[:,129+s] transposed
    static constexpr int N_BU  = N_A + 32 * 128;      // [128] stacked first-layer bias (goes into U)
    static constexpr int NQ_W1 = N_BU + 128;          // nqm.0^T [64][32]
    static constexpr int NQ_B1 = NQ_W1 + 64 * 32;
    static constexpr int NQ_W2 = NQ_B1 + 32;          // nqm.2^T [32][32]
    static constexpr int NQ_B2 = NQ_W2 + 32 * 32;
    static constexpr int NQ_W3 = NQ_B2 + 32;          // nqm.4^T [32][16] (12 used), pre-scaled by 1/sdk
    static constexpr int NQ_B3 = NQ_W3 + 32 * 16;     // [16], pre-scaled by 1/sdk
    // ---- edge kernel (shared-memory resident block, contiguous) ----
    static constexpr int E_BEGIN = NQ_B3 + 16;
    static constexpr int E_WB  = E_BEGIN;             // [32][128]: W1[:,161+s] transposed (p_j.r part)
    static constexpr int E_WD  = E_WB + 32 * 128;     // [128]: W1[:,0] (distance column)
    static constexpr int E_2Q  = E_WD + 128;          // eqkm.2^T [32][32]
    static constexpr int E_2QB = E_2Q + 32 * 32;      // [32]
    static constexpr int E_2P  = E_2QB + 32;          // epkm.2^T [32][32]
    static constexpr int E_2PB = E_2P + 32 * 32;
    static constexpr int E_2V  = E_2PB + 32;          // evm.2^T [64][64]
    static constexpr int E_2VB = E_2V + 64 * 64;      // [64]
    static constexpr int E_3Q  = E_2VB + 64;          // eqkm.4^T [32][4]  (3 used)
    static constexpr int E_3QB = E_3Q + 32 * 4;       // [4]
    static constexpr int E_3P  = E_3QB + 4;           // epkm.4^T [32][12] (9 used)
    static constexpr int E_3PB = E_3P + 32 * 12;      // [12]
    static constexpr int E_3V  = E_3PB + 12;          // evm.4^T [64][64]
    static constexpr int E_3VB = E_3V + 64 * 64;      // [64]
    static constexpr int E_END = E_3VB + 64;
    static constexpr int E_SIZE = E_END - E_BEGIN;    // 15184 floats = 60736 B
    // ---- per-atom output projections ----
    static constexpr int O_Q1  = E_END;               // qpm.0^T [64][32]
    static constexpr int O_Q1B = O_Q1 + 64 * 32;
    static constexpr int O_Q2  = O_Q1B + 32;          // qpm.2^T [32][32]
    static constexpr int O_Q2B = O_Q2 + 32 * 32;
    static constexpr int O_Q3  = O_Q2B + 32;          // qpm.4^T [32][32]
    static constexpr int O_Q3B = O_Q3 + 32 * 32;
    static constexpr int O_P   = O_Q3B + 32;          // ppm.0^T [64][32]
    static constexpr int SIZE  = ((O_P + 64 * 32 + 31) / 32) * 32;
};
static_assert(LayerLayout::E_BEGIN % 4 == 0 && LayerLayout::E_SIZE % 4 == 0, "edge block must be float4 aligned");

// Global (non-layer) weights
struct HeadLayout {
    static constexpr int MAXQ0 = 128;                 // max supported q0 feature width
    static constexpr int EM_W1 = 0;                   // em.0^T [MAXQ0][32]
    static constexpr int EM_B1 = EM_W1 + MAXQ0 * 32;
    static constexpr int EM_W2 = EM_B1 + 32;
    static constexpr int EM_B2 = EM_W2 + 32 * 32;
    static constexpr int EM_W3 = EM_B2 + 32;
    static constexpr int EM_B3 = EM_W3 + 32 * 32;
    static constexpr int SAM_W1 = EM_B3 + 32;         // spl.sam.0^T [64][32]
    static constexpr int SAM_B1 = SAM_W1 + 64 * 32;
    static constexpr int SAM_W2 = SAM_B1 + 32;
    static constexpr int SAM_B2 = SAM_W2 + 32 * 32;
    static constexpr int SAM_W3 = SAM_B2 + 32;        // spl.sam.4^T [32][8]
    static constexpr int SAM_B3 = SAM_W3 + 32 * 8;    // [8]
    static constexpr int ZDM_W1 = SAM_B3 + 8;         // spl.zdm.0^T [128][32]
    static constexpr int ZDM_B1 = ZDM_W1 + 128 * 32;
    static constexpr int ZDM_W2 = ZDM_B1 + 32;
    static constexpr int ZDM_B2 = ZDM_W2 + 32 * 32;
    static constexpr int ZDM_W3 = ZDM_B2 + 32;
    static constexpr int ZDM_B3 = ZDM_W3 + 32 * 32;
    static constexpr int ZDV_W  = ZDM_B3 + 32;        // spl.zdm_vec.0^T [128][32]
    static constexpr int DM_W1  = ZDV_W + 128 * 32;   // dm.0^T [64][32]  (one-layer decoder: dm.0^T [64][8], bias in DM_B1[0..8))
    static constexpr int DM_B1  = DM_W1 + 64 * 32;
    static constexpr int DM_W2  = DM_B1 + 32;
    static constexpr int DM_B2  = DM_W2 + 32 * 32;
    static constexpr int DM_W3  = DM_B2 + 32;         // dm.4^T [32][8] (5 used)
    static constexpr int DM_B3  = DM_W3 + 32 * 8;     // [8]
    // architecture of the two heads (model/save/i_v3_1*/model.py has one Linear where the other checkpoints have
    // Linear-ELU-Linear-ELU-Linear), read by the kernels as warp-uniform values
    static constexpr int META_EM_LAYERS = DM_B3 + 8;  // 1.0f or 3.0f
    static constexpr int META_DM_LAYERS = META_EM_LAYERS + 1;
    static constexpr int META_NUM_OUT   = META_EM_LAYERS + 2;   // logits per residue, 1..8
    static constexpr int SIZE   = META_EM_LAYERS + 8;
};

// node-kernel outputs
constexpr int NODE_T_STRIDE = 128;                    // T_j, gathered per edge (row 0 = sink = 0)
constexpr int NODE_C_STRIDE = 528;                    // U(128) | A_x(128) | A_y(128) | A_z(128) | Q(12)+pad(4)
constexpr int NODE_C_Q = 512;
constexpr int NODE_Z_STRIDE = 256;                    // attention sums of the edge kernel: Zq(2x32) | Zp(3x2x32)
// node scratch of one layer: T | C | Z, each with n_atoms + 1 rows (row 0 = sink)
__host__ __device__ inline size_t node_scratch_floats(int n_atoms) {
    return ((size_t)n_atoms + 1) * (NODE_T_STRIDE + NODE_C_STRIDE + NODE_Z_STRIDE);
}
inline float *node_Z(float *node_scratch, int n_atoms) { return node_scratch + ((size_t)n_atoms + 1) * (NODE_T_STRIDE + NODE_C_STRIDE); }

__device__ __forceinline__ float elu(float x) { return x > 0.f ? x : (expf(x) - 1.0f); }

__device__ __forceinline__ float dist_exact(float dx, float dy, float dz) {
    // bit-exact with torch.norm(R, dim=-1) on CPU for 3-vectors
    return __fsqrt_rn(__fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx))));
}

void set_error(const char *fmt, ...);
int check_cuda(cudaError_t e, const char *what);
// Per-device launch setup of a kernel: opt in to `dyn_smem` bytes of dynamic shared memory (a per-device function
// attribute) the first time the kernel is launched on the current device, and return that device's SM count.
int device_setup(const void *kernel, int dyn_smem, int *n_sm_out);
// Device word (one per device, allocated on first use) that receives the stage id of a tensor-core wait that timed out
// when a launch is not given a status word of its own (staged API); nullptr if it cannot be allocated.
int *device_watchdog_word();

#define PESTO_CUDA(call)                                             \
    do {                                                             \
        int _rc = ::pesto::check_cuda((call), #call);                \
        if (_rc != PESTO_OK) return _rc;                             \
    } while (0)

// Launch `kernel` on `st` with programmatic dependent launch allowed: its CTAs may start while the previous kernel of the
// stream is finishing and run until their griddepcontrol.wait (tc::pdl_wait); everything before that point must depend on
// the model only.  Kernels that never execute pdl_wait / pdl_launch_dependents behave as under a normal launch.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// launch wrappers implemented in the individual .cu files
int launch_knn(const float *X, int n_atoms, const int32_t *seg_off, int n_seg, int k, int base,
               int64_t *ids_out, float *d_out, float *r_out, void *scratch, cudaStream_t st);
int launch_prologue(const float *head_w, int q0_dim, const float *X, const int64_t *ids1, int ids_cols,
                    const float *q0, int n_atoms, float *state, int32_t *ids32, float *geom,
                    void *scratch4, cudaStream_t st);
int launch_state_update_fp32(const float *layer_w, int nn, int n_atoms, const int32_t *ids32, const float *geom,
                             const float *state_in, float *state_out, float *node_scratch, cudaStream_t st,
                             cudaEvent_t *ev = nullptr);
int launch_node(const float *layer_w, int n_atoms, const float *state_in, float *node_scratch, cudaStream_t st);
int launch_edge_tc_layer(const float *lw, const void *tcw, int nn, int n_atoms, const int32_t *ids32, const float *geom,
                         const float *state_in, float *node_scratch, float *Z, int mode, cudaStream_t st, int *wd = nullptr);
int launch_state_update_tc(const float *lw, const void *tcw, int nn, int n_atoms, const int32_t *ids32, const float *geom,
                           const float *state_in, float *state_out, float *node_scratch, float *Z, int mode,
                           cudaStream_t st, cudaEvent_t *ev, int *wd = nullptr);
int launch_node_umma(const void *img_tail, const void *img_head, const float *state_prev, const float *Z, float *state_new,
                     int n_atoms, float *node_scratch, int mode, cudaStream_t st, int *wd = nullptr);
size_t node_tc_layer_bytes();
void pack_node_tc_layer(const float *layer_blob_host, void *dst_host);
size_t tc_edge_bytes();          // the per-layer image block is [edge images | node images]
size_t tc_layer_bytes();
void pack_tc_layer(const float *layer_blob_host, void *dst_host);
int launch_residue_index(const float *M, int n_atoms, int n_res, int32_t *rid, int32_t *flags, cudaStream_t st);
int launch_pool_decode(const float *head_w, const float *state, const int32_t *rid, int n_atoms, int n_res,
                       float *z, void *scratch, int32_t *poison, cudaStream_t st);      // poison: status words 1..4 of the forward (or nullptr)
int launch_unpack_state(const float *state, int n_atoms, float *q, float *p, cudaStream_t st);
size_t pool_scratch_bytes(int n_atoms, int n_res);
size_t knn_scratch_bytes(int n_atoms, int n_seg);

}  // namespace pesto
