// extern "C" entry points of libpesto_b200.so (declared in include/pesto_b200.h) and the host-side
// weight packer that turns the reference's state-dict tensors into the device layout of common.cuh.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "common.cuh"

namespace pesto {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_cuda(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return PESTO_OK;
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return PESTO_ECUDA;
}

int device_setup(const void *kernel, int dyn_smem, int *n_sm_out) {
    // cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute of a kernel, and so is the SM count the
    // persistent grids are sized with: both are cached per (kernel, device), not per process
    static std::mutex mu;
    static std::map<std::pair<const void *, int>, int> done;       // -> SM count
    int dev = 0;
    PESTO_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    auto it = done.find({kernel, dev});
    if (it == done.end()) {
        if (dyn_smem > 0) PESTO_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_smem));
        int n_sm = 0;
        PESTO_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        it = done.emplace(std::make_pair(kernel, dev), n_sm).first;
    }
    *n_sm_out = it->second;
    return PESTO_OK;
}

int *device_watchdog_word() {
    static std::mutex mu;
    static std::map<int, int *> words;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    auto it = words.find(dev);
    if (it == words.end()) {
        int *p = nullptr;
        if (cudaMalloc((void **)&p, sizeof(int)) != cudaSuccess || cudaMemset(p, 0, sizeof(int)) != cudaSuccess) return nullptr;
        it = words.emplace(dev, p).first;
    }
    return it->second;
}

}  // namespace pesto

using namespace pesto;

struct pesto_model {
    int n_layers = 0;
    int q0_dim = 0;
    int em_layers = 3, dm_layers = 3, n_out = PESTO_NUM_OUT;   // found from the tensors at finalize
    std::vector<int> nn;
    std::map<std::string, std::vector<float>> tensors;
    float *d_blob = nullptr;       // [n_layers * LayerLayout::SIZE | HeadLayout::SIZE]
    unsigned char *d_tc = nullptr; // tensor-core operand images, one block per layer
    bool finalized = false;
    const float *layer(int l) const { return d_blob + (size_t)l * LayerLayout::SIZE; }
    const float *head() const { return d_blob + (size_t)n_layers * LayerLayout::SIZE; }
    const void *layer_tc(int l) const { return d_tc ? d_tc + (size_t)l * tc_layer_bytes() : nullptr; }
};

namespace {

struct Packer {
    const pesto_model *m;
    bool ok = true;
    const float *get(const std::string &key, size_t numel) {
        auto it = m->tensors.find(key);
        if (it == m->tensors.end()) {
            if (ok) set_error("model_finalize: missing tensor '%s'", key.c_str());
            ok = false;
            return nullptr;
        }
        if (it->second.size() != numel) {
            if (ok) set_error("model_finalize: tensor '%s' has %zu elements, expected %zu", key.c_str(), it->second.size(), numel);
            ok = false;
            return nullptr;
        }
        return it->second.data();
    }
    // dst[k * ld + o] = W[o][k] * scale      (W row-major [n_out][n_in])
    void transposed(float *dst, int ld, const std::string &key, int n_out, int n_in, float scale = 1.f) {
        const float *w = get(key, (size_t)n_out * n_in);
        if (!w) return;
        for (int o = 0; o < n_out; ++o)
            for (int k = 0; k < n_in; ++k) dst[(size_t)k * ld + o] = w[(size_t)o * n_in + k] * scale;
    }
    void vec(float *dst, const std::string &key, int n, float scale = 1.f) {
        const float *b = get(key, (size_t)n);
        if (!b) return;
        for (int i = 0; i < n; ++i) dst[i] = b[i] * scale;
    }
};

void pack_layer(Packer &pk, int l, float *dst) {
    using L = LayerLayout;
    const std::string p = "sum." + std::to_string(l) + ".su.";
    // stacked first layer [eqkm.0 ; epkm.0 ; evm.0]: 128 x 193
    std::vector<float> W1(128 * 193, 0.f), b1(128, 0.f);
    const char *names[3] = {"eqkm", "epkm", "evm"};
    const int rows[3] = {32, 32, 64}, row0[3] = {0, 32, 64};
    for (int i = 0; i < 3; ++i) {
        const float *w = pk.get(p + names[i] + ".0.weight", (size_t)rows[i] * 193);
        const float *b = pk.get(p + names[i] + ".0.bias", (size_t)rows[i]);
        if (!w || !b) return;
        memcpy(&W1[(size_t)row0[i] * 193], w, sizeof(float) * rows[i] * 193);
        memcpy(&b1[row0[i]], b, sizeof(float) * rows[i]);
    }
    for (int o = 0; o < 128; ++o) {
        const float *w = &W1[(size_t)o * 193];
        dst[L::E_WD + o] = w[0];
        dst[L::N_BU + o] = b1[o];
        for (int s = 0; s < 32; ++s) {
            dst[L::N_TU + s * 256 + 128 + o] = w[1 + s];            // U: q_i
            dst[L::N_TU + (32 + s) * 256 + 128 + o] = w[33 + s];    // U: |p_i|
            dst[L::N_TU + s * 256 + o] = w[65 + s];                 // T: q_j
            dst[L::N_TU + (32 + s) * 256 + o] = w[97 + s];          // T: |p_j|
            dst[L::N_A + s * 128 + o] = w[129 + s];                 // A: p_i . r
            dst[L::E_WB + s * 128 + o] = w[161 + s];                // p_j . r
        }
    }
    float sdk = std::sqrt((float)NK);
    auto it = pk.m->tensors.find(p + "sdk");
    if (it != pk.m->tensors.end() && it->second.size() == 1) sdk = it->second[0];
    const float isdk = 1.0f / sdk;
    pk.transposed(dst + L::NQ_W1, 32, p + "nqm.0.weight", 32, 64);
    pk.vec(dst + L::NQ_B1, p + "nqm.0.bias", 32);
    pk.transposed(dst + L::NQ_W2, 32, p + "nqm.2.weight", 32, 32);
    pk.vec(dst + L::NQ_B2, p + "nqm.2.bias", 32);
    pk.transposed(dst + L::NQ_W3, 16, p + "nqm.4.weight", 12, 32, isdk);
    pk.vec(dst + L::NQ_B3, p + "nqm.4.bias", 12, isdk);
    pk.transposed(dst + L::E_2Q, 32, p + "eqkm.2.weight", 32, 32);
    pk.vec(dst + L::E_2QB, p + "eqkm.2.bias", 32);
    pk.transposed(dst + L::E_2P, 32, p + "epkm.2.weight", 32, 32);
    pk.vec(dst + L::E_2PB, p + "epkm.2.bias", 32);
    pk.transposed(dst + L::E_2V, 64, p + "evm.2.weight", 64, 64);
    pk.vec(dst + L::E_2VB, p + "evm.2.bias", 64);
    pk.transposed(dst + L::E_3Q, 4, p + "eqkm.4.weight", 3, 32);
    pk.vec(dst + L::E_3QB, p + "eqkm.4.bias", 3);
    pk.transposed(dst + L::E_3P, 12, p + "epkm.4.weight", 9, 32);
    pk.vec(dst + L::E_3PB, p + "epkm.4.bias", 9);
    pk.transposed(dst + L::E_3V, 64, p + "evm.4.weight", 64, 64);
    pk.vec(dst + L::E_3VB, p + "evm.4.bias", 64);
    pk.transposed(dst + L::O_Q1, 32, p + "qpm.0.weight", 32, 64);
    pk.vec(dst + L::O_Q1B, p + "qpm.0.bias", 32);
    pk.transposed(dst + L::O_Q2, 32, p + "qpm.2.weight", 32, 32);
    pk.vec(dst + L::O_Q2B, p + "qpm.2.bias", 32);
    pk.transposed(dst + L::O_Q3, 32, p + "qpm.4.weight", 32, 32);
    pk.vec(dst + L::O_Q3B, p + "qpm.4.bias", 32);
    pk.transposed(dst + L::O_P, 32, p + "ppm.0.weight", 32, 64);
}

// The depth of em / dm and the number of logits are read off the checkpoint: `em.2.weight` / `dm.2.weight` exist only in
// the three-layer heads (model/save/i_v3_1*/model.py:9-13,20-22 has single Linear layers), the last dm bias has N2 entries.
bool head_architecture(pesto_model *m) {
    m->em_layers = m->tensors.count("em.2.weight") ? 3 : 1;
    m->dm_layers = m->tensors.count("dm.2.weight") ? 3 : 1;
    auto it = m->tensors.find(m->dm_layers == 3 ? "dm.4.bias" : "dm.0.bias");
    if (it == m->tensors.end() || it->second.empty() || it->second.size() > 8) {
        set_error("model_finalize: the decoder's output bias (dm.%d.bias) is missing or has more than 8 entries", m->dm_layers == 3 ? 4 : 0);
        return false;
    }
    m->n_out = (int)it->second.size();
    return true;
}

void pack_head(Packer &pk, float *dst) {
    using H = HeadLayout;
    const int q0 = pk.m->q0_dim, n_out = pk.m->n_out;
    dst[H::META_EM_LAYERS] = (float)pk.m->em_layers;
    dst[H::META_DM_LAYERS] = (float)pk.m->dm_layers;
    dst[H::META_NUM_OUT] = (float)n_out;
    pk.transposed(dst + H::EM_W1, 32, "em.0.weight", 32, q0);
    pk.vec(dst + H::EM_B1, "em.0.bias", 32);
    if (pk.m->em_layers == 3) {
        pk.transposed(dst + H::EM_W2, 32, "em.2.weight", 32, 32);
        pk.vec(dst + H::EM_B2, "em.2.bias", 32);
        pk.transposed(dst + H::EM_W3, 32, "em.4.weight", 32, 32);
        pk.vec(dst + H::EM_B3, "em.4.bias", 32);
    }
    pk.transposed(dst + H::SAM_W1, 32, "spl.sam.0.weight", 32, 64);
    pk.vec(dst + H::SAM_B1, "spl.sam.0.bias", 32);
    pk.transposed(dst + H::SAM_W2, 32, "spl.sam.2.weight", 32, 32);
    pk.vec(dst + H::SAM_B2, "spl.sam.2.bias", 32);
    pk.transposed(dst + H::SAM_W3, 8, "spl.sam.4.weight", 8, 32);
    pk.vec(dst + H::SAM_B3, "spl.sam.4.bias", 8);
    pk.transposed(dst + H::ZDM_W1, 32, "spl.zdm.0.weight", 32, 128);
    pk.vec(dst + H::ZDM_B1, "spl.zdm.0.bias", 32);
    pk.transposed(dst + H::ZDM_W2, 32, "spl.zdm.2.weight", 32, 32);
    pk.vec(dst + H::ZDM_B2, "spl.zdm.2.bias", 32);
    pk.transposed(dst + H::ZDM_W3, 32, "spl.zdm.4.weight", 32, 32);
    pk.vec(dst + H::ZDM_B3, "spl.zdm.4.bias", 32);
    pk.transposed(dst + H::ZDV_W, 32, "spl.zdm_vec.0.weight", 32, 128);
    if (pk.m->dm_layers == 3) {
        pk.transposed(dst + H::DM_W1, 32, "dm.0.weight", 32, 64);
        pk.vec(dst + H::DM_B1, "dm.0.bias", 32);
        pk.transposed(dst + H::DM_W2, 32, "dm.2.weight", 32, 32);
        pk.vec(dst + H::DM_B2, "dm.2.bias", 32);
        pk.transposed(dst + H::DM_W3, 8, "dm.4.weight", n_out, 32);
        pk.vec(dst + H::DM_B3, "dm.4.bias", n_out);
    } else {
        pk.transposed(dst + H::DM_W1, 8, "dm.0.weight", n_out, 64);
        pk.vec(dst + H::DM_B1, "dm.0.bias", n_out);
    }
}

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

struct Workspace {
    float *state_a, *state_b, *geom, *node;
    int32_t *ids32, *rid, *status;
    void *pool;
    size_t total;
};

Workspace carve_workspace(void *base, int n_atoms, int n_res) {
    Workspace w;
    size_t o = 0;
    auto take = [&](size_t bytes) {
        void *p = base ? (char *)base + o : nullptr;
        o += align_up(bytes);
        return p;
    };
    const size_t rows = (size_t)n_atoms + 1;
    w.state_a = (float *)take(rows * SR * sizeof(float));
    w.state_b = (float *)take(rows * SR * sizeof(float));
    w.ids32 = (int32_t *)take((size_t)n_atoms * KMAX * sizeof(int32_t));
    w.geom = (float *)take((size_t)n_atoms * KMAX * 4 * sizeof(float));
    w.node = (float *)take(node_scratch_floats(n_atoms) * sizeof(float));
    w.rid = (int32_t *)take((size_t)n_atoms * sizeof(int32_t));
    w.status = (int32_t *)take(PESTO_STATUS_WORDS * sizeof(int32_t));
    w.pool = take(pool_scratch_bytes(n_atoms, n_res));
    w.total = o;
    return w;
}

bool valid_model(const pesto_model *m, const char *who) {
    if (!m || !m->finalized) {
        set_error("%s: model is NULL or not finalized", who);
        return false;
    }
    return true;
}

}  // namespace

extern "C" {

int pesto_abi_version(void) { return 1; }
const char *pesto_last_error(void) { return g_err; }

size_t pesto_knn_scratch_bytes(int n_atoms, int n_seg) { return knn_scratch_bytes(n_atoms, n_seg); }

int pesto_knn(const float *X, int n_atoms, const int32_t *seg_off, int n_seg, int k, int base, int64_t *ids_out,
              float *d_out, float *r_out, void *scratch, void *stream) {
    if (!X || !seg_off || !ids_out || !scratch || n_atoms < 0 || n_seg < 1) {
        set_error("pesto_knn: null pointer or bad sizes (n_atoms=%d, n_seg=%d)", n_atoms, n_seg);
        return PESTO_EINVAL;
    }
    if (k < 1 || k > KMAX) {
        set_error("pesto_knn: k=%d outside [1, %d]", k, KMAX);
        return PESTO_EINVAL;
    }
    return launch_knn(X, n_atoms, seg_off, n_seg, k, base ? 1 : 0, ids_out, d_out, r_out, scratch, (cudaStream_t)stream);
}
int pesto_knn_launch_count(void) { return 4; }

pesto_model_t *pesto_model_create(int n_layers, const int32_t *nn_per_layer_host, int q0_dim) {
    if (n_layers < 1 || !nn_per_layer_host || q0_dim < 1 || q0_dim > HeadLayout::MAXQ0) {
        set_error("pesto_model_create: bad arguments (n_layers=%d, q0_dim=%d, max q0_dim %d)", n_layers, q0_dim, HeadLayout::MAXQ0);
        return nullptr;
    }
    for (int l = 0; l < n_layers; ++l) {
        int nn = nn_per_layer_host[l];
        if (nn != 8 && nn != 16 && nn != 32 && nn != 64) {
            set_error("pesto_model_create: layer %d has nn=%d; supported: 8, 16, 32, 64", l, nn);
            return nullptr;
        }
    }
    pesto_model *m = new pesto_model();
    m->n_layers = n_layers;
    m->q0_dim = q0_dim;
    m->nn.assign(nn_per_layer_host, nn_per_layer_host + n_layers);
    return m;
}

int pesto_model_set_tensor(pesto_model_t *m, const char *key, const float *data_host, int64_t numel) {
    if (!m || !key || !data_host || numel < 0) {
        set_error("pesto_model_set_tensor: null argument");
        return PESTO_EINVAL;
    }
    if (m->finalized) {
        set_error("pesto_model_set_tensor: model already finalized");
        return PESTO_ESTATE;
    }
    m->tensors[key].assign(data_host, data_host + numel);
    return PESTO_OK;
}

int pesto_model_finalize(pesto_model_t *m) {
    if (!m) {
        set_error("pesto_model_finalize: null model");
        return PESTO_EINVAL;
    }
    if (m->finalized) return PESTO_OK;
    if (!head_architecture(m)) return PESTO_ESTATE;
    const size_t n_float = (size_t)m->n_layers * LayerLayout::SIZE + HeadLayout::SIZE;
    std::vector<float> blob(n_float, 0.f);
    Packer pk{m};
    for (int l = 0; l < m->n_layers && pk.ok; ++l) pack_layer(pk, l, blob.data() + (size_t)l * LayerLayout::SIZE);
    if (pk.ok) pack_head(pk, blob.data() + (size_t)m->n_layers * LayerLayout::SIZE);
    if (!pk.ok) return PESTO_ESTATE;
    PESTO_CUDA(cudaMalloc((void **)&m->d_blob, n_float * sizeof(float)));
    PESTO_CUDA(cudaMemcpy(m->d_blob, blob.data(), n_float * sizeof(float), cudaMemcpyHostToDevice));
    const size_t tcb = tc_layer_bytes();
    if (tcb) {
        std::vector<unsigned char> tc((size_t)m->n_layers * tcb, 0);
        for (int l = 0; l < m->n_layers; ++l)
            pack_tc_layer(blob.data() + (size_t)l * LayerLayout::SIZE, tc.data() + (size_t)l * tcb);
        PESTO_CUDA(cudaMalloc((void **)&m->d_tc, tc.size()));
        PESTO_CUDA(cudaMemcpy(m->d_tc, tc.data(), tc.size(), cudaMemcpyHostToDevice));
    }
    m->tensors.clear();
    m->finalized = true;
    return PESTO_OK;
}

void pesto_model_destroy(pesto_model_t *m) {
    if (!m) return;
    if (m->d_blob) cudaFree(m->d_blob);
    if (m->d_tc) cudaFree(m->d_tc);
    delete m;
}

int pesto_model_num_layers(const pesto_model_t *m) { return m ? m->n_layers : 0; }
int pesto_model_num_out(const pesto_model_t *m) { return (m && m->finalized) ? m->n_out : 0; }
int pesto_model_layer_nn(const pesto_model_t *m, int layer) {
    return (m && layer >= 0 && layer < m->n_layers) ? m->nn[layer] : 0;
}

int pesto_prologue(const pesto_model_t *m, const float *X, const int64_t *ids1, int ids_cols, const float *q0,
                   int n_atoms, float *state, int32_t *ids32, float *geom, void *scratch8, void *stream) {
    if (!valid_model(m, "pesto_prologue")) return PESTO_ESTATE;
    if (!X || !ids1 || !q0 || !state || !ids32 || !geom || !scratch8 || n_atoms < 1 || ids_cols < 1 || ids_cols > KMAX) {
        set_error("pesto_prologue: null pointer or bad sizes (n_atoms=%d, ids_cols=%d)", n_atoms, ids_cols);
        return PESTO_EINVAL;
    }
    return launch_prologue(m->head(), m->q0_dim, X, ids1, ids_cols, q0, n_atoms, state, ids32, geom, scratch8,
                           (cudaStream_t)stream);
}

size_t pesto_node_scratch_bytes(int n_atoms) {   // per-atom factors T | C, the attention sums Z and the fp16 p planes of one layer
    return node_scratch_floats(n_atoms) * sizeof(float);
}

static int state_update_impl(const pesto_model_t *m, int layer, int n_atoms, const int32_t *ids32, const float *geom,
                             const float *state_in, float *state_out, void *node_scratch, int mode, void *stream,
                             cudaEvent_t *ev) {
    if (!valid_model(m, "pesto_state_update")) return PESTO_ESTATE;
    if (layer < 0 || layer >= m->n_layers || n_atoms < 1 || !ids32 || !geom || !state_in || !state_out || !node_scratch ||
        state_in == state_out) {
        set_error("pesto_state_update: bad arguments (layer=%d of %d, n_atoms=%d)", layer, m->n_layers, n_atoms);
        return PESTO_EINVAL;
    }
    if (mode == PESTO_MODE_FP32)
        return launch_state_update_fp32(m->layer(layer), m->nn[layer], n_atoms, ids32, geom, state_in, state_out,
                                        (float *)node_scratch, (cudaStream_t)stream, ev);
    if (mode == PESTO_MODE_BF16X3 || mode == PESTO_MODE_BF16)
        return launch_state_update_tc(m->layer(layer), m->layer_tc(layer), m->nn[layer], n_atoms, ids32, geom, state_in,
                                      state_out, (float *)node_scratch, node_Z((float *)node_scratch, n_atoms), mode,
                                      (cudaStream_t)stream, ev);
    set_error("pesto_state_update: unknown mode %d", mode);
    return PESTO_EINVAL;
}

int pesto_state_update(const pesto_model_t *m, int layer, int n_atoms, const int32_t *ids32, const float *geom,
                       const float *state_in, float *state_out, void *node_scratch, int mode, void *stream) {
    return state_update_impl(m, layer, n_atoms, ids32, geom, state_in, state_out, node_scratch, mode, stream, nullptr);
}

int pesto_state_update_timed(const pesto_model_t *m, int layer, int n_atoms, const int32_t *ids32, const float *geom,
                             const float *state_in, float *state_out, void *node_scratch, int mode, void *stream,
                             float *ms_node_host, float *ms_edge_host) {
    cudaEvent_t ev[3];
    for (int i = 0; i < 3; ++i) PESTO_CUDA(cudaEventCreate(&ev[i]));
    int rc = state_update_impl(m, layer, n_atoms, ids32, geom, state_in, state_out, node_scratch, mode, stream, ev);
    if (rc == PESTO_OK) rc = check_cuda(cudaEventSynchronize(ev[2]), "cudaEventSynchronize");
    float a = 0.f, b = 0.f;
    if (rc == PESTO_OK) rc = check_cuda(cudaEventElapsedTime(&a, ev[0], ev[1]), "cudaEventElapsedTime");
    if (rc == PESTO_OK) rc = check_cuda(cudaEventElapsedTime(&b, ev[1], ev[2]), "cudaEventElapsedTime");
    for (int i = 0; i < 3; ++i) cudaEventDestroy(ev[i]);
    if (ms_node_host) *ms_node_host = a;
    if (ms_edge_host) *ms_edge_host = b;
    return rc;
}

int pesto_edge_kernel_timed(const pesto_model_t *m, int layer, int n_atoms, const int32_t *ids32, const float *geom,
                            const float *state_in, void *node_scratch, int mode, int reps, void *stream,
                            float *ms_per_launch_host) {
    if (!valid_model(m, "pesto_edge_kernel_timed")) return PESTO_ESTATE;
    if (layer < 0 || layer >= m->n_layers || n_atoms < 1 || !ids32 || !geom || !state_in || !node_scratch || reps < 1 ||
        !ms_per_launch_host || (mode != PESTO_MODE_F16X3 && mode != PESTO_MODE_F16)) {
        set_error("pesto_edge_kernel_timed: bad arguments (tensor-core modes only)");
        return PESTO_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    float *ns = (float *)node_scratch;
    const void *nimg = (const void *)((const unsigned char *)m->layer_tc(layer) + tc_edge_bytes());
    int rc = launch_node_umma(nullptr, nimg, state_in, nullptr, nullptr, n_atoms, ns, mode, st);      // the layer's per-atom factors
    if (rc != PESTO_OK) return rc;
    cudaEvent_t ev[2];
    for (int i = 0; i < 2; ++i) PESTO_CUDA(cudaEventCreate(&ev[i]));
    rc = launch_edge_tc_layer(m->layer(layer), m->layer_tc(layer), m->nn[layer], n_atoms, ids32, geom, state_in, ns, node_Z(ns, n_atoms), mode, st);
    if (rc == PESTO_OK) rc = check_cuda(cudaEventRecord(ev[0], st), "cudaEventRecord");
    for (int r = 0; r < reps && rc == PESTO_OK; ++r)
        rc = launch_edge_tc_layer(m->layer(layer), m->layer_tc(layer), m->nn[layer], n_atoms, ids32, geom, state_in, ns, node_Z(ns, n_atoms), mode, st);
    if (rc == PESTO_OK) rc = check_cuda(cudaEventRecord(ev[1], st), "cudaEventRecord");
    if (rc == PESTO_OK) rc = check_cuda(cudaEventSynchronize(ev[1]), "cudaEventSynchronize");
    float ms = 0.f;
    if (rc == PESTO_OK) rc = check_cuda(cudaEventElapsedTime(&ms, ev[0], ev[1]), "cudaEventElapsedTime");
    for (int i = 0; i < 2; ++i) cudaEventDestroy(ev[i]);
    *ms_per_launch_host = ms / (float)reps;
    return rc;
}

int pesto_residue_index(const float *M, int n_atoms, int n_res, int32_t *rid, int32_t *flags, void *stream) {
    if (!M || !rid || !flags || n_atoms < 1 || n_res < 1) {
        set_error("pesto_residue_index: null pointer or bad sizes");
        return PESTO_EINVAL;
    }
    return launch_residue_index(M, n_atoms, n_res, rid, flags, (cudaStream_t)stream);
}

size_t pesto_pool_scratch_bytes(int n_atoms, int n_res) { return pool_scratch_bytes(n_atoms, n_res); }

int pesto_pool_decode(const pesto_model_t *m, const float *state, const int32_t *rid, int n_atoms, int n_res, float *z,
                      void *scratch, void *stream) {
    if (!valid_model(m, "pesto_pool_decode")) return PESTO_ESTATE;
    if (!state || !rid || !z || !scratch || n_atoms < 1 || n_res < 1) {
        set_error("pesto_pool_decode: null pointer or bad sizes");
        return PESTO_EINVAL;
    }
    return launch_pool_decode(m->head(), state, rid, n_atoms, n_res, z, scratch, nullptr, (cudaStream_t)stream);
}

int pesto_unpack_state(const float *state, int n_atoms, float *q, float *p, void *stream) {
    if (!state || !q || !p || n_atoms < 0) {
        set_error("pesto_unpack_state: null pointer");
        return PESTO_EINVAL;
    }
    return launch_unpack_state(state, n_atoms, q, p, (cudaStream_t)stream);
}

size_t pesto_forward_workspace_bytes(int n_atoms, int n_res) {
    if (n_atoms < 1 || n_res < 1) return 0;
    return carve_workspace(nullptr, n_atoms, n_res).total;
}

int pesto_forward(const pesto_model_t *m, const float *X, const int64_t *ids1, int ids_cols, const float *q0,
                  const float *M, const int32_t *rid, int n_atoms, int n_res, float *z, void *workspace,
                  size_t workspace_bytes, int mode, void *stream) {
    if (!valid_model(m, "pesto_forward")) return PESTO_ESTATE;
    if (!X || !ids1 || !q0 || !z || !workspace || n_atoms < 1 || n_res < 1 || ((M == nullptr) == (rid == nullptr))) {
        set_error("pesto_forward: null pointer, bad sizes, or not exactly one of M / rid given");
        return PESTO_EINVAL;
    }
    if (ids_cols < 1 || ids_cols > KMAX) {
        set_error("pesto_forward: ids_topk has %d columns, supported 1..%d", ids_cols, KMAX);
        return PESTO_EINVAL;
    }
    for (int l = 0; l < m->n_layers; ++l)
        if (m->nn[l] > ids_cols) {
            set_error("pesto_forward: layer %d needs nn=%d neighbours but ids_topk has %d columns", l, m->nn[l], ids_cols);
            return PESTO_EINVAL;
        }
    if (((uintptr_t)workspace & 255) != 0) {
        set_error("pesto_forward: workspace must be 256-byte aligned");
        return PESTO_EINVAL;
    }
    Workspace w = carve_workspace(workspace, n_atoms, n_res);
    if (workspace_bytes < w.total) {
        set_error("pesto_forward: workspace too small (%zu < %zu bytes)", workspace_bytes, w.total);
        return PESTO_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    PESTO_CUDA(cudaMemsetAsync(w.status, 0, PESTO_STATUS_WORDS * sizeof(int32_t), st));
    int *wd = w.status + 3;      // a tensor-core stage that never completes records its id here -> NaN logits, pesto_forward_status
    int rc = launch_prologue(m->head(), m->q0_dim, X, ids1, ids_cols, q0, n_atoms, w.state_a, w.ids32, w.geom, w.status, st);
    if (rc != PESTO_OK) return rc;
    float *cur = w.state_a, *nxt = w.state_b;
    if (mode == PESTO_MODE_FP32) {
        for (int l = 0; l < m->n_layers; ++l) {
            rc = pesto_state_update(m, l, n_atoms, w.ids32, w.geom, cur, nxt, w.node, mode, stream);
            if (rc != PESTO_OK) return rc;
            float *t = cur; cur = nxt; nxt = t;
        }
    } else if (mode == PESTO_MODE_BF16X3 || mode == PESTO_MODE_BF16) {
        // tensor-core path: the per-atom tail of layer l and the per-atom head of layer l+1 share one launch
        float *Z = node_Z(w.node, n_atoms);
        auto node_img = [&](int l) { return (const void *)((const unsigned char *)m->layer_tc(l) + tc_edge_bytes()); };
        rc = launch_node_umma(nullptr, node_img(0), cur, nullptr, nullptr, n_atoms, w.node, mode, st, wd);
        if (rc != PESTO_OK) return rc;
        for (int l = 0; l < m->n_layers; ++l) {
            rc = launch_edge_tc_layer(m->layer(l), m->layer_tc(l), m->nn[l], n_atoms, w.ids32, w.geom, cur, w.node, Z, mode, st, wd);
            if (rc != PESTO_OK) return rc;
            const bool more = l + 1 < m->n_layers;
            rc = launch_node_umma(node_img(l), more ? node_img(l + 1) : nullptr, cur, Z, nxt, n_atoms, w.node, mode, st, wd);
            if (rc != PESTO_OK) return rc;
            float *t = cur; cur = nxt; nxt = t;
        }
    } else {
        set_error("pesto_forward: unknown mode %d", mode);
        return PESTO_EINVAL;
    }
    const int32_t *rid_dev = rid;
    if (M) {
        rc = launch_residue_index(M, n_atoms, n_res, w.rid, w.status + 2, st);
        if (rc != PESTO_OK) return rc;
        rid_dev = w.rid;
    }
    // status[1] = id out of range, [2] = membership not one-hot, [3] = a tensor-core stage timed out: each poisons z with NaN
    // ([4] = residue index out of range, set by the pool kernels)
    return launch_pool_decode(m->head(), cur, rid_dev, n_atoms, n_res, z, w.pool, w.status + 1, st);
}

size_t pesto_forward_status_offset(int n_atoms, int n_res) {
    if (n_atoms < 1 || n_res < 1) return 0;
    char *fake = reinterpret_cast<char *>(uintptr_t(1) << 20);
    return (size_t)(reinterpret_cast<char *>(carve_workspace(fake, n_atoms, n_res).status) - fake);
}

int pesto_forward_status(const void *workspace, int n_atoms, int n_res, int32_t *status_host, void *stream) {
    if (!workspace || !status_host || n_atoms < 1 || n_res < 1) {
        set_error("pesto_forward_status: null pointer or bad sizes");
        return PESTO_EINVAL;
    }
    Workspace w = carve_workspace(const_cast<void *>(workspace), n_atoms, n_res);
    cudaStream_t st = (cudaStream_t)stream;
    PESTO_CUDA(cudaMemcpyAsync(status_host, w.status, PESTO_STATUS_WORDS * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    PESTO_CUDA(cudaStreamSynchronize(st));
    status_host[0] = 0;                                    // (word 0 is geometry scratch, not a flag)
    if (status_host[1]) { set_error("pesto_forward: a neighbour id in ids_topk is out of range [0, n_atoms]"); return PESTO_EINPUT; }
    if (status_host[2]) { set_error("pesto_forward: a row of the membership matrix M is not one-hot"); return PESTO_EINPUT; }
    if (status_host[4]) { set_error("pesto_forward: a residue index is out of range [0, n_res)"); return PESTO_EINPUT; }
    if (status_host[3] >= 100) {
        set_error("pesto_forward: the state left the range of the fp16 operand planes (|q| or |p| > 2^14, or NaN): run this "
                  "model / input in mode fp32; the logits are NaN");
        return PESTO_EINPUT;
    }
    if (status_host[3]) {
        set_error("pesto_forward: tensor-core stage %d never completed (watchdog): the logits are NaN", status_host[3]);
        return PESTO_ECUDA;
    }
    return PESTO_OK;
}

int pesto_debug_watchdog(int32_t *value_host) {
    int *p = device_watchdog_word();
    if (!p || !value_host) {
        set_error("pesto_debug_watchdog: no device word");
        return PESTO_ECUDA;
    }
    PESTO_CUDA(cudaMemcpy(value_host, p, sizeof(int), cudaMemcpyDeviceToHost));
    PESTO_CUDA(cudaMemset(p, 0, sizeof(int)));
    return PESTO_OK;
}

int pesto_forward_launch_count(const pesto_model_t *m, int dense_m, int mode) {
    if (!m) return 0;
    return 3 + 2 * m->n_layers + (mode == PESTO_MODE_FP32 ? 0 : 1) + (dense_m ? 1 : 0) + 5;
}

}  // extern "C"
