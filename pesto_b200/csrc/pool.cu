// StatePoolLayer + decoder (src/model_operations.py:197-213, model/model.py:46-50) as a segmented softmax over
// the atoms of each residue.  The reference builds a dense [N, R, 8] tensor; here the membership
// matrix is reduced once to a residue column per atom and every residue is handled by one warp.
#include "common.cuh"

namespace pesto {

namespace {

using H = HeadLayout;
constexpr unsigned FULL = 0xffffffffu;

// dense one-hot M[N,R] -> rid[N].  One warp per atom row.
__global__ void __launch_bounds__(256)
residue_index_kernel(const float *__restrict__ M, int n_atoms, int n_res, int32_t *__restrict__ rid,
                     int32_t *__restrict__ flags) {
    int lane = threadIdx.x & 31;
    int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= n_atoms) return;
    const float *row = M + (size_t)i * n_res;
    int cnt = 0, col = -1;
    for (int c = lane; c < n_res; c += 32) {
        float v = __ldg(row + c);
        if (v != 0.f) {
            ++cnt;
            col = c;
            if (v != 1.f) cnt += 2;
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        cnt += __shfl_xor_sync(FULL, cnt, o);
        col = max(col, __shfl_xor_sync(FULL, col, o));
    }
    if (lane == 0) {
        rid[i] = col < 0 ? 0 : col;
        if (cnt != 1) atomicOr(flags, 1);     // not a one-hot row: unsupported membership
    }
}

// per atom: attention logits a = sam([q, |p|]) (8 = Nh x {scalar, vector}); residue histogram; order check
__global__ void __launch_bounds__(256)
pool_logits_kernel(const float *__restrict__ hw, const float *__restrict__ state, const int32_t *__restrict__ rid,
                   int n_atoms, int n_res, float *__restrict__ alog, int32_t *__restrict__ cnt,
                   int32_t *__restrict__ flags) {
    int lane = threadIdx.x & 31;
    int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= n_atoms) return;
    const float *sr = state + (size_t)(i + 1) * SR;
    float q = sr[lane], x = sr[32 + lane], y = sr[64 + lane], z = sr[96 + lane];
    float pn = sqrtf(x * x + y * y + z * z);
    float h = hw[H::SAM_B1 + lane];
#pragma unroll
    for (int k = 0; k < 32; ++k) {
        h = fmaf(__shfl_sync(FULL, q, k), __ldg(hw + H::SAM_W1 + k * 32 + lane), h);
        h = fmaf(__shfl_sync(FULL, pn, k), __ldg(hw + H::SAM_W1 + (32 + k) * 32 + lane), h);
    }
    h = elu(h);
    float g = hw[H::SAM_B2 + lane];
#pragma unroll
    for (int k = 0; k < 32; ++k) g = fmaf(__shfl_sync(FULL, h, k), __ldg(hw + H::SAM_W2 + k * 32 + lane), g);
    g = elu(g);
    float o = hw[H::SAM_B3 + (lane & 7)];
#pragma unroll
    for (int k = 0; k < 32; ++k) o = fmaf(__shfl_sync(FULL, g, k), __ldg(hw + H::SAM_W3 + k * 8 + (lane & 7)), o);
    if (lane < 8) alog[(size_t)i * 8 + lane] = o;
    if (lane == 0) {
        int r = rid[i];
        if (r < 0 || r >= n_res) {
            atomicOr(flags + 1, 1);
        } else {
            atomicAdd(cnt + r, 1);
            if (i > 0 && rid[i - 1] > r) atomicOr(flags, 1);   // atoms of a residue are not one contiguous run
        }
    }
}

// off = exclusive scan of cnt (single block)
__global__ void __launch_bounds__(1024)
scan_kernel(const int32_t *__restrict__ cnt, int n_res, int32_t *__restrict__ off) {
    __shared__ int warp_sums[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < n_res; base += 1024) {
        int idx = base + threadIdx.x;
        int v = idx < n_res ? cnt[idx] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int ws = warp_sums[lane], wi = ws;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(FULL, wi, o);
                if (lane >= o) wi += t;
            }
            warp_sums[lane] = wi - ws;     // exclusive prefix of warp totals
        }
        __syncthreads();
        int c = carry;
        if (idx < n_res) off[idx] = c + warp_sums[warp] + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = c + warp_sums[warp] + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) off[n_res] = carry;
}

// only for non-contiguous residues: scatter atom indices into their residue's slot range ...
__global__ void fill_kernel(const int32_t *__restrict__ rid, int n_atoms, int n_res, const int32_t *__restrict__ off,
                            int32_t *__restrict__ cursor, int32_t *__restrict__ perm_tmp,
                            const int32_t *__restrict__ flags) {
    if (!flags[0]) return;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_atoms) return;
    int r = rid[i];
    if (r < 0 || r >= n_res) return;
    perm_tmp[off[r] + atomicAdd(cursor + r, 1)] = i;
}
// ... and order every range by atom index (rank counting), so that the summation order is deterministic
__global__ void __launch_bounds__(256)
sort_segments_kernel(int n_res, const int32_t *__restrict__ off, const int32_t *__restrict__ perm_tmp,
                     int32_t *__restrict__ perm, const int32_t *__restrict__ flags) {
    if (!flags[0]) return;
    int lane = threadIdx.x & 31;
    int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= n_res) return;
    int a = off[r], b = off[r + 1];
    for (int e = a + lane; e < b; e += 32) {
        int v = perm_tmp[e], rank = 0;
        for (int f = a; f < b; ++f) rank += perm_tmp[f] < v;
        perm[a + rank] = v;
    }
}

// one warp per residue: softmax over its atoms, weighted sums, zdm / zdm_vec, decoder
__global__ void __launch_bounds__(256)
residue_kernel(const float *__restrict__ hw, const float *__restrict__ state, const float *__restrict__ alog,
               const int32_t *__restrict__ off, const int32_t *__restrict__ perm, const int32_t *__restrict__ flags,
               int32_t *__restrict__ poison, int n_res, float *__restrict__ z) {
    int lane = threadIdx.x & 31;
    int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= n_res) return;
    // invalid input (index out of range, membership not one-hot) and a tensor-core stage that never completed cannot be
    // reported without a host sync: the logits are poisoned instead; pesto_forward_status reads the flags (the Python side
    // raises from them wherever it synchronises anyway)
    if (flags[1] || (poison && (poison[0] || poison[1] || poison[2]))) {
        if (flags[1] && poison && r == 0 && lane == 0) poison[3] = 1;      // residue index out of range -> status word 4
        const int n_out = (int)hw[H::META_NUM_OUT];
        if (lane < n_out) z[(size_t)r * n_out + lane] = __int_as_float(0x7fc00000);
        return;
    }
    const bool unsorted = flags[0] != 0;
    const int a = off[r], b = off[r + 1];
    float m[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) m[u] = -INFINITY;
    for (int e = a; e < b; ++e) {
        int atom = unsorted ? perm[e] : e;
        const float4 lo = __ldg(reinterpret_cast<const float4 *>(alog + (size_t)atom * 8));
        const float4 hi = __ldg(reinterpret_cast<const float4 *>(alog + (size_t)atom * 8 + 4));
        m[0] = fmaxf(m[0], lo.x); m[1] = fmaxf(m[1], lo.y); m[2] = fmaxf(m[2], lo.z); m[3] = fmaxf(m[3], lo.w);
        m[4] = fmaxf(m[4], hi.x); m[5] = fmaxf(m[5], hi.y); m[6] = fmaxf(m[6], hi.z); m[7] = fmaxf(m[7], hi.w);
    }
    float den[8], qh[4], ph[3][4];
#pragma unroll
    for (int u = 0; u < 8; ++u) den[u] = 0.f;
#pragma unroll
    for (int h = 0; h < 4; ++h) qh[h] = ph[0][h] = ph[1][h] = ph[2][h] = 0.f;
    for (int e = a; e < b; ++e) {
        int atom = unsorted ? perm[e] : e;
        const float4 lo = __ldg(reinterpret_cast<const float4 *>(alog + (size_t)atom * 8));
        const float4 hi = __ldg(reinterpret_cast<const float4 *>(alog + (size_t)atom * 8 + 4));
        const float av[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
        float ex[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            ex[u] = expf(av[u] - m[u]);
            den[u] += ex[u];
        }
        const float *sr = state + (size_t)(atom + 1) * SR;
        const float q = sr[lane], x = sr[32 + lane], y = sr[64 + lane], zz = sr[96 + lane];
#pragma unroll
        for (int h = 0; h < 4; ++h) {                         // logit index h*2 + t, t = 0 scalar / 1 vector
            qh[h] = fmaf(ex[2 * h], q, qh[h]);
            ph[0][h] = fmaf(ex[2 * h + 1], x, ph[0][h]);
            ph[1][h] = fmaf(ex[2 * h + 1], y, ph[1][h]);
            ph[2][h] = fmaf(ex[2 * h + 1], zz, ph[2][h]);
        }
    }
#pragma unroll
    for (int h = 0; h < 4; ++h) {
        const float iq = b > a ? 1.0f / den[2 * h] : 0.f, ip = b > a ? 1.0f / den[2 * h + 1] : 0.f;
        qh[h] *= iq;
        ph[0][h] *= ip; ph[1][h] *= ip; ph[2][h] *= ip;
    }
    // zdm on qh flattened s-major (index s*4 + h), zdm_vec on ph likewise
    float y1 = hw[H::ZDM_B1 + lane], pr0 = 0.f, pr1 = 0.f, pr2 = 0.f;
    for (int s = 0; s < 32; ++s) {
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const int k = s * 4 + h;
            y1 = fmaf(__shfl_sync(FULL, qh[h], s), __ldg(hw + H::ZDM_W1 + k * 32 + lane), y1);
            const float wv = __ldg(hw + H::ZDV_W + k * 32 + lane);
            pr0 = fmaf(__shfl_sync(FULL, ph[0][h], s), wv, pr0);
            pr1 = fmaf(__shfl_sync(FULL, ph[1][h], s), wv, pr1);
            pr2 = fmaf(__shfl_sync(FULL, ph[2][h], s), wv, pr2);
        }
    }
    y1 = elu(y1);
    float y2 = hw[H::ZDM_B2 + lane];
#pragma unroll
    for (int k = 0; k < 32; ++k) y2 = fmaf(__shfl_sync(FULL, y1, k), __ldg(hw + H::ZDM_W2 + k * 32 + lane), y2);
    y2 = elu(y2);
    float qr = hw[H::ZDM_B3 + lane];
#pragma unroll
    for (int k = 0; k < 32; ++k) qr = fmaf(__shfl_sync(FULL, y2, k), __ldg(hw + H::ZDM_W3 + k * 32 + lane), qr);
    const float prn = sqrtf(pr0 * pr0 + pr1 * pr1 + pr2 * pr2);
    // decoder dm([qr, |pr|])
    const int n_out = (int)hw[H::META_NUM_OUT];
    if (hw[H::META_DM_LAYERS] < 2.f) {                          // one Linear (model/save/i_v3_1*/model.py:20-22)
        float o = hw[H::DM_B1 + (lane & 7)];
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            o = fmaf(__shfl_sync(FULL, qr, k), __ldg(hw + H::DM_W1 + k * 8 + (lane & 7)), o);
            o = fmaf(__shfl_sync(FULL, prn, k), __ldg(hw + H::DM_W1 + (32 + k) * 8 + (lane & 7)), o);
        }
        if (lane < n_out) z[(size_t)r * n_out + lane] = o;
        return;
    }
    float d1 = hw[H::DM_B1 + lane];
#pragma unroll
    for (int k = 0; k < 32; ++k) {
        d1 = fmaf(__shfl_sync(FULL, qr, k), __ldg(hw + H::DM_W1 + k * 32 + lane), d1);
        d1 = fmaf(__shfl_sync(FULL, prn, k), __ldg(hw + H::DM_W1 + (32 + k) * 32 + lane), d1);
    }
    d1 = elu(d1);
    float d2 = hw[H::DM_B2 + lane];
#pragma unroll
    for (int k = 0; k < 32; ++k) d2 = fmaf(__shfl_sync(FULL, d1, k), __ldg(hw + H::DM_W2 + k * 32 + lane), d2);
    d2 = elu(d2);
    float o = hw[H::DM_B3 + (lane & 7)];
#pragma unroll
    for (int k = 0; k < 32; ++k) o = fmaf(__shfl_sync(FULL, d2, k), __ldg(hw + H::DM_W3 + k * 8 + (lane & 7)), o);
    if (lane < n_out) z[(size_t)r * n_out + lane] = o;
}

struct PoolScratch {
    float *alog;
    int32_t *cnt, *cursor, *flags, *off, *perm_tmp, *perm;
    size_t zero_bytes;
};

PoolScratch carve(void *scratch, int n_atoms, int n_res) {
    PoolScratch p;
    char *c = (char *)scratch;
    p.alog = (float *)c;                 c += (size_t)n_atoms * 8 * sizeof(float);
    p.cnt = (int32_t *)c;                c += (size_t)n_res * sizeof(int32_t);
    p.cursor = (int32_t *)c;             c += (size_t)n_res * sizeof(int32_t);
    p.flags = (int32_t *)c;              c += 4 * sizeof(int32_t);
    p.zero_bytes = (size_t)(c - (char *)p.cnt);
    p.off = (int32_t *)c;                c += ((size_t)n_res + 4) * sizeof(int32_t);
    p.perm_tmp = (int32_t *)c;           c += (size_t)n_atoms * sizeof(int32_t);
    p.perm = (int32_t *)c;
    return p;
}

}  // namespace

size_t pool_scratch_bytes(int n_atoms, int n_res) {
    return (size_t)n_atoms * (8 * sizeof(float) + 2 * sizeof(int32_t)) + ((size_t)n_res * 3 + 8) * sizeof(int32_t) + 256;
}

int launch_residue_index(const float *M, int n_atoms, int n_res, int32_t *rid, int32_t *flags, cudaStream_t st) {
    PESTO_CUDA(cudaMemsetAsync(flags, 0, sizeof(int32_t), st));
    residue_index_kernel<<<(n_atoms + 7) / 8, 256, 0, st>>>(M, n_atoms, n_res, rid, flags);
    PESTO_CUDA(cudaGetLastError());
    return PESTO_OK;
}

int launch_pool_decode(const float *head_w, const float *state, const int32_t *rid, int n_atoms, int n_res, float *z,
                       void *scratch, int32_t *poison, cudaStream_t st) {
    PoolScratch p = carve(scratch, n_atoms, n_res);
    PESTO_CUDA(cudaMemsetAsync(p.cnt, 0, p.zero_bytes, st));
    pool_logits_kernel<<<(n_atoms + 7) / 8, 256, 0, st>>>(head_w, state, rid, n_atoms, n_res, p.alog, p.cnt, p.flags);
    scan_kernel<<<1, 1024, 0, st>>>(p.cnt, n_res, p.off);
    fill_kernel<<<(n_atoms + 255) / 256, 256, 0, st>>>(rid, n_atoms, n_res, p.off, p.cursor, p.perm_tmp, p.flags);
    sort_segments_kernel<<<(n_res + 7) / 8, 256, 0, st>>>(n_res, p.off, p.perm_tmp, p.perm, p.flags);
    residue_kernel<<<(n_res + 7) / 8, 256, 0, st>>>(head_w, state, p.alog, p.off, p.perm, p.flags, poison, n_res, z);
    PESTO_CUDA(cudaGetLastError());
    return PESTO_OK;
}

}  // namespace pesto
