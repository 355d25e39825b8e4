// Exact k-nearest-neighbour topology.  Replaces extract_topology (src/data_encoding.py:87-102) and, with
// base = 1, the index shift + sink padding of collate_batch_features (src/dataset.py:100-109).
//
// One warp per query atom scans the atoms of its own structure (coordinates stay L1/L2 resident: a
// 32768-atom structure is 384 KB) and keeps the 64 best (distance, index) keys sorted across the warp
// (two 64-bit keys per lane).  The N x N distance / displacement tensors of the reference are never
// written.  Distances are bit-exact with the reference's fp32 recipe; ties are broken by index.
//
// Masking (src/data_encoding.py:93): entries with D < 1e-2 get + max(D of the structure), i.e. they rank
// after every unmasked entry.  The main kernel therefore selects among unmasked entries only and flags
// the (rare: structures with <= 64 atoms, duplicated atoms) rows that need masked entries; two small
// follow-up kernels compute max(D) for flagged structures and redo flagged rows with the full keys.
#include "common.cuh"

namespace pesto {

namespace {

constexpr int KNN_WARPS = 8;
constexpr unsigned long long KEY_NONE = 0xFFFFFFFFFFFFFFFFull;
constexpr float MASK_THR = 1e-2f;

struct TopK {
    unsigned long long k0, k1;   // sorted positions lane and 32 + lane
    unsigned long long thr;      // key at position 63 (warp-uniform)
};

__device__ __forceinline__ void topk_init(TopK &t) { t.k0 = t.k1 = t.thr = KEY_NONE; }

// insert key c (warp-uniform, c < t.thr) into the distributed sorted list
__device__ __forceinline__ void topk_insert(TopK &t, unsigned long long c, int lane) {
    unsigned long long up0 = __shfl_up_sync(0xffffffffu, t.k0, 1);
    unsigned long long up1 = __shfl_up_sync(0xffffffffu, t.k1, 1);
    unsigned long long cross = __shfl_sync(0xffffffffu, t.k0, 31);
    unsigned long long prev1 = lane == 0 ? cross : up1;
    unsigned long long n0 = t.k0, n1 = t.k1;
    if (t.k0 > c) n0 = (lane > 0 && up0 > c) ? up0 : c;
    if (t.k1 > c) n1 = (prev1 > c) ? prev1 : c;
    t.k0 = n0;
    t.k1 = n1;
    t.thr = __shfl_sync(0xffffffffu, t.k1, 31);
}

__device__ __forceinline__ void topk_offer(TopK &t, unsigned long long key, int lane) {
    unsigned m = __ballot_sync(0xffffffffu, key < t.thr);
    while (m) {
        int src = __ffs(m) - 1;
        m &= m - 1;
        unsigned long long c = __shfl_sync(0xffffffffu, key, src);
        if (c < t.thr) topk_insert(t, c, lane);
    }
}

__device__ __forceinline__ int find_segment(const int32_t *__restrict__ seg_off, int n_seg, int i) {
    int lo = 0, hi = n_seg;   // invariant: seg_off[lo] <= i < seg_off[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (seg_off[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}

// MODE 0: unmasked entries only.  MODE 1: all entries with key distance D' = D + maxd * (D < 1e-2).
template <int MODE>
__device__ __forceinline__ void scan_range(TopK &t, float &rowmax, const float *__restrict__ X, int lo, int a, int b,
                                           float xi, float yi, float zi, float maxd, int lane) {
    for (int c0 = a; c0 < b; c0 += 32) {
        int j = c0 + lane;
        unsigned long long key = KEY_NONE;
        if (j < b) {
            float dx = __ldg(X + 3 * (size_t)j + 0) - xi;
            float dy = __ldg(X + 3 * (size_t)j + 1) - yi;
            float dz = __ldg(X + 3 * (size_t)j + 2) - zi;
            float d = dist_exact(dx, dy, dz);
            bool masked = d < MASK_THR;
            if (MODE == 1) {
                if (masked) d = __fadd_rn(d, maxd);
                key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)(j - lo);
            } else if (!masked) {
                rowmax = fmaxf(rowmax, d);
                key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)(j - lo);
            }
        }
        topk_offer(t, key, lane);
    }
}

template <int MODE>
__device__ __forceinline__ void scan_segment(TopK &t, float &rowmax, const float *__restrict__ X, int lo, int hi, int i,
                                             float xi, float yi, float zi, float maxd, int lane) {
    // chain-ordered structures have most neighbours close in index: scan outwards from i in 64-atom steps (both sides
    // alternately) so that the list fills with near-final entries and the threshold tightens early -- every later
    // candidate that passes costs a warp-wide insertion -- then the rest of the structure
    constexpr int W = 256, STEP = 64;
    const int c0 = max(lo, (i / 32) * 32 - 32);            // first block: [c0, c0 + 96) contains i
    int left = c0, right = min(hi, c0 + 96);
    scan_range<MODE>(t, rowmax, X, lo, left, right, xi, yi, zi, maxd, lane);
#pragma unroll 1
    for (int s = 0; s < W / STEP; ++s) {
        const int nl = max(lo, left - STEP), nr = min(hi, right + STEP);
        if (nl < left) scan_range<MODE>(t, rowmax, X, lo, nl, left, xi, yi, zi, maxd, lane);
        if (nr > right) scan_range<MODE>(t, rowmax, X, lo, right, nr, xi, yi, zi, maxd, lane);
        left = nl;
        right = nr;
    }
    scan_range<MODE>(t, rowmax, X, lo, lo, left, xi, yi, zi, maxd, lane);
    scan_range<MODE>(t, rowmax, X, lo, right, hi, xi, yi, zi, maxd, lane);
}

__device__ __forceinline__ void write_row(const TopK &t, const float *__restrict__ X, int i, int lo, int n, int k,
                                          int base, float xi, float yi, float zi, int64_t *__restrict__ ids_out,
                                          float *__restrict__ d_out, float *__restrict__ r_out, int lane) {
    int kk = min(k, n);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        int pos = lane + 32 * half;
        if (pos >= k) continue;
        unsigned long long key = half ? t.k1 : t.k0;
        size_t o = (size_t)i * k + pos;
        if (pos < kk) {
            int jl = (int)(key & 0xffffffffu);
            float d = __uint_as_float((unsigned)(key >> 32));
            ids_out[o] = base ? (int64_t)(jl + lo + 1) : (int64_t)jl;
            if (d_out) d_out[o] = d;
            if (r_out) {
                int j = jl + lo;
                r_out[3 * o + 0] = __fdiv_rn(__ldg(X + 3 * (size_t)j + 0) - xi, d);
                r_out[3 * o + 1] = __fdiv_rn(__ldg(X + 3 * (size_t)j + 1) - yi, d);
                r_out[3 * o + 2] = __fdiv_rn(__ldg(X + 3 * (size_t)j + 2) - zi, d);
            }
        } else {
            ids_out[o] = base ? 0 : -1;
            if (d_out) d_out[o] = 0.f;
            if (r_out) r_out[3 * o + 0] = r_out[3 * o + 1] = r_out[3 * o + 2] = 0.f;
        }
    }
}

__global__ void __launch_bounds__(KNN_WARPS * 32)
knn_main_kernel(const float *__restrict__ X, int n_atoms, const int32_t *__restrict__ seg_off, int n_seg, int k, int base,
                int64_t *__restrict__ ids_out, float *__restrict__ d_out, float *__restrict__ r_out,
                int32_t *__restrict__ seg_flag, int32_t *__restrict__ row_flag) {
    int lane = threadIdx.x & 31;
    int i = blockIdx.x * KNN_WARPS + (threadIdx.x >> 5);
    if (i >= n_atoms) return;
    int s = find_segment(seg_off, n_seg, i);
    int lo = seg_off[s], hi = seg_off[s + 1], n = hi - lo;
    float xi = __ldg(X + 3 * (size_t)i), yi = __ldg(X + 3 * (size_t)i + 1), zi = __ldg(X + 3 * (size_t)i + 2);
    TopK t;
    topk_init(t);
    float rowmax = 0.f;
    scan_segment<0>(t, rowmax, X, lo, hi, i, xi, yi, zi, 0.f, lane);
#pragma unroll
    for (int o = 16; o; o >>= 1) rowmax = fmaxf(rowmax, __shfl_xor_sync(0xffffffffu, rowmax, o));
    int kk = min(min(k, KMAX), n);
    // key at sorted position kk-1
    unsigned long long last = __shfl_sync(0xffffffffu, (kk - 1) < 32 ? t.k0 : t.k1, (kk - 1) & 31);
    bool incomplete = (last == KEY_NONE) || !(__uint_as_float((unsigned)(last >> 32)) < rowmax);
    if (incomplete) {   // needs masked entries or may tie with them: redo with full keys later
        if (lane == 0) {
            row_flag[i] = 1;
            seg_flag[s] = 1;
        }
        return;
    }
    if (lane == 0) row_flag[i] = 0;
    write_row(t, X, i, lo, n, k, base, xi, yi, zi, ids_out, d_out, r_out, lane);
}

__global__ void __launch_bounds__(KNN_WARPS * 32)
knn_segmax_kernel(const float *__restrict__ X, int n_atoms, const int32_t *__restrict__ seg_off, int n_seg,
                  const int32_t *__restrict__ seg_flag, unsigned *__restrict__ seg_maxd) {
    int lane = threadIdx.x & 31;
    int i = blockIdx.x * KNN_WARPS + (threadIdx.x >> 5);
    if (i >= n_atoms) return;
    int s = find_segment(seg_off, n_seg, i);
    if (!seg_flag[s]) return;
    int lo = seg_off[s], hi = seg_off[s + 1];
    float xi = __ldg(X + 3 * (size_t)i), yi = __ldg(X + 3 * (size_t)i + 1), zi = __ldg(X + 3 * (size_t)i + 2);
    float m = 0.f;
    for (int j = lo + lane; j < hi; j += 32)
        m = fmaxf(m, dist_exact(__ldg(X + 3 * (size_t)j) - xi, __ldg(X + 3 * (size_t)j + 1) - yi,
                                __ldg(X + 3 * (size_t)j + 2) - zi));
#pragma unroll
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) atomicMax(seg_maxd + s, __float_as_uint(m));   // non-negative floats order like their bits
}

__global__ void __launch_bounds__(KNN_WARPS * 32)
knn_exact_rows_kernel(const float *__restrict__ X, int n_atoms, const int32_t *__restrict__ seg_off, int n_seg, int k,
                      int base, int64_t *__restrict__ ids_out, float *__restrict__ d_out, float *__restrict__ r_out,
                      const int32_t *__restrict__ row_flag, const unsigned *__restrict__ seg_maxd) {
    int lane = threadIdx.x & 31;
    int i = blockIdx.x * KNN_WARPS + (threadIdx.x >> 5);
    if (i >= n_atoms || !row_flag[i]) return;
    int s = find_segment(seg_off, n_seg, i);
    int lo = seg_off[s], hi = seg_off[s + 1], n = hi - lo;
    float maxd = __uint_as_float(seg_maxd[s]);
    float xi = __ldg(X + 3 * (size_t)i), yi = __ldg(X + 3 * (size_t)i + 1), zi = __ldg(X + 3 * (size_t)i + 2);
    TopK t;
    topk_init(t);
    float rowmax = 0.f;
    scan_segment<1>(t, rowmax, X, lo, hi, i, xi, yi, zi, maxd, lane);
    write_row(t, X, i, lo, n, k, base, xi, yi, zi, ids_out, d_out, r_out, lane);
}

}  // namespace

size_t knn_scratch_bytes(int n_atoms, int n_seg) {
    return ((size_t)n_seg * 2 + (size_t)n_atoms) * sizeof(int32_t) + 256;
}

int launch_knn(const float *X, int n_atoms, const int32_t *seg_off, int n_seg, int k, int base, int64_t *ids_out,
               float *d_out, float *r_out, void *scratch, cudaStream_t st) {
    if (n_atoms <= 0) return PESTO_OK;
    int32_t *seg_flag = (int32_t *)scratch;
    unsigned *seg_maxd = (unsigned *)(seg_flag + n_seg);
    int32_t *row_flag = (int32_t *)(seg_maxd + n_seg);
    PESTO_CUDA(cudaMemsetAsync(scratch, 0, (size_t)n_seg * 2 * sizeof(int32_t), st));
    int grid = (n_atoms + KNN_WARPS - 1) / KNN_WARPS;
    knn_main_kernel<<<grid, KNN_WARPS * 32, 0, st>>>(X, n_atoms, seg_off, n_seg, k, base, ids_out, d_out, r_out,
                                                     seg_flag, row_flag);
    knn_segmax_kernel<<<grid, KNN_WARPS * 32, 0, st>>>(X, n_atoms, seg_off, n_seg, seg_flag, seg_maxd);
    knn_exact_rows_kernel<<<grid, KNN_WARPS * 32, 0, st>>>(X, n_atoms, seg_off, n_seg, k, base, ids_out, d_out, r_out,
                                                           row_flag, seg_maxd);
    PESTO_CUDA(cudaGetLastError());
    return PESTO_OK;
}

}  // namespace pesto
