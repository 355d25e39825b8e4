// Exact k-nearest-neighbour topology.  Replaces extract_topology (src/data_encoding.py:87-102) and, with
// base = 1, the index shift + sink padding of collate_batch_features (src/dataset.py:100-109).
//
// One warp per query atom scans the atoms of its own structure (coordinates stay L1/L2 resident: a
// 32768-atom structure is 384 KB) and keeps the 64 best (distance, index) keys sorted across the warp
// (two 64-bit keys per lane); a chunk with many candidates is sorted and merged in (bitonic), a few are inserted
// singly.  The N x N distance / displacement tensors of the reference are never written.  Distances are bit-exact
// with the reference's fp32 recipe; ties are broken by index.
//
// Pruning: a pre-pass stores the bounding box of every aligned 32-atom chunk (chain-ordered atoms: four residues, a
// ~10 A box).  After the query's own neighbourhood in the chain has filled the list, every other chunk is first tested
// box-against-threshold, 32 chunks per warp instruction (one per lane); only chunks whose box comes within the current
// 64th distance are scanned (about ten of the 78 chunks of a 2 500-atom structure).  The box distance is built from the
// same monotone fp32 operations as the atom distance, so it never exceeds the distance of an atom inside the box and
// the selection stays exact.
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

// The 64 best keys, sorted ascending, interleaved over the warp: lane l holds positions 2 l and 2 l + 1 (a shift by one
// position then needs one cross-lane move, and a lane's two output ids are one 16-byte store).
struct TopK {
    unsigned long long k0, k1;   // sorted positions 2 * lane and 2 * lane + 1
    unsigned long long thr;      // key at position 63 (warp-uniform)
};

__device__ __forceinline__ void topk_init(TopK &t) { t.k0 = t.k1 = t.thr = KEY_NONE; }

__device__ __forceinline__ unsigned long long umin64(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
__device__ __forceinline__ unsigned long long umax64(unsigned long long a, unsigned long long b) { return a < b ? b : a; }

// insert key c (warp-uniform, c < t.thr) into the distributed sorted list
__device__ __forceinline__ void topk_insert(TopK &t, unsigned long long c, int lane) {
    const unsigned long long up = __shfl_up_sync(0xffffffffu, t.k1, 1);       // position 2 * lane - 1
    const bool g0 = t.k0 > c, g1 = t.k1 > c, gu = lane > 0 && up > c;
    const unsigned long long n0 = g0 ? (gu ? up : c) : t.k0;
    const unsigned long long n1 = g1 ? (g0 ? t.k0 : c) : t.k1;
    t.k0 = n0;
    t.k1 = n1;
    t.thr = __shfl_sync(0xffffffffu, t.k1, 31);
}

// merge the 32 keys of a chunk (one per lane, any order, KEY_NONE = no candidate) into the list: bitonic sort of the
// chunk, min against the reversed upper half of the list (the 64 smallest of the union, as a bitonic sequence), bitonic
// merge.  About the cost of eight single insertions.
__device__ __forceinline__ void topk_merge(TopK &t, unsigned long long key, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, key, j);
            const bool keep_min = ((lane & j) == 0) == ((lane & k) == 0);
            key = keep_min ? umin64(key, o) : umax64(key, o);
        }
    }
    // list position p >= 32 meets chunk rank 63 - p; positions < 32 meet the padding
    const unsigned long long ca = __shfl_sync(0xffffffffu, key, (63 - 2 * lane) & 31);
    const unsigned long long cb = __shfl_sync(0xffffffffu, key, (62 - 2 * lane) & 31);
    if (lane >= 16) {
        t.k0 = umin64(t.k0, ca);
        t.k1 = umin64(t.k1, cb);
    }
#pragma unroll
    for (int jl = 16; jl > 0; jl >>= 1) {                                      // position distance 2 * jl
        const unsigned long long o0 = __shfl_xor_sync(0xffffffffu, t.k0, jl);
        const unsigned long long o1 = __shfl_xor_sync(0xffffffffu, t.k1, jl);
        const bool keep_min = (lane & jl) == 0;
        t.k0 = keep_min ? umin64(t.k0, o0) : umax64(t.k0, o0);
        t.k1 = keep_min ? umin64(t.k1, o1) : umax64(t.k1, o1);
    }
    const unsigned long long a = umin64(t.k0, t.k1), b = umax64(t.k0, t.k1);
    t.k0 = a;
    t.k1 = b;
    t.thr = __shfl_sync(0xffffffffu, t.k1, 31);
}

constexpr int MERGE_MIN = 8;     // candidates of one chunk from which the merge is cheaper than single insertions

__device__ __forceinline__ void topk_offer(TopK &t, unsigned long long key, int lane) {
    unsigned m = __ballot_sync(0xffffffffu, key < t.thr);
    if (__popc(m) >= MERGE_MIN) {
        topk_merge(t, key, lane);
        return;
    }
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

// one aligned chunk g (atoms 32 g .. 32 g + 31, clipped to the structure [lo, hi)), unmasked entries only
__device__ __forceinline__ void scan_chunk(TopK &t, float &rowmax, const float *__restrict__ X, int lo, int hi, int g,
                                           float xi, float yi, float zi, int lane) {
    const int j = g * 32 + lane;
    unsigned long long key = KEY_NONE;
    if (j >= lo && j < hi) {
        float dx = __ldg(X + 3 * (size_t)j + 0) - xi;
        float dy = __ldg(X + 3 * (size_t)j + 1) - yi;
        float dz = __ldg(X + 3 * (size_t)j + 2) - zi;
        float d = dist_exact(dx, dy, dz);
        if (!(d < MASK_THR)) {
            rowmax = fmaxf(rowmax, d);
            key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)(j - lo);
        }
    }
    topk_offer(t, key, lane);
}

__device__ __forceinline__ float thr_distance(const TopK &t) { return __uint_as_float((unsigned)(t.thr >> 32)); }   // NaN while the list is not full

// MODE 0 scan with box pruning.  Returns rowmax = +inf when a chunk was skipped: every atom of a skipped chunk is
// unmasked and strictly farther than the 64th entry, which is all the caller asks of rowmax.
__device__ __forceinline__ void scan_segment_pruned(TopK &t, float &rowmax, const float *__restrict__ X,
                                                    const float4 *__restrict__ boxes, int lo, int hi, int i, float xi,
                                                    float yi, float zi, int lane) {
    const int glo = lo >> 5, ghi = (hi - 1) >> 5, gi = i >> 5;
    constexpr int WIN = 2;                                      // the query's chunk and two either side, nearest first
    scan_chunk(t, rowmax, X, lo, hi, gi, xi, yi, zi, lane);
#pragma unroll 1
    for (int s = 1; s <= WIN; ++s) {
        if (gi - s >= glo) scan_chunk(t, rowmax, X, lo, hi, gi - s, xi, yi, zi, lane);
        if (gi + s <= ghi) scan_chunk(t, rowmax, X, lo, hi, gi + s, xi, yi, zi, lane);
    }
    bool skipped = false;
#pragma unroll 1
    for (int g0 = glo; g0 <= ghi; g0 += 32) {
        const int g = g0 + lane;
        float bd = 0.f;
        const bool want = g <= ghi && (g < gi - WIN || g > gi + WIN);
        if (want) {
            const float4 mn = __ldg(boxes + 2 * (size_t)g), mx = __ldg(boxes + 2 * (size_t)g + 1);
            // fl(x_j - xi) >= fl(mn - xi) and fl(xi - x_j) >= fl(xi - mx) for every atom of the chunk (rounding is monotone)
            const float bx = fmaxf(fmaxf(mn.x - xi, xi - mx.x), 0.f);
            const float by = fmaxf(fmaxf(mn.y - yi, yi - mx.y), 0.f);
            const float bz = fmaxf(fmaxf(mn.z - zi, zi - mx.z), 0.f);
            bd = dist_exact(bx, by, bz);
        }
        const bool pass = want && !(bd > thr_distance(t));
        skipped |= want && !pass;
        unsigned m = __ballot_sync(0xffffffffu, pass);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const float b = __shfl_sync(0xffffffffu, bd, src);
            if (b > thr_distance(t)) {                          // the threshold has tightened since the ballot
                skipped = true;
                continue;
            }
            scan_chunk(t, rowmax, X, lo, hi, g0 + src, xi, yi, zi, lane);
        }
    }
    if (__any_sync(0xffffffffu, skipped)) rowmax = __int_as_float(0x7f800000);
}

// bounding box of every aligned 32-atom chunk: boxes[2 g] = min xyz, boxes[2 g + 1] = max xyz (NaN coordinates are ignored:
// such atoms never enter a full list)
__global__ void __launch_bounds__(KNN_WARPS * 32)
knn_boxes_kernel(const float *__restrict__ X, int n_atoms, float4 *__restrict__ boxes) {
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * KNN_WARPS + (threadIdx.x >> 5);
    if (g >= (n_atoms + 31) >> 5) return;
    const int j = g * 32 + lane;
    const float inf = __int_as_float(0x7f800000);
    float mnx = inf, mny = inf, mnz = inf, mxx = -inf, mxy = -inf, mxz = -inf;
    if (j < n_atoms) {
        mnx = mxx = __ldg(X + 3 * (size_t)j + 0);
        mny = mxy = __ldg(X + 3 * (size_t)j + 1);
        mnz = mxz = __ldg(X + 3 * (size_t)j + 2);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
        mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o));
        mnz = fminf(mnz, __shfl_xor_sync(0xffffffffu, mnz, o));
        mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
        mxz = fmaxf(mxz, __shfl_xor_sync(0xffffffffu, mxz, o));
    }
    if (lane == 0) {
        boxes[2 * (size_t)g] = make_float4(mnx, mny, mnz, 0.f);
        boxes[2 * (size_t)g + 1] = make_float4(mxx, mxy, mxz, 0.f);
    }
}

__device__ __forceinline__ void write_row(const TopK &t, const float *__restrict__ X, int i, int lo, int n, int k,
                                          int base, float xi, float yi, float zi, int64_t *__restrict__ ids_out,
                                          float *__restrict__ d_out, float *__restrict__ r_out, int lane) {
    int kk = min(k, n);
#pragma unroll
    for (int slot = 0; slot < 2; ++slot) {
        int pos = 2 * lane + slot;
        if (pos >= k) continue;
        unsigned long long key = slot ? t.k1 : t.k0;
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
                int32_t *__restrict__ seg_flag, int32_t *__restrict__ row_flag, const float4 *__restrict__ boxes) {
    int lane = threadIdx.x & 31;
    int i = blockIdx.x * KNN_WARPS + (threadIdx.x >> 5);
    if (i >= n_atoms) return;
    int s = find_segment(seg_off, n_seg, i);
    int lo = seg_off[s], hi = seg_off[s + 1], n = hi - lo;
    float xi = __ldg(X + 3 * (size_t)i), yi = __ldg(X + 3 * (size_t)i + 1), zi = __ldg(X + 3 * (size_t)i + 2);
    TopK t;
    topk_init(t);
    float rowmax = 0.f;
    scan_segment_pruned(t, rowmax, X, boxes, lo, hi, i, xi, yi, zi, lane);
#pragma unroll
    for (int o = 16; o; o >>= 1) rowmax = fmaxf(rowmax, __shfl_xor_sync(0xffffffffu, rowmax, o));
    int kk = min(min(k, KMAX), n);
    // key at sorted position kk-1
    unsigned long long last = __shfl_sync(0xffffffffu, ((kk - 1) & 1) ? t.k1 : t.k0, (kk - 1) >> 1);
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

// scratch: [chunk boxes: 2 float4 per 32 atoms][seg_flag n_seg][seg_maxd n_seg][row_flag n_atoms]
static size_t knn_box_bytes(int n_atoms) { return (size_t)((n_atoms + 31) / 32) * 2 * sizeof(float4); }

size_t knn_scratch_bytes(int n_atoms, int n_seg) {
    return knn_box_bytes(n_atoms) + ((size_t)n_seg * 2 + (size_t)n_atoms) * sizeof(int32_t) + 256;
}

int launch_knn(const float *X, int n_atoms, const int32_t *seg_off, int n_seg, int k, int base, int64_t *ids_out,
               float *d_out, float *r_out, void *scratch, cudaStream_t st) {
    if (n_atoms <= 0) return PESTO_OK;
    if ((uintptr_t)scratch & 15) {
        set_error("pesto_knn: scratch must be 16-byte aligned");
        return PESTO_EINVAL;
    }
    float4 *boxes = (float4 *)scratch;
    int32_t *seg_flag = (int32_t *)((char *)scratch + knn_box_bytes(n_atoms));
    unsigned *seg_maxd = (unsigned *)(seg_flag + n_seg);
    int32_t *row_flag = (int32_t *)(seg_maxd + n_seg);
    PESTO_CUDA(cudaMemsetAsync(seg_flag, 0, (size_t)n_seg * 2 * sizeof(int32_t), st));
    int grid = (n_atoms + KNN_WARPS - 1) / KNN_WARPS;
    int n_chunks = (n_atoms + 31) / 32;
    knn_boxes_kernel<<<(n_chunks + KNN_WARPS - 1) / KNN_WARPS, KNN_WARPS * 32, 0, st>>>(X, n_atoms, boxes);
    knn_main_kernel<<<grid, KNN_WARPS * 32, 0, st>>>(X, n_atoms, seg_off, n_seg, k, base, ids_out, d_out, r_out,
                                                     seg_flag, row_flag, boxes);
    knn_segmax_kernel<<<grid, KNN_WARPS * 32, 0, st>>>(X, n_atoms, seg_off, n_seg, seg_flag, seg_maxd);
    knn_exact_rows_kernel<<<grid, KNN_WARPS * 32, 0, st>>>(X, n_atoms, seg_off, n_seg, k, base, ids_out, d_out, r_out,
                                                           row_flag, seg_maxd);
    PESTO_CUDA(cudaGetLastError());
    return PESTO_OK;
}

}  // namespace pesto
