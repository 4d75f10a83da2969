// Pair generation on the device: which views become reference views and which neighbours each one is
// matched against - the callers' side of the path, O(V^2) on the flattened 4x4 world-to-camera poses.
//   ldp_kcenters_kernel    reference core/selection.py:36-54  select_cameras_kcenters (numpy, float32)
//   ldp_knn_kernel         reference core/selection.py:57-70  nearest_neighbors (torch.cdist + topk)
// k-centres mirrors numpy's float32 arithmetic operation by operation (sequential column sums for mean / std, the
// 8-accumulator pairwise row sum of np.linalg.norm over 16 columns, correctly rounded sqrt and division), so the greedy
// sequence of arg-max picks is numpy's.  Included by ldp_api.cu (unity build).
#include "ldp_device.cuh"

namespace ldp {

constexpr int KC_THREADS = 1024;
constexpr int KC_MAX_PER_THREAD = 8;      // views per thread: V <= 8192
constexpr int KC_DIM = 16;

// np.add.reduce over 16 contiguous float32: r[j] = a[j] + a[8 + j], then ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7))
// np.einsum("nd,nd->n") over 16 float32 products, as numpy's x86-64 wheels compute it (einsum_sumprod.c.src,
// sum_of_products_contig_contig_outstride0_two; the einsum loops are built for the baseline SIMD width, 4 lanes, multiply then
// add): lane j accumulates p[12+j], p[8+j], p[4+j], p[j] in that order, then (l0 + l1) + (l2 + l3).  Measured against np.einsum
// (20 000 random rows: identical).  It decides the first k-centres pick when row norms tie, as they do on symmetric rings.
__device__ __forceinline__ float einsum_row16(const float (&p)[KC_DIM]) {
    float l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) l[j] = __fadd_rn(__fadd_rn(__fadd_rn(p[12 + j], p[8 + j]), p[4 + j]), p[j]);
    return __fadd_rn(__fadd_rn(l[0], l[1]), __fadd_rn(l[2], l[3]));
}
__device__ __forceinline__ float pairwise16(const float (&a)[KC_DIM]) {
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(a[j], a[8 + j]);
    return __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])), __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
}

// first index wins among equal values (np.argmax)
__device__ __forceinline__ void argmax_merge(float& v, int& i, float ov, int oi) {
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
}

// SMEM: the poses live in shared memory (rows padded to 17 floats: conflict-free row reads), n * 68 bytes; otherwise the
// normalised rows go to the caller's scratch in global memory.
constexpr int KC_PAD = KC_DIM + 1;
template <bool SMEM>
__global__ void __launch_bounds__(KC_THREADS, 1)
ldp_kcenters_kernel(const float* __restrict__ X, int n, int k, float* __restrict__ Xn, int32_t* __restrict__ out_sorted,
                    int32_t* __restrict__ out_order)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* xs = reinterpret_cast<float*>(smem_raw);                  // [n][KC_PAD] when SMEM
    __shared__ float s_mu[KC_DIM], s_sigma[KC_DIM], s_c[KC_DIM];
    __shared__ float s_v[32];
    __shared__ int s_i[32];
    __shared__ int s_pick;
    __shared__ int red_i[32];
    grid_dependency_sync();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rs = SMEM ? KC_PAD : KC_DIM;                           // row stride of the working copy
    float* R = SMEM ? xs : Xn;
    if (SMEM) {
        for (int e = tid; e < n * KC_DIM; e += KC_THREADS) xs[(e >> 4) * KC_PAD + (e & 15)] = X[e];
        __syncthreads();
    }
    const float* src = SMEM ? xs : X;
    // ---- mu = X.mean(axis=0), sigma = X.std(axis=0) + 1e-8: column sums run over the rows in order (numpy reduces the
    // outer axis of a C-contiguous array row by row)
    if (tid < KC_DIM) {
        float s = 0.f;
#pragma unroll 8
        for (int i = 0; i < n; ++i) s = __fadd_rn(s, src[(size_t)i * rs + tid]);
        const float mu = __fdiv_rn(s, (float)n);
        float s2 = 0.f;
#pragma unroll 8
        for (int i = 0; i < n; ++i) {
            const float d = __fsub_rn(src[(size_t)i * rs + tid], mu);
            s2 = __fadd_rn(s2, __fmul_rn(d, d));
        }
        s_mu[tid] = mu;
        s_sigma[tid] = __fadd_rn(__fsqrt_rn(__fdiv_rn(s2, (float)n)), 1e-8f);
    }
    __syncthreads();
    // ---- Xn = (X - mu) / sigma; first = argmax of the squared row norms
    float dist[KC_MAX_PER_THREAD];
    float bv = -INFINITY;
    int bi = 0x7fffffff;
#pragma unroll
    for (int u = 0; u < KC_MAX_PER_THREAD; ++u) {
        const int i = tid + u * KC_THREADS;
        dist[u] = -INFINITY;
        if (i < n) {
            float a[KC_DIM];
#pragma unroll
            for (int j = 0; j < KC_DIM; ++j) {
                const float v = __fdiv_rn(__fsub_rn(src[(size_t)i * rs + j], s_mu[j]), s_sigma[j]);
                R[(size_t)i * rs + j] = v;                            // a thread rewrites only its own rows
                a[j] = __fmul_rn(v, v);
            }
            argmax_merge(bv, bi, einsum_row16(a), i);
        }
    }
    auto block_argmax = [&](float v, int i) -> int {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, v, o);
            const int oi = __shfl_xor_sync(0xffffffffu, i, o);
            argmax_merge(v, i, ov, oi);
        }
        if (lane == 0) { s_v[warp] = v; s_i[warp] = i; }
        __syncthreads();
        if (warp == 0) {
            v = s_v[lane]; i = s_i[lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, v, o);
                const int oi = __shfl_xor_sync(0xffffffffu, i, o);
                argmax_merge(v, i, ov, oi);
            }
            if (lane == 0) s_pick = i;
        }
        __syncthreads();
        return s_pick;
    };
    int c = block_argmax(bv, bi);                                     // (its barriers also publish the normalised rows)
    // ---- greedy k-centres: dist = min(dist, |Xn - Xn[c]|), picked views drop to -inf
    for (int it = 0; it < k; ++it) {
        if (tid == 0) out_order[it] = c;
        if (tid < KC_DIM) s_c[tid] = R[(size_t)c * rs + tid];
        __syncthreads();
        bv = -INFINITY; bi = 0x7fffffff;
#pragma unroll
        for (int u = 0; u < KC_MAX_PER_THREAD; ++u) {
            const int i = tid + u * KC_THREADS;
            if (i < n) {
                float a[KC_DIM];
#pragma unroll
                for (int j = 0; j < KC_DIM; ++j) {
                    const float d = __fsub_rn(R[(size_t)i * rs + j], s_c[j]);
                    a[j] = __fmul_rn(d, d);
                }
                const float d = __fsqrt_rn(pairwise16(a));
                float cur = (it == 0) ? d : ((d < dist[u] || d != d) ? d : dist[u]);      // np.minimum (NaN propagates)
                if (i == c) cur = -INFINITY;
                if (dist[u] == -INFINITY && it > 0) cur = -INFINITY;                      // picked earlier
                dist[u] = cur;
                argmax_merge(bv, bi, cur, i);
            }
        }
        if (it + 1 < k) c = block_argmax(bv, bi);
    }
    __syncthreads();
    // ---- sorted(centers): ordered compaction of the picked views (dist == -inf)
    int written = 0;
#pragma unroll
    for (int u = 0; u < KC_MAX_PER_THREAD; ++u) {
        if (u * KC_THREADS >= n) break;                                // uniform
        const int i = u * KC_THREADS + tid;
        const int flag = (i < n && dist[u] == -INFINITY) ? 1 : 0;
        int total;
        const int ex = block_exclusive_scan(flag, red_i, &total);
        if (flag) out_sorted[written + ex] = i;
        written += total;
        __syncthreads();
    }
}

// One CTA per view: distances to every view, then torch.topk's own selection of the k smallest.
// The distances are torch.cdist's float32 values (reference core/selection.py:66; ATen _euclidean_dist, used above 25 rows):
//   d(i, j) = sqrt(max(x1_[i] . x2_[j], 0)),  x1_ = [-2 x, |x|^2, 1],  x2_ = [x, 1, |x|^2]
// with the K = 18 dot product accumulated as a chain of float32 FMAs in index order (what the sgemm does) and the row norms
// summed as eight lanes (a[i] + a[i + 8]) added left to right -- both orders measured against torch 2.11 (oracle:
// cdist_squared_f32).  The cancellation error of this formula (~1e-6 on poses of norm ~10) decides the order of the left /
// right neighbours of a ring camera, so it is reproduced, not avoided; the diagonal is +inf (dist.fill_diagonal_).
// Where float32 distances are EXACTLY equal - a third of the rows of a ring scene - the reference's order is whatever
// torch.topk's CPU kernel leaves: (value, index) pairs, a value-only comparator (NaN last) and, for k * 64 <= n,
// std::partial_sort(begin, begin + k, end), else std::nth_element(begin, begin + k - 1, end) + std::sort(begin, begin + k - 1).
// Those are deterministic sequences of moves; libstdc++'s are restated below one for one (oracle: topk_smallest_like_torch,
// pinned against torch.topk and the reference's frozen tables), run by one thread per view on the row in shared memory.
constexpr int KNN_MAX_K = 16;
constexpr int KNN_THREADS = 128;
constexpr int KNN_CHUNK = 2048;                 // heap path: distances staged per chunk (any n)
constexpr int KNN_SELECT_MAX = 64 * KNN_MAX_K;  // selection path: n < 64 k <= 1024 pairs in shared memory
struct KnnPair { float v; int i; };
__device__ __forceinline__ bool knn_lt(const KnnPair& a, const KnnPair& b) {
    return ((a.v == a.v) && (b.v != b.v)) || a.v < b.v;
}
// bits/stl_heap.h: __push_heap, __adjust_heap, __make_heap, __pop_heap; bits/stl_algo.h: __heap_select
__device__ void knn_adjust_heap(KnnPair* v, int hole, int len, KnnPair value) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (knn_lt(v[child], v[child - 1])) --child;
        v[hole] = v[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        v[hole] = v[child - 1];
        hole = child - 1;
    }
    int parent = (hole - 1) / 2;
    while (hole > top && knn_lt(v[parent], value)) {
        v[hole] = v[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    v[hole] = value;
}
__device__ void knn_make_heap(KnnPair* v, int len) {
    if (len < 2) return;
    for (int parent = (len - 2) / 2; ; --parent) {
        knn_adjust_heap(v, parent, len, v[parent]);
        if (parent == 0) return;
    }
}
__device__ void knn_sort_heap(KnnPair* v, int len) {
    while (len > 1) {
        --len;
        const KnnPair value = v[len];
        v[len] = v[0];
        knn_adjust_heap(v, 0, len, value);
    }
}
__device__ void knn_heap_select(KnnPair* v, int middle, int last) {      // [0, middle) heap over [0, last)
    knn_make_heap(v, middle);
    for (int i = middle; i < last; ++i) {
        if (knn_lt(v[i], v[0])) {
            const KnnPair value = v[i];
            v[i] = v[0];
            knn_adjust_heap(v, 0, middle, value);
        }
    }
}
// bits/stl_algo.h: __move_median_to_first + __unguarded_partition (pivot at first), __insertion_sort, __introselect
__device__ int knn_partition_pivot(KnnPair* v, int first, int last) {
    const int a = first + 1, b = first + (last - first) / 2, c = last - 1;
    int m;
    if (knn_lt(v[a], v[b])) m = knn_lt(v[b], v[c]) ? b : (knn_lt(v[a], v[c]) ? c : a);
    else m = knn_lt(v[a], v[c]) ? a : (knn_lt(v[b], v[c]) ? c : b);
    { const KnnPair t = v[first]; v[first] = v[m]; v[m] = t; }
    int lo = first + 1, hi = last;
    for (;;) {
        while (knn_lt(v[lo], v[first])) ++lo;
        --hi;
        while (knn_lt(v[first], v[hi])) --hi;
        if (!(lo < hi)) return lo;
        const KnnPair t = v[lo]; v[lo] = v[hi]; v[hi] = t;
        ++lo;
    }
}
__device__ void knn_insertion_sort(KnnPair* v, int first, int last) {
    for (int i = first + 1; i < last; ++i) {
        const KnnPair val = v[i];
        if (knn_lt(val, v[first])) {
            for (int q = i; q > first; --q) v[q] = v[q - 1];          // std::move_backward
            v[first] = val;
        } else {                                                     // __unguarded_linear_insert
            int q = i;
            while (knn_lt(val, v[q - 1])) { v[q] = v[q - 1]; --q; }
            v[q] = val;
        }
    }
}
__device__ void knn_nth_element(KnnPair* v, int nth, int n) {
    int first = 0, last = n;
    if (first == last || nth == last) return;
    int depth = (31 - __clz(n)) * 2;
    while (last - first > 3) {
        if (depth == 0) {
            knn_heap_select(v + first, nth + 1 - first, last - first);
            const KnnPair t = v[first]; v[first] = v[nth]; v[nth] = t;
            return;
        }
        --depth;
        const int cut = knn_partition_pivot(v, first, last);
        if (cut <= nth) first = cut; else last = cut;
    }
    knn_insertion_sort(v, first, last);
}

__device__ __forceinline__ float cdist_row_norm16(const float* __restrict__ x) {
    float l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) l[i] = __fadd_rn(__fmul_rn(x[i], x[i]), __fmul_rn(x[i + 8], x[i + 8]));
    float s = l[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) s = __fadd_rn(s, l[i]);
    return s;
}
__device__ __forceinline__ float cdist_value(const float* xi_m2, float ni, const float* __restrict__ X, int j) {
    float xj[KC_DIM];
#pragma unroll
    for (int c = 0; c < KC_DIM; ++c) xj[c] = X[(size_t)j * KC_DIM + c];
    const float nj = cdist_row_norm16(xj);
    float acc = __fmul_rn(xi_m2[0], xj[0]);                        // fma(a, b, 0)
#pragma unroll
    for (int c = 1; c < KC_DIM; ++c) acc = __fmaf_rn(xi_m2[c], xj[c], acc);
    acc = __fmaf_rn(ni, 1.0f, acc);
    acc = __fmaf_rn(1.0f, nj, acc);
    return __fsqrt_rn(fmaxf(acc, 0.0f));                           // clamp_min_(0).sqrt_()
}
// Up to 25 views torch.cdist does not go through the matrix product: cdist_impl takes the mm formulation only for r1 > 25 || r2 > 25
// and otherwise runs the direct kernel (ATen DistanceOpsKernel.cpp, tdist_calc): a sequential float32 sum over the columns of
// (a - b) * (a - b), multiply then add, then sqrt (oracle: cdist_f32, measured against torch).
constexpr int CDIST_MM_ABOVE = 25;
__device__ __forceinline__ float cdist_value_direct(const float* xi, const float* __restrict__ X, int j) {
    float agg = 0.f;
#pragma unroll
    for (int c = 0; c < KC_DIM; ++c) {
        const float d = fabsf(__fsub_rn(xi[c], X[(size_t)j * KC_DIM + c]));
        agg = __fadd_rn(agg, __fmul_rn(d, d));
    }
    return __fsqrt_rn(agg);
}
__global__ void __launch_bounds__(KNN_THREADS)
ldp_knn_kernel(const float* __restrict__ X, int n, int k, long long* __restrict__ out)
{
    __shared__ KnnPair s_pair[KNN_SELECT_MAX];                     // selection path: the whole row; heap path: the heap
    __shared__ float s_val[KNN_CHUNK];
    grid_dependency_sync();
    const int row = blockIdx.x, tid = threadIdx.x;
    float xi[KC_DIM];
#pragma unroll
    for (int j = 0; j < KC_DIM; ++j) xi[j] = X[(size_t)row * KC_DIM + j];
    const float ni = cdist_row_norm16(xi);
    const bool direct = n <= CDIST_MM_ABOVE;                               // (such a row never takes the heap path: k * 64 > n)
    if (!direct) {
#pragma unroll
        for (int j = 0; j < KC_DIM; ++j) xi[j] = __fmul_rn(xi[j], -2.0f);  // x1.mul(-2): exact
    }
    if ((long long)k * 64 <= (long long)n) {
        // std::partial_sort: only an element smaller than the heap's top moves anything, so the row is streamed
        for (int base = 0; base < n; base += KNN_CHUNK) {
            const int cnt = min(KNN_CHUNK, n - base);
            for (int q = tid; q < cnt; q += KNN_THREADS) s_val[q] = (base + q == row) ? INFINITY : cdist_value(xi, ni, X, base + q);
            __syncthreads();
            if (tid == 0) {
                int q = 0;
                if (base == 0) {
                    for (; q < k; ++q) { s_pair[q].v = s_val[q]; s_pair[q].i = q; }        // k <= 16 < KNN_CHUNK
                    knn_make_heap(s_pair, k);
                }
                for (; q < cnt; ++q) {
                    KnnPair e; e.v = s_val[q]; e.i = base + q;
                    if (knn_lt(e, s_pair[0])) knn_adjust_heap(s_pair, 0, k, e);            // __pop_heap(first, middle, i)
                }
            }
            __syncthreads();
        }
        if (tid == 0) knn_sort_heap(s_pair, k);
    } else {                                                       // n < 64 k: the pairs fit
        for (int j = tid; j < n; j += KNN_THREADS) {
            s_pair[j].v = (j == row) ? INFINITY : (direct ? cdist_value_direct(xi, X, j) : cdist_value(xi, ni, X, j));
            s_pair[j].i = j;
        }
        __syncthreads();
        if (tid == 0) {
            knn_nth_element(s_pair, k - 1, n);
            knn_insertion_sort(s_pair, 0, k - 1);                  // std::sort of k - 1 <= 15 elements is its insertion sort
        }
    }
    __syncthreads();
    if (tid < k) out[(size_t)row * k + tid] = (long long)s_pair[tid].i;
}

}  // namespace ldp
