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
            argmax_merge(bv, bi, pairwise16(a), i);
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

// One warp per view: distances to every other view, the k smallest in ascending order (ties: lower index first).
// The distances are torch.cdist's own float32 values (reference core/selection.py:66; ATen _euclidean_dist, used above 25 rows):
//   d(i, j) = sqrt(max(x1_[i] . x2_[j], 0)),  x1_ = [-2 x, |x|^2, 1],  x2_ = [x, 1, |x|^2]
// with the K = 18 dot product accumulated as a chain of float32 FMAs in index order (what the sgemm does) and the row norms
// summed as eight lanes (a[i] + a[i + 8]) added left to right -- both orders measured against torch 2.11 (oracle:
// cdist_squared_f32).  The cancellation error of this formula (~1e-6 on poses of norm ~10) is what decides the order of the
// left / right neighbours of a ring camera, so it is reproduced, not avoided.  Where two float32 distances are EXACTLY equal
// the reference's order is whatever std::nth_element / std::partial_sort leave (torch.topk); here the lower index comes first.
constexpr int KNN_MAX_K = 16;
__device__ __forceinline__ float cdist_row_norm16(const float* __restrict__ x) {
    float l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) l[i] = __fadd_rn(__fmul_rn(x[i], x[i]), __fmul_rn(x[i + 8], x[i + 8]));
    float s = l[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) s = __fadd_rn(s, l[i]);
    return s;
}
__global__ void __launch_bounds__(256)
ldp_knn_kernel(const float* __restrict__ X, int n, int k, long long* __restrict__ out)
{
    grid_dependency_sync();
    const int lane = threadIdx.x & 31, row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    float xi[KC_DIM];
#pragma unroll
    for (int j = 0; j < KC_DIM; ++j) xi[j] = X[(size_t)row * KC_DIM + j];
    const float ni = cdist_row_norm16(xi);
#pragma unroll
    for (int j = 0; j < KC_DIM; ++j) xi[j] = __fmul_rn(xi[j], -2.0f);      // x1.mul(-2): exact
    float bv[KNN_MAX_K];
    int bi[KNN_MAX_K];
#pragma unroll
    for (int t = 0; t < KNN_MAX_K; ++t) { bv[t] = INFINITY; bi[t] = 0x7fffffff; }
    for (int j = lane; j < n; j += 32) {
        if (j == row) continue;                                    // dist.fill_diagonal_(inf)
        float xj[KC_DIM];
#pragma unroll
        for (int c = 0; c < KC_DIM; ++c) xj[c] = X[(size_t)j * KC_DIM + c];
        const float nj = cdist_row_norm16(xj);
        float acc = __fmul_rn(xi[0], xj[0]);                       // fma(a, b, 0)
#pragma unroll
        for (int c = 1; c < KC_DIM; ++c) acc = __fmaf_rn(xi[c], xj[c], acc);
        acc = __fmaf_rn(ni, 1.0f, acc);
        acc = __fmaf_rn(1.0f, nj, acc);
        float v = __fsqrt_rn(fmaxf(acc, 0.0f));                    // clamp_min_(0).sqrt_()
        int vi = j;
        // insert into this lane's ascending list (indices ascend within a lane, so ties keep the earlier one first)
#pragma unroll
        for (int t = 0; t < KNN_MAX_K; ++t) {
            if (t < k && v < bv[t]) { const float tv = bv[t]; const int ti = bi[t]; bv[t] = v; bi[t] = vi; v = tv; vi = ti; }
        }
    }
    // merge the 32 sorted lists: k rounds of a warp arg-min over the lanes' heads
    for (int t = 0; t < k; ++t) {
        float v = bv[0];
        int i = bi[0], who = lane;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, v, o);
            const int oi = __shfl_xor_sync(0xffffffffu, i, o), ow = __shfl_xor_sync(0xffffffffu, who, o);
            if (ov < v || (ov == v && oi < i)) { v = ov; i = oi; who = ow; }
        }
        if (lane == 0) out[(size_t)row * k + t] = (long long)i;
        if (lane == who) {                                          // pop the head
#pragma unroll
            for (int s = 0; s + 1 < KNN_MAX_K; ++s) { bv[s] = bv[s + 1]; bi[s] = bi[s + 1]; }
            bv[KNN_MAX_K - 1] = INFINITY; bi[KNN_MAX_K - 1] = 0x7fffffff;
        }
    }
}

}  // namespace ldp
