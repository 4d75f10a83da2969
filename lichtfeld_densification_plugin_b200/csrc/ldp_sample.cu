// K1: certainty-weighted + coverage correspondence sampling, one CTA per reference view.
//
// Replaces, for every reference view of the launch,
//   * the per-pixel best neighbour           reference core/pipeline.py:634-635   (torch.max over neighbours)
//   * select_samples_with_coverage           reference core/sampling.py:8-53
//       - cap / border mask / f32 normalise  :12-14,23-29
//       - np.random.choice(replace=False, p) :31-32  == numpy legacy RandomState.choice: rejection rounds of
//         inverse-CDF draws on a sequential f64 cumsum (numpy/random/mtrand.pyx), restated in
//         oracle/densify_oracle.py:legacy_choice_no_replace
//       - per-tile best pixel ("coverage")   :34-50
//       - np.unique(concat)                  :52
//
// Data flow per view (N = H*W pixels, nn neighbours):
//   phase 1  stream nn certainty planes once (float4, evict-first)  -> w[N] f32, bestk[N] u8 in the
//            L2-resident workspace; block f64 sum -> s (or the caller's override)
//   phase 2  re-read w (L2): p = fl32(w/s) as f64, per-chunk sums in SHARED memory, per-tile arg-max
//   rounds   in-place f64 prefix over the chunk table (shared) -> per draw: binary search in shared
//            memory, then one 128-byte read of the chunk's weights and a <=chunk-long sequential scan
//            that reproduces numpy's `searchsorted(cdf/cdf[-1], u, 'right')` comparison exactly
//            (__ddiv_rn); first-occurrence dedupe by atomicOr on a bitmap; found pixels are zeroed
//            and their mass subtracted from the chunk table (exact, see DESIGN.md "exact f64 sums").
//   finish   coverage picks OR-ed into the bitmap; ordered bitmap compaction = sorted unique sel_idx.
#include "ldp_device.cuh"

namespace ldp {

constexpr int K1_THREADS = 1024;

struct K1Shared {
    double red_d[32];
    int red_i[32];
    int n_found;
    int n_new_base;
    int status;
    int flags;
};

__device__ __forceinline__ unsigned long long cov_key(float p, int idx) {
    return ((unsigned long long)__float_as_uint(p) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)idx);
}

// descending bitonic sort of n (power of two) 64-bit keys in shared memory
__device__ void bitonic_sort_desc(unsigned long long* a, int n) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const int l = i ^ j;
                if (l > i) {
                    const unsigned long long x = a[i], y = a[l];
                    const bool desc = ((i & k) == 0);
                    if (desc ? (x < y) : (x > y)) { a[i] = y; a[l] = x; }
                }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(K1_THREADS, 1)
ldp_sample_kernel(const ldp_params P, const ldp_ref_desc* __restrict__ refs, const double* __restrict__ uniforms,
                  const Workspace ws, const ldp_outputs out, const SampleGeom G)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* csum = reinterpret_cast<double*>(smem_raw);                                  // [nchunk]
    unsigned long long* bins = reinterpret_cast<unsigned long long*>(csum + G.nchunk);   // [nbins_pow2]
    __shared__ K1Shared sh;
    __shared__ const float* s_cert[LDP_MAX_NN];

    const int r = blockIdx.x;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int T = blockDim.x;
    const int N = G.N;
    const ldp_ref_desc* rd = refs + r;
    const int nn = rd->nn;

    float* __restrict__ w = ws.w + (size_t)r * ws.n_pad;
    uint8_t* __restrict__ bk = ws.bestk + (size_t)r * ws.n_pad;
    uint32_t* __restrict__ bitmap = ws.bitmap + (size_t)r * ws.n_words;
    int32_t* __restrict__ found = ws.found + (size_t)r * ws.found_cap;
    int32_t* __restrict__ sel = (out.sel_idx ? out.sel_idx : ws.sel) + (size_t)r * ws.sel_cap;

    if (tid < LDP_MAX_NN) s_cert[tid] = (tid < nn) ? rd->cert[tid] : nullptr;
    if (tid == 0) { sh.n_found = 0; sh.n_new_base = 0; sh.status = LDP_REF_OK; sh.flags = 0; ws.kept[r] = 0; }
    for (int i = tid; i < (int)ws.n_words; i += T) bitmap[i] = 0u;
    for (int i = tid; i < G.nchunk; i += T) csum[i] = 0.0;
    int nb_pow2 = 1;
    while (nb_pow2 < G.nbins) nb_pow2 <<= 1;
    for (int i = tid; i < nb_pow2; i += T) bins[i] = 0ull;
    __syncthreads();

    auto finish_empty = [&](int status) {
        if (tid == 0) {
            out.status[r] = status;
            out.n_samples[r] = 0;
            if (out.uniforms_used) out.uniforms_used[r] = 0;
            if (out.rounds) out.rounds[r] = 0;
        }
    };
    if (nn <= 0) { finish_empty(LDP_REF_NO_NEIGHBOURS); return; }

    // ------------------------------------------------------------------ phase 1: stream certainties
    const float cap = P.sample_cap;
    const int W = P.W, H = P.H, border = P.border;
    const int nquad = (N + 3) >> 2;
    double lsum = 0.0;
    int lbad = 0;
    for (int q = tid; q < nquad; q += T) {
        const int px = q << 2;
        float best[4];
        int bi[4] = {0, 0, 0, 0};
        if (G.vec) {
            float4 v = ld_stream4(s_cert[0] + px);
            best[0] = v.x; best[1] = v.y; best[2] = v.z; best[3] = v.w;
#pragma unroll 4
            for (int k = 1; k < nn; ++k) {
                const float4 c = ld_stream4(s_cert[k] + px);
                if (c.x > best[0]) { best[0] = c.x; bi[0] = k; }
                if (c.y > best[1]) { best[1] = c.y; bi[1] = k; }
                if (c.z > best[2]) { best[2] = c.z; bi[2] = k; }
                if (c.w > best[3]) { best[3] = c.w; bi[3] = k; }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) best[j] = (px + j < N) ? __ldcs(s_cert[0] + px + j) : 0.f;
            for (int k = 1; k < nn; ++k) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float c = (px + j < N) ? __ldcs(s_cert[k] + px + j) : 0.f;
                    if (c > best[j]) { best[j] = c; bi[j] = k; }
                }
            }
        }
        int y = px / W, x = px - y * W;
        float wv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float c = (best[j] > cap) ? cap : best[j];          // torch.clamp(max=cap): NaN stays NaN
            float m = 1.f;
            if (!P.no_filter)
                m = (x >= border && x <= W - 1 - border && y >= border && y <= H - 1 - border) ? 1.f : 0.f;
            float v = c * m;
            if (px + j >= N) v = 0.f;
            wv[j] = v;
            lbad |= (v != v) ? 1 : 0;
            lbad |= (v < 0.f) ? 2 : 0;
            lsum += (double)v;
            if (++x == W) { x = 0; ++y; }
        }
        *reinterpret_cast<float4*>(w + px) = make_float4(wv[0], wv[1], wv[2], wv[3]);
        *reinterpret_cast<uchar4*>(bk + px) = make_uchar4((unsigned char)bi[0], (unsigned char)bi[1],
                                                          (unsigned char)bi[2], (unsigned char)bi[3]);
    }
    const double wsum64 = block_sum(lsum, sh.red_d);
    const int bad = block_sum(lbad ? ((lbad & 1) | ((lbad & 2) << 15)) : 0, sh.red_i);   // low half: NaN count, high: negatives
    if (P.no_filter) return;   // top-M selection is done by ldp_topm_kernel

    float s = (float)wsum64;
    if (rd->weight_sum_override > 0.f) s = rd->weight_sum_override;
    if (tid == 0 && out.weight_sum) out.weight_sum[r] = s;
    if (bad & 0xffff) { finish_empty(LDP_REF_BAD_WEIGHTS); return; }       // NaN: `s <= 0` is False, choice raises
    if (!(s > 0.f)) { finish_empty(LDP_REF_EMPTY); return; }               // core/sampling.py:27-28
    if (bad >> 16) { finish_empty(LDP_REF_BAD_WEIGHTS); return; }

    // ------------------------------------------------------------------ phase 2: chunk sums of p, tile arg-max
    // (the workspace writes of phase 1 are visible after the barriers inside block_sum)
    const int cs = G.chunk_shift;
    const int gl = min(32, (1 << cs) >> 2);            // lanes that share one chunk
    double ltot = 0.0;
    int lpos = 0;
    int lemin = 0x7fffffff;
    const int nunit = (N + 127) >> 7;                  // 128 pixels per warp-iteration
    for (int u = (tid >> 5); u < nunit; u += (T >> 5)) {
        const int px = (u << 7) + (lane << 2);
        double a = 0.0;
        if (px < N) {
            const float4 v = *reinterpret_cast<const float4*>(w + px);
            const float wv[4] = {v.x, v.y, v.z, v.w};
            int y = px / W, x = px - y * W;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float p = __fdiv_rn(wv[j], s);                     // core/sampling.py:29 (f32 division)
                if (p > 0.f) {
                    a += (double)p;
                    ++lpos;
                    const int e = (int)((__float_as_uint(p) >> 23) & 0xffu);
                    lemin = min(lemin, e);
                    const int b = (x / G.tile) * G.nby + (y / G.tile);
                    const unsigned long long key = cov_key(p, px + j);
                    if (bins[b] < key) atomicMax(&bins[b], key);
                }
                if (++x == W) { x = 0; ++y; }
            }
        }
        ltot += a;
        for (int o = 1; o < gl; o <<= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (px < N && (lane & (gl - 1)) == 0) {
            if (cs <= 7) csum[px >> cs] = a;
            else atomicAdd(&csum[px >> cs], a);
        }
    }
    const double ptotal = block_sum(ltot, sh.red_d);
    const int npos = block_sum(lpos, sh.red_i);
    const int emin = block_min(lemin, sh.red_i);
    // numpy's checks in RandomState.choice, in numpy's order
    if (fabs(ptotal - 1.0) > 3.4526698300124393e-4) { finish_empty(LDP_REF_PSUM); return; }
    if (npos < G.size) { finish_empty(LDP_REF_FEWER_NONZERO); return; }
    // exactness of every f64 partial sum: all p are multiples of 2^(emin-150) and the total stays below
    // 2^53 of those units  <=>  ilogb(total) - (emin - 150) < 53
    int inexact = 0;
    if (npos > 0) {
        const int etot = ilogb(ptotal);
        if (emin == 0 || etot - (emin - 150) >= 53) inexact = 1;
    }

    // ------------------------------------------------------------------ rejection rounds (numpy legacy choice)
    const int size = G.size;
    int n_have = 0;
    int drawn = 0;
    int rounds = 0;
    const double* U = uniforms ? uniforms + (size_t)r * (size_t)P.uniforms_per_ref : nullptr;
    const int L = (G.nchunk + T - 1) / T;
    const int seg0 = min(tid * L, G.nchunk), seg1 = min(seg0 + L, G.nchunk);
    int fail = 0;
    while (n_have < size) {
        const int cnt = size - n_have;
        if (P.rng_mode == LDP_RNG_EXPLICIT && (int64_t)drawn + cnt > P.uniforms_per_ref) { fail = LDP_REF_UNIFORMS_EXHAUSTED; break; }
        if (rounds >= 64) { fail = LDP_REF_ROUNDS_EXCEEDED; break; }
        // (a) in-place inclusive prefix of the chunk sums
        double loc = 0.0;
        for (int i = seg0; i < seg1; ++i) loc += csum[i];
        double total;
        double run = block_exclusive_scan(loc, sh.red_d, &total);
        for (int i = seg0; i < seg1; ++i) { run += csum[i]; csum[i] = run; }
        __syncthreads();
        total = csum[G.nchunk - 1];
        // (b) draws
        for (int d = tid; d < cnt; d += T) {
            const double u = (P.rng_mode == LDP_RNG_EXPLICIT) ? U[drawn + d]
                                                              : philox_uniform(P.seed, rd->rng_stream, (uint32_t)(drawn + d));
            const double t = u * total;
            int lo = 0, hi = G.nchunk - 1;               // first chunk with prefix > t (approximate)
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (csum[mid] > t) hi = mid; else lo = mid + 1;
            }
            int c = lo;
            // exact predicate of searchsorted(cdf, u, 'right'): fl(prefix / total) > u
            while (c > 0 && __ddiv_rn(csum[c - 1], total) > u) --c;
            while (c < G.nchunk - 1 && !(__ddiv_rn(csum[c], total) > u)) ++c;
            const double tlo = t * (1.0 - 1.0 / 1125899906842624.0);
            double cum = (c > 0) ? csum[c - 1] : 0.0;
            int idx = -1, last = -1;
            for (; c < G.nchunk && idx < 0; ++c) {
                const int p0 = c << cs, p1 = min(p0 + (1 << cs), N);
                for (int b = p0; b < p1 && idx < 0; b += 32) {
                    float4 v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        v[j] = (b + 4 * j < p1) ? *reinterpret_cast<const float4*>(w + b + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float e4[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
                        for (int m = 0; m < 4; ++m) {
                            if (idx < 0 && e4[m] > 0.f) {
                                const float p = __fdiv_rn(e4[m], s);
                                if (p > 0.f) {
                                    cum += (double)p;
                                    last = b + 4 * j + m;
                                    if (cum > tlo && __ddiv_rn(cum, total) > u) idx = last;
                                }
                            }
                        }
                    }
                }
            }
            if (idx < 0) idx = last;                     // only reachable when sums are inexact
            if (idx >= 0) {
                const uint32_t bit = 1u << (idx & 31);
                const uint32_t old = atomicOr(&bitmap[idx >> 5], bit);
                if (!(old & bit)) {
                    const int pos = atomicAdd(&sh.n_found, 1);
                    found[pos] = idx;
                }
            }
        }
        __syncthreads();
        const int n_now = sh.n_found;
        drawn += cnt;
        ++rounds;
        // (c) back to chunk sums (adjacent difference, in place), remove the found mass, zero the weights
        double prev = (seg0 > 0 && seg0 < G.nchunk) ? csum[seg0 - 1] : 0.0;
        __syncthreads();
        for (int i = seg1 - 1; i >= seg0; --i) {
            const double below = (i > seg0) ? csum[i - 1] : prev;
            csum[i] = csum[i] - below;
        }
        __syncthreads();
        for (int e = n_have + tid; e < n_now; e += T) {
            const int idx = found[e];
            const float p = __fdiv_rn(w[idx], s);
            atomicAdd(&csum[idx >> cs], -(double)p);
            w[idx] = 0.f;
        }
        __threadfence_block();
        __syncthreads();
        n_have = n_now;
    }
    if (fail) { finish_empty(fail); return; }

    // ------------------------------------------------------------------ coverage picks (core/sampling.py:34-50)
    {
        int lp = 0;
        for (int i = tid; i < G.nbins; i += T) lp += (bins[i] != 0ull) ? 1 : 0;
        const int nb_pos = block_sum(lp, sh.red_i);
        if (nb_pos > G.cov_budget) {          // budget binds: keep the cov_budget best tiles (descending weight)
            bitonic_sort_desc(bins, nb_pow2);
        }
        const int take = min(nb_pos, G.cov_budget);
        const int lim = (nb_pos > G.cov_budget) ? take : G.nbins;
        for (int i = tid; i < lim; i += T) {
            const unsigned long long key = bins[i];
            if (key != 0ull) {
                const int idx = (int)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull));
                atomicOr(&bitmap[idx >> 5], 1u << (idx & 31));
            }
        }
        __syncthreads();
    }

    // ------------------------------------------------------------------ ordered compaction == np.unique(concat)
    {
        const int nwords = (N + 31) >> 5;
        const int Lw = (nwords + T - 1) / T;
        const int w0 = min(tid * Lw, nwords), w1 = min(w0 + Lw, nwords);
        int cntb = 0;
        for (int i = w0; i < w1; ++i) cntb += __popc(bitmap[i]);
        int totalS;
        int pos = block_exclusive_scan(cntb, sh.red_i, &totalS);
        for (int i = w0; i < w1; ++i) {
            uint32_t m = bitmap[i];
            while (m) {
                const int b = __ffs(m) - 1;
                m &= m - 1;
                if (pos < (int)ws.sel_cap) sel[pos] = (i << 5) + b;
                ++pos;
            }
        }
        if (tid == 0) {
            out.status[r] = LDP_REF_OK | (inexact ? LDP_REF_INEXACT_SCAN : 0);
            out.n_samples[r] = min(totalS, (int)ws.sel_cap);
            if (out.uniforms_used) out.uniforms_used[r] = drawn;
            if (out.rounds) out.rounds[r] = rounds;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// no_filter branch of the sampler (reference core/sampling.py:15-21): the M pixels of largest capped
// certainty, in descending order.  np.argsort is unstable, so the order among equal certainties is
// implementation-defined in the reference; here ties are broken by ascending pixel index.
// One CTA per view; runs after ldp_sample_kernel's phase 1 (w = capped certainty, no border mask).
//   1. 4-pass 8-bit radix select over the f32 keys -> value of the M-th largest, #strictly greater
//   2. ordered gather of {key > kth} U {first (M - #greater) pixels with key == kth}
//   3. bitonic sort of the M (key, ~index) pairs, descending
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t f32_order_key(float f) {
    if (f != f) return 0u;                      // numpy sorts NaN last in argsort(-x)
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__global__ void __launch_bounds__(K1_THREADS, 1)
ldp_topm_kernel(const ldp_params P, const ldp_ref_desc* __restrict__ refs, const Workspace ws,
                const ldp_outputs out, const SampleGeom G)
{
    __shared__ int hist[256];
    __shared__ int red_i[32];
    __shared__ uint32_t s_prefix;
    __shared__ int s_remaining;
    const int r = blockIdx.x, tid = threadIdx.x, T = blockDim.x, N = G.N;
    const float* __restrict__ w = ws.w + (size_t)r * ws.n_pad;
    int32_t* __restrict__ sel = (out.sel_idx ? out.sel_idx : ws.sel) + (size_t)r * ws.sel_cap;
    unsigned long long* keys = ws.topk_keys + (size_t)r * ws.topk_cap;
    const int M = min(P.matches_per_ref, N);
    if (refs[r].nn <= 0 || M <= 0) {
        if (tid == 0) { out.status[r] = (refs[r].nn <= 0) ? LDP_REF_NO_NEIGHBOURS : LDP_REF_EMPTY; out.n_samples[r] = 0; }
        return;
    }
    // 1. radix select (most significant byte first)
    if (tid == 0) { s_prefix = 0u; s_remaining = M; }
    __syncthreads();
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        for (int i = tid; i < 256; i += T) hist[i] = 0;
        __syncthreads();
        const uint32_t prefix = s_prefix;
        const uint32_t pmask = (pass == 0) ? 0u : (0xFFFFFFFFu << (shift + 8));
        for (int i = tid; i < N; i += T) {
            const uint32_t k = f32_order_key(w[i]);
            if ((k & pmask) == prefix) atomicAdd(&hist[(k >> shift) & 0xffu], 1);
        }
        __syncthreads();
        if (tid == 0) {
            int rem = s_remaining;
            int b = 255;
            for (; b > 0; --b) {
                if (hist[b] >= rem) break;
                rem -= hist[b];
            }
            s_prefix = prefix | ((uint32_t)b << shift);
            s_remaining = rem;            // how many of the elements equal-so-far are still needed
        }
        __syncthreads();
    }
    const uint32_t kth = s_prefix;
    const int need_eq = s_remaining;
    // 2. ordered gather
    const int Lp = (N + T - 1) / T;
    const int p0 = min(tid * Lp, N), p1 = min(p0 + Lp, N);
    int c_gt = 0, c_eq = 0;
    for (int i = p0; i < p1; ++i) {
        const uint32_t k = f32_order_key(w[i]);
        c_gt += (k > kth);
        c_eq += (k == kth);
    }
    int tot_gt, tot_eq;
    int o_gt = block_exclusive_scan(c_gt, red_i, &tot_gt);
    int o_eq = block_exclusive_scan(c_eq, red_i, &tot_eq);
    for (int i = p0; i < p1; ++i) {
        const uint32_t k = f32_order_key(w[i]);
        if (k > kth) { keys[o_gt++] = ((unsigned long long)k << 32) | (0xFFFFFFFFu - (uint32_t)i); }
        else if (k == kth) {
            if (o_eq < need_eq) keys[tot_gt + o_eq] = ((unsigned long long)k << 32) | (0xFFFFFFFFu - (uint32_t)i);
            ++o_eq;
        }
    }
    int n2 = 1;
    while (n2 < M) n2 <<= 1;
    for (int i = M + tid; i < n2; i += T) keys[i] = 0ull;
    __threadfence_block();
    __syncthreads();
    // 3. sort descending (global memory, L2-resident)
    bitonic_sort_desc(keys, n2);
    for (int i = tid; i < M; i += T) sel[i] = (int)(0xFFFFFFFFu - (uint32_t)(keys[i] & 0xFFFFFFFFull));
    if (tid == 0) {
        out.status[r] = LDP_REF_OK;
        out.n_samples[r] = M;
        if (out.uniforms_used) out.uniforms_used[r] = 0;
        if (out.rounds) out.rounds[r] = 0;
    }
}

}  // namespace ldp
