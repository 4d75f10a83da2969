// K_A  front: stream + normalise in ONE persistent kernel (replaces ldp_stream_kernel + ldp_prep_kernel on the vector path).
//
//   reference core/pipeline.py:634-635 (torch.max over neighbours), core/sampling.py:12-14,23-29 (cap, border mask, f32
//   weights, s = weights.sum(), p = weights / s), :34-50 (coverage walk == per-tile arg-max of p)
//
// Why one kernel: the normaliser s of a view is known only when ALL of its certainty values have been read, and p = w / s
// is what every later stage consumes.  Two kernels meant writing w (4 B/px), reading it back and writing p (another
// 8 B/px through L2) plus a grid-wide barrier in between.  Here a CTA parks the weights of the tiles it has streamed in
// shared memory, announces each tile on a per-view arrival counter, goes on streaming, and turns a parked tile into p the
// moment its view's normaliser is published -- so the arithmetic of the second half runs under the HBM stream of the first.
//
// Structure (one CTA = 512 threads, 2 CTAs per SM, tiles of 2048 pixels = one quad per thread, dynamic tickets):
//   * TMA: thread 0 issues one cp.async.bulk per neighbour plane and tile (8 KB each, L2 evict-first: read-once data) into
//     a 2-stage shared-memory ring; completion is signalled on an mbarrier per stage (complete_tx::bytes).  The loads of the
//     next tile are in flight while the current one is consumed and while older tiles are normalised.
//   * stream phase (per tile): per-pixel max over neighbours, first index on ties, NaN propagating; cap; border mask; f64
//     partial sum + NaN / negative flags; winning neighbour (u8) to the workspace; the weights stay in shared memory (park
//     ring of 4 tiles).  The tile's partial goes to its slot of partial[r][] (stored as ~bits, so that a zeroed slot means
//     "not there yet"), then arrive[r] += 1; the CTA that completes a view sums the slots in fixed order (deterministic s;
//     it simply re-reads a slot whose store is still in flight, so nobody needs a fence) and publishes (epoch, flags, s)
//     as ONE 64-bit word.
//   * normalise phase (per parked tile whose flag is up): p = fl32(w / s) (hoisted correctly-rounded reciprocal + two FMA
//     residual corrections == IEEE division, scratch/divcheck.cu), written ONCE to the workspace; f64 chunk sums by warp
//     shuffles; per-tile arg-max keys (p bits << 32 | ~index) in shared memory then global atomicMax; #positive, smallest
//     exponent (exactness test of the draw kernel).
//   No L2 round trip sits on the critical path of a tile: the ticket of a refill is requested one iteration before the TMA
//   that uses it is issued, the arrival counter's return value is looked at one iteration later, and the poll of the
//   oldest parked view's word is issued before the stream phase and read after it.
//   Deadlock freedom: a CTA only ever BLOCKS on a view's flag while it holds no ticket it has not announced.  When the park
//   ring is full and the oldest parked view is not complete, the tiles still sitting in the TMA stages are streamed IN
//   PLACE (the weights overwrite plane 0 of the stage, which then counts as a parked tile and is refilled only after it has
//   been normalised) and announced first.  Tickets are taken only for a stage that is about to be filled, never ahead.
//   So every ticket taken is announced without waiting for anybody, all views complete, and every wait ends - whatever
//   subset of the grid is resident (other kernels of other streams may hold SMs for a while).
#include "ldp_device.cuh"

namespace ldp {

constexpr int KF_THREADS = 512;
constexpr int KF_TILE = KF_THREADS * 4;      // pixels per tile
constexpr int KF_STAGES = 2;                 // TMA ring
constexpr int KF_PARK = 4;                   // parked tiles (weights waiting for their view's normaliser)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
// 1-D bulk copy global -> shared through the TMA unit, read-once data (L2 evict-first)
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}
__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

#ifdef LDP_FRONT_CLOCKS
#define FCLK_DECL long long fc_t = clock64(), fc_acc[12] = {0,0,0,0,0,0,0,0,0,0,0,0}
#define FCLK(slot) do { if (threadIdx.x == 0) { const long long n_ = clock64(); fc_acc[slot] += n_ - fc_t; fc_t = n_; } } while (0)
#define FCNT(slot) do { if (threadIdx.x == 0) fc_acc[slot] += 1; } while (0)
#define FCLK_FLUSH do { if (threadIdx.x == 0) for (int q_ = 0; q_ < 12; ++q_) atomicAdd(reinterpret_cast<unsigned long long*>(ws.dbgclk) + q_, (unsigned long long)fc_acc[q_]); } while (0)
#else
#define FCLK_DECL do { } while (0)
#define FCLK(slot) do { } while (0)
#define FCNT(slot) do { } while (0)
#define FCLK_FLUSH do { } while (0)
#endif

struct FrontCtl {                       // static shared memory of the front kernel
    unsigned long long full_bar[KF_STAGES];
    int tile[KF_STAGES];                // ticket loaded (or being loaded) into each stage
    int tile_nn[KF_STAGES];             // neighbour planes of that tile's view
    double red_d[KF_THREADS / 32];
    float red_f[KF_THREADS / 32];
    int red_i[KF_THREADS / 32];
    int red_j[KF_THREADS / 32];
    int ent_r[KF_PARK + KF_STAGES], ent_blk[KF_PARK + KF_STAGES], ent_loc[KF_PARK + KF_STAGES];   // FIFO of parked tiles: view, tile,
                                        // where the weights are (0..KF_PARK-1: park slot, KF_PARK + s: plane 0 of stage s)
    int last_r;                         // >= 0: this CTA completed that view (it sums the partials and publishes s)
    int ready;                          // how many of the oldest parked tiles (0..2) have their view's normaliser
    int bad, bad2;
    float s, s2;
};

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long front_word(int epoch, int bad, float s) {
    return ((unsigned long long)(uint32_t)epoch << 34) | ((unsigned long long)(bad & 3) << 32) | (unsigned long long)__float_as_uint(s);
}

// PRO: raw matcher planes -- clamp(min = certainty_floor) and x mask_a are applied as the values are read
//      (core/pipeline.py:405-417); launches with warped neighbour masks (mask_b) take the two-kernel path.
template <bool PRO>
__global__ void __launch_bounds__(KF_THREADS, 2)
ldp_front_kernel(const ldp_params P, const ldp_ref_desc* __restrict__ refs, const Workspace ws, const ldp_outputs out,
                 const SampleGeom G)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ FrontCtl ctl;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = G.N, W = P.W, H = P.H;
    const int tpv = (int)ws.nblk;                        // tiles per view
    const int total = tpv * P.n_refs;
    const int nn_stage = G.front_nn;                     // planes a stage holds
    float* stage_base = reinterpret_cast<float*>(smem_raw);                                   // [KF_STAGES][nn_stage][KF_TILE]
    float* park_base = stage_base + (size_t)KF_STAGES * nn_stage * KF_TILE;                   // [KF_PARK][KF_TILE]
    unsigned long long* lb = reinterpret_cast<unsigned long long*>(park_base + (size_t)KF_PARK * KF_TILE);   // [prep_lb_cap]
    // per-view plane pointers and neighbour counts, cached so that issuing a tile's copies never waits on a descriptor read
    const float** c_cert = reinterpret_cast<const float**>(lb + G.prep_lb_cap);                               // [n_refs][nn_stage]
    int* c_nn = reinterpret_cast<int*>(c_cert + (size_t)(G.front_cache ? P.n_refs * nn_stage : 0));           // [n_refs]
    const float cap = P.sample_cap;
    const int border = P.border;
    constexpr int KF_ENT = KF_PARK + KF_STAGES;
    uint64_t policy = 0;

    // ---- thread 0's private control state
    int pend_tk[KF_STAGES];                              // ticket requested for a stage whose TMA has not been issued yet
#pragma unroll
    for (int q = 0; q < KF_STAGES; ++q) pend_tk[q] = 0;
    int pend_old = 0, pend_r = -1;                       // arrival counter value returned for view pend_r (-1: none pending)

    auto issue_tile = [&](int stg, int t) {              // thread 0 only: arm the stage's barrier and start the copies
        ctl.tile[stg] = t;
        if (t >= total) return;
        const int r = t / tpv, blk = t - r * tpv;
        const ldp_ref_desc* rd = refs + r;
        const int nn = G.front_cache ? c_nn[r] : min(max(rd->nn, 0), nn_stage);
        ctl.tile_nn[stg] = nn;
        const int px0 = blk * KF_TILE;
        const uint32_t bytes = (uint32_t)(min(KF_TILE, N - px0) * (int)sizeof(float));
        const uint32_t bar = smem_u32(&ctl.full_bar[stg]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the stage was read / written through the generic proxy
        mbar_expect_tx(bar, bytes * (uint32_t)nn);
        for (int k = 0; k < nn; ++k) {
            const float* src = G.front_cache ? c_cert[(size_t)r * nn_stage + k] : rd->cert[k];
            tma_load_1d(smem_u32(stage_base + ((size_t)stg * nn_stage + k) * KF_TILE), src + px0, bytes, bar, policy);
        }
    };

    if (tid == 0) {
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
        for (int s = 0; s < KF_STAGES; ++s) mbar_init(smem_u32(&ctl.full_bar[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        ctl.last_r = -1;
    }
    for (int i = tid; i < G.prep_lb_cap; i += KF_THREADS) lb[i] = 0ull;
    if (G.front_cache) {                                   // host-written descriptors: no kernel produces them
        for (int i = tid; i < P.n_refs; i += KF_THREADS) c_nn[i] = min(max(refs[i].nn, 0), nn_stage);
        for (int i = tid; i < P.n_refs * nn_stage; i += KF_THREADS) c_cert[i] = refs[i / nn_stage].cert[i % nn_stage];
    }
    __syncthreads();
    grid_dependency_sync();
    if (tid == 0) {
        for (int s = 0; s < KF_STAGES; ++s) issue_tile(s, atomicAdd(ws.ticket, 1));
    }
    __syncthreads();

    // FIFO of parked tiles and the state of the stages: identical in every thread of the CTA
    int head = 0, npark = 0;
    uint32_t free_slots = (1u << KF_PARK) - 1u;          // park slots not in use
    uint32_t busy = 0u;                                  // stages whose plane 0 holds a parked tile
    uint32_t refill = 0u;                                // stages whose next ticket is requested (thread 0: pend_tk) but not issued
    uint32_t uses[KF_STAGES];                            // fills consumed per stage (mbarrier phase parity)
#pragma unroll
    for (int q = 0; q < KF_STAGES; ++q) uses[q] = 0u;

    // thread 0: issue the TMA of every stage whose ticket has come back (callers follow with a CTA barrier)
    auto flush_refills = [&]() {
#pragma unroll
        for (int q = 0; q < KF_STAGES; ++q) if ((refill >> q) & 1u) issue_tile(q, pend_tk[q]);
    };
    // thread 0: ask for the next ticket of a stage; the value is first looked at by the next flush_refills
    auto request_ticket = [&](int stg) {
#pragma unroll
        for (int q = 0; q < KF_STAGES; ++q) if (q == stg) pend_tk[q] = atomicAdd(ws.ticket, 1);
    };

    // ---- the CTA that completed view r: fixed-order sum of the tile partials -> s, published as one 64-bit word.
    //      A slot still zero means its (fence-less) store is in flight: read it again.
    auto publish_view = [&](int r) {
        if (warp == 0) {
            const ldp_ref_desc* rd = refs + r;
            unsigned long long* slots = reinterpret_cast<unsigned long long*>(ws.partial + (size_t)r * tpv);
            double a = 0.0;
            int f = 0;
            for (int i = lane; i < tpv; i += 32) {
                unsigned long long v = ld_relaxed_u64(slots + i);
                int fl = __ldcg(ws.bflags + (size_t)r * tpv + i);
                while (v == 0ull || fl == 0) { __nanosleep(32); v = ld_relaxed_u64(slots + i); fl = __ldcg(ws.bflags + (size_t)r * tpv + i); }
                a += __longlong_as_double((long long)~v);
                f |= fl;
                slots[i] = 0ull;                           // self-cleaning: a relaunch finds every slot empty again
                ws.bflags[(size_t)r * tpv + i] = 0;
            }
            a = warp_sum(a);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) f |= __shfl_xor_sync(0xffffffffu, f, o);
            if (lane == 0) {
                f &= 7;
                float s = (float)a;
                if (rd->weight_sum_override > 0.f) s = rd->weight_sum_override;
                ws.rstat[r].s = s;
                ws.rstat[r].bad = f;
                if (out.weight_sum) out.weight_sum[r] = s;
                ws.arrive[r] = 0;
                // what the normalise phase needs travels in the word itself: bit 0 = unusable view (NaN / negative weight,
                // no neighbours, s <= 0: the draw kernel reports which), bits of s
                ws.vword[r] = front_word(G.epoch, (f != 0 || !(s > 0.f)) ? 1 : 0, s);
            }
        }
    };
    // thread 0 looks at the arrival counter it got back for the tile announced one iteration ago
    auto resolve_arrival = [&]() {                        // thread 0 only
        ctl.last_r = (pend_r >= 0 && pend_old == tpv - 1) ? pend_r : -1;
        pend_r = -1;
    };

    // ---- normalise one parked tile: reference core/sampling.py:29 + chunk sums, coverage keys, exactness statistics
    auto normalise = [&](int loc, int r, int blk, float s) {
        const float* pk = (loc < KF_PARK) ? park_base + (size_t)loc * KF_TILE
                                          : stage_base + (size_t)(loc - KF_PARK) * nn_stage * KF_TILE;
        const int base = blk * KF_TILE;
        const int px = base + tid * 4;
        const int end = min(base + KF_TILE, N);
        const int y_first = (int)div_magic((uint32_t)base, G.w_magic), y_last = (int)div_magic((uint32_t)(end - 1), G.w_magic);
        const int ty0 = (int)__umulhi((uint32_t)y_first, G.t_magic32), ty1 = (int)__umulhi((uint32_t)y_last, G.t_magic32);
        const int nlb = (ty1 - ty0 + 1) * G.nbx;
        const float yr = __frcp_rn(s);
        const bool fast_div = (s >= 1.0f) && (s < 3.0e8f);
        float* __restrict__ w = ws.w + (size_t)r * ws.n_pad;
        double* __restrict__ csum = ws.csum + (size_t)r * ws.nchunk_pad;
        double* __restrict__ csum0 = ws.csum0 + (size_t)r * ws.nchunk_pad;
        int lpos = 0;
        uint32_t lmin1 = 0xFFFFFFFFu;
        double a = 0.0;
        unsigned long long key0 = 0ull;                    // best (p bits, ~index) of this quad for coverage tile bin0
        int bin0 = -1 - lane;                              // (lanes without a pixel: a bin of their own)
        if (px < N) {
            const float4 v = *reinterpret_cast<const float4*>(pk + tid * 4);
            float p0 = div_by(v.x, s, yr), p1 = div_by(v.y, s, yr), p2 = div_by(v.z, s, yr), p3 = div_by(v.w, s, yr);
            uint32_t cmin = min(min(__float_as_uint(p0) - 1u, __float_as_uint(p1) - 1u), min(__float_as_uint(p2) - 1u, __float_as_uint(p3) - 1u));
            if (cmin < 0x12000000u || !fast_div) {         // a non-zero quotient below 2^-91 (or an unusual s): IEEE division
                p0 = __fdiv_rn(v.x, s); p1 = __fdiv_rn(v.y, s); p2 = __fdiv_rn(v.z, s); p3 = __fdiv_rn(v.w, s);
                cmin = min(min(__float_as_uint(p0) - 1u, __float_as_uint(p1) - 1u), min(__float_as_uint(p2) - 1u, __float_as_uint(p3) - 1u));
            }
            lmin1 = cmin;
            const float pmin = fminf(fminf(p0, p1), fminf(p2, p3));
            if (pmin > 0.f) lpos = 4;
            else lpos = ((p0 > 0.f) ? 1 : 0) + ((p1 > 0.f) ? 1 : 0) + ((p2 > 0.f) ? 1 : 0) + ((p3 > 0.f) ? 1 : 0);
            *reinterpret_cast<float4*>(w + px) = make_float4(p0, p1, p2, p3);
            a = ((double)p0 + (double)p1) + ((double)p2 + (double)p3);
            // per-tile arg-max of p, lowest index on ties: 64-bit key, exact pre-filter, rare CAS (a quad lies in one row)
            const int y = (int)div_magic((uint32_t)px, G.w_magic), x = px - y * W;
            const int tx0 = (int)__umulhi((uint32_t)x, G.t_magic32);
            const int nsplit = (tx0 + 1) * G.tile - x;
            const int b0 = ((int)__umulhi((uint32_t)y, G.t_magic32) - ty0) * G.nbx + tx0;
            const bool s1 = nsplit > 1, s2 = nsplit > 2, s3 = nsplit > 3;
            const float m0 = fmaxf(fmaxf(p0, s1 ? p1 : 0.f), fmaxf(s2 ? p2 : 0.f, s3 ? p3 : 0.f));
            const unsigned long long k0 = ((unsigned long long)__float_as_uint(m0) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)px);
            const int j0 = (p0 == m0) ? 0 : (s1 && p1 == m0) ? 1 : (s2 && p2 == m0) ? 2 : 3;
            key0 = (m0 > 0.f) ? k0 - (unsigned long long)j0 : 0ull;
            bin0 = b0;
            if (!s3) {                                     // the quad straddles a tile edge: its tail goes to the next tile (rare)
                const float m1 = fmaxf(fmaxf(s1 ? 0.f : p1, s2 ? 0.f : p2), p3);
                const unsigned long long k1 = ((unsigned long long)__float_as_uint(m1) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)px);
                if (k1 > lb[b0 + 1]) {
                    const int j1 = (!s1 && p1 == m1) ? 1 : (!s2 && p2 == m1) ? 2 : 3;
                    if (m1 > 0.f) atomicMax(&lb[b0 + 1], k1 - (unsigned long long)j1);
                }
            }
        } else if (px < (int)ws.n_pad) {                  // row padding stays zero: the draw kernel's 32-byte scans read it
            *reinterpret_cast<float4*>(w + px) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // per-tile arg-max: the lanes of a warp walk along the image row, so lanes of one coverage tile are neighbours --
        // a segmented max-scan leaves each tile's best key in its last lane, which alone touches the shared-memory table
        // (one 64-bit CAS per warp and tile instead of one per quad: the quads of a saturated region all tie)
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long ok = __shfl_up_sync(0xffffffffu, key0, o);
            const int ob = __shfl_up_sync(0xffffffffu, bin0, o);
            if (lane >= o && ob == bin0 && ok > key0) key0 = ok;
        }
        {
            const int nb = __shfl_down_sync(0xffffffffu, bin0, 1);
            if ((lane == 31 || nb != bin0) && key0 != 0ull && key0 > lb[bin0]) atomicMax(&lb[bin0], key0);
        }
        // chunk sums (a warp covers 128 consecutive pixels)
        const int cs = G.chunk_shift;
        if (cs <= 7) {
            const int gl = (1 << cs) >> 2;
            for (int o = 1; o < gl; o <<= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (px < N && (lane & (gl - 1)) == 0) { csum[px >> cs] = a; csum0[px >> cs] = a; }
        } else {                                           // chunks wider than a warp row (zeroed by the host memset)
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (px < N && lane == 0) { atomicAdd(&csum[px >> cs], a); atomicAdd(&csum0[px >> cs], a); }
        }
        // positives / smallest positive exponent: one combined block reduction
        lpos = warp_sum(lpos);
        int m1s = warp_min((int)(lmin1 ^ 0x80000000u));    // unsigned order through the signed reduction
        if (lane == 0) { ctl.red_i[warp] = lpos; ctl.red_j[warp] = m1s; }
        __syncthreads();                                   // also: every key of this tile is in lb
        if (warp == 0) {
            int np = (lane < KF_THREADS / 32) ? ctl.red_i[lane] : 0;
            int mm = (lane < KF_THREADS / 32) ? ctl.red_j[lane] : 0x7fffffff;
            np = warp_sum(np);
            mm = warp_min(mm);
            if (lane == 0) {
                const uint32_t min1 = (uint32_t)mm ^ 0x80000000u;
                if (np) atomicAdd(&ws.rstat[r].npos, np);
                if (min1 != 0xFFFFFFFFu) atomicMax(&ws.rstat[r].emin_inv, 255 - (int)(((min1 + 1u) >> 23) & 0xffu));
            }
        }
        unsigned long long* gb = ws.gbins + (size_t)r * ws.bins_cap;
        for (int i = tid; i < nlb; i += KF_THREADS) {
            const unsigned long long key = lb[i];
            if (key) {
                atomicMax(&gb[(ty0 + i / G.nbx) * G.nbx + (i % G.nbx)], key);
                lb[i] = 0ull;                              // ready for the next tile
            }
        }
        __syncthreads();
    };

    // after the decision "the oldest parked tile is ready" has been broadcast in ctl (ready, s, bad): normalise it and pop
    auto pop_ready = [&](bool second) {
        const float s = second ? ctl.s2 : ctl.s;
        const int bad = second ? ctl.bad2 : ctl.bad;
        const int r = ctl.ent_r[head], pblk = ctl.ent_blk[head], loc = ctl.ent_loc[head];
        __syncthreads();                                   // ctl.* may be rewritten
        if (!bad) normalise(loc, r, pblk, s);              // else: the draw kernel reports the view's status
        head = (head + 1) % KF_ENT;
        --npark;
        if (loc < KF_PARK) {
            free_slots |= 1u << loc;
        } else {                                           // the stage is free again: ask for its next ticket
            const int stg = loc - KF_PARK;
            busy &= ~(1u << stg);
            refill |= 1u << stg;
            if (tid == 0) request_ticket(stg);
        }
    };
    // blocking form (only ever called when this CTA holds no ticket it has not announced): wait for the oldest parked view
    auto normalise_oldest_blocking = [&]() {
        if (tid == 0) {
            resolve_arrival();
        }
        __syncthreads();
        if (ctl.last_r >= 0) publish_view(ctl.last_r);
        if (tid == 0) {
            const unsigned long long* vw = ws.vword + ctl.ent_r[head];
            unsigned long long v = ld_relaxed_u64(vw);
            while ((int)(v >> 34) != G.epoch) { __nanosleep(64); v = ld_relaxed_u64(vw); }
            ctl.ready = 1;
            ctl.s = __uint_as_float((uint32_t)v);
            ctl.bad = (int)((v >> 32) & 3ull);
        }
        __syncthreads();
        pop_ready(false);
    };

    int cur = 0;
    FCLK_DECL;
    for (;;) {
        // ---- the next stage to consume: any stage with a tile in it (or on its way).  If there is none: first turn requested
        //      tickets into copies; then, if a stage is occupied by a parked tile, normalise the oldest parked tile
        //      (blocking: nothing un-announced is held at that point); else all tickets are gone.
        int stg = -1;
#pragma unroll
        for (int q = 0; q < KF_STAGES; ++q) {
            const int c = (cur + q) % KF_STAGES;
            if (stg < 0 && !(((busy | refill) >> c) & 1u) && ctl.tile[c] < total) stg = c;
        }
        if (stg < 0) {
            if (refill) {
                if (tid == 0) flush_refills();
                refill = 0u;
                __syncthreads();
                FCLK(7);
                continue;
            }
            if (busy == 0u) break;
            normalise_oldest_blocking();
            FCLK(6); FCNT(10);
            continue;
        }
        cur = (stg + 1) % KF_STAGES;
        const int t = ctl.tile[stg];
        const int nn = ctl.tile_nn[stg];
        const int r = t / tpv, blk = t - r * tpv;
        const ldp_ref_desc* rd = refs + r;
        const int base = blk * KF_TILE;
        const int px = base + tid * 4;
        // thread 0: the poll of the oldest parked view's word travels while the tile is streamed
        unsigned long long polled = 0ull, polled2 = 0ull;
        if (tid == 0 && npark > 0) polled = ld_relaxed_u64(ws.vword + ctl.ent_r[head]);
        if (tid == 0 && npark > 1) polled2 = ld_relaxed_u64(ws.vword + ctl.ent_r[(head + 1) % KF_ENT]);
        if (tid < KF_TILE / 32) {                          // the draw kernels' selection bitmap of this tile's pixels
            const int wi = base / 32 + tid;
            if (wi < (int)ws.n_words) ws.bitmap[(size_t)r * ws.n_words + wi] = 0u;
        }
        uint32_t parity = 0u;
#pragma unroll
        for (int q = 0; q < KF_STAGES; ++q) if (q == stg) { parity = uses[q] & 1u; ++uses[q]; }
        FCLK(0);
        mbar_wait(smem_u32(&ctl.full_bar[stg]), parity);
        FCLK(1); FCNT(8);
        // ---- stream phase
        double lsum = 0.0;
        float lmin = 0.f;
        float wv[4] = {0.f, 0.f, 0.f, 0.f};
        int bi[4] = {0, 0, 0, 0};
        if (px < N && nn > 0) {
            const float* sp = stage_base + (size_t)stg * nn_stage * KF_TILE + tid * 4;
            float ma[4] = {1.f, 1.f, 1.f, 1.f};
            bool has_a = false;
            float fl = 0.f;
            if (PRO) {
                fl = P.certainty_floor;
                const uint8_t* mk = rd->mask_a;
                has_a = mk != nullptr;
                if (has_a) {
                    const int mw = rd->mask_w, mh = rd->mask_h;
                    if (mw == W && mh == H && (reinterpret_cast<uintptr_t>(mk) & 3u) == 0) {
                        const uint32_t m = __ldg(reinterpret_cast<const uint32_t*>(mk + px));
                        ma[0] = (float)(m & 0xffu); ma[1] = (float)((m >> 8) & 0xffu); ma[2] = (float)((m >> 16) & 0xffu); ma[3] = (float)(m >> 24);
                    } else {
                        const float msx = rd->mask_sx, msy = rd->mask_sy;
                        const int y = (int)div_magic((uint32_t)px, G.w_magic), x = px - y * W;
                        const int my = min((int)floorf(__fmul_rn((float)y, msy)), mh - 1);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int mx = min((int)floorf(__fmul_rn((float)(x + j), msx)), mw - 1);
                            ma[j] = (float)__ldg(mk + (size_t)my * mw + mx);
                        }
                    }
                }
            }
            auto post = [&](float4& c) {
                if (PRO) {
                    c.x = (c.x < fl) ? fl : c.x; c.y = (c.y < fl) ? fl : c.y;           // torch.clamp(min=): NaN stays NaN
                    c.z = (c.z < fl) ? fl : c.z; c.w = (c.w < fl) ? fl : c.w;
                    if (has_a) { c.x = __fmul_rn(c.x, ma[0]); c.y = __fmul_rn(c.y, ma[1]); c.z = __fmul_rn(c.z, ma[2]); c.w = __fmul_rn(c.w, ma[3]); }
                }
            };
            float4 c0 = *reinterpret_cast<const float4*>(sp);
            post(c0);
            wv[0] = c0.x; wv[1] = c0.y; wv[2] = c0.z; wv[3] = c0.w;
#pragma unroll 4
            for (int k = 1; k < nn; ++k) {                 // NaN-propagating max, first index wins
                float4 c = *reinterpret_cast<const float4*>(sp + (size_t)k * KF_TILE);
                post(c);
                if (c.x > wv[0] || c.x != c.x) { wv[0] = c.x; bi[0] = k; }
                if (c.y > wv[1] || c.y != c.y) { wv[1] = c.y; bi[1] = k; }
                if (c.z > wv[2] || c.z != c.z) { wv[2] = c.z; bi[2] = k; }
                if (c.w > wv[3] || c.w != c.w) { wv[3] = c.w; bi[3] = k; }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) wv[j] = (wv[j] > cap) ? cap : wv[j];            // torch.clamp(max=cap): NaN stays NaN
            const int y = (int)div_magic((uint32_t)px, G.w_magic), x = px - y * W;
            const bool interior = y >= border && y <= H - 1 - border && x >= border && x + 3 <= W - 1 - border && px + 3 < N;
            if (!interior) quad_border_weights(wv, px, x, y, W, H, border, N);
            lmin = fminf(fminf(wv[0], wv[1]), fminf(wv[2], wv[3]));
            lsum = (widen_f32(wv[0]) + widen_f32(wv[1])) + (widen_f32(wv[2]) + widen_f32(wv[3]));
            *reinterpret_cast<uint32_t*>(ws.bestk + (size_t)r * ws.n_pad + px) =
                (uint32_t)bi[0] | ((uint32_t)bi[1] << 8) | ((uint32_t)bi[2] << 16) | ((uint32_t)bi[3] << 24);
        }
        // where the weights wait for their normaliser: a park slot, or -- none free -- plane 0 of the stage itself (in place:
        // a thread overwrites only the quad it has just read)
        const int loc = free_slots ? (__ffs(free_slots) - 1) : KF_PARK + stg;
        float* dstw = (loc < KF_PARK) ? park_base + (size_t)loc * KF_TILE : stage_base + (size_t)stg * nn_stage * KF_TILE;
        *reinterpret_cast<float4*>(dstw + tid * 4) = make_float4(wv[0], wv[1], wv[2], wv[3]);
        if (loc < KF_PARK) { free_slots &= ~(1u << loc); } else { busy |= 1u << stg; }
        const int ent = (head + npark) % KF_ENT;
        const uint32_t refill_before = refill;
        if (loc < KF_PARK) refill |= 1u << stg;
        // ---- the tile's partial sum and flags
        lsum = warp_sum(lsum);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lmin = fminf(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
        if (lane == 0) { ctl.red_d[warp] = lsum; ctl.red_f[warp] = lmin; }
        __syncthreads();                                   // the stage is consumed; partials are in shared memory
        FCLK(2);
        if (loc >= KF_PARK) FCNT(9);
        if (tid == 0) {
            // copies of the stages whose ticket was requested an iteration ago; then this stage's next ticket (looked at next time)
#pragma unroll
            for (int q = 0; q < KF_STAGES; ++q) if ((refill_before >> q) & 1u) issue_tile(q, pend_tk[q]);
            if (loc < KF_PARK) request_ticket(stg);
            resolve_arrival();                             // the tile announced one iteration ago
            // announce this tile: slot stores without a fence (the summing CTA re-reads a slot that is still empty), counter
            double bsum = 0.0;
            float bmin = 0.f;
            for (int q = 0; q < KF_THREADS / 32; ++q) { bsum += ctl.red_d[q]; bmin = fminf(bmin, ctl.red_f[q]); }
            reinterpret_cast<unsigned long long*>(ws.partial)[(size_t)r * tpv + blk] = ~(unsigned long long)__double_as_longlong(bsum);
            ws.bflags[(size_t)r * tpv + blk] = 8 | ((bsum != bsum) ? 1 : 0) | ((bmin < 0.f) ? 2 : 0) | ((nn <= 0) ? 4 : 0);
            pend_old = atomicAdd(ws.arrive + r, 1);
            pend_r = r;
            ctl.ent_r[ent] = r; ctl.ent_blk[ent] = blk; ctl.ent_loc[ent] = loc;
            // the poll issued before the stream phase
            const bool rdy = npark > 0 && (int)(polled >> 34) == G.epoch;
            const bool rdy2 = rdy && npark > 1 && (int)(polled2 >> 34) == G.epoch;
            ctl.ready = rdy ? (rdy2 ? 2 : 1) : 0;
            if (rdy) { ctl.s = __uint_as_float((uint32_t)polled); ctl.bad = (int)((polled >> 32) & 3ull); }
            if (rdy2) { ctl.s2 = __uint_as_float((uint32_t)polled2); ctl.bad2 = (int)((polled2 >> 32) & 3ull); }
        }
        refill &= ~refill_before;
        __syncthreads();
        FCLK(3);
        ++npark;
        if (ctl.last_r >= 0) { publish_view(ctl.last_r); FCLK(4); }
        const int nready = ctl.ready;
        if (nready >= 1) { pop_ready(false); FCLK(5); FCNT(11); }
        if (nready >= 2) { pop_ready(true); FCLK(5); FCNT(11); }
    }
    // ---- no tickets left: the pending arrival, then the parked tiles in order
    while (npark > 0) { normalise_oldest_blocking(); FCLK(6); FCNT(10); }
    FCLK_FLUSH;
    if (tid == 0) resolve_arrival();
    __syncthreads();
    if (ctl.last_r >= 0) publish_view(ctl.last_r);
    // ---- self-cleaning ticket counter: the last CTA to leave resets it (ldp_debug_launch_stream relaunches without memset)
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(ws.ticket + 1, 1) == (int)gridDim.x - 1) { ws.ticket[0] = 0; ws.ticket[1] = 0; }
    }
}

}  // namespace ldp
