// K2: per-sample decode -> colour -> Sampson gate -> DLT triangulation -> reprojection / cheirality /
//     parallax filters.  One thread per sampled pixel, grid = (sample tiles, reference views).
// K3: ordered stream compaction of the kept samples into the packed point cloud.
//
// Replaces reference core/pipeline.py:652-780 and core/geometry.py:58-141 (see the per-step citations).
// Arithmetic mirrors the reference's dtype flow and operation order:
//   * decode / pixel scaling / DLT rows / reprojection / parallax in f32 with NO fused multiply-add
//     except where the reference's sgemm uses one (the 4-term projection dot products);
//   * Sampson distance and the bilinear colour weights in f64;
//   * the 4x4 null vector: the reference calls LAPACK's f32 SVD; here it is the eigenvector of the
//     smallest eigenvalue of A^T A by shifted inverse iteration in f64 (both sit ~1e-7 relative from the
//     exact singular vector of the same f32 matrix; tolerance 1e-4 rel / 1e-5 abs).
// Per-pair camera constants (P1,P2,C1,C2,F, pixel scales, group ids) are staged in shared memory.
#include "ldp_device.cuh"

namespace ldp {

constexpr int K2_THREADS = 128;
constexpr int K3_THREADS = 1024;
// np.degrees on float32 multiplies by f32(180) / f32(pi) evaluated in f32 (measured)
#define RAD2DEG_F32 57.295776367187500f

struct PairConst {          // one neighbour, staged in shared memory
    float P2[12];
    float C2[3];
    float F[9];
    float sxB, syB;
    int group;
    const float* warp;
    const float* cert;
};
struct RefConst {
    float P1[12];
    float C1[3];
    float sxA, syA, sx_img, sy_img;
    int img_w, img_h, nn;
    const uint8_t* image;
};

// Robust fallback: cyclic Jacobi eigen-decomposition of the symmetric 4x4 M (f64); returns the eigenvector
// of the smallest eigenvalue.  Only reached when inverse iteration has not converged (sigma_4 ~ sigma_3:
// grossly inconsistent matches), so it is kept out of line to protect the fast path's register budget.
__device__ __noinline__ void jacobi_smallest_eigvec4(const double* __restrict__ Min, double* __restrict__ vout) {
    double a[4][4], V[4][4];
    a[0][0] = Min[0]; a[1][0] = a[0][1] = Min[1]; a[1][1] = Min[2];
    a[2][0] = a[0][2] = Min[3]; a[2][1] = a[1][2] = Min[4]; a[2][2] = Min[5];
    a[3][0] = a[0][3] = Min[6]; a[3][1] = a[1][3] = Min[7]; a[3][2] = a[2][3] = Min[8]; a[3][3] = Min[9];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
    const double tr = a[0][0] + a[1][1] + a[2][2] + a[3][3];
    const double tiny = tr * tr * 1e-34;
    for (int sweep = 0; sweep < 16; ++sweep) {
        const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[0][3] * a[0][3] +
                           a[1][2] * a[1][2] + a[1][3] * a[1][3] + a[2][3] * a[2][3];
        if (!(off > tiny)) break;
#pragma unroll
        for (int p = 0; p < 3; ++p) {
#pragma unroll
            for (int q = p + 1; q < 4; ++q) {
                const double apq = a[p][q];
                if (fabs(apq) > 1e-300) {
                    const double tau = (a[q][q] - a[p][p]) / (2.0 * apq);
                    const double t = ((tau >= 0.0) ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                    const double c = rsqrt(1.0 + t * t), sn = t * c;
                    a[p][p] -= t * apq;
                    a[q][q] += t * apq;
                    a[p][q] = a[q][p] = 0.0;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (k != p && k != q) {
                            const double akp = a[k][p], akq = a[k][q];
                            a[k][p] = a[p][k] = c * akp - sn * akq;
                            a[k][q] = a[q][k] = sn * akp + c * akq;
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const double vkp = V[k][p], vkq = V[k][q];
                        V[k][p] = c * vkp - sn * vkq;
                        V[k][q] = sn * vkp + c * vkq;
                    }
                }
            }
        }
    }
    int j = 0;
    double best = a[0][0];
    if (a[1][1] < best) { best = a[1][1]; j = 1; }
    if (a[2][2] < best) { best = a[2][2]; j = 2; }
    if (a[3][3] < best) { best = a[3][3]; j = 3; }
#pragma unroll
    for (int k = 0; k < 4; ++k) vout[k] = (j == 0) ? V[k][0] : (j == 1) ? V[k][1] : (j == 2) ? V[k][2] : V[k][3];
}

// smallest-eigenvalue eigenvector of M = A^T A (A 4x4 given row-major, f32 values held in f64).
// Fast path: inverse iteration on the Cholesky factor of M + mu*I (same eigenvectors; converges at
// (sigma_4/sigma_3)^2 per step: 3-5 steps on consistent matches).  Not converged after INVIT_MAX steps -> Jacobi.
constexpr int INVIT_MAX = 8;
__device__ __forceinline__ void null_vector4(const double A[16], double v[4]) {
    double m00 = 0, m10 = 0, m11 = 0, m20 = 0, m21 = 0, m22 = 0, m30 = 0, m31 = 0, m32 = 0, m33 = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const double a = A[4 * r], b = A[4 * r + 1], c = A[4 * r + 2], d = A[4 * r + 3];
        m00 = fma(a, a, m00); m10 = fma(b, a, m10); m11 = fma(b, b, m11);
        m20 = fma(c, a, m20); m21 = fma(c, b, m21); m22 = fma(c, c, m22);
        m30 = fma(d, a, m30); m31 = fma(d, b, m31); m32 = fma(d, c, m32); m33 = fma(d, d, m33);
    }
    const double tr = m00 + m11 + m22 + m33;
    if (!(tr > 0.0) || !isfinite(tr)) { v[0] = v[1] = v[2] = 0.0; v[3] = 1.0; if (!(tr == tr)) v[0] = tr; return; }
    // M + mu*I has the same eigenvectors; the shift keeps the Cholesky pivots positive when A is
    // numerically rank-3 (noise-free correspondences).
    const double mu = tr * 1e-13;
    double d0 = m00 + mu;
    const double i0 = rsqrt(d0);
    const double l10 = m10 * i0, l20 = m20 * i0, l30 = m30 * i0;
    double d1 = m11 + mu - l10 * l10; d1 = fmax(d1, mu * 1e-3);
    const double i1 = rsqrt(d1);
    const double l21 = (m21 - l20 * l10) * i1, l31 = (m31 - l30 * l10) * i1;
    double d2 = m22 + mu - l20 * l20 - l21 * l21; d2 = fmax(d2, mu * 1e-3);
    const double i2 = rsqrt(d2);
    const double l32 = (m32 - l30 * l20 - l31 * l21) * i2;
    double d3 = m33 + mu - l30 * l30 - l31 * l31 - l32 * l32; d3 = fmax(d3, mu * 1e-3);
    const double i3 = rsqrt(d3);
    double x0 = 0.0, x1 = 0.0, x2 = 0.0, x3 = 1.0;
    bool converged = false;
    for (int it = 0; it < INVIT_MAX; ++it) {
        // L y = x
        const double y0 = x0 * i0;
        const double y1 = (x1 - l10 * y0) * i1;
        const double y2 = (x2 - l20 * y0 - l21 * y1) * i2;
        const double y3 = (x3 - l30 * y0 - l31 * y1 - l32 * y2) * i3;
        // L^T z = y
        const double z3 = y3 * i3;
        const double z2 = (y2 - l32 * z3) * i2;
        const double z1 = (y1 - l21 * z2 - l31 * z3) * i1;
        const double z0 = (y0 - l10 * z1 - l20 * z2 - l30 * z3) * i0;
        const double inv = rsqrt(z0 * z0 + z1 * z1 + z2 * z2 + z3 * z3);
        const double sgn = (z0 * x0 + z1 * x1 + z2 * x2 + z3 * x3) < 0.0 ? -inv : inv;
        const double n0 = z0 * sgn, n1 = z1 * sgn, n2 = z2 * sgn, n3 = z3 * sgn;
        const double e0 = n0 - x0, e1 = n1 - x1, e2 = n2 - x2, e3 = n3 - x3;
        x0 = n0; x1 = n1; x2 = n2; x3 = n3;
        if (e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3 < 1e-26) { converged = true; break; }
    }
    if (converged) {
        v[0] = x0; v[1] = x1; v[2] = x2; v[3] = x3;
    } else {
        const double Ms[10] = {m00, m10, m11, m20, m21, m22, m30, m31, m32, m33};
        jacobi_smallest_eigvec4(Ms, v);
    }
}

// X @ P^T row: the reference's sgemm accumulates the K=4 products with FMAs in index order
// (measured against numpy/OpenBLAS); same here.
__device__ __forceinline__ float proj_row(const float* p, float X0, float X1, float X2, float X3) {
    float acc = __fmul_rn(X0, p[0]);
    acc = __fmaf_rn(X1, p[1], acc);
    acc = __fmaf_rn(X2, p[2], acc);
    acc = __fmaf_rn(X3, p[3], acc);
    return acc;
}

// core/geometry.py:91-104
__device__ __forceinline__ float reproj_err(const float* P, float X0, float X1, float X2, float X3,
                                            float u, float v, float* zout) {
    const float q0 = proj_row(P, X0, X1, X2, X3), q1 = proj_row(P + 4, X0, X1, X2, X3), q2 = proj_row(P + 8, X0, X1, X2, X3);
    *zout = q2;
    const float z = fmaxf(q2, 1e-12f);
    const float du = __fsub_rn(__fdiv_rn(q0, z), u), dv = __fsub_rn(__fdiv_rn(q1, z), v);
    return __fsqrt_rn(__fadd_rn(__fmul_rn(du, du), __fmul_rn(dv, dv)));
}

__global__ void __launch_bounds__(K2_THREADS)
ldp_geometry_kernel(const ldp_params P, const ldp_ref_desc* __restrict__ refs, const Workspace ws, const ldp_outputs out,
                    const int have_bestk)
{
    __shared__ RefConst rc;
    __shared__ PairConst pc[LDP_MAX_NN];
    const int r = blockIdx.y;
    const int S = out.n_samples[r];
    const int i0 = blockIdx.x * blockDim.x;
    if (i0 >= S) return;
    const ldp_ref_desc* rd = refs + r;
    {   // stage the view's camera constants
        const int t = threadIdx.x;
        if (t < 12) rc.P1[t] = rd->P1[t];
        if (t < 3) rc.C1[t] = rd->C1[t];
        if (t == 0) {
            rc.sxA = rd->sxA; rc.syA = rd->syA; rc.sx_img = rd->sx_img; rc.sy_img = rd->sy_img;
            rc.img_w = rd->img_w; rc.img_h = rd->img_h; rc.nn = rd->nn; rc.image = rd->image;
        }
        const int nn = rd->nn;
        for (int e = t; e < nn * 12; e += blockDim.x) pc[e / 12].P2[e % 12] = rd->P2[e / 12][e % 12];
        for (int e = t; e < nn * 9; e += blockDim.x) pc[e / 9].F[e % 9] = rd->F[e / 9][e % 9];
        for (int e = t; e < nn * 3; e += blockDim.x) pc[e / 3].C2[e % 3] = rd->C2[e / 3][e % 3];
        for (int e = t; e < nn; e += blockDim.x) {
            pc[e].sxB = rd->sxB[e]; pc[e].syB = rd->syB[e]; pc[e].group = rd->group[e];
            pc[e].warp = rd->warp[e]; pc[e].cert = rd->cert[e];
        }
    }
    __syncthreads();
    const int i = i0 + threadIdx.x;
    const bool active = i < S;
    int keep = 0, good = 0, grp = 0;
    if (active) {
        const int32_t* sel = (out.sel_idx ? out.sel_idx : ws.sel) + (size_t)r * ws.sel_cap;
        const int idx = sel[i];
        int k = 0;                                                                // core/pipeline.py:634-635,652
        if (have_bestk) {
            k = ws.bestk[(size_t)r * ws.n_pad + idx];
        } else {                           // stage entry point: arg-max over neighbours at the sampled pixel only
            float best = __ldg(pc[0].cert + idx);
            for (int q = 1; q < rc.nn; ++q) {
                const float c = __ldg(pc[q].cert + idx);
                if (c > best) { best = c; k = q; }
            }
        }
        const PairConst& pk = pc[k];
        grp = pk.group;
        const float4 wv = __ldg(reinterpret_cast<const float4*>(pk.warp) + idx);  // core/pipeline.py:636-640,653
        const float wm1 = (float)(P.w_match - 1), hm1 = (float)(P.h_match - 1);
        // core/pipeline.py:655-656 and 701-702: ((x + 1.0) * 0.5) * (w_match - 1), f32 op by op
        const float xA = __fmul_rn(__fmul_rn(__fadd_rn(wv.x, 1.0f), 0.5f), wm1);
        const float yA = __fmul_rn(__fmul_rn(__fadd_rn(wv.y, 1.0f), 0.5f), hm1);
        const float xB = __fmul_rn(__fmul_rn(__fadd_rn(wv.z, 1.0f), 0.5f), wm1);
        const float yB = __fmul_rn(__fmul_rn(__fadd_rn(wv.w, 1.0f), 0.5f), hm1);

        // ---- colour: core/pipeline.py:661-679 (f64 weights, clipped corners)
        float cr, cg, cb;
        {
            const float fx = __fmul_rn(xA, rc.sx_img), fy = __fmul_rn(yA, rc.sy_img);
            const int iw = rc.img_w, ih = rc.img_h;
            // floor(...).astype(int32) then clip; clamp in float first so the cast cannot overflow
            const int x0 = min(max((int)fminf(fmaxf(floorf(fx), -1.f), (float)iw), 0), iw - 1);
            const int y0 = min(max((int)fminf(fmaxf(floorf(fy), -1.f), (float)ih), 0), ih - 1);
            const int x1 = min(x0 + 1, iw - 1), y1 = min(y0 + 1, ih - 1);
            const double dfx = (double)fx, dfy = (double)fy;
            const double ax = __dsub_rn((double)x1, dfx), bx = __dsub_rn(dfx, (double)x0);
            const double ay = __dsub_rn((double)y1, dfy), by = __dsub_rn(dfy, (double)y0);
            const double w00 = __dmul_rn(ax, ay), w01 = __dmul_rn(bx, ay), w10 = __dmul_rn(ax, by), w11 = __dmul_rn(bx, by);
            const uint8_t* im = rc.image;
            const uint8_t* t00 = im + ((size_t)y0 * iw + x0) * 3;
            const uint8_t* t01 = im + ((size_t)y0 * iw + x1) * 3;
            const uint8_t* t10 = im + ((size_t)y1 * iw + x0) * 3;
            const uint8_t* t11 = im + ((size_t)y1 * iw + x1) * 3;
            float col[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                double acc = __dmul_rn((double)__ldg(t00 + c), w00);
                acc = __dadd_rn(acc, __dmul_rn((double)__ldg(t01 + c), w01));
                acc = __dadd_rn(acc, __dmul_rn((double)__ldg(t10 + c), w10));
                acc = __dadd_rn(acc, __dmul_rn((double)__ldg(t11 + c), w11));
                col[c] = (float)__ddiv_rn(acc, 255.0);                            // .astype(np.float32) at :754
            }
            cr = col[0]; cg = col[1]; cb = col[2];
        }
        // ---- full-resolution pixel coordinates: core/pipeline.py:681-683, 697-703
        const float uA = __fmul_rn(xA, rc.sxA), vA = __fmul_rn(yA, rc.syA);
        const float uB = __fmul_rn(xB, pk.sxB), vB = __fmul_rn(yB, pk.syB);

        // ---- Sampson gate in f64: core/geometry.py:133-141, core/pipeline.py:708-727
        good = 1;
        if (!P.no_filter && P.sampson_thresh > 0.0) {
            const double a0 = uA, a1 = vA, b0 = uB, b1 = vB;
            const float* F = pk.F;
            const double l0 = (double)F[0] * a0 + (double)F[1] * a1 + (double)F[2];
            const double l1 = (double)F[3] * a0 + (double)F[4] * a1 + (double)F[5];
            const double l2 = (double)F[6] * a0 + (double)F[7] * a1 + (double)F[8];
            const double m0 = (double)F[0] * b0 + (double)F[3] * b1 + (double)F[6];
            const double m1 = (double)F[1] * b0 + (double)F[4] * b1 + (double)F[7];
            const double num = b0 * l0 + b1 * l1 + l2;
            const double den = l0 * l0 + l1 * l1 + m0 * m0 + m1 * m1 + 1e-12;
            const double se = (num * num) / den;
            good = (se < P.sampson_thresh) ? 1 : 0;
        }

        float X0 = 0.f, X1 = 0.f, X2 = 0.f, X3 = 0.f, err = __int_as_float(0x7fc00000);
        if (good) {        // the reference triangulates only the Sampson survivors (core/pipeline.py:721-733)
        // ---- DLT rows (f32, multiply then subtract): core/geometry.py:72-75
        double A[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            A[j]      = (double)__fsub_rn(__fmul_rn(uA, rc.P1[8 + j]), rc.P1[j]);
            A[4 + j]  = (double)__fsub_rn(__fmul_rn(vA, rc.P1[8 + j]), rc.P1[4 + j]);
            A[8 + j]  = (double)__fsub_rn(__fmul_rn(uB, pk.P2[8 + j]), pk.P2[j]);
            A[12 + j] = (double)__fsub_rn(__fmul_rn(vB, pk.P2[8 + j]), pk.P2[4 + j]);
        }
        double v[4];
        null_vector4(A, v);
        // core/geometry.py:85-87: w = where(|Xh3| < 1e-12, 1e-12, Xh3); X = Xh / w
        {
            const float w32 = (float)v[3];
            if (fabsf(w32) < 1e-12f) {
                X0 = __fdiv_rn((float)v[0], 1e-12f); X1 = __fdiv_rn((float)v[1], 1e-12f);
                X2 = __fdiv_rn((float)v[2], 1e-12f); X3 = __fdiv_rn(w32, 1e-12f);
            } else {
                const double iw = 1.0 / v[3];
                X0 = (float)(v[0] * iw); X1 = (float)(v[1] * iw); X2 = (float)(v[2] * iw); X3 = 1.0f;
            }
        }
        // ---- reprojection errors + cheirality: core/geometry.py:91-110, core/pipeline.py:735-737
        float z1, z2;
        const float e1 = reproj_err(rc.P1, X0, X1, X2, X3, uA, vA, &z1);
        const float e2 = reproj_err(pk.P2, X0, X1, X2, X3, uB, vB, &z2);
        err = (e1 != e1 || e2 != e2) ? __int_as_float(0x7fc00000) : fmaxf(e1, e2);   // np.maximum propagates NaN
        if (P.no_filter) {                                                        // core/pipeline.py:739-743
            keep = (isfinite(X0) && isfinite(X1) && isfinite(X2) && isfinite(X3) && isfinite(err)) ? 1 : 0;
        } else {                                                                  // core/pipeline.py:745-749
            keep = (good && err <= P.reproj_thresh && z1 > 0.0f && z2 > 0.0f) ? 1 : 0;
            if (keep && P.min_parallax_deg > 0.0f) {                              // core/geometry.py:113-119
                float a0 = __fsub_rn(X0, rc.C1[0]), a1 = __fsub_rn(X1, rc.C1[1]), a2 = __fsub_rn(X2, rc.C1[2]);
                float b0 = __fsub_rn(X0, pk.C2[0]), b1 = __fsub_rn(X1, pk.C2[1]), b2 = __fsub_rn(X2, pk.C2[2]);
                const float na = __fadd_rn(__fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(a0, a0), __fmul_rn(a1, a1)), __fmul_rn(a2, a2))), 1e-12f);
                const float nb = __fadd_rn(__fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(b0, b0), __fmul_rn(b1, b1)), __fmul_rn(b2, b2))), 1e-12f);
                a0 = __fdiv_rn(a0, na); a1 = __fdiv_rn(a1, na); a2 = __fdiv_rn(a2, na);
                b0 = __fdiv_rn(b0, nb); b1 = __fdiv_rn(b1, nb); b2 = __fdiv_rn(b2, nb);
                float d = __fadd_rn(__fadd_rn(__fmul_rn(a0, b0), __fmul_rn(a1, b1)), __fmul_rn(a2, b2));
                const bool dnan = (d != d);
                d = fminf(fmaxf(d, -1.0f), 1.0f);
                const float ang = __fmul_rn((float)acos((double)d), RAD2DEG_F32);
                keep = (!dnan && ang >= P.min_parallax_deg) ? 1 : 0;
            }
        }
        }   // good
        const size_t o = (size_t)r * ws.sel_cap + i;
        ws.pt0[o] = make_float4(X0, X1, X2, err);
        float dcert = 0.f;
        if (P.collect_debug) {                                                    // core/pipeline.py:761-769
            const float c = __ldg(pk.cert + idx);
            const float denom = (P.sample_cap > 1e-6f) ? P.sample_cap : 1.0f;
            dcert = fminf(fmaxf(__fdiv_rn(c, denom), 0.f), 1.f);
            ws.dbgm[o] = make_float4(fminf(fmaxf(xA, 0.f), wm1), fminf(fmaxf(yA, 0.f), hm1),
                                     fminf(fmaxf(xB, 0.f), wm1), fminf(fmaxf(yB, 0.f), hm1));
        }
        ws.pt1[o] = make_float4(cr, cg, cb, dcert);
        ws.flags[o] = (uint8_t)(keep | (good << 1) | (grp << 2));
        if (out.sample_flags) out.sample_flags[o] = (uint8_t)(keep | (good << 1) | (grp << 2));
        if (out.sample_xyzerr) reinterpret_cast<float4*>(out.sample_xyzerr)[o] = make_float4(X0, X1, X2, err);
    }
    // kept-point count of the view (one atomic per warp)
    const unsigned km = __ballot_sync(0xffffffffu, keep);
    if ((threadIdx.x & 31) == 0 && km) atomicAdd(&ws.kept[r], __popc(km));
}

// ---------------------------------------------------------------------------------------------
// K3: one CTA per reference view.  Output order of the reference (core/pipeline.py:685-695,753-780):
// neighbour groups in order of first appearance over the sample order, samples in sample order inside
// a group, only kept samples.  Rank of a kept sample inside its group = ordered ballot scan.
// The view's base offset is the sum of the kept counts of the views before it (ws.kept from K2).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(K3_THREADS, 1)
ldp_pack_kernel(const ldp_params P, const ldp_ref_desc* __restrict__ refs, const Workspace ws, const ldp_outputs out)
{
    __shared__ int s_first[LDP_MAX_NN];      // first sample position per group
    __shared__ int s_count[LDP_MAX_NN];      // kept per group
    __shared__ int s_base[LDP_MAX_NN];       // output base per group (relative to the view)
    __shared__ int s_run[LDP_MAX_NN];        // running kept count per group across tiles
    __shared__ int s_wcnt[32][LDP_MAX_NN];   // per-warp kept count per group in this tile
    __shared__ int s_order[LDP_MAX_NN];
    __shared__ long long s_red[32];
    __shared__ long long s_off;
    const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, T = blockDim.x;
    const int nwarp = T >> 5;
    const int S = out.n_samples[r];
    const uint8_t* flags = ws.flags + (size_t)r * ws.sel_cap;

    // base offset of this view
    long long part = 0;
    for (int q = tid; q < r; q += T) part += (long long)ws.kept[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) s_red[warp] = part;
    if (tid < LDP_MAX_NN) { s_first[tid] = 0x7fffffff; s_count[tid] = 0; s_run[tid] = 0; }
    __syncthreads();
    if (warp == 0) {
        long long v = (lane < nwarp) ? s_red[lane] : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) s_off = v;
    }
    // pass 1: first appearance + kept count per group
    for (int i = tid; i < S; i += T) {
        const int f = flags[i];
        const int g = f >> 2;
        atomicMin(&s_first[g], i);
        if (f & 1) atomicAdd(&s_count[g], 1);
    }
    __syncthreads();
    if (tid == 0) {
        // groups in first-appearance order (insertion sort of <= 16 entries)
        int n = 0;
        for (int g = 0; g < LDP_MAX_NN; ++g) if (s_first[g] != 0x7fffffff) s_order[n++] = g;
        for (int a = 1; a < n; ++a) {
            const int g = s_order[a];
            int b = a - 1;
            while (b >= 0 && s_first[s_order[b]] > s_first[g]) { s_order[b + 1] = s_order[b]; --b; }
            s_order[b + 1] = g;
        }
        int acc = 0;
        for (int a = 0; a < n; ++a) { s_base[s_order[a]] = acc; acc += s_count[s_order[a]]; }
        for (int a = 0; a < LDP_MAX_NN; ++a) {
            out.group_order[(size_t)r * LDP_MAX_NN + a] = (a < n) ? s_order[a] : -1;
            out.group_count[(size_t)r * LDP_MAX_NN + a] = s_count[a];
        }
        out.ref_offset[r] = s_off;
        if (r == (int)gridDim.x - 1) out.ref_offset[r + 1] = s_off + acc;
    }
    __syncthreads();
    const long long off = s_off;
    // pass 2: ordered scatter, tile by tile
    for (int t0 = 0; t0 < S; t0 += T) {
        const int i = t0 + tid;
        const int f = (i < S) ? flags[i] : 0;
        const int keep = f & 1, g = f >> 2;
        // per-warp, per-group ballots
        const unsigned same = __match_any_sync(0xffffffffu, keep ? g : -1);
        const int rank_in_warp = __popc(same & ((1u << lane) - 1u));
        for (int e = lane; e < LDP_MAX_NN; e += 32) s_wcnt[warp][e] = 0;
        __syncwarp();
        if (keep && rank_in_warp == 0) s_wcnt[warp][g] = __popc(same);
        __syncthreads();
        if (keep) {
            int before = s_run[g];
            for (int ww = 0; ww < warp; ++ww) before += s_wcnt[ww][g];
            const long long dst = off + s_base[g] + before + rank_in_warp;
            if (dst < out.capacity) {
                const size_t o = (size_t)r * ws.sel_cap + i;
                const float4 a = ws.pt0[o], b = ws.pt1[o];
                out.xyz[dst * 3 + 0] = a.x; out.xyz[dst * 3 + 1] = a.y; out.xyz[dst * 3 + 2] = a.z;
                out.rgb[dst * 3 + 0] = b.x; out.rgb[dst * 3 + 1] = b.y; out.rgb[dst * 3 + 2] = b.z;
                out.err[dst] = a.w;
                if (P.collect_debug && out.dbg_matches) {
                    reinterpret_cast<float4*>(out.dbg_matches)[dst] = ws.dbgm[o];
                    if (out.dbg_cert) out.dbg_cert[dst] = b.w;
                }
            }
        }
        __syncthreads();
        if (tid < LDP_MAX_NN) {
            int add = 0;
            for (int ww = 0; ww < nwarp; ++ww) add += s_wcnt[ww][tid];
            s_run[tid] += add;
        }
        __syncthreads();
    }
}

}  // namespace ldp
