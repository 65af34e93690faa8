// Batched closed-form rigid solve from pairs of UME matrices
// (replaces utils/loc_utils.py:292-335,346-350 `batch_estimate_transform_ume_old`).
//
// Eight lanes per hypothesis: a C x 4 matrix row is one float4 [m, x, y, z]; the lanes stride
// over the C rows twice (weighted centroids, then the 3x3 cross moment; the rows stay in L1),
// combine with width-8 shuffles, and every lane then runs the 3x3 decomposition redundantly.
// The SVD is a one-sided Jacobi on 3x3; the reference's  R = U diag(1,1,det(U Vh)) Vh  equals
// u1 v1^T + u2 v2^T + (u1 x u2)(v1 x v2)^T  for the two leading singular pairs, which needs
// neither the third singular vector nor a determinant.
#include "ume_common.cuh"

namespace ume {
namespace {

UME_DEVI float group8_sum(float v) {
    v += __shfl_xor_sync(UME_FULL_MASK, v, 4);
    v += __shfl_xor_sync(UME_FULL_MASK, v, 2);
    v += __shfl_xor_sync(UME_FULL_MASK, v, 1);
    return v;
}

struct V3 { float x, y, z; };
UME_DEVI float dot(const V3& a, const V3& b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
UME_DEVI V3 cross(const V3& a, const V3& b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
UME_DEVI V3 axpy(float s, const V3& a, const V3& b) { return {fmaf(s, a.x, b.x), fmaf(s, a.y, b.y), fmaf(s, a.z, b.z)}; }
UME_DEVI V3 scale(const V3& a, float s) { return {a.x * s, a.y * s, a.z * s}; }

// One Jacobi rotation making columns p,q of B orthogonal; V accumulates the rotations.
UME_DEVI bool rotate_pair(V3& bp, V3& bq, V3& vp, V3& vq) {
    const float alpha = dot(bp, bp), beta = dot(bq, bq), gamma = dot(bp, bq);
    if (!(fabsf(gamma) > 1e-8f * sqrtf(alpha * beta)) || gamma == 0.f) return false;
    const float zeta = (beta - alpha) / (2.f * gamma);
    const float t = copysignf(1.f, zeta) / (fabsf(zeta) + sqrtf(fmaf(zeta, zeta, 1.f)));
    const float c = rsqrtf(fmaf(t, t, 1.f)), s = c * t;
    const V3 nbp = {c * bp.x - s * bq.x, c * bp.y - s * bq.y, c * bp.z - s * bq.z};
    const V3 nbq = {s * bp.x + c * bq.x, s * bp.y + c * bq.y, s * bp.z + c * bq.z};
    const V3 nvp = {c * vp.x - s * vq.x, c * vp.y - s * vq.y, c * vp.z - s * vq.z};
    const V3 nvq = {s * vp.x + c * vq.x, s * vp.y + c * vq.y, s * vp.z + c * vq.z};
    bp = nbp; bq = nbq; vp = nvp; vq = nvq;
    return true;
}

UME_DEVI V3 normalized_or(const V3& a, const V3& fallback) {
    const float n2 = dot(a, a);
    if (n2 > 1e-30f && isfinite(n2)) return scale(a, rsqrtf(n2));
    return fallback;
}

// A (row-major 3x3) -> the proper frames L = [l1 l2 l1xl2], Rr = [r1 r2 r1xr2] with A = L diag(s1, s2, +-s3) Rr^T
// (the third "singular value" carries the sign of det(U Vh)): the reference's R = U diag(1,1,det(U Vh)) Vh is L Rr^T.
UME_DEVI void svd_frames(const float A[9], V3 l[3], V3 r[3]) {
    // pre-scale: Jacobi is scale invariant, but squares of tiny / huge entries are not
    float amax = 0.f;
#pragma unroll
    for (int i = 0; i < 9; ++i) amax = fmaxf(amax, fabsf(A[i]));
    const float sc = (amax > 0.f && isfinite(amax)) ? 1.f / amax : 0.f;
    V3 b0 = {A[0] * sc, A[3] * sc, A[6] * sc};      // columns of A
    V3 b1 = {A[1] * sc, A[4] * sc, A[7] * sc};
    V3 b2 = {A[2] * sc, A[5] * sc, A[8] * sc};
    V3 v0 = {1.f, 0.f, 0.f}, v1 = {0.f, 1.f, 0.f}, v2 = {0.f, 0.f, 1.f};
    for (int sweep = 0; sweep < 12; ++sweep) {
        bool any = rotate_pair(b0, b1, v0, v1);
        any |= rotate_pair(b0, b2, v0, v2);
        any |= rotate_pair(b1, b2, v1, v2);
        if (!any) break;
    }
    // order by singular value (column norm), descending
    float n0 = dot(b0, b0), n1 = dot(b1, b1), n2 = dot(b2, b2);
    if (n0 < n1) { V3 t = b0; b0 = b1; b1 = t; t = v0; v0 = v1; v1 = t; float f = n0; n0 = n1; n1 = f; }
    if (n0 < n2) { V3 t = b0; b0 = b2; b2 = t; t = v0; v0 = v2; v2 = t; float f = n0; n0 = n2; n2 = f; }
    if (n1 < n2) { V3 t = b1; b1 = b2; b2 = t; t = v1; v1 = v2; v2 = t; float f = n1; n1 = n2; n2 = f; }
    // right singular vectors: re-orthonormalise the accumulated rotations
    const V3 r1 = normalized_or(v0, V3{1.f, 0.f, 0.f});
    const V3 r2 = normalized_or(axpy(-dot(r1, v1), r1, v1), V3{0.f, 1.f, 0.f});
    const V3 r3 = cross(r1, r2);
    // left singular vectors of the two leading pairs; a vanishing singular value falls back to the
    // matching right vector (A = 0 then gives R = I, as LAPACK's U = V = I does)
    const V3 l1 = normalized_or(b0, r1);
    V3 l2 = normalized_or(axpy(-dot(l1, b1), l1, b1), V3{0.f, 0.f, 0.f});
    if (dot(l2, l2) == 0.f || n1 <= 1e-12f * n0) {
        V3 cand = axpy(-dot(l1, r2), l1, r2);
        if (dot(cand, cand) < 1e-6f) cand = axpy(-dot(l1, r3), l1, r3);
        l2 = normalized_or(cand, V3{0.f, 1.f, 0.f});
    }
    l[0] = l1; l[1] = l2; l[2] = cross(l1, l2);
    r[0] = r1; r[1] = r2; r[2] = r3;
}

// R = L Rr^T, row-major
UME_DEVI void rotation_from_frames(const V3 l[3], const V3 r[3], float R[9]) {
    const float lx[3] = {l[0].x, l[1].x, l[2].x}, ly[3] = {l[0].y, l[1].y, l[2].y}, lz[3] = {l[0].z, l[1].z, l[2].z};
    const float rx[3] = {r[0].x, r[1].x, r[2].x}, ry[3] = {r[0].y, r[1].y, r[2].y}, rz[3] = {r[0].z, r[1].z, r[2].z};
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        R[0] = fmaf(lx[k], rx[k], R[0]); R[1] = fmaf(lx[k], ry[k], R[1]); R[2] = fmaf(lx[k], rz[k], R[2]);
        R[3] = fmaf(ly[k], rx[k], R[3]); R[4] = fmaf(ly[k], ry[k], R[4]); R[5] = fmaf(ly[k], rz[k], R[5]);
        R[6] = fmaf(lz[k], rx[k], R[6]); R[7] = fmaf(lz[k], ry[k], R[7]); R[8] = fmaf(lz[k], rz[k], R[8]);
    }
}

// A (row-major 3x3) -> R = U diag(1,1,det(U Vh)) Vh, row-major.
UME_DEVI void rotation_from_cross_moment(const float A[9], float R[9]) {
    V3 l[3], r[3];
    svd_frames(A, l, r);
    rotation_from_frames(l, r, R);
}

__global__ void __launch_bounds__(256) rigid_kernel(const float* __restrict__ G, const float* __restrict__ H,
                                                    const int64_t* __restrict__ gi, const int64_t* __restrict__ hi,
                                                    const float* __restrict__ offG, const float* __restrict__ offH,
                                                    int64_t total, int nG, int nH, int nm, int C, float* __restrict__ T) {
    const int l8 = threadIdx.x & 7;
    const int64_t hyp = (int64_t)blockIdx.x * 32 + (threadIdx.x >> 3);
    const bool valid = hyp < total;
    const int64_t hc = valid ? hyp : 0;
    const int64_t b = hc / nm, i = hc % nm;
    int64_t ig = gi ? gi[hc] : i, ih = hi ? hi[hc] : i;
    ig = ig < 0 ? 0 : (ig >= nG ? nG - 1 : ig);            // out-of-range indices are clamped, never read OOB
    ih = ih < 0 ? 0 : (ih >= nH ? nH - 1 : ih);
    const float* Gm = G + ((size_t)b * nG + ig) * C * 4;
    const float* Hm = H + ((size_t)b * nH + ih) * C * 4;

    // pass 1 (utils/loc_utils.py:312-320): sums for the weighted centroids
    float s_mg2 = 0.f, s_mgmh = 0.f, s_gmg[3] = {0.f, 0.f, 0.f}, s_hmg[3] = {0.f, 0.f, 0.f};
    for (int c = l8; c < C; c += 8) {
        const float4 g = ldg_f4(Gm + (size_t)c * 4), h = ldg_f4(Hm + (size_t)c * 4);
        s_mg2 = fmaf(g.x, g.x, s_mg2);
        s_mgmh = fmaf(g.x, h.x, s_mgmh);
        s_gmg[0] = fmaf(g.y, g.x, s_gmg[0]); s_gmg[1] = fmaf(g.z, g.x, s_gmg[1]); s_gmg[2] = fmaf(g.w, g.x, s_gmg[2]);
        s_hmg[0] = fmaf(h.y, g.x, s_hmg[0]); s_hmg[1] = fmaf(h.z, g.x, s_hmg[1]); s_hmg[2] = fmaf(h.w, g.x, s_hmg[2]);
    }
    s_mg2 = group8_sum(s_mg2);
    s_mgmh = group8_sum(s_mgmh);
#pragma unroll
    for (int d = 0; d < 3; ++d) { s_gmg[d] = group8_sum(s_gmg[d]); s_hmg[d] = group8_sum(s_hmg[d]); }
    const float den_l = (s_mg2 + 1e-16f) + 1e-16f;          // :312 then :319
    const float den_r = s_mgmh + 1e-16f;                    // :320
    float wlc[3], wrc[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) { wlc[d] = s_gmg[d] / den_l; wrc[d] = s_hmg[d] / den_r; }

    // pass 2 (:322-326): A = (right^T left)^T = sum_c left_c (x) right_c
    float A[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int c = l8; c < C; c += 8) {
        const float4 g = ldg_f4(Gm + (size_t)c * 4), h = ldg_f4(Hm + (size_t)c * 4);
        const float lf[3] = {fmaf(-wlc[0], g.x, g.y), fmaf(-wlc[1], g.x, g.z), fmaf(-wlc[2], g.x, g.w)};
        const float rt[3] = {fmaf(-wrc[0], h.x, h.y), fmaf(-wrc[1], h.x, h.z), fmaf(-wrc[2], h.x, h.w)};
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int k = 0; k < 3; ++k) A[r * 3 + k] = fmaf(lf[r], rt[k], A[r * 3 + k]);
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) A[k] = group8_sum(A[k]);

    float R[9];
    rotation_from_cross_moment(A, R);

    // translation (:332), in double for the final combination of ~50 m magnitudes
    double b2[3];
#pragma unroll
    for (int j = 0; j < 3; ++j)
        b2[j] = (double)wrc[j] - ((double)wlc[0] * R[0 * 3 + j] + (double)wlc[1] * R[1 * 3 + j] + (double)wlc[2] * R[2 * 3 + j]);
    if (offG) {   // moments were relative to offG / offH: t = b2' + o_h - R^T o_g
        const float* og = offG + ((size_t)b * nG + ig) * 3;
        const float* oh = offH + ((size_t)b * nH + ih) * 3;
#pragma unroll
        for (int j = 0; j < 3; ++j)
            b2[j] += (double)oh[j] - ((double)og[0] * R[0 * 3 + j] + (double)og[1] * R[1 * 3 + j] + (double)og[2] * R[2 * 3 + j]);
    }
    if (valid && l8 < 4) {   // lane r writes row r of T: [R^T | b2 ; 0 0 0 1]  (:346-349)
        float4 row;
        if (l8 < 3) row = make_float4(R[0 * 3 + l8], R[1 * 3 + l8], R[2 * 3 + l8], (float)b2[l8 == 0 ? 0 : (l8 == 1 ? 1 : 2)]);
        else row = make_float4(0.f, 0.f, 0.f, 1.f);
        *reinterpret_cast<float4*>(T + (size_t)hyp * 16 + l8 * 4) = row;
    }
}

// Backward of the solve above for the training losses (loss.py:137-190 differentiates through
// batch_estimate_transform_ume_old): gT (nb,4,4) -> gG, gH (nb,C,4).  Same decomposition as the forward:
// eight lanes per hypothesis, the rows re-read from L1.  With A = L diag(s) Rr^T in proper frames (s3 signed) and
// R = L Rr^T:  dR = L Z Rr^T,  Z antisymmetric,  Z_ij = (dP_ij - dP_ji) / (s_i + s_j),  dP = L^T dA Rr  — no
// 1 / (s_i^2 - s_j^2) as in a generic SVD backward, so equal singular values are harmless.
__global__ void __launch_bounds__(256) rigid_backward_kernel(const float* __restrict__ G, const float* __restrict__ H,
                                                             const float* __restrict__ gT, int64_t total, int C,
                                                             float* __restrict__ gG, float* __restrict__ gH) {
    const int l8 = threadIdx.x & 7;
    const int64_t hyp = (int64_t)blockIdx.x * 32 + (threadIdx.x >> 3);
    const bool valid = hyp < total;
    const int64_t hc = valid ? hyp : 0;
    const float* Gm = G + (size_t)hc * C * 4;
    const float* Hm = H + (size_t)hc * C * 4;
    // forward, pass 1
    float s_mg2 = 0.f, s_mgmh = 0.f, s_gmg[3] = {0.f, 0.f, 0.f}, s_hmg[3] = {0.f, 0.f, 0.f};
    for (int c = l8; c < C; c += 8) {
        const float4 g = ldg_f4(Gm + (size_t)c * 4), h = ldg_f4(Hm + (size_t)c * 4);
        s_mg2 = fmaf(g.x, g.x, s_mg2);
        s_mgmh = fmaf(g.x, h.x, s_mgmh);
        s_gmg[0] = fmaf(g.y, g.x, s_gmg[0]); s_gmg[1] = fmaf(g.z, g.x, s_gmg[1]); s_gmg[2] = fmaf(g.w, g.x, s_gmg[2]);
        s_hmg[0] = fmaf(h.y, g.x, s_hmg[0]); s_hmg[1] = fmaf(h.z, g.x, s_hmg[1]); s_hmg[2] = fmaf(h.w, g.x, s_hmg[2]);
    }
    s_mg2 = group8_sum(s_mg2);
    s_mgmh = group8_sum(s_mgmh);
#pragma unroll
    for (int d = 0; d < 3; ++d) { s_gmg[d] = group8_sum(s_gmg[d]); s_hmg[d] = group8_sum(s_hmg[d]); }
    const float den_l = (s_mg2 + 1e-16f) + 1e-16f;
    const float den_r = s_mgmh + 1e-16f;
    float wl[3], wr[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) { wl[d] = s_gmg[d] / den_l; wr[d] = s_hmg[d] / den_r; }
    // forward, pass 2
    float A[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int c = l8; c < C; c += 8) {
        const float4 g = ldg_f4(Gm + (size_t)c * 4), h = ldg_f4(Hm + (size_t)c * 4);
        const float lf[3] = {fmaf(-wl[0], g.x, g.y), fmaf(-wl[1], g.x, g.z), fmaf(-wl[2], g.x, g.w)};
        const float rt[3] = {fmaf(-wr[0], h.x, h.y), fmaf(-wr[1], h.x, h.z), fmaf(-wr[2], h.x, h.w)};
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int k = 0; k < 3; ++k) A[r * 3 + k] = fmaf(lf[r], rt[k], A[r * 3 + k]);
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) A[k] = group8_sum(A[k]);
    V3 L[3], Rr[3];
    svd_frames(A, L, Rr);
    float R[9];
    rotation_from_frames(L, Rr, R);
    // signed singular values s_i = l_i^T A r_i
    float sv[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const V3 Ar = {fmaf(A[0], Rr[i].x, fmaf(A[1], Rr[i].y, A[2] * Rr[i].z)), fmaf(A[3], Rr[i].x, fmaf(A[4], Rr[i].y, A[5] * Rr[i].z)),
                       fmaf(A[6], Rr[i].x, fmaf(A[7], Rr[i].y, A[8] * Rr[i].z))};
        sv[i] = dot(L[i], Ar);
    }
    // T[:3,:3] = R^T, T[:3,3] = b2 = wr - wl R
    const float* gt = gT + (size_t)hc * 16;
    float gR[9], gb[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) gb[j] = __ldg(gt + j * 4 + 3);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int k = 0; k < 3; ++k) gR[r * 3 + k] = __ldg(gt + k * 4 + r) - wl[r] * gb[k];
    float gwl[3], gwr[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        gwl[r] = -(gb[0] * R[r * 3 + 0] + gb[1] * R[r * 3 + 1] + gb[2] * R[r * 3 + 2]);
        gwr[r] = gb[r];
    }
    // gZ = L^T gR Rr ; gP = antisymmetrised and divided ; gA = L gP Rr^T
    float gZ[9];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        // row vector l_i^T gR
        const float u0 = L[i].x * gR[0] + L[i].y * gR[3] + L[i].z * gR[6];
        const float u1 = L[i].x * gR[1] + L[i].y * gR[4] + L[i].z * gR[7];
        const float u2 = L[i].x * gR[2] + L[i].y * gR[5] + L[i].z * gR[8];
#pragma unroll
        for (int j = 0; j < 3; ++j) gZ[i * 3 + j] = u0 * Rr[j].x + u1 * Rr[j].y + u2 * Rr[j].z;
    }
    float gP[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const float smax = fmaxf(fabsf(sv[0]), 1e-30f);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = i + 1; j < 3; ++j) {
            const float den = sv[i] + sv[j];
            const float k = (fabsf(den) > 1e-7f * smax) ? (gZ[i * 3 + j] - gZ[j * 3 + i]) / den : 0.f;   // (rank <= 1: R is not unique)
            gP[i * 3 + j] = k;
            gP[j * 3 + i] = -k;
        }
    float gA[9];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const float lr[3] = {r == 0 ? L[0].x : (r == 1 ? L[0].y : L[0].z), r == 0 ? L[1].x : (r == 1 ? L[1].y : L[1].z),
                             r == 0 ? L[2].x : (r == 1 ? L[2].y : L[2].z)};
        // row r of L gP
        const float w0 = lr[0] * gP[0] + lr[1] * gP[3] + lr[2] * gP[6];
        const float w1 = lr[0] * gP[1] + lr[1] * gP[4] + lr[2] * gP[7];
        const float w2 = lr[0] * gP[2] + lr[1] * gP[5] + lr[2] * gP[8];
        gA[r * 3 + 0] = w0 * Rr[0].x + w1 * Rr[1].x + w2 * Rr[2].x;
        gA[r * 3 + 1] = w0 * Rr[0].y + w1 * Rr[1].y + w2 * Rr[2].y;
        gA[r * 3 + 2] = w0 * Rr[0].z + w1 * Rr[1].z + w2 * Rr[2].z;
    }
    // pass 3: the centroid weights also collect  - sum_c mg_c g_left_c  and  - sum_c mh_c g_right_c
    float t_l[3] = {0.f, 0.f, 0.f}, t_r[3] = {0.f, 0.f, 0.f};
    for (int c = l8; c < C; c += 8) {
        const float4 g = ldg_f4(Gm + (size_t)c * 4), h = ldg_f4(Hm + (size_t)c * 4);
        const float lf[3] = {fmaf(-wl[0], g.x, g.y), fmaf(-wl[1], g.x, g.z), fmaf(-wl[2], g.x, g.w)};
        const float rt[3] = {fmaf(-wr[0], h.x, h.y), fmaf(-wr[1], h.x, h.z), fmaf(-wr[2], h.x, h.w)};
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const float gl = gA[r * 3 + 0] * rt[0] + gA[r * 3 + 1] * rt[1] + gA[r * 3 + 2] * rt[2];      // g_left_c[r]
            const float gr = gA[0 * 3 + r] * lf[0] + gA[1 * 3 + r] * lf[1] + gA[2 * 3 + r] * lf[2];      // g_right_c[r]
            t_l[r] = fmaf(g.x, gl, t_l[r]);
            t_r[r] = fmaf(h.x, gr, t_r[r]);
        }
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) { gwl[r] -= group8_sum(t_l[r]); gwr[r] -= group8_sum(t_r[r]); }
    const float nl_gwl = s_gmg[0] * gwl[0] + s_gmg[1] * gwl[1] + s_gmg[2] * gwl[2];
    const float nr_gwr = s_hmg[0] * gwr[0] + s_hmg[1] * gwr[1] + s_hmg[2] * gwr[2];
    const float il = 1.f / den_l, ir = 1.f / den_r;
    // pass 4: the rows' gradients
    if (valid) {
        for (int c = l8; c < C; c += 8) {
            const float4 g = ldg_f4(Gm + (size_t)c * 4), h = ldg_f4(Hm + (size_t)c * 4);
            const float lf[3] = {fmaf(-wl[0], g.x, g.y), fmaf(-wl[1], g.x, g.z), fmaf(-wl[2], g.x, g.w)};
            const float rt[3] = {fmaf(-wr[0], h.x, h.y), fmaf(-wr[1], h.x, h.z), fmaf(-wr[2], h.x, h.w)};
            float gl[3], gr[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                gl[r] = gA[r * 3 + 0] * rt[0] + gA[r * 3 + 1] * rt[1] + gA[r * 3 + 2] * rt[2];
                gr[r] = gA[0 * 3 + r] * lf[0] + gA[1 * 3 + r] * lf[1] + gA[2 * 3 + r] * lf[2];
            }
            const float gv[3] = {g.y, g.z, g.w}, hv[3] = {h.y, h.z, h.w};
            float g_mg = -2.f * g.x * nl_gwl * il * il - h.x * nr_gwr * ir * ir;
            float g_mh = -g.x * nr_gwr * ir * ir;
            float og[3], oh[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                g_mg += -gl[r] * wl[r] + gv[r] * gwl[r] * il + hv[r] * gwr[r] * ir;
                g_mh += -gr[r] * wr[r];
                og[r] = gl[r] + g.x * gwl[r] * il;
                oh[r] = gr[r] + g.x * gwr[r] * ir;
            }
            *reinterpret_cast<float4*>(gG + ((size_t)hyp * C + c) * 4) = make_float4(g_mg, og[0], og[1], og[2]);
            *reinterpret_cast<float4*>(gH + ((size_t)hyp * C + c) * 4) = make_float4(g_mh, oh[0], oh[1], oh[2]);
        }
    }
}

// utils/eval_utils.py:60-76: one thread per pair of rotations.
__global__ void rotation_error_kernel(const float* __restrict__ R, const float* __restrict__ Rh, int64_t n, int sr, int sh,
                                      float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* a = R + i * sr;
    const float* b = Rh + i * sh;
    const int la = (sr >= 12) ? 4 : 3, lb = (sh >= 12) ? 4 : 3;   // row pitch: 3 (packed 3x3) or 4 (block of a 4x4 transform)
    // trace(R_hat R^T) = sum_jk R_hat[j][k] R[j][k]   (the einsum of :66-67 after the matmul of :65)
    float tr = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        float row = 0.f;                                   // (R_hat R^T)[j][j], a 3-term dot as torch.matmul forms it
#pragma unroll
        for (int k = 0; k < 3; ++k) row = fmaf(b[j * lb + k], a[j * la + k], row);
        tr += row;
    }
    tr = fminf(fmaxf(tr, -1.f), 3.f);
    out[i] = acosf((tr - 1.f) * 0.5f) * (180.f / 3.14159265358979323846f);
}

}  // namespace
}  // namespace ume

extern "C" int ume_rotation_error_deg_f32(const float* R, const float* R_hat, int64_t n, int stride_R, int stride_R_hat,
                                          float* out, void* stream_) {
    using namespace ume;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    UME_REQUIRE(n >= 0, UME_ERR_BAD_ARG, "ume_rotation_error_deg_f32: negative size");
    if (n == 0) return UME_OK;
    UME_REQUIRE(R && R_hat && out, UME_ERR_BAD_ARG, "ume_rotation_error_deg_f32: null pointer");
    UME_REQUIRE((stride_R == 9 || stride_R == 16) && (stride_R_hat == 9 || stride_R_hat == 16), UME_ERR_BAD_ARG,
                "ume_rotation_error_deg_f32: stride must be 9 (packed 3x3) or 16 (4x4 transforms)");
    UME_REQUIRE((n + 255) / 256 < 0x7fffffffll, UME_ERR_UNSUPPORTED, "ume_rotation_error_deg_f32: too many rotations");
    rotation_error_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(R, R_hat, n, stride_R, stride_R_hat, out);
    count_launch();
    return check_launch("rotation_error_kernel");
}

extern "C" int ume_rigid_solve_f32(const float* G, const float* H, const int64_t* gi, const int64_t* hi,
                                   const float* offG, const float* offH, int B, int nG, int nH, int nm, int C,
                                   float* T, void* stream_) {
    using namespace ume;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    UME_REQUIRE(B >= 0 && nm >= 0, UME_ERR_BAD_ARG, "ume_rigid_solve_f32: negative size");
    if (B == 0 || nm == 0) return UME_OK;
    UME_REQUIRE(G && H && T, UME_ERR_BAD_ARG, "ume_rigid_solve_f32: null pointer");
    UME_REQUIRE(nG >= 1 && nH >= 1 && C >= 1, UME_ERR_BAD_ARG, "ume_rigid_solve_f32: nG, nH, C must be >= 1");
    UME_REQUIRE((offG == nullptr) == (offH == nullptr), UME_ERR_BAD_ARG, "ume_rigid_solve_f32: offG and offH go together");
    UME_REQUIRE(gi || nm <= nG, UME_ERR_BAD_ARG, "ume_rigid_solve_f32: nm > nG without an index");
    UME_REQUIRE(hi || nm <= nH, UME_ERR_BAD_ARG, "ume_rigid_solve_f32: nm > nH without an index");
    UME_REQUIRE(reinterpret_cast<uintptr_t>(G) % 16 == 0 && reinterpret_cast<uintptr_t>(H) % 16 == 0 &&
                    reinterpret_cast<uintptr_t>(T) % 16 == 0, UME_ERR_BAD_ARG, "ume_rigid_solve_f32: pointers not 16-byte aligned");
    const int64_t total = (int64_t)B * nm;
    const int64_t blocks = (total + 31) / 32;
    UME_REQUIRE(blocks < 0x7fffffffll, UME_ERR_UNSUPPORTED, "ume_rigid_solve_f32: too many hypotheses");
    ProfScope prof(UME_PROF_RIGID, stream);
    rigid_kernel<<<(unsigned)blocks, 256, 0, stream>>>(G, H, gi, hi, offG, offH, total, nG, nH, nm, C, T);
    count_launch();
    return check_launch("rigid_kernel");
}

extern "C" int ume_rigid_solve_backward_f32(const float* G, const float* H, const float* gT, int64_t nb, int C, float* gG,
                                            float* gH, void* stream_) {
    using namespace ume;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    UME_REQUIRE(nb >= 0, UME_ERR_BAD_ARG, "ume_rigid_solve_backward_f32: negative size");
    if (nb == 0) return UME_OK;
    UME_REQUIRE(G && H && gT && gG && gH, UME_ERR_BAD_ARG, "ume_rigid_solve_backward_f32: null pointer");
    UME_REQUIRE(C >= 1, UME_ERR_BAD_ARG, "ume_rigid_solve_backward_f32: C must be >= 1");
    UME_REQUIRE(reinterpret_cast<uintptr_t>(G) % 16 == 0 && reinterpret_cast<uintptr_t>(H) % 16 == 0 &&
                    reinterpret_cast<uintptr_t>(gG) % 16 == 0 && reinterpret_cast<uintptr_t>(gH) % 16 == 0,
                UME_ERR_BAD_ARG, "ume_rigid_solve_backward_f32: pointers not 16-byte aligned");
    const int64_t blocks = (nb + 31) / 32;
    UME_REQUIRE(blocks < 0x7fffffffll, UME_ERR_UNSUPPORTED, "ume_rigid_solve_backward_f32: too many hypotheses");
    rigid_backward_kernel<<<(unsigned)blocks, 256, 0, stream>>>(G, H, gT, nb, C, gG, gH);
    count_launch();
    return check_launch("rigid_backward_kernel");
}
