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

// A (row-major 3x3) -> R = U diag(1,1,det(U Vh)) Vh, row-major.
UME_DEVI void rotation_from_cross_moment(const float A[9], float R[9]) {
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
    const V3 l3 = cross(l1, l2);
    const float lx[3] = {l1.x, l2.x, l3.x}, ly[3] = {l1.y, l2.y, l3.y}, lz[3] = {l1.z, l2.z, l3.z};
    const float rx[3] = {r1.x, r2.x, r3.x}, ry[3] = {r1.y, r2.y, r3.y}, rz[3] = {r1.z, r2.z, r3.z};
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        R[0] = fmaf(lx[k], rx[k], R[0]); R[1] = fmaf(lx[k], ry[k], R[1]); R[2] = fmaf(lx[k], rz[k], R[2]);
        R[3] = fmaf(ly[k], rx[k], R[3]); R[4] = fmaf(ly[k], ry[k], R[4]); R[5] = fmaf(ly[k], rz[k], R[5]);
        R[6] = fmaf(lz[k], rx[k], R[6]); R[7] = fmaf(lz[k], ry[k], R[7]); R[8] = fmaf(lz[k], rz[k], R[8]);
    }
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
