// Backward of the subspace distance for the training losses: loss.py:84-118 differentiates
// through utils/loc_utils.py:8-15 (thin QR -> projector P = Q Q^T -> cdist / sqrt 2).  The reference lets autograd
// walk torch.linalg.qr, the (B,n,C,C) projectors and cdist; here the chain is closed-form on the 4x4 Gram blocks
// the forward already works with, and no projector is ever formed:
//
//   D_ij = sqrt(4 - |S_ij|_F^2),  S_ij = Q1_i^T Q2_j  (4x4)      =>  dD_ij = -tr(dP1_i P2_j + P1_i dP2_j) / (2 D_ij)
//   gradient wrt the projector P1_i:  G_i = sum_j w_ij P2_j,  w_ij = -gD_ij / (2 D_ij)     (0 where D_ij = 0, as torch.cdist)
//   P = F (F^T F)^-1 F^T  =>  gF = 2 (I - P) G (F^+)^T = 2 (M - Q (Q^T M)) R^-T,   M = G Q,  R = Q^T F  (F = Q R)
//   M1_i = sum_j w_ij Q2_j S_ij^T   and   M2_j = sum_i w_ij Q1_i S_ij :
// two GEMM-shaped passes over the same Gram blocks as the forward (cdist_bwd_m_kernel, once per side), then a
// warp-per-matrix epilogue with the 4x4 triangular solve (proj_bwd_kernel).  Any orthonormal basis of the column
// space gives the same P, hence the same gradient: the bases are the forward's (ume_orthonormalize_f32).
#include "ume_common.cuh"

namespace ume {
namespace {

constexpr int kTI = 16, kTJ = 16;      // keypoints of side A / side B per tile

UME_DEVI float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(UME_FULL_MASK, v, o);
    return v;
}

// M[b,i,a,c] = sum_j w(i,j) sum_k QB[b,j,k,c] S_ij[a][k],  S_ij[a][k] = sum_c QA[b,i,a,c] QB[b,j,k,c].
// D / gD are (B,nA,nB) (transposed = 0) or (B,nB,nA) (transposed = 1: side A is the distance matrix's column side).
// 256 threads: phase 1 thread = (i, j) of the tile -> its 4x4 block, scaled by w; phase 2 thread = (i, 16 channels
// apart) -> accumulates M over the tile's j.  CPT = channels per thread in phase 2 (C <= 16 CPT).
template <int CPT>
__global__ void __launch_bounds__(256) cdist_bwd_m_kernel(const float* __restrict__ QA, const float* __restrict__ QB,
                                                          const float* __restrict__ D, const float* __restrict__ gD, int transposed,
                                                          int nA, int nB, int C, float* __restrict__ M) {
    extern __shared__ float bw_smem[];
    const int RS = C + 1;                    // row stride (one basis vector)
    const int KS = 4 * RS + 1;               // keypoint stride: 16 consecutive keypoints fall into distinct banks
    float* sA = bw_smem;                     // [kTI][KS]
    float* sB = sA + kTI * KS;               // [kTJ][KS]
    float* sT = sB + kTJ * KS;               // [kTI][kTJ][17]
    const int b = blockIdx.y;
    const int i0 = blockIdx.x * kTI;
    const int t = threadIdx.x;
    const int il = t >> 4, jl = t & 15;
    const float* QAb = QA + (size_t)b * nA * 4 * C;
    const float* QBb = QB + (size_t)b * nB * 4 * C;
    for (int e = t; e < kTI * 4 * C; e += 256) {
        const int kp = e / (4 * C), rem = e % (4 * C);
        sA[kp * KS + (rem / C) * RS + rem % C] = (i0 + kp < nA) ? __ldg(QAb + (size_t)(i0 + kp) * 4 * C + rem) : 0.f;
    }
    float acc[CPT][4];
#pragma unroll
    for (int u = 0; u < CPT; ++u) acc[u][0] = acc[u][1] = acc[u][2] = acc[u][3] = 0.f;
    for (int j0 = 0; j0 < nB; j0 += kTJ) {
        __syncthreads();                     // the previous tile's readers are done (and sA is visible)
        for (int e = t; e < kTJ * 4 * C; e += 256) {
            const int kp = e / (4 * C), rem = e % (4 * C);
            sB[kp * KS + (rem / C) * RS + rem % C] = (j0 + kp < nB) ? __ldg(QBb + (size_t)(j0 + kp) * 4 * C + rem) : 0.f;
        }
        __syncthreads();
        // phase 1: the 4x4 Gram block of (il, jl), times w
        float S[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a) S[a][0] = S[a][1] = S[a][2] = S[a][3] = 0.f;
        const float* pa = sA + il * KS;
        const float* pb = sB + jl * KS;
        for (int c = 0; c < C; ++c) {
            float av[4], bv[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) { av[a] = pa[a * RS + c]; bv[a] = pb[a * RS + c]; }
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int k = 0; k < 4; ++k) S[a][k] = fmaf(av[a], bv[k], S[a][k]);
        }
        float w = 0.f;
        const int i = i0 + il, j = j0 + jl;
        if (i < nA && j < nB) {
            const size_t at = transposed ? ((size_t)b * nB + j) * nA + i : ((size_t)b * nA + i) * nB + j;
            const float d = __ldg(D + at);
            w = (d > 0.f) ? -__ldg(gD + at) / (2.f * d) : 0.f;
        }
        float* pt = sT + (il * kTJ + jl) * 17;
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int k = 0; k < 4; ++k) pt[a * 4 + k] = w * S[a][k];
        __syncthreads();
        // phase 2: M[il][a][c] += sum_j sum_k QB[j][k][c] T[il][j][a][k] for c = jl, jl + 16, ...
#pragma unroll 4
        for (int jj = 0; jj < kTJ; ++jj) {
            const float* tt = sT + (il * kTJ + jj) * 17;
            float tv[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) tv[e] = tt[e];
#pragma unroll
            for (int u = 0; u < CPT; ++u) {
                const int c = jl + 16 * u;
                if (c < C) {
                    const float* qb = sB + jj * KS + c;
                    const float q0 = qb[0], q1 = qb[RS], q2 = qb[2 * RS], q3 = qb[3 * RS];
#pragma unroll
                    for (int a = 0; a < 4; ++a)
                        acc[u][a] = fmaf(q3, tv[a * 4 + 3], fmaf(q2, tv[a * 4 + 2], fmaf(q1, tv[a * 4 + 1], fmaf(q0, tv[a * 4 + 0], acc[u][a]))));
                }
            }
        }
    }
    const int i = i0 + il;
    if (i < nA) {
        float* Mo = M + ((size_t)b * nA + i) * 4 * C;
#pragma unroll
        for (int u = 0; u < CPT; ++u) {
            const int c = jl + 16 * u;
            if (c < C) {
#pragma unroll
                for (int a = 0; a < 4; ++a) Mo[(size_t)a * C + c] = acc[u][a];
            }
        }
    }
}

// gF = 2 (M - Q (Q^T M)) R^-T with R = Q^T F.  One warp per matrix; lane owns rows lane, lane + 32, ...
template <int RPL>
__global__ void __launch_bounds__(256) proj_bwd_kernel(const float* __restrict__ F, const float* __restrict__ Qt,
                                                       const float* __restrict__ M, int64_t nmat, int C, float* __restrict__ gF) {
    const int lane = threadIdx.x & 31;
    const int64_t mat = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (mat >= nmat) return;
    float f[RPL][4], q[RPL][4], m[RPL][4];
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
        const int c = lane + 32 * r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < C) v = ldg_f4(F + ((size_t)mat * C + c) * 4);
        f[r][0] = v.x; f[r][1] = v.y; f[r][2] = v.z; f[r][3] = v.w;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            q[r][a] = (c < C) ? __ldg(Qt + ((size_t)mat * 4 + a) * C + c) : 0.f;
            m[r][a] = (c < C) ? __ldg(M + ((size_t)mat * 4 + a) * C + c) : 0.f;
        }
    }
    float QM[4][4], R[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float s = 0.f, rr = 0.f;
#pragma unroll
            for (int r = 0; r < RPL; ++r) { s = fmaf(q[r][a], m[r][k], s); rr = fmaf(q[r][a], f[r][k], rr); }
            QM[a][k] = warp_sum(s);
            R[a][k] = warp_sum(rr);
        }
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
        const int c = lane + 32 * r;
        float y[4], x[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            y[k] = 2.f * (m[r][k] - (q[r][0] * QM[0][k] + q[r][1] * QM[1][k] + q[r][2] * QM[2][k] + q[r][3] * QM[3][k]));
        // x R^T = y with R upper triangular: sum_{k >= a} x[k] R[a][k] = y[a], back substitution from a = 3
#pragma unroll
        for (int a = 3; a >= 0; --a) {
            float s = y[a];
#pragma unroll
            for (int k = a + 1; k < 4; ++k) s = fmaf(-R[a][k], x[k], s);
            x[a] = s / R[a][a];
        }
        if (c < C) *reinterpret_cast<float4*>(gF + ((size_t)mat * C + c) * 4) = make_float4(x[0], x[1], x[2], x[3]);
    }
}

template <int CPT>
int launch_m(const float* QA, const float* QB, const float* D, const float* gD, int transposed, int B, int nA, int nB, int C,
             float* M, cudaStream_t stream) {
    const size_t smem = ((size_t)(kTI + kTJ) * (4 * (C + 1) + 1) + (size_t)kTI * kTJ * 17) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(cdist_bwd_m_kernel<CPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    UME_REQUIRE(e == cudaSuccess, UME_ERR_CUDA, "cdist_bwd_m_kernel: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    dim3 grid((unsigned)((nA + kTI - 1) / kTI), (unsigned)B);
    cdist_bwd_m_kernel<CPT><<<grid, 256, smem, stream>>>(QA, QB, D, gD, transposed, nA, nB, C, M);
    count_launch();
    return check_launch("cdist_bwd_m_kernel");
}

int launch_m_any(const float* QA, const float* QB, const float* D, const float* gD, int transposed, int B, int nA, int nB, int C,
                 float* M, cudaStream_t stream) {
    if (C <= 16) return launch_m<1>(QA, QB, D, gD, transposed, B, nA, nB, C, M, stream);
    if (C <= 32) return launch_m<2>(QA, QB, D, gD, transposed, B, nA, nB, C, M, stream);
    if (C <= 64) return launch_m<4>(QA, QB, D, gD, transposed, B, nA, nB, C, M, stream);
    return launch_m<8>(QA, QB, D, gD, transposed, B, nA, nB, C, M, stream);
}

int launch_proj(const float* F, const float* Qt, const float* M, int64_t nmat, int C, float* gF, cudaStream_t stream) {
    const int64_t blocks = (nmat + 7) / 8;
    UME_REQUIRE(blocks < 0x7fffffffll, UME_ERR_UNSUPPORTED, "proj_bwd_kernel: too many matrices");
    const int rpl = (C + 31) / 32;
    if (rpl == 1) proj_bwd_kernel<1><<<(unsigned)blocks, 256, 0, stream>>>(F, Qt, M, nmat, C, gF);
    else if (rpl == 2) proj_bwd_kernel<2><<<(unsigned)blocks, 256, 0, stream>>>(F, Qt, M, nmat, C, gF);
    else proj_bwd_kernel<4><<<(unsigned)blocks, 256, 0, stream>>>(F, Qt, M, nmat, C, gF);
    count_launch();
    return check_launch("proj_bwd_kernel");
}

}  // namespace
}  // namespace ume

extern "C" size_t ume_cdist_backward_workspace_bytes(int B, int n1, int n2, int C) {
    if (B <= 0 || n1 <= 0 || n2 <= 0 || C <= 0) return 0;
    return ume::align_up((size_t)B * n1 * 4 * C * sizeof(float), 256) + ume::align_up((size_t)B * n2 * 4 * C * sizeof(float), 256) + 256;
}

extern "C" int ume_cdist_backward_f32(const float* F1, const float* F2, const float* Qt1, const float* Qt2, const float* D,
                                      const float* gD, int B, int n1, int n2, int C, float* gF1, float* gF2, void* ws,
                                      size_t ws_bytes, void* stream_) {
    using namespace ume;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    UME_REQUIRE(B >= 0 && n1 >= 0 && n2 >= 0, UME_ERR_BAD_ARG, "ume_cdist_backward_f32: negative size");
    if (B == 0 || n1 == 0 || n2 == 0) return UME_OK;
    UME_REQUIRE(F1 && F2 && Qt1 && Qt2 && D && gD && gF1 && gF2, UME_ERR_BAD_ARG, "ume_cdist_backward_f32: null pointer");
    UME_REQUIRE(C >= 4 && C <= 128, UME_ERR_UNSUPPORTED, "ume_cdist_backward_f32: C = %d not in [4,128]", C);
    UME_REQUIRE(B <= 65535, UME_ERR_UNSUPPORTED, "ume_cdist_backward_f32: more than 65535 batch entries");
    UME_REQUIRE(reinterpret_cast<uintptr_t>(F1) % 16 == 0 && reinterpret_cast<uintptr_t>(F2) % 16 == 0 &&
                    reinterpret_cast<uintptr_t>(gF1) % 16 == 0 && reinterpret_cast<uintptr_t>(gF2) % 16 == 0,
                UME_ERR_BAD_ARG, "ume_cdist_backward_f32: pointers not 16-byte aligned");
    UME_REQUIRE(ws && ws_bytes >= ume_cdist_backward_workspace_bytes(B, n1, n2, C), UME_ERR_WORKSPACE,
                "ume_cdist_backward_f32: workspace too small");
    Workspace w(ws, ws_bytes);
    float* M1 = w.take<float>((size_t)B * n1 * 4 * C);
    float* M2 = w.take<float>((size_t)B * n2 * 4 * C);
    int rc = launch_m_any(Qt1, Qt2, D, gD, 0, B, n1, n2, C, M1, stream);
    if (rc != UME_OK) return rc;
    rc = launch_m_any(Qt2, Qt1, D, gD, 1, B, n2, n1, C, M2, stream);
    if (rc != UME_OK) return rc;
    rc = launch_proj(F1, Qt1, M1, (int64_t)B * n1, C, gF1, stream);
    if (rc != UME_OK) return rc;
    return launch_proj(F2, Qt2, M2, (int64_t)B * n2, C, gF2, stream);
}
