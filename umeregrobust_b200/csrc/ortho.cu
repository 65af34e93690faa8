// Subspace descriptor: orthonormal basis of the column space of each C x 4 UME matrix
// (replaces torch.linalg.qr at utils/loc_utils.py:9,11,338,341).
//
// One warp per matrix.  Lane l owns rows l, l+32, ... (a row is one float4 = [m0, mx, my, mz]), so
// the load is one coalesced 16-byte access per lane per 32 rows.  Classical Gram-Schmidt with one
// re-orthogonalisation pass (CGS2): all projections of a pass are independent warp reductions and
// overlap; orthogonality is at rounding level for the condition numbers seen here (first-order
// moments in absolute coordinates are strongly correlated with the zeroth-order column).
// Output is written transposed, (4, C): the K-major operand of the distance GEMM.
#include "ume_common.cuh"

#include <cuda_fp16.h>

namespace ume {
namespace {

constexpr int kMaxRowsPerLane = 8;   // C <= 256

UME_DEVI float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(UME_FULL_MASK, v, o);
    return v;
}

template <int RPL>
__global__ void __launch_bounds__(256) ortho_kernel(const float* __restrict__ F, int64_t nmat, int C,
                                                    float* __restrict__ Qt, __half* __restrict__ Qh,
                                                    int32_t* __restrict__ rank_out) {
    const int lane = threadIdx.x & 31;
    const int64_t mat = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (mat >= nmat) return;
    const float* Fm = F + mat * C * 4;
    float a[RPL][4];   // input columns (by row)
    float q[RPL][4];   // orthonormal columns built so far
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
        const int c = lane + 32 * r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < C) v = ldg_f4(Fm + (size_t)c * 4);
        a[r][0] = v.x; a[r][1] = v.y; a[r][2] = v.z; a[r][3] = v.w;
        q[r][0] = q[r][1] = q[r][2] = q[r][3] = 0.f;
    }
    // scale-free rank test: compare each residual with the norm of its own column
    int rank = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float v[RPL];
        float n0 = 0.f;
#pragma unroll
        for (int r = 0; r < RPL; ++r) { v[r] = a[r][j]; n0 = fmaf(v[r], v[r], n0); }
        n0 = warp_sum(n0);
        // pre-scale the column to unit length so that huge / tiny inputs do not overflow the dots
        const float inv0 = (n0 > 0.f && isfinite(n0)) ? rsqrtf(n0) : 0.f;
#pragma unroll
        for (int r = 0; r < RPL; ++r) v[r] *= inv0;
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
            float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < j; ++i) {
#pragma unroll
                for (int r = 0; r < RPL; ++r) d[i] = fmaf(q[r][i], v[r], d[i]);
            }
#pragma unroll
            for (int i = 0; i < j; ++i) d[i] = warp_sum(d[i]);
#pragma unroll
            for (int i = 0; i < j; ++i) {
#pragma unroll
                for (int r = 0; r < RPL; ++r) v[r] = fmaf(-d[i], q[r][i], v[r]);
            }
        }
        float n1 = 0.f;
#pragma unroll
        for (int r = 0; r < RPL; ++r) n1 = fmaf(v[r], v[r], n1);
        n1 = warp_sum(n1);
        // v had unit length before the projections: n1 is the squared sine of the angle between
        // column j and the span of the previous ones
        bool ok = (inv0 > 0.f) && (n1 > 1e-10f);
        if (!ok) {
            // rank-deficient: complete with the canonical unit vector least represented so far
            // (lowest row wins ties) -> all-zero input yields e0..e3 like LAPACK's Householder QR
            float best = 2.f;
            int best_c = 0;
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                const int c = lane + 32 * r;
                float s = 0.f;
#pragma unroll
                for (int i = 0; i < j; ++i) s = fmaf(q[r][i], q[r][i], s);
                if (c < C && s < best) { best = s; best_c = c; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ob = __shfl_xor_sync(UME_FULL_MASK, best, o);
                const int oc = __shfl_xor_sync(UME_FULL_MASK, best_c, o);
                if (ob < best || (ob == best && oc < best_c)) { best = ob; best_c = oc; }
            }
#pragma unroll
            for (int r = 0; r < RPL; ++r) v[r] = (lane + 32 * r == best_c) ? 1.f : 0.f;
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int i = 0; i < j; ++i) {
#pragma unroll
                    for (int r = 0; r < RPL; ++r) d[i] = fmaf(q[r][i], v[r], d[i]);
                }
#pragma unroll
                for (int i = 0; i < j; ++i) d[i] = warp_sum(d[i]);
#pragma unroll
                for (int i = 0; i < j; ++i) {
#pragma unroll
                    for (int r = 0; r < RPL; ++r) v[r] = fmaf(-d[i], q[r][i], v[r]);
                }
            }
            n1 = 0.f;
#pragma unroll
            for (int r = 0; r < RPL; ++r) n1 = fmaf(v[r], v[r], n1);
            n1 = warp_sum(n1);
        } else {
            ++rank;
        }
        const float inv1 = rsqrtf(n1);
#pragma unroll
        for (int r = 0; r < RPL; ++r) q[r][j] = v[r] * inv1;
    }
    if (Qt) {
        float* Qm = Qt + mat * 4 * C;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                const int c = lane + 32 * r;
                if (c < C) Qm[(size_t)j * C + c] = q[r][j];
            }
        }
    }
    if (Qh) {
        // the distance GEMM's operand (cdist_tc.cu): row j = [hi (C) | lo (C)] halves of 256 q, written here so
        // that no separate split pass (an extra read and write of the descriptors) is needed
        __half* Hm = Qh + mat * 4 * 2 * C;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                const int c = lane + 32 * r;
                if (c < C) {
                    const float v = q[r][j] * 256.f;
                    const __half hi = __float2half_rn(v);
                    Hm[(size_t)j * 2 * C + c] = hi;
                    Hm[(size_t)j * 2 * C + C + c] = __float2half_rn(v - __half2float(hi));
                }
            }
        }
    }
    if (rank_out && lane == 0) rank_out[mat] = rank;
}

// Dp[i] = scale * sqrt(max(8 - 2 |Q1_i^T Q2_i|_F^2, 0)); one warp per pair.
__global__ void __launch_bounds__(256) pair_dist_kernel(const float* __restrict__ Q1, const float* __restrict__ Q2,
                                                        int64_t nmat, int C, float scale, float* __restrict__ Dp) {
    const int lane = threadIdx.x & 31;
    const int64_t mat = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (mat >= nmat) return;
    const float* A = Q1 + mat * 4 * C;
    const float* Bm = Q2 + mat * 4 * C;
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
    for (int c = lane; c < C; c += 32) {
        float av[4], bv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { av[i] = __ldg(A + (size_t)i * C + c); bv[i] = __ldg(Bm + (size_t)i * C + c); }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) s[i][j] = fmaf(av[i], bv[j], s[i][j]);
    }
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float d = warp_sum(s[i][j]);
            tot = fmaf(d, d, tot);
        }
    if (lane == 0) Dp[mat] = scale * sqrtf(fmaxf(8.f - 2.f * tot, 0.f));
}

}  // namespace
}  // namespace ume

static int orthonormalize(const float* F, int64_t nmat, int C, float* Qt, void* Qh, int32_t* rank, void* stream_,
                         const char* who) {
    using namespace ume;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    UME_REQUIRE(nmat >= 0, UME_ERR_BAD_ARG, "%s: negative nmat", who);
    if (nmat == 0) return UME_OK;
    UME_REQUIRE(F && (Qt || Qh), UME_ERR_BAD_ARG, "%s: null pointer", who);
    UME_REQUIRE(C >= 4 && C <= 32 * kMaxRowsPerLane, UME_ERR_UNSUPPORTED,
                "%s: C = %d not in [4,256] (a C x 4 matrix needs C >= 4 for a rank-4 basis)", who, C);
    UME_REQUIRE(reinterpret_cast<uintptr_t>(F) % 16 == 0, UME_ERR_BAD_ARG, "%s: F not 16-byte aligned", who);
    const int wpb = 8;
    const int64_t blocks = (nmat + wpb - 1) / wpb;
    UME_REQUIRE(blocks < 0x7fffffffll, UME_ERR_UNSUPPORTED, "%s: too many matrices", who);
    const int rpl = (C + 31) / 32;
    __half* H = static_cast<__half*>(Qh);
    ProfScope prof(UME_PROF_ORTHO, stream);
    if (rpl == 1) ortho_kernel<1><<<(unsigned)blocks, wpb * 32, 0, stream>>>(F, nmat, C, Qt, H, rank);
    else if (rpl == 2) ortho_kernel<2><<<(unsigned)blocks, wpb * 32, 0, stream>>>(F, nmat, C, Qt, H, rank);
    else if (rpl <= 4) ortho_kernel<4><<<(unsigned)blocks, wpb * 32, 0, stream>>>(F, nmat, C, Qt, H, rank);
    else ortho_kernel<8><<<(unsigned)blocks, wpb * 32, 0, stream>>>(F, nmat, C, Qt, H, rank);
    count_launch();
    return check_launch("ortho_kernel");
}

extern "C" int ume_orthonormalize_f32(const float* F, int64_t nmat, int C, float* Qt, int32_t* rank, void* stream) {
    UME_REQUIRE(Qt || nmat == 0, UME_ERR_BAD_ARG, "ume_orthonormalize_f32: null pointer");
    return orthonormalize(F, nmat, C, Qt, nullptr, rank, stream, "ume_orthonormalize_f32");
}

extern "C" int ume_orthonormalize_split_f32(const float* F, int64_t nmat, int C, float* Qt, void* Qh, int32_t* rank,
                                            void* stream) {
    return orthonormalize(F, nmat, C, Qt, Qh, rank, stream, "ume_orthonormalize_split_f32");
}

extern "C" int ume_pair_dist_f32(const float* Qt1, const float* Qt2, int64_t nmat, int C, float scale, float* Dp,
                                 void* stream_) {
    using namespace ume;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    UME_REQUIRE(nmat >= 0, UME_ERR_BAD_ARG, "ume_pair_dist_f32: negative nmat");
    if (nmat == 0) return UME_OK;
    UME_REQUIRE(Qt1 && Qt2 && Dp, UME_ERR_BAD_ARG, "ume_pair_dist_f32: null pointer");
    UME_REQUIRE(C >= 1, UME_ERR_BAD_ARG, "ume_pair_dist_f32: C < 1");
    const int64_t blocks = (nmat + 7) / 8;
    UME_REQUIRE(blocks < 0x7fffffffll, UME_ERR_UNSUPPORTED, "ume_pair_dist_f32: too many matrices");
    pair_dist_kernel<<<(unsigned)blocks, 256, 0, stream>>>(Qt1, Qt2, nmat, C, scale, Dp);
    count_launch();
    return check_launch("pair_dist_kernel");
}
