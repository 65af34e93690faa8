// All-pairs subspace distance with fused row arg-min — fp32 SIMT implementation (impl 0).
// Replaces utils/loc_utils.py:12-13 and evaluate.py:224 through D^2 = 4 - |Q1^T Q2|_F^2: instead of
// the reference's C^2-long projector vectors (2 n^2 C^2 flops) only the 4n x 4n Gram matrix of
// the basis vectors is formed (2 (4n)^2 C flops) and each 4x4 block is squared and summed in the
// epilogue.  The tcgen05 tensor-core implementation (impl 1) lives in cdist_tc.cu.
#include "ume_common.cuh"

namespace ume {

int cdist_tc_launch(const float* Qt1, const float* Qt2, int B, int n1, int n2, int C, float* D, int64_t* argmin,
                    float* dmin, void* ws, size_t ws_bytes, cudaStream_t stream);
size_t cdist_tc_workspace_bytes(int B, int n1, int n2, int C);
int cdist_tc_launch_split(const void* Qh1, const void* Qh2, int B, int n1, int n2, int C, float* D, int64_t* argmin,
                          float* dmin, cudaStream_t stream);

namespace {

constexpr int kTile = 32;                 // keypoints per tile side -> 128 basis vectors
constexpr int kRows = kTile * 4;
constexpr int kLd = kRows + 4;            // padded, keeps rows 16-byte aligned for LDS.128

// dst[c][r] = src[(row0 + r) * C + c], zero beyond nrows.  src rows are basis vectors (C floats).
UME_DEVI void load_tile_transposed(float* dst, const float* __restrict__ src, int nrows_valid, int C) {
    const int c4n = C >> 2;
    for (int i = threadIdx.x; i < kRows * c4n; i += blockDim.x) {
        const int r = i / c4n, c4 = i % c4n;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < nrows_valid) v = ldg_f4(src + (size_t)r * C + c4 * 4);
        dst[(c4 * 4 + 0) * kLd + r] = v.x;
        dst[(c4 * 4 + 1) * kLd + r] = v.y;
        dst[(c4 * 4 + 2) * kLd + r] = v.z;
        dst[(c4 * 4 + 3) * kLd + r] = v.w;
    }
}

__global__ void __launch_bounds__(256, 2) cdist_simt_kernel(const float* __restrict__ Q1, const float* __restrict__ Q2,
                                                           int n1, int n2, int C, float* __restrict__ D,
                                                           int64_t* __restrict__ argmin, float* __restrict__ dmin) {
    extern __shared__ float smem[];
    float* As = smem;                       // [C][kLd]
    float* Bs = smem + (size_t)C * kLd;     // [C][kLd]
    const int b = blockIdx.y;
    const int i0 = blockIdx.x * kTile;
    const int ti = threadIdx.x >> 4, tj = threadIdx.x & 15;
    const float* Q1b = Q1 + (size_t)b * n1 * 4 * C;
    const float* Q2b = Q2 + (size_t)b * n2 * 4 * C;

    load_tile_transposed(As, Q1b + (size_t)i0 * 4 * C, min(kTile, n1 - i0) * 4, C);

    float best[2] = {INFINITY, INFINITY};
    int best_j[2] = {0x7fffffff, 0x7fffffff};

    for (int j0 = 0; j0 < n2; j0 += kTile) {
        __syncthreads();                    // previous tile's Bs fully consumed (and As visible)
        load_tile_transposed(Bs, Q2b + (size_t)j0 * 4 * C, min(kTile, n2 - j0) * 4, C);
        __syncthreads();
        float acc[2][2][4][4];
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int m = 0; m < 4; ++m)
#pragma unroll
                    for (int k = 0; k < 4; ++k) acc[p][q][m][k] = 0.f;
#pragma unroll 4
        for (int c = 0; c < C; ++c) {
            const float4 a0 = *reinterpret_cast<const float4*>(As + (size_t)c * kLd + 4 * ti);
            const float4 a1 = *reinterpret_cast<const float4*>(As + (size_t)c * kLd + 64 + 4 * ti);
            const float4 b0 = *reinterpret_cast<const float4*>(Bs + (size_t)c * kLd + 4 * tj);
            const float4 b1 = *reinterpret_cast<const float4*>(Bs + (size_t)c * kLd + 64 + 4 * tj);
            const float av[2][4] = {{a0.x, a0.y, a0.z, a0.w}, {a1.x, a1.y, a1.z, a1.w}};
            const float bv[2][4] = {{b0.x, b0.y, b0.z, b0.w}, {b1.x, b1.y, b1.z, b1.w}};
#pragma unroll
            for (int p = 0; p < 2; ++p)
#pragma unroll
                for (int q = 0; q < 2; ++q)
#pragma unroll
                    for (int m = 0; m < 4; ++m)
#pragma unroll
                        for (int k = 0; k < 4; ++k) acc[p][q][m][k] = fmaf(av[p][m], bv[q][k], acc[p][q][m][k]);
        }
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const int i = i0 + ti + 16 * p;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int j = j0 + tj + 16 * q;
                float s = 0.f;
#pragma unroll
                for (int m = 0; m < 4; ++m)
#pragma unroll
                    for (int k = 0; k < 4; ++k) s = fmaf(acc[p][q][m][k], acc[p][q][m][k], s);
                const float d = sqrtf(fmaxf(4.f - s, 0.f));
                if (i < n1 && j < n2) {
                    if (D) D[((size_t)b * n1 + i) * n2 + j] = d;
                    if (d < best[p]) { best[p] = d; best_j[p] = j; }
                }
            }
        }
    }
    if (argmin || dmin) {
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            float bd = best[p];
            int bj = best_j[p];
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {            // the 16 tj lanes of a row share a half-warp
                const float od = __shfl_xor_sync(UME_FULL_MASK, bd, o);
                const int oj = __shfl_xor_sync(UME_FULL_MASK, bj, o);
                if (od < bd || (od == bd && oj < bj)) { bd = od; bj = oj; }
            }
            const int i = i0 + ti + 16 * p;
            if (tj == 0 && i < n1) {
                if (argmin) argmin[(size_t)b * n1 + i] = (bj == 0x7fffffff) ? 0 : bj;   // all-NaN row -> 0
                if (dmin) dmin[(size_t)b * n1 + i] = bd;
            }
        }
    }
}

}  // namespace
}  // namespace ume

extern "C" size_t ume_cdist_workspace_bytes(int B, int n1, int n2, int C, int impl) {
    if (impl == 1) return ume::cdist_tc_workspace_bytes(B, n1, n2, C);
    return 0;
}

extern "C" int ume_cdist_f32(const float* Qt1, const float* Qt2, int B, int n1, int n2, int C, int impl, float* D,
                             int64_t* argmin, float* dmin, void* ws, size_t ws_bytes, void* stream_) {
    using namespace ume;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    UME_REQUIRE(B >= 0 && n1 >= 0 && n2 >= 0, UME_ERR_BAD_ARG, "ume_cdist_f32: negative size");
    if (B == 0 || n1 == 0) return UME_OK;
    UME_REQUIRE(n2 >= 1 || (!argmin && !dmin), UME_ERR_BAD_ARG, "ume_cdist_f32: arg-min over an empty row (n2 = 0)");
    if (n2 == 0) return UME_OK;
    UME_REQUIRE(Qt1 && Qt2, UME_ERR_BAD_ARG, "ume_cdist_f32: null pointer");
    UME_REQUIRE(C >= 4 && C % 4 == 0 && C <= 128, UME_ERR_UNSUPPORTED, "ume_cdist_f32: C = %d must be a multiple of 4 in [4,128]", C);
    UME_REQUIRE(B <= 65535, UME_ERR_UNSUPPORTED, "ume_cdist_f32: B = %d > 65535", B);
    UME_REQUIRE(reinterpret_cast<uintptr_t>(Qt1) % 16 == 0 && reinterpret_cast<uintptr_t>(Qt2) % 16 == 0, UME_ERR_BAD_ARG,
                "ume_cdist_f32: descriptors not 16-byte aligned");
    if (impl == 1) return cdist_tc_launch(Qt1, Qt2, B, n1, n2, C, D, argmin, dmin, ws, ws_bytes, stream);
    UME_REQUIRE(impl == 0, UME_ERR_BAD_ARG, "ume_cdist_f32: impl = %d (0 = SIMT fp32, 1 = tcgen05)", impl);
    const size_t smem = (size_t)2 * C * kLd * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(cdist_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    UME_REQUIRE(e == cudaSuccess, UME_ERR_CUDA, "cdist: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    dim3 grid((unsigned)((n1 + kTile - 1) / kTile), (unsigned)B);
    ProfScope prof(UME_PROF_CDIST, stream);
    cdist_simt_kernel<<<grid, 256, smem, stream>>>(Qt1, Qt2, n1, n2, C, D, argmin, dmin);
    count_launch();
    return check_launch("cdist_simt_kernel");
}

extern "C" int ume_cdist_split_f16(const void* Qh1, const void* Qh2, int B, int n1, int n2, int C, float* D, int64_t* argmin,
                                   float* dmin, void* stream_) {
    using namespace ume;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    UME_REQUIRE(B >= 0 && n1 >= 0 && n2 >= 0, UME_ERR_BAD_ARG, "ume_cdist_split_f16: negative size");
    if (B == 0 || n1 == 0) return UME_OK;
    UME_REQUIRE(n2 >= 1 || (!argmin && !dmin), UME_ERR_BAD_ARG, "ume_cdist_split_f16: arg-min over an empty row (n2 = 0)");
    if (n2 == 0) return UME_OK;
    UME_REQUIRE(Qh1 && Qh2, UME_ERR_BAD_ARG, "ume_cdist_split_f16: null pointer");
    UME_REQUIRE(B <= 65535, UME_ERR_UNSUPPORTED, "ume_cdist_split_f16: B = %d > 65535", B);
    return cdist_tc_launch_split(Qh1, Qh2, B, n1, n2, C, D, argmin, dmin, stream);
}
