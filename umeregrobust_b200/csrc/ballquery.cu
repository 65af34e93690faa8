// pytorch3d.ops.ball_query replacement (evaluate.py:51, utils/loc_utils.py:383-384): the first K
// rows of p2, in row order, with dist2 < radius^2.  Same neighbourhood machinery as the fused
// moment kernel (neighbors.cuh); the selected entries are then sorted by row index in shared
// memory (bitonic) so that the output order is the reference's scan order.
#include "neighbors.cuh"

namespace ume {
namespace {

constexpr int kNT = 256;

struct BallQueryParams {
    GridView grid;
    const float* p1;      // (B,P1,3)
    const float* p2;      // (B,P2,3)
    int64_t* idx;         // (B,P1,K) or null
    float* dists;         // (B,P1,K) or null
    float* nn;            // (B,P1,K,3) or null
    int32_t* count;       // (B,P1) or null
    int P1, K, cap;
    float radius;
};

template <bool kFma>
__global__ void __launch_bounds__(kNT) ball_query_kernel(BallQueryParams p) {
    extern __shared__ float4 list[];
    __shared__ CollectSmem sm;
    const int q = blockIdx.x;
    const int b = q / p.P1;
    const int N = p.grid.N;
    const GridHeader h = p.grid.hdr[b];
    const int* cs = p.grid.cell_start + (size_t)b * (p.grid.cells_cap + 1);
    const float4* sorted_b = p.grid.sorted + (size_t)b * N;
    const float* p2b = p.p2 + (size_t)b * N * 3;
    const float kx = p.p1[(size_t)q * 3 + 0], ky = p.p1[(size_t)q * 3 + 1], kz = p.p1[(size_t)q * 3 + 2];
    const int K = p.K;

    const int used = collect_neighbors<kFma, kNT, true>(
        sm, list, p.cap, h, cs, sorted_b, N, kx, ky, kz, p.radius, K, [&](int len, int) {
            // K <= cap, so this runs exactly once with the complete neighbourhood
            int P = 1;
            while (P < len) P <<= 1;
            __syncthreads();
            for (int t = len + threadIdx.x; t < P; t += kNT) list[t] = make_float4(0.f, 0.f, 0.f, __int_as_float(0x7fffffff));
            __syncthreads();
            for (int k = 2; k <= P; k <<= 1) {
                for (int j = k >> 1; j > 0; j >>= 1) {
                    for (int t = threadIdx.x; t < P; t += kNT) {
                        const int u = t ^ j;
                        if (u > t) {
                            const float4 a = list[t], c = list[u];
                            const bool asc = (t & k) == 0;
                            if ((__float_as_int(a.w) > __float_as_int(c.w)) == asc) { list[t] = c; list[u] = a; }
                        }
                    }
                    __syncthreads();
                }
            }
            for (int k = threadIdx.x; k < K; k += kNT) {
                const size_t o = (size_t)q * K + k;
                if (k < len) {
                    const float4 e = list[k];
                    const int j = __float_as_int(e.w);
                    if (p.idx) p.idx[o] = j;
                    if (p.dists) p.dists[o] = dist2_ordered<kFma>(e.x, e.y, e.z);
                    if (p.nn) {
                        p.nn[o * 3 + 0] = __ldg(p2b + (size_t)j * 3 + 0);
                        p.nn[o * 3 + 1] = __ldg(p2b + (size_t)j * 3 + 1);
                        p.nn[o * 3 + 2] = __ldg(p2b + (size_t)j * 3 + 2);
                    }
                } else {
                    if (p.idx) p.idx[o] = -1;
                    if (p.dists) p.dists[o] = 0.f;
                    if (p.nn) { p.nn[o * 3 + 0] = 0.f; p.nn[o * 3 + 1] = 0.f; p.nn[o * 3 + 2] = 0.f; }
                }
            }
        });
    if (threadIdx.x == 0 && p.count) p.count[q] = used;
}

}  // namespace
}  // namespace ume

extern "C" size_t ume_ball_query_workspace_bytes(int B, int P1, int P2, int K) {
    (void)P1; (void)K;
    if (B <= 0 || P2 <= 0) return 0;
    return ume::grid_workspace_bytes(B, P2, ume::kCellsCap) + 256;
}

extern "C" int ume_ball_query_f32(const float* p1, const float* p2, int B, int P1, int P2, int K, float radius,
                                  unsigned flags, int64_t* idx, float* dists, float* nn, int32_t* count, void* ws,
                                  size_t ws_bytes, void* stream_) {
    using namespace ume;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    UME_REQUIRE(B >= 0 && P1 >= 0 && P2 >= 0, UME_ERR_BAD_ARG, "ume_ball_query_f32: negative size");
    if (B == 0 || P1 == 0) return UME_OK;
    UME_REQUIRE(p1 && p2, UME_ERR_BAD_ARG, "ume_ball_query_f32: null pointer");
    UME_REQUIRE(P2 >= 1, UME_ERR_BAD_ARG, "ume_ball_query_f32: empty cloud (P2 = 0)");
    UME_REQUIRE(K >= 1 && K <= 8192, UME_ERR_UNSUPPORTED, "ume_ball_query_f32: K = %d not in [1,8192]", K);
    UME_REQUIRE(P2 <= kMaxPoints, UME_ERR_UNSUPPORTED, "ume_ball_query_f32: P2 = %d > %d", P2, kMaxPoints);
    UME_REQUIRE((size_t)B * P1 < 0x7fffffffull, UME_ERR_UNSUPPORTED, "ume_ball_query_f32: B*P1 too large");
    UME_REQUIRE(ws && ws_bytes >= ume_ball_query_workspace_bytes(B, P1, P2, K), UME_ERR_WORKSPACE,
                "ume_ball_query_f32: workspace too small (%zu needed, %zu given)",
                ume_ball_query_workspace_bytes(B, P1, P2, K), ws_bytes);
    Workspace w(ws, ws_bytes);
    BallQueryParams p;
    int rc = grid_build(p2, p1, B, P2, P1, fabsf(radius), fabsf(radius) / ((flags & UME_FLAG_CELL_DIV2) ? 2.f : 1.f), kCellsCap, w, &p.grid, stream);
    if (rc != UME_OK) return rc;
    p.p1 = p1; p.p2 = p2; p.idx = idx; p.dists = dists; p.nn = nn; p.count = count;
    p.P1 = P1; p.K = K; p.radius = radius;
    int cap = 1024;
    while (cap < K) cap <<= 1;
    p.cap = cap;
    const size_t smem = (size_t)cap * sizeof(float4);
    const bool fma = (flags & UME_FLAG_FMA_DIST) != 0;
    auto kern = fma ? ball_query_kernel<true> : ball_query_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    UME_REQUIRE(e == cudaSuccess, UME_ERR_CUDA, "ball_query: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    ProfScope prof(UME_PROF_BALLQUERY, stream);
    kern<<<(unsigned)((size_t)B * P1), kNT, smem, stream>>>(p);
    count_launch();
    return check_launch("ball_query_kernel");
}
