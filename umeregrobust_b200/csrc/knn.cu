// pytorch3d.ops.knn_points(K=1) + knn_gather replacement (evaluate.py:272-275): nearest cloud row
// for every query, lower row index on ties, optional copy of that row's feature vector.
//
// One warp per query over the search grid: rings of cells around the query's cell are visited in
// order of increasing Chebyshev radius; the search stops once the best distance found is closer
// than any unvisited ring can be.  Distances use the same ordered fp32 arithmetic as ball_query.
#include "ume_common.cuh"

namespace ume {
namespace {

struct Knn1Params {
    GridView grid;
    const float* q;     // (B,P1,3)
    const float* x;     // (B,P2,U) or null
    int64_t* idx;
    float* d2;
    float* out;
    int P1, U;
};

template <bool kFma>
__global__ void __launch_bounds__(256) knn1_kernel(Knn1Params p) {
    const int lane = threadIdx.x & 31;
    const int64_t qi = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int b = blockIdx.y;
    if (qi >= p.P1) return;
    const int N = p.grid.N;
    const GridHeader h = p.grid.hdr[b];
    const int* cs = p.grid.cell_start + (size_t)b * (p.grid.cells_cap + 1);
    const float4* sorted_b = p.grid.sorted + (size_t)b * N;
    const size_t qo = ((size_t)b * p.P1 + qi);
    const float kx = p.q[qo * 3 + 0], ky = p.q[qo * 3 + 1], kz = p.q[qo * 3 + 2];
    const int cx = cell_coord(kx, h.ox, h.inv_s, h.nx), cy = cell_coord(ky, h.oy, h.inv_s, h.ny),
              cz = cell_coord(kz, h.oz, h.inv_s, h.nz);
    float best = INFINITY;
    int best_j = 0x7fffffff;
    const int max_ring = max(h.nx, max(h.ny, h.nz));
    for (int ring = 0; ring <= max_ring; ++ring) {
        // every point in a cell at Chebyshev ring >= `ring` is at least (ring-1)*s + (distance of
        // the query to its own cell wall) away; (ring-1)*s is a safe lower bound
        if (ring >= 2) {
            const float lb = (float)(ring - 1) * h.s * 0.9999f;
            if (best < lb * lb) break;
        }
        const int z0 = max(cz - ring, 0), z1 = min(cz + ring, h.nz - 1);
        const int y0 = max(cy - ring, 0), y1 = min(cy + ring, h.ny - 1);
        for (int iz = z0; iz <= z1; ++iz) {
            for (int iy = y0; iy <= y1; ++iy) {
                const bool shell_row = (abs(iz - cz) == ring) || (abs(iy - cy) == ring);
                const int base = (iz * h.ny + iy) * h.nx;
                // full x range on shell rows, only the two end cells otherwise
                int xr[2][2];
                int nr = 0;
                if (shell_row) {
                    xr[0][0] = max(cx - ring, 0); xr[0][1] = min(cx + ring, h.nx - 1); nr = 1;
                } else {
                    if (cx - ring >= 0) { xr[nr][0] = xr[nr][1] = cx - ring; ++nr; }
                    if (cx + ring <= h.nx - 1 && ring > 0) { xr[nr][0] = xr[nr][1] = cx + ring; ++nr; }
                }
                for (int r = 0; r < nr; ++r) {
                    const int s = cs[base + xr[r][0]], e = cs[base + xr[r][1] + 1];
                    for (int t = s + lane; t < e; t += 32) {
                        const float4 c = __ldg(&sorted_b[t]);
                        const float d = dist2_ordered<kFma>(__fsub_rn(kx, c.x), __fsub_rn(ky, c.y), __fsub_rn(kz, c.z));
                        const int j = __float_as_int(c.w);
                        if (d < best || (d == best && j < best_j)) { best = d; best_j = j; }
                    }
                }
            }
        }
        // warp-combine after every ring so that the stop test is uniform
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float od = __shfl_xor_sync(UME_FULL_MASK, best, o);
            const int oj = __shfl_xor_sync(UME_FULL_MASK, best_j, o);
            if (od < best || (od == best && oj < best_j)) { best = od; best_j = oj; }
        }
    }
    if (best_j == 0x7fffffff) best_j = 0;      // only when every distance is NaN
    if (lane == 0) {
        if (p.idx) p.idx[qo] = best_j;
        if (p.d2) p.d2[qo] = best;
    }
    if (p.out && p.x) {
        const float* src = p.x + ((size_t)b * N + best_j) * p.U;
        float* dst = p.out + qo * p.U;
        for (int u = lane; u < p.U; u += 32) dst[u] = __ldg(src + u);
    }
}

}  // namespace
}  // namespace ume

extern "C" size_t ume_knn1_workspace_bytes(int B, int P1, int P2) {
    (void)P1;
    if (B <= 0 || P2 <= 0) return 0;
    return ume::grid_workspace_bytes(B, P2, ume::kCellsCap) + 256;
}

extern "C" int ume_knn1_gather_f32(const float* q, const float* pcl, const float* x, int B, int P1, int P2, int U,
                                   unsigned flags, int64_t* idx, float* d2, float* out, void* ws, size_t ws_bytes,
                                   void* stream_) {
    using namespace ume;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    UME_REQUIRE(B >= 0 && P1 >= 0 && P2 >= 0, UME_ERR_BAD_ARG, "ume_knn1_gather_f32: negative size");
    if (B == 0 || P1 == 0) return UME_OK;
    UME_REQUIRE(q && pcl, UME_ERR_BAD_ARG, "ume_knn1_gather_f32: null pointer");
    UME_REQUIRE(P2 >= 1, UME_ERR_BAD_ARG, "ume_knn1_gather_f32: K = 1 > P2 = 0");
    UME_REQUIRE(P2 <= kMaxPoints, UME_ERR_UNSUPPORTED, "ume_knn1_gather_f32: P2 = %d > %d", P2, kMaxPoints);
    UME_REQUIRE(B <= 65535, UME_ERR_UNSUPPORTED, "ume_knn1_gather_f32: B > 65535");
    UME_REQUIRE(!out || (x && U >= 1), UME_ERR_BAD_ARG, "ume_knn1_gather_f32: out needs x and U >= 1");
    UME_REQUIRE(ws && ws_bytes >= ume_knn1_workspace_bytes(B, P1, P2), UME_ERR_WORKSPACE,
                "ume_knn1_gather_f32: workspace too small");
    Workspace w(ws, ws_bytes);
    Knn1Params p;
    // the grid covers the CLOUD here (a query's nearest row can be anywhere, queries outside the
    // box clamp to border cells), with the finest cells the table allows
    int rc = grid_build(pcl, pcl, B, P2, P2, /*expand=*/0.f, /*cell=*/0.f, kCellsCap, w, &p.grid, stream);
    if (rc != UME_OK) return rc;
    p.q = q; p.x = x; p.idx = idx; p.d2 = d2; p.out = out; p.P1 = P1; p.U = U;
    dim3 grid((unsigned)((P1 + 7) / 8), (unsigned)B);
    ProfScope prof(UME_PROF_KNN, stream);
    if (flags & UME_FLAG_FMA_DIST) knn1_kernel<true><<<grid, 256, 0, stream>>>(p);
    else knn1_kernel<false><<<grid, 256, 0, stream>>>(p);
    count_launch();
    return check_launch("knn1_kernel");
}
