// Hypothesis selection by feature correlation: the reference's FeatureCorrelator
// (utils/loc_utils.py:579-681, driven by evaluate.py:20-47,287-296), general K nearest neighbours
// (pytorch3d.ops.knn_points as used at utils/loc_utils.py:580,623) and feature_spatial_var (:579-585).
//
// Everything is built on one device routine: an exact K-nearest search of one query per THREAD over
// the search grid (rings of cells of growing Chebyshev radius, stop when the K-th best distance beats
// the next ring's lower bound), the K best kept sorted by (dist2, row) in a small per-thread array —
// ties go to the lower row, as a row-order scan with strict '<' does in pytorch3d.
//
// Correlation score (pc_corr, utils/loc_utils.py:592-619):
//   score[h] = (1/Ns) sum_i sum_{k<K} cauchy(|T_h p_i - q_nn(i,k)|; sigma) <wf_src_i, wf_tgt_nn(i,k)>
// One thread per (source point, hypothesis): the source points are taken in CELL ORDER of their own
// grid, so the 128 threads of a CTA are spatial neighbours; under the same rigid transform they land
// in the same few target cells and share candidates and target feature rows through L1.  The source
// feature (C floats) stays in registers across the thread's loop over hypotheses.
#include "ume_common.cuh"
#include "corr_tile.cuh"

#include <algorithm>

namespace ume {
namespace {

constexpr int kCorrThreads = 128;

// Sorted insertion into the K best (ascending by (d, j)); n = current fill.
template <int KMAX>
UME_DEVI void topk_insert(float (&bd)[KMAX], int (&bj)[KMAX], int& n, int K, float d, int j) {
    if (n == K) {
        const float wd = bd[K - 1];
        if (!(d < wd || (d == wd && j < bj[K - 1]))) return;
    }
    int pos = (n < K) ? n : K - 1;
    while (pos > 0) {
        const float pd = bd[pos - 1];
        const int pj = bj[pos - 1];
        if (!(d < pd || (d == pd && j < pj))) break;
        bd[pos] = pd;
        bj[pos] = pj;
        --pos;
    }
    bd[pos] = d;
    bj[pos] = j;
    if (n < K) ++n;
}

// K best in a per-thread array (local memory), any K <= KMAX.
template <int KMAX>
struct ArrayTopK {
    float bd[KMAX];
    int bj[KMAX];
    int n, K;
    UME_DEVI void reset(int k) { n = 0; K = k; }
    UME_DEVI bool full() const { return n == K; }
    UME_DEVI float worst() const { return bd[K - 1]; }
    UME_DEVI int size() const { return n; }
    UME_DEVI void consider(float d, int j) { topk_insert<KMAX>(bd, bj, n, K, d, j); }
    template <typename F>
    UME_DEVI void for_each(F f) const {
        for (int k = 0; k < n; ++k) f(bd[k], bj[k]);
    }
};

// K best in REGISTERS (K is a compile-time constant): sorted ascending, empty slots hold +inf.
// Insertion is one fully unrolled carry pass — no local memory, no data-dependent loop.
template <int K_>
struct RegTopK {
    float bd[K_];
    int bj[K_];
    UME_DEVI void reset(int) {
#pragma unroll
        for (int k = 0; k < K_; ++k) { bd[k] = INFINITY; bj[k] = 0x7fffffff; }
    }
    UME_DEVI bool full() const { return bj[K_ - 1] != 0x7fffffff; }
    UME_DEVI float worst() const { return bd[K_ - 1]; }
    UME_DEVI void consider(float d, int j) {
        if (!(d < bd[K_ - 1] || (d == bd[K_ - 1] && j < bj[K_ - 1]))) return;
#pragma unroll
        for (int k = 0; k < K_; ++k) {
            const bool lt = d < bd[k] || (d == bd[k] && j < bj[k]);
            const float td = lt ? bd[k] : d;
            const int tj = lt ? bj[k] : j;
            bd[k] = lt ? d : bd[k];
            bj[k] = lt ? j : bj[k];
            d = td;
            j = tj;
        }
    }
    template <typename F>
    UME_DEVI void for_each(F f) const {
#pragma unroll
        for (int k = 0; k < K_; ++k)
            if (bj[k] != 0x7fffffff) f(bd[k], bj[k]);          // fewer than K rows in the cloud: empty slots
    }
};

// Exact K nearest rows of one cloud for query (qx,qy,qz).  dist2 in pytorch3d's arithmetic.
template <bool kFma, typename Top>
UME_DEVI void grid_knn(const GridHeader& h, const int* __restrict__ cs, const float4* __restrict__ sorted_b,
                       float qx, float qy, float qz, Top& top) {
    const int cx = cell_coord(qx, h.ox, h.inv_s, h.nx), cy = cell_coord(qy, h.oy, h.inv_s, h.ny),
              cz = cell_coord(qz, h.oz, h.inv_s, h.nz);
    const int max_ring = max(h.nx, max(h.ny, h.nz));
    // squared distance from the query to the grid's box (0 inside): every row is at least that far
    const float ox = fmaxf(fmaxf(h.ox - qx, qx - h.hx), 0.f), oy = fmaxf(fmaxf(h.oy - qy, qy - h.hy), 0.f),
                oz = fmaxf(fmaxf(h.oz - qz, qz - h.hz), 0.f);
    const float out2 = (ox * ox + oy * oy + oz * oz) * 0.9999f;
    for (int ring = 0; ring <= max_ring; ++ring) {
        if (ring >= 1 && top.full()) {
            // A row not yet visited lies in a cell at Chebyshev distance >= ring from the (clamped)
            // centre cell: along that axis it is >= (ring-1)*s + (query's overshoot on that axis) away,
            // along the others >= the overshoot, hence dist^2 >= ((ring-1)*s)^2 + |overshoot|^2.
            // Far-away queries (bad hypotheses) therefore stop after a few rings too.
            const float lb = (float)(ring - 1) * h.s * 0.9999f;
            if (top.worst() < lb * lb + out2) break;
        }
        const int z0 = max(cz - ring, 0), z1 = min(cz + ring, h.nz - 1);
        const int y0 = max(cy - ring, 0), y1 = min(cy + ring, h.ny - 1);
        for (int iz = z0; iz <= z1; ++iz) {
            const bool zshell = (iz - cz == ring) || (cz - iz == ring);
            // distance from the query to the slab of cell layer iz (slightly under-estimated)
            const float zlo = h.oz + (float)iz * h.s, zhi = zlo + h.s;
            const float dz = fmaxf(fmaxf(zlo - qz, qz - zhi), 0.f) * 0.9999f;
            if (top.full() && dz * dz >= top.worst()) continue;
            for (int iy = y0; iy <= y1; ++iy) {
                const bool shell = zshell || (iy - cy == ring) || (cy - iy == ring);
                const int base = (iz * h.ny + iy) * h.nx;
                // shell rows: the whole x range; inner rows: only the two end cells
                int xa[2], xb[2], nr = 0;
                if (shell) {
                    xa[0] = max(cx - ring, 0); xb[0] = min(cx + ring, h.nx - 1); nr = 1;
                } else {
                    if (cx - ring >= 0) { xa[nr] = xb[nr] = cx - ring; ++nr; }
                    if (cx + ring <= h.nx - 1) { xa[nr] = xb[nr] = cx + ring; ++nr; }
                }
                if (top.full()) {
                    // prune by the distance from the query to the cell row's box: nothing farther than the
                    // current K-th best can enter the list; on full rows also trim the x range
                    const float ylo = h.oy + (float)iy * h.s, yhi = ylo + h.s;
                    const float dy = fmaxf(fmaxf(ylo - qy, qy - yhi), 0.f) * 0.9999f;
                    const float rem = top.worst() - dz * dz - dy * dy;
                    if (rem <= 0.f) continue;
                    const float half = sqrtf(rem) * 1.0001f + h.s * 1e-3f;
                    const int tx0 = cell_coord(qx - half, h.ox, h.inv_s, h.nx), tx1 = cell_coord(qx + half, h.ox, h.inv_s, h.nx);
                    for (int r = 0; r < nr; ++r) { xa[r] = max(xa[r], tx0); xb[r] = min(xb[r], tx1); }
                }
                for (int r = 0; r < nr; ++r) {
                    if (xb[r] < xa[r]) continue;
                    const int s = __ldg(&cs[base + xa[r]]), e = __ldg(&cs[base + xb[r] + 1]);
                    for (int t = s; t < e; ++t) {
                        const float4 c = __ldg(&sorted_b[t]);
                        const float d = dist2_ordered<kFma>(__fsub_rn(qx, c.x), __fsub_rn(qy, c.y), __fsub_rn(qz, c.z));
                        top.consider(d, __float_as_int(c.w));
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------- knn_points, any K <= 64
template <int KMAX, bool kFma>
__global__ void __launch_bounds__(128) knn_kernel(GridView grid, const float* __restrict__ q, int P1, int K,
                                                  int64_t* __restrict__ idx, float* __restrict__ d2) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P1) return;
    const GridHeader h = grid.hdr[b];
    const int* cs = grid.cell_start + (size_t)b * (grid.cells_cap + 1);
    const float4* sorted_b = grid.sorted + (size_t)b * grid.N;
    const size_t qo = (size_t)b * P1 + i;
    ArrayTopK<KMAX> top;
    top.reset(K);
    grid_knn<kFma>(h, cs, sorted_b, q[qo * 3 + 0], q[qo * 3 + 1], q[qo * 3 + 2], top);
    for (int k = 0; k < K; ++k) {
        if (idx) idx[qo * K + k] = (k < top.n) ? top.bj[k] : 0;
        if (d2) d2[qo * K + k] = (k < top.n) ? top.bd[k] : 0.f;
    }
}

// ---------------------------------------------------------------- feature_spatial_var
// out[i] = mean_{k=1..K-1} |f_i - f_nn(i,k)|_2 over the K nearest rows of the cloud itself, the
// nearest (normally the point itself) dropped (utils/loc_utils.py:579-585).
template <int KMAX, bool kFma>
__global__ void __launch_bounds__(128) spatial_var_kernel(GridView grid, const float* __restrict__ feat, int C, int K,
                                                          float* __restrict__ out) {
    const int b = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int N = grid.N;
    const GridHeader h = grid.hdr[b];
    if (t >= h.n_sorted) return;
    const int* cs = grid.cell_start + (size_t)b * (grid.cells_cap + 1);
    const float4* sorted_b = grid.sorted + (size_t)b * N;
    const float4 me = sorted_b[t];                           // cell order: neighbouring threads, neighbouring points
    const int i = __float_as_int(me.w);
    ArrayTopK<KMAX> top;
    top.reset(K);
    grid_knn<kFma>(h, cs, sorted_b, me.x, me.y, me.z, top);
    const float* fb = feat + (size_t)b * N * C;
    const float* fi = fb + (size_t)i * C;
    float acc = 0.f;
    for (int k = 1; k < top.n; ++k) {
        const float* fj = fb + (size_t)top.bj[k] * C;
        float s = 0.f;
        for (int c = 0; c < C; c += 4) {
            const float4 a = ldg_f4(fi + c), v = ldg_f4(fj + c);
            const float dx = a.x - v.x, dy = a.y - v.y, dz = a.z - v.z, dw = a.w - v.w;
            s = fmaf(dx, dx, s); s = fmaf(dy, dy, s); s = fmaf(dz, dz, s); s = fmaf(dw, dw, s);
        }
        acc += sqrtf(s);
    }
    out[(size_t)b * N + i] = (K > 1) ? acc / (float)(K - 1) : 0.f;
}

// ---------------------------------------------------------------- weighted features
// wf[i,:] = (f[i,:] - m[:]) * w[i]   (utils/loc_utils.py:649-650)
__global__ void weight_features_kernel(const float* __restrict__ f, const float* __restrict__ m, const float* __restrict__ w,
                                       int64_t rows, int C, float* __restrict__ out) {
    const int c4n = C >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows * c4n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / c4n;
        const int c = (int)(i % c4n) * 4;
        const float4 a = ldg_f4(f + r * C + c), mm = ldg_f4(m + c);
        const float ww = __ldg(w + r);
        *reinterpret_cast<float4*>(out + r * C + c) =
            make_float4((a.x - mm.x) * ww, (a.y - mm.y) * ww, (a.z - mm.z) * ww, (a.w - mm.w) * ww);
    }
}

// ---------------------------------------------------------------- correlation score
struct CorrParams {
    GridView src_grid;       // B = 1: the source cloud in cell order (spatial coherence of the CTA)
    GridView tgt_grid;
    const float* wf_src;     // (Ns, C) weighted source features, original row order
    const float* wf_tgt;     // (Nt, C)
    const float* T;          // (n_hyp, 4, 4)
    float* partial;          // (n_hyp, gridDim.x)
    int n_hyp, K;
    float inv_sigma;
};

#ifndef UME_CORR_MINB
#define UME_CORR_MINB 6            // CTAs per SM the register allocation is capped for
#endif

template <int C4, typename Top, bool kFma>
__global__ void __launch_bounds__(kCorrThreads, UME_CORR_MINB) corr_score_kernel(CorrParams p) {
    __shared__ float s_red[kCorrThreads / 32];
    // the thread's source feature row lives in shared memory ([chunk][thread]: conflict-free 16-byte
    // reads), not in 4*C4 registers: the K-best list already takes 2K of them
    __shared__ float4 s_sf[C4][kCorrThreads];
    const GridHeader hs = p.src_grid.hdr[0];
    const GridHeader ht = p.tgt_grid.hdr[0];
    const int* cs = p.tgt_grid.cell_start;
    const float4* tgt_sorted = p.tgt_grid.sorted;
    const int t = blockIdx.x * kCorrThreads + threadIdx.x;
    const bool active = t < hs.n_sorted;
    float4 me = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) {
        me = p.src_grid.sorted[t];
        const float* row = p.wf_src + (size_t)__float_as_int(me.w) * (C4 * 4);
#pragma unroll
        for (int c = 0; c < C4; ++c) s_sf[c][threadIdx.x] = ldg_f4(row + 4 * c);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int hyp = blockIdx.y; hyp < p.n_hyp; hyp += gridDim.y) {
        const float* T = p.T + (size_t)hyp * 16;
        float acc = 0.f;
        if (active) {
            // source_points @ R^T + t  (utils/loc_utils.py:626), row-times-matrix in fp32
            const float qx = fmaf(me.z, __ldg(T + 2), fmaf(me.y, __ldg(T + 1), me.x * __ldg(T + 0))) + __ldg(T + 3);
            const float qy = fmaf(me.z, __ldg(T + 6), fmaf(me.y, __ldg(T + 5), me.x * __ldg(T + 4))) + __ldg(T + 7);
            const float qz = fmaf(me.z, __ldg(T + 10), fmaf(me.y, __ldg(T + 9), me.x * __ldg(T + 8))) + __ldg(T + 11);
            Top top;
            top.reset(p.K);
            grid_knn<kFma>(ht, cs, tgt_sorted, qx, qy, qz, top);
            top.for_each([&](float dk, int jk) {
                const float* row = p.wf_tgt + (size_t)jk * (C4 * 4);
                float v = 0.f;
#pragma unroll
                for (int c = 0; c < C4; ++c) {
                    const float4 g = ldg_f4(row + 4 * c);
                    const float4 a = s_sf[c][threadIdx.x];
                    v = fmaf(a.x, g.x, v); v = fmaf(a.y, g.y, v); v = fmaf(a.z, g.z, v); v = fmaf(a.w, g.w, v);
                }
                const float e = sqrtf(dk) * p.inv_sigma;             // |p - q| / sigma
                acc = fmaf(v, 1.f / fmaf(e, e, 1.f), acc);            // cauchy_kernel (:588-589)
            });
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(UME_FULL_MASK, acc, o);
        if (lane == 0) s_red[warp] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            float s = 0.f;
            for (int w = 0; w < kCorrThreads / 32; ++w) s += s_red[w];
            p.partial[(size_t)hyp * gridDim.x + blockIdx.x] = s;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------- correlation score, tile version
// (design notes at the top of corr_tile.cuh)
struct TileParams {
    GridView src_grid;       // B = 1: the source cloud in cell order
    GridView tgt_grid;
    const float* wf_src;     // (Ns, C)
    const float* wf_tgt;     // (Nt, C)
    const float* T;          // (n_hyp, 4, 4)
    float* partial;          // (n_hyp, nslots)
    int* tiles;              // [0] = number of tiles, then (start, len) pairs
    unsigned long long* stats;   // optional: [0] lanes served by the tile path, [1] by the ring-search fallback
    int n_hyp, K, nslots;
    float inv_sigma;
};

// Tiles = runs of <= 32 consecutive entries of the source's cell-sorted array that do not cross a cell
// row (a tile that wrapped from the end of one row to the start of the next would span the cloud).
__global__ void __launch_bounds__(1024) corr_tiles_kernel(GridView g, int* __restrict__ tiles, int max_tiles) {
    __shared__ int warp_tot[32];
    __shared__ int running;
    const GridHeader h = g.hdr[0];
    const int nrows = h.ny * h.nz;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (t == 0) running = 0;
    __syncthreads();
    for (int base = 0; base < nrows; base += 1024) {
        const int r = base + t;
        int s = 0, n = 0;
        if (r < nrows) {
            s = g.cell_start[r * h.nx];
            n = g.cell_start[(r + 1) * h.nx] - s;
        }
        const int nt = (n + 31) >> 5;
        int incl = nt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(UME_FULL_MASK, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) warp_tot[w] = incl;
        __syncthreads();
        int wbase = 0;
        for (int k = 0; k < w; ++k) wbase += warp_tot[k];
        const int run0 = running;
        int at = run0 + wbase + incl - nt;
        for (int k = 0; k < nt && at < max_tiles; ++k, ++at) {
            tiles[1 + 2 * at] = s + 32 * k;
            tiles[2 + 2 * at] = min(32, n - 32 * k);
        }
        __syncthreads();
        if (t == 1023) running = run0 + wbase + incl;
        __syncthreads();
    }
    if (t == 0) tiles[0] = min(running, max_tiles);
}

#ifndef UME_TILE_RHO
#define UME_TILE_RHO 1.3f           // staged radius = this x the K-NN radius estimated from the tile's own density
#endif

// One query (ax, ay, az) against the M staged candidates, lane = candidate (S slots per lane): finds the K
// nearest among those closer than sqrt(m2) and adds lane k's k-th neighbour contribution to `acc`.  Returns
// false (warp-uniform) when the K nearest are not all inside the margin.  `qi` = the query's row of sm.sf.
template <int S, int C4, bool kFma>
UME_DEVI bool tile_one_query(tile::WarpSmem<C4>& sm, int M, int K, float ax, float ay, float az, float m2, float tau0, int qi,
                             const float* __restrict__ wf_tgt, float inv_sigma, float& acc) {
    const int lane = threadIdx.x & 31;
    const unsigned lt = lanemask_lt();
    float d[S];
#pragma unroll
    for (int j = 0; j < S; ++j) {
        const int c = j * 32 + lane;
        const float4 cd = sm.cand[c < M ? c : 0];
        const float v = dist2_ordered<kFma>(__fsub_rn(ax, cd.x), __fsub_rn(ay, cd.y), __fsub_rn(az, cd.z));
        d[j] = c < M ? v : INFINITY;
    }
    auto count_lt = [&](float tau) {
        int n = 0;
#pragma unroll
        for (int j = 0; j < S; ++j) n += (d[j] < tau) ? 1 : 0;
        return __reduce_add_sync(UME_FULL_MASK, n);
    };
    // a threshold with exactly K candidates below it, inside [0, margin^2]
    float lo = 0.f, hi = m2, tau = m2;
    int n_hi = count_lt(m2);
    if (n_hi < K) return false;                                // the K nearest are not all inside the margin
    int n_tau = n_hi, n_lo = 0;
    if (n_hi > K) {
        float probe = fminf(tau0, 0.5f * m2);
        bool done = false;
        for (int it = 0; it < 64 && !done; ++it) {
            const int n = count_lt(probe);
            if (n == K) { tau = probe; n_tau = n; done = true; break; }
            if (n < K) { lo = probe; n_lo = n; } else { hi = probe; n_hi = n; }
            const float mid = 0.5f * (lo + hi);
            if (!(mid > lo && mid < hi)) { tau = hi; n_tau = n_hi; done = true; break; }   // no float strictly inside: ties straddle K
            probe = mid;
        }
        if (!done) return false;                               // (K or more candidates at distance ~0: leave it to the ring search)
    }
    // selected: everything below tau; when ties straddle K (n_tau > K) the candidates in [lo, tau) are all
    // EQUAL in distance and the (K - n_lo) lowest rows among them win
    unsigned pick = 0;                                         // bit j: this lane's slot j is a neighbour
    if (n_tau == K) {
#pragma unroll
        for (int j = 0; j < S; ++j) pick |= (d[j] < tau) ? (1u << j) : 0u;
    } else {
        int rows[S];
        unsigned tied = 0;
#pragma unroll
        for (int j = 0; j < S; ++j) {
            const int c = j * 32 + lane;
            rows[j] = __float_as_int(sm.cand[c < M ? c : 0].w);
            pick |= (d[j] < lo) ? (1u << j) : 0u;
            tied |= (d[j] >= lo && d[j] < tau) ? (1u << j) : 0u;
        }
        for (int need = K - n_lo; need > 0; --need) {          // rare: one warp-wide arg-min per tied winner
            int best = 0x7fffffff;
#pragma unroll
            for (int j = 0; j < S; ++j) best = ((tied >> j) & 1u) ? min(best, rows[j]) : best;
            best = __reduce_min_sync(UME_FULL_MASK, best);
#pragma unroll
            for (int j = 0; j < S; ++j)
                if (((tied >> j) & 1u) && rows[j] == best) { tied &= ~(1u << j); pick |= 1u << j; }
        }
    }
    // squeeze the K neighbours into the list
    int base = 0;
    __syncwarp();
#pragma unroll
    for (int j = 0; j < S; ++j) {
        const bool take = (pick >> j) & 1u;
        const unsigned m = __ballot_sync(UME_FULL_MASK, take);
        if (take) {
            const int at = base + __popc(m & lt);
            sm.sel_d2[at] = d[j];
            sm.sel_row[at] = __float_as_int(sm.cand[j * 32 + lane].w);
        }
        base += __popc(m);
    }
    __syncwarp();
    // K dot products in parallel: lane k takes the k-th neighbour
    if (lane < K) {
        const float dk = sm.sel_d2[lane];
        const float* row = wf_tgt + (size_t)sm.sel_row[lane] * (C4 * 4);
        float v = 0.f;
#pragma unroll
        for (int c = 0; c < C4; ++c) {
            const float4 g = ldg_f4(row + 4 * c);
            const float4 a = sm.sf[qi][c];                     // broadcast
            v = fmaf(a.x, g.x, v); v = fmaf(a.y, g.y, v); v = fmaf(a.z, g.z, v); v = fmaf(a.w, g.w, v);
        }
        const float e = sqrtf(dk) * inv_sigma;                 // |p - q| / sigma
        acc = fmaf(v, 1.f / fmaf(e, e, 1.f), acc);              // cauchy_kernel (:588-589)
    }
    return true;
}

template <int C4, bool kFma>
UME_DEVI bool tile_one_query_any(tile::WarpSmem<C4>& sm, int M, int K, float ax, float ay, float az, float m2, float tau0,
                                 int qi, const float* __restrict__ wf_tgt, float inv_sigma, float& acc) {
    if (M <= 32 * 4) return tile_one_query<4, C4, kFma>(sm, M, K, ax, ay, az, m2, tau0, qi, wf_tgt, inv_sigma, acc);
    if (M <= 32 * 7) return tile_one_query<7, C4, kFma>(sm, M, K, ax, ay, az, m2, tau0, qi, wf_tgt, inv_sigma, acc);
    return tile_one_query<tile::kSlotsMax, C4, kFma>(sm, M, K, ax, ay, az, m2, tau0, qi, wf_tgt, inv_sigma, acc);
}

// squared distance from q to the nearest face of the cell block [c0, c1] that has target points beyond it
// (faces on the grid's boundary do not count: every target point lies inside the grid), shrunk by a safety
// factor, and capped by the distance to the block's farthest corner (a finite range for the bisection)
UME_DEVI float block_margin2(const GridHeader& ht, float qx, float qy, float qz, int cx0, int cx1, int cy0, int cy1, int cz0, int cz1) {
    const float inf = INFINITY;
    const float xl = ht.ox + (float)cx0 * ht.s, xh = ht.ox + (float)(cx1 + 1) * ht.s;
    const float yl = ht.oy + (float)cy0 * ht.s, yh = ht.oy + (float)(cy1 + 1) * ht.s;
    const float zl = ht.oz + (float)cz0 * ht.s, zh = ht.oz + (float)(cz1 + 1) * ht.s;
    float mg = fminf(fminf(fminf(cx0 == 0 ? inf : qx - xl, cx1 == ht.nx - 1 ? inf : xh - qx),
                           fminf(cy0 == 0 ? inf : qy - yl, cy1 == ht.ny - 1 ? inf : yh - qy)),
                     fminf(cz0 == 0 ? inf : qz - zl, cz1 == ht.nz - 1 ? inf : zh - qz));
    const float fx = fmaxf(fabsf(qx - xl), fabsf(qx - xh)), fy = fmaxf(fabsf(qy - yl), fabsf(qy - yh)),
                fz = fmaxf(fabsf(qz - zl), fabsf(qz - zh));
    const float far = sqrtf(fx * fx + fy * fy + fz * fz) * 1.001f + 1e-3f * ht.s;
    mg = fminf(fmaxf(mg * 0.9999f - 1e-4f * ht.s, 0.f), far);
    return mg * mg;
}

template <int C4, typename Top, bool kFma>
__global__ void __launch_bounds__(32 * tile::kWarps, 4) corr_tile_kernel(TileParams p) {
    extern __shared__ __align__(16) unsigned char tile_smem_raw[];
    using WS = tile::WarpSmem<C4>;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WS& sm = reinterpret_cast<WS*>(tile_smem_raw)[warp];
    const GridHeader ht = p.tgt_grid.hdr[0];
    const int* cs = p.tgt_grid.cell_start;
    const float4* tgt_sorted = p.tgt_grid.sorted;
    const float4* src_sorted = p.src_grid.sorted;
    const int n_tiles = p.tiles[0];
    const int slot = blockIdx.x * tile::kWarps + warp;
    const int K = p.K;
    unsigned long long n_fast = 0, n_slow = 0;

    for (int hyp = blockIdx.y; hyp < p.n_hyp; hyp += gridDim.y) {
        const float* T = p.T + (size_t)hyp * 16;
        const float t00 = __ldg(T + 0), t01 = __ldg(T + 1), t02 = __ldg(T + 2), t03 = __ldg(T + 3);
        const float t10 = __ldg(T + 4), t11 = __ldg(T + 5), t12 = __ldg(T + 6), t13 = __ldg(T + 7);
        const float t20 = __ldg(T + 8), t21 = __ldg(T + 9), t22 = __ldg(T + 10), t23 = __ldg(T + 11);
        float score = 0.f;                                     // this slot's share of the hypothesis, fixed tile order
        for (int tl = slot; tl < n_tiles; tl += p.nslots) {
            const int t0 = p.tiles[1 + 2 * tl], tn = p.tiles[2 + 2 * tl];
            const bool active = lane < tn;
            float4 me = make_float4(0.f, 0.f, 0.f, 0.f);
            __syncwarp();                                      // the previous tile's readers are done with sm
            if (active) {
                me = __ldg(&src_sorted[t0 + lane]);
                const float* row = p.wf_src + (size_t)__float_as_int(me.w) * (C4 * 4);
#pragma unroll
                for (int c = 0; c < C4; ++c) sm.sf[lane][c] = ldg_f4(row + 4 * c);
            }
            // source_points @ R^T + t  (utils/loc_utils.py:626), row-times-matrix in fp32
            const float qx = fmaf(me.z, t02, fmaf(me.y, t01, me.x * t00)) + t03;
            const float qy = fmaf(me.z, t12, fmaf(me.y, t11, me.x * t10)) + t13;
            const float qz = fmaf(me.z, t22, fmaf(me.y, t21, me.x * t20)) + t23;
            // the tile's landing zone and the radius that should hold K target points around any of its queries
            const float bx0 = tile::warp_min(active ? qx : INFINITY), bx1 = tile::warp_max(active ? qx : -INFINITY);
            const float by0 = tile::warp_min(active ? qy : INFINITY), by1 = tile::warp_max(active ? qy : -INFINITY);
            const float bz0 = tile::warp_min(active ? qz : INFINITY), bz1 = tile::warp_max(active ? qz : -INFINITY);
            const float dx = bx1 - bx0, dy = by1 - by0, dz = bz1 - bz0;
            const float area = fmaxf(dx * dy, fmaxf(dx * dz, dy * dz));
            const float rk = sqrtf((float)K * area / (3.14159265f * (float)tn));     // K-NN radius at the tile's own density
            float rho = fminf(fmaxf(UME_TILE_RHO * rk, 0.5f * ht.s), 6.f * ht.s);
            const bool finite = isfinite(dx) && isfinite(dy) && isfinite(dz);
            int M = -1;
            int cx0 = 0, cx1 = 0, cy0 = 0, cy1 = 0, cz0 = 0, cz1 = 0;
            if (finite) {
                // too many candidates: half the radius; too few: twice the radius (at most three stagings)
                bool shrunk = false, grown = false;
                for (int attempt = 0; attempt < 3; ++attempt) {
                    cx0 = cell_coord(bx0 - rho, ht.ox, ht.inv_s, ht.nx); cx1 = cell_coord(bx1 + rho, ht.ox, ht.inv_s, ht.nx);
                    cy0 = cell_coord(by0 - rho, ht.oy, ht.inv_s, ht.ny); cy1 = cell_coord(by1 + rho, ht.oy, ht.inv_s, ht.ny);
                    cz0 = cell_coord(bz0 - rho, ht.oz, ht.inv_s, ht.nz); cz1 = cell_coord(bz1 + rho, ht.oz, ht.inv_s, ht.nz);
                    M = tile::stage_block<C4>(sm, ht, cs, tgt_sorted, cx0, cx1, cy0, cy1, cz0, cz1);
                    if (M < 0 && !grown) { rho *= 0.5f; shrunk = true; }
                    else if (M >= 0 && M < 2 * K && !shrunk) { rho *= 2.f; grown = true; }
                    else break;
                }
            }
            unsigned failed = 0xffffffffu >> (32 - tn);        // queries still to be served by the ring search
            float acc = 0.f;
            if (M >= K) {                                      // warp-uniform
                const float margin2 = block_margin2(ht, qx, qy, qz, cx0, cx1, cy0, cy1, cz0, cz1);
                const float tau0 = rk * rk;
                unsigned again = 0u;                           // queries whose K nearest reach past their margin
                for (int i = 0; i < tn; ++i) {
                    const float ax = __shfl_sync(UME_FULL_MASK, qx, i), ay = __shfl_sync(UME_FULL_MASK, qy, i),
                                az = __shfl_sync(UME_FULL_MASK, qz, i), m2 = __shfl_sync(UME_FULL_MASK, margin2, i);
                    if (!tile_one_query_any<C4, kFma>(sm, M, K, ax, ay, az, m2, tau0, i, p.wf_tgt, p.inv_sigma, acc)) again |= 1u << i;
                }
                // second chance, query by query: a block of its own around the query, twice the radius
                failed = 0u;
                const float rho2 = fminf(2.2f * fmaxf(rho, rk), 8.f * ht.s);
                while (again) {
                    const int i = __ffs(again) - 1;
                    again &= again - 1;
                    const float ax = __shfl_sync(UME_FULL_MASK, qx, i), ay = __shfl_sync(UME_FULL_MASK, qy, i),
                                az = __shfl_sync(UME_FULL_MASK, qz, i);
                    const int ax0 = cell_coord(ax - rho2, ht.ox, ht.inv_s, ht.nx), ax1 = cell_coord(ax + rho2, ht.ox, ht.inv_s, ht.nx);
                    const int ay0 = cell_coord(ay - rho2, ht.oy, ht.inv_s, ht.ny), ay1 = cell_coord(ay + rho2, ht.oy, ht.inv_s, ht.ny);
                    const int az0 = cell_coord(az - rho2, ht.oz, ht.inv_s, ht.nz), az1 = cell_coord(az + rho2, ht.oz, ht.inv_s, ht.nz);
                    const int M2 = tile::stage_block<C4>(sm, ht, cs, tgt_sorted, ax0, ax1, ay0, ay1, az0, az1);
                    bool ok = false;
                    if (M2 >= K) {
                        const float m2 = block_margin2(ht, ax, ay, az, ax0, ax1, ay0, ay1, az0, az1);
                        ok = tile_one_query_any<C4, kFma>(sm, M2, K, ax, ay, az, m2, 4.f * tau0, i, p.wf_tgt, p.inv_sigma, acc);
                    }
                    if (!ok) failed |= 1u << i;
                }
            }
            if (active && ((failed >> lane) & 1u)) {
                // sparse or far-away landing zone, more candidates than fit: the exact ring search of round 1
                Top top;
                top.reset(K);
                grid_knn<kFma>(ht, cs, tgt_sorted, qx, qy, qz, top);
                top.for_each([&](float dk, int jk) {
                    const float* row = p.wf_tgt + (size_t)jk * (C4 * 4);
                    float v = 0.f;
#pragma unroll
                    for (int c = 0; c < C4; ++c) {
                        const float4 g = ldg_f4(row + 4 * c);
                        const float4 a = sm.sf[lane][c];
                        v = fmaf(a.x, g.x, v); v = fmaf(a.y, g.y, v); v = fmaf(a.z, g.z, v); v = fmaf(a.w, g.w, v);
                    }
                    const float e = sqrtf(dk) * p.inv_sigma;
                    acc = fmaf(v, 1.f / fmaf(e, e, 1.f), acc);
                });
            }
            if (p.stats) {
                n_fast += tn - __popc(failed);
                n_slow += __popc(failed);
                if (lane == 0) { atomicAdd(&p.stats[2], (unsigned long long)max(M, 0)); atomicAdd(&p.stats[3], 1ull); if (M < 0) atomicAdd(&p.stats[4], 1ull); if (M >= 0 && M < K) atomicAdd(&p.stats[5], 1ull); }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(UME_FULL_MASK, acc, o);
            score += acc;
        }
        if (lane == 0) p.partial[(size_t)hyp * p.nslots + slot] = score;
    }
    if (p.stats && lane == 0) {
        atomicAdd(&p.stats[0], n_fast);
        atomicAdd(&p.stats[1], n_slow);
    }
}

// score[h] = sum_b partial[h][b] / Ns, then the arg-max (first index on ties).  One CTA.
__global__ void __launch_bounds__(256) corr_finalize_kernel(const float* __restrict__ partial, int n_hyp, int nb, float inv_ns,
                                                            float* __restrict__ score, int64_t* __restrict__ best) {
    __shared__ float s_v[256];
    __shared__ int s_i[256];
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int h = threadIdx.x; h < n_hyp; h += 256) {
        float s = 0.f;
        for (int b = 0; b < nb; ++b) s += partial[(size_t)h * nb + b];     // fixed order: deterministic
        s *= inv_ns;
        score[h] = s;
        if (s > bv) { bv = s; bi = h; }
    }
    s_v[threadIdx.x] = bv;
    s_i[threadIdx.x] = bi;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const float ov = s_v[threadIdx.x + o];
            const int oi = s_i[threadIdx.x + o];
            if (ov > s_v[threadIdx.x] || (ov == s_v[threadIdx.x] && oi < s_i[threadIdx.x])) { s_v[threadIdx.x] = ov; s_i[threadIdx.x] = oi; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0 && best) *best = (s_i[0] == 0x7fffffff) ? 0 : s_i[0];
}

// K = 20 (the reference's corr_num_nn) keeps the K best in registers; other K use the array version.
template <int C4, bool kFma>
void launch_corr_c(const CorrParams& p, int K, dim3 grid, cudaStream_t stream) {
    if (K == 20) corr_score_kernel<C4, RegTopK<20>, kFma><<<grid, kCorrThreads, 0, stream>>>(p);
    else corr_score_kernel<C4, ArrayTopK<32>, kFma><<<grid, kCorrThreads, 0, stream>>>(p);
}
void launch_corr(const CorrParams& p, int C, int K, bool fma, dim3 grid, cudaStream_t stream) {
    if (C == 32) {
        if (fma) launch_corr_c<8, true>(p, K, grid, stream);
        else launch_corr_c<8, false>(p, K, grid, stream);
    } else {
        if (fma) launch_corr_c<16, true>(p, K, grid, stream);
        else launch_corr_c<16, false>(p, K, grid, stream);
    }
}

template <bool kFma>
int launch_knn(const GridView& g, const float* q, int B, int P1, int K, int64_t* idx, float* d2, cudaStream_t stream) {
    dim3 grid((unsigned)((P1 + 127) / 128), (unsigned)B);
    if (K <= 16) knn_kernel<16, kFma><<<grid, 128, 0, stream>>>(g, q, P1, K, idx, d2);
    else if (K <= 32) knn_kernel<32, kFma><<<grid, 128, 0, stream>>>(g, q, P1, K, idx, d2);
    else knn_kernel<64, kFma><<<grid, 128, 0, stream>>>(g, q, P1, K, idx, d2);
    count_launch();
    return check_launch("knn_kernel");
}

}  // namespace
}  // namespace ume

// ================================================================== C ABI
extern "C" size_t ume_knn_workspace_bytes(int B, int P1, int P2) {
    (void)P1;
    if (B <= 0 || P2 <= 0) return 0;
    return ume::grid_workspace_bytes(B, P2, ume::kCellsCap) + 256;
}

extern "C" int ume_knn_f32(const float* q, const float* pcl, int B, int P1, int P2, int K, unsigned flags, int64_t* idx,
                           float* d2, void* ws, size_t ws_bytes, void* stream_) {
    using namespace ume;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    UME_REQUIRE(B >= 0 && P1 >= 0 && P2 >= 0, UME_ERR_BAD_ARG, "ume_knn_f32: negative size");
    if (B == 0 || P1 == 0) return UME_OK;
    UME_REQUIRE(q && pcl, UME_ERR_BAD_ARG, "ume_knn_f32: null pointer");
    UME_REQUIRE(K >= 1 && K <= 64, UME_ERR_UNSUPPORTED, "ume_knn_f32: K = %d not in [1,64]", K);
    UME_REQUIRE(K <= P2, UME_ERR_BAD_ARG, "ume_knn_f32: K = %d > P2 = %d", K, P2);
    UME_REQUIRE(P2 <= kMaxPoints && B <= 65535, UME_ERR_UNSUPPORTED, "ume_knn_f32: size not supported");
    UME_REQUIRE(ws && ws_bytes >= ume_knn_workspace_bytes(B, P1, P2), UME_ERR_WORKSPACE, "ume_knn_f32: workspace too small");
    Workspace w(ws, ws_bytes);
    GridView g;
    int rc = grid_build(pcl, pcl, B, P2, P2, 0.f, -(float)max(2, K / 8), kCellsCap, w, &g, stream);
    if (rc != UME_OK) return rc;
    ProfScope prof(UME_PROF_KNN, stream);
    return (flags & UME_FLAG_FMA_DIST) ? launch_knn<true>(g, q, B, P1, K, idx, d2, stream)
                                       : launch_knn<false>(g, q, B, P1, K, idx, d2, stream);
}

extern "C" size_t ume_feature_spatial_var_workspace_bytes(int B, int N) {
    if (B <= 0 || N <= 0) return 0;
    return ume::grid_workspace_bytes(B, N, ume::kCellsCap) + 256;
}

extern "C" int ume_feature_spatial_var_f32(const float* pts, const float* feat, int B, int N, int C, int knn,
                                           unsigned flags, float* out, void* ws, size_t ws_bytes, void* stream_) {
    using namespace ume;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    UME_REQUIRE(B >= 0 && N >= 0, UME_ERR_BAD_ARG, "ume_feature_spatial_var_f32: negative size");
    if (B == 0 || N == 0) return UME_OK;
    UME_REQUIRE(pts && feat && out, UME_ERR_BAD_ARG, "ume_feature_spatial_var_f32: null pointer");
    UME_REQUIRE(knn >= 1 && knn <= 64 && knn <= N, UME_ERR_UNSUPPORTED, "ume_feature_spatial_var_f32: knn = %d not in [1, min(64, N)]", knn);
    UME_REQUIRE(C >= 4 && C % 4 == 0, UME_ERR_UNSUPPORTED, "ume_feature_spatial_var_f32: C = %d must be a multiple of 4", C);
    UME_REQUIRE(N <= kMaxPoints && B <= 65535, UME_ERR_UNSUPPORTED, "ume_feature_spatial_var_f32: size not supported");
    UME_REQUIRE(ws && ws_bytes >= ume_feature_spatial_var_workspace_bytes(B, N), UME_ERR_WORKSPACE,
                "ume_feature_spatial_var_f32: workspace too small");
    Workspace w(ws, ws_bytes);
    GridView g;
    int rc = grid_build(pts, pts, B, N, N, 0.f, -(float)max(2, knn / 8), kCellsCap, w, &g, stream);
    if (rc != UME_OK) return rc;
    ProfScope prof(UME_PROF_KNN, stream);
    cudaMemsetAsync(out, 0, (size_t)B * N * sizeof(float), stream);      // rows with non-finite coordinates
    dim3 grid((unsigned)((N + 127) / 128), (unsigned)B);
    const bool fma = (flags & UME_FLAG_FMA_DIST) != 0;
    if (fma) spatial_var_kernel<64, true><<<grid, 128, 0, stream>>>(g, feat, C, knn, out);
    else spatial_var_kernel<64, false><<<grid, 128, 0, stream>>>(g, feat, C, knn, out);
    count_launch();
    return check_launch("spatial_var_kernel");
}

extern "C" int ume_weight_features_f32(const float* f, const float* mean, const float* w, int64_t rows, int C, float* out,
                                       void* stream_) {
    using namespace ume;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (rows <= 0) return UME_OK;
    UME_REQUIRE(f && mean && w && out, UME_ERR_BAD_ARG, "ume_weight_features_f32: null pointer");
    UME_REQUIRE(C >= 4 && C % 4 == 0, UME_ERR_UNSUPPORTED, "ume_weight_features_f32: C must be a multiple of 4");
    const int64_t total = rows * (C / 4);
    weight_features_kernel<<<(unsigned)std::min<int64_t>((total + 255) / 256, 148 * 8), 256, 0, stream>>>(f, mean, w, rows, C, out);
    count_launch();
    return check_launch("weight_features_kernel");
}

namespace ume {
namespace {

// diagnostics (tools/bench_corr.py): lanes served by the tile path / by the fallback, staged candidates, tiles,
// tiles whose block did not fit, tiles with fewer than K candidates
__device__ unsigned long long g_corr_stats[8];
bool g_corr_stats_on = false;

constexpr int kTileSlotsX = 37;       // CTAs along the tiles: kTileSlotsX * tile::kWarps warp slots share the tiles

template <int C4, bool kFma>
int launch_tile(const TileParams& p, int K, dim3 grid, cudaStream_t stream) {
    const size_t smem = sizeof(tile::WarpSmem<C4>) * tile::kWarps;
    auto k20 = corr_tile_kernel<C4, RegTopK<20>, kFma>;
    auto kany = corr_tile_kernel<C4, ArrayTopK<32>, kFma>;
    cudaError_t e = cudaFuncSetAttribute(K == 20 ? k20 : kany, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    UME_REQUIRE(e == cudaSuccess, UME_ERR_CUDA, "corr_tile_kernel: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    if (K == 20) k20<<<grid, 32 * tile::kWarps, smem, stream>>>(p);
    else kany<<<grid, 32 * tile::kWarps, smem, stream>>>(p);
    return UME_OK;
}

}  // namespace
}  // namespace ume

extern "C" size_t ume_corr_scores_workspace_bytes(int Ns, int Nt, int n_hyp) {
    if (Ns <= 0 || Nt <= 0 || n_hyp <= 0) return 0;
    const size_t nb = (size_t)(Ns + ume::kCorrThreads - 1) / ume::kCorrThreads;
    const size_t nslots = (size_t)ume::kTileSlotsX * ume::tile::kWarps;
    const size_t max_tiles = (size_t)Ns / 32 + ume::kCellsCap + 1;
    return ume::grid_workspace_bytes(1, Ns, ume::kCellsCap) + ume::grid_workspace_bytes(1, Nt, ume::kCellsCap) +
           ume::align_up((size_t)n_hyp * std::max(nb, nslots) * sizeof(float), 256) +
           ume::align_up((1 + 2 * max_tiles) * sizeof(int), 256) + 1024;
}

extern "C" int ume_corr_scores_f32(const float* src_pts, const float* tgt_pts, const float* wf_src, const float* wf_tgt,
                                   const float* T, int Ns, int Nt, int C, int n_hyp, int K, float sigma, unsigned flags,
                                   float* score, int64_t* best, void* ws, size_t ws_bytes, void* stream_) {
    using namespace ume;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    UME_REQUIRE(Ns >= 0 && Nt >= 0 && n_hyp >= 0, UME_ERR_BAD_ARG, "ume_corr_scores_f32: negative size");
    if (n_hyp == 0) return UME_OK;
    UME_REQUIRE(src_pts && tgt_pts && wf_src && wf_tgt && T && score, UME_ERR_BAD_ARG, "ume_corr_scores_f32: null pointer");
    UME_REQUIRE(Ns >= 1 && Nt >= 1, UME_ERR_BAD_ARG, "ume_corr_scores_f32: empty cloud");
    UME_REQUIRE(C == 32 || C == 64, UME_ERR_UNSUPPORTED, "ume_corr_scores_f32: C = %d (32 or 64 supported)", C);
    UME_REQUIRE(K >= 1 && K <= 32 && K <= Nt, UME_ERR_UNSUPPORTED, "ume_corr_scores_f32: K = %d not in [1, min(32, Nt)]", K);
    UME_REQUIRE(sigma > 0.f, UME_ERR_BAD_ARG, "ume_corr_scores_f32: sigma must be positive");
    UME_REQUIRE(Ns <= kMaxPoints && Nt <= kMaxPoints, UME_ERR_UNSUPPORTED, "ume_corr_scores_f32: cloud too large");
    UME_REQUIRE(ws && ws_bytes >= ume_corr_scores_workspace_bytes(Ns, Nt, n_hyp), UME_ERR_WORKSPACE,
                "ume_corr_scores_f32: workspace too small");
    Workspace w(ws, ws_bytes);
    const bool fma = (flags & UME_FLAG_FMA_DIST) != 0;
    if (!(flags & UME_FLAG_CORR_THREAD)) {
        // tile version (corr_tile.cuh): a warp per tile of 32 neighbouring source points, candidates staged once
        TileParams p;
        int rc = grid_build(src_pts, src_pts, 1, Ns, Ns, 0.f, -8.f, kCellsCap, w, &p.src_grid, stream);
        if (rc != UME_OK) return rc;
        rc = grid_build(tgt_pts, tgt_pts, 1, Nt, Nt, 0.f, -4.f, kCellsCap, w, &p.tgt_grid, stream);
        if (rc != UME_OK) return rc;
        const int max_tiles = Ns / 32 + kCellsCap + 1;
        p.nslots = kTileSlotsX * tile::kWarps;
        p.partial = w.take<float>((size_t)n_hyp * p.nslots);
        p.tiles = w.take<int>((size_t)1 + 2 * max_tiles);
        UME_REQUIRE(w.ok(), UME_ERR_WORKSPACE, "ume_corr_scores_f32: workspace too small");
        p.stats = nullptr;
        if (g_corr_stats_on) cudaGetSymbolAddress(reinterpret_cast<void**>(&p.stats), g_corr_stats);
        p.wf_src = wf_src; p.wf_tgt = wf_tgt; p.T = T; p.n_hyp = n_hyp; p.K = K; p.inv_sigma = 1.f / sigma;
        ProfScope prof(UME_PROF_CORR, stream);
        corr_tiles_kernel<<<1, 1024, 0, stream>>>(p.src_grid, p.tiles, max_tiles);
        // hypothesis groups: a few CTAs per SM in flight, several waves for balance
        const int gy = max(1, min(n_hyp, 64));
        dim3 grid((unsigned)kTileSlotsX, (unsigned)gy);
        if (C == 32) rc = fma ? launch_tile<8, true>(p, K, grid, stream) : launch_tile<8, false>(p, K, grid, stream);
        else rc = fma ? launch_tile<16, true>(p, K, grid, stream) : launch_tile<16, false>(p, K, grid, stream);
        if (rc != UME_OK) return rc;
        corr_finalize_kernel<<<1, 256, 0, stream>>>(p.partial, n_hyp, p.nslots, 1.f / (float)Ns, score, best);
        count_launch(3);
        return check_launch("corr_tile_kernel");
    }
    CorrParams p;
    int rc = grid_build(src_pts, src_pts, 1, Ns, Ns, 0.f, -8.f, kCellsCap, w, &p.src_grid, stream);
    if (rc != UME_OK) return rc;
    rc = grid_build(tgt_pts, tgt_pts, 1, Nt, Nt, 0.f, -(float)max(2, K / 8), kCellsCap, w, &p.tgt_grid, stream);
    if (rc != UME_OK) return rc;
    const int nb = (Ns + kCorrThreads - 1) / kCorrThreads;
    p.partial = w.take<float>((size_t)n_hyp * nb);
    UME_REQUIRE(w.ok(), UME_ERR_WORKSPACE, "ume_corr_scores_f32: workspace too small");
    p.wf_src = wf_src; p.wf_tgt = wf_tgt; p.T = T; p.n_hyp = n_hyp; p.K = K; p.inv_sigma = 1.f / sigma;
    // hypothesis groups: enough CTAs for ~8 waves of the 148 SMs
    int gy = (148 * 8 * 4 + nb - 1) / nb;
    gy = max(1, min(gy, min(n_hyp, 65535)));
    dim3 grid((unsigned)nb, (unsigned)gy);
    ProfScope prof(UME_PROF_CORR, stream);
    launch_corr(p, C, K, fma, grid, stream);
    corr_finalize_kernel<<<1, 256, 0, stream>>>(p.partial, n_hyp, nb, 1.f / (float)Ns, score, best);
    count_launch(2);
    return check_launch("corr_score_kernel");
}

// Diagnostics of the tile kernel (not part of the measured path): enable / read-and-reset the counters.
extern "C" int ume_corr_stats(int enable, uint64_t* out8_host) {
    using namespace ume;
    g_corr_stats_on = enable != 0;
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (out8_host) {
        cudaError_t e = cudaMemcpyFromSymbol(out8_host, g_corr_stats, sizeof(z));
        UME_REQUIRE(e == cudaSuccess, UME_ERR_CUDA, "ume_corr_stats: %s", cudaGetErrorString(e));
    }
    cudaError_t e = cudaMemcpyToSymbol(g_corr_stats, z, sizeof(z));
    UME_REQUIRE(e == cudaSuccess, UME_ERR_CUDA, "ume_corr_stats: %s", cudaGetErrorString(e));
    return UME_OK;
}
