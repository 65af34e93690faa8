// Hypothesis selection by feature correlation: the reference's FeatureCorrelator
// (utils/loc_utils.py:579-681, driven by evaluate.py:20-47,287-296), general K nearest neighbours
// (pytorch3d.ops.knn_points as used at utils/loc_utils.py:580,623) and feature_spatial_var (:579-585).
//
// Everything is built on one device routine: an exact K-nearest search of one query per THREAD over
// the search grid (rings of cells of growing Chebyshev radius, stop when the K-th best distance beats
// the next ring's lower bound), the K best ordered by (dist2, row) — ties go to the lower row, as a row-order
// scan with strict '<' does in pytorch3d — kept sorted in a per-thread array (knn_points, any K) or unsorted
// in a shared-memory column with the worst entry tracked (the hypothesis scorer, which only sums over them).
//
// Correlation score (pc_corr, utils/loc_utils.py:592-619):
//   score[h] = (1/Ns) sum_i sum_{k<K} cauchy(|T_h p_i - q_nn(i,k)|; sigma) <wf_src_i, wf_tgt_nn(i,k)>
// One thread per (source point, hypothesis) for the search: the source points are taken in CELL ORDER of
// their own grid, so the 128 threads of a CTA are spatial neighbours; under the same rigid transform they
// land in the same few target cells and share candidates through L1.  The K x 32 feature dot products of a
// warp's queries are then formed by the warp together, C/4 lanes per (query, neighbour): a target feature
// row is read as one 128-byte line instead of 32 lanes gathering 16 bytes from 32 different rows.
#include "ume_common.cuh"

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
    static constexpr int kCap = KMAX;
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

// K best UNSORTED in a shared-memory column (entries `stride` apart), the worst of them tracked through per-group
// maxima in registers: an insertion overwrites the worst entry, re-reads the S entries of that entry's group and
// compares G group maxima — ~30 instructions instead of the ~125 of a carry pass through a sorted register list.
// (Profile of the round-1 kernel: 72 % of all issued instructions were that carry pass, because a warp runs it
// whenever ANY of its 32 queries inserts.)  An entry is ONE 64-bit key, (dist2 bits << 32) | row: squared distances
// are non-negative, so unsigned key order is the (dist2, row) lexicographic order used everywhere — one compare,
// one shared-memory access per entry, ties by row for free.
template <int K_>
struct SmemGroupTopK {
    static constexpr int kCap = K_;
    static constexpr int S = 5, G = (K_ + S - 1) / S;
    unsigned long long* col;
    int stride, n;
    unsigned long long gkey[G];      // per group: the largest key and where it sits
    int gp[G];
    unsigned long long wkey;         // the worst of all
    int wp;
    float wd;
    UME_DEVI SmemGroupTopK(unsigned long long* c, int st) : col(c), stride(st), n(0), wkey(~0ull), wp(0), wd(INFINITY) {}
    UME_DEVI void reset(int) { n = 0; wkey = ~0ull; wp = 0; wd = INFINITY; }
    UME_DEVI bool full() const { return n == K_; }
    UME_DEVI float worst() const { return wd; }
    UME_DEVI static unsigned long long pack(float d, int j) {
        return ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)j;
    }
    UME_DEVI void group_max(int g, unsigned long long& mk, int& mp) const {
        mp = g * S;
        mk = col[mp * stride];
#pragma unroll
        for (int i = 1; i < S; ++i) {
            const int pos = g * S + i;
            if (pos < K_) {
                const unsigned long long e = col[pos * stride];
                if (e > mk) { mk = e; mp = pos; }
            }
        }
    }
    UME_DEVI void global_max() {
        wkey = gkey[0]; wp = gp[0];
#pragma unroll
        for (int g = 1; g < G; ++g)
            if (gkey[g] > wkey) { wkey = gkey[g]; wp = gp[g]; }
        wd = __uint_as_float((unsigned)(wkey >> 32));
    }
    UME_DEVI void consider(float d, int j) {
        const unsigned long long key = pack(d, j);
        if (n < K_) {
            col[n * stride] = key;
            if (++n == K_) {
#pragma unroll
                for (int g = 0; g < G; ++g) group_max(g, gkey[g], gp[g]);
                global_max();
            }
            return;
        }
        if (!(key < wkey)) return;
        col[wp * stride] = key;
        const int g = wp / S;
        unsigned long long mk; int mp;
        group_max(g, mk, mp);
#pragma unroll
        for (int i = 0; i < G; ++i)
            if (i == g) { gkey[i] = mk; gp[i] = mp; }
        global_max();
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
        // fewer than K usable rows: padded with index 0 / distance 0, as pytorch3d's knn_points pads (its outputs are
        // zero-initialised); callers mask by the row count
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
    // (rows with non-finite coordinates are not in the grid: with fewer than K usable rows the mean is over what was found)
    out[(size_t)b * N + i] = (top.n > 1) ? acc / (float)(top.n - 1) : 0.f;
}

// The reference's knn = 50 (utils/loc_utils.py:647-648): the K best unsorted in a shared-memory column (SmemGroupTopK)
// instead of a sorted per-thread array in local memory; the nearest entry (the one the reference drops as index 0 of
// the sorted list) is the minimum key.
template <int K_, bool kFma>
__global__ void __launch_bounds__(128) spatial_var_smem_kernel(GridView grid, const float* __restrict__ feat, int C,
                                                               float* __restrict__ out) {
    extern __shared__ unsigned long long sv_keys[];          // [K_][128]
    const int b = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int N = grid.N;
    const GridHeader h = grid.hdr[b];
    if (t >= h.n_sorted) return;
    const int* cs = grid.cell_start + (size_t)b * (grid.cells_cap + 1);
    const float4* sorted_b = grid.sorted + (size_t)b * N;
    const float4 me = sorted_b[t];
    const int i = __float_as_int(me.w);
    unsigned long long* col = sv_keys + threadIdx.x;
    SmemGroupTopK<K_> top(col, 128);
    grid_knn<kFma>(h, cs, sorted_b, me.x, me.y, me.z, top);
    unsigned long long kmin = ~0ull;
    for (int k = 0; k < top.n; ++k) kmin = min(kmin, col[k * 128]);
    const float* fb = feat + (size_t)b * N * C;
    const float* fi = fb + (size_t)i * C;
    float acc = 0.f;
    for (int k = 0; k < top.n; ++k) {
        const unsigned long long e = col[k * 128];
        if (e == kmin) continue;                               // (keys are distinct: one entry per target row)
        const float* fj = fb + (size_t)(unsigned)(e & 0xffffffffull) * C;
        float s = 0.f;
        for (int c = 0; c < C; c += 4) {
            const float4 a = ldg_f4(fi + c), v = ldg_f4(fj + c);
            const float dx = a.x - v.x, dy = a.y - v.y, dz = a.z - v.z, dw = a.w - v.w;
            s = fmaf(dx, dx, s); s = fmaf(dy, dy, s); s = fmaf(dz, dz, s); s = fmaf(dw, dw, s);
        }
        acc += sqrtf(s);
    }
    out[(size_t)b * N + i] = (top.n > 1) ? acc / (float)(top.n - 1) : 0.f;
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

#ifndef UME_CORR_TGT_PPC
#define UME_CORR_TGT_PPC(K) ((float)max(2, (K) / 8))   // target points per grid cell the ring search walks
#endif
#ifndef UME_CORR_SRC_PPC
#define UME_CORR_SRC_PPC 8.f                            // source points per cell: the CTA's queries are neighbours
#endif
#ifndef UME_CORR_MINB
#define UME_CORR_MINB 6            // CTAs per SM the register allocation is capped for
#endif

// TopKind: 0 = SmemGroupTopK<20> in the kernel's own neighbour columns (the reference's K), 1 = ArrayTopK<32> (any K <= 32)
template <int C4, int TopKind, bool kFma>
__global__ void __launch_bounds__(kCorrThreads, UME_CORR_MINB) corr_score_kernel(CorrParams p) {
    constexpr int KCAP = TopKind == 0 ? 20 : 32;
    // the K neighbours of every thread's query, [k][thread]: squared distance (then weight) and target row.  The
    // dot products are formed by the warp together — C4 lanes per (query, neighbour) read one feature row as one
    // 128-byte line (one L1 wavefront) instead of 32 lanes gathering 16 bytes from 32 different lines.  Nothing in
    // the hypothesis loop synchronises the CTA: every warp writes its own partial sum.
    __shared__ unsigned long long s_key[KCAP][kCorrThreads];   // (dist2 bits, then weight bits) << 32 | target row
    const GridHeader hs = p.src_grid.hdr[0];
    const GridHeader ht = p.tgt_grid.hdr[0];
    const int* cs = p.tgt_grid.cell_start;
    const float4* tgt_sorted = p.tgt_grid.sorted;
    const int t = blockIdx.x * kCorrThreads + threadIdx.x;
    const bool active = t < hs.n_sorted;
    float4 me = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) me = p.src_grid.sorted[t];
    const int my_row = __float_as_int(me.w);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int PP = 32 / C4;
    const int gq = lane / C4, ch = lane % C4;
    const int K = p.K;
    const int slot = blockIdx.x * (kCorrThreads / 32) + warp, nslots = gridDim.x * (kCorrThreads / 32);
    for (int hyp = blockIdx.y; hyp < p.n_hyp; hyp += gridDim.y) {
        const float* T = p.T + (size_t)hyp * 16;
        int n_found = 0;
        if (active) {
            // source_points @ R^T + t  (utils/loc_utils.py:626), row-times-matrix in fp32
            const float qx = fmaf(me.z, __ldg(T + 2), fmaf(me.y, __ldg(T + 1), me.x * __ldg(T + 0))) + __ldg(T + 3);
            const float qy = fmaf(me.z, __ldg(T + 6), fmaf(me.y, __ldg(T + 5), me.x * __ldg(T + 4))) + __ldg(T + 7);
            const float qz = fmaf(me.z, __ldg(T + 10), fmaf(me.y, __ldg(T + 9), me.x * __ldg(T + 8))) + __ldg(T + 11);
            if (TopKind == 0) {
                SmemGroupTopK<20> top(&s_key[0][threadIdx.x], kCorrThreads);
                grid_knn<kFma>(ht, cs, tgt_sorted, qx, qy, qz, top);
                n_found = top.n;
            } else {
                ArrayTopK<32> top;
                top.reset(K);
                grid_knn<kFma>(ht, cs, tgt_sorted, qx, qy, qz, top);
                n_found = top.n;
                for (int k = 0; k < n_found; ++k) s_key[k][threadIdx.x] = SmemGroupTopK<20>::pack(top.bd[k], top.bj[k]);
            }
            for (int k = 0; k < n_found; ++k) {
                const unsigned long long e = s_key[k][threadIdx.x];
                const float r = sqrtf(__uint_as_float((unsigned)(e >> 32))) * p.inv_sigma;       // |p - q| / sigma
                s_key[k][threadIdx.x] = ((unsigned long long)__float_as_uint(1.f / fmaf(r, r, 1.f)) << 32) | (e & 0xffffffffull);   // cauchy_kernel (:588-589)
            }
        }
        for (int k = n_found; k < K; ++k) s_key[k][threadIdx.x] = 0ull;   // (fewer than K rows in the cloud: weight 0, row 0)
        __syncwarp();
        float acc = 0.f;
        for (int q = 0; q < 32; ++q) {
            const int row_q = __shfl_sync(UME_FULL_MASK, my_row, q);
            const float4 a = ldg_f4(p.wf_src + (size_t)row_q * (C4 * 4) + 4 * ch);
            const int col = warp * 32 + q;
#pragma unroll 5
            for (int k0 = 0; k0 < K; k0 += PP) {
                const int k = min(k0 + gq, K - 1);
                const unsigned long long e = s_key[k][col];
                const float wgt = (k0 + gq < K) ? __uint_as_float((unsigned)(e >> 32)) : 0.f;
                const float4 g = ldg_f4(p.wf_tgt + (size_t)(unsigned)(e & 0xffffffffull) * (C4 * 4) + 4 * ch);
                float v = a.x * g.x;
                v = fmaf(a.y, g.y, v); v = fmaf(a.z, g.z, v); v = fmaf(a.w, g.w, v);
                acc = fmaf(v, wgt, acc);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(UME_FULL_MASK, acc, o);
        if (lane == 0) p.partial[(size_t)hyp * nslots + slot] = acc;
        __syncwarp();
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

// K = 20 (the reference's corr_num_nn) keeps the K best in the kernel's shared-memory columns; other K use the array version.
template <int C4, bool kFma>
void launch_corr_c(const CorrParams& p, int K, dim3 grid, cudaStream_t stream) {
    if (K == 20) corr_score_kernel<C4, 0, kFma><<<grid, kCorrThreads, 0, stream>>>(p);
    else corr_score_kernel<C4, 1, kFma><<<grid, kCorrThreads, 0, stream>>>(p);
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
    cudaError_t me = cudaMemsetAsync(out, 0, (size_t)B * N * sizeof(float), stream);      // rows with non-finite coordinates
    UME_REQUIRE(me == cudaSuccess, UME_ERR_CUDA, "ume_feature_spatial_var_f32: cudaMemsetAsync: %s", cudaGetErrorString(me));
    dim3 grid((unsigned)((N + 127) / 128), (unsigned)B);
    const bool fma = (flags & UME_FLAG_FMA_DIST) != 0;
    if (knn == 50) {
        const size_t smem = (size_t)50 * 128 * sizeof(unsigned long long);
        auto kern = fma ? spatial_var_smem_kernel<50, true> : spatial_var_smem_kernel<50, false>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        UME_REQUIRE(e == cudaSuccess, UME_ERR_CUDA, "ume_feature_spatial_var_f32: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        kern<<<grid, 128, smem, stream>>>(g, feat, C, out);
    } else if (fma) spatial_var_kernel<64, true><<<grid, 128, 0, stream>>>(g, feat, C, knn, out);
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

extern "C" size_t ume_corr_scores_workspace_bytes(int Ns, int Nt, int n_hyp) {
    if (Ns <= 0 || Nt <= 0 || n_hyp <= 0) return 0;
    const size_t nb = ((size_t)(Ns + ume::kCorrThreads - 1) / ume::kCorrThreads) * (ume::kCorrThreads / 32);   // one partial per warp
    return ume::grid_workspace_bytes(1, Ns, ume::kCellsCap) + ume::grid_workspace_bytes(1, Nt, ume::kCellsCap) +
           ume::align_up((size_t)n_hyp * nb * sizeof(float), 256) + 1024;
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
    CorrParams p;
    int rc = grid_build(src_pts, src_pts, 1, Ns, Ns, 0.f, -UME_CORR_SRC_PPC, kCellsCap, w, &p.src_grid, stream);
    if (rc != UME_OK) return rc;
    rc = grid_build(tgt_pts, tgt_pts, 1, Nt, Nt, 0.f, -UME_CORR_TGT_PPC(K), kCellsCap, w, &p.tgt_grid, stream);
    if (rc != UME_OK) return rc;
    const int nb = (Ns + kCorrThreads - 1) / kCorrThreads;
    const int nslots = nb * (kCorrThreads / 32);                      // one partial sum per warp
    p.partial = w.take<float>((size_t)n_hyp * nslots);
    UME_REQUIRE(w.ok(), UME_ERR_WORKSPACE, "ume_corr_scores_f32: workspace too small");
    p.wf_src = wf_src; p.wf_tgt = wf_tgt; p.T = T; p.n_hyp = n_hyp; p.K = K; p.inv_sigma = 1.f / sigma;
    // hypothesis groups: enough CTAs for ~8 waves of the 148 SMs
    int gy = (148 * 8 * 4 + nb - 1) / nb;
    gy = max(1, min(gy, min(n_hyp, 65535)));
    dim3 grid((unsigned)nb, (unsigned)gy);
    const bool fma = (flags & UME_FLAG_FMA_DIST) != 0;
    ProfScope prof(UME_PROF_CORR, stream);
    launch_corr(p, C, K, fma, grid, stream);
    corr_finalize_kernel<<<1, 256, 0, stream>>>(p.partial, n_hyp, nslots, 1.f / (float)Ns, score, best);
    count_launch(2);
    return check_launch("corr_score_kernel");
}
