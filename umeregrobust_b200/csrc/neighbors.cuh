// CTA-level neighbourhood collection shared by the ball-query and the fused moment kernels.
//
// One CTA serves one query.  Candidates come from the search grid (contiguous runs of the
// cell-sorted point array, one run per (y,z) cell row); membership is decided by the exact fp32
// distance test; "the first K rows in row order" (pytorch3d ball_query semantics, evaluate.py:51)
// is recovered WITHOUT scanning in row order: the K smallest row indices among the in-radius
// candidates are found with a two-level counting select (histogram over index bins, then a bitmap
// inside the crossing bin), which works for any number of hits.
#pragma once
#include "ume_common.cuh"

namespace ume {

static constexpr int kMaxRows = 64;       // (2*div+2)^2 <= 36 cell rows per query
static constexpr int kHistBins = 512;
static constexpr int kBitmapWords = 128;  // bin width <= 4096 indices  ->  N <= 512*4096 (= 2 M points per cloud)
static constexpr int kChunkCap = 512;     // chunk-table window: runs of <= 32 candidates handed to the warps

struct CollectSmem {
    int seg_start[kMaxRows];
    int seg_prefix[kMaxRows + 1];
    int seg_chunk0[kMaxRows + 1];         // exclusive prefix of the rows' chunk counts
    unsigned chunk[kChunkCap];            // (position in the sorted array << 5) | (candidates - 1)
    unsigned chunk_mask[kChunkCap];       // hit ballot of every chunk of the window (scan_mark)
    int chunk_off[kChunkCap];             // list slot of the chunk's first hit (chunk_offsets)
    unsigned hist[kHistBins];
    unsigned bitmap[kBitmapWords];
    int warp_cnt[32];
    int count;        // hits appended so far (may exceed the list capacity)
    int sel_bin, sel_below, sel_T;
    int nrows, total, nchunks;
};

// One row's share of the chunk table for the window of chunk ids [w0, w0 + kChunkCap): the run
// (start s, n candidates) is cut into pieces of 32; c0 = id of its first chunk.
UME_DEVI void fill_chunks(CollectSmem& sm, int s, int n, int c0, int w0) {
    const int nch = (n + 31) >> 5;
    const int k1 = min(nch, w0 + kChunkCap - c0);
    for (int k = max(0, w0 - c0); k < k1; ++k)
        sm.chunk[c0 + k - w0] = ((unsigned)(s + 32 * k) << 5) | (unsigned)(min(32, n - 32 * k) - 1);
}

// ---------------------------------------------------------------- phase 0: candidate runs
// The cells a query's ball can touch, as <= kMaxRows runs of the cell-sorted array: one run per
// (y,z) cell row, its x range trimmed to the chord of the (slightly inflated) ball.
struct RowSetup {
    int cx0, cx1, cy0, cz0, nyr, nrows;
    float rr2, mx;
};

UME_DEVI RowSetup row_setup(const GridHeader& h, float kx, float ky, float kz, float radius) {
    RowSetup rs;
    const float r = fabsf(radius);
    const float mx = r * 1e-4f + fabsf(kx) * 1e-6f + 1e-7f;
    const float my = r * 1e-4f + fabsf(ky) * 1e-6f + 1e-7f;
    const float mz = r * 1e-4f + fabsf(kz) * 1e-6f + 1e-7f;
    rs.cx0 = cell_coord(kx - r - mx, h.ox, h.inv_s, h.nx);
    rs.cx1 = cell_coord(kx + r + mx, h.ox, h.inv_s, h.nx);
    rs.cy0 = cell_coord(ky - r - my, h.oy, h.inv_s, h.ny);
    const int cy1 = cell_coord(ky + r + my, h.oy, h.inv_s, h.ny);
    rs.cz0 = cell_coord(kz - r - mz, h.oz, h.inv_s, h.nz);
    const int cz1 = cell_coord(kz + r + mz, h.oz, h.inv_s, h.nz);
    rs.nyr = cy1 - rs.cy0 + 1;
    rs.nrows = min(rs.nyr * (cz1 - rs.cz0 + 1), kMaxRows);   // the cap cannot bind: cell >= radius/2 (grid_params_kernel)
    const float rr = (r + 4.f * (mx + my + mz));
    rs.rr2 = rr * rr;
    rs.mx = mx;
    return rs;
}

// run of candidate row `row` (< rs.nrows): start `s` in the sorted array and length `n`
UME_DEVI void row_run(const RowSetup& rs, const GridHeader& h, const int* __restrict__ cs, float kx, float ky,
                      float kz, int row, int& s, int& n) {
    s = 0;
    n = 0;
    const int iy = rs.cy0 + row % rs.nyr, iz = rs.cz0 + row / rs.nyr;
    // prune rows / trim the x range by the distance from the query to the cell slab; the
    // outermost cells also hold clamped coordinates, so they are treated as unbounded
    float dy = 0.f, dz = 0.f;
    {
        float lo = h.oy + (float)iy * h.s, hi = lo + h.s;
        if (iy > 0 && ky < lo) dy = lo - ky;
        if (iy < h.ny - 1 && ky > hi) dy = ky - hi;
        lo = h.oz + (float)iz * h.s; hi = lo + h.s;
        if (iz > 0 && kz < lo) dz = lo - kz;
        if (iz < h.nz - 1 && kz > hi) dz = kz - hi;
    }
    const float rem2 = rs.rr2 - dy * dy - dz * dz;
    if (rem2 > 0.f) {
        const float half = sqrtf(rem2) * 1.0001f + rs.mx;
        const int tx0 = max(rs.cx0, cell_coord(kx - half, h.ox, h.inv_s, h.nx));
        const int tx1 = min(rs.cx1, cell_coord(kx + half, h.ox, h.inv_s, h.nx));
        if (tx1 >= tx0) {
            const int base = (iz * h.ny + iy) * h.nx;
            s = cs[base + tx0];
            n = cs[base + tx1 + 1] - s;
        }
    }
}

template <int NT>
UME_DEVI void collect_rows(CollectSmem& sm, const GridHeader& h, const int* __restrict__ cs, float kx,
                           float ky, float kz, float radius) {
    static_assert(NT >= 2 * kMaxRows && kMaxRows == 64, "one thread per candidate row (two whole warps), the others clear the histogram");
    const bool row_thread = threadIdx.x < kMaxRows;       // warp-uniform: kMaxRows is a multiple of 32
    const int row = threadIdx.x;
    int s = 0, n = 0, incl = 0, cincl = 0;
    if (row_thread) {
        const RowSetup rs = row_setup(h, kx, ky, kz, radius);
        const int nrows = rs.nrows;
        if (row < nrows) {
            row_run(rs, h, cs, kx, ky, kz, row, s, n);
            sm.seg_start[row] = s;
        }
        // inclusive prefixes of the run lengths and of the chunk counts over the (<= 64) rows: two
        // warps, shuffle scans, the second warp fixed up after the barrier
        const int lane = threadIdx.x & 31;
        incl = n;
        cincl = (n + 31) >> 5;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(UME_FULL_MASK, incl, o);
            const int c = __shfl_up_sync(UME_FULL_MASK, cincl, o);
            if (lane >= o) { incl += v; cincl += c; }
        }
        if (threadIdx.x == 31) { sm.warp_cnt[0] = incl; sm.warp_cnt[1] = cincl; }
        if (threadIdx.x == 0) { sm.nrows = nrows; sm.count = 0; sm.seg_prefix[0] = 0; sm.seg_chunk0[0] = 0; }
    } else {
        // the warps without rows clear the row-index histogram the scan fills (select_kth_index)
        for (int i = threadIdx.x - kMaxRows; i < kHistBins; i += NT - kMaxRows) sm.hist[i] = 0;
    }
    __syncthreads();                                       // one barrier, reached by every thread on the same path
    if (row_thread) {
        if (threadIdx.x >= 32) { incl += sm.warp_cnt[0]; cincl += sm.warp_cnt[1]; }
        sm.seg_prefix[row + 1] = incl;
        sm.seg_chunk0[row + 1] = cincl;
        if (row == kMaxRows - 1) { sm.total = incl; sm.nchunks = cincl; }   // rows >= nrows are empty
        fill_chunks(sm, s, n, cincl - ((n + 31) >> 5), 0);
    }
    __syncthreads();
}

// ---------------------------------------------------------------- candidate walk
// visit(hit, ex, ey, ez, d2, row_index) is called by ALL threads of the CTA the same number of
// times (so it may use warp votes and __syncthreads); e = point - query.
template <bool kFma, int NT, typename Visit>
UME_DEVI void scan_candidates(const CollectSmem& sm, const float4* __restrict__ sorted_b, float kx, float ky,
                              float kz, float r2, Visit visit) {
    const int total = sm.total;
    int seg = 0;
    int seg_end = sm.seg_prefix[1];                       // run bounds live in registers; smem is only
    int seg_off = sm.seg_start[0];                        // touched when a thread crosses into the next run
    for (int base = 0; base < total; base += 2 * NT) {
        const int j0 = base + threadIdx.x, j1 = j0 + NT;
        const bool v0 = j0 < total, v1 = j1 < total;
        float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f), c1 = c0;
        if (v0) {
            while (j0 >= seg_end) { ++seg; seg_end = sm.seg_prefix[seg + 1]; seg_off = sm.seg_start[seg] - sm.seg_prefix[seg]; }
            c0 = __ldg(&sorted_b[j0 + seg_off]);
        }
        if (v1) {
            while (j1 >= seg_end) { ++seg; seg_end = sm.seg_prefix[seg + 1]; seg_off = sm.seg_start[seg] - sm.seg_prefix[seg]; }
            c1 = __ldg(&sorted_b[j1 + seg_off]);
        }
        const float ex0 = __fsub_rn(c0.x, kx), ey0 = __fsub_rn(c0.y, ky), ez0 = __fsub_rn(c0.z, kz);
        const float ex1 = __fsub_rn(c1.x, kx), ey1 = __fsub_rn(c1.y, ky), ez1 = __fsub_rn(c1.z, kz);
        const float d0 = dist2_ordered<kFma>(ex0, ey0, ez0);
        const float d1 = dist2_ordered<kFma>(ex1, ey1, ez1);
        visit(v0 && (d0 < r2), ex0, ey0, ez0, d0, __float_as_int(c0.w));
        if (base + NT < total) visit(v1 && (d1 < r2), ex1, ey1, ez1, d1, __float_as_int(c1.w));   // block-uniform
    }
}

// ---------------------------------------------------------------- first pass: mark, offsets, fill
// The candidates of the query are cut into chunks of <= 32 consecutive entries of one cell row
// (chunk table, built once per query by the row threads).  The list must not depend on how the
// warps happen to be scheduled — its order decides the order of the fp32 sums downstream, and the
// results are meant to be bit-reproducible — so hits are placed at positions that are a pure
// function of the chunk table:
//   scan_mark      every warp takes chunks warp, warp + NW, ...: two per iteration, the next two
//                  already loading (four 16-byte loads in flight per lane); exact distance test; the
//                  chunk's hit ballot goes to chunk_mask[], every hit is counted in the histogram of
//                  (row index >> shift): level 1 of the counting select comes for free;
//   chunk_offsets  exclusive prefix of the ballots' populations = the slot of each chunk's first hit;
//   scan_fill      chunks with hits are read again (L1 hits) and their hits written to their slots.
// Trip counts differ between warps, so nothing inside scan_mark / scan_fill may synchronise the CTA.
template <bool kFma, int NT>
UME_DEVI void scan_mark(CollectSmem& sm, const float4* __restrict__ sorted_b, float kx, float ky, float kz, float r2,
                        int nch, int shift) {
    constexpr int NW = NT / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    auto fetch = [&](int c, float4& v, int& cnt) {
        v = make_float4(0.f, 0.f, 0.f, 0.f);
        cnt = 0;
        if (c < nch) {                                   // warp-uniform
            const unsigned e = sm.chunk[c];
            cnt = (int)(e & 31u) + 1;
            if (lane < cnt) v = __ldg(&sorted_b[(e >> 5) + lane]);
        }
    };
    float4 a, b;
    int ca, cb;
    fetch(warp, a, ca);
    fetch(warp + NW, b, cb);
    for (int c = warp; c < nch; c += 2 * NW) {
        float4 na, nb;
        int nca, ncb;
        fetch(c + 2 * NW, na, nca);
        fetch(c + 3 * NW, nb, ncb);
        const float ax = __fsub_rn(a.x, kx), ay = __fsub_rn(a.y, ky), az = __fsub_rn(a.z, kz);
        const float bx = __fsub_rn(b.x, kx), by = __fsub_rn(b.y, ky), bz = __fsub_rn(b.z, kz);
        const bool ha = (lane < ca) && (dist2_ordered<kFma>(ax, ay, az) < r2);
        const bool hb = (lane < cb) && (dist2_ordered<kFma>(bx, by, bz) < r2);
        const unsigned ma = __ballot_sync(UME_FULL_MASK, ha), mb = __ballot_sync(UME_FULL_MASK, hb);
        if (ha) atomicAdd(&sm.hist[__float_as_int(a.w) >> shift], 1u);     // level 1 of the counting select
        if (hb) atomicAdd(&sm.hist[__float_as_int(b.w) >> shift], 1u);
        if (lane == 0) {
            sm.chunk_mask[c] = ma;
            if (c + NW < nch) sm.chunk_mask[c + NW] = mb;
        }
        a = na; b = nb;
        ca = nca; cb = ncb;
    }
}

// chunk_off[c] = sm.count + hits of the chunks before c; sm.count += hits of the window.
// Called by the whole CTA between two barriers of its own.
template <int NT>
UME_DEVI void chunk_offsets(CollectSmem& sm, int nch) {
    static_assert(kChunkCap % NT == 0, "chunks per thread");
    constexpr int per = kChunkCap / NT, NW = NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __syncthreads();                                       // the window's ballots are visible
    const int base = sm.count;
    int v[per], s = 0;
#pragma unroll
    for (int i = 0; i < per; ++i) {
        v[i] = (tid * per + i < nch) ? __popc(sm.chunk_mask[tid * per + i]) : 0;
        s += v[i];
    }
    int incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(UME_FULL_MASK, incl, o);
        if (lane >= o) incl += u;
    }
    if (lane == 31) sm.warp_cnt[warp] = incl;
    __syncthreads();
    int wbase = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        const int c = sm.warp_cnt[w];
        if (w < warp) wbase += c;
        tot += c;
    }
    int at = base + wbase + incl - s;
#pragma unroll
    for (int i = 0; i < per; ++i) {
        if (tid * per + i < nch) sm.chunk_off[tid * per + i] = at;
        at += v[i];
    }
    __syncthreads();                                       // everyone has read sm.count / warp_cnt
    if (tid == 0) sm.count = base + tot;
}

template <int NT>
UME_DEVI void scan_fill(const CollectSmem& sm, float4* list, int cap, const float4* __restrict__ sorted_b, float kx,
                        float ky, float kz, int nch) {
    constexpr int NW = NT / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = lanemask_lt();
#pragma unroll 2
    for (int c = warp; c < nch; c += NW) {
        const unsigned m = sm.chunk_mask[c];
        const int off = sm.chunk_off[c];
        if (m == 0u || off >= cap) continue;             // warp-uniform
        if ((m >> lane) & 1u) {
            const int pos = off + __popc(m & lt);
            if (pos < cap) {
                const float4 v = __ldg(&sorted_b[(sm.chunk[c] >> 5) + lane]);
                list[pos] = make_float4(__fsub_rn(v.x, kx), __fsub_rn(v.y, ky), __fsub_rn(v.z, kz), v.w);
            }
        }
    }
}

// ---------------------------------------------------------------- counting select
// Given that more than K candidates are in radius, find T = the K-th smallest row index among
// them.  sm.hist already holds the histogram of (row index >> shift) over ALL in-radius candidates
// (filled by scan_mark, visible after a barrier).  `each(f)` must call f(row_index) once per
// in-radius candidate (any thread, any order): it resolves the bin the K-th index falls in.
template <int NT, typename Each>
UME_DEVI int select_kth_index(CollectSmem& sm, int K, int shift, Each each) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {
        constexpr int per = kHistBins / NT;
        const int lo = tid * per;
        int s = 0;
#pragma unroll
        for (int i = 0; i < per; ++i) s += (int)sm.hist[lo + i];
        int incl = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int v = __shfl_up_sync(UME_FULL_MASK, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) sm.warp_cnt[warp] = incl;
        __syncthreads();
        int wbase = 0;
        for (int w = 0; w < warp; ++w) wbase += sm.warp_cnt[w];
        const int excl = wbase + incl - s;
        if (excl < K && K <= excl + s) {
            int run = excl;
            for (int i = 0; i < per; ++i) {
                const int hcount = (int)sm.hist[lo + i];
                if (run + hcount >= K) {
                    sm.sel_bin = lo + i;
                    sm.sel_below = run;
                    break;
                }
                run += hcount;
            }
        }
    }
    const int nwords = max(1, (1 << shift) >> 5);
    for (int i = tid; i < nwords; i += NT) sm.bitmap[i] = 0;
    __syncthreads();
    const int bin = sm.sel_bin;
    const int need = K - sm.sel_below;                 // >= 1 indices wanted from the crossing bin
    const unsigned wmask = (1u << shift) - 1u;
    each([&](int idx) {
        if ((idx >> shift) == bin) {
            const unsigned off = (unsigned)idx & wmask;
            atomicOr(&sm.bitmap[off >> 5], 1u << (off & 31));
        }
    });
    __syncthreads();
    if (warp == 0) {
        int run = 0;
        bool done = false;
        for (int t = 0; t * 32 < nwords && !done; ++t) {
            const unsigned word = (t * 32 + lane < nwords) ? sm.bitmap[t * 32 + lane] : 0u;
            const int pc = __popc(word);
            int incl = pc;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int v = __shfl_up_sync(UME_FULL_MASK, incl, o);
                if (lane >= o) incl += v;
            }
            const int excl = run + incl - pc;
            const bool mine = (excl < need) && (need <= excl + pc);
            if (mine) {
                unsigned wd = word;
                for (int k = need - excl; k > 1; --k) wd &= wd - 1;   // drop the k-1 lowest set bits
                const int bit = __ffs(wd) - 1;
                sm.sel_T = (bin << shift) + (t * 32 + lane) * 32 + bit;
            }
            done = __any_sync(UME_FULL_MASK, mine);
            run += __shfl_sync(UME_FULL_MASK, incl, 31);
        }
    }
    __syncthreads();
    return sm.sel_T;
}

// In-place stable-free compaction of list[0..count): keep entries whose row index <= T.
template <int NT>
UME_DEVI int compact_list(CollectSmem& sm, float4* list, int count, int T) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = NT / 32;
    int out = 0;
    for (int base = 0; base < count; base += NT) {
        const int i = base + tid;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        bool keep = false;
        if (i < count) {
            v = list[i];
            keep = __float_as_int(v.w) <= T;
        }
        const unsigned m = __ballot_sync(UME_FULL_MASK, keep);
        if (lane == 0) sm.warp_cnt[warp] = __popc(m);
        __syncthreads();
        int wbase = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const int c = sm.warp_cnt[w];
            if (w < warp) wbase += c;
            tot += c;
        }
        if (keep) list[out + wbase + __popc(m & lanemask_lt())] = v;
        out += tot;
        __syncthreads();
    }
    return out;
}

// ---------------------------------------------------------------- driver
// Collects the neighbourhood of one query and hands it to `flush(len, T)`: list[0..len) holds
// (ex, ey, ez, row index) entries, of which those with row index <= T are the neighbours
// (T = INT_MAX when every entry counts).  kCompact = true squeezes the list first so that all `len`
// entries are neighbours (the ball-query kernel sorts them); kCompact = false leaves the filtering
// to the consumer (the moment kernel predicates its gather: no extra pass, no barriers).
// `flush` runs once unless the hits overflow the list AND K > cap.
// Returns the number of neighbours = min(K, #in radius).
template <bool kFma, int NT, bool kCompact, typename Flush>
UME_DEVI int collect_neighbors(CollectSmem& sm, float4* list, int cap, const GridHeader& h,
                               const int* __restrict__ cs, const float4* __restrict__ sorted_b, int N, float kx,
                               float ky, float kz, float radius, int K, Flush flush) {
    const float r2 = __fmul_rn(radius, radius);
    collect_rows<NT>(sm, h, cs, kx, ky, kz, radius);
    const int nchunks = sm.nchunks;
    const int shift = max(0, 23 - __clz(N - 1));         // smallest shift with (N-1) >> shift < kHistBins
    for (int w0 = 0;;) {
        const int nch = min(kChunkCap, nchunks - w0);
        scan_mark<kFma, NT>(sm, sorted_b, kx, ky, kz, r2, nch, shift);
        chunk_offsets<NT>(sm, nch);
        scan_fill<NT>(sm, list, cap, sorted_b, kx, ky, kz, nch);
        w0 += kChunkCap;
        if (w0 >= nchunks) break;                        // the usual case: one window
        __syncthreads();
        if (threadIdx.x < kMaxRows)
            fill_chunks(sm, sm.seg_start[threadIdx.x], sm.seg_prefix[threadIdx.x + 1] - sm.seg_prefix[threadIdx.x],
                        sm.seg_chunk0[threadIdx.x], w0);
        __syncthreads();
    }
    __syncthreads();
    const int count = sm.count;
    if (count <= cap) {
        int len = count, T = 0x7fffffff;
        if (count > K) {
            T = select_kth_index<NT>(sm, K, shift, [&](auto f) {
                for (int i = threadIdx.x; i < count; i += NT) f(__float_as_int(list[i].w));
            });
            if (kCompact) {
                len = compact_list<NT>(sm, list, count, T);
                T = 0x7fffffff;
            }
        }
        flush(len, T);
        return count > K ? K : count;
    }
    // Overflow: more hits than the list holds.  Everything is recomputed from the candidates.
    int T = 0x7fffffff;
    if (count > K) {
        T = select_kth_index<NT>(sm, K, shift, [&](auto f) {
            scan_candidates<kFma, NT>(sm, sorted_b, kx, ky, kz, r2,
                                      [&](bool hit, float, float, float, float, int idx) {
                                          if (hit) f(idx);
                                      });
        });
    }
    __syncthreads();
    int len = 0;   // block-uniform running length of the list
    scan_candidates<kFma, NT>(sm, sorted_b, kx, ky, kz, r2,
                              [&](bool hit, float ex, float ey, float ez, float, int idx) {
                                  // slots in thread order (prefix over the warps' ballots): the list, and with
                                  // it the order of the sums, does not depend on the warps' timing
                                  const bool acc = hit && idx <= T;
                                  const unsigned m = __ballot_sync(UME_FULL_MASK, acc);
                                  const int warp = threadIdx.x >> 5;
                                  if ((threadIdx.x & 31) == 0) sm.warp_cnt[warp] = __popc(m);
                                  __syncthreads();
                                  int wbase = 0, n_acc = 0;
#pragma unroll
                                  for (int w = 0; w < NT / 32; ++w) {
                                      const int c = sm.warp_cnt[w];
                                      if (w < warp) wbase += c;
                                      n_acc += c;
                                  }
                                  __syncthreads();
                                  if (len + n_acc > cap) {
                                      flush(len, 0x7fffffff);
                                      __syncthreads();
                                      len = 0;
                                  }
                                  if (acc) list[len + wbase + __popc(m & lanemask_lt())] = make_float4(ex, ey, ez, __int_as_float(idx));
                                  len += n_acc;
                              });
    __syncthreads();
    flush(len, 0x7fffffff);
    return count > K ? K : count;
}

}  // namespace ume
