// Hypothesis scoring, tile version (utils/loc_utils.py:592-631).
//
// Round 1 scored with one THREAD per (source point, hypothesis): a private ring search over the
// target grid and a private sorted list of the K best — 14 of 32 lanes active in the candidate loop,
// 7 of 32 in the insertion, 79 ms for 2500 hypotheses x 10 000 points.  Here a WARP owns a tile of <= 32
// source points that are neighbours in space (consecutive entries of one cell row of the source's
// cell-sorted array) and walks the hypotheses; per (tile, hypothesis):
//   1. the tile is transformed; its bounding box, grown by rho (from the tile's own point density),
//      selects a block of target cells whose points — one to three hundred — are staged ONCE in
//      shared memory;
//   2. the tile's queries are then served one after the other by the WHOLE warp, lane = candidate:
//      every lane computes the squared distances of its (<= kSlots) candidates to the query into
//      registers; the K nearest are found by bisecting a distance threshold with warp-wide counts
//      (compare + REDUX, no memory traffic, no sorted lists, no divergence) until exactly K
//      candidates lie below it; they are squeezed into a K-entry list (ballot + prefix) and lanes
//      0..K-1 each form one feature dot product — the K dot products of a query run in parallel;
//   3. exactness: a query accepts its K only if the K-th distance is smaller than its distance to
//      the faces of the staged block (nothing outside can be closer); queries that cannot be served
//      this way (sparse or far-away landing zone, more candidates than fit) are left to the exact
//      round-1 ring search, run by their own lanes at the end of the tile.
// Ties in distance are broken by the lower row, as pytorch3d's row-order scan with strict '<' does.
// (A first tile version kept per-lane histograms and neighbour lists in shared memory, one query per
// lane: shared-memory atomics at 2 cycles per lane, then read-modify-write chains, 6 warps per SM —
// it was slower than round 1.)
#pragma once
#include "ume_common.cuh"

namespace ume {
namespace tile {

constexpr int kWarps = 4;            // warps (= tiles) per CTA
constexpr int kSlotsMax = 10;        // candidates per lane
constexpr int kCap = 32 * kSlotsMax; // staged candidates per tile
constexpr int kKMax = 32;
constexpr int kMaxRuns = 64;

template <int C4>
struct WarpSmem {
    float4 cand[kCap];                       // (x, y, z, row)
    float sel_d2[kKMax];                     // the current query's K neighbours: squared distance ...
    int sel_row[kKMax];                      // ... and target row
    int run_start[kMaxRuns], run_off[kMaxRuns + 1];
    float4 sf[32][C4 + 1];                   // the tile's weighted source features, [query][chunk] (padded)
};

UME_DEVI int f2key(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
UME_DEVI float key2f(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }
UME_DEVI float warp_min(float v) { return key2f(__reduce_min_sync(UME_FULL_MASK, f2key(v))); }
UME_DEVI float warp_max(float v) { return key2f(__reduce_max_sync(UME_FULL_MASK, f2key(v))); }

// (d, j) < (d', j') in the order of a row-order scan with strict '<'
UME_DEVI bool before(float d, int j, float d2, int j2) { return d < d2 || (d == d2 && j < j2); }

// Stage the target points of the cell block [c0, c1] (inclusive, per axis) into sm.cand.  Returns the
// number of candidates, or -1 when the block has more runs or points than fit.
template <int C4>
UME_DEVI int stage_block(WarpSmem<C4>& sm, const GridHeader& h, const int* __restrict__ cs,
                         const float4* __restrict__ sorted, int cx0, int cx1, int cy0, int cy1, int cz0, int cz1) {
    const int lane = threadIdx.x & 31;
    const int nyr = cy1 - cy0 + 1, nrows = nyr * (cz1 - cz0 + 1);
    if (nrows > kMaxRuns) return -1;
    int total = 0;
    for (int r0 = 0; r0 < nrows; r0 += 32) {
        const int r = r0 + lane;
        int s = 0, n = 0;
        if (r < nrows) {
            const int base = ((cz0 + r / nyr) * h.ny + cy0 + r % nyr) * h.nx;
            s = __ldg(&cs[base + cx0]);
            n = __ldg(&cs[base + cx1 + 1]) - s;
        }
        int incl = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(UME_FULL_MASK, incl, o);
            if (lane >= o) incl += v;
        }
        if (r < nrows) { sm.run_start[r] = s; sm.run_off[r] = total + incl - n; }
        total += __shfl_sync(UME_FULL_MASK, incl, 31);
    }
    if (lane == 0) sm.run_off[nrows] = total;
    __syncwarp();
    if (total > kCap) return -1;
    // every lane copies whole runs (a run is a handful of consecutive float4)
    for (int r = lane; r < nrows; r += 32) {
        const int s = sm.run_start[r], o = sm.run_off[r], n = sm.run_off[r + 1] - o;
        for (int k = 0; k < n; ++k) sm.cand[o + k] = __ldg(&sorted[s + k]);
    }
    __syncwarp();
    return total;
}

}  // namespace tile
}  // namespace ume
