// Fused radius-neighbourhood gather + UME moment build, ONE WARP PER KEYPOINT (evaluate.py:50-60).
//
// No shared-memory neighbour list and no CTA barrier: a warp walks its keypoint's candidates (the
// runs of the cell-sorted array the grid hands out, 32 per step, lane = candidate) twice.
//   pass 1  exact fp32 distance test, number of hits, 256-bin histogram of the hits' row indices;
//           if more than K points are in radius the histogram gives the bin [T_lo, T_hi) holding
//           the K-th smallest row index (refined with another histogram pass over that bin only
//           while it holds more than 32 hits: bins shrink 256x per level, so <= 3 levels);
//   pass 2  hits with row < T_lo are neighbours for certain: they are squeezed into a 128-entry
//           ring in shared memory (ballot + prefix) and consumed from there RPW rows per warp
//           instruction — LPR lanes x LDG.128 per feature row — into 16 packed fp32 accumulators
//           per lane; hits inside [T_lo, T_hi) (<= 32, one per lane) are ranked at the end and the
//           smallest `need` of them follow.
// "The first K rows in row order" (pytorch3d ball_query) is therefore reproduced exactly for any
// K, N and hit count, and the warps of an SM never wait for each other: 32 independent gather
// streams per SM hide the L2 latency that a CTA-per-keypoint kernel serialises behind barriers.
// One single-warp CTA per keypoint (32 resident per SM): the hardware block scheduler refills a
// warp slot the moment its keypoint is done, so the 30x spread in neighbourhood sizes never leaves
// a slot idle.  (UME_WARPK_PERSISTENT=1 builds the variant where warps pull keypoints from a global
// counter instead; it measured slower: its extra live registers make ptxas serialise the gather loads.)
#pragma once
#include "neighbors.cuh"

// the kernel is one body for three modes: code after a mode's early exit is unreachable in that
// instantiation only
#pragma nv_diag_suppress code_is_unreachable

namespace ume {
namespace warpk {

#ifndef UME_WARPK_WARPS
#define UME_WARPK_WARPS 1          // warps per CTA
#endif
#ifndef UME_WARPK_MINB
#define UME_WARPK_MINB 32          // CTAs per SM the register allocation is capped for
#endif
#ifndef UME_WARPK_PERSISTENT
#define UME_WARPK_PERSISTENT 0     // 1: warps pull keypoints from a global counter; 0: one keypoint per warp
#endif
#ifndef UME_WARPK_D1
#define UME_WARPK_D1 4             // candidate chunks in flight per warp in the histogram passes
#endif
#ifndef UME_WARPK_D2
#define UME_WARPK_D2 2             // ... in the gather pass (next to the feature-row loads)
#endif
#ifndef UME_WARPK_UNROLL
#define UME_WARPK_UNROLL 4         // feature-row loads in flight per lane
#endif
#if UME_WARPK_PERSISTENT
#define UME_WARPK_NEXT continue
#else
#define UME_WARPK_NEXT break
#endif
constexpr int kWarps = UME_WARPK_WARPS;
constexpr int kChunks = 128;       // chunk-table window per warp
constexpr int kBins = 256;
constexpr int kRing = 128;
constexpr int kMaybe = 32;
constexpr int kPad = 4;          // chunk-table entries past the end that prefetches may read

struct WarpSmem {
    int seg_start[kMaxRows];
    int seg_n[kMaxRows];
    int seg_c0[kMaxRows];
    unsigned chunk[kChunks + 2 * kPad];   // (position in the sorted array << 5) | (candidates - 1), padded
    unsigned hist[kBins];
    float4 ring[kRing];
    float4 maybe[kMaybe];
};

struct Params {
    GridView grid;
    const float* kpts;    // (B,n,3)
    const float* feat;    // (B,N,C)
    const float* kpts2;   // clouds [Bs, B): keypoints / features of a second batch handled by the same launch
    const float* feat2;   //   (null and Bs = B for a single batch); F, Fc, count cover all B clouds
    int Bs;
    float* F;             // (B,n,C,4)
    float* Fc;            // (B,n,C,4) or null
    int32_t* count;       // (B,n) or null
    unsigned long long* next;   // work counter (zeroed before the launch; persistent builds only)
    const float* gF;      // kBackward: (B,n,C,4) gradient of the RAW moments [sum f | sum f x^T]
    float* grad_feat;     // kBackward: (B,N,C), accumulated into
    int raw;              // kForward: 1 = leave out the normalisation of evaluate.py:59
    int n, K;
    long long total;      // B*n
    float radius;
};

// ---- shared memory through explicit 32-bit addresses: one register holds the warp's base, every
// access is a single LDS/STS/ATOMS (nvcc otherwise re-derives the shared window base per access)
UME_DEVI unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
UME_DEVI unsigned lds_u32(unsigned a) {
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
UME_DEVI float4 lds_f4(unsigned a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
UME_DEVI void sts_u32(unsigned a, unsigned v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v)); }
UME_DEVI void sts_f4_if(bool p, unsigned a, float x, float y, float z, float w) {
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %0, 0; @q st.shared.v4.f32 [%1], {%2, %3, %4, %5}; }"
                 ::"r"((unsigned)p), "r"(a), "f"(x), "f"(y), "f"(z), "f"(w));
}
UME_DEVI void inc_if(bool p, unsigned a) {
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %0, 0; @q red.shared.add.u32 [%1], 1; }" ::"r"((unsigned)p), "r"(a));
}

// chunk-table window [w0, w0 + kChunks) from the per-row runs (lanes own rows lane, lane + 32)
UME_DEVI void fill_window(WarpSmem& sm, int nrows, int nchunks, int w0) {
    const int lane = threadIdx.x & 31;
    for (int row = lane; row < nrows; row += 32) {
        const int s = sm.seg_start[row], n = sm.seg_n[row], c0 = sm.seg_c0[row];
        const int nch = (n + 31) >> 5;
        const int k1 = min(nch, w0 + kChunks - c0);
        for (int k = max(0, w0 - c0); k < k1; ++k)
            sm.chunk[c0 + k - w0] = ((unsigned)(s + 32 * k) << 5) | (unsigned)(min(32, n - 32 * k) - 1);
    }
    if (lane < kPad) sm.chunk[min(kChunks, nchunks - w0) + lane] = 0u;   // prefetches past the end read entry 0
    __syncwarp();
}

// Walk every candidate of the query: visit(hit, ex, ey, ez, row) is called by the converged warp once
// per chunk (lane = candidate; e = point - query).  D chunks are in flight: the loop is unrolled by
// D so that the buffers rotate without register moves, the table is padded so that the prefetch
// needs no bounds check, and lanes past the end of a short chunk re-read its last entry so that
// the load needs no predicate.
template <bool kFma, int D, typename Visit>
UME_DEVI void scan(WarpSmem& sm, int nrows, int nchunks, int& loaded_w0, const float4* __restrict__ sorted_b,
                   float kx, float ky, float kz, float r2, Visit visit) {
    static_assert(D <= kPad, "table padding");
    const int lane = threadIdx.x & 31;
    const unsigned chunk_a = smem_u32(sm.chunk);
    for (int w0 = 0; w0 < nchunks; w0 += kChunks) {
        if (loaded_w0 != w0) {
            __syncwarp();
            fill_window(sm, nrows, nchunks, w0);
            loaded_w0 = w0;
        }
        const int nch = min(kChunks, nchunks - w0);
        float4 buf[D];
        int last[D];                                     // candidates - 1 of the chunk in buf[i]
        auto fetch = [&](int c, int i) {
            const unsigned e = lds_u32(chunk_a + 4u * (unsigned)c);
            last[i] = (int)(e & 31u);
            buf[i] = __ldg(sorted_b + ((e >> 5) + (unsigned)min(lane, last[i])));
        };
#pragma unroll
        for (int i = 0; i < D; ++i) fetch(i, i);
        for (int c = 0; c < nch; c += D) {
#pragma unroll
            for (int i = 0; i < D; ++i) {
                if (i > 0 && c + i >= nch) break;            // warp-uniform
                const float ex = __fsub_rn(buf[i].x, kx), ey = __fsub_rn(buf[i].y, ky), ez = __fsub_rn(buf[i].z, kz);
                const bool hit = (dist2_ordered<kFma>(ex, ey, ez) < r2) & (lane <= last[i]);
                visit(hit, ex, ey, ez, __float_as_int(buf[i].w));
                fetch(c + i + D, i);
            }
        }
    }
}

enum Mode { kForward = 0, kBackward = 1, kCountOnly = 2 };

UME_DEVI void red_add_f4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// MODE kForward : F (and Fc, count) as documented in include/umereg_b200.h
//      kBackward: the same neighbourhoods, walked the same way; instead of reading a neighbour's
//                 feature row the warp adds  gF[i,c,0] + gF[i,c,1:4] . x_j  to grad_feat[j,c]
//                 (d/df of the raw moments; 8 lanes x RED.128 per row)
//      kCountOnly: pass 1 only -> count
template <int LPR, bool kFma, int MODE>
__global__ void __launch_bounds__(32 * kWarps, UME_WARPK_MINB) moments_warp_kernel(Params p) {
    constexpr int C = 4 * LPR;
    constexpr int RPW = 32 / LPR;                 // feature rows per warp instruction
    constexpr int U = (UME_WARPK_UNROLL * RPW <= 32) ? UME_WARPK_UNROLL : (32 / RPW);
    static_assert(U >= 1 && 2 * RPW * U + 32 <= kRing && kRing % (RPW * U) == 0,
                  "a batch must fit the ring next to one chunk of hits");
    __shared__ WarpSmem smem[kWarps];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpSmem& sm = smem[warp];
    const unsigned lt = lanemask_lt();
    const int sub = lane / LPR, l = lane % LPR;

  for (;;) {                                      // nothing in here synchronises the CTA
#if UME_WARPK_PERSISTENT
    unsigned long long q = 0;
    if (lane == 0) q = atomicAdd(p.next, 1ull);
    q = __shfl_sync(UME_FULL_MASK, q, 0);
    if (q >= (unsigned long long)p.total) break;
    __syncwarp();
#else
    const unsigned long long q = (unsigned long long)blockIdx.x * kWarps + warp;
    if (q >= (unsigned long long)p.total) break;
#endif

    const int b = (int)(q / p.n);
    const int N = p.grid.N;
    const GridHeader h = p.grid.hdr[b];
    const int* cs = p.grid.cell_start + (size_t)b * (p.grid.cells_cap + 1);
    const float4* sorted_b = p.grid.sorted + (size_t)b * N;
    const float* feat_b = (b < p.Bs) ? p.feat + (size_t)b * N * C : p.feat2 + (size_t)(b - p.Bs) * N * C;
    const float* kp = (b < p.Bs) ? p.kpts + q * 3 : p.kpts2 + (q - (unsigned long long)p.Bs * p.n) * 3;
    const float kx = kp[0], ky = kp[1], kz = kp[2];
    const float r2 = __fmul_rn(p.radius, p.radius);
    const int K = p.K;

    // ---- candidate runs and the chunk table
    const RowSetup rs = row_setup(h, kx, ky, kz, p.radius);
    const int nrows = rs.nrows;
    int nchunks = 0;
    for (int r0 = 0; r0 < nrows; r0 += 32) {      // one or two rounds
        const int row = r0 + lane;
        int s = 0, n = 0;
        if (row < nrows) row_run(rs, h, cs, kx, ky, kz, row, s, n);
        const int nch = (n + 31) >> 5;
        int incl = nch;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(UME_FULL_MASK, incl, o);
            if (lane >= o) incl += v;
        }
        if (row < nrows) {
            sm.seg_start[row] = s;
            sm.seg_n[row] = n;
            sm.seg_c0[row] = nchunks + incl - nch;
        }
        nchunks += __shfl_sync(UME_FULL_MASK, incl, 31);
    }
#pragma unroll
    for (int i = lane; i < kBins; i += 32) sm.hist[i] = 0;
    __syncwarp();
    int loaded_w0 = -1;

    // ---- pass 1: hits and the level-0 histogram of their row indices
    int shift = max(0, 24 - __clz(N - 1));        // smallest shift with (N-1) >> shift < kBins
    int my_hits = 0;
    const unsigned hist_a = smem_u32(sm.hist);
    scan<kFma, UME_WARPK_D1>(sm, nrows, nchunks, loaded_w0, sorted_b, kx, ky, kz, r2,
                  [&](bool hit, float, float, float, int row) {
                      my_hits += hit ? 1 : 0;
                      inc_if(hit, hist_a + 4u * (unsigned)(row >> shift));
                  });
    const int hits = __reduce_add_sync(UME_FULL_MASK, my_hits);
    __syncwarp();

    // ---- the bin [T_lo, T_hi) of the K-th smallest row index, `need` of its hits are neighbours
    int T_lo = 0x7fffffff, T_hi = 0x7fffffff, need = 0;
    const bool saturated = hits > K;
    if (saturated) {
        int below = 0, lo = 0;
        for (;;) {
            // lane owns bins [8 lane, 8 lane + 8)
            const uint4 h0 = *reinterpret_cast<const uint4*>(&sm.hist[8 * lane]);
            const uint4 h1 = *reinterpret_cast<const uint4*>(&sm.hist[8 * lane + 4]);
            const int hv[8] = {(int)h0.x, (int)h0.y, (int)h0.z, (int)h0.w, (int)h1.x, (int)h1.y, (int)h1.z, (int)h1.w};
            int s8 = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) s8 += hv[i];
            int incl = s8;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(UME_FULL_MASK, incl, o);
                if (lane >= o) incl += v;
            }
            const int excl = below + incl - s8;
            const bool mine = (excl < K) && (K <= excl + s8);       // exactly one lane
            int bin = 0, bin_below = 0, bin_cnt = 0;
            if (mine) {
                int run = excl;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (run < K && K <= run + hv[i]) { bin = 8 * lane + i; bin_below = run; bin_cnt = hv[i]; }
                    run += hv[i];
                }
            }
            const int src = __ffs(__ballot_sync(UME_FULL_MASK, mine)) - 1;
            bin = __shfl_sync(UME_FULL_MASK, bin, src);
            bin_below = __shfl_sync(UME_FULL_MASK, bin_below, src);
            bin_cnt = __shfl_sync(UME_FULL_MASK, bin_cnt, src);
            T_lo = lo + (bin << shift);
            T_hi = T_lo + (1 << shift);
            need = K - bin_below;
            if (bin_cnt <= kMaybe || shift == 0) break;
            // refine: histogram of the hits inside [T_lo, T_hi) with 256x narrower bins
            below = bin_below;
            lo = T_lo;
            const int hi = T_hi;
            shift = max(0, shift - 8);
            __syncwarp();
#pragma unroll
            for (int i = lane; i < kBins; i += 32) sm.hist[i] = 0;
            __syncwarp();
            scan<kFma, UME_WARPK_D1>(sm, nrows, nchunks, loaded_w0, sorted_b, kx, ky, kz, r2,
                          [&](bool hit, float, float, float, int row) {
                              const bool in = hit & (row >= lo) & (row < hi);
                              inc_if(in, hist_a + 4u * (unsigned)((in ? row - lo : 0) >> shift));
                          });
            __syncwarp();
        }
    }

    if (MODE == kCountOnly) {
        if (lane == 0 && p.count) p.count[q] = min(hits, K);
        UME_WARPK_NEXT;
    }

    // ---- pass 2: gather + moments (kBackward: scatter of the moment gradients)
    float2 a01[4], a23[4];                         // [moment 1,x,y,z] of channels (0,1) and (2,3) of this lane
#pragma unroll
    for (int j = 0; j < 4; ++j) { a01[j] = make_float2(0.f, 0.f); a23[j] = make_float2(0.f, 0.f); }
    const float* fl = feat_b + 4 * l;
    float* gl = nullptr;
    if (MODE == kBackward) {
        // a01[j] / a23[j] hold the gradient instead: [g0 + g1 . k, g1x, g1y, g1z] per channel, so that
        // d/df_j = a[0] + a[1] ex + a[2] ey + a[3] ez with e = x_j - k
        gl = p.grad_feat + (size_t)b * N * C + 4 * l;
        const float4* g = reinterpret_cast<const float4*>(p.gF) + (size_t)q * C + 4 * l;
        const float4 g0 = __ldg(g + 0), g1 = __ldg(g + 1), g2 = __ldg(g + 2), g3 = __ldg(g + 3);
        auto c0 = [&](const float4& v) { return fmaf(v.w, kz, fmaf(v.z, ky, fmaf(v.y, kx, v.x))); };
        a01[0] = make_float2(c0(g0), c0(g1)); a23[0] = make_float2(c0(g2), c0(g3));
        a01[1] = make_float2(g0.y, g1.y);     a23[1] = make_float2(g2.y, g3.y);
        a01[2] = make_float2(g0.z, g1.z);     a23[2] = make_float2(g2.z, g3.z);
        a01[3] = make_float2(g0.w, g1.w);     a23[3] = make_float2(g2.w, g3.w);
    }
    auto scatter = [&](const float4& nb) {
        const float2 nx = make_float2(nb.x, nb.x), ny = make_float2(nb.y, nb.y), nz = make_float2(nb.z, nb.z);
        const float2 v01 = __ffma2_rn(a01[3], nz, __ffma2_rn(a01[2], ny, __ffma2_rn(a01[1], nx, a01[0])));
        const float2 v23 = __ffma2_rn(a23[3], nz, __ffma2_rn(a23[2], ny, __ffma2_rn(a23[1], nx, a23[0])));
        red_add_f4(gl + (size_t)__float_as_int(nb.w) * C, v01.x, v01.y, v23.x, v23.y);
    };
    auto accumulate = [&](const float4& nb, const float4& f) {
        const float2 f01 = make_float2(f.x, f.y), f23 = make_float2(f.z, f.w);
        const float2 nx = make_float2(nb.x, nb.x), ny = make_float2(nb.y, nb.y), nz = make_float2(nb.z, nb.z);
        // Blackwell's two-wide fp32 pipe: 2 FADD2 + 6 FFMA2 for the 16 scalar updates
        a01[0] = __fadd2_rn(a01[0], f01);     a23[0] = __fadd2_rn(a23[0], f23);
        a01[1] = __ffma2_rn(f01, nx, a01[1]); a23[1] = __ffma2_rn(f23, nx, a23[1]);
        a01[2] = __ffma2_rn(f01, ny, a01[2]); a23[2] = __ffma2_rn(f23, ny, a23[2]);
        a01[3] = __ffma2_rn(f01, nz, a01[3]); a23[3] = __ffma2_rn(f23, nz, a23[3]);
    };
    int wpos = 0, rpos = 0;                        // ring cursors (warp-uniform, free running)
    const unsigned ring_a = smem_u32(sm.ring), maybe_a = smem_u32(sm.maybe);
    const unsigned ring_sub = ring_a + 16u * (unsigned)sub;
    auto consume = [&]() {
        while (wpos - rpos >= RPW * U) {
            // a batch never wraps: RPW * U divides kRing.  Row indices first, all U feature loads in
            // flight, then the offsets are re-read from the ring as the rows arrive (registers)
            const unsigned ra = ring_sub + 16u * (unsigned)(rpos & (kRing - 1));
            if (MODE == kBackward) {
#pragma unroll
                for (int u = 0; u < U; ++u) scatter(lds_f4(ra + 16u * (u * RPW)));
            } else {
                float4 f[U];
#pragma unroll
                for (int u = 0; u < U; ++u) f[u] = ldg_f4(fl + (size_t)lds_u32(ra + 16u * (u * RPW) + 12u) * C);
#pragma unroll
                for (int u = 0; u < U; ++u) accumulate(lds_f4(ra + 16u * (u * RPW)), f[u]);
            }
            rpos += RPW * U;
        }
    };
    auto push = [&](bool take, float ex, float ey, float ez, int row) {
        const unsigned m = __ballot_sync(UME_FULL_MASK, take);
        if (m) {
            sts_f4_if(take, ring_a + 16u * (unsigned)((wpos + __popc(m & lt)) & (kRing - 1)), ex, ey, ez, __int_as_float(row));
            wpos += __popc(m);
            __syncwarp();
            consume();
        }
    };
    int n_maybe = 0;
    scan<kFma, UME_WARPK_D2>(sm, nrows, nchunks, loaded_w0, sorted_b, kx, ky, kz, r2,
                  [&](bool hit, float ex, float ey, float ez, int row) {
                      const bool take = hit & (row < T_lo);
                      if (saturated) {
                          const bool maybe = hit & (row >= T_lo) & (row < T_hi);
                          const unsigned mm = __ballot_sync(UME_FULL_MASK, maybe);
                          if (mm) {                // at most kMaybe of these per query, or exactly one (shift == 0)
                              const int at = n_maybe + __popc(mm & lt);
                              sts_f4_if(maybe & (at < kMaybe), maybe_a + 16u * (unsigned)min(at, kMaybe - 1), ex, ey, ez, __int_as_float(row));
                              n_maybe += __popc(mm);
                          }
                      }
                      push(take, ex, ey, ez, row);
                  });
    __syncwarp();
    if (n_maybe > 0) {
        // the `need` smallest row indices of the crossing bin
        float4 mine = make_float4(0.f, 0.f, 0.f, __int_as_float(0x7fffffff));
        if (lane < n_maybe) mine = lds_f4(maybe_a + 16u * (unsigned)lane);
        const int my_row = __float_as_int(mine.w);
        int rank = 0;
        for (int j = 0; j < n_maybe; ++j) rank += (__shfl_sync(UME_FULL_MASK, my_row, j) < my_row) ? 1 : 0;
        push(lane < n_maybe && rank < need, mine.x, mine.y, mine.z, my_row);
    }
    // what is left in the ring: less than one full batch, predicated
    for (; rpos < wpos; rpos += RPW) {
        const int e = rpos + sub;
        if (e < wpos) {
            const float4 nb = lds_f4(ring_a + 16u * (unsigned)(e & (kRing - 1)));
            if (MODE == kBackward) scatter(nb);
            else accumulate(nb, ldg_f4(fl + (size_t)__float_as_int(nb.w) * C));
        }
    }
    if (MODE == kBackward) UME_WARPK_NEXT;

    // ---- combine the RPW row groups, normalise, un-centre, store
    float acc[4][4];                               // [channel within lane][moment]
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[0][j] = a01[j].x; acc[1][j] = a01[j].y; acc[2][j] = a23[j].x; acc[3][j] = a23[j].y; }
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] += __shfl_xor_sync(UME_FULL_MASK, acc[i][j], o);
    float f0 = (acc[0][0] + acc[1][0]) + (acc[2][0] + acc[3][0]);
#pragma unroll
    for (int o = 1; o < LPR; o <<= 1) f0 += __shfl_xor_sync(UME_FULL_MASK, f0, o);
    const float den = p.raw ? 1.f : f0 + 1e-6f;    // evaluate.py:59
    if (lane < LPR) {
        float4* Fo = reinterpret_cast<float4*>(p.F) + (size_t)q * C + 4 * lane;
        float4* Fco = p.Fc ? reinterpret_cast<float4*>(p.Fc) + (size_t)q * C + 4 * lane : nullptr;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float m0 = acc[i][0];
            // un-centre: sum f x = sum f (x-k) + k sum f
            Fo[i] = make_float4(m0 / den, (acc[i][1] + kx * m0) / den, (acc[i][2] + ky * m0) / den, (acc[i][3] + kz * m0) / den);
            if (Fco) Fco[i] = make_float4(m0 / den, acc[i][1] / den, acc[i][2] / den, acc[i][3] / den);
        }
    }
    if (lane == 0 && p.count) p.count[q] = min(hits, K);
#if !UME_WARPK_PERSISTENT
    break;
#endif
  }
}

template <int LPR, int MODE>
int launch(const Params& p, bool fma, cudaStream_t stream) {
    auto kern = fma ? moments_warp_kernel<LPR, true, MODE> : moments_warp_kernel<LPR, false, MODE>;
    long long grid = (p.total + kWarps - 1) / kWarps;
#if UME_WARPK_PERSISTENT
    static int ctas_per_sm[2] = {0, 0}, sms = 0;   // per template instance; benign race (same values)
    if (!ctas_per_sm[fma]) {
        int dev = 0, n = 0, per = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, kern, 32 * kWarps, 0);
        UME_REQUIRE(e == cudaSuccess && n > 0 && per > 0, UME_ERR_CUDA, "moments_warp: occupancy query: %s", cudaGetErrorString(e));
        sms = n;
        ctas_per_sm[fma] = per;
    }
    if (grid > (long long)sms * ctas_per_sm[fma]) grid = (long long)sms * ctas_per_sm[fma];
    cudaError_t e = cudaMemsetAsync(p.next, 0, sizeof(unsigned long long), stream);
    UME_REQUIRE(e == cudaSuccess, UME_ERR_CUDA, "moments_warp: cudaMemsetAsync: %s", cudaGetErrorString(e));
#endif
    kern<<<(unsigned)grid, 32 * kWarps, 0, stream>>>(p);
    count_launch();
    return check_launch("moments_warp_kernel");
}

template <int MODE>
int launch_c(const Params& p, int C, bool fma, cudaStream_t stream) {
    switch (C) {
        case 16: return launch<4, MODE>(p, fma, stream);
        case 32: return launch<8, MODE>(p, fma, stream);
        case 64: return launch<16, MODE>(p, fma, stream);
        default: return launch<32, MODE>(p, fma, stream);
    }
}

}  // namespace warpk
}  // namespace ume
