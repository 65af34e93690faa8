// Fused radius-neighbourhood gather + UME moment build (replaces evaluate.py:50-60).
//
// One CTA per keypoint.  The neighbourhood is collected into shared memory as (point - keypoint,
// row index) entries (neighbors.cuh); the gather then streams the neighbours' feature rows from
// HBM/L2 with 128-bit loads — a C=32 row is exactly one 128-byte line, 8 lanes x float4 — and
// accumulates the C x 4 matrix [sum f | sum f (x-k)^T] in registers.  Partial sums are combined
// with warp shuffles and one shared-memory pass, un-centred (F1 = F1c + k F0), normalised as the
// reference does and written as one 16-byte row per channel.  The (B,n,K,C) gather tensor the
// reference materialises (evaluate.py:54-55) never exists.
#include "neighbors.cuh"
#include "moments_warp.cuh"

namespace ume {

namespace {

// Tunables (kernel-variant experiments build with -D overrides; the defaults are the measured best)
#ifndef UME_MOMENTS_NT
#define UME_MOMENTS_NT 256        // threads per keypoint CTA
#endif
#ifndef UME_MOMENTS_MINB
#define UME_MOMENTS_MINB 5        // CTAs per SM the register allocation is capped for
#endif
#ifndef UME_GATHER_UNROLL
#define UME_GATHER_UNROLL 2       // feature-row loads in flight per lane
#endif
constexpr int kNT = UME_MOMENTS_NT;
constexpr int kGU = UME_GATHER_UNROLL;
constexpr int kNW = kNT / 32;

struct MomentsParams {
    GridView grid;
    const float* kpts;    // (B,n,3)
    const float* feat;    // (B,N,C)
    float* F;             // (B,n,C,4)
    float* Fc;            // (B,n,C,4) or null
    int32_t* count;       // (B,n) or null
    const float* kpts2;   // clouds [Bs, B): keypoints / features of a second batch handled by the same launch
    const float* feat2;   //   (null and Bs = B for a single batch)
    int Bs;
    int n, C, K, cap;
    float radius;
};

// ---- accumulators: vectorised (C = 4*LPR, LPR lanes x float4 per feature row) ----------------
template <int LPR>
struct VecAcc {
    static constexpr int C = 4 * LPR;
    static constexpr int RPW = 32 / LPR;       // feature rows per warp instruction
    // accumulators as packed pairs for Blackwell's two-wide fp32 pipe (FADD2 / FFMA2):
    // a01[j] = moment j of channels (0,1), a23[j] = channels (2,3); moments j = 1, x, y, z
    float2 a01[4], a23[4];

    UME_DEVI void set_channels(int) {}
    UME_DEVI void clear() {
#pragma unroll
        for (int j = 0; j < 4; ++j) { a01[j] = make_float2(0.f, 0.f); a23[j] = make_float2(0.f, 0.f); }
    }
    static UME_DEVI void fma2(float2& acc, const float2& x, const float2& y) {
        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(reinterpret_cast<uint64_t&>(acc))
            : "l"(reinterpret_cast<const uint64_t&>(x)), "l"(reinterpret_cast<const uint64_t&>(y)));
    }
    static UME_DEVI void add2(float2& acc, const float2& x) {
        asm("add.rn.f32x2 %0, %0, %1;" : "+l"(reinterpret_cast<uint64_t&>(acc)) : "l"(reinterpret_cast<const uint64_t&>(x)));
    }
    // 2 FADD2 + 6 FFMA2 per feature float4 (16 scalar fp32 operations)
    UME_DEVI void add(const float4& nb, const float4& f) {
        const float2 f01 = make_float2(f.x, f.y), f23 = make_float2(f.z, f.w);
        const float2 nx = make_float2(nb.x, nb.x), ny = make_float2(nb.y, nb.y), nz = make_float2(nb.z, nb.z);
        add2(a01[0], f01); add2(a23[0], f23);
        fma2(a01[1], f01, nx); fma2(a23[1], f23, nx);
        fma2(a01[2], f01, ny); fma2(a23[2], f23, ny);
        fma2(a01[3], f01, nz); fma2(a23[3], f23, nz);
    }
    // Every warp owns a contiguous slice of the list.  If the list holds non-neighbours (row index
    // > T) the warp first squeezes them out of its slice in place (ballot + prefix, no barrier: the
    // write position never passes the read position), then streams the feature rows of what is
    // left: RPW rows per warp instruction, kGU instructions in flight per lane, no predication.
    UME_DEVI void gather(float4* list, int len, int T, const float* __restrict__ feat_b) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int per = (len + kNW - 1) / kNW;
        const int lo = min(len, warp * per);
        int cnt = min(len, lo + per) - lo;
        float4* seg = list + lo;
        if (T != 0x7fffffff) {
            const unsigned lt = lanemask_lt();
            int out = 0;
            for (int i = 0; i < cnt; i += 32) {
                const bool in = i + lane < cnt;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (in) v = seg[i + lane];          // (lanes past the end read nothing: slot 0 may be rewritten by now)
                const bool keep = in && __float_as_int(v.w) <= T;
                const unsigned m = __ballot_sync(UME_FULL_MASK, keep);
                __syncwarp();                  // every lane's read of this round is done before any lane overwrites a slot
                if (keep) seg[out + __popc(m & lt)] = v;
                out += __popc(m);
            }
            cnt = out;
            __syncwarp();
        }
        const int sub = lane / LPR, l = lane % LPR;
        const float* fl = feat_b + 4 * l;
        int e = sub;
        for (; e + (kGU - 1) * RPW < cnt; e += kGU * RPW) {
            float4 nb[kGU], f[kGU];
#pragma unroll
            for (int u = 0; u < kGU; ++u) nb[u] = seg[e + u * RPW];
#pragma unroll
            for (int u = 0; u < kGU; ++u) f[u] = ldg_f4(fl + (size_t)__float_as_int(nb[u].w) * C);
#pragma unroll
            for (int u = 0; u < kGU; ++u) add(nb[u], f[u]);
        }
        for (; e < cnt; e += RPW) {
            const float4 nb = seg[e];
            add(nb, ldg_f4(fl + (size_t)__float_as_int(nb.w) * C));
        }
    }
    // combine the RPW row groups of the warp, then lanes [0,LPR) hold the warp's C x 4 partial
    UME_DEVI void store(float4* red_w) {
        float a[4][4];                         // [channel within lane][moment]
#pragma unroll
        for (int j = 0; j < 4; ++j) { a[0][j] = a01[j].x; a[1][j] = a01[j].y; a[2][j] = a23[j].x; a[3][j] = a23[j].y; }
#pragma unroll
        for (int o = LPR; o < 32; o <<= 1)
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) a[i][j] += __shfl_xor_sync(UME_FULL_MASK, a[i][j], o);
        const int lane = threadIdx.x & 31;
        if (lane < LPR) {
#pragma unroll
            for (int i = 0; i < 4; ++i) red_w[4 * lane + i] = make_float4(a[i][0], a[i][1], a[i][2], a[i][3]);
        }
    }
};

// ---- accumulators: any C <= 32*CJ, one feature row per warp instruction, scalar loads --------
template <int CJ>
struct GenAcc {
    float a[CJ][4];
    int C;
    UME_DEVI void set_channels(int c) { C = c; }
    UME_DEVI void clear() {
#pragma unroll
        for (int i = 0; i < CJ; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) a[i][j] = 0.f;
    }
    UME_DEVI void add_row(const float4& nb, const float* __restrict__ row) {
        const int lane = threadIdx.x & 31;
#pragma unroll
        for (int i = 0; i < CJ; ++i) {
            const int c = lane + 32 * i;
            const float f = (c < C) ? __ldg(row + c) : 0.f;
            a[i][0] += f;
            a[i][1] = fmaf(f, nb.x, a[i][1]);
            a[i][2] = fmaf(f, nb.y, a[i][2]);
            a[i][3] = fmaf(f, nb.z, a[i][3]);
        }
    }
    UME_DEVI void gather(float4* list, int len, int T, const float* __restrict__ feat_b) {
        const int warp = threadIdx.x >> 5;
        int e = warp;
        for (; e + kNW < len; e += 2 * kNW) {
            const float4 n0 = list[e], n1 = list[e + kNW];
            if (__float_as_int(n0.w) <= T) add_row(n0, feat_b + (size_t)__float_as_int(n0.w) * C);   // warp-uniform
            if (__float_as_int(n1.w) <= T) add_row(n1, feat_b + (size_t)__float_as_int(n1.w) * C);
        }
        for (; e < len; e += kNW) {
            const float4 n0 = list[e];
            if (__float_as_int(n0.w) <= T) add_row(n0, feat_b + (size_t)__float_as_int(n0.w) * C);
        }
    }
    UME_DEVI void store(float4* red_w) {
        const int lane = threadIdx.x & 31;
#pragma unroll
        for (int i = 0; i < CJ; ++i) {
            const int c = lane + 32 * i;
            if (c < C) red_w[c] = make_float4(a[i][0], a[i][1], a[i][2], a[i][3]);
        }
    }
};

template <typename Acc, bool kFma>
__global__ void __launch_bounds__(kNT, UME_MOMENTS_MINB) moments_kernel(MomentsParams p) {
    extern __shared__ float4 list[];            // cap entries; reused as reduction scratch
    __shared__ CollectSmem sm;
    __shared__ float s_f0[256];

    const int q = blockIdx.x;
    const int b = q / p.n;
    const int C = p.C;
    const GridHeader h = p.grid.hdr[b];
    const int* cs = p.grid.cell_start + (size_t)b * (p.grid.cells_cap + 1);
    const float4* sorted_b = p.grid.sorted + (size_t)b * p.grid.N;
    const float* feat_b = (b < p.Bs) ? p.feat + (size_t)b * p.grid.N * C : p.feat2 + (size_t)(b - p.Bs) * p.grid.N * C;
    const float* kp = (b < p.Bs) ? p.kpts + (size_t)q * 3 : p.kpts2 + ((size_t)q - (size_t)p.Bs * p.n) * 3;
    const float kx = kp[0], ky = kp[1], kz = kp[2];

    Acc acc;
    acc.clear();
    acc.set_channels(C);

    const int used = collect_neighbors<kFma, kNT, false>(
        sm, list, p.cap, h, cs, sorted_b, p.grid.N, kx, ky, kz, p.radius, p.K, [&](int len, int T) {
            __syncthreads();                     // list entries of every warp are visible
            acc.gather(list, len, T, feat_b);
        });

    __syncthreads();                             // everyone is done reading the list
    const int warp = threadIdx.x >> 5;
    acc.store(list + (size_t)warp * C);
    __syncthreads();
    const int c = threadIdx.x;
    float4 row = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < C) {
#pragma unroll
        for (int w = 0; w < kNW; ++w) {
            const float4 v = list[(size_t)w * C + c];
            row.x += v.x; row.y += v.y; row.z += v.z; row.w += v.w;
        }
        s_f0[c] = row.x;
    }
    __syncthreads();
    if (c < C) {
        float s = 0.f;
        for (int i = 0; i < C; ++i) s += s_f0[i];          // same order in every thread
        const float den = s + 1e-6f;                        // evaluate.py:59
        const size_t o = ((size_t)q * C + c) * 4;
        float4 out;
        out.x = row.x / den;
        out.y = (row.y + kx * row.x) / den;                // un-centre: sum f x = sum f (x-k) + k sum f
        out.z = (row.z + ky * row.x) / den;
        out.w = (row.w + kz * row.x) / den;
        *reinterpret_cast<float4*>(p.F + o) = out;
        if (p.Fc) *reinterpret_cast<float4*>(p.Fc + o) = make_float4(row.x / den, row.y / den, row.z / den, row.w / den);
    }
    if (threadIdx.x == 0 && p.count) p.count[q] = used;
}

template <typename Acc>
int launch_moments(const MomentsParams& p, int B, bool fma, cudaStream_t stream) {
    const size_t smem = (size_t)p.cap * sizeof(float4);
    auto kern = fma ? moments_kernel<Acc, true> : moments_kernel<Acc, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    UME_REQUIRE(e == cudaSuccess, UME_ERR_CUDA, "moments: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    ProfScope prof(UME_PROF_MOMENTS, stream);
    kern<<<(unsigned)((size_t)B * p.n), kNT, smem, stream>>>(p);
    count_launch();
    return check_launch("moments_kernel");
}

}  // namespace

static constexpr long long kWarpKernelMinKeypoints = 3072;
#ifndef UME_WARPK_CELL_DIV
#define UME_WARPK_CELL_DIV 2       // search-grid cell = radius / this for the warp-per-keypoint kernel
#endif

static int moments_cap(int C, int K) {
    (void)K;
    int cap = 2048;
    while (cap < kNW * C) cap *= 2;              // reduction scratch must fit in the list
    return cap;
}

}  // namespace ume

extern "C" size_t ume_moments_workspace_bytes(int B, int N, int n, int C, int K) {
    (void)n; (void)C; (void)K;
    if (B <= 0 || N <= 0) return 0;
    return ume::grid_workspace_bytes(B, N, ume::kCellsCap) + 256;
}

extern "C" int ume_moments_f32(const float* pts, const float* kpts, const float* feat, int B, int N, int n,
                               int C, int K, float radius, unsigned flags, float* F, float* Fc,
                               int32_t* count, void* ws, size_t ws_bytes, void* stream_) {
    using namespace ume;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    UME_REQUIRE(B >= 0 && N >= 0 && n >= 0, UME_ERR_BAD_ARG, "ume_moments_f32: negative size");
    if (B == 0 || n == 0) return UME_OK;
    UME_REQUIRE(pts && kpts && feat && F, UME_ERR_BAD_ARG, "ume_moments_f32: null pointer");
    UME_REQUIRE(N >= 1, UME_ERR_BAD_ARG, "ume_moments_f32: empty cloud (N = 0)");
    UME_REQUIRE(C >= 1 && C <= 256, UME_ERR_UNSUPPORTED, "ume_moments_f32: C = %d not in [1,256]", C);
    UME_REQUIRE(K >= 1, UME_ERR_BAD_ARG, "ume_moments_f32: K = %d < 1", K);
    UME_REQUIRE(N <= kMaxPoints, UME_ERR_UNSUPPORTED, "ume_moments_f32: N = %d > %d", N, kMaxPoints);
    UME_REQUIRE((size_t)B * n < 0x7fffffffull, UME_ERR_UNSUPPORTED, "ume_moments_f32: B*n too large");
    UME_REQUIRE(ws && ws_bytes >= ume_moments_workspace_bytes(B, N, n, C, K), UME_ERR_WORKSPACE,
                "ume_moments_f32: workspace too small (%zu needed, %zu given)",
                ume_moments_workspace_bytes(B, N, n, C, K), ws_bytes);
    Workspace w(ws, ws_bytes);
    MomentsParams p;
    const bool aligned = reinterpret_cast<uintptr_t>(feat) % 16 == 0;
    const bool warp_ok = aligned && (C == 16 || C == 32 || C == 64 || C == 128);
    const bool raw = (flags & UME_FLAG_RAW_MOMENTS) != 0;
    UME_REQUIRE(!raw || warp_ok, UME_ERR_UNSUPPORTED, "ume_moments_f32: raw moments need C in {16,32,64,128} (C = %d)", C);
    // A warp takes ~80 us per keypoint (latency-bound, hidden by the 32 x 148 warps in flight); a 256-thread
    // CTA takes ~16 us.  Launches too small to fill the warp slots are faster on the CTA kernel.
    const bool small = (long long)B * n < kWarpKernelMinKeypoints;
    const bool use_warp = (!((flags & UME_FLAG_CTA_MOMENTS) || (small && !(flags & UME_FLAG_WARP_MOMENTS))) || raw) && warp_ok;
    // cells of radius/2 for the warp kernel (its two candidate scans pay for every candidate twice)
    const float cell = fabsf(radius) / (use_warp ? (float)UME_WARPK_CELL_DIV : ((flags & UME_FLAG_CELL_DIV2) ? 2.f : 1.f));
    int rc = grid_build(pts, kpts, B, N, n, fabsf(radius), cell, kCellsCap, w, &p.grid, stream);
    if (rc != UME_OK) return rc;
    p.kpts = kpts; p.feat = feat; p.F = F; p.Fc = Fc; p.count = count;
    p.kpts2 = nullptr; p.feat2 = nullptr; p.Bs = B;
    p.n = n; p.C = C; p.K = K; p.cap = moments_cap(C, K); p.radius = radius;
    const bool fma = (flags & UME_FLAG_FMA_DIST) != 0;
    if (use_warp) {
        // one warp per keypoint (moments_warp.cuh): the default for the channel counts it is built for
        warpk::Params wp;
        wp.grid = p.grid; wp.kpts = kpts; wp.feat = feat; wp.F = F; wp.Fc = Fc; wp.count = count;
        wp.kpts2 = nullptr; wp.feat2 = nullptr; wp.Bs = B;
        wp.gF = nullptr; wp.grad_feat = nullptr; wp.raw = raw ? 1 : 0;
        wp.n = n; wp.K = K; wp.total = (long long)B * n; wp.radius = radius;
        wp.next = w.take<unsigned long long>(1);
        UME_REQUIRE(w.ok(), UME_ERR_WORKSPACE, "ume_moments_f32: workspace too small for the work counter");
        ProfScope prof(UME_PROF_MOMENTS, stream);
        return warpk::launch_c<warpk::kForward>(wp, C, fma, stream);
    }
    const bool vec = (C % 4 == 0) && ((C & (C - 1)) == 0) && C >= 4 && C <= 128 &&
                     (reinterpret_cast<uintptr_t>(feat) % 16 == 0);
    if (vec) {
        switch (C) {
            case 4: return launch_moments<VecAcc<1>>(p, B, fma, stream);
            case 8: return launch_moments<VecAcc<2>>(p, B, fma, stream);
            case 16: return launch_moments<VecAcc<4>>(p, B, fma, stream);
            case 32: return launch_moments<VecAcc<8>>(p, B, fma, stream);
            case 64: return launch_moments<VecAcc<16>>(p, B, fma, stream);
            default: return launch_moments<VecAcc<32>>(p, B, fma, stream);
        }
    }
    if (C <= 32) return launch_moments<GenAcc<1>>(p, B, fma, stream);
    if (C <= 64) return launch_moments<GenAcc<2>>(p, B, fma, stream);
    if (C <= 128) return launch_moments<GenAcc<4>>(p, B, fma, stream);
    return launch_moments<GenAcc<8>>(p, B, fma, stream);
}

// Source and target batch of a registration step in ONE grid build and ONE moment launch: 2B clouds, no tail
// between the two sides, half the launches.  Kernel choice as in ume_moments_f32.
extern "C" int ume_moments_pair_f32(const float* pts1, const float* kpts1, const float* feat1, const float* pts2,
                                    const float* kpts2, const float* feat2, int B, int N, int n, int C, int K, float radius,
                                    unsigned flags, float* F, float* Fc, int32_t* count, void* ws, size_t ws_bytes,
                                    void* stream_) {
    using namespace ume;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    UME_REQUIRE(B >= 0 && N >= 0 && n >= 0, UME_ERR_BAD_ARG, "ume_moments_pair_f32: negative size");
    if (B == 0 || n == 0) return UME_OK;
    UME_REQUIRE(pts1 && kpts1 && feat1 && pts2 && kpts2 && feat2 && F, UME_ERR_BAD_ARG, "ume_moments_pair_f32: null pointer");
    UME_REQUIRE(N >= 1 && K >= 1, UME_ERR_BAD_ARG, "ume_moments_pair_f32: N = %d, K = %d", N, K);
    UME_REQUIRE(N <= kMaxPoints && (size_t)2 * B * n < 0x7fffffffull && 2 * B <= 65535, UME_ERR_UNSUPPORTED,
                "ume_moments_pair_f32: size not supported");
    UME_REQUIRE((C == 16 || C == 32 || C == 64 || C == 128) && reinterpret_cast<uintptr_t>(feat1) % 16 == 0 &&
                    reinterpret_cast<uintptr_t>(feat2) % 16 == 0,
                UME_ERR_UNSUPPORTED, "ume_moments_pair_f32: C = %d must be 16, 32, 64 or 128 (16-byte aligned rows)", C);
    UME_REQUIRE(ws && ws_bytes >= ume_moments_workspace_bytes(2 * B, N, n, C, K), UME_ERR_WORKSPACE,
                "ume_moments_pair_f32: workspace too small (%zu needed, %zu given)", ume_moments_workspace_bytes(2 * B, N, n, C, K),
                ws_bytes);
    Workspace w(ws, ws_bytes);
    const bool raw = (flags & UME_FLAG_RAW_MOMENTS) != 0;
    const bool fma = (flags & UME_FLAG_FMA_DIST) != 0;
    // launches too small to fill the warp slots are faster on the CTA-per-keypoint kernel (see ume_moments_f32)
    const bool small = (long long)2 * B * n < kWarpKernelMinKeypoints;
    const bool use_warp = !((flags & UME_FLAG_CTA_MOMENTS) || (small && !(flags & UME_FLAG_WARP_MOMENTS))) || raw;
    if (!use_warp) {
        MomentsParams p;
        const float cell = fabsf(radius) / ((flags & UME_FLAG_CELL_DIV2) ? 2.f : 1.f);
        int rc = grid_build(pts1, kpts1, B, N, n, fabsf(radius), cell, kCellsCap, w, &p.grid, stream, pts2, kpts2, B);
        if (rc != UME_OK) return rc;
        p.kpts = kpts1; p.feat = feat1; p.kpts2 = kpts2; p.feat2 = feat2; p.Bs = B;
        p.F = F; p.Fc = Fc; p.count = count;
        p.n = n; p.C = C; p.K = K; p.cap = moments_cap(C, K); p.radius = radius;
        switch (C) {
            case 16: return launch_moments<VecAcc<4>>(p, 2 * B, fma, stream);
            case 32: return launch_moments<VecAcc<8>>(p, 2 * B, fma, stream);
            case 64: return launch_moments<VecAcc<16>>(p, 2 * B, fma, stream);
            default: return launch_moments<VecAcc<32>>(p, 2 * B, fma, stream);
        }
    }
    warpk::Params wp;
    int rc = grid_build(pts1, kpts1, B, N, n, fabsf(radius), fabsf(radius) / (float)UME_WARPK_CELL_DIV, kCellsCap, w, &wp.grid,
                        stream, pts2, kpts2, B);
    if (rc != UME_OK) return rc;
    wp.kpts = kpts1; wp.feat = feat1; wp.kpts2 = kpts2; wp.feat2 = feat2; wp.Bs = B;
    wp.F = F; wp.Fc = Fc; wp.count = count;
    wp.gF = nullptr; wp.grad_feat = nullptr; wp.raw = raw ? 1 : 0;
    wp.n = n; wp.K = K; wp.total = (long long)2 * B * n; wp.radius = radius;
    wp.next = w.take<unsigned long long>(1);
    UME_REQUIRE(w.ok(), UME_ERR_WORKSPACE, "ume_moments_pair_f32: workspace too small for the work counter");
    ProfScope prof(UME_PROF_MOMENTS, stream);
    return warpk::launch_c<warpk::kForward>(wp, C, fma, stream);
}

// Shared front end of the two auxiliary entry points below (same checks and grid as ume_moments_f32).
static int moments_aux(const float* pts, const float* kpts, const float* gF, int B, int N, int n, int C, int K,
                       float radius, unsigned flags, float* grad_feat, int32_t* count, void* ws, size_t ws_bytes,
                       void* stream_, bool backward, const char* who) {
    using namespace ume;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    UME_REQUIRE(B >= 0 && N >= 0 && n >= 0, UME_ERR_BAD_ARG, "%s: negative size", who);
    if (B == 0 || n == 0) return UME_OK;
    UME_REQUIRE(pts && kpts, UME_ERR_BAD_ARG, "%s: null pointer", who);
    UME_REQUIRE(N >= 1 && K >= 1, UME_ERR_BAD_ARG, "%s: N = %d, K = %d", who, N, K);
    UME_REQUIRE(N <= kMaxPoints && (size_t)B * n < 0x7fffffffull, UME_ERR_UNSUPPORTED, "%s: size not supported", who);
    UME_REQUIRE(ws && ws_bytes >= ume_moments_workspace_bytes(B, N, n, C, K), UME_ERR_WORKSPACE, "%s: workspace too small", who);
    Workspace w(ws, ws_bytes);
    warpk::Params wp;
    int rc = grid_build(pts, kpts, B, N, n, fabsf(radius), fabsf(radius) / (float)UME_WARPK_CELL_DIV, kCellsCap, w, &wp.grid, stream);
    if (rc != UME_OK) return rc;
    wp.kpts = kpts; wp.feat = nullptr; wp.F = nullptr; wp.Fc = nullptr; wp.count = count;
    wp.kpts2 = nullptr; wp.feat2 = nullptr; wp.Bs = B;
    wp.gF = gF; wp.grad_feat = grad_feat; wp.raw = 1;
    wp.n = n; wp.K = K; wp.total = (long long)B * n; wp.radius = radius;
    wp.next = w.take<unsigned long long>(1);
    UME_REQUIRE(w.ok(), UME_ERR_WORKSPACE, "%s: workspace too small for the work counter", who);
    const bool fma = (flags & UME_FLAG_FMA_DIST) != 0;
    ProfScope prof(UME_PROF_MOMENTS, stream);
    return backward ? warpk::launch_c<warpk::kBackward>(wp, C, fma, stream)
                    : warpk::launch_c<warpk::kCountOnly>(wp, 32, fma, stream);
}

extern "C" int ume_moments_backward_f32(const float* pts, const float* kpts, const float* gF, int B, int N, int n, int C,
                                        int K, float radius, unsigned flags, float* grad_feat, void* ws, size_t ws_bytes,
                                        void* stream) {
    UME_REQUIRE(gF && grad_feat, UME_ERR_BAD_ARG, "ume_moments_backward_f32: null pointer");
    UME_REQUIRE((C == 16 || C == 32 || C == 64 || C == 128) && reinterpret_cast<uintptr_t>(grad_feat) % 16 == 0 &&
                    reinterpret_cast<uintptr_t>(gF) % 16 == 0,
                UME_ERR_UNSUPPORTED, "ume_moments_backward_f32: C = %d must be 16, 32, 64 or 128 (16-byte aligned rows)", C);
    return moments_aux(pts, kpts, gF, B, N, n, C, K, radius, flags, grad_feat, nullptr, ws, ws_bytes, stream, true,
                       "ume_moments_backward_f32");
}

extern "C" int ume_neighbor_count_f32(const float* pts, const float* kpts, int B, int N, int n, int K, float radius,
                                      unsigned flags, int32_t* count, void* ws, size_t ws_bytes, void* stream) {
    UME_REQUIRE(count, UME_ERR_BAD_ARG, "ume_neighbor_count_f32: null pointer");
    return moments_aux(pts, kpts, nullptr, B, N, n, 32, K, radius, flags, nullptr, count, ws, ws_bytes, stream, false,
                       "ume_neighbor_count_f32");
}
