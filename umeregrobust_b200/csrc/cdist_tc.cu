// All-pairs subspace distance on the 5th-gen tensor cores (impl 1): S = A B^T with
// A = (4 n1) x C, B = (4 n2) x C (descriptor basis vectors as rows), D = sqrt(4 - sum of squares of
// each 4x4 block of S), row arg-min fused.  Replaces utils/loc_utils.py:12-13 + evaluate.py:224.
//
// Precision: one reduced-precision product alone (TF32: 10 mantissa bits) gives ~7e-4 error in D
// (BASELINE.md §3).  Every operand is therefore split a * 2^8 = hi + lo into two FP16 numbers
// (hi = fp16(256 a), lo = fp16(256 a - hi): 22 mantissa bits together; the scale keeps the small
// entries of an orthonormal basis out of the fp16 subnormals) and each K step issues three MMAs,
// hi*hi + hi*lo + lo*hi, into the same fp32 TMEM accumulator — fp16 x fp16 products are exact in fp32,
// the dropped lo*lo term is 2^-24 relative: fp32-grade, and kind::f16 runs at twice the kind::tf32
// rate (round 1 used a 3 x TF32 split: 12 MMAs of K = 8 per tile, now 6 of K = 16).
// The operands arrive pre-split as rows of [hi (C) | lo (C)] halves: C = 32 makes a row exactly one
// 128-byte swizzle atom.  `ume_orthonormalize_split_f32` writes them directly; the generic
// `ume_cdist_f32` entry splits fp32 descriptors in a small pre-pass.
//
// Kernel shape (one CTA per SM, 320 threads):
//   warp 0      TMA producer: the CTA's A tiles once, then the B tiles of every n-block through a
//               ring of shared-memory stages (128-byte swizzle, mbarrier complete_tx)
//   warp 1      TMEM allocation + single-thread tcgen05.mma issue; tcgen05.commit releases the smem
//               stage and publishes the accumulator stage
//   warps 2..9  epilogue, two warps per TMEM lane quadrant (one per 128-row A block): tcgen05.ld (one
//               TMEM lane = one row of S per thread; all 128 columns in flight before the wait), squares, 4x4 block
//               sums (columns in-thread, rows with a reduce-scatter over 4 lanes), sqrt, 128-byte
//               coalesced D stores, running row arg-min; double-buffered against the next MMAs
// K = C is tiny (6 MMAs per 128 x 128 tile at C = 32), so the kernel lives or dies by how fast the
// accumulators are drained: with one epilogue warp per quadrant and a wait after every tcgen05.ld
// the drain took ~3000 cycles per tile against ~1100 cycles of MMA (round 1 / first fp16 version).
#include "ume_common.cuh"

#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>

namespace ume {
namespace {

constexpr int kTileM = 128;            // rows of S per MMA (= TMEM lanes) = 32 source keypoints
constexpr int kTileN = 128;            // columns of S per MMA                = 32 target keypoints
constexpr int kChunkK = 64;            // halves per 128-byte swizzle row
constexpr int kChunkBytes = kTileM * kChunkK * 2;   // 16 KB: one (128 rows x 64 halves) box
constexpr float kOperandScale = 256.f;               // operands are 2^8 a: accumulators hold 2^16 <a,b>
constexpr float kInvScale4 = 1.f / 4294967296.f;     // 2^-32: un-scales a sum of squared accumulators
constexpr int kEpiWarps = 8;             // two per TMEM lane quadrant: their tcgen05.ld latencies overlap
constexpr int kThreads = 64 + 32 * kEpiWarps;

// ---------------------------------------------------------------- PTX wrappers
UME_DEVI uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

UME_DEVI void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
UME_DEVI void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
UME_DEVI void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
UME_DEVI void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    }
}
UME_DEVI void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
UME_DEVI void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
UME_DEVI void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
UME_DEVI void tcgen05_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
UME_DEVI void tcgen05_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// issue only: the registers are valid after tmem_ld_wait()
UME_DEVI void tmem_ld_32x32b_x32_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
UME_DEVI void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

UME_DEVI void tmem_ld_32x32b_x32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, 128-byte swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// start address >> 4 in [0,14), LBO = 1 in [16,30), SBO = 1024 B >> 4 in [32,46), version 1 in
// [46,48), layout type SWIZZLE_128B (= 2) in [61,64).
UME_DEVI uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3fffu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor: D = F32 (1 << 4), A = B = F16 (format 0 in [7,10) and [10,13)), both
// K-major, N >> 3 in [17,23), M >> 4 in [24,29).  One instruction covers K = 16 halves = 32 bytes.
constexpr uint32_t kInstrDesc = (1u << 4) | (0u << 7) | (0u << 10) | ((kTileN >> 3) << 17) | ((kTileM >> 4) << 24);

// ---------------------------------------------------------------- pre-pass: a -> [hi | lo] halves
// in: rows x C fp32, out: rows x 2C halves (hi in columns [0,C), lo in [C,2C)), both of 256 a.
__global__ void split_f16_kernel(const float* __restrict__ in, __half* __restrict__ out, int64_t rows, int C) {
    const int c4n = C >> 2;
    const int64_t total = rows * c4n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / c4n;
        const int c4 = (int)(i % c4n);
        const float4 a = ldg_f4(in + r * C + c4 * 4);
        const float v[4] = {a.x * kOperandScale, a.y * kOperandScale, a.z * kOperandScale, a.w * kOperandScale};
        __half hi[4], lo[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            hi[k] = __float2half_rn(v[k]);
            lo[k] = __float2half_rn(v[k] - __half2float(hi[k]));
        }
        *reinterpret_cast<uint2*>(out + r * 2 * C + c4 * 4) = *reinterpret_cast<const uint2*>(hi);
        *reinterpret_cast<uint2*>(out + r * 2 * C + C + c4 * 4) = *reinterpret_cast<const uint2*>(lo);
    }
}

// ---------------------------------------------------------------- main kernel
struct TcParams {
    int n1, n2;
    float* D;
    int64_t* argmin;
    float* dmin;
};

// KC = C / 32 swizzle chunks (of 64 halves) per operand row [hi | lo]; MB = 128-row blocks of A per CTA;
// S = B stages.
template <int KC, int MB, int S>
struct TcSmem {
    static constexpr int kABytes = MB * KC * kChunkBytes;
    static constexpr int kBStageBytes = KC * kChunkBytes;
    static constexpr int kTotal = kABytes + S * kBStageBytes + 1024;   // + alignment slack
};

template <int KC, int MB, int S>
__global__ void __launch_bounds__(kThreads, 1)
cdist_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smA = smem;                                           // [MB][KC] chunks
    uint8_t* smB = smem + TcSmem<KC, MB, S>::kABytes;              // [S][KC] chunks
    __shared__ uint64_t bar_a_full, bar_full[S], bar_empty[S], bar_tmem_full[2], bar_tmem_empty[2];
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const int m0 = blockIdx.x * MB * kTileM;                       // first row of A handled by this CTA (within batch b)
    const int rowsA = 4 * p.n1, rowsB = 4 * p.n2;
    const int NT = (rowsB + kTileN - 1) / kTileN;
    constexpr int kAccCols = MB * kTileN;                          // TMEM columns per accumulator stage

    if (threadIdx.x == 0) {
        mbar_init(&bar_a_full, 1);
        for (int s = 0; s < S; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&bar_tmem_full[a], 1); mbar_init(&bar_tmem_empty[a], kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapB)) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(2 * kAccCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            mbar_expect_tx(&bar_a_full, TcSmem<KC, MB, S>::kABytes);
            for (int mb = 0; mb < MB; ++mb)
                for (int c = 0; c < KC; ++c)
                    tma_load_2d(smA + (mb * KC + c) * kChunkBytes, &mapA, &bar_a_full, c * kChunkK,
                                b * rowsA + m0 + mb * kTileM);
            for (int t = 0; t < NT; ++t) {
                const int s = t % S, round = t / S;
                mbar_wait(&bar_empty[s], (round & 1) ^ 1);
                mbar_expect_tx(&bar_full[s], TcSmem<KC, MB, S>::kBStageBytes);
                for (int c = 0; c < KC; ++c)
                    tma_load_2d(smB + (s * KC + c) * kChunkBytes, &mapB, &bar_full[s], c * kChunkK, b * rowsB + t * kTileN);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            mbar_wait(&bar_a_full, 0);
            for (int t = 0; t < NT; ++t) {
                const int s = t % S, round = t / S;
                const int acc = t & 1, round2 = t >> 1;
                mbar_wait(&bar_tmem_empty[acc], (round2 & 1) ^ 1);
                mbar_wait(&bar_full[s], round & 1);
                tcgen05_fence_after();
                const uint32_t bBase = smem_u32(smB + s * KC * kChunkBytes);
                // K step j of a row (16 halves = 32 bytes): chunk j / 4, byte offset 32 (j % 4) inside the swizzle
                // atom; hi = steps [0, 2 KC), lo = steps [2 KC, 4 KC)
                constexpr int kHalfSteps = 2 * KC;
#pragma unroll
                for (int mb = 0; mb < MB; ++mb) {
                    const uint32_t aBase = smem_u32(smA + mb * KC * kChunkBytes);
                    const uint32_t d = tmem_base + acc * kAccCols + mb * kTileN;
                    uint32_t accum = 0;
                    // (a half, b half): hi*hi, hi*lo, lo*hi
#pragma unroll
                    for (int pr = 0; pr < 3; ++pr) {
                        const int ah = (pr == 2) ? 1 : 0, bh = (pr == 1) ? 1 : 0;
#pragma unroll
                        for (int kk = 0; kk < kHalfSteps; ++kk) {
                            const int ja = ah * kHalfSteps + kk, jb = bh * kHalfSteps + kk;
                            const uint64_t ad = umma_desc_sw128(aBase + (ja >> 2) * kChunkBytes + (ja & 3) * 32);
                            const uint64_t bd = umma_desc_sw128(bBase + (jb >> 2) * kChunkBytes + (jb & 3) * 32);
                            tcgen05_mma_f16(d, ad, bd, kInstrDesc, accum);
                            accum = 1;
                        }
                    }
                }
                tcgen05_commit(&bar_empty[s]);          // smem stage reusable once these MMAs have read it
                tcgen05_commit(&bar_tmem_full[acc]);    // accumulators complete
            }
        }
    } else {
        // ===================== epilogue (warps 2..9) =====================
        // Two warps per TMEM lane quadrant (a warp may only read the quadrant warp % 4): group 0 drains
        // the accumulators of A block 0, group 1 those of A block 1, so every scheduler has two
        // epilogue warps whose tcgen05.ld latencies overlap; a warp puts all four 32-column loads of its rows
        // in flight before it waits.
        static_assert(MB == 2, "one A block per epilogue-warp group");
        const int quad = warp & 3;                      // TMEM lane quadrant this warp may read
        const int mb = (warp - 2) >> 2;                 // the A block this warp finishes
        const int row = quad * 32 + lane;               // row of the 128-row block = TMEM lane
        const int r4 = lane & 3;
        const int hb = (r4 >> 1) & 1, lb = r4 & 1;
        const int jsub = 16 * hb + 8 * lb;              // the 8 target keypoints (of 32) this lane finishes
        float best = INFINITY;
        int best_j = 0x7fffffff;
        const bool vec_ok = (p.n2 & 3) == 0;
        const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16) + mb * kTileN;
        const int i = (m0 + mb * kTileM + row) >> 2;    // source keypoint of this lane

        for (int t = 0; t < NT; ++t) {
            const int acc = t & 1, round2 = t >> 1;
            mbar_wait(&bar_tmem_full[acc], round2 & 1);
            tcgen05_fence_after();
            const uint32_t col0 = lane_base + acc * kAccCols;
            float part[32];                              // per target keypoint of the tile: sum over its 4 columns
            // all 128 columns of this warp's rows in flight at once, one wait: the TMEM latency is paid once
            // per tile instead of once per 32 columns
            uint32_t v[4][32];
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) tmem_ld_32x32b_x32_issue(col0 + ch * 32, v[ch]);
            tmem_ld_wait();
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    const float v0 = __uint_as_float(v[ch][4 * g]), v1 = __uint_as_float(v[ch][4 * g + 1]);
                    const float v2 = __uint_as_float(v[ch][4 * g + 2]), v3 = __uint_as_float(v[ch][4 * g + 3]);
                    float sq = v0 * v0;
                    sq = fmaf(v1, v1, sq);
                    sq = fmaf(v2, v2, sq);
                    sq = fmaf(v3, v3, sq);
                    part[ch * 8 + g] = sq;
                }
            }
            // every accumulator this warp owns is in registers: hand the stage back to the MMA warp
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_tmem_empty[acc]);
            // reduce-scatter over the 4 lanes (= 4 basis rows of one source keypoint)
            float q16[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const float send = hb ? part[k] : part[k + 16];
                const float keep = hb ? part[k + 16] : part[k];
                q16[k] = keep + __shfl_xor_sync(UME_FULL_MASK, send, 2);
            }
            float s8[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float send = lb ? q16[k] : q16[k + 8];
                const float keep = lb ? q16[k + 8] : q16[k];
                s8[k] = keep + __shfl_xor_sync(UME_FULL_MASK, send, 1);
            }
            const int jbase = t * (kTileN / 4) + jsub;           // first target keypoint of this lane's 8
            float d8[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) d8[k] = sqrtf(fmaxf(fmaf(-s8[k], kInvScale4, 4.f), 0.f));
            if (i < p.n1) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (jbase + k < p.n2 && d8[k] < best) { best = d8[k]; best_j = jbase + k; }
                }
                if (p.D) {
                    float* dst = p.D + ((size_t)b * p.n1 + i) * p.n2 + jbase;
                    if (vec_ok && jbase + 8 <= p.n2) {
                        *reinterpret_cast<float4*>(dst) = make_float4(d8[0], d8[1], d8[2], d8[3]);
                        *reinterpret_cast<float4*>(dst + 4) = make_float4(d8[4], d8[5], d8[6], d8[7]);
                    } else {
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            if (jbase + k < p.n2) dst[k] = d8[k];
                    }
                }
            }
        }
        if (p.argmin || p.dmin) {
            float bd = best;
            int bj = best_j;
#pragma unroll
            for (int o = 2; o > 0; o >>= 1) {
                const float od = __shfl_xor_sync(UME_FULL_MASK, bd, o);
                const int oj = __shfl_xor_sync(UME_FULL_MASK, bj, o);
                if (od < bd || (od == bd && oj < bj)) { bd = od; bj = oj; }
            }
            if (r4 == 0 && i < p.n1) {
                if (p.argmin) p.argmin[(size_t)b * p.n1 + i] = (bj == 0x7fffffff) ? 0 : bj;
                if (p.dmin) p.dmin[(size_t)b * p.n1 + i] = bd;
            }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * kAccCols) : "memory");
    }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// rows x (2C) half matrix, box = 64 halves x 128 rows, 128-byte swizzle
int make_map(CUtensorMap* map, const __half* base, uint64_t rows, int C) {
    EncodeTiledFn fn = encode_tiled_fn();
    UME_REQUIRE(fn != nullptr, UME_ERR_CUDA, "cdist_tc: cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)(2 * C), (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)(2 * C) * sizeof(__half)};
    cuuint32_t box[2] = {(cuuint32_t)kChunkK, (cuuint32_t)kTileM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    UME_REQUIRE(r == CUDA_SUCCESS, UME_ERR_CUDA, "cdist_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return UME_OK;
}

template <int KC, int MB, int S>
int launch_tc(const CUtensorMap& mapA, const CUtensorMap& mapB, const TcParams& p, int B, cudaStream_t stream) {
    auto kern = cdist_tc_kernel<KC, MB, S>;
    constexpr int smem = TcSmem<KC, MB, S>::kTotal;
    static_assert(smem <= 227 * 1024, "shared memory budget");
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    UME_REQUIRE(e == cudaSuccess, UME_ERR_CUDA, "cdist_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    dim3 grid((unsigned)((4 * p.n1 + MB * kTileM - 1) / (MB * kTileM)), (unsigned)B);
    kern<<<grid, kThreads, smem, stream>>>(mapA, mapB, p);
    count_launch();
    return check_launch("cdist_tc_kernel");
}

}  // namespace

size_t cdist_tc_workspace_bytes(int B, int n1, int n2, int C) {
    if (B <= 0) return 0;
    return align_up((size_t)B * 4 * n1 * 2 * C * sizeof(__half), 256) + align_up((size_t)B * 4 * n2 * 2 * C * sizeof(__half), 256) + 256;
}

// Qh1 / Qh2: pre-split operands, rows of [hi (C) | lo (C)] halves of 256 q (see split_f16_kernel).
int cdist_tc_launch_split(const void* Qh1, const void* Qh2, int B, int n1, int n2, int C, float* D, int64_t* argmin,
                          float* dmin, cudaStream_t stream) {
    UME_REQUIRE(C == 32 || C == 64, UME_ERR_UNSUPPORTED, "cdist (tcgen05) supports C = 32 or 64, got %d", C);
    UME_REQUIRE((int64_t)B * 4 * (int64_t)(n1 > n2 ? n1 : n2) < 0x7fffffffll, UME_ERR_UNSUPPORTED, "cdist (tcgen05): too many rows");
    UME_REQUIRE(reinterpret_cast<uintptr_t>(Qh1) % 16 == 0 && reinterpret_cast<uintptr_t>(Qh2) % 16 == 0, UME_ERR_BAD_ARG,
                "cdist (tcgen05): operands not 16-byte aligned");
    const int64_t rows1 = (int64_t)B * 4 * n1, rows2 = (int64_t)B * 4 * n2;
    CUtensorMap mapA, mapB;
    int rc = make_map(&mapA, static_cast<const __half*>(Qh1), (uint64_t)rows1, C);
    if (rc != UME_OK) return rc;
    rc = make_map(&mapB, static_cast<const __half*>(Qh2), (uint64_t)rows2, C);
    if (rc != UME_OK) return rc;
    TcParams p;
    p.n1 = n1; p.n2 = n2; p.D = D; p.argmin = argmin; p.dmin = dmin;
    ProfScope prof(UME_PROF_CDIST, stream);
    if (C == 32) return launch_tc<1, 2, 4>(mapA, mapB, p, B, stream);
    return launch_tc<2, 2, 3>(mapA, mapB, p, B, stream);
}

int cdist_tc_launch(const float* Qt1, const float* Qt2, int B, int n1, int n2, int C, float* D, int64_t* argmin,
                    float* dmin, void* ws, size_t ws_bytes, cudaStream_t stream) {
    UME_REQUIRE(C == 32 || C == 64, UME_ERR_UNSUPPORTED, "ume_cdist_f32: impl 1 (tcgen05) supports C = 32 or 64, got %d", C);
    UME_REQUIRE(ws && ws_bytes >= cdist_tc_workspace_bytes(B, n1, n2, C), UME_ERR_WORKSPACE,
                "ume_cdist_f32: impl 1 workspace too small (%zu needed, %zu given)", cdist_tc_workspace_bytes(B, n1, n2, C), ws_bytes);
    Workspace w(ws, ws_bytes);
    const int64_t rows1 = (int64_t)B * 4 * n1, rows2 = (int64_t)B * 4 * n2;
    __half* A2 = w.take<__half>((size_t)rows1 * 2 * C);
    __half* B2 = w.take<__half>((size_t)rows2 * 2 * C);
    {
        ProfScope prof(UME_PROF_CDIST, stream);
        split_f16_kernel<<<(unsigned)std::min<int64_t>((rows1 * (C / 4) + 255) / 256, 148 * 16), 256, 0, stream>>>(Qt1, A2, rows1, C);
        split_f16_kernel<<<(unsigned)std::min<int64_t>((rows2 * (C / 4) + 255) / 256, 148 * 16), 256, 0, stream>>>(Qt2, B2, rows2, C);
        count_launch(2);
        int rc = check_launch("split_f16_kernel");
        if (rc != UME_OK) return rc;
    }
    return cdist_tc_launch_split(A2, B2, B, n1, n2, C, D, argmin, dmin, stream);
}

}  // namespace ume
