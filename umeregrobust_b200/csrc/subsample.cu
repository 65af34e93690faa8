// Distance-weighted match sub-sampling on the device: evaluate.py:233-245 draws
// `ume_n_samples` of the matches WITHOUT replacement with probability proportional to
// exp((1 - d) / tau) on the host (np.random.choice: a D2H copy and a sync per pair).  Successive
// sampling without replacement is distributed like "the k largest of log-weight + Gumbel noise", so
// here every match gets the key (1 - d) / tau - log(-log u), u ~ U(0,1) from a counter-based
// generator (Philox4x32-10 keyed by (seed, pair, match) — or from an array the caller passes, which
// is how the parity tests pin the selection), and one CTA per pair picks the k largest keys with an
// exact radix select in shared memory.  The survivors are written in ascending match order, so the
// result is a pure function of (d, tau, u).
#include "ume_common.cuh"

namespace ume {
namespace {

constexpr int kSelThreads = 512;

// Philox4x32-10 (Salmon et al., SC'11): counter = (i, b, 0, 0), key = seed
UME_DEVI uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

// order-preserving map float -> unsigned (larger float, larger key); NaN keys sort lowest
UME_DEVI unsigned ordered_key(float f) {
    if (f != f) return 0u;
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(kSelThreads)
gumbel_topk_kernel(const float* __restrict__ d, const float* __restrict__ u_in, int n, int k, float inv_tau,
                   unsigned long long seed, int64_t* __restrict__ idx_out) {
    extern __shared__ unsigned s_key[];                 // n ordered keys
    __shared__ unsigned s_hist[256];
    __shared__ unsigned s_prefix, s_want;
    __shared__ int s_warp[kSelThreads / 32];
    __shared__ int s_run_gt, s_run_eq;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* db = d + (size_t)b * n;
    for (int i = tid; i < n; i += kSelThreads) {
        float u;
        if (u_in) {
            u = u_in[(size_t)b * n + i];
        } else {
            const uint4 r = philox4x32_10(make_uint4((unsigned)i, (unsigned)b, 0u, 0u),
                                          make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
            u = ((float)(r.x >> 8) + 0.5f) * (1.0f / 16777216.0f);      // (0,1), 24 bits
        }
        u = fminf(fmaxf(u, 1e-20f), 1.0f - 1e-7f);
        const float key = (1.0f - db[i]) * inv_tau - logf(-logf(u));
        s_key[i] = ordered_key(key);
    }
    if (tid == 0) { s_prefix = 0u; s_want = (unsigned)k; }
    __syncthreads();
    // radix select of the k-th LARGEST key: 4 digits of 8 bits, most significant first
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int i = tid; i < 256; i += kSelThreads) s_hist[i] = 0u;
        __syncthreads();
        const unsigned prefix = s_prefix;
        const unsigned himask = (shift == 24) ? 0u : (0xffffffffu << (shift + 8));
        for (int i = tid; i < n; i += kSelThreads) {
            const unsigned key = s_key[i];
            if ((key & himask) == prefix) atomicAdd(&s_hist[(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            unsigned want = s_want, run = 0u;
            int dig = 255;
            for (; dig > 0; --dig) {
                if (run + s_hist[dig] >= want) break;
                run += s_hist[dig];
            }
            s_prefix = prefix | ((unsigned)dig << shift);
            s_want = want - run;                          // still wanted among the keys with this digit
        }
        __syncthreads();
    }
    const unsigned thr = s_prefix;                        // the k-th largest key
    const int want_eq = (int)s_want;                      // how many keys == thr belong to the selection
    if (tid == 0) { s_run_gt = 0; s_run_eq = 0; }
    __syncthreads();
    // survivors in ascending match order: every key > thr, and the first want_eq keys == thr
    int64_t* out = idx_out + (size_t)b * k;
    for (int base = 0; base < n; base += kSelThreads) {
        const int i = base + tid;
        const unsigned key = (i < n) ? s_key[i] : 0u;
        const bool gt = (i < n) && key > thr, eq = (i < n) && key == thr;
        const unsigned mg = __ballot_sync(UME_FULL_MASK, gt), me = __ballot_sync(UME_FULL_MASK, eq);
        if (lane == 0) s_warp[warp] = __popc(mg) | (__popc(me) << 16);
        __syncthreads();
        int g_before = 0, e_before = 0, g_tot = 0, e_tot = 0;
#pragma unroll
        for (int w = 0; w < kSelThreads / 32; ++w) {
            const int v = s_warp[w], g = v & 0xffff, e = v >> 16;
            if (w < warp) { g_before += g; e_before += e; }
            g_tot += g; e_tot += e;
        }
        const unsigned lt = lanemask_lt();
        const int my_g = s_run_gt + g_before + __popc(mg & lt);     // keys > thr before this one
        const int my_e = s_run_eq + e_before + __popc(me & lt);     // keys == thr before this one
        // slot = (# selected before me) = (# gt before me) + min(# eq before me, want_eq)
        if (gt) out[my_g + min(my_e, want_eq)] = i;
        else if (eq && my_e < want_eq) out[my_g + my_e] = i;
        __syncthreads();
        if (tid == 0) { s_run_gt += g_tot; s_run_eq += e_tot; }
        __syncthreads();
    }
}

}  // namespace
}  // namespace ume

extern "C" int ume_gumbel_topk_f32(const float* d, const float* u, int B, int n, int k, float tau, uint64_t seed,
                                   int64_t* idx, void* stream_) {
    using namespace ume;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    UME_REQUIRE(B >= 0 && n >= 0 && k >= 0, UME_ERR_BAD_ARG, "ume_gumbel_topk_f32: negative size");
    if (B == 0 || k == 0) return UME_OK;
    UME_REQUIRE(d && idx, UME_ERR_BAD_ARG, "ume_gumbel_topk_f32: null pointer");
    UME_REQUIRE(k <= n, UME_ERR_BAD_ARG, "ume_gumbel_topk_f32: k = %d > n = %d", k, n);
    UME_REQUIRE(tau > 0.f, UME_ERR_BAD_ARG, "ume_gumbel_topk_f32: tau must be positive");
    UME_REQUIRE(n <= 49152, UME_ERR_UNSUPPORTED, "ume_gumbel_topk_f32: n = %d > 49152 matches per pair", n);
    UME_REQUIRE(B <= 2147483647 / 1, UME_ERR_UNSUPPORTED, "ume_gumbel_topk_f32: B too large");
    const size_t smem = (size_t)n * sizeof(unsigned);
    cudaError_t e = cudaFuncSetAttribute(gumbel_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    UME_REQUIRE(e == cudaSuccess, UME_ERR_CUDA, "ume_gumbel_topk_f32: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    gumbel_topk_kernel<<<(unsigned)B, kSelThreads, smem, stream>>>(d, u, n, k, 1.0f / tau, (unsigned long long)seed, idx);
    count_launch();
    return check_launch("gumbel_topk_kernel");
}
