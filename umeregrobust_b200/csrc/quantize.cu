// Voxel de-duplication: MinkowskiEngine `ME.utils.sparse_quantize(coordinates, return_index=True,
// quantization_size=q)` as called at evaluate.py:261-264 (and kitti_dataset.py:416, nuscenes_dataset.py:424).
//
// Every point gets the integer voxel floor(p / q) (fp32 division, like torch.floor(coords / q)); the
// FIRST row of each occupied voxel survives, survivors keep their row order.  (ME's CPU hash map
// inserts rows sequentially, so its unique_map is the ascending list of first occurrences.)
//
//   1. insert   open-addressing hash table keyed by the packed voxel (3 x 21 bits), value = smallest
//               row index seen (atomicCAS on the key, atomicMin on the value)
//   2. flag     row i survives iff the table value of its voxel is i; per-CTA survivor counts
//   3. scan     exclusive scan of the per-CTA counts (one CTA)
//   4. scatter  survivors written in row order (ballot/prefix inside the CTA + its base)
#include "ume_common.cuh"

namespace ume {
namespace {

constexpr unsigned long long kEmpty = 0xffffffffffffffffull;
constexpr int kQT = 256;                  // threads per CTA = rows per CTA in flag / scatter
constexpr int kRange = 1 << 20;           // voxel coordinates must lie in [-2^20, 2^20)

struct QuantParams {
    const float* pts;         // (N,3)
    unsigned long long* keys; // [cap]
    int* vals;                // [cap]
    int* cta_count;           // [nblk + 1]
    int* status;              // [0] = survivors, [1] = sticky out-of-range flag
    int64_t* index;           // (N) survivors' rows, ascending
    int32_t* coords;          // (N,3) survivors' voxel coordinates (may be null)
    int N, nblk;
    unsigned cap_mask;
    float q;
};

UME_DEVI bool voxel_key(const QuantParams& p, int i, unsigned long long& key, int& vx, int& vy, int& vz) {
    const float fx = floorf(__fdiv_rn(p.pts[(size_t)i * 3 + 0], p.q));
    const float fy = floorf(__fdiv_rn(p.pts[(size_t)i * 3 + 1], p.q));
    const float fz = floorf(__fdiv_rn(p.pts[(size_t)i * 3 + 2], p.q));
    const bool ok = fx >= -(float)kRange && fx < (float)kRange && fy >= -(float)kRange && fy < (float)kRange &&
                    fz >= -(float)kRange && fz < (float)kRange;      // NaN fails every comparison
    vx = ok ? (int)fx : 0; vy = ok ? (int)fy : 0; vz = ok ? (int)fz : 0;
    key = ((unsigned long long)(unsigned)(vx + kRange) << 42) | ((unsigned long long)(unsigned)(vy + kRange) << 21) |
          (unsigned long long)(unsigned)(vz + kRange);
    return ok;
}

UME_DEVI unsigned hash_slot(unsigned long long key, unsigned mask) {
    key ^= key >> 33; key *= 0xff51afd7ed558ccdull; key ^= key >> 33; key *= 0xc4ceb9fe1a85ec53ull; key ^= key >> 33;
    return (unsigned)key & mask;
}

__global__ void __launch_bounds__(kQT) quant_clear_kernel(QuantParams p) {
    const size_t i = (size_t)blockIdx.x * kQT + threadIdx.x;
    if (i <= p.cap_mask) { p.keys[i] = kEmpty; p.vals[i] = 0x7fffffff; }
    if (i < 2) p.status[i] = 0;
}

__global__ void __launch_bounds__(kQT) quant_insert_kernel(QuantParams p) {
    const int i = blockIdx.x * kQT + threadIdx.x;
    if (i >= p.N) return;
    unsigned long long key;
    int vx, vy, vz;
    if (!voxel_key(p, i, key, vx, vy, vz)) { atomicExch(&p.status[1], 1); return; }
    unsigned s = hash_slot(key, p.cap_mask);
    for (;;) {
        const unsigned long long old = atomicCAS(&p.keys[s], kEmpty, key);
        if (old == kEmpty || old == key) { atomicMin(&p.vals[s], i); return; }
        s = (s + 1) & p.cap_mask;                 // the table is at least twice the row count: terminates
    }
}

UME_DEVI bool survives(const QuantParams& p, int i, int& vx, int& vy, int& vz) {
    unsigned long long key;
    if (i >= p.N || !voxel_key(p, i, key, vx, vy, vz)) return false;
    unsigned s = hash_slot(key, p.cap_mask);
    while (p.keys[s] != key) s = (s + 1) & p.cap_mask;
    return p.vals[s] == i;
}

__global__ void __launch_bounds__(kQT) quant_flag_kernel(QuantParams p) {
    int vx, vy, vz;
    const int n = __syncthreads_count(survives(p, blockIdx.x * kQT + threadIdx.x, vx, vy, vz));
    if (threadIdx.x == 0) p.cta_count[blockIdx.x] = n;
}

// exclusive scan of cta_count[0..nblk) in place, total -> status[0] (one CTA, any nblk)
__global__ void __launch_bounds__(1024) quant_scan_kernel(QuantParams p) {
    __shared__ int warp_sum[32];
    __shared__ int carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < p.nblk; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = (i < p.nblk) ? p.cta_count[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(UME_FULL_MASK, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = warp_sum[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(UME_FULL_MASK, wi, o);
                if (lane >= o) wi += t;
            }
            warp_sum[lane] = wi - w;              // exclusive prefix of the warp totals
        }
        __syncthreads();
        const int excl = carry + warp_sum[warp] + incl - v;
        if (i < p.nblk) p.cta_count[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) p.status[0] = p.status[1] ? -1 : carry;   // -1: a coordinate was out of range
}

__global__ void __launch_bounds__(kQT) quant_scatter_kernel(QuantParams p) {
    __shared__ int warp_cnt[kQT / 32];
    const int i = blockIdx.x * kQT + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int vx, vy, vz;
    const bool keep = survives(p, i, vx, vy, vz);
    const unsigned m = __ballot_sync(UME_FULL_MASK, keep);
    if (lane == 0) warp_cnt[warp] = __popc(m);
    __syncthreads();
    int at = p.cta_count[blockIdx.x] + __popc(m & lanemask_lt());
    for (int w = 0; w < warp; ++w) at += warp_cnt[w];
    if (keep) {
        p.index[at] = i;
        if (p.coords) { p.coords[(size_t)at * 3 + 0] = vx; p.coords[(size_t)at * 3 + 1] = vy; p.coords[(size_t)at * 3 + 2] = vz; }
    }
}

unsigned table_cap(int N) {
    unsigned cap = 1024;
    while (cap < 2u * (unsigned)N) cap <<= 1;
    return cap;
}

}  // namespace
}  // namespace ume

extern "C" size_t ume_voxel_unique_workspace_bytes(int N) {
    if (N <= 0) return 0;
    const size_t cap = ume::table_cap(N);
    const size_t nblk = ((size_t)N + ume::kQT - 1) / ume::kQT;
    return ume::align_up(cap * 8, 256) + ume::align_up(cap * 4, 256) + ume::align_up((nblk + 1) * 4, 256) + 512;
}

extern "C" int ume_voxel_unique_f32(const float* pts, int N, float voxel, int64_t* index, int32_t* coords,
                                    int32_t* count, void* ws, size_t ws_bytes, void* stream_) {
    using namespace ume;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    UME_REQUIRE(N >= 0, UME_ERR_BAD_ARG, "ume_voxel_unique_f32: negative size");
    UME_REQUIRE(count, UME_ERR_BAD_ARG, "ume_voxel_unique_f32: null count");
    UME_REQUIRE(voxel > 0.f, UME_ERR_BAD_ARG, "ume_voxel_unique_f32: voxel size %g <= 0", (double)voxel);
    if (N == 0) {
        cudaError_t e = cudaMemsetAsync(count, 0, sizeof(int32_t), stream);
        UME_REQUIRE(e == cudaSuccess, UME_ERR_CUDA, "ume_voxel_unique_f32: %s", cudaGetErrorString(e));
        return UME_OK;
    }
    UME_REQUIRE(pts && index, UME_ERR_BAD_ARG, "ume_voxel_unique_f32: null pointer");
    UME_REQUIRE(N <= (1 << 29), UME_ERR_UNSUPPORTED, "ume_voxel_unique_f32: N = %d too large", N);
    UME_REQUIRE(ws && ws_bytes >= ume_voxel_unique_workspace_bytes(N), UME_ERR_WORKSPACE,
                "ume_voxel_unique_f32: workspace too small (%zu needed, %zu given)",
                ume_voxel_unique_workspace_bytes(N), ws_bytes);
    Workspace w(ws, ws_bytes);
    QuantParams p;
    const unsigned cap = table_cap(N);
    p.pts = pts; p.N = N; p.q = voxel; p.cap_mask = cap - 1;
    p.nblk = (N + kQT - 1) / kQT;
    p.keys = w.take<unsigned long long>(cap);
    p.vals = w.take<int>(cap);
    p.cta_count = w.take<int>((size_t)p.nblk + 1);
    p.status = w.take<int>(2);
    p.index = index; p.coords = coords;
    ProfScope prof(UME_PROF_KNN, stream);
    quant_clear_kernel<<<(cap + kQT - 1) / kQT, kQT, 0, stream>>>(p);
    quant_insert_kernel<<<p.nblk, kQT, 0, stream>>>(p);
    quant_flag_kernel<<<p.nblk, kQT, 0, stream>>>(p);
    quant_scan_kernel<<<1, 1024, 0, stream>>>(p);
    quant_scatter_kernel<<<p.nblk, kQT, 0, stream>>>(p);
    count_launch(5);
    cudaError_t e = cudaMemcpyAsync(count, p.status, sizeof(int32_t), cudaMemcpyDeviceToDevice, stream);
    UME_REQUIRE(e == cudaSuccess, UME_ERR_CUDA, "ume_voxel_unique_f32: %s", cudaGetErrorString(e));
    return check_launch("voxel_unique");
}
