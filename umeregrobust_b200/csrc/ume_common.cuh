// Shared device/host helpers for the umereg_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/umereg_b200.h"

#define UME_DEVI __device__ __forceinline__
#define UME_FULL_MASK 0xffffffffu

namespace ume {

// ---------------------------------------------------------------- host-side error plumbing
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int check_launch(const char* what);   // cudaGetLastError -> UME_OK / UME_ERR_CUDA
const char* last_error();
uint64_t launches();
int prof_begin(int slot, cudaStream_t stream);            // -1 when profiling is off
void prof_end(int slot, int token, cudaStream_t stream);

// RAII bracket: events around everything launched on `stream` during its lifetime
struct ProfScope {
    int slot, token;
    cudaStream_t stream;
    ProfScope(int s, cudaStream_t st) : slot(s), token(prof_begin(s, st)), stream(st) {}
    ~ProfScope() { prof_end(slot, token, stream); }
};

#define UME_REQUIRE(cond, status, ...)            \
    do {                                          \
        if (!(cond)) {                            \
            ::ume::set_error(__VA_ARGS__);        \
            return (status);                      \
        }                                         \
    } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over the caller's workspace.
struct Workspace {
    char* base;
    size_t size;
    size_t used;
    Workspace(void* p, size_t n) : base(static_cast<char*>(p)), size(n), used(0) {}
    template <typename T>
    T* take(size_t count) {
        used = align_up(used, 256);
        T* p = reinterpret_cast<T*>(base + used);
        used += count * sizeof(T);
        return p;
    }
    bool ok() const { return base != nullptr && used <= size; }
};

// ---------------------------------------------------------------- search grid
// Per-cloud uniform grid over the (radius-expanded) bounding box of the QUERIES.  Points outside
// the box cannot be within `radius` of any query and are dropped at build time.  The grid only
// pre-filters candidates: membership is always decided by the exact fp32 distance test, so the
// cell size is free (it is enlarged until the table fits `cells_cap`).
struct GridHeader {
    float ox, oy, oz;      // lower corner of the domain
    float hx, hy, hz;      // upper corner of the domain
    float inv_s;           // 1 / cell size
    float s;               // cell size
    int nx, ny, nz;
    int ncells;
    int n_sorted;          // points kept (inside the domain)
    int pad[3];
};
static_assert(sizeof(GridHeader) == 64, "GridHeader must stay 64 bytes");

struct GridView {
    const GridHeader* hdr;     // [B]
    const int* cell_start;     // [B][cells_cap + 1]
    const float4* sorted;      // [B][N]  (x, y, z, row index bits), grouped by cell
    int cells_cap;
    int N;
};

// Monotone non-decreasing in x (fp32 sub, mul by a positive constant, floor, clamp): a point and a
// query interval [lo,hi] evaluated with this same function can never disagree about coverage.
UME_DEVI int cell_coord(float x, float o, float inv_s, int n) {
    float f = floorf((x - o) * inv_s);
    int c = (f > 0.0f) ? ((f >= (float)n) ? n - 1 : (int)f) : 0;   // NaN -> 0
    return c;
}

// pytorch3d BallQueryKernel / knn arithmetic: diff = p1 - p2 per axis, dist2 accumulated in axis
// order starting from 0.  kFma=false: every product and sum separately rounded; kFma=true:
// dist2 = fma(diff, diff, dist2) as nvcc contracts it.  (e = point - query; the sign is irrelevant.)
template <bool kFma>
UME_DEVI float dist2_ordered(float ex, float ey, float ez) {
    if (kFma) {
        float s = __fmul_rn(ex, ex);                 // fma(ex,ex,0) == rn(ex*ex)
        s = __fmaf_rn(ey, ey, s);
        s = __fmaf_rn(ez, ez, s);
        return s;
    } else {
        float s = __fmul_rn(ex, ex);                 // 0 + ex*ex is exact
        s = __fadd_rn(s, __fmul_rn(ey, ey));
        s = __fadd_rn(s, __fmul_rn(ez, ez));
        return s;
    }
}

UME_DEVI float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

UME_DEVI unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// Host entry points implemented in grid.cu
size_t grid_workspace_bytes(int B, int N, int cells_cap);
// Builds the grid for clouds `pts` (B,N,3) over the bounding box of `q` (B,nq,3) grown by
// `expand`; cell size `cell` (0: the finest the table allows; < 0: about -cell points per cell for
// a surface-like cloud).  Carves its buffers out of `ws` and fills `view`.  With (pts2, q2, B2) the grids of a
// second batch of B2 clouds (same N, nq) are built by the same launches: clouds [B, B + B2) of the view.
int grid_build(const float* pts, const float* q, int B, int N, int nq, float expand, float cell,
               int cells_cap, Workspace& ws, GridView* view, cudaStream_t stream, const float* pts2 = nullptr,
               const float* q2 = nullptr, int B2 = 0);

// Cells per cloud.  8192 keeps the binning kernel's shared-memory histogram at 32 KB (4 CTAs per SM)
// and the cell table L1/L2-friendly; a KITTI-shape cloud at cell = radius needs ~4.6 k cells.
#ifndef UME_CELLS_CAP
#define UME_CELLS_CAP 8192
#endif
static constexpr int kCellsCap = UME_CELLS_CAP;
static constexpr int kMaxPoints = 2097152;

}  // namespace ume
