// Search-grid construction: bins each cloud's points into uniform cells around the queries.
// HBM-bound integer/byte work: one coalesced pass to count, one to scatter.
#include "ume_common.cuh"

#include <float.h>

namespace ume {

namespace {

UME_DEVI int float_order_key(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
UME_DEVI float float_from_key(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }

__global__ void grid_init_kernel(int* __restrict__ bbox, int B) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B * 6) bbox[i] = (i % 6 < 3) ? INT_MAX : INT_MIN;   // [min xyz | max xyz] as ordered keys
}

// clouds [0, Bs) come from the first array, clouds [Bs, B) from the second one (a source / target pair of batches
// handled by ONE launch; Bs = B and a null second array for a single batch)
UME_DEVI const float* cloud_of(const float* a, const float* a2, int Bs, int b, size_t stride) {
    return (b < Bs) ? a + (size_t)b * stride : a2 + (size_t)(b - Bs) * stride;
}

__global__ void grid_bbox_kernel(const float* __restrict__ q, const float* __restrict__ q2, int Bs, int nq, int* __restrict__ bbox) {
    const int b = blockIdx.y;
    const float* qb = cloud_of(q, q2, Bs, b, (size_t)nq * 3);
    int mn[3] = {INT_MAX, INT_MAX, INT_MAX}, mx[3] = {INT_MIN, INT_MIN, INT_MIN};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += gridDim.x * blockDim.x) {
        float x = qb[i * 3 + 0], y = qb[i * 3 + 1], z = qb[i * 3 + 2];
        if (isfinite(x) && isfinite(y) && isfinite(z)) {
            int kx = float_order_key(x), ky = float_order_key(y), kz = float_order_key(z);
            mn[0] = min(mn[0], kx); mx[0] = max(mx[0], kx);
            mn[1] = min(mn[1], ky); mx[1] = max(mx[1], ky);
            mn[2] = min(mn[2], kz); mx[2] = max(mx[2], kz);
        }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        mn[d] = __reduce_min_sync(UME_FULL_MASK, mn[d]);
        mx[d] = __reduce_max_sync(UME_FULL_MASK, mx[d]);
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            atomicMin(&bbox[b * 6 + d], mn[d]);
            atomicMax(&bbox[b * 6 + 3 + d], mx[d]);
        }
    }
}

__global__ void grid_params_kernel(const int* __restrict__ bbox, GridHeader* __restrict__ hdr, int B,
                                   float expand, float cell, int cells_cap, int N) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    GridHeader h;
    float lo[3], hi[3];
    bool valid = true;
    float amax = 0.f;
    for (int d = 0; d < 3; ++d) {
        int kmin = bbox[b * 6 + d], kmax = bbox[b * 6 + 3 + d];
        if (kmin == INT_MAX || kmax == INT_MIN) valid = false;
        lo[d] = float_from_key(kmin);
        hi[d] = float_from_key(kmax);
        amax = fmaxf(amax, fmaxf(fabsf(lo[d]), fabsf(hi[d])));
    }
    const float r = fabsf(expand);
    if (!valid || !isfinite(r)) {
        // no usable query: an empty domain (nothing passes x >= +inf)
        h.ox = h.oy = h.oz = INFINITY;
        h.hx = h.hy = h.hz = -INFINITY;
        h.inv_s = 1.f; h.s = 1.f; h.nx = h.ny = h.nz = 1; h.ncells = 1; h.n_sorted = 0;
        h.pad[0] = h.pad[1] = h.pad[2] = 0;
        hdr[b] = h;
        return;
    }
    const float margin = r * 1e-3f + amax * 1e-5f + 1e-6f;
    float ext[3];
    for (int d = 0; d < 3; ++d) {
        lo[d] -= (r + margin);
        hi[d] += (r + margin);
        ext[d] = hi[d] - lo[d];
    }
    float emax = fmaxf(ext[0], fmaxf(ext[1], ext[2]));
    float s = cell;
    if (s < 0.f) {
        // automatic, about -cell points per cell for a surface-like (planar) cloud of N points
        const float fl = emax * 1e-3f;
        float e0 = fmaxf(ext[0], fl), e1 = fmaxf(ext[1], fl), e2 = fmaxf(ext[2], fl);
        const float emin = fminf(e0, fminf(e1, e2));
        s = sqrtf(e0 * e1 * e2 / emin * (-cell) / (float)max(N, 1));
    } else if (!(s > 0.f)) {
        // automatic: the finest grid the table allows (the loop below coarsens it until it fits)
        const float fl = emax * 1e-3f;
        s = cbrtf(fmaxf(ext[0], fl) * fmaxf(ext[1], fl) * fmaxf(ext[2], fl) / (float)cells_cap);
    }
    if (!(s > emax * 1e-6f)) s = emax;          // absurdly small cell: a single cell
    if (!(s > 0.f)) s = 1.f;
    // Radius-query grids (cell and expand given) over a slab-like domain — the queries' z range is at
    // most 2 r, the LiDAR case — get ONE layer of cells in z: the table shrinks to nx*ny cells (a
    // finer xy grid fits), a query touches a handful of cell rows instead of rows x layers, and the
    // exact distance test sorts out z anyway.  (cell_coord clamps z to layer 0; neighbors.cuh treats
    // the outermost layers as unbounded.)  The k-NN grids keep cubic cells: their ring bounds need them.
    const bool flat = (cell > 0.f) && (r > 0.f) && (ext[2] <= 4.1f * r);
    int n[3];
    for (int it = 0; it < 256; ++it) {
        double prod = 1.0;
        for (int d = 0; d < 3; ++d) {
            float c = floorf(ext[d] / s) + 1.f;
            n[d] = (c < 1.f) ? 1 : (c > 1.0e6f ? 1000000 : (int)c);
            if (d == 2 && flat) n[d] = 1;
            prod *= (double)n[d];
        }
        if (prod <= (double)cells_cap) break;
        s *= 1.26f;
        if (it == 255) { n[0] = n[1] = n[2] = 1; s = emax; }
    }
    h.ox = lo[0]; h.oy = lo[1]; h.oz = lo[2];
    h.hx = hi[0]; h.hy = hi[1]; h.hz = hi[2];
    h.s = s; h.inv_s = 1.0f / s;
    h.nx = n[0]; h.ny = n[1]; h.nz = n[2];
    h.ncells = n[0] * n[1] * n[2];
    h.n_sorted = 0;
    h.pad[0] = h.pad[1] = h.pad[2] = 0;
    hdr[b] = h;
}

UME_DEVI bool point_cell(const GridHeader& h, float x, float y, float z, int* cell) {
    bool inside = (x >= h.ox) && (x <= h.hx) && (y >= h.oy) && (y <= h.hy) && (z >= h.oz) && (z <= h.hz);
    if (!inside) return false;
    int cx = cell_coord(x, h.ox, h.inv_s, h.nx);
    int cy = cell_coord(y, h.oy, h.inv_s, h.ny);
    int cz = cell_coord(z, h.oz, h.inv_s, h.nz);
    *cell = (cz * h.ny + cy) * h.nx + cx;
    return true;
}


// ---------------------------------------------------------------- binning
// A STABLE counting sort by cell: inside a cell the points keep their row order, so the sorted array —
// and with it every sum the moment kernels form by walking it — is a pure function of the input
// (bit-reproducible from launch to launch; the first version ranked with shared-memory atomics alone
// and was not).  Each CTA owns a contiguous slice of one cloud and a shared-memory histogram:
//   pass A  every point takes a PROVISIONAL rank inside its cell with one shared-memory atomic
//           (fast, but the order the atomics land in is arbitrary);
//   scan    exclusive scan of the slice's per-cell counts: the slice's points, grouped by cell, fit a
//           shared-memory array of 16-bit slice-local row numbers (everything stays coalesced in HBM);
//   pass B  every point drops its row number into its cell's group at its provisional rank;
//   pass C  every point reads its group (about three members on average) and counts the members
//           with a smaller row index: that is its STABLE rank, whatever order the atomics took.
// The CTA's per-cell counts go to a (slice, cell) table; the scan kernel turns them into cell_start[]
// and, in place, into every slice's first slot inside each cell.  No global atomics, no memset, and
// the scatter pass is position = slice_base[slice][cell] + rank.
constexpr int kRankThreads = 512;

// slices per cloud: about four CTAs per SM in total, at least ~2048 rows per slice
static int grid_slices(int B, int N) {
    int G = (4 * 148) / (B > 0 ? B : 1);
    // ... and at most ~14k rows per slice: the slice's shared-memory footprint (CTAs per SM) and the in-cell
    // ranking pass (quadratic in the slice's rows per cell) grow with the slice
    if (G < (N + 13999) / 14000) G = (N + 13999) / 14000;
    const int cap = N / 2048 < 64 ? N / 2048 : 64;
    if (G > cap) G = cap;
    const int need = (N + 65534) / 65535;                    // slice-local row numbers are 16 bits wide
    if (G < need) G = need;
    return G < 1 ? 1 : G;
}

__global__ void __launch_bounds__(kRankThreads)
grid_rank_kernel(const float* __restrict__ pts, const float* __restrict__ pts2, int Bs, int N, int per,
                 const GridHeader* __restrict__ hdr, int* __restrict__ slice_cnt, int cells_cap, int* __restrict__ cell_of,
                 int* __restrict__ rank_of) {
    // 16-bit shared memory throughout (a slice holds < 65536 rows): cells_cap + 2 per-cell counters — bumped
    // two to a 32-bit word by the atomics, then turned in place into exclusive starts — and `per` slots
    // for the slice's row numbers grouped by cell.  43 KB at the usual sizes: four CTAs per SM.
    extern __shared__ unsigned s_words[];
    unsigned short* s_cnt = reinterpret_cast<unsigned short*>(s_words);
    unsigned short* tmp = s_cnt + cells_cap + 2;
    __shared__ int s_warp[kRankThreads / 32];
    __shared__ int s_running;
    const int b = blockIdx.y, G = gridDim.x, g = blockIdx.x;
    const GridHeader h = hdr[b];
    const int ncells = h.ncells;
    const int lo = min(N, g * per), hi = min(N, lo + per);
    const float* pb = cloud_of(pts, pts2, Bs, b, (size_t)N * 3);
    int* cell_b = cell_of + (size_t)b * N;
    int* rank_b = rank_of + (size_t)b * N;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int w = tid; w <= (ncells + 1) / 2; w += kRankThreads) s_words[w] = 0u;
    if (tid == 0) s_running = 0;
    __syncthreads();
    // pass A: cell and provisional rank
    for (int i = lo + tid; i < hi; i += kRankThreads) {
        const float x = pb[(size_t)i * 3 + 0], y = pb[(size_t)i * 3 + 1], z = pb[(size_t)i * 3 + 2];
        int cell;
        if (point_cell(h, x, y, z, &cell)) {
            cell_b[i] = cell;
            const unsigned old = atomicAdd(&s_words[cell >> 1], (cell & 1) ? 65536u : 1u);
            rank_b[i] = (int)((old >> (16 * (cell & 1))) & 0xffffu);
        } else {
            cell_b[i] = -1;
        }
    }
    __syncthreads();
    // the slice's per-cell counts -> global table; exclusive scan in place (s_cnt[c] = first slot of cell c)
    int* out = slice_cnt + ((size_t)b * G + g) * cells_cap;
    for (int base = 0; base <= ncells; base += kRankThreads) {
        const int c = base + tid;
        const int v = (c < ncells) ? (int)s_cnt[c] : 0;
        if (c < ncells) out[c] = v;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(UME_FULL_MASK, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        int wbase = 0;
        for (int k = 0; k < warp; ++k) wbase += s_warp[k];
        const int run0 = s_running;
        if (c <= ncells) s_cnt[c] = (unsigned short)(run0 + wbase + incl - v);
        __syncthreads();
        if (tid == kRankThreads - 1) s_running = run0 + wbase + incl;
        __syncthreads();
    }
    // pass B: slice-local row numbers grouped by cell, in the (arbitrary) order of the provisional ranks
    for (int i = lo + tid; i < hi; i += kRankThreads) {
        const int c = cell_b[i];                             // written by this same thread
        if (c >= 0) tmp[(int)s_cnt[c] + rank_b[i]] = (unsigned short)(i - lo);
    }
    __syncthreads();
    // pass C: stable rank = members of the group with a smaller row number
    for (int i = lo + tid; i < hi; i += kRankThreads) {
        const int c = cell_b[i];
        if (c >= 0) {
            const int s0 = s_cnt[c], n = (int)s_cnt[c + 1] - s0;
            const int me = i - lo;
            int r = 0;
            for (int k = 0; k < n; ++k) r += ((int)tmp[s0 + k] < me) ? 1 : 0;
            rank_b[i] = r;
        }
    }
}

// One CTA per cloud: per-cell totals over the slices, exclusive scan over the cells in use
// (cell_start), and in place the first slot of every slice inside each cell.
__global__ void __launch_bounds__(1024) grid_scan_kernel(int* __restrict__ slice_cnt, int G, int* __restrict__ cell_start,
                                                         GridHeader* __restrict__ hdr, int cells_cap) {
    const int b = blockIdx.x;
    const int ncells = hdr[b].ncells;
    int* tab = slice_cnt + (size_t)b * G * cells_cap;
    int* cs = cell_start + (size_t)b * (cells_cap + 1);
    __shared__ int warp_tot[32];
    __shared__ int running;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (t == 0) running = 0;
    __syncthreads();
    for (int base = 0; base < ncells; base += 1024) {
        const int c = base + t;
        int v = 0;
        if (c < ncells)
            for (int g = 0; g < G; ++g) v += tab[(size_t)g * cells_cap + c];
        int incl = v;
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
        if (c < ncells) {
            int at = run0 + wbase + incl - v;
            cs[c] = at;
            for (int g = 0; g < G; ++g) {
                const int n = tab[(size_t)g * cells_cap + c];
                tab[(size_t)g * cells_cap + c] = at;
                at += n;
            }
        }
        __syncthreads();
        if (t == 1023) running = run0 + wbase + incl;
        __syncthreads();
    }
    if (t == 0) {
        cs[ncells] = running;
        hdr[b].n_sorted = running;
    }
}

__global__ void grid_scatter_kernel(const float* __restrict__ pts, const float* __restrict__ pts2, int Bs, int N, int per, int G,
                                    const int* __restrict__ slice_base,
                                    int cells_cap, const int* __restrict__ cell_of, const int* __restrict__ rank_of,
                                    float4* __restrict__ sorted) {
    const int b = blockIdx.y;
    const float* pb = cloud_of(pts, pts2, Bs, b, (size_t)N * 3);
    const int* tab = slice_base + (size_t)b * G * cells_cap;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        const int c = cell_of[(size_t)b * N + i];
        if (c >= 0) {
            const int pos = __ldg(&tab[(size_t)(i / per) * cells_cap + c]) + rank_of[(size_t)b * N + i];
            sorted[(size_t)b * N + pos] = make_float4(pb[(size_t)i * 3 + 0], pb[(size_t)i * 3 + 1], pb[(size_t)i * 3 + 2], __int_as_float(i));
        }
    }
}

}  // namespace

size_t grid_workspace_bytes(int B, int N, int cells_cap) {
    size_t s = 0;
    s = align_up(s, 256) + (size_t)B * sizeof(GridHeader);
    s = align_up(s, 256) + (size_t)B * 6 * sizeof(int);
    s = align_up(s, 256) + (size_t)B * grid_slices(B, N) * cells_cap * sizeof(int);
    s = align_up(s, 256) + (size_t)B * (cells_cap + 1) * sizeof(int);
    s = align_up(s, 256) + (size_t)B * N * sizeof(int);
    s = align_up(s, 256) + (size_t)B * N * sizeof(int);
    s = align_up(s, 256) + (size_t)B * N * sizeof(float4);
    return align_up(s, 256);
}

int grid_build(const float* pts, const float* q, int B, int N, int nq, float expand, float cell,
               int cells_cap, Workspace& ws, GridView* view, cudaStream_t stream, const float* pts2, const float* q2, int B2) {
    const int Bs = B;                         // clouds [0, Bs) from (pts, q), clouds [Bs, Bs + B2) from (pts2, q2)
    B += (pts2 != nullptr) ? B2 : 0;
    const int G = grid_slices(B, N);
    const int per = (N + G - 1) / G;
    GridHeader* hdr = ws.take<GridHeader>(B);
    int* bbox = ws.take<int>((size_t)B * 6);
    int* slice_cnt = ws.take<int>((size_t)B * G * cells_cap);
    int* cell_start = ws.take<int>((size_t)B * (cells_cap + 1));
    int* cell_of = ws.take<int>((size_t)B * N);
    int* rank_of = ws.take<int>((size_t)B * N);
    float4* sorted = ws.take<float4>((size_t)B * N);
    UME_REQUIRE(ws.ok(), UME_ERR_WORKSPACE, "grid_build: workspace too small (%zu needed, %zu given)",
                ws.used, ws.size);
    UME_REQUIRE(B <= 65535, UME_ERR_UNSUPPORTED, "grid_build: more than 65535 clouds per call");
    const size_t smem = (size_t)(cells_cap + 2 + per) * sizeof(unsigned short) + 16;
    cudaError_t e = cudaFuncSetAttribute(grid_rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    UME_REQUIRE(e == cudaSuccess, UME_ERR_CUDA, "grid_build: cudaFuncSetAttribute: %s", cudaGetErrorString(e));

    ProfScope prof(UME_PROF_GRID, stream);
    grid_init_kernel<<<(B * 6 + 127) / 128, 128, 0, stream>>>(bbox, B);
    if (nq > 0) {
        dim3 g((unsigned)min((nq + 255) / 256, 64), (unsigned)B);
        grid_bbox_kernel<<<g, 256, 0, stream>>>(q, q2, Bs, nq, bbox);
    }
    grid_params_kernel<<<(B + 127) / 128, 128, 0, stream>>>(bbox, hdr, B, expand, cell, cells_cap, N);
    grid_rank_kernel<<<dim3((unsigned)G, (unsigned)B), kRankThreads, smem, stream>>>(pts, pts2, Bs, N, per, hdr, slice_cnt, cells_cap,
                                                                                   cell_of, rank_of);
    grid_scan_kernel<<<B, 1024, 0, stream>>>(slice_cnt, G, cell_start, hdr, cells_cap);
    dim3 gs((unsigned)min((N + 255) / 256, 296), (unsigned)B);
    grid_scatter_kernel<<<gs, 256, 0, stream>>>(pts, pts2, Bs, N, per, G, slice_cnt, cells_cap, cell_of, rank_of, sorted);
    count_launch(nq > 0 ? 6 : 5);
    view->hdr = hdr;
    view->cell_start = cell_start;
    view->sorted = sorted;
    view->cells_cap = cells_cap;
    view->N = N;
    return check_launch("grid_build");
}

}  // namespace ume
