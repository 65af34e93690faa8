/*
 * umereg_b200 — C ABI of the B200-native UME descriptor-and-registration hot path.
 *
 * The reference (yuvalH9/UMERegRobust) has no FFI / plugin boundary: it is pure Python whose
 * native work happens inside third-party wheels.  This header is therefore the boundary a
 * maintainer would bind (ctypes stub shown in INTEGRATION.md); every entry point names the
 * reference call it replaces (file:line into the reference tree).
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name ends in `_host`; float32 row-major,
 *     contiguous; nothing is allocated behind the caller's back: scratch memory comes in through
 *     (`ws`, `ws_bytes`) sized by the matching `*_workspace_bytes` query;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), never synchronises,
 *     and is safe to call concurrently from different host threads on different streams with
 *     different workspaces;
 *   - return value: UME_OK (0) or a negative ume_status; `ume_last_error()` holds a thread-local
 *     human-readable message for the last failure.  No exception crosses this boundary.
 *   - B = clouds in the batch, N = points per cloud, n = keypoints per cloud, C = feature
 *     channels, K = max neighbours, M = 4 monomials [1,x,y,z].
 */
#ifndef UMEREG_B200_H_
#define UMEREG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UME_ABI_VERSION 1

typedef enum ume_status {
    UME_OK = 0,
    UME_ERR_BAD_ARG = -1,      /* null pointer, non-positive size, unsupported C / K        */
    UME_ERR_WORKSPACE = -2,    /* ws == NULL or ws_bytes too small                          */
    UME_ERR_UNSUPPORTED = -3,  /* size outside what this build supports (see each function) */
    UME_ERR_CUDA = -4          /* a CUDA runtime call / kernel launch failed                */
} ume_status;

/* flags */
#define UME_FLAG_FMA_DIST      1u  /* dist2 with fused multiply-add (what nvcc -fmad=true makes of
                                      pytorch3d's CUDA kernel); default: separately rounded mul/add
                                      (pytorch3d CPU build, torch/numpy restatements)           */
#define UME_FLAG_CELL_DIV2     2u  /* search grid with cell = radius/2 instead of radius        */
#define UME_FLAG_RAW_MOMENTS   8u  /* ume_moments_f32: F = [sum f | sum f x^T] without the division of
                                      evaluate.py:59 (training path: utils/loc_utils.py:157-161 with
                                      normalized_ume=False, and the differentiable wrapper)          */
#define UME_FLAG_CTA_MOMENTS   4u  /* ume_moments_f32: force the CTA-per-keypoint kernel (the default
                                      for C in {16,32,64,128} and B*n >= 3072 keypoints is the
                                      warp-per-keypoint kernel)                                      */
#define UME_FLAG_WARP_MOMENTS  16u /* ume_moments_f32: warp-per-keypoint kernel also for small launches */

int ume_abi_version(void);
const char* ume_last_error(void);
const char* ume_status_string(int status);

/* ---------------------------------------------------------------- neighbour search
 * Replaces pytorch3d.ops.ball_query as called at evaluate.py:51 and utils/loc_utils.py:383-384:
 * for every query the FIRST K rows of p2 (in row order) with dist2 < radius^2 (strict).
 *   p1 (B,P1,3) queries, p2 (B,P2,3) cloud
 *   idx   (B,P1,K) int64, -1 padded           (may be NULL)
 *   dists (B,P1,K) squared distances, 0 padded (may be NULL)
 *   nn    (B,P1,K,3) neighbour xyz, 0 padded   (may be NULL)
 *   count (B,P1) int32 number of neighbours found (may be NULL)
 * Limits: K <= 8192, P2 <= 8388608. */
size_t ume_ball_query_workspace_bytes(int B, int P1, int P2, int K);
int ume_ball_query_f32(const float* p1, const float* p2, int B, int P1, int P2, int K, float radius,
                       unsigned flags, int64_t* idx, float* dists, float* nn, int32_t* count,
                       void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------- fused gather + UME moments
 * Replaces evaluate.py:50-60 `my_ume_generation` (ball_query + -1 padding + torch.gather of the
 * (B,n,K,C) neighbour features + the two reductions + normalisation) and the same math in
 * utils/loc_utils.py:365-372 (`ume_kp_layer.ume_mat`).  The (B,n,K,C) tensor is never formed.
 *   pts (B,N,3), kpts (B,n,3), feat (B,N,C)
 *   F   (B,n,C,4)  = [sum f | sum f x^T] / (sum_c sum f + 1e-6), absolute coordinates (as the reference)
 *   Fc  (B,n,C,4)  same matrix with coordinates relative to the keypoint (may be NULL); it spans
 *                  the same column space as F and is better conditioned
 *   count (B,n) int32 neighbours used per keypoint (may be NULL)
 * Limits: C multiple of 4 in [4,128] or any C <= 256 (slower path); N <= 8388608; any K >= 1. */
size_t ume_moments_workspace_bytes(int B, int N, int n, int C, int K);
int ume_moments_f32(const float* pts, const float* kpts, const float* feat, int B, int N, int n,
                    int C, int K, float radius, unsigned flags, float* F, float* Fc, int32_t* count,
                    void* ws, size_t ws_bytes, void* stream);

/* The source and the target batch of a registration step (evaluate.py:206-207 calls my_ume_generation twice) in ONE
 * search-grid build and ONE moment launch.  pts1/kpts1/feat1 and pts2/kpts2/feat2 as in ume_moments_f32, the same
 * B, N, n, C on both sides; F, Fc (may be NULL), count (may be NULL) hold 2B clouds: [0,B) side 1, [B,2B) side 2.
 * Results are bit-identical to two ume_moments_f32 calls.  C in {16,32,64,128}; workspace:
 * ume_moments_workspace_bytes(2B, N, n, C, K). */
int ume_moments_pair_f32(const float* pts1, const float* kpts1, const float* feat1, const float* pts2, const float* kpts2,
                         const float* feat2, int B, int N, int n, int C, int K, float radius, unsigned flags, float* F,
                         float* Fc, int32_t* count, void* ws, size_t ws_bytes, void* stream);

/* Gradient of the RAW moments with respect to the features (the backward pass of the
 * training-time UME generation, utils/loc_utils.py:86-188, which the reference gets from autograd
 * through its materialised (B,n,K,C) gather):
 *   grad_feat[b,j,c] += sum over keypoints i whose neighbourhood holds row j of
 *                       gF[b,i,c,0] + gF[b,i,c,1:4] . pts[b,j,:]
 * Same neighbourhoods as ume_moments_f32 with the same (K, radius, flags); grad_feat (B,N,C) must be
 * initialised by the caller (it is accumulated into with vector atomics, so the last bits depend
 * on the order of arrival).  C in {16,32,64,128}. */
int ume_moments_backward_f32(const float* pts, const float* kpts, const float* gF, int B, int N, int n, int C,
                             int K, float radius, unsigned flags, float* grad_feat, void* ws, size_t ws_bytes,
                             void* stream);

/* count[b,i] = min(K, number of rows of pts[b] with dist2 < radius^2 from kpts[b,i]) without building
 * anything else: the dense-neighbourhood filter of utils/loc_utils.py:119 ((bq_idxs > -1).sum >= min_nn)
 * without the (B,n,K) index tensor.  Workspace: ume_moments_workspace_bytes. */
int ume_neighbor_count_f32(const float* pts, const float* kpts, int B, int N, int n, int K, float radius,
                           unsigned flags, int32_t* count, void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------- subspace descriptor
 * Replaces the two torch.linalg.qr calls of utils/loc_utils.py:9,11 (and :338,341): an
 * orthonormal basis of the column space of each C x 4 matrix.  The projector Q Q^T does not
 * depend on the basis chosen, so no particular QR sign convention is reproduced.
 *   F  (nmat, C, 4)  ->  Qt (nmat, 4, C)   (basis vectors as ROWS: the K-major GEMM operand)
 * Rank-deficient input: missing directions are completed with canonical unit vectors
 * (all-zero input gives e0..e3, like LAPACK); `rank` (nmat) int32 may be NULL. */
int ume_orthonormalize_f32(const float* F, int64_t nmat, int C, float* Qt, int32_t* rank, void* stream);

/* Same, additionally (or only: Qt may be NULL) emitting the distance kernel's tensor-core operand
 *   Qh (nmat, 4, 2C) __half: row = [hi (C) | lo (C)], hi = fp16(256 q), lo = fp16(256 q - hi)
 * (see ume_cdist_split_f16).  The fused pipeline uses this to skip the split pre-pass. */
int ume_orthonormalize_split_f32(const float* F, int64_t nmat, int C, float* Qt, void* Qh, int32_t* rank,
                                 void* stream);

/* ---------------------------------------------------------------- all-pairs subspace distance
 * Replaces utils/loc_utils.py:12-13 (P = QQ^T, torch.cdist(P1.flatten, P2.flatten)/sqrt 2) and the
 * arg-min of evaluate.py:224 through the identity D^2 = 4 - |Q1^T Q2|_F^2.
 *   Qt1 (B,n1,4,C), Qt2 (B,n2,4,C)
 *   D      (B,n1,n2) (may be NULL)
 *   argmin (B,n1) int64 = first index of the row minimum (may be NULL)
 *   dmin   (B,n1) the row minimum (may be NULL)
 * `impl`: 0 = fp32 SIMT kernel, 1 = tcgen05 tensor-core kernel (operands split into two FP16 numbers,
 *         three kind::f16 products per term: fp32-grade; descriptor entries must satisfy |q| < 255).
 * Limits: C multiple of 4, C <= 128. */
size_t ume_cdist_workspace_bytes(int B, int n1, int n2, int C, int impl);
int ume_cdist_f32(const float* Qt1, const float* Qt2, int B, int n1, int n2, int C, int impl, float* D,
                  int64_t* argmin, float* dmin, void* ws, size_t ws_bytes, void* stream);

/* The tensor-core distance kernel on pre-split operands (ume_orthonormalize_split_f32's Qh):
 *   Qh1 (B,n1,4,2C), Qh2 (B,n2,4,2C) __half.  Same outputs as ume_cdist_f32; no workspace.  C = 32 or 64. */
int ume_cdist_split_f16(const void* Qh1, const void* Qh2, int B, int n1, int n2, int C, float* D, int64_t* argmin,
                        float* dmin, void* stream);

/* Distance between CORRESPONDING descriptors: Dp[i] = scale * sqrt(8 - 2 |Q1_i^T Q2_i|_F^2).
 * utils/loc_utils.py:344 uses scale = 0.707 (literally), ume_cdist uses 1/sqrt(2). */
int ume_pair_dist_f32(const float* Qt1, const float* Qt2, int64_t nmat, int C, float scale, float* Dp,
                      void* stream);

/* ---------------------------------------------------------------- rigid solve
 * Replaces utils/loc_utils.py:292-335,346-350 `batch_estimate_transform_ume_old(G, H)`:
 * weighted centroids, 3x3 cross moment over the C rows, SVD, det fix, translation.
 *   G, H: moment matrices, row stride C*4.  Hypothesis i of batch b uses
 *         G[b, gi[b,i]] and H[b, hi[b,i]]  (gi / hi int64, either may be NULL = identity i)
 *   nG, nH: matrices per batch entry in G / H;  nm: hypotheses per batch entry
 *   T (B,nm,4,4): T[:3,:3] = R^T, T[:3,3] = b2  (so tgt ~ T[:3,:3] src + T[:3,3])
 *   offG (B,nG,3) / offH (B,nH,3) (both or neither NULL): G / H hold moments RELATIVE to these
 *         points (the `Fc` output of ume_moments_f32); the solve then runs on the small centred
 *         numbers and t is un-centred at the end. */
int ume_rigid_solve_f32(const float* G, const float* H, const int64_t* gi, const int64_t* hi,
                        const float* offG, const float* offH, int B, int nG, int nH, int nm, int C,
                        float* T, void* stream);

/* ---------------------------------------------------------------- backward kernels of the training row
 * loss.py:137-190 (`CubeRegistrationLoss`) differentiates through batch_estimate_transform_ume_old
 * (utils/loc_utils.py:292-335): gradient of the solve above for the pairing G[i] <-> H[i].
 *   G, H (nb,C,4), gT (nb,4,4) gradient wrt T (only T[:3,:] matters) -> gG, gH (nb,C,4).
 * Closed form through the proper singular frames of the 3x3 cross moment (no 1/(s_i^2 - s_j^2) terms). */
int ume_rigid_solve_backward_f32(const float* G, const float* H, const float* gT, int64_t nb, int C, float* gG,
                                 float* gH, void* stream);

/* loss.py:84-118 (`UMEContrastiveLoss`) differentiates through ume_cdist (utils/loc_utils.py:8-15: thin QR ->
 * P = Q Q^T -> cdist / sqrt 2): gradient of D (B,n1,n2) wrt the UME matrices, closed form on the 4x4 Gram
 * blocks (no (B,n,C,C) projector, no QR autograd).
 *   F1 (B,n1,C,4), F2 (B,n2,C,4) full-rank UME matrices; Qt1 (B,n1,4,C), Qt2 (B,n2,4,C) their bases from
 *   ume_orthonormalize_f32; D, gD (B,n1,n2) the forward's distances and their gradient -> gF1, gF2.
 * Entries with D = 0 contribute nothing (as torch.cdist's backward).  4 <= C <= 128. */
size_t ume_cdist_backward_workspace_bytes(int B, int n1, int n2, int C);
int ume_cdist_backward_f32(const float* F1, const float* F2, const float* Qt1, const float* Qt2, const float* D,
                           const float* gD, int B, int n1, int n2, int C, float* gF1, float* gF2, void* ws,
                           size_t ws_bytes, void* stream);

/* Replaces utils/eval_utils.py:60-76 `relative_rotation_error(R, R_hat)`:
 *   out[i] = acos((clamp(trace(R_hat_i R_i^T), -1, 3) - 1) / 2) * 180 / pi   (degrees)
 * R, R_hat: rotation matrices, row-major, `stride` floats apart: 9 = packed (n,3,3) arrays (row pitch 3),
 * 16 = the rotation blocks of (n,4,4) transforms read in place (row pitch 4). */
int ume_rotation_error_deg_f32(const float* R, const float* R_hat, int64_t n, int stride_R, int stride_R_hat,
                               float* out, void* stream);

/* ---------------------------------------------------------------- match sub-sampling
 * Replaces evaluate.py:233-245 (`filter_by_ume_dist_cond`): draw k of the n matches of every pair
 * WITHOUT replacement with probability proportional to exp((1 - d) / tau), which the reference does
 * on the host with np.random.choice (a D2H copy and a sync per pair).  Gumbel-top-k: the k largest
 * of (1 - d)/tau - log(-log u), u ~ U(0,1) from Philox4x32-10 keyed by (seed, pair, match), or taken
 * from `u` (B,n) when it is not NULL (parity tests feed the uniforms in as data).
 *   d (B,n) match distances -> idx (B,k) int64: the selected matches in ascending order.
 * Same distribution as the reference's sampler, not the same random stream.  Limits: n <= 49152. */
int ume_gumbel_topk_f32(const float* d, const float* u, int B, int n, int k, float tau, uint64_t seed,
                        int64_t* idx, void* stream);

/* ---------------------------------------------------------------- nearest-neighbour feature transfer
 * Replaces pytorch3d.ops.knn_points(K=1) + knn_gather at evaluate.py:272-275.
 *   q (B,P1,3) queries, p (B,P2,3) cloud, x (B,P2,U) features of p (may be NULL)
 *   idx (B,P1) int64 (may be NULL), d2 (B,P1) (may be NULL), out (B,P1,U) = x[idx] (may be NULL)
 * Ties: the lower row index wins (as a row-order scan with strict '<'). */
size_t ume_knn1_workspace_bytes(int B, int P1, int P2);
int ume_knn1_gather_f32(const float* q, const float* p, const float* x, int B, int P1, int P2, int U,
                        unsigned flags, int64_t* idx, float* d2, float* out, void* ws, size_t ws_bytes,
                        void* stream);

/* ---------------------------------------------------------------- general K nearest neighbours
 * Replaces pytorch3d.ops.knn_points as used at utils/loc_utils.py:580,623 (K = 50 self-query, K = 20
 * cross-query): squared L2, the K smallest in ascending order, lower row index on ties.
 *   q (B,P1,3), p (B,P2,3) -> idx (B,P1,K) int64 (may be NULL), d2 (B,P1,K) (may be NULL).
 * Limits: 1 <= K <= min(64, P2). */
size_t ume_knn_workspace_bytes(int B, int P1, int P2);
int ume_knn_f32(const float* q, const float* p, int B, int P1, int P2, int K, unsigned flags, int64_t* idx,
                float* d2, void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------- hypothesis selection
 * Replaces utils/loc_utils.py:579-585 `feature_spatial_var(pts, feat, knn)`:
 *   out[b,i] = mean over the knn-1 nearest other rows j of |feat_i - feat_j|_2.   knn <= 64, C % 4 == 0. */
size_t ume_feature_spatial_var_workspace_bytes(int B, int N);
int ume_feature_spatial_var_f32(const float* pts, const float* feat, int B, int N, int C, int knn,
                                unsigned flags, float* out, void* ws, size_t ws_bytes, void* stream);

/* out[i,:] = (f[i,:] - mean[:]) * w[i]   (utils/loc_utils.py:649-650).  f (rows,C), mean (C), w (rows). */
int ume_weight_features_f32(const float* f, const float* mean, const float* w, int64_t rows, int C, float* out,
                            void* stream);

/* Replaces the scoring loop of FeatureCorrelator.feature_corr_hypothesis_test
 * (utils/loc_utils.py:651-662 -> pc_corr_cost_pytorch3d :621-631 -> pc_corr :592-619) for ALL
 * hypotheses in one launch (the reference chunks them by `batch`):
 *   score[h] = (1/Ns) sum_i sum_{k<K} 1/(1 + (|T_h p_i - q_nn|/sigma)^2) <wf_src_i, wf_tgt_nn>
 * with q_nn the K nearest target rows of the transformed source point.
 *   src_pts (Ns,3), tgt_pts (Nt,3), wf_src (Ns,C), wf_tgt (Nt,C), T (n_hyp,4,4) row-major
 *   score (n_hyp), best (1) int64 = arg-max (first index on ties; may be NULL).
 * Limits: C = 32 or 64, K <= 32. */
size_t ume_corr_scores_workspace_bytes(int Ns, int Nt, int n_hyp);
int ume_corr_scores_f32(const float* src_pts, const float* tgt_pts, const float* wf_src, const float* wf_tgt,
                        const float* T, int Ns, int Nt, int C, int n_hyp, int K, float sigma, unsigned flags,
                        float* score, int64_t* best, void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------- voxel de-duplication
 * Replaces MinkowskiEngine `ME.utils.sparse_quantize(coordinates, return_index=True, quantization_size=q)`
 * as called at evaluate.py:261-264: voxel = floor(p / q) (fp32 division); the FIRST row of every
 * occupied voxel survives, survivors keep their row order.
 *   pts (N,3) -> index (N) int64: the first *count entries are the surviving rows, ascending;
 *   coords (N,3) int32: their voxel coordinates (may be NULL); count (1) int32 on the device:
 *   number of survivors, or -1 when a coordinate was NaN or outside [-2^20, 2^20) voxels.
 * The caller reads `count` back to size its result (the reference call is a sync point too). */
size_t ume_voxel_unique_workspace_bytes(int N);
int ume_voxel_unique_f32(const float* pts, int N, float voxel, int64_t* index, int32_t* coords, int32_t* count,
                         void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------- optimal matching (host)
 * Replaces scipy.optimize.linear_sum_assignment at evaluate.py:216-222 (hungarian_matching_flag, off in
 * every shipped config), which the reference also runs on the host over `D[b].cpu().numpy()`.
 * HOST pointers (the only `_host` entry point): cost (n_rows, n_cols) row-major float32, finite or
 * +inf; outputs min(n_rows, n_cols) pairs with row_ind ascending — scipy's output convention.
 * Shortest augmenting paths in double precision: the minimum-cost assignment. */
int ume_linear_sum_assignment_host_f32(const float* cost_host, int n_rows, int n_cols, int64_t* row_ind_host,
                                       int64_t* col_ind_host);

/* ---------------------------------------------------------------- stage profiler
 * When enabled, every stage brackets its kernel launches with CUDA events on the launching
 * stream; ume_profile_read() synchronises those events and returns the accumulated device time
 * and the number of brackets of one stage since the last reset.  Used by bench.py for the
 * per-kernel durations behind the roofline figures; off by default. */
#define UME_PROF_GRID      0   /* search-grid build (bbox, count, scan, scatter)   */
#define UME_PROF_MOMENTS   1   /* fused gather + moment kernel                     */
#define UME_PROF_ORTHO     2   /* descriptor orthonormalisation                    */
#define UME_PROF_CDIST     3   /* distance GEMM + arg-min                          */
#define UME_PROF_RIGID     4   /* rigid solve                                      */
#define UME_PROF_BALLQUERY 5
#define UME_PROF_KNN       6
#define UME_PROF_CORR      7   /* hypothesis-selection scores                      */
#define UME_PROF_SLOTS     8
void ume_profile_enable(int on);
void ume_profile_reset(void);
int ume_profile_read(int slot, double* total_ms, uint64_t* launches);

/* ---------------------------------------------------------------- counters
 * Number of kernels this library has launched since load (all threads); bench.py reports the
 * difference over the timed region as `gpu_launches`. */
uint64_t ume_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* UMEREG_B200_H_ */
