/*
 * ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the three pytorch3d==0.7.7 ops the reference's UME hot path calls
 * (pytorch3d is an un-vendored dependency: requirements.txt:3; it is not installable offline,
 * so its published algorithm is restated here).  Call sites in the reference that anchor the
 * semantics:
 *   ball_query : evaluate.py:51, utils/loc_utils.py:383-384
 *   knn_points : evaluate.py:272,274, utils/loc_utils.py:580,623
 *
 * ball_query semantics (pytorch3d/csrc/ball_query): for every query i of batch n, walk p2 in
 * ROW ORDER j = 0..P2-1, accumulate dist2 = sum_d (p1[d]-p2[d])^2 in fp32 in d-order, keep j when
 * dist2 < radius*radius (strict), stop after K hits.  Unfilled slots: idx = -1, dist = 0.
 * Neighbours are therefore NOT sorted by distance and NOT the nearest K.
 *
 * knn_points semantics: squared L2, K smallest, ascending; among equal distances the lower row
 * index wins (the scan replaces only on strict '<').
 *
 * Arithmetic mode: `use_fma == 0` evaluates diff*diff and the adds as separately rounded fp32
 * operations (what pytorch3d's CPU build and a torch/numpy restatement do); `use_fma != 0`
 * contracts dist2 += diff*diff into one fused multiply-add per axis (what nvcc's default
 * -fmad=true does to pytorch3d's CUDA kernel).  Build with -ffp-contract=off so the compiler
 * never chooses for us.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

static inline float dist2_plain(const float* a, const float* b) {
    float d0 = a[0] - b[0];
    float d1 = a[1] - b[1];
    float d2 = a[2] - b[2];
    float s = 0.0f;
    s = s + d0 * d0;
    s = s + d1 * d1;
    s = s + d2 * d2;
    return s;
}

static inline float dist2_fused(const float* a, const float* b) {
    float d0 = a[0] - b[0];
    float d1 = a[1] - b[1];
    float d2 = a[2] - b[2];
    float s = 0.0f;
    s = __builtin_fmaf(d0, d0, s);
    s = __builtin_fmaf(d1, d1, s);
    s = __builtin_fmaf(d2, d2, s);
    return s;
}

/* p1: (B,P1,3)  p2: (B,P2,3)  idx: (B,P1,K) int64  dists: (B,P1,K) f32  nn: (B,P1,K,3) or NULL
 * scan_len (optional, (B,P1) int64): number of p2 rows visited before the walk stopped. */
int oracle_ball_query_f32(const float* p1, const float* p2, int64_t B, int64_t P1, int64_t P2,
                          int64_t K, float radius, int use_fma, int64_t* idx, float* dists,
                          float* nn, int64_t* scan_len, int num_threads) {
    const float r2 = radius * radius;
#ifdef _OPENMP
    if (num_threads > 0) omp_set_num_threads(num_threads);
#endif
    const int64_t total = B * P1;
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t q = 0; q < total; ++q) {
        const int64_t b = q / P1;
        const float* query = p1 + q * 3;
        const float* cloud = p2 + b * P2 * 3;
        int64_t* oi = idx + q * K;
        float* od = dists + q * K;
        float* on = nn ? nn + q * K * 3 : NULL;
        int64_t cnt = 0, j = 0;
        for (; j < P2 && cnt < K; ++j) {
            const float d = use_fma ? dist2_fused(query, cloud + j * 3)
                                    : dist2_plain(query, cloud + j * 3);
            if (d < r2) {
                oi[cnt] = j;
                od[cnt] = d;
                if (on) {
                    on[cnt * 3 + 0] = cloud[j * 3 + 0];
                    on[cnt * 3 + 1] = cloud[j * 3 + 1];
                    on[cnt * 3 + 2] = cloud[j * 3 + 2];
                }
                ++cnt;
            }
        }
        if (scan_len) scan_len[q] = j;
        for (int64_t k = cnt; k < K; ++k) {
            oi[k] = -1;
            od[k] = 0.0f;
            if (on) { on[k * 3] = 0.0f; on[k * 3 + 1] = 0.0f; on[k * 3 + 2] = 0.0f; }
        }
    }
    return 0;
}

/* K-nearest by insertion into a sorted window; strict '<' keeps the lower row index on ties. */
int oracle_knn_points_f32(const float* p1, const float* p2, int64_t B, int64_t P1, int64_t P2,
                          int64_t K, int use_fma, int64_t* idx, float* dists, int num_threads) {
#ifdef _OPENMP
    if (num_threads > 0) omp_set_num_threads(num_threads);
#endif
    if (K > P2) return -1;
    const int64_t total = B * P1;
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t q = 0; q < total; ++q) {
        const int64_t b = q / P1;
        const float* query = p1 + q * 3;
        const float* cloud = p2 + b * P2 * 3;
        int64_t* oi = idx + q * K;
        float* od = dists + q * K;
        int64_t filled = 0;
        for (int64_t j = 0; j < P2; ++j) {
            const float d = use_fma ? dist2_fused(query, cloud + j * 3)
                                    : dist2_plain(query, cloud + j * 3);
            if (filled < K || d < od[filled - 1]) {
                int64_t pos = filled < K ? filled : K - 1;
                while (pos > 0 && d < od[pos - 1]) {
                    od[pos] = od[pos - 1];
                    oi[pos] = oi[pos - 1];
                    --pos;
                }
                od[pos] = d;
                oi[pos] = j;
                if (filled < K) ++filled;
            }
        }
    }
    return 0;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
