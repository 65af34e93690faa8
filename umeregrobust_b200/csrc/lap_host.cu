// Optimal one-to-one matching of keypoints ("Hungarian" option, evaluate.py:216-222).
//
// The reference runs scipy.optimize.linear_sum_assignment on the HOST, on a copy of the distance
// matrix (`D[b_idx].cpu().numpy()`), once per pair; the option is off in every shipped config.  This
// is the same boundary: a host routine over a host copy of D — the shortest-augmenting-path
// algorithm with dual variables (Jonker-Volgenant as restated by Crouse, "On implementing 2D
// rectangular assignment algorithms", IEEE TAES 2016; scipy's solver implements the same paper), in
// double precision like scipy, O(n^3) worst case.  It is not a fallback for anything: the CUDA path
// produces D, this consumes it.
#include "ume_common.cuh"

#include <limits>
#include <vector>

namespace ume {
namespace {

// rows <= cols.  cost(i,j) = c[i*ld_i + j*ld_j].  col4row[i] = column assigned to row i.
bool solve_lap(const float* c, int nr, int nc, size_t ld_i, size_t ld_j, std::vector<int>& col4row) {
    const double inf = std::numeric_limits<double>::infinity();
    std::vector<double> u(nr, 0.0), v(nc, 0.0), shortest(nc);
    std::vector<int> row4col(nc, -1), path(nc), remaining(nc);
    std::vector<char> in_sr(nr), in_sc(nc);
    col4row.assign(nr, -1);
    for (int cur = 0; cur < nr; ++cur) {
        std::fill(shortest.begin(), shortest.end(), inf);
        std::fill(path.begin(), path.end(), -1);
        std::fill(in_sr.begin(), in_sr.end(), 0);
        std::fill(in_sc.begin(), in_sc.end(), 0);
        // columns in reverse order, so that among equal reduced costs the LOWEST column is met last
        // and wins the `<=`-free comparison below the way scipy's implementation does
        int n_rem = nc;
        for (int j = 0; j < nc; ++j) remaining[j] = nc - 1 - j;
        int sink = -1, i = cur;
        double min_val = 0.0;
        while (sink < 0) {
            in_sr[i] = 1;
            double lowest = inf;
            int at = -1;
            const float* ci = c + (size_t)i * ld_i;
            for (int k = 0; k < n_rem; ++k) {
                const int j = remaining[k];
                const double r = min_val + (double)ci[(size_t)j * ld_j] - u[i] - v[j];
                if (r < shortest[j]) { shortest[j] = r; path[j] = i; }
                // prefer an unassigned column among ties: the path ends sooner
                if (shortest[j] < lowest || (shortest[j] == lowest && row4col[j] < 0)) { lowest = shortest[j]; at = k; }
            }
            min_val = lowest;
            if (!(min_val < inf)) return false;                   // no finite-cost completion
            const int j = remaining[at];
            if (row4col[j] < 0) sink = j;
            else i = row4col[j];
            in_sc[j] = 1;
            remaining[at] = remaining[--n_rem];
        }
        // dual update
        u[cur] += min_val;
        for (int r = 0; r < nr; ++r)
            if (in_sr[r] && r != cur) u[r] += min_val - shortest[col4row[r]];
        for (int j = 0; j < nc; ++j)
            if (in_sc[j]) v[j] -= min_val - shortest[j];
        // augment along the alternating path back to `cur`
        int j = sink;
        for (;;) {
            const int r = path[j];
            row4col[j] = r;
            std::swap(col4row[r], j);
            if (r == cur) break;
        }
    }
    return true;
}

}  // namespace
}  // namespace ume

extern "C" int ume_linear_sum_assignment_host_f32(const float* cost_host, int n_rows, int n_cols, int64_t* row_ind_host,
                                                  int64_t* col_ind_host) {
    using namespace ume;
    UME_REQUIRE(n_rows >= 0 && n_cols >= 0, UME_ERR_BAD_ARG, "ume_linear_sum_assignment_host_f32: negative size");
    const int k = n_rows < n_cols ? n_rows : n_cols;
    if (k == 0) return UME_OK;
    UME_REQUIRE(cost_host && row_ind_host && col_ind_host, UME_ERR_BAD_ARG, "ume_linear_sum_assignment_host_f32: null pointer");
    for (size_t t = 0; t < (size_t)n_rows * n_cols; ++t)
        UME_REQUIRE(cost_host[t] == cost_host[t] && cost_host[t] != -std::numeric_limits<float>::infinity(), UME_ERR_BAD_ARG,
                    "ume_linear_sum_assignment_host_f32: cost matrix holds NaN or -inf");
    std::vector<int> a;
    if (n_rows <= n_cols) {
        UME_REQUIRE(solve_lap(cost_host, n_rows, n_cols, (size_t)n_cols, 1, a), UME_ERR_BAD_ARG,
                    "ume_linear_sum_assignment_host_f32: cost matrix is infeasible");
        for (int i = 0; i < n_rows; ++i) { row_ind_host[i] = i; col_ind_host[i] = a[i]; }
    } else {
        // more rows than columns: solve the transpose, then list the pairs by ascending row
        UME_REQUIRE(solve_lap(cost_host, n_cols, n_rows, 1, (size_t)n_cols, a), UME_ERR_BAD_ARG,
                    "ume_linear_sum_assignment_host_f32: cost matrix is infeasible");
        std::vector<int> col_of_row(n_rows, -1);
        for (int j = 0; j < n_cols; ++j) col_of_row[a[j]] = j;
        int t = 0;
        for (int i = 0; i < n_rows; ++i)
            if (col_of_row[i] >= 0) { row_ind_host[t] = i; col_ind_host[t] = col_of_row[i]; ++t; }
    }
    return UME_OK;
}
