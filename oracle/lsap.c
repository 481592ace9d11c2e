/* ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the rectangular linear-sum-assignment solver the reference calls at
 * reference src/matcher.py:136 (`scipy.optimize.linear_sum_assignment`).  The solver itself is a
 * third-party dependency that is NOT under /root/reference: SciPy (unpinned in the reference's
 * requirements.txt:2; scipy 1.18.1 is installed in this image, only as a compiled .so).  Its
 * published algorithm is the shortest-augmenting-path method of D. F. Crouse, "On implementing 2D
 * rectangular assignment algorithms", IEEE T-AES 52(4), 2016, as implemented in SciPy's
 * `rectangular_lsap`:
 *   - costs are float64; if there are more rows than columns the problem is transposed;
 *   - for every row, a Dijkstra-like scan over the not-yet-scanned columns, kept in an array
 *     `remaining` that is initialised in DESCENDING column order and shrunk by swap-with-last;
 *   - the next column is the one with the smallest reduced path cost; ties are broken toward a
 *     column that is still unassigned (which ends the search), otherwise the first one met;
 *   - dual variables u, v are updated, then the assignment is augmented along `path`;
 *   - the result is returned sorted by original row index.
 * Parity pin: tests/test_oracle_lsap.py compares this file against the installed SciPy on random
 * rectangular float matrices, on integer matrices full of ties, and on the known answers listed in
 * SURVEY.md §8(a.1) (the reference itself ships no tests or golden vectors).
 *
 * Build: oracle/build_oracle.sh (gcc -O2 -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* Solve one nr x nc problem with nr <= nc on a row-major float64 cost.  col4row[nr] out. */
static int lsap_core(int nr, int nc, const double *cost, int64_t *col4row, int64_t *row4col,
                     double *u, double *v, double *spc, int64_t *path, int64_t *remaining,
                     unsigned char *SR, unsigned char *SC)
{
    for (int i = 0; i < nr; i++) { u[i] = 0.0; col4row[i] = -1; }
    for (int j = 0; j < nc; j++) { v[j] = 0.0; row4col[j] = -1; path[j] = -1; }

    for (int cur = 0; cur < nr; cur++) {
        /* --- augmenting path search from row `cur` --- */
        double min_val = 0.0;
        int64_t i = cur;
        int num_remaining = nc;
        for (int it = 0; it < nc; it++) remaining[it] = nc - it - 1;
        memset(SR, 0, (size_t)nr);
        memset(SC, 0, (size_t)nc);
        for (int j = 0; j < nc; j++) spc[j] = INFINITY;

        int64_t sink = -1;
        while (sink == -1) {
            int index = -1;
            double lowest = INFINITY;
            SR[i] = 1;
            for (int it = 0; it < num_remaining; it++) {
                int64_t j = remaining[it];
                double r = min_val + cost[i * (int64_t)nc + j] - u[i] - v[j];
                if (r < spc[j]) { path[j] = i; spc[j] = r; }
                if (spc[j] < lowest || (spc[j] == lowest && row4col[j] == -1)) {
                    lowest = spc[j];
                    index = it;
                }
            }
            min_val = lowest;
            if (min_val == INFINITY) return -1; /* infeasible */
            int64_t j = remaining[index];
            if (row4col[j] == -1) sink = j; else i = row4col[j];
            SC[j] = 1;
            remaining[index] = remaining[--num_remaining];
        }

        /* --- dual update --- */
        u[cur] += min_val;
        for (int r = 0; r < nr; r++)
            if (SR[r] && r != cur) u[r] += min_val - spc[col4row[r]];
        for (int j = 0; j < nc; j++)
            if (SC[j]) v[j] -= min_val - spc[j];

        /* --- augment --- */
        int64_t j = sink;
        for (;;) {
            int64_t r = path[j];
            row4col[j] = r;
            int64_t t = col4row[r]; col4row[r] = j; j = t;
            if (r == cur) break;
        }
    }
    return 0;
}

/* cost32: row-major float32 [n_rows, n_cols] (the reference feeds float32, SciPy casts to float64).
 * Writes min(n_rows, n_cols) pairs (row_ind sorted ascending, col_ind).  Returns the pair count,
 * or -1 if infeasible / out of memory. */
int lsap_oracle_f32(int n_rows, int n_cols, const float *cost32, int64_t *row_ind, int64_t *col_ind)
{
    const int transpose = n_cols < n_rows;
    const int nr = transpose ? n_cols : n_rows;
    const int nc = transpose ? n_rows : n_cols;
    if (n_rows < 0 || n_cols < 0) return -1;
    if (nr == 0 || nc == 0) return 0;

    double *cost = (double *)malloc(sizeof(double) * (size_t)nr * nc);
    double *u = (double *)malloc(sizeof(double) * nr), *v = (double *)malloc(sizeof(double) * nc);
    double *spc = (double *)malloc(sizeof(double) * nc);
    int64_t *col4row = (int64_t *)malloc(sizeof(int64_t) * nr);
    int64_t *row4col = (int64_t *)malloc(sizeof(int64_t) * nc);
    int64_t *path = (int64_t *)malloc(sizeof(int64_t) * nc);
    int64_t *remaining = (int64_t *)malloc(sizeof(int64_t) * nc);
    unsigned char *SR = (unsigned char *)malloc((size_t)nr), *SC = (unsigned char *)malloc((size_t)nc);
    int rc = -1;
    if (cost && u && v && spc && col4row && row4col && path && remaining && SR && SC) {
        if (transpose) {
            for (int i = 0; i < nr; i++)
                for (int j = 0; j < nc; j++) cost[(size_t)i * nc + j] = (double)cost32[(size_t)j * n_cols + i];
        } else {
            for (size_t k = 0; k < (size_t)nr * nc; k++) cost[k] = (double)cost32[k];
        }
        rc = lsap_core(nr, nc, cost, col4row, row4col, u, v, spc, path, remaining, SR, SC);
        if (rc == 0) {
            if (transpose) {
                /* rows of the original problem are our columns: emit sorted by original row */
                int k = 0;
                for (int j = 0; j < nc; j++)
                    if (row4col[j] != -1) { row_ind[k] = j; col_ind[k] = row4col[j]; k++; }
                rc = k;
            } else {
                for (int i = 0; i < nr; i++) { row_ind[i] = i; col_ind[i] = col4row[i]; }
                rc = nr;
            }
        }
    }
    free(cost); free(u); free(v); free(spc); free(col4row); free(row4col); free(path);
    free(remaining); free(SR); free(SC);
    return rc;
}
