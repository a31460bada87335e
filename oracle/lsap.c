/*
 * lsap.c — plain-C restatement of scipy.optimize.linear_sum_assignment (TEST INFRASTRUCTURE).
 *
 * The reference calls scipy per image (A2/models/matcher.py:246, A1/models/matcher.py:90); scipy is a
 * third-party dependency absent from /root/reference (requirements.txt:6 lists it unpinned; this
 * image has scipy 1.18.1, which is the pin).  This file restates scipy's published algorithm — the
 * shortest-augmenting-path solver of D. F. Crouse, "On implementing 2D rectangular assignment
 * algorithms", IEEE TAES 52(4), 2016, as implemented in scipy/optimize/rectangular_lsap — including
 * its tie rule and output ordering.  tests/test_oracle_lsap.py pins it against the
 * installed scipy on float, tie-heavy integer and 1000x1000 cases.
 *
 * cost: row-major [nr, nc] doubles.  Outputs a[k], b[k] for k < min(nr,nc): row index (ascending)
 * and its assigned column.  Returns 0, or -1 if infeasible / invalid input.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static int solve_rows_le_cols(int nr, int nc, const double* cost, int64_t* col4row) {
  /* requires nr <= nc */
  double* u = (double*)calloc(nr, sizeof(double));
  double* v = (double*)calloc(nc, sizeof(double));
  double* spc = (double*)malloc(sizeof(double) * nc);
  int64_t* path = (int64_t*)malloc(sizeof(int64_t) * nc);
  int64_t* row4col = (int64_t*)malloc(sizeof(int64_t) * nc);
  int64_t* remaining = (int64_t*)malloc(sizeof(int64_t) * nc);
  char* SR = (char*)malloc(nr);
  char* SC = (char*)malloc(nc);
  int rc = 0;
  for (int i = 0; i < nr; ++i) col4row[i] = -1;
  for (int j = 0; j < nc; ++j) row4col[j] = -1;

  for (int cur = 0; cur < nr && rc == 0; ++cur) {
    double min_val = 0.0;
    int64_t i = cur;
    int num_remaining = nc;
    for (int it = 0; it < nc; ++it) remaining[it] = nc - it - 1; /* reverse fill */
    memset(SR, 0, nr);
    memset(SC, 0, nc);
    for (int j = 0; j < nc; ++j) spc[j] = INFINITY;
    int64_t sink = -1;
    while (sink == -1) {
      int64_t index = -1;
      double lowest = INFINITY;
      SR[i] = 1;
      for (int it = 0; it < num_remaining; ++it) {
        int64_t j = remaining[it];
        double r = min_val + cost[i * (int64_t)nc + j] - u[i] - v[j];
        if (r < spc[j]) {
          path[j] = i;
          spc[j] = r;
        }
        /* tie rule: prefer a column that is still unassigned */
        if (spc[j] < lowest || (spc[j] == lowest && row4col[j] == -1)) {
          lowest = spc[j];
          index = it;
        }
      }
      min_val = lowest;
      if (min_val == INFINITY) { rc = -1; break; }
      int64_t j = remaining[index];
      if (row4col[j] == -1) sink = j; else i = row4col[j];
      SC[j] = 1;
      remaining[index] = remaining[--num_remaining];
    }
    if (rc != 0) break;
    u[cur] += min_val;
    for (int r = 0; r < nr; ++r)
      if (SR[r] && r != cur) u[r] += min_val - spc[col4row[r]];
    for (int j = 0; j < nc; ++j)
      if (SC[j]) v[j] -= min_val - spc[j];
    int64_t j = sink;
    for (;;) {
      int64_t r = path[j];
      row4col[j] = r;
      int64_t t = col4row[r];
      col4row[r] = j;
      j = t;
      if (r == cur) break;
    }
  }
  free(u); free(v); free(spc); free(path); free(row4col); free(remaining); free(SR); free(SC);
  return rc;
}

int oracle_lsap(int nr, int nc, const double* cost, int64_t* a, int64_t* b) {
  if (nr <= 0 || nc <= 0) return 0;
  for (int64_t k = 0; k < (int64_t)nr * nc; ++k)
    if (isnan(cost[k]) || cost[k] == -INFINITY) return -1;
  if (nr <= nc) {
    int rc = solve_rows_le_cols(nr, nc, cost, b);
    for (int i = 0; i < nr; ++i) a[i] = i;
    return rc;
  }
  /* more rows than columns: solve the transpose, then order by row index (scipy argsort) */
  double* ct = (double*)malloc(sizeof(double) * (size_t)nr * nc);
  for (int i = 0; i < nr; ++i)
    for (int j = 0; j < nc; ++j) ct[(int64_t)j * nr + i] = cost[(int64_t)i * nc + j];
  int64_t* row4col_t = (int64_t*)malloc(sizeof(int64_t) * nc); /* for each column: its row */
  int rc = solve_rows_le_cols(nc, nr, ct, row4col_t);
  if (rc == 0) {
    /* pairs (row4col_t[j], j) sorted by row: rows are distinct, counting pass over rows */
    int64_t* colofrow = (int64_t*)malloc(sizeof(int64_t) * nr);
    for (int i = 0; i < nr; ++i) colofrow[i] = -1;
    for (int j = 0; j < nc; ++j) colofrow[row4col_t[j]] = j;
    int k = 0;
    for (int i = 0; i < nr; ++i)
      if (colofrow[i] >= 0) { a[k] = i; b[k] = colofrow[i]; ++k; }
    free(colofrow);
  }
  free(ct); free(row4col_t);
  return rc;
}
