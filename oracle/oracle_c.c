/*
 * Plain-C restatement of the bixverse co-expression hot path (CPU oracle).
 * TEST INFRASTRUCTURE ONLY - never linked into or called by the product library.
 *
 * It follows the reference's algorithm SHAPE as visible at its call sites:
 *   rank (average ties) -> standardise -> FULL p x p Gram -> scalar row-major
 *   gather of the upper triangle      src/rust/src/base/r_cors_similarity.rs:251-268
 *   two of those + Fisher z           src/rust/src/methods/r_diffcor.rs:62-66
 *   dense A*A + normalise             src/rust/src/methods/r_coremo.rs:54,
 *                                     formulas R/functions_stats.R:15-37
 * The arithmetic itself lives in the un-vendored crate bixverse-rs 0.4.4
 * (src/rust/Cargo.lock:173-176); parity pins: see oracle/oracle.py header.
 * PARITY UNPINNED for the Fisher-z constants, ties/NaN and the TOM diagonal.
 *
 * Build: make -C oracle   (gcc -O2 -shared)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double v; int64_t i; } pair_t;

static int cmp_pair(const void* a, const void* b) {
  double x = ((const pair_t*)a)->v, y = ((const pair_t*)b)->v;
  return (x > y) - (x < y);
}

/* 1-based average-tie ranks of one column; any NaN -> all NaN */
void orc_rank_average(const double* v, int64_t n, double* out) {
  pair_t* s = (pair_t*)malloc(sizeof(pair_t) * (size_t)(n > 0 ? n : 1));
  int has_nan = 0;
  for (int64_t i = 0; i < n; ++i) { s[i].v = v[i] + 0.0; s[i].i = i; if (v[i] != v[i]) has_nan = 1; }
  if (has_nan) { for (int64_t i = 0; i < n; ++i) out[i] = NAN; free(s); return; }
  qsort(s, (size_t)n, sizeof(pair_t), cmp_pair);
  int64_t a = 0;
  while (a < n) {
    int64_t b = a;
    while (b + 1 < n && s[b + 1].v == s[a].v) ++b;
    double r = 0.5 * (double)(a + b) + 1.0;
    for (int64_t t = a; t <= b; ++t) out[s[t].i] = r;
    a = b + 1;
  }
  free(s);
}

void orc_rank_matrix_col(const double* x, int64_t n, int64_t p, double* out) {
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t j = 0; j < p; ++j) orc_rank_average(x + j * n, n, out + j * n);
}

/* dense p x p correlation, column-major in/out; full square like the reference */
void orc_cor_dense(const double* x, int64_t n, int64_t p, int spearman, double* out) {
  double* z = (double*)malloc(sizeof(double) * (size_t)n * (size_t)p);
  if (spearman) orc_rank_matrix_col(x, n, p, z); else memcpy(z, x, sizeof(double) * (size_t)n * (size_t)p);
#pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < p; ++j) {
    double* c = z + j * n; double m = 0.0, ss = 0.0;
    for (int64_t i = 0; i < n; ++i) m += c[i];
    m /= (double)n;
    for (int64_t i = 0; i < n; ++i) { c[i] -= m; ss += c[i] * c[i]; }
    double nr = sqrt(ss);
    for (int64_t i = 0; i < n; ++i) c[i] /= nr;
  }
#pragma omp parallel for schedule(dynamic, 8)
  for (int64_t j = 0; j < p; ++j)
    for (int64_t i = 0; i < p; ++i) {
      const double *a = z + i * n, *b = z + j * n; double s = 0.0;
      for (int64_t k = 0; k < n; ++k) s += a[k] * b[k];
      out[i + j * p] = s;
    }
  free(z);
}

int64_t orc_packed_offset(int64_t i, int64_t p, int shift) {
  return shift ? i * (2 * p - i - 1) / 2 : i * (2 * p - i + 1) / 2;
}

/* r_helpers.rs:126-131 */
void orc_dense_to_upper(const double* m, int shift, int64_t p, double* out) {
  int64_t idx = 0;
  for (int64_t i = 0; i < p; ++i)
    for (int64_t j = i + (shift ? 1 : 0); j < p; ++j) out[idx++] = m[i + j * p];
}

/* r_helpers.rs:79-90 */
void orc_upper_to_dense(const double* v, int shift, int64_t p, double* out) {
  int64_t idx = 0;
  for (int64_t i = 0; i < p; ++i)
    for (int64_t j = i; j < p; ++j) {
      if (shift && i == j) out[i + j * p] = 1.0;
      else { out[i + j * p] = v[idx]; out[j + i * p] = v[idx]; ++idx; }
    }
}

void orc_cor_upper(const double* x, int64_t n, int64_t p, int spearman, int shift, double* out) {
  double* d = (double*)malloc(sizeof(double) * (size_t)p * (size_t)p);
  orc_cor_dense(x, n, p, spearman, d);
  orc_dense_to_upper(d, shift, p, out);
  free(d);
}

void orc_diffcor(const double* ra, const double* rb, int64_t len, int64_t na, int64_t nb, double c,
                 double* z, double* pv) {
  double se = sqrt(c / (double)(na - 3) + c / (double)(nb - 3));
#pragma omp parallel for schedule(static)
  for (int64_t t = 0; t < len; ++t) {
    z[t] = (atanh(ra[t]) - atanh(rb[t])) / se;
    pv[t] = erfc(fabs(z[t]) / sqrt(2.0));
  }
}

/* calc_tom: version 1 / 2, signed flag; input dense symmetric, column-major */
void orc_tom(const double* a_in, int64_t p, int is_signed, int version, double* out) {
  double* a = (double*)malloc(sizeof(double) * (size_t)p * (size_t)p);
  double* k = (double*)malloc(sizeof(double) * (size_t)p);
  for (int64_t t = 0; t < p * p; ++t) a[t] = is_signed ? a_in[t] : fabs(a_in[t]);
  for (int64_t i = 0; i < p; ++i) { double s = 0.0; for (int64_t j = 0; j < p; ++j) s += fabs(a[i + j * p]); k[i] = s; }
#pragma omp parallel for schedule(dynamic, 8)
  for (int64_t j = 0; j < p; ++j)
    for (int64_t i = 0; i < p; ++i) {
      if (i == j) { out[i + j * p] = 1.0; continue; }
      double l = 0.0;
      for (int64_t u = 0; u < p; ++u) if (u != i && u != j) l += a[i + u * p] * a[u + j * p];
      double aij = a[i + j * p], mk = k[i] < k[j] ? k[i] : k[j];
      out[i + j * p] = version == 1 ? (aij + l) / (mk + 1.0 - fabs(aij)) : 0.5 * (aij + l / (mk + fabs(aij)));
    }
  free(a); free(k);
}

/* ---- "next" rows of the scope table ----------------------------------------------------------- */

/* cor(x_i, y_j): x is n x p, y is n x q (column-major), out p x q; src/rust/src/base/r_cors_similarity.rs:114-122 */
void orc_cor2(const double* x, int64_t n, int64_t p, const double* y, int64_t q, int spearman, double* out) {
  double* zx = (double*)malloc(sizeof(double) * (size_t)n * (size_t)p);
  double* zy = (double*)malloc(sizeof(double) * (size_t)n * (size_t)q);
  for (int side = 0; side < 2; ++side) {
    const double* src = side ? y : x;
    double* z = side ? zy : zx;
    const int64_t cols = side ? q : p;
    for (int64_t j = 0; j < cols; ++j) {
      double* col = z + j * n;
      if (spearman) orc_rank_average(src + j * n, n, col); else memcpy(col, src + j * n, sizeof(double) * (size_t)n);
      double m = 0.0, ss = 0.0;
      for (int64_t i = 0; i < n; ++i) m += col[i];
      m /= (double)n;
      for (int64_t i = 0; i < n; ++i) { col[i] -= m; ss += col[i] * col[i]; }
      ss = sqrt(ss);
      for (int64_t i = 0; i < n; ++i) col[i] /= ss;
    }
  }
  for (int64_t j = 0; j < q; ++j)
    for (int64_t i = 0; i < p; ++i) {
      double s = 0.0;
      for (int64_t t = 0; t < n; ++t) s += zx[t + i * n] * zy[t + j * n];
      out[i + j * p] = s;
    }
  free(zx); free(zy);
}

static int cmp_desc(const void* a, const void* b) {
  double x = *(const double*)a, y = *(const double*)b;
  return (x < y) - (x > y);
}
static double kth_best(const double* v, int64_t len, int64_t stride, int k) {
  double* t = (double*)malloc(sizeof(double) * (size_t)(len > 0 ? len : 1));
  int64_t m = 0;
  for (int64_t e = 0; e < len; ++e) { double a = fabs(v[e * stride]); if (a == a) t[m++] = a; }
  qsort(t, (size_t)m, sizeof(double), cmp_desc);
  double r = (m >= k) ? t[k - 1] : -INFINITY;
  free(t);
  return r;
}
/* reciprocal best hits on |cor2| (src/rust/src/methods/r_rbh.rs:187-272): hits origin-major, 0-based; returns their
 * number; at most `cap` are written */
int64_t orc_rbh_cor(const double* x, int64_t n, int64_t p, const double* y, int64_t q, int k_best, int spearman,
                    double min_similarity, int32_t* oi, int32_t* oj, double* osim, int64_t cap) {
  double* sim = (double*)malloc(sizeof(double) * (size_t)p * (size_t)q);
  double* rt = (double*)malloc(sizeof(double) * (size_t)p);
  double* ct = (double*)malloc(sizeof(double) * (size_t)q);
  orc_cor2(x, n, p, y, q, spearman, sim);
  for (int64_t i = 0; i < p; ++i) rt[i] = kth_best(sim + i, q, p, k_best);
  for (int64_t j = 0; j < q; ++j) ct[j] = kth_best(sim + j * p, p, 1, k_best);
  int64_t cnt = 0;
  for (int64_t i = 0; i < p; ++i)
    for (int64_t j = 0; j < q; ++j) {
      double v = fabs(sim[i + j * p]);
      if (v >= rt[i] && v >= ct[j] && v > min_similarity) {
        if (cnt < cap) { oi[cnt] = (int32_t)i; oj[cnt] = (int32_t)j; osim[cnt] = v; }
        ++cnt;
      }
    }
  free(sim); free(rt); free(ct);
  return cnt;
}

/* RBF kernels on distances (src/rust/src/base/r_rbf.rs:36-48): 1 gaussian, 2 bump, 3 inverse quadratic */
void orc_rbf(const double* d, int64_t len, double eps, int type, double* out) {
  for (int64_t t = 0; t < len; ++t) {
    double e2 = (eps * d[t]) * (eps * d[t]);
    if (type == 1) out[t] = exp(-e2);
    else if (type == 3) out[t] = 1.0 / (1.0 + e2);
    else out[t] = (d[t] < 1.0 / eps) ? exp(-1.0 / (1.0 - e2) + 1.0) : 0.0;
  }
}
