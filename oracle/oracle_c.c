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
  /* rs_tom level: the matrix as passed; abs() for unsigned networks is the R wrappers' job (R/functions_stats.R:52-54) */
  for (int64_t t = 0; t < p * p; ++t) a[t] = a_in[t];
  for (int64_t i = 0; i < p; ++i) { double s = 0.0; for (int64_t j = 0; j < p; ++j) s += is_signed ? fabs(a[i + j * p]) : a[i + j * p]; k[i] = s; }
#pragma omp parallel for schedule(dynamic, 8)
  for (int64_t j = 0; j < p; ++j)
    for (int64_t i = 0; i < p; ++i) {
      if (i == j) { out[i + j * p] = 1.0; continue; }
      double l = 0.0;
      for (int64_t u = 0; u < p; ++u) if (u != i && u != j) l += a[i + u * p] * a[u + j * p];
      double aij = a[i + j * p], mk = k[i] < k[j] ? k[i] : k[j];
      double den = is_signed ? fabs(aij) : aij;
      out[i + j * p] = version == 1 ? (aij + l) / (mk + 1.0 - den) : 0.5 * (aij + l / (mk + den));
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

/* ============================================================================================================
 * Timed CPU baseline ("reference arm" of bench.py): the reference's algorithm SHAPE on all host cores.
 *   rank (rayon over columns in the crate -> OpenMP here) -> standardise -> FULL p x p Gram with a threaded BLAS
 *   dgemm (faer GEMM in the crate -> OpenBLAS here, handed in as a function pointer because the only OpenBLAS in
 *   the image is the ILP64 build bundled with numpy) -> triangle gather (r_cors_similarity.rs:258-265) -> |r|^beta.
 * Deliberately a STRONG baseline: the gather is threaded and reads the symmetric matrix by contiguous columns
 * (the reference's is a single-threaded stride-p row walk), so the GPU/CPU ratio is conservative.
 * ============================================================================================================ */
#ifdef _OPENMP
#include <omp.h>
#endif
#include <time.h>

/* cblas_dgemm of an ILP64 OpenBLAS (scipy_cblas_dgemm64_): layout 102 = column major, trans 111 = N / 112 = T */
typedef void (*orc_dgemm_fn)(int layout, int transa, int transb, int64_t m, int64_t n, int64_t k, double alpha,
                             const double* a, int64_t lda, const double* b, int64_t ldb, double beta, double* c, int64_t ldc);

static double orc_now(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* bottom-up merge sort of (value, index) pairs, insertion-sorted runs of 16 (3x faster than qsort at n = 1000) */
static void sort_pairs(pair_t* s, pair_t* tmp, int64_t n) {
  for (int64_t b = 0; b < n; b += 16) {
    const int64_t e = b + 16 < n ? b + 16 : n;
    for (int64_t i = b + 1; i < e; ++i) {
      pair_t v = s[i];
      int64_t j = i;
      while (j > b && s[j - 1].v > v.v) { s[j] = s[j - 1]; --j; }
      s[j] = v;
    }
  }
  pair_t *src = s, *dst = tmp;
  for (int64_t w = 16; w < n; w <<= 1) {
    for (int64_t b = 0; b < n; b += 2 * w) {
      int64_t i = b, m = b + w < n ? b + w : n, j = m, e = b + 2 * w < n ? b + 2 * w : n, o = b;
      while (i < m && j < e) dst[o++] = (src[j].v < src[i].v) ? src[j++] : src[i++];
      while (i < m) dst[o++] = src[i++];
      while (j < e) dst[o++] = src[j++];
    }
    pair_t* t = src; src = dst; dst = t;
  }
  if (src != s) memcpy(s, src, sizeof(pair_t) * (size_t)n);
}

/* centred, unit-norm ranks (Spearman) or values (Pearson) of every column into z (n x p) */
static void orc_standardised(const double* x, int64_t n, int64_t p, int spearman, double* z) {
#pragma omp parallel
  {
    pair_t* s = (pair_t*)malloc(sizeof(pair_t) * (size_t)(2 * n + 2));
    pair_t* tmp = s + n + 1;
#pragma omp for schedule(dynamic, 32)
    for (int64_t j = 0; j < p; ++j) {
      const double* v = x + j * n;
      double* c = z + j * n;
      if (spearman) {
        int has_nan = 0;
        for (int64_t i = 0; i < n; ++i) { s[i].v = v[i] + 0.0; s[i].i = i; if (v[i] != v[i]) has_nan = 1; }
        if (has_nan) { for (int64_t i = 0; i < n; ++i) c[i] = NAN; continue; }
        sort_pairs(s, tmp, n);
        int64_t a = 0;
        while (a < n) {
          int64_t b = a;
          while (b + 1 < n && s[b + 1].v == s[a].v) ++b;
          const double r = 0.5 * (double)(a + b) + 1.0;
          for (int64_t t = a; t <= b; ++t) c[s[t].i] = r;
          a = b + 1;
        }
      } else {
        memcpy(c, v, sizeof(double) * (size_t)n);
      }
      double m = 0.0, ss = 0.0;
      for (int64_t i = 0; i < n; ++i) m += c[i];
      m /= (double)n;
      for (int64_t i = 0; i < n; ++i) { c[i] -= m; ss += c[i] * c[i]; }
      const double inv = 1.0 / sqrt(ss);
      for (int64_t i = 0; i < n; ++i) c[i] *= inv;
    }
    free(s);
  }
}

/* rs_cor_upper_triangle + |r|^beta on all host cores.  out: packed (shift) vector; secs[4]: rank+standardise, dgemm,
 * gather+power, total.  Returns 0, or -1 when a buffer cannot be allocated. */
int orc_cor_upper_pipeline(const double* x, int64_t n, int64_t p, int spearman, int shift, double beta,
                           orc_dgemm_fn dgemm, double* out, double* secs) {
  const double t0 = orc_now();
  double* z = (double*)malloc(sizeof(double) * (size_t)n * (size_t)p);
  double* c = (double*)malloc(sizeof(double) * (size_t)p * (size_t)p);
  if (!z || !c) { free(z); free(c); return -1; }
  orc_standardised(x, n, p, spearman, z);
  const double t1 = orc_now();
  dgemm(102, 112, 111, p, p, n, 1.0, z, n, z, n, 0.0, c, p); /* FULL square, as column_pairwise_cor */
  const double t2 = orc_now();
  const int s = shift ? 1 : 0;
  const int bi = (beta == floor(beta) && beta >= 1.0 && beta <= 64.0) ? (int)beta : 0;
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t i = 0; i < p; ++i) {
    double* row = out + orc_packed_offset(i, p, s);
    const double* col = c + i * p; /* symmetric: column i == row i, contiguous */
    for (int64_t j = i + s; j < p; ++j) {
      double r = col[j];
      if (beta != 1.0) {
        double a = fabs(r);
        if (bi == 6) { const double q = a * a; r = q * q * q; }
        else r = pow(a, beta);
      }
      row[j - i - s] = r;
    }
  }
  const double t3 = orc_now();
  free(z); free(c);
  if (secs) { secs[0] = t1 - t0; secs[1] = t2 - t1; secs[2] = t3 - t2; secs[3] = t3 - t0; }
  return 0;
}

/* calculate_tom_from_exp shape on all host cores (bounded-sample TOM baseline): correlation as above (dense), signed
 * |r|^beta affinity with unit diagonal, A*A by dgemm, normalisation (v1 / v2), packed shift = 1 output. */
int orc_tom_from_exp_pipeline(const double* x, int64_t n, int64_t p, int spearman, int is_signed, int version, double beta,
                              orc_dgemm_fn dgemm, double* out_packed, double* secs) {
  const double t0 = orc_now();
  double* z = (double*)malloc(sizeof(double) * (size_t)n * (size_t)p);
  double* a = (double*)malloc(sizeof(double) * (size_t)p * (size_t)p);
  double* l = (double*)malloc(sizeof(double) * (size_t)p * (size_t)p);
  double* k = (double*)malloc(sizeof(double) * (size_t)p);
  if (!z || !a || !l || !k) { free(z); free(a); free(l); free(k); return -1; }
  orc_standardised(x, n, p, spearman, z);
  dgemm(102, 112, 111, p, p, n, 1.0, z, n, z, n, 0.0, a, p);
#pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < p; ++j) {
    double s = 0.0;
    for (int64_t i = 0; i < p; ++i) {
      double r = a[i + j * p];
      if (i == j) r = 1.0;
      else if (beta != 1.0) { const double w = pow(fabs(r), beta); r = is_signed ? copysign(w, r) : w; }
      else if (!is_signed) r = fabs(r);
      a[i + j * p] = r;
      s += is_signed ? fabs(r) : r;
    }
    k[j] = s;
  }
  const double t1 = orc_now();
  dgemm(102, 111, 111, p, p, p, 1.0, a, p, a, p, 0.0, l, p);
  const double t2 = orc_now();
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t i = 0; i < p; ++i) {
    double* row = out_packed + orc_packed_offset(i, p, 1);
    for (int64_t j = i + 1; j < p; ++j) {
      const double aij = a[j + i * p], den = is_signed ? fabs(aij) : aij;
      const double sh = l[j + i * p] - aij * (a[i + i * p] + a[j + j * p]);
      const double mk = k[i] < k[j] ? k[i] : k[j];
      row[j - i - 1] = version == 1 ? (aij + sh) / (mk + 1.0 - den) : 0.5 * (aij + sh / (mk + den));
    }
  }
  const double t3 = orc_now();
  free(z); free(a); free(l); free(k);
  if (secs) { secs[0] = t1 - t0; secs[1] = t2 - t1; secs[2] = t3 - t2; secs[3] = t3 - t0; }
  return 0;
}

/* torchrun exports OMP_NUM_THREADS=1 to its children; the CPU arm is a host job and takes the cores it is given */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
