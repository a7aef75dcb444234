// Small HBM-bound helper kernels: TOM connectivity / diagonal, sparse CSR
// densification and the sparse-Gram -> Pearson finaliser.
#pragma once

#include "common.cuh"

namespace bixcor {

// k_i = sum_j |a_ij| (diagonal included, as the 4x4 known-answer test of
// inst/tinytest/test_cor_modules.R:9-46 requires); unsigned TOM works on |A|
// (calculate_tom takes abs() in R first, R/functions_stats.R:52-54), so the
// matrix is replaced by its absolute value in place.  One CTA per column.
__global__ void __launch_bounds__(256)
tom_connectivity_kernel(double* __restrict__ a, int64_t p, int64_t col0, int make_abs, double* __restrict__ kvec) {
  __shared__ double scr[32];
  const int64_t i = col0 + blockIdx.x;
  double* col = a + i * p;
  double s = 0.0;
  for (int64_t j = threadIdx.x; j < p; j += blockDim.x) {
    const double v = fabs(col[j]);
    s += v;
    if (make_abs) col[j] = v;
  }
  s = block_sum(s, scr);
  if (threadIdx.x == 0) kvec[i] = s;
}

__global__ void extract_diag_kernel(const double* __restrict__ a, int64_t p, double* __restrict__ d) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < p) d[i] = a[i * p + i];
}

// rs_cov2cor (src/rust/src/base/r_cors_similarity.rs:135-142): r_ij = c_ij / sqrt(c_ii c_jj), column-major.
__global__ void __launch_bounds__(256)
cov2cor_kernel(const double* __restrict__ cov, int64_t p, double* __restrict__ out) {
  const int64_t j = blockIdx.y;
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= p) return;
  out[i + j * p] = cov[i + j * p] / sqrt(cov[i + i * p] * cov[j + j * p]);
}

__global__ void fill_kernel(double* __restrict__ x, int64_t n, double v) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) x[i] = v;
}

// Scatter the cells [c0, c0+ncell) of a cells x genes CSR matrix into a dense
// K-major panel D[gene][cell - c0] (ld = ldk doubles, pre-zeroed).  One warp per cell.
__global__ void __launch_bounds__(256)
csr_densify_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                   const float* __restrict__ data, int64_t c0, int ncell, int64_t genes, double* __restrict__ panel,
                   int64_t ldk) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= ncell) return;
  const int64_t b = indptr[c0 + warp], e = indptr[c0 + warp + 1];
  for (int64_t q = b + lane; q < e; q += 32) {
    const int64_t g = indices[q];
    if (g >= 0 && g < genes) panel[g * ldk + warp] = (double)data[q];
  }
}

// colsum[g] += sum over the panel row g.  One CTA per gene.
__global__ void __launch_bounds__(128)
panel_rowsum_kernel(const double* __restrict__ panel, int64_t ldk, int ncell, double* __restrict__ colsum) {
  __shared__ double scr[32];
  const int64_t g = blockIdx.x;
  double s = 0.0;
  for (int i = threadIdx.x; i < ncell; i += blockDim.x) s += panel[g * ldk + i];
  s = block_sum(s, scr);
  if (threadIdx.x == 0) colsum[g] += s;
}

// Pearson from the raw Gram G = X'X and column sums s over N cells, packed shift = 1:
// r_ij = (G_ij - s_i s_j / N) / sqrt(var_i var_j), var_i = G_ii - s_i^2 / N.
__global__ void __launch_bounds__(256)
sparse_finalise_kernel(const double* __restrict__ gram, const double* __restrict__ colsum, double n_cells,
                       int64_t genes, double* __restrict__ out_packed) {
  const int64_t i = blockIdx.y;
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j <= i || j >= genes) return;
  const double si = colsum[i], sj = colsum[j];
  const double vi = gram[i * genes + i] - si * si / n_cells;
  const double vj = gram[j * genes + j] - sj * sj / n_cells;
  const double cov = gram[i * genes + j] - si * sj / n_cells;
  out_packed[packed_row_offset(i, genes, 1) + (j - i - 1)] = cov / sqrt(vi * vj);
}

}  // namespace bixcor
