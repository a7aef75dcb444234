// Small HBM-bound helper kernels: TOM connectivity / diagonal, sparse CSR
// densification and the sparse-Gram -> Pearson finaliser.
#pragma once

#include "common.cuh"
#include "epilogue.cuh"

namespace bixcor {

// Connectivity of calc_tom (crate methods::coremo, call site src/rust/src/methods/r_coremo.rs:46-59; formulas
// R/functions_stats.R:15-37), diagonal included as the 4x4 known-answer test of
// inst/tinytest/test_cor_modules.R:9-46 requires.  signed: k_i = sum_j |a_ij|; unsigned: k_i = sum_j a_ij - the
// matrix is used AS PASSED: abs() belongs to the R wrappers calculate_tom / calculate_tom_from_exp
// (R/functions_stats.R:52-54,114-116), cor_module_tom (R/methods_coexp_cor.R:151) hands rs_tom the raw correlation.
// One CTA per column.
__global__ void __launch_bounds__(256)
tom_connectivity_kernel(const double* __restrict__ a, int64_t p, int64_t col0, int take_abs, double* __restrict__ kvec) {
  __shared__ double scr[32];
  const int64_t i = col0 + blockIdx.x;
  const double* col = a + i * p;
  double s = 0.0;
  for (int64_t j = threadIdx.x; j < p; j += blockDim.x) {
    const double v = col[j];
    s += take_abs ? fabs(v) : v;
  }
  s = block_sum(s, scr);
  if (threadIdx.x == 0) kvec[i] = s;
}

// diagonal of A from its hi / lo operand planes (records [K hi][K lo] per gene)
__global__ void extract_diag_planes_kernel(const void* __restrict__ planes, int64_t ld, int K, int half, int64_t p,
                                           double* __restrict__ d) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= p) return;
  if (half) {
    const __half* r = reinterpret_cast<const __half*>(planes) + i * ld;
    d[i] = ((double)__half2float(r[i]) + (double)__half2float(r[K + i])) * (1.0 / kF16Scale);
  } else {
    const float* r = reinterpret_cast<const float*>(planes) + i * ld;
    d[i] = (double)r[i] + (double)r[K + i];
  }
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

// rs_rbf_function / rs_rbf_function_mat (src/rust/src/base/r_rbf.rs:36-83): elementwise phi(d).
__global__ void __launch_bounds__(256)
rbf_elementwise_kernel(const double* __restrict__ d, int64_t len, int mode, double eps, double* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < len; i += stride)
    out[i] = rbf_apply(mode, eps, d[i]);
}

// rs_rbf_iterate_epsilons (r_rbf.rs:118-131): for every epsilon the column sums of the symmetric affinity
// matrix rebuilt from the packed distance vector (diagonal := 1 when shift, as rs_upper_triangle_to_dense does).
// Two deterministic passes over the packed vector, all epsilons at once (NE_MAX accumulators per thread):
//   rows:    k[i] += sum_{j > i} phi(d_ij)   one warp per row, contiguous run
//   columns: k[j] += sum_{i < j} phi(d_ij)   one thread per column, consecutive threads read consecutive words,
//            rows split into chunks whose partial sums are reduced in a fixed order by the caller.
constexpr int kRbfMaxEps = 32;
struct RbfEps {
  int n;
  double eps[kRbfMaxEps];
};

__global__ void __launch_bounds__(256)
rbf_rowsum_kernel(const double* __restrict__ packed, int64_t p, int shift, int mode, RbfEps E, double* __restrict__ k) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * 8 + warp;
  if (i >= p) return;
  const double* row = packed + packed_row_offset(i, p, shift);
  const int64_t len = p - i - shift;  // elements (i, j) for j >= i + shift
  double acc[kRbfMaxEps];
#pragma unroll
  for (int e = 0; e < kRbfMaxEps; ++e) acc[e] = 0.0;
  for (int64_t q = lane; q < len; q += 32) {
    const double d = row[q];
    const bool diag = (shift == 0 && q == 0);  // the stored diagonal is counted once, by the column pass
#pragma unroll
    for (int e = 0; e < kRbfMaxEps; ++e)
      if (e < E.n && !diag) acc[e] += rbf_apply(mode, E.eps[e], d);
  }
#pragma unroll
  for (int e = 0; e < kRbfMaxEps; ++e) {
    if (e < E.n) {
      const double s = warp_sum(acc[e]);
      if (lane == 0) k[(int64_t)e * p + i] = s + (shift ? 1.0 : 0.0);
    }
  }
}

// partial[(chunk * ne + e) * p + j] = sum over rows i in [chunk*rows_per_chunk, ...) with i < j (i <= j if !shift)
__global__ void __launch_bounds__(128)
rbf_colsum_kernel(const double* __restrict__ packed, int64_t p, int shift, int mode, RbfEps E, int rows_per_chunk,
                  double* __restrict__ partial) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int chunk = blockIdx.y;
  const int64_t i0 = (int64_t)chunk * rows_per_chunk;
  if (j >= p) return;
  const int64_t i1 = min((int64_t)(i0 + rows_per_chunk), shift ? j : j + 1);
  double acc[kRbfMaxEps];
#pragma unroll
  for (int e = 0; e < kRbfMaxEps; ++e) acc[e] = 0.0;
  for (int64_t i = i0; i < i1; ++i) {
    const double d = packed[packed_row_offset(i, p, shift) + (j - i - shift)];
#pragma unroll
    for (int e = 0; e < kRbfMaxEps; ++e)
      if (e < E.n) acc[e] += rbf_apply(mode, E.eps[e], d);
  }
#pragma unroll
  for (int e = 0; e < kRbfMaxEps; ++e)
    if (e < E.n) partial[((int64_t)chunk * E.n + e) * p + j] = acc[e];
}

// k[e][j] += sum_chunk partial[chunk][e][j] in chunk order (deterministic)
__global__ void rbf_reduce_kernel(const double* __restrict__ partial, int64_t p, int ne, int nchunks, double* __restrict__ k) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int e = blockIdx.y;
  if (j >= p) return;
  double s = k[(int64_t)e * p + j];
  for (int c = 0; c < nchunks; ++c) s += partial[((int64_t)c * ne + e) * p + j];
  k[(int64_t)e * p + j] = s;
}

// rs_coremo_stability: drop sample `s` from every gene's column (column-major n x p -> (n-1) x p).
__global__ void __launch_bounds__(128)
remove_row_kernel(const double* __restrict__ x, int64_t n, int64_t s, double* __restrict__ out) {
  const int64_t j = blockIdx.x;
  for (int64_t i = threadIdx.x; i < n - 1; i += blockDim.x) out[j * (n - 1) + i] = x[j * n + (i < s ? i : i + 1)];
}

// Pair-list path (rs_pairwise_gene_cors_mc, src/rust/src/meta_cell/r_mc_processing.rs:258-283): only the genes
// that occur in the pair lists are densified, gene-major (slot s holds its `cells` values contiguously).
// slot_of[g] = dense slot of gene g or -1.  CSR (rows = cells): one warp per cell.
__global__ void __launch_bounds__(256)
csr_gather_genes_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                        const float* __restrict__ data, int64_t cells, int64_t genes,
                        const int32_t* __restrict__ slot_of, const int32_t* __restrict__ cell_pos,
                        double* __restrict__ dense, int64_t ld) {
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= cells) return;
  const int64_t dst = cell_pos ? cell_pos[warp] : warp;  // cells_to_keep: position among the kept cells, -1 = dropped
  if (dst < 0) return;
  const int64_t b = indptr[warp], e = indptr[warp + 1];
  for (int64_t q = b + lane; q < e; q += 32) {
    const int64_t g = indices[q];
    if (g < 0 || g >= genes) continue;
    const int s = slot_of[g];
    if (s >= 0) dense[(int64_t)s * ld + dst] = (double)data[q];
  }
}
// CSC (columns = genes): one CTA per used gene (gene_of_slot[s]).
__global__ void __launch_bounds__(256)
csc_gather_genes_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                        const float* __restrict__ data, int64_t cells, const int32_t* __restrict__ gene_of_slot,
                        const int32_t* __restrict__ cell_pos, double* __restrict__ dense, int64_t ld) {
  const int64_t s = blockIdx.x;
  const int64_t g = gene_of_slot[s];
  const int64_t b = indptr[g], e = indptr[g + 1];
  for (int64_t q = b + threadIdx.x; q < e; q += blockDim.x) {
    const int64_t cidx = indices[q];
    if (cidx < 0 || cidx >= cells) continue;
    const int64_t dst = cell_pos ? cell_pos[cidx] : cidx;
    if (dst >= 0) dense[s * ld + dst] = (double)data[q];
  }
}
// out[k] = <Z[slot1[k]], Z[slot2[k]]> over K (unit-norm centred columns -> the correlation).  One warp per pair.
__global__ void __launch_bounds__(256)
pair_dot_kernel(const double* __restrict__ z, int64_t K, const int32_t* __restrict__ s1, const int32_t* __restrict__ s2,
                int64_t npairs, double* __restrict__ out) {
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= npairs) return;
  const double2* a = reinterpret_cast<const double2*>(z + (int64_t)s1[warp] * K);
  const double2* b = reinterpret_cast<const double2*>(z + (int64_t)s2[warp] * K);
  double acc0 = 0.0, acc1 = 0.0;
  for (int64_t i = lane; i < K / 2; i += 32) {
    const double2 u = a[i], v = b[i];
    acc0 = fma(u.x, v.x, acc0);
    acc1 = fma(u.y, v.y, acc1);
  }
  const double sum = warp_sum(acc0 + acc1);
  if (lane == 0) out[warp] = sum;
}

// 1/|d| of every gene straight from its int8 digit record [K h][K l] (d = 16h + l, exact in int64): lets a
// multi-process driver all-gather only the operand records and rebuild the norms locally.  One warp per gene.
__global__ void __launch_bounds__(256)
inv_norm_from_digits_kernel(const int8_t* __restrict__ rec, int64_t K, int64_t p, double* __restrict__ inv) {
  const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (g >= p) return;
  const int* h = reinterpret_cast<const int*>(rec + g * 2 * K);
  const int* l = reinterpret_cast<const int*>(rec + g * 2 * K + K);
  long long ss = 0;
  for (int64_t i = lane; i < K / 4; i += 32) {
    const int hw = h[i], lw = l[i];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int d = 16 * (int)(int8_t)(hw >> (8 * b)) + (int)(int8_t)(lw >> (8 * b));
      ss += (long long)d * d;
    }
  }
  ss = warp_sum_ll(ss);
  if (lane == 0) inv[g] = 1.0 / sqrt((double)ss);
}

// Fisher-z step of the differential correlation (calculate_diff_correlation::<f64>, src/rust/src/methods/r_diffcor.rs:65-73)
// over a packed range: z = (atanh r_a - atanh r_b) * inv_se, p = erfc(|z| / sqrt 2).  A full-occupancy streaming kernel:
// the transcendental chains are latency bound inside the 8 epilogue warps of a Gram CTA (measured 10-18 ms fused
// against ~4 ms split at 20000 genes), so the Grams write r_a / r_b through the lean packed epilogue and this kernel
// finishes.
__global__ void __launch_bounds__(256)
diffcor_zp_kernel(const double* __restrict__ ra, const double* __restrict__ rb, int64_t len, double inv_se,
                  double* __restrict__ z, double* __restrict__ pv) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < len; i += stride) {
    // atanh a - atanh b = log(((1 + a)(1 - b)) / ((1 - a)(1 + b))) / 2: one log and one division instead of two atanh
    // (same +-inf / NaN outcomes at |r| = 1)
    const double a = ra[i], b = rb[i];
    const double zz = 0.5 * log(((1.0 + a) * (1.0 - b)) / ((1.0 - a) * (1.0 + b))) * inv_se;
    z[i] = zz;
    pv[i] = erfc(fabs(zz) * 0.70710678118654752440);
  }
}

// cor_module_tom (R/methods_coexp_cor.R:118-164) unpacks the stored triangle before rs_tom
// (rs_upper_triangle_to_dense, src/rust/src/base/r_helpers.rs:75-94: diagonal := 1 when shift, mirror).  On the
// device: one CTA per row copies the contiguous packed run (upper part + diagonal), then a tiled transpose fills
// the lower part with coalesced reads and writes.
__global__ void __launch_bounds__(256)
unpack_upper_rows_kernel(const double* __restrict__ packed, int64_t p, int shift, double* __restrict__ dense) {
  const int64_t i = blockIdx.x;
  const double* row = packed + packed_row_offset(i, p, shift);
  double* dst = dense + i * p;
  if (shift && threadIdx.x == 0) dst[i] = 1.0;
  for (int64_t j = i + shift + threadIdx.x; j < p; j += blockDim.x) dst[j] = row[j - i - shift];
}
__global__ void __launch_bounds__(256)
mirror_upper_kernel(double* __restrict__ dense, int64_t p) {  // dense[j][i] = dense[i][j] for j > i, 32 x 32 tiles
  __shared__ double t[32][33];
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bj < bi) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8 threads
  for (int r = ty; r < 32; r += 8) {
    const int64_t i = (int64_t)bi * 32 + r, j = (int64_t)bj * 32 + tx;
    t[r][tx] = (i < p && j < p) ? dense[i * p + j] : 0.0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int64_t j = (int64_t)bj * 32 + r, i = (int64_t)bi * 32 + tx;  // write row j, column i
    if (j < p && i < p && j > i) dense[j * p + i] = t[tx][r];
  }
}

// binary32 transport of a <= 1e-5-class result (3xTF32 / 3xF16 kinds): finished values are rounded to float before they
// cross PCIe (half the bytes) and the host staging threads widen them into the caller's f64 buffer.  A streaming pass
// over one band: 8 bytes read + 4 written per element, grid-stride, fully coalesced.
__global__ void __launch_bounds__(256)
narrow_f64_f32_kernel(const double* __restrict__ src, float* __restrict__ dst, int64_t count) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += stride) dst[t] = (float)src[t];
}

// Reciprocal best hits on a correlation block (rs_rbh_cor, src/rust/src/methods/r_rbh.rs:187-272).  sim is the
// column-major p x q matrix of cor2; one warp per line (a row i: stride p over j; a column j: contiguous) finds the
// k-th largest |r| of the line (with multiplicity; NaN never counts; fewer than k values: -inf).
__global__ void __launch_bounds__(256)
rbh_kth_best_kernel(const double* __restrict__ sim, int64_t p, int64_t q, int k, double* __restrict__ row_thr,
                    double* __restrict__ col_thr) {
  const int64_t line = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (line >= p + q) return;
  const bool is_row = line < p;
  const double* base = is_row ? sim + line : sim + (line - p) * p;
  const int64_t len = is_row ? q : p, stride = is_row ? p : 1;
  double prev = INFINITY, thr = -INFINITY;
  int remaining = k;
  bool first = true;
  while (remaining > 0) {
    double m = -INFINITY;
    int cnt = 0;
    for (int64_t e = lane; e < len; e += 32) {
      const double v = fabs(base[e * stride]);
      if (v == v && (first || v < prev)) {
        if (v > m) {
          m = v;
          cnt = 1;
        } else if (v == m) {
          ++cnt;
        }
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const double om = __shfl_xor_sync(0xffffffffu, m, o);
      const int oc = __shfl_xor_sync(0xffffffffu, cnt, o);
      if (om > m) {
        m = om;
        cnt = oc;
      } else if (om == m) {
        cnt += oc;
      }
    }
    if (cnt == 0) break;  // line exhausted before the k-th value
    if (cnt >= remaining) {
      thr = m;
      break;
    }
    remaining -= cnt;
    prev = m;
    first = false;
  }
  if (lane == 0) (is_row ? row_thr[line] : col_thr[line - p]) = thr;
}
// sim := |sim|; hit[i + j p] = |r_ij| is among the k best of row i AND of column j AND exceeds min_similarity
__global__ void __launch_bounds__(256)
rbh_mark_kernel(double* __restrict__ sim, int64_t p, int64_t q, const double* __restrict__ row_thr,
                const double* __restrict__ col_thr, double min_similarity, unsigned char* __restrict__ hit) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= p * q) return;
  const int64_t i = e % p, j = e / p;
  const double v = fabs(sim[e]);
  sim[e] = v;
  hit[e] = (v >= row_thr[i] && v >= col_thr[j] && v > min_similarity) ? 1 : 0;
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

// Same scatter into the TF32 hi / lo planes of the fast path (gene-major record [K hi][K lo], pre-zeroed):
// hi = tf32(v), lo = tf32(v - hi).  The column sums are accumulated from the exact f32 values (one atomic per
// nonzero would be order dependent, so they are taken from the planes by panel_rowsum_tf32_kernel instead).
__global__ void __launch_bounds__(256)
csr_densify_tf32_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                        const float* __restrict__ data, int64_t c0, int ncell, int64_t genes, float* __restrict__ planes,
                        int64_t K) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= ncell) return;
  const int64_t b = indptr[c0 + warp], e = indptr[c0 + warp + 1];
  for (int64_t q = b + lane; q < e; q += 32) {
    const int64_t g = indices[q];
    if (g < 0 || g >= genes) continue;
    const float v = data[q];
    const float hi = to_tf32(v);
    planes[g * 2 * K + warp] = hi;
    planes[g * 2 * K + K + warp] = to_tf32(v - hi);
  }
}

// binary16 form of the panel (BIXCOR_PREC_F16X3): hi = f16(2^10 v), lo = f16(2^10 v - hi); |v| >= 63 does not fit
// and raises *overflow (the caller reports it instead of returning infinities).
__global__ void __launch_bounds__(256)
csr_densify_f16_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                       const float* __restrict__ data, int64_t c0, int ncell, int64_t genes, __half* __restrict__ planes,
                       int64_t K, int* __restrict__ overflow) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= ncell) return;
  const int64_t b = indptr[c0 + warp], e = indptr[c0 + warp + 1];
  for (int64_t q = b + lane; q < e; q += 32) {
    const int64_t g = indices[q];
    if (g < 0 || g >= genes) continue;
    const double v = (double)data[q];
    if (fabs(v) >= 63.0) atomicExch(overflow, 1);
    __half hi, lo;
    split_f16(v, hi, lo);
    planes[g * 2 * K + warp] = hi;
    planes[g * 2 * K + K + warp] = lo;
  }
}
__global__ void __launch_bounds__(128)
panel_rowsum_f16_kernel(const __half* __restrict__ planes, int64_t K, int ncell, double* __restrict__ colsum) {
  __shared__ double scr[32];
  const int64_t g = blockIdx.x;
  double s = 0.0;
  for (int i = threadIdx.x; i < ncell; i += blockDim.x)
    s += (double)__half2float(planes[g * 2 * K + i]) + (double)__half2float(planes[g * 2 * K + K + i]);
  s = block_sum(s, scr);
  if (threadIdx.x == 0) colsum[g] += s * (1.0 / kF16Scale);
}

__global__ void __launch_bounds__(128)
panel_rowsum_tf32_kernel(const float* __restrict__ planes, int64_t K, int ncell, double* __restrict__ colsum) {
  __shared__ double scr[32];
  const int64_t g = blockIdx.x;
  double s = 0.0;
  for (int i = threadIdx.x; i < ncell; i += blockDim.x) s += (double)planes[g * 2 * K + i] + (double)planes[g * 2 * K + K + i];
  s = block_sum(s, scr);
  if (threadIdx.x == 0) colsum[g] += s;
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
