// Fused Gram epilogues shared by the FP64 (DMMA) and tcgen05 (TF32x3 / int8) kernels.
//
// Every epilogue receives one Gram value v = G[i][j] and writes it where the
// reference's result layout wants it:
//   EPI_PACKED   packed upper triangle, order of src/rust/src/base/r_helpers.rs:126-131,
//                optional soft-threshold power (north-star extension, beta == 1 is a no-op)
//   EPI_DENSE    dense column-major p x p (rs_cor, r_cors_similarity.rs:70-77), mirrored
//   EPI_DIFFCOR  Fisher-z differential correlation (calculate_diff_correlation,
//                src/rust/src/methods/r_diffcor.rs:65-73): r_a, r_b, z, two-sided p
//   EPI_TOM      topological overlap normalisation (calc_tom, r_coremo.rs:54; formulas
//                R/functions_stats.R:15-37)
//   EPI_ACCUM    G += v (sparse single-cell partial Gram)
#pragma once

#include "common.cuh"

namespace bixcor {

enum Epi : int { EPI_PACKED = 0, EPI_DENSE = 1, EPI_DIFFCOR = 2, EPI_TOM = 3, EPI_ACCUM = 4 };

struct EpiParams {
  int64_t p;
  double* out;
  int64_t out_base;  // packed: offset of out[0] (and of r_a..pv[0]) in the full packed vector
  int64_t ldo;       // dense leading dimension
  int shift;
  int mirror;        // dense: also write the transposed element
  int diag_one;      // dense: force 1.0 on the diagonal
  double scale;      // dense / packed: value multiplier applied first (1/(n-1) for covariance); 1 = none
  double beta;
  int beta_int;      // > 0: beta is this small integer (multiplication chain instead of pow)
  int keep_sign;
  int abs_out;       // beta == 1 only: write |r| (the abs() of calculate_tom_from_exp for unsigned networks, R/functions_stats.R:114-116)
  double* r_a;
  double* r_b;
  double* z;
  double* pv;
  double inv_se;
  const double* A;     // TOM: dense p x p affinity (symmetric), ld = p; nullptr: a_ij is rebuilt from the hi / lo operand planes
  const void* A_planes;  // TOM without the f64 matrix: gene-major records [K hi][K lo] (float, or __half scaled by 2^10)
  int64_t planes_ld;     // elements per record (2 K)
  int planes_K;          // padded K of a plane
  int planes_half;       // binary16 planes
  const double* kvec;  // TOM: connectivity
  const double* dvec;  // TOM: diagonal of A
  int tom_v2;
  int tom_signed;   // signed: |a_ij| in the denominators; unsigned: a_ij as passed (R/functions_stats.R:15-31)
  int tom_packed;
  int out_s32;       // EPI_PACKED on the exact int8 kind: `out` is an int32 vector in the packed layout that receives the exact
                     // integer Gram S_ij; the host staging threads turn it into r = S / (|d_i| |d_j|) (and the power) while they
                     // move it into the caller's pageable buffer - half the bytes cross PCIe (csrc/bixcor.cu, expand_packed_s32)
  int upper_tiles;   // the launch's tile list holds upper-triangle tiles only (validity rule of the multicast super-tiles)
  // RBF affinity transform (crate core::math::rbf behind src/rust/src/base/r_rbf.rs and the
  // distance -> affinity step of rs_coremo_stability, src/rust/src/methods/r_coremo.rs:211-226)
  int rbf_mode;        // 0 none, BIXCOR_RBF_GAUSSIAN / _BUMP / _INVERSE_QUADRATIC
  double rbf_eps;
  int rbf_from_cor;    // input is a correlation: distance d = 1 - |r| first
  int rbf_complement;  // write 1 - phi(d) (rs_coremo_stability returns the RBF *distance*)
};

enum RbfMode : int { RBF_NONE = 0, RBF_GAUSSIAN = 1, RBF_BUMP = 2, RBF_INVERSE_QUADRATIC = 3 };

// phi(d) for the three kernels of the reference (rs_rbf_function): gaussian exp(-(eps d)^2), bump
// exp(-1 / (1 - (eps d)^2) + 1) for d < 1/eps and 0 beyond, inverse quadratic 1 / (1 + (eps d)^2).
__device__ __forceinline__ double rbf_apply(int mode, double eps, double d) {
  const double t = eps * d;
  const double t2 = t * t;
  if (mode == RBF_GAUSSIAN) return exp(-t2);
  if (mode == RBF_BUMP) return (d < 1.0 / eps) ? exp(-1.0 / (1.0 - t2) + 1.0) : 0.0;
  return 1.0 / (1.0 + t2);
}

__device__ __forceinline__ double epi_power(const EpiParams& E, double r) {
  if (E.rbf_mode) {  // warp uniform
    const double d = E.rbf_from_cor ? 1.0 - fabs(r) : r;
    const double a = rbf_apply(E.rbf_mode, E.rbf_eps, d);
    return E.rbf_complement ? 1.0 - a : a;
  }
  if (E.beta == 1.0) return E.abs_out ? fabs(r) : r;
  double a = fabs(r);
  const int bi = E.beta_int;  // warp-uniform: the branches below never diverge
  if (bi > 0) {
    const double sq = a * a;
    if (bi == 2) {
      a = sq;
    } else if (bi == 3) {
      a = sq * a;
    } else if (bi == 4) {
      a = sq * sq;
    } else if (bi == 6) {
      a = sq * sq * sq;
    } else if (bi == 8) {
      const double q = sq * sq;
      a = q * q;
    } else {
      double acc = 1.0, base = a;
      int e = bi;
      while (e) {
        if (e & 1) acc *= base;
        base *= base;
        e >>= 1;
      }
      a = acc;
    }
  } else {
    a = pow(a, E.beta);
  }
  return E.keep_sign ? copysign(a, r) : a;
}

// Per-row constants hoisted out of the column loop.
struct EpiRow {
  int i;
  int64_t ro;  // packed: index of (i, j) is ro + j
  double k_i, d_i;
};

template <int EPI>
__device__ __forceinline__ EpiRow epi_row(const EpiParams& E, int64_t i) {
  EpiRow R;
  R.i = (int)i;
  const int shift = (EPI == EPI_DIFFCOR || (EPI == EPI_TOM)) ? 1 : E.shift;
  R.ro = packed_row_offset(i, E.p, shift) - E.out_base - i - shift;
  R.k_i = 0.0;
  R.d_i = 0.0;
  if (EPI == EPI_TOM) {
    R.k_i = E.kvec[i];
    R.d_i = E.dvec[i];
  }
  return R;
}

// pass: 0 for everything except the second Gram of EPI_DIFFCOR (pass 1).  In pass 0 of
// EPI_DIFFCOR the value is r_a and is parked in its final location.
// Returns the value written to the primary location.  MIRROR_INLINE = false leaves the
// transposed write of the dense modes to the caller (epi_mirror), so that a kernel can issue
// it from a thread mapping in which it is coalesced.
template <int EPI, bool MIRROR_INLINE = true>
__device__ __forceinline__ double epi_store(const EpiParams& E, const EpiRow& R, int j, double v, int pass) {
  const int i = R.i;
  if (EPI == EPI_PACKED) {
    const double w = epi_power(E, v * E.scale);
    if (j >= i + E.shift) E.out[R.ro + j] = w;
    return w;
  } else if (EPI == EPI_DENSE) {
    double w = epi_power(E, v * E.scale);
    if (i == j && E.diag_one) w = 1.0;
    if (E.mirror && j < i) return w;  // lower half of a diagonal tile: filled by the mirror write
    E.out[(int64_t)i * E.ldo + j] = w;
    if (MIRROR_INLINE && E.mirror && i != j) E.out[(int64_t)j * E.ldo + i] = w;
    return w;
  } else if (EPI == EPI_ACCUM) {
    // One thread owns an element within a launch and launches are stream ordered, so the fire-and-forget
    // reduction (RED.ADD.F64 at the L2) is deterministic; it saves the load round trip of a read-modify-write.
    atomicAdd(&E.out[(int64_t)i * E.ldo + j], v);
    return v;
  } else if (EPI == EPI_DIFFCOR) {
    if (j > i) {
      if (pass == 0) {
        E.r_a[R.ro + j] = v;
      } else {
        const double ra = E.r_a[R.ro + j];  // parked by this very thread in pass 0
        const double zz = (atanh(ra) - atanh(v)) * E.inv_se;
        E.r_b[R.ro + j] = v;
        E.z[R.ro + j] = zz;
        E.pv[R.ro + j] = erfc(fabs(zz) * 0.70710678118654752440);
      }
    }
    return v;
  } else {  // EPI_TOM
    double a;
    if (E.A) {
      a = E.A[(int64_t)i * E.p + j];
    } else if (E.planes_half) {  // hi + lo carries 22 significant bits: inside the 1e-5 class this path serves
      const __half* r = reinterpret_cast<const __half*>(E.A_planes) + (int64_t)i * E.planes_ld;
      a = ((double)__half2float(r[j]) + (double)__half2float(r[E.planes_K + j])) * (1.0 / kF16Scale);
    } else {
      const float* r = reinterpret_cast<const float*>(E.A_planes) + (int64_t)i * E.planes_ld;
      a = (double)r[j] + (double)r[E.planes_K + j];
    }
    const double absa = E.tom_signed ? fabs(a) : a;
    const double shared = v - a * (R.d_i + E.dvec[j]);
    const double mk = fmin(R.k_i, E.kvec[j]);
    double tom = E.tom_v2 ? 0.5 * (a + shared / (mk + absa)) : (a + shared) / (mk + 1.0 - absa);
    if (i == j) tom = 1.0;
    if (E.tom_packed) {
      if (j > i) E.out[R.ro + j] = tom;
    } else if (!(E.mirror && j < i)) {
      E.out[(int64_t)i * E.ldo + j] = tom;
      if (MIRROR_INLINE && E.mirror && i != j) E.out[(int64_t)j * E.ldo + i] = tom;
    }
    return tom;
  }
}

// does this epilogue configuration need the separate transposed write?
template <int EPI>
__device__ __forceinline__ bool epi_wants_mirror(const EpiParams& E) {
  if (EPI == EPI_DENSE) return E.mirror != 0;
  if (EPI == EPI_TOM) return E.mirror != 0 && !E.tom_packed;
  return false;
}

__device__ __forceinline__ void epi_mirror(const EpiParams& E, int i, int j, double w) {
  if (j > i) E.out[(int64_t)j * E.ldo + i] = w;
}

}  // namespace bixcor
