// Shared device/host helpers for libbixcor (sm_100a only).
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace bixcor {

constexpr int kTile = 128;  // block-row granularity of the triangular tiling

// Packed upper-triangle offset of the first stored element of row i
// (order of src/rust/src/base/r_helpers.rs:126-131 in the reference).
__host__ __device__ __forceinline__ int64_t packed_row_offset(int64_t i, int64_t p, int shift) {
  return shift ? (i * (2 * p - i - 1)) / 2 : (i * (2 * p - i + 1)) / 2;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum; `scratch` holds >= 32 doubles.  All threads get the result.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect scratch reuse across consecutive calls
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double t = (lane < nwarp) ? scratch[lane] : 0.0;
  return warp_sum(t);
}

__device__ __forceinline__ long long block_sum_ll(long long v, long long* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  v = warp_sum_ll(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  long long t = (lane < nwarp) ? scratch[lane] : 0ll;
  return warp_sum_ll(t);
}

// Soft-threshold extension: beta == 1 is the identity (sign kept, bit-identical).
__device__ __forceinline__ double soft_power(double r, double beta, int keep_sign) {
  if (beta == 1.0) return r;
  double a = pow(fabs(r), beta);
  return keep_sign ? copysign(a, r) : a;
}

// round-to-nearest FP32 -> TF32 (low 13 mantissa bits cleared)
__device__ __forceinline__ float to_tf32(float f) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(f));
  return __uint_as_float(u);
}

// binary16 hi / lo split of the f16 tensor kind (BIXCOR_PREC_F16X3); the planes carry 2^10 z
constexpr double kF16Scale = 1024.0;  // keeps lo = f16(s z - hi) out of the subnormal range for |z| > 1e-4
__device__ __forceinline__ void split_f16(double z, __half& hi, __half& lo) {
  const double zs = z * kF16Scale;
  hi = __float2half_rn((float)zs);
  lo = __float2half_rn((float)(zs - (double)__half2float(hi)));
}

}  // namespace bixcor
