// Exact-path Gram kernel: FP64 tensor-core MMA (mma.sync DMMA family; tcgen05 has
// no FP64 kind) over upper-triangle 128x128 tiles with fused epilogues.
//
//   G[i][j] = sum_k R[i][k] * C[j][k]          (both operands K-major)
//
// R == C == Z gives the correlation Gram Z'Z of the standardised matrix
// (replaces the faer GEMM under column_pairwise_cor,
// src/rust/src/base/r_cors_similarity.rs:74), R == C == A gives the A*A of
// calc_tom (src/rust/src/methods/r_coremo.rs:54).  Only tiles on/above the
// diagonal are visited and the epilogue writes straight into the reference's
// packed triangular order (src/rust/src/base/r_helpers.rs:126-131), so the
// dense p x p intermediate + scalar gather of r_cors_similarity.rs:257-265
// never exists.
//
// Tiling: CTA 128x128 or 128x64 (see Shape<CFG>), K step 16 per stage, 4-stage
// cp.async pipeline.  Shared layout per operand per stage is
// [k8 group][row][8 doubles]: one LDS.128 per thread fetches the two k values a
// thread feeds to the MMA (k permuted identically for both operands), bank
// conflict free.
#pragma once

#include "common.cuh"
#include "epilogue.cuh"

namespace bixcor {
namespace dmma {

constexpr int BM = 128, BK = 16, STAGES = 4;

// Kernel shapes.  All use DMMA.8x8x4 (every FP64 mma.sync shape lowers to it on sm_100a).
//   CFG 0: CTA 128x128, 8 warps (2x4), warp tile 64x32, 1 CTA/SM   (2 warps per scheduler)
//   CFG 1: CTA 128x64,  8 warps (4x2), warp tile 32x32, 2 CTAs/SM  (4 warps per scheduler; one CTA's
//          epilogue / pipeline fill overlaps the other CTA's main loop)
//   CFG 2: CTA 128x128, 16 warps (4x4), warp tile 32x32, 1 CTA/SM  (4 warps per scheduler)
template <int CFG>
struct Shape;
template <>
struct Shape<0> {
  static constexpr int WM = 2, WN = 4, BN = 128, MIN_CTAS = 1;
};
template <>
struct Shape<1> {
  static constexpr int WM = 4, WN = 2, BN = 64, MIN_CTAS = 2;
};
template <>
struct Shape<2> {
  static constexpr int WM = 4, WN = 4, BN = 128, MIN_CTAS = 1;
};
template <int CFG>
constexpr int threads_of() { return 32 * Shape<CFG>::WM * Shape<CFG>::WN; }
template <int CFG>
constexpr int smem_bytes_of() { return STAGES * (BM + Shape<CFG>::BN) * BK * (int)sizeof(double); }

struct Params {
  // operand set 0 (rows / cols may alias), K multiple of BK, rows zero padded in K
  const double* R0;
  const double* C0;
  int64_t ld0;
  int K0;
  // operand set 1 (second Gram of the differential correlation)
  const double* R1;
  const double* C1;
  int64_t ld1;
  int K1;
  int p;              // number of row genes of G
  int pc;             // number of column genes of G (== p for the symmetric Gram)
  const int2* tiles;  // (block row, block col) list of 128 x 128 tiles
  EpiParams E;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void mma_8x8x4(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c[0]), "+d"(c[1])
      : "d"(a), "d"(b));
}

// Load one pipeline stage: rows [row0, row0+128) of R and [col0, col0+BN) of C, k in [k0, k0+16).
// Shared layout per operand: [k8 group (2)][row][8 doubles].
template <int CFG>
__device__ __forceinline__ void load_stage(double* stage, const double* __restrict__ R, const double* __restrict__ C,
                                           int64_t ld, int row0, int col0, int k0, int p, int pc, int tid) {
  constexpr int BN = Shape<CFG>::BN, THREADS = threads_of<CFG>();
  constexpr int CHUNKS = (BM + BN) * 8;  // 16-byte chunks of one stage
#pragma unroll
  for (int it = 0; it < CHUNKS / THREADS; ++it) {
    const int c = tid + it * THREADS;
    const bool isB = c >= BM * 8;
    const int cc = isB ? c - BM * 8 : c;
    const int row = cc >> 3, q = cc & 7;
    const int grow = (isB ? col0 : row0) + row;
    const bool valid = grow < (isB ? pc : p);
    const double* src = (isB ? C : R) + (int64_t)(valid ? grow : 0) * ld + k0 + q * 2;
    double* dst = stage + (isB ? BM * BK : 0) + (q >> 2) * ((isB ? BN : BM) * 8) + row * 8 + (q & 3) * 2;
    cp_async16(dst, src, valid);
  }
}

template <int CFG>
__device__ __forceinline__ void gram_mainloop(double (&acc)[(BM / Shape<CFG>::WM) / 8][(Shape<CFG>::BN / Shape<CFG>::WN) / 8][2],
                                              double* smem, const double* __restrict__ R, const double* __restrict__ C,
                                              int64_t ld, int K, int row0, int col0, int p, int pc, int tid) {
  constexpr int WN = Shape<CFG>::WN, BN = Shape<CFG>::BN;
  constexpr int WTM = BM / Shape<CFG>::WM, WTN = BN / WN, MT = WTM / 8, NT = WTN / 8;
  constexpr int STAGE_DOUBLES = (BM + BN) * BK;
  const int lane = tid & 31, warp = tid >> 5;
  const int wm = warp / WN, wn = warp % WN;
  const int g = lane >> 2, t = lane & 3;
  const int KT = K / BK;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < KT) load_stage<CFG>(smem + s * STAGE_DOUBLES, R, C, ld, row0, col0, s * BK, p, pc, tid);
    cp_async_commit();
  }
  for (int kt = 0; kt < KT; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int kn = kt + STAGES - 1;
      if (kn < KT) load_stage<CFG>(smem + (kn % STAGES) * STAGE_DOUBLES, R, C, ld, row0, col0, kn * BK, p, pc, tid);
      cp_async_commit();
    }
    const double* sA = smem + (kt % STAGES) * STAGE_DOUBLES;
    const double* sB = sA + BM * BK;
#pragma unroll
    for (int c8 = 0; c8 < BK / 8; ++c8) {
      double2 a[MT], b[NT];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
        a[mt] = *reinterpret_cast<const double2*>(sA + c8 * (BM * 8) + (wm * WTM + mt * 8 + g) * 8 + 2 * t);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
        b[nt] = *reinterpret_cast<const double2*>(sB + c8 * (BN * 8) + (wn * WTN + nt * 8 + g) * 8 + 2 * t);
      // k slice outermost: MT*NT independent MMAs between two MMAs on the same accumulator
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma_8x8x4(acc[mt][nt], a[mt].x, b[nt].x);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma_8x8x4(acc[mt][nt], a[mt].y, b[nt].y);
    }
  }
  cp_async_wait<0>();
  __syncthreads();  // smem may be refilled by a following mainloop
}

template <int EPI, int CFG>
__global__ void __launch_bounds__(threads_of<CFG>(), Shape<CFG>::MIN_CTAS) gram_dmma_kernel(const Params P) {
  constexpr int WN = Shape<CFG>::WN, BN = Shape<CFG>::BN;
  constexpr int WTM = BM / Shape<CFG>::WM, WTN = BN / WN, MT = WTM / 8, NT = WTN / 8;
  constexpr int SPLIT = 128 / BN;  // CTAs per 128 x 128 list tile
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x;
  const int2 tile = P.tiles[blockIdx.x / SPLIT];
  const int row0 = tile.x * BM, col0 = tile.y * 128 + (blockIdx.x % SPLIT) * BN;
  const int lane = tid & 31, warp = tid >> 5;
  const int wm = warp / WN, wn = warp % WN;
  const int g = lane >> 2, t = lane & 3;
  const int64_t p = P.p;
  const EpiParams& E = P.E;
  if (col0 >= P.pc) return;  // right half of a ragged last tile

  double acc[MT][NT][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

  constexpr int NPASS = (EPI == EPI_DIFFCOR) ? 2 : 1;
#pragma unroll 1
  for (int pass = 0; pass < NPASS; ++pass) {
    if (pass == 0)
      gram_mainloop<CFG>(acc, smem, P.R0, P.C0, P.ld0, P.K0, row0, col0, P.p, P.pc, tid);
    else
      gram_mainloop<CFG>(acc, smem, P.R1, P.C1, P.ld1, P.K1, row0, col0, P.p, P.pc, tid);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const int64_t i = row0 + wm * WTM + mt * 8 + g;
      if (i < p) {
        const EpiRow R = epi_row<EPI>(E, i);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int j = col0 + wn * WTN + nt * 8 + 2 * t + e;
            if (j < P.pc) epi_store<EPI>(E, R, j, acc[mt][nt][e], pass);
          }
      }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
    }
  }
}

}  // namespace dmma
}  // namespace bixcor
