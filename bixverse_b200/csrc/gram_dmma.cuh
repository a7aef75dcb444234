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
// Tiling: CTA 128x128, 8 warps as 2(M) x 4(N), warp tile 64x32, K step 16 per
// stage, 4-stage cp.async pipeline.  Shared layout per operand per stage is
// [k8 group][row][8 doubles]: one LDS.128 per thread fetches the two k values a
// thread feeds to the MMA (k permuted identically for both operands), bank
// conflict free.
#pragma once

#include "common.cuh"
#include "epilogue.cuh"

namespace bixcor {
namespace dmma {

constexpr int BM = 128, BN = 128, BK = 16, STAGES = 4, THREADS = 256;
constexpr int OPERAND_DOUBLES = BM * BK;           // one operand tile of one stage
constexpr int STAGE_DOUBLES = 2 * OPERAND_DOUBLES;  // A + B
constexpr int SMEM_BYTES = STAGES * STAGE_DOUBLES * (int)sizeof(double);  // 128 KB

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
  int p;              // number of genes (rows and cols of G)
  const int2* tiles;  // (block row, block col) list
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

__device__ __forceinline__ void mma_16x8x8(double (&c0)[2], double (&c1)[2], double a0, double a1, double a2,
                                           double a3, double b0, double b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+d"(c0[0]), "+d"(c0[1]), "+d"(c1[0]), "+d"(c1[1])
      : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
}
__device__ __forceinline__ void mma_8x8x4(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma_16x8x16(double (&c0)[2], double (&c1)[2], const double2& a0lo,
                                            const double2& a1lo, const double2& a0hi, const double2& a1hi,
                                            const double2& blo, const double2& bhi) {
  // k slices (t, t+4, t+8, t+12) <- physical (lo.x, lo.y, hi.x, hi.y); rows g (a0*) and g+8 (a1*)
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
      "{%12,%13,%14,%15}, {%0,%1,%2,%3};"
      : "+d"(c0[0]), "+d"(c0[1]), "+d"(c1[0]), "+d"(c1[1])
      : "d"(a0lo.x), "d"(a1lo.x), "d"(a0lo.y), "d"(a1lo.y), "d"(a0hi.x), "d"(a1hi.x), "d"(a0hi.y), "d"(a1hi.y),
        "d"(blo.x), "d"(blo.y), "d"(bhi.x), "d"(bhi.y));
}

// Load one pipeline stage: rows [row0, row0+128) of R and [col0, col0+128) of C, k in [k0, k0+16).
__device__ __forceinline__ void load_stage(double* stage, const double* __restrict__ R, const double* __restrict__ C,
                                           int64_t ld, int row0, int col0, int k0, int p, int tid) {
#pragma unroll
  for (int it = 0; it < (2 * BM * BK / 2) / THREADS; ++it) {  // 2048 16-byte chunks / 256 threads
    const int c = tid + it * THREADS;
    const int which = c >> 10;
    const int cc = c & 1023;
    const int row = cc >> 3, q = cc & 7;
    const int grow = (which ? col0 : row0) + row;
    const bool valid = grow < p;
    const double* src = (which ? C : R) + (int64_t)(valid ? grow : 0) * ld + k0 + q * 2;
    double* dst = stage + which * OPERAND_DOUBLES + (q >> 2) * (BM * 8) + row * 8 + (q & 3) * 2;
    cp_async16(dst, src, valid);
  }
}

// MMA_VARIANT: 0 = m8n8k4 (two dependent MMAs per accumulator back to back), 1 = m16n8k8,
// 2 = m16n8k16, 3 = m8n8k4 with the k slice as the outer loop (32 independent MMAs between
// two MMAs on the same accumulator)
template <int MMA_VARIANT>
__device__ __forceinline__ void gram_mainloop(double (&acc)[8][4][2], double* smem, const double* __restrict__ R,
                                              const double* __restrict__ C, int64_t ld, int K, int row0, int col0,
                                              int p, int tid) {
  const int lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;
  const int g = lane >> 2, t = lane & 3;
  const int KT = K / BK;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < KT) load_stage(smem + s * STAGE_DOUBLES, R, C, ld, row0, col0, s * BK, p, tid);
    cp_async_commit();
  }
  for (int kt = 0; kt < KT; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int kn = kt + STAGES - 1;
      if (kn < KT) load_stage(smem + (kn % STAGES) * STAGE_DOUBLES, R, C, ld, row0, col0, kn * BK, p, tid);
      cp_async_commit();
    }
    const double* sA = smem + (kt % STAGES) * STAGE_DOUBLES;
    const double* sB = sA + OPERAND_DOUBLES;
    if (MMA_VARIANT == 2) {
      double2 a_lo[8], a_hi[8], b_lo[4], b_hi[4];
#pragma unroll
      for (int mt = 0; mt < 8; ++mt) {
        const double* pa = sA + (wm * 64 + mt * 8 + g) * 8 + 2 * t;
        a_lo[mt] = *reinterpret_cast<const double2*>(pa);
        a_hi[mt] = *reinterpret_cast<const double2*>(pa + BM * 8);
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const double* pb = sB + (wn * 32 + nt * 8 + g) * 8 + 2 * t;
        b_lo[nt] = *reinterpret_cast<const double2*>(pb);
        b_hi[nt] = *reinterpret_cast<const double2*>(pb + BN * 8);
      }
#pragma unroll
      for (int m2 = 0; m2 < 4; ++m2)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
          mma_16x8x16(acc[2 * m2][nt], acc[2 * m2 + 1][nt], a_lo[2 * m2], a_lo[2 * m2 + 1], a_hi[2 * m2],
                      a_hi[2 * m2 + 1], b_lo[nt], b_hi[nt]);
    } else {
#pragma unroll
      for (int c8 = 0; c8 < BK / 8; ++c8) {
        double2 a[8], b[4];
#pragma unroll
        for (int mt = 0; mt < 8; ++mt)
          a[mt] = *reinterpret_cast<const double2*>(sA + c8 * (BM * 8) + (wm * 64 + mt * 8 + g) * 8 + 2 * t);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
          b[nt] = *reinterpret_cast<const double2*>(sB + c8 * (BN * 8) + (wn * 32 + nt * 8 + g) * 8 + 2 * t);
        if (MMA_VARIANT == 1) {
#pragma unroll
          for (int m2 = 0; m2 < 4; ++m2)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
              mma_16x8x8(acc[2 * m2][nt], acc[2 * m2 + 1][nt], a[2 * m2].x, a[2 * m2 + 1].x, a[2 * m2].y,
                         a[2 * m2 + 1].y, b[nt].x, b[nt].y);
        } else if (MMA_VARIANT == 3) {
#pragma unroll
          for (int mt = 0; mt < 8; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) mma_8x8x4(acc[mt][nt], a[mt].x, b[nt].x);
#pragma unroll
          for (int mt = 0; mt < 8; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) mma_8x8x4(acc[mt][nt], a[mt].y, b[nt].y);
        } else {
#pragma unroll
          for (int mt = 0; mt < 8; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
              mma_8x8x4(acc[mt][nt], a[mt].x, b[nt].x);
              mma_8x8x4(acc[mt][nt], a[mt].y, b[nt].y);
            }
        }
      }
    }
  }
  cp_async_wait<0>();
  __syncthreads();  // smem may be refilled by a following mainloop
}

template <int EPI, int MMA_VARIANT>
__global__ void __launch_bounds__(THREADS, 1) gram_dmma_kernel(const Params P) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x;
  const int2 tile = P.tiles[blockIdx.x];
  const int row0 = tile.x * BM, col0 = tile.y * BN;
  const int lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;
  const int g = lane >> 2, t = lane & 3;
  const int64_t p = P.p;
  const EpiParams& E = P.E;

  double acc[8][4][2];
#pragma unroll
  for (int mt = 0; mt < 8; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

  constexpr int NPASS = (EPI == EPI_DIFFCOR) ? 2 : 1;
#pragma unroll 1
  for (int pass = 0; pass < NPASS; ++pass) {
    if (pass == 0)
      gram_mainloop<MMA_VARIANT>(acc, smem, P.R0, P.C0, P.ld0, P.K0, row0, col0, P.p, tid);
    else
      gram_mainloop<MMA_VARIANT>(acc, smem, P.R1, P.C1, P.ld1, P.K1, row0, col0, P.p, tid);
#pragma unroll
    for (int mt = 0; mt < 8; ++mt) {
      const int64_t i = row0 + wm * 64 + mt * 8 + g;
      if (i < p) {
        const EpiRow R = epi_row<EPI>(E, i);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int64_t j = col0 + wn * 32 + nt * 8 + 2 * t + e;
            if (j < p) epi_store<EPI>(E, R, j, acc[mt][nt][e], pass);
          }
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
    }
  }
}

}  // namespace dmma
}  // namespace bixcor
