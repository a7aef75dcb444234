// tcgen05 / TMEM Gram kernels: the fast 3xTF32 path and the exact int8 Spearman path.
//
//   G[i][j] = sum_k R[i][k] * C[j][k]   with both operands K-major, exactly the shape of
//   Z'Z (column_pairwise_cor, src/rust/src/base/r_cors_similarity.rs:74) and of the A*A of
//   calc_tom (src/rust/src/methods/r_coremo.rs:54).
//
// Operands live in HBM as two "digit planes" per matrix:
//   KIND_TF32  plane 0 = hi = tf32(z), plane 1 = lo = tf32(z - hi); the product is recovered as
//              hi*hi + hi*lo + lo*hi (3 MMAs, FP32 accumulation in one TMEM accumulator).
//   KIND_I8    Spearman only.  d = 2*rank - (n+1) is an integer, d = 16*h + l with |h| <= 127,
//              l in [-8,7]; plane 0 = h, plane 1 = l (int8).  Three exact int32 accumulators
//              hh, hl+lh, ll; the epilogue recombines S = 256*hh + 16*(hl+lh) + ll exactly in
//              FP64 and scales by 1/(|d_i||d_j|): an integer-exact Gram at int8 tensor rate.
//
// Structure (persistent, warp specialised, one CTA per SM):
//   warp 0   TMA producer: 4 x cp.async.bulk.tensor (A hi/lo, B hi/lo tiles of 128 rows x 128 B,
//            SWIZZLE_128B) per pipeline stage, completion on an mbarrier (complete_tx)
//   warp 1   MMA issuer: one elected lane issues tcgen05.mma.cta_group::1 (M=128, N=128) with
//            shared-memory descriptors; tcgen05.commit releases smem stages / publishes accumulators
//   warp 2   TMEM allocator
//   warps 4-11 epilogue (two per scheduler): drain the accumulators with tcgen05.ld into registers
//            (releasing TMEM to the MMA warp at once), then run the fused epilogue (epilogue.cuh)
//            through a swizzled shared-memory transpose so stores are contiguous row runs
#pragma once

#include <cuda.h>

#include "common.cuh"
#include "epilogue.cuh"

namespace bixcor {
namespace umma {

enum Kind : int { KIND_TF32 = 0, KIND_I8 = 1 };

constexpr int BM = 128, BN = 128;
constexpr int ROW_BYTES = 128;                    // one SWIZZLE_128B row: 32 tf32 or 128 int8
constexpr int PLANE_TILE_BYTES = BM * ROW_BYTES;  // 16 KB
constexpr int STAGE_BYTES = 4 * PLANE_TILE_BYTES;  // A hi, A lo, B hi, B lo
constexpr int STAGES = 3;
constexpr int EPI_WARPS = 8;                                  // two per scheduler: (TMEM lane quarter) x (column half)
constexpr int THREADS = 128 + 32 * EPI_WARPS;
constexpr int EPI_COLS = BN / (EPI_WARPS / 4);                // accumulator columns owned by one epilogue warp
constexpr int EPI_CHUNK = 16;                                 // columns transposed per step
constexpr int EPI_STAGE_BYTES = EPI_WARPS * 32 * EPI_CHUNK * 8;  // one 32 x 16 double transpose buffer per warp
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + EPI_STAGE_BYTES;

struct Params {
  int ntiles;
  const int2* tiles;
  int npass;         // 2 for the differential correlation (operand set 1 in pass 1)
  int kblocks[2];    // 128-byte K blocks per pass
  const double* inv_norm[2];  // KIND_I8: 1/|d| per gene and pass
  int zero;  // always 0; opaque to the compiler (scheduling fence, see the epilogue drain)
  int dbg;  // profiling aid (env BIXCOR_UMMA_DEBUG): 1 = drain only, 2 = generic path without global stores, 3 = no MMA issue,
            // 4 = lean path without global stores, 5 = lean path without phase B, 6 = no TMA and no MMA (epilogue alone)
  EpiParams E;
};

// ---------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) __trap();
  }
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

template <int KIND>
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  if (KIND == KIND_TF32) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}

// 32 lanes x 16 consecutive 32-bit columns -> 16 registers per thread
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format): start address >> 4,
// LBO = 1 (unused for swizzled K-major), SBO = 1024 B (8 rows x 128 B), version 1, layout 2.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// instruction descriptor: dense, K-major A and B, M = 128, N = 128
template <int KIND>
__host__ __device__ constexpr uint32_t make_idesc() {
  return (KIND == KIND_TF32 ? (1u << 4) | (2u << 7) | (2u << 10)    // D = F32, A = B = TF32
                            : (2u << 4) | (1u << 7) | (1u << 10))   // D = S32, A = B = signed int8
         | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

template <int KIND>
struct KindTraits;
template <>
struct KindTraits<KIND_TF32> {
  static constexpr int NACC = 1;        // one FP32 accumulator
  static constexpr int ACC_STAGES = 2;  // double buffered in TMEM
  // The tensor core adds into its FP32 accumulator with truncation, a bias that grows linearly with
  // the number of accumulations.  The MMA therefore only accumulates CHUNK_KB K-blocks (128 samples)
  // at a time; the epilogue warps drain every chunk from TMEM and sum the chunks in FP32 registers
  // with round-to-nearest (measured: 1e-5 -> <1e-6 at n = 1000, and independent of K).
  static constexpr int CHUNK_KB = 4;
};
template <>
struct KindTraits<KIND_I8> {
  static constexpr int NACC = 3;  // hh, hl+lh, ll (exact int32)
  static constexpr int ACC_STAGES = 1;
  static constexpr int CHUNK_KB = 1 << 30;  // integer accumulation is exact: one chunk = whole K
};


// Lean epilogue of one interior tile of the packed correlation for one epilogue warp: 32 rows x
// EPI_COLS columns, nothing predicated.  racc holds the drained accumulators (lane = row).  Row-side
// factors are applied with lane = row, column-side factors and the power after the shared-memory
// transpose with lane = (row parity, column); the 16 row offsets of a lane are 32-bit, computed once
// per tile and pinned in registers, so a store costs one 64-bit add.
template <int KIND, bool PW6>
__device__ __forceinline__ void packed_interior(const uint32_t (&racc)[EPI_COLS], double* stg, int lane, double rowf,
                                                bool tf32_scaled, const double* colv, char* wbase, int a, int dbg) {
  const int rr = lane >> 4, cc = lane & 15;
  uint32_t roff[16];
#pragma unroll
  for (int k2 = 0; k2 < 16; ++k2) {
    const int m = 2 * k2 + rr;
    roff[k2] = (uint32_t)(m * a - (m * (m - 1)) / 2 + cc) << 3;
    asm volatile("" : "+r"(roff[k2]));  // keep it: ptxas otherwise rematerialises the product per store
  }
  double colf[EPI_COLS / EPI_CHUNK];
#pragma unroll
  for (int c = 0; c < EPI_COLS / EPI_CHUNK; ++c) colf[c] = (KIND == KIND_I8) ? __ldg(colv + c * EPI_CHUNK + cc) : 1.0;
  const uint32_t stg_s = smem_u32(stg);
  uint32_t st_addr[EPI_CHUNK / 2];  // 16-byte unit u of row `lane` lives at unit u ^ (lane & 7)
#pragma unroll
  for (int u = 0; u < EPI_CHUNK / 2; ++u)
    st_addr[u] = stg_s + (uint32_t)lane * (EPI_CHUNK * 8) + ((((uint32_t)u) ^ ((uint32_t)lane & 7u)) << 4);
  // row r = 2*k2 + rr keeps unit (cc >> 1) at (cc >> 1) ^ (r & 7) = xq ^ ((2*k2) & 7)
  const uint32_t xq = (uint32_t)((cc >> 1) ^ rr);
  uint32_t ld_off[4];
#pragma unroll
  for (int t = 0; t < 4; ++t)
    ld_off[t] = (uint32_t)rr * (EPI_CHUNK * 8) + (uint32_t)(cc & 1) * 8 + ((xq ^ (uint32_t)(2 * t)) << 4);
  const char* const sb = reinterpret_cast<const char*>(stg);
#pragma unroll
  for (int c = 0; c < EPI_COLS; c += EPI_CHUNK) {
#pragma unroll
    for (int q = 0; q < EPI_CHUNK; q += 2) {  // phase A (lane = row)
      double v0, v1;
      if (KIND == KIND_I8) {
        v0 = (double)(int)racc[c + q] * rowf;
        v1 = (double)(int)racc[c + q + 1] * rowf;
      } else {
        v0 = (double)__uint_as_float(racc[c + q]);
        v1 = (double)__uint_as_float(racc[c + q + 1]);
        if (tf32_scaled) {
          v0 *= rowf;
          v1 *= rowf;
        }
      }
      asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(st_addr[q >> 1]), "d"(v0), "d"(v1) : "memory");
    }
    __syncwarp();
    const double cf = colf[c / EPI_CHUNK];
    if (dbg == 5) {  // profiling: no phase B at all
      __syncwarp();
      continue;
    }
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) {  // phase B (lane = 2 rows x 16 columns): two 128-byte runs per store
      double v = *reinterpret_cast<const double*>(sb + ld_off[k2 & 3] + k2 * (2 * EPI_CHUNK * 8));
      if (KIND == KIND_I8) v *= cf;
      if (PW6) {
        const double sq = v * v;
        v = sq * sq * sq;
      }
      if (dbg == 4) {  // profiling: everything but the global store
        asm volatile("" ::"d"(v));
        continue;
      }
      *reinterpret_cast<double*>(wbase + roff[k2] + (uint32_t)(c * 8)) = v;
    }
    __syncwarp();
  }
}

template <int KIND, int EPI>
__global__ void __launch_bounds__(THREADS, 1)
gram_umma_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1, const Params P) {
  using T = KindTraits<KIND>;
  constexpr int TMEM_COLS = (T::NACC * T::ACC_STAGES * BN <= 256) ? 256 : 512;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B needs 1024-byte alignment
  const uint32_t bar_base = base + STAGES * STAGE_BYTES;
  // barriers: full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], then the TMEM base address slot
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
  double* epi_stage_all = reinterpret_cast<double*>(smem_raw + (bar_base + 256u - smem_u32(smem_raw)));
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map0) : "memory");
    if (P.npass > 1) asm volatile("prefetch.tensormap [%0];" ::"l"(&map1) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), EPI_WARPS);  // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < P.ntiles; t += gridDim.x) {
        const int2 tile = P.tiles[t];
        for (int pass = 0; pass < P.npass; ++pass) {
          const CUtensorMap* map = pass ? &map1 : &map0;
          const int kb_n = P.kblocks[pass];
          const int kelems = (KIND == KIND_TF32) ? 32 : 128;
          for (int kb = 0; kb < kb_n; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            if (P.dbg == 6) {  // profiling: no operand traffic at all
              mbar_arrive(full_bar(stage));
            } else {
              mbar_expect_tx(full_bar(stage), STAGE_BYTES);
              const uint32_t s0 = base + stage * STAGE_BYTES;
              tma_load_3d(s0 + 0 * PLANE_TILE_BYTES, map, full_bar(stage), kb * kelems, 0, tile.x * BM);
              tma_load_3d(s0 + 1 * PLANE_TILE_BYTES, map, full_bar(stage), kb * kelems, 1, tile.x * BM);
              tma_load_3d(s0 + 2 * PLANE_TILE_BYTES, map, full_bar(stage), kb * kelems, 0, tile.y * BN);
              tma_load_3d(s0 + 3 * PLANE_TILE_BYTES, map, full_bar(stage), kb * kelems, 1, tile.y * BN);
            }
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ==============================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc<KIND>();
      int stage = 0, acc_stage = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int t = blockIdx.x; t < P.ntiles; t += gridDim.x) {
        for (int pass = 0; pass < P.npass; ++pass) {
          const int kb_n = P.kblocks[pass];
          for (int kb0 = 0; kb0 < kb_n; kb0 += T::CHUNK_KB) {
            const int kb1 = (kb_n - kb0 > T::CHUNK_KB) ? kb0 + T::CHUNK_KB : kb_n;
            mbar_wait(tempty_bar(acc_stage), acc_phase ^ 1u);  // epilogue has drained this accumulator set
            tc_fence_after();
            const uint32_t acc0 = tmem_base + (uint32_t)(acc_stage * T::NACC * BN);
            for (int kb = kb0; kb < kb1; ++kb) {
              mbar_wait(full_bar(stage), phase);
              tc_fence_after();
              const uint32_t s0 = base + stage * STAGE_BYTES;
              const uint64_t dA0 = make_smem_desc(s0 + 0 * PLANE_TILE_BYTES);
              const uint64_t dA1 = make_smem_desc(s0 + 1 * PLANE_TILE_BYTES);
              const uint64_t dB0 = make_smem_desc(s0 + 2 * PLANE_TILE_BYTES);
              const uint64_t dB1 = make_smem_desc(s0 + 3 * PLANE_TILE_BYTES);
#pragma unroll
              for (int k = 0; k < 4; ++k) {             // 4 x 32-byte K steps inside the 128-byte swizzle row
                const uint64_t ko = (uint64_t)(k * 2);   // +32 bytes, in 16-byte descriptor units
                const uint32_t first = (kb == kb0 && k == 0) ? 0u : 1u;
                if (P.dbg == 3 || P.dbg == 6) continue;
                if (KIND == KIND_TF32) {
                  tc_mma<KIND>(acc0, dA0 + ko, dB0 + ko, idesc, first);  // hi*hi
                  tc_mma<KIND>(acc0, dA0 + ko, dB1 + ko, idesc, 1u);     // hi*lo
                  tc_mma<KIND>(acc0, dA1 + ko, dB0 + ko, idesc, 1u);     // lo*hi
                } else {
                  tc_mma<KIND>(acc0 + 0 * BN, dA0 + ko, dB0 + ko, idesc, first);  // h*h
                  tc_mma<KIND>(acc0 + 1 * BN, dA0 + ko, dB1 + ko, idesc, first);  // h*l
                  tc_mma<KIND>(acc0 + 1 * BN, dA1 + ko, dB0 + ko, idesc, 1u);     // l*h
                  tc_mma<KIND>(acc0 + 2 * BN, dA1 + ko, dB1 + ko, idesc, first);  // l*l
                }
              }
              tc_commit(empty_bar(stage));  // smem stage reusable once these MMAs have read it
              if (++stage == STAGES) {
                stage = 0;
                phase ^= 1u;
              }
            }
            tc_commit(tfull_bar(acc_stage));  // chunk accumulators complete -> epilogue warps
            if (++acc_stage == T::ACC_STAGES) {
              acc_stage = 0;
              acc_phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ============================ epilogue ================================
    // Warp e = warp - 4 owns TMEM lanes [32*(e%4), +32) (its hardware lane quarter) and the
    // accumulator columns [EPI_COLS*(e/4), +EPI_COLS) of every tile.
    constexpr bool DIRECT = (KIND == KIND_I8 && EPI == EPI_PACKED);
    static_assert(EPI_CHUNK == 16, "the direct path loads 16 accumulator columns per step");
    const int ew = (warp - 4) & 3;
    const int ch = (warp - 4) >> 2;
    int acc_stage = 0;
    uint32_t acc_phase = 0;
    const EpiParams& E = P.E;
    double* stg = epi_stage_all + (warp - 4) * (32 * EPI_CHUNK);
    const bool want_mirror = epi_wants_mirror<EPI>(E);
    const int pp = (int)E.p;
    const int dbg = P.dbg;
    // loop invariants pulled out of the constant bank once
    double* const out = E.out;
    const int64_t ldo = E.ldo;
    const int shift = E.shift;
    const int diag_one = E.diag_one;
    const int keep_sign = E.keep_sign;
    const double scale = E.scale;
    // 1 identity, 6 = WGCNA default, 0 generic (any other exponent, or an RBF transform)
    const int pw_mode = E.rbf_mode ? 0 : ((E.beta == 1.0) ? 1 : (E.beta_int == 6 ? 6 : 0));
    for (int t = blockIdx.x; t < P.ntiles; t += gridDim.x) {
      const int2 tile = P.tiles[t];
      const int row_base = tile.x * BM + ew * 32;
      const int i = row_base + lane;
      const int col_base = tile.y * BN + ch * EPI_COLS;
      for (int pass = 0; pass < P.npass; ++pass) {
        // Interior tiles of the packed correlation (97% of the tiles at p = 20000) lie strictly above the
        // diagonal and fully inside the matrix: nothing is predicated and they take the lean path below.
        const bool interior = (EPI == EPI_PACKED) && tile.x < tile.y && tile.y * BN + BN <= pp && pw_mode != 0 && !keep_sign && (dbg == 0 || dbg >= 3);
        // DIRECT kernels keep the accumulators of the few non-interior tiles in TMEM and finish them chunk
        // by chunk (the MMA warp waits a little longer on 3% of the tiles); in exchange the 64-register
        // drain array only lives on the lean path and the kernel stays free of local-memory spills.
        const bool direct = DIRECT && !interior;
        uint32_t racc[EPI_COLS];
        const int kb_n = P.kblocks[pass];
        uint32_t tacc_direct = 0;
        if (direct) {
          mbar_wait(tfull_bar(acc_stage), acc_phase);
          tc_fence_after();
          tacc_direct = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc_stage * T::NACC * BN + ch * EPI_COLS);
        } else {
          // ---- drain: TMEM -> registers, chunk by chunk (FP32 round-to-nearest / exact int32 sums)
          for (int kb0 = 0; kb0 < kb_n; kb0 += T::CHUNK_KB) {
            mbar_wait(tfull_bar(acc_stage), acc_phase);
            tc_fence_after();
            const uint32_t tacc = tmem_base + ((uint32_t)(ew * 32) << 16) +
                                  (uint32_t)(acc_stage * T::NACC * BN + ch * EPI_COLS);
            if (KIND == KIND_I8) {
              // `dep` is always 0 (P.zero is a runtime 0): it chains each batch of loads to the previous
              // batch's sums (all 16 of them, or ptxas only hurries the one it needs), otherwise ptxas hoists all 12 tcgen05.ld (192 registers) and spills.
              uint32_t dep = 0;
#pragma unroll
              for (int c = 0; c < EPI_COLS; c += 16) {
                uint32_t r0[16], r1[16], r2[16];
                tmem_ld16(tacc + c + dep, r0);
                tmem_ld16(tacc + BN + c + dep, r1);
                tmem_ld16(tacc + 2 * BN + c + dep, r2);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 16; ++q)  // S = 256*hh + 16*(hl+lh) + ll, exact in int32 for n <= 1860
                  racc[c + q] = (uint32_t)(256 * (int)r0[q] + 16 * (int)r1[q] + (int)r2[q]);
                uint32_t any = 0;
#pragma unroll
                for (int q = 0; q < 16; ++q) any |= racc[c + q];
                dep = any & (uint32_t)P.zero;
              }
            } else {
#pragma unroll
              for (int c = 0; c < EPI_COLS; c += 32) {
                uint32_t r0[32];
                tmem_ld32(tacc + c, r0);
                tmem_ld_wait();
                if (kb0 == 0) {
#pragma unroll
                  for (int q = 0; q < 32; ++q) racc[c + q] = r0[q];
                } else {
#pragma unroll
                  for (int q = 0; q < 32; ++q)
                    racc[c + q] = __float_as_uint(__uint_as_float(racc[c + q]) + __uint_as_float(r0[q]));
                }
              }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc_stage));  // TMEM buffer free: the MMA warp moves on
            if (++acc_stage == T::ACC_STAGES) {
              acc_stage = 0;
              acc_phase ^= 1u;
            }
          }
        }
        auto release_direct = [&]() {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(acc_stage));
          if (++acc_stage == T::ACC_STAGES) {
            acc_stage = 0;
            acc_phase ^= 1u;
          }
        };
        if (dbg == 1) {
          if (direct) release_direct();
          continue;
        }
        // ---- lean path: row-side factors are applied with lane = row, column-side factors and the power
        //      after the transpose with lane = column; the 16 row offsets of a lane are 32-bit and computed
        //      once per tile, so a store costs one address add.
        if (interior) {
          const EpiRow R0 = epi_row<EPI>(E, row_base);  // warp uniform
          double rowf = scale;
          if (KIND == KIND_I8) rowf *= P.inv_norm[pass][i];
          const double* colv = (KIND == KIND_I8) ? P.inv_norm[pass] + col_base : nullptr;
          char* const wbase = reinterpret_cast<char*>(out + (R0.ro + col_base));
          const int a = pp - row_base - shift - 1;  // ro(i + 1) - ro(i) = p - i - shift - 1
          if (pw_mode == 6) packed_interior<KIND, true>(racc, stg, lane, rowf, scale != 1.0, colv, wbase, a, dbg);
          else packed_interior<KIND, false>(racc, stg, lane, rowf, scale != 1.0, colv, wbase, a, dbg);
          continue;
        }
        // ---- fused epilogue from registers.  Each 32 x 16 block is transposed through shared
        //      memory (XOR-swizzled, conflict free) so that one store instruction writes two
        //      contiguous 128-byte runs of the packed rows instead of 32 scattered doubles.
        const bool row_ok = i < pp;
        EpiRow R;
        R.i = i;
        R.ro = 0;
        R.k_i = R.d_i = 0.0;
        double inv_i = 0.0;
        if (row_ok) {
          R = epi_row<EPI>(E, i);
          if (KIND == KIND_I8) inv_i = P.inv_norm[pass][i];
        }
        const int rr = lane >> 4, cc = lane & 15;
        const double* inv_col = (KIND == KIND_I8) ? P.inv_norm[pass] : nullptr;
#pragma unroll
        for (int c = 0; c < EPI_COLS; c += EPI_CHUNK) {
          const int jb = col_base + c;
          // blocks entirely below the diagonal of a diagonal tile hold nothing to store (warp uniform)
          if (EPI != EPI_ACCUM && jb + EPI_CHUNK - 1 < row_base &&
              (EPI == EPI_PACKED || EPI == EPI_DIFFCOR || want_mirror || (EPI == EPI_TOM && E.tom_packed)))
            continue;
          // phase A (lane = row): finish the Gram value.  All mode decisions are warp uniform and sit
          // OUTSIDE the element loops, so each loop body is a handful of FP64 instructions.
          double vals[EPI_CHUNK];
          if (DIRECT) {  // (KIND_I8) straight from TMEM
            uint32_t r0[16], r1[16], r2[16];
            tmem_ld16(tacc_direct + c, r0);
            tmem_ld16(tacc_direct + BN + c, r1);
            tmem_ld16(tacc_direct + 2 * BN + c, r2);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < EPI_CHUNK; ++q)
              vals[q] = (double)(256 * (int)r0[q] + 16 * (int)r1[q] + (int)r2[q]) * inv_i;
          } else {
#pragma unroll
            for (int q = 0; q < EPI_CHUNK; ++q)
              vals[q] = (KIND == KIND_I8) ? (double)(int)racc[c + q] * inv_i : (double)__uint_as_float(racc[c + q]);
          }
          if (KIND == KIND_I8) {
#pragma unroll
            for (int q = 0; q < EPI_CHUNK; ++q) vals[q] *= (jb + q < pp) ? __ldg(inv_col + jb + q) : 0.0;
          }
          if ((EPI == EPI_PACKED || EPI == EPI_DENSE) && scale != 1.0) {
#pragma unroll
            for (int q = 0; q < EPI_CHUNK; ++q) vals[q] *= scale;
          }
          if ((EPI == EPI_PACKED || EPI == EPI_DENSE) && pw_mode == 6) {
#pragma unroll
            for (int q = 0; q < EPI_CHUNK; ++q) {
              const double sq = vals[q] * vals[q];
              const double w = sq * sq * sq;
              vals[q] = keep_sign ? copysign(w, vals[q]) : w;
            }
          }
#pragma unroll
          for (int q = 0; q < EPI_CHUNK; ++q) stg[lane * EPI_CHUNK + (q ^ (lane & 15))] = vals[q];
          if ((EPI == EPI_PACKED || EPI == EPI_DENSE) && pw_mode == 0) {
            // generic exponent: transform the staged values in place (keeps vals[] in registers above)
#pragma unroll 1
            for (int q = 0; q < EPI_CHUNK; ++q) {
              double* sp = stg + lane * EPI_CHUNK + (q ^ (lane & 15));
              *sp = epi_power(E, *sp);  // (scale already applied above)
            }
          }
          __syncwarp();
          // phase B (lane = 2 rows x 16 columns): contiguous 128-byte runs per row.  Shuffles and
          // shared loads are issued for all 16 row pairs before the predicated stores (no branches).
          const int j = jb + cc;
          const bool col_ok = (j < pp) && (dbg != 2);
          if (EPI == EPI_PACKED) {  // diagonal / ragged tiles only (interior tiles took the path above)
#pragma unroll 2
            for (int k2 = 0; k2 < 16; ++k2) {
              const int r = 2 * k2 + rr;
              const int ir = row_base + r;
              const int64_t ro_r = __shfl_sync(0xffffffffu, R.ro, r);
              const double v = stg[r * EPI_CHUNK + (cc ^ (r & 15))];
              if (col_ok && (ir < pp) && (j >= ir + shift)) out[ro_r + j] = v;
            }
          } else if (EPI == EPI_DENSE) {
#pragma unroll 8
            for (int r2 = 0; r2 < 32; r2 += 2) {
              const int r = r2 + rr;
              const int ir = row_base + r;
              double v = stg[r * EPI_CHUNK + (cc ^ (r & 15))];
              if (ir == j && diag_one) v = 1.0;
              if (col_ok && ir < pp && !(want_mirror && j < ir)) out[(int64_t)ir * ldo + j] = v;
            }
          } else {
#pragma unroll 2
            for (int r2 = 0; r2 < 32; r2 += 2) {
              const int r = r2 + rr;
              EpiRow Rr;
              Rr.i = row_base + r;
              Rr.ro = __shfl_sync(0xffffffffu, R.ro, r);
              if (EPI == EPI_TOM) {
                Rr.k_i = __shfl_sync(0xffffffffu, R.k_i, r);
                Rr.d_i = __shfl_sync(0xffffffffu, R.d_i, r);
              } else {
                Rr.k_i = Rr.d_i = 0.0;
              }
              if (Rr.i < pp && col_ok) {
                const double v = stg[r * EPI_CHUNK + (cc ^ (r & 15))];
                const double w = epi_store<EPI, false>(E, Rr, j, v, pass);
                if (want_mirror) stg[r * EPI_CHUNK + (cc ^ (r & 15))] = w;
              }
            }
          }
          __syncwarp();
          if (want_mirror) {  // transposed element: consecutive lanes = consecutive rows i -> coalesced
            if (row_ok) {
#pragma unroll 4
              for (int q = 0; q < EPI_CHUNK; ++q) {
                const int jj = jb + q;
                if (jj < pp && jj > i) out[(int64_t)jj * ldo + i] = stg[lane * EPI_CHUNK + (q ^ (lane & 15))];
              }
            }
            __syncwarp();
          }
        }
        if (direct) release_direct();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline char* err_buf() {
  static char buf[512] = "";
  return buf;
}
inline const char* last_error() { return err_buf(); }

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = (EncodeTiledFn)p;
  return fn;
}

// 3-D map over the gene-major operand [row][plane][K]: dims (K, 2 planes, rows), box (128 bytes of K,
// 1 plane, 128 rows), SWIZZLE_128B; out-of-range rows are zero filled (ragged last tile).
inline int make_map(CUtensorMap* map, int kind, const void* ptr, int64_t rows, int K) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    snprintf(err_buf(), 512, "cuTensorMapEncodeTiled entry point not available");
    return -1;
  }
  const int es = kind == KIND_TF32 ? 4 : 1;
  cuuint64_t dims[3] = {(cuuint64_t)K, 2, (cuuint64_t)rows};
  cuuint64_t strides[2] = {(cuuint64_t)K * es, (cuuint64_t)K * es * 2};
  cuuint32_t box[3] = {(cuuint32_t)(ROW_BYTES / es), 1, (cuuint32_t)BM};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, kind == KIND_TF32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 3,
                  const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(err_buf(), 512, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld K=%d es=%d)", (int)r,
             (long long)rows, K, es);
    return -1;
  }
  return 0;
}

template <int KIND, int EPI>
inline int launch_one(cudaStream_t stream, int sm_count, const CUtensorMap& m0, const CUtensorMap& m1, const Params& P) {
  auto kern = gram_umma_kernel<KIND, EPI>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) {
    snprintf(err_buf(), 512, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return -1;
  }
  const int grid = P.ntiles < sm_count ? P.ntiles : sm_count;
  kern<<<grid, THREADS, SMEM_BYTES, stream>>>(m0, m1, P);
  e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(err_buf(), 512, "gram_umma_kernel launch: %s", cudaGetErrorString(e));
    return -1;
  }
  return 0;
}

// operand set s: gene-major digit planes ([row][plane][K]), K elements (padded), inv_norm (KIND_I8)
inline int launch_gram_umma(cudaStream_t stream, int sm_count, int kind, int epi, int64_t rows, const void* planes0,
                            int K0, const double* inv0, const void* planes1, int K1, const double* inv1,
                            const int2* tiles, int ntiles, const EpiParams& E) {
  if (ntiles <= 0) return 0;
  const int kelems = kind == KIND_TF32 ? 32 : 128;
  if (K0 % kelems != 0 || (planes1 && K1 % kelems != 0)) {
    snprintf(err_buf(), 512, "operand K must be padded to %d elements", kelems);
    return -1;
  }
  CUtensorMap m0, m1;
  if (make_map(&m0, kind, planes0, rows, K0)) return -1;
  if (planes1) {
    if (make_map(&m1, kind, planes1, rows, K1)) return -1;
  } else {
    m1 = m0;
  }
  Params P;
  P.ntiles = ntiles;
  P.tiles = tiles;
  P.npass = planes1 ? 2 : 1;
  P.kblocks[0] = K0 / kelems;
  P.kblocks[1] = planes1 ? K1 / kelems : 0;
  P.inv_norm[0] = inv0;
  P.inv_norm[1] = inv1;
  P.E = E;
  P.zero = 0;
  {
    const char* d = getenv("BIXCOR_UMMA_DEBUG");
    P.dbg = d ? atoi(d) : 0;
  }
#define BIXCOR_UMMA_CASE(KIND, EPI) \
  if (kind == KIND && epi == EPI) return launch_one<KIND, EPI>(stream, sm_count, m0, m1, P);
  BIXCOR_UMMA_CASE(KIND_TF32, EPI_PACKED)
  BIXCOR_UMMA_CASE(KIND_TF32, EPI_DENSE)
  BIXCOR_UMMA_CASE(KIND_TF32, EPI_DIFFCOR)
  BIXCOR_UMMA_CASE(KIND_TF32, EPI_TOM)
  BIXCOR_UMMA_CASE(KIND_TF32, EPI_ACCUM)
  BIXCOR_UMMA_CASE(KIND_I8, EPI_PACKED)
  BIXCOR_UMMA_CASE(KIND_I8, EPI_DENSE)
  BIXCOR_UMMA_CASE(KIND_I8, EPI_DIFFCOR)
#undef BIXCOR_UMMA_CASE
  snprintf(err_buf(), 512, "no tcgen05 kernel for kind %d / epilogue %d", kind, epi);
  return -1;
}

}  // namespace umma
}  // namespace bixcor
