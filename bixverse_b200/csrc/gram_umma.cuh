// tcgen05 / TMEM Gram kernels: the fast 3xTF32 path, the exact int8 Spearman path and the int8 passes of the
// FP64-class Ozaki path.
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
//              The same kind runs the Ozaki passes (BIXCOR_PREC_OZAKI): operand records hold six base-128
//              digit planes, a pass multiplies two plane pairs and adds w * (128^2 hh + 128 (hl+lh) + ll)
//              into an FP64 matrix (EPI_ACCUM, RED.ADD.F64).
//
// Structure (persistent, warp specialised, one CTA per SM):
//   warp 0   TMA producer: 4 x cp.async.bulk.tensor (A hi/lo, B hi/lo tiles of 128 rows x 128 B,
//            SWIZZLE_128B) per pipeline stage, completion on an mbarrier (complete_tx)
//   warp 1   MMA issuer: one elected lane issues tcgen05.mma.cta_group::1 (M=128, N=128) - or, as the leader of a
//            CTA pair, cta_group::2 (M=256) - with shared-memory descriptors; tcgen05.commit releases smem
//            stages / publishes accumulators
//   warp 2   TMEM allocator
//   warps 4-11 epilogue (two per scheduler): drain the accumulators with tcgen05.ld into registers
//            (releasing TMEM to the MMA warp at once), then run the fused epilogue (epilogue.cuh)
//            through a swizzled shared-memory transpose so stores are contiguous row runs
#pragma once

#include <cuda.h>

#include <vector>

#include "common.cuh"
#include "epilogue.cuh"

namespace bixcor {
namespace umma {

enum Kind : int { KIND_TF32 = 0, KIND_I8 = 1, KIND_F16 = 2 };
// KIND_F16: the same hi / lo split as KIND_TF32 on binary16 planes (hi = f16(s z), lo = f16(s z - hi), s = 2^10 so that
// the low part stays out of the subnormal range; the epilogue unscales by 2^-20): 22 instead of 21 significant bits,
// half the operand bytes and twice the MMA rate of the TF32 form.  FP32 accumulation with the same chunked promotion.
constexpr double kF16Unscale = 1.0 / 1048576.0;

// Developer instrumentation (ablation modes of Params::dbg inside the lean epilogue, clock64 phase counters)
// costs registers in the hot loops, so it is compiled in only with -DBIXCOR_UMMA_PROFILE (scripts/perf_probe.py,
// profiles/umma_probe_r01.txt were produced that way).
// Lean epilogue of interior packed tiles: through the swizzled shared-memory transpose (default: every store
// instruction writes two full 128-byte runs) or straight from the tcgen05.ld.16x256b fragment layout
// (-DBIXCOR_UMMA_LEAN_FRAG: no STS/LDS, but a store instruction then touches 8 rows x 64 bytes and the kernel
// is 17% slower - measured, profiles/umma_probe_r01.txt; kept for that ablation).
#ifdef BIXCOR_UMMA_LEAN_FRAG
constexpr bool kFragEpilogue = true;
#else
constexpr bool kFragEpilogue = false;
#endif
// int8 kind with N = 256 MMAs (-DBIXCOR_UMMA_I8_WIDE=1): the B operand of one instruction is the hi plane AND the lo
// plane of the column genes (they are adjacent in shared memory), so a K step takes two MMAs (h x [h|l], l x [h|l])
// instead of four; every A plane is read from shared memory once instead of twice (-33 % operand reads per CTA of a
// pair), at the price of a fourth accumulator (hl and lh no longer share one): all 512 TMEM columns.
#ifndef BIXCOR_UMMA_I8_WIDE
#define BIXCOR_UMMA_I8_WIDE 0
#endif
constexpr bool kI8Wide = BIXCOR_UMMA_I8_WIDE != 0;
#ifndef BIXCOR_UMMA_TILE_SYNC_DEFAULT
#define BIXCOR_UMMA_TILE_SYNC_DEFAULT 1  // measured (run 45): 20k TOM GEMM f16 25.4 -> 23.4 ms, tf32 49.1 -> 42.7 ms
#endif
#ifdef BIXCOR_UMMA_PROFILE
constexpr bool kProfile = true;
#else
constexpr bool kProfile = false;
#endif

constexpr int BM = 128, BN = 128;
constexpr int ROW_BYTES = 128;                    // one SWIZZLE_128B row: 32 tf32 or 128 int8
constexpr int PLANE_TILE_BYTES = BM * ROW_BYTES;  // 16 KB
#ifndef BIXCOR_UMMA_EPI_WARPS
#define BIXCOR_UMMA_EPI_WARPS 8
#endif
constexpr int EPI_WARPS = BIXCOR_UMMA_EPI_WARPS;              // 8: two per scheduler, (TMEM lane quarter) x (column half); 16: four
// CTAS = 1: one CTA per 128 x 128 tile (cta_group::1).  CTAS = 2: a CTA pair (one TPC) shares a 256 x 128
// super-tile with tcgen05.mma.cta_group::2: each CTA stages its own 128 rows of A and HALF of the B tile, the
// tensor cores of both SMs read both halves.  Per-SM operand staging drops from 64 KB to 48 KB per K block (L2
// and TMA traffic -25 %).  Measured (profiles/umma_probe_r01.txt): +3 % on the packed correlation, a win on the
// Ozaki passes whose operands stream from HBM/L2, a loss on the long-K TF32 TOM (cross-CTA drain handshake).
constexpr int THREADS = 128 + 32 * EPI_WARPS;
constexpr int EPI_COLS = BN / (EPI_WARPS / 4);                // accumulator columns owned by one epilogue warp
constexpr int EPI_CHUNK = 16;                                 // columns transposed per step
// Lean epilogue with bulk-async stores (-DBIXCOR_UMMA_BULK=<columns per store: 16 / 32 / 64>): every lane finishes
// BULK_COLS values of ITS OWN row, parks them in shared memory and hands the run to the copy engine
// (cp.async.bulk.global.shared::cta), so the warp issues neither the transposing LDS nor the STG of the
// default lean epilogue.  Row layout in shared memory: 16 bytes of lead padding + the run; a row whose
// global address is 8 (mod 16) sits 8 bytes earlier so that elements 1 .. BULK_COLS-2 are 16-byte aligned on
// both sides, its first and last element go out as scalar stores.
#ifdef BIXCOR_UMMA_BULK
constexpr int BULK_COLS = BIXCOR_UMMA_BULK;
#ifdef BIXCOR_UMMA_BULK_NBUF
constexpr int BULK_NBUF = BIXCOR_UMMA_BULK_NBUF;
#else
constexpr int BULK_NBUF = 1;
#endif
#else
constexpr int BULK_COLS = 0;
constexpr int BULK_NBUF = 0;
#endif
constexpr int BULK_ROW_BYTES = BULK_COLS * 8 + 16;
constexpr int BULK_BUF_BYTES = 32 * BULK_ROW_BYTES;
constexpr int BULK_WARP_BYTES = BULK_COLS ? BULK_NBUF * BULK_BUF_BYTES + EPI_COLS * 8 /*column factors*/ : 0;
constexpr int EPI_WARP_STAGE_BYTES = (BULK_WARP_BYTES > 32 * EPI_CHUNK * 8) ? BULK_WARP_BYTES : 32 * EPI_CHUNK * 8;
constexpr int EPI_STAGE_BYTES = EPI_WARPS * EPI_WARP_STAGE_BYTES;  // one 32 x 16 double transpose buffer per warp (or the bulk rows)
constexpr int SMEM_LIMIT = 232448;  // 227 KB per CTA
template <int CTAS>
struct Cfg {
  static constexpr int B_TILE_BYTES = PLANE_TILE_BYTES / CTAS;
  static constexpr int STAGE_BYTES = 2 * PLANE_TILE_BYTES + 2 * B_TILE_BYTES;  // A hi, A lo, B hi, B lo
  static constexpr int FIT = (SMEM_LIMIT - 1024 - 256 - EPI_STAGE_BYTES) / STAGE_BYTES;  // smem left after the epilogue staging
  static constexpr int STAGES = FIT > 4 ? 4 : FIT;
};
static_assert(Cfg<1>::STAGES >= 2 && Cfg<2>::STAGES >= 2, "operand pipeline needs two stages");
template <int CTAS>
constexpr int smem_bytes_of() {
  return Cfg<CTAS>::STAGES * Cfg<CTAS>::STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + EPI_STAGE_BYTES;
}
constexpr int TILE_SECOND_INVALID = 1 << 30;  // CTAS = 2: flag on tile.y, the pair's second tile holds nothing
// Flag on tile.y (packed correlation, interior tiles): the tile is computed TRANSPOSED - the column genes are the M
// operand (TMEM lanes), the row genes the N operand (TMEM columns).  A warp then holds, for one output row, 32
// consecutive columns in its 32 lanes and writes 256 contiguous bytes per store straight from registers: no
// shared-memory transpose at all.  A CTA pair covers one block row x TWO consecutive block columns (tile.y, tile.y + 1).
constexpr int TILE_SWAPPED = 1 << 29;
// Compile-time option (-DBIXCOR_UMMA_SWAP=1); the default keeps interior tiles untransposed with the shared-memory
// transpose of packed_interior.  Measured (profiles/umma_probe_r01.txt, run 36): correct and bit-identical on the int8
// path, but no faster (0.834 vs 0.814 ms) - the epilogue is not limited by its shared-memory traffic or LSU
// instruction count.  Only one of the two lean paths is compiled so that the kernel stays within its registers.
#ifndef BIXCOR_UMMA_SWAP
#define BIXCOR_UMMA_SWAP 0
#endif
constexpr bool kSwapInterior = BIXCOR_UMMA_SWAP != 0;
constexpr int TILE_FLAGS = TILE_SECOND_INVALID | TILE_SWAPPED;

struct Params {
  int ntiles;
  const int2* tiles;  // CTAS = 1: (block row, block col); CTAS = 2: (first block row of the pair, block col | flags)
  int npass;         // 2 for the differential correlation (operand set 1 in pass 1)
  int kblocks[2];    // 128-byte K blocks per pass
  const double* inv_norm[2];  // KIND_I8: 1/|d| per gene and pass
  unsigned long long* prof;  // optional (env BIXCOR_UMMA_PROF): per epilogue warp cycles {wait, drain, lean epilogue, tiles}
  int plane_a0, plane_b0;  // first digit plane of the row / column operand inside the gene record (0 unless Ozaki groups)
  int zero;  // always 0; opaque to the compiler (scheduling fence, see the epilogue drain)
  int tile_sync_kb;     // > 0: additionally re-align every tile_sync_kb K blocks (env BIXCOR_UMMA_TILE_SYNC_KB, experiment)
  unsigned* tile_sync;  // long-K launches: global arrival counter, the TMA producers of all CTAs start every tile together
                        // so that CTAs sharing an operand panel stream it through L2 in step (nullptr = off)
  int prefetch_kb;  // long-K launches: the producer asks the TMA unit to pull the operand boxes of K block kb + prefetch_kb into L2
  int mc_nb, mc_rb0, mc_rb1, mc_upper;  // MC kernels: number of block columns, block rows of this launch, upper-triangle-only
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

// ---- cluster variants used by the CTA-pair kernel ----
// Default (.release.cta / .acquire.cta) semantics on purpose, like CUTLASS' ClusterBarrier: a .cluster-scope
// release compiles to MEMBAR.ALL.GPU in front of every arrive and an acquire to CCTL.IVALL after every wait,
// which cost the producer more than a K block.  What these barriers order is async-proxy traffic (TMA bytes,
// tcgen05 reads/writes), which complete_tx and the tcgen05 fences cover.
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_cluster(uint32_t cluster_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_addr), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) __trap();
  }
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load of a CTA pair member: data lands in the executing CTA's shared memory, the completion bytes are
// signalled on `bar_cluster`, a shared::cluster address that may belong to the peer (leader) CTA.
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1,
                                                 int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {  // arrives on `bar` in BOTH CTAs of the pair
  asm volatile(
      "{\n\t.reg .b16 m;\n\tmov.b16 m, 3;\n\t"
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}"
      ::"r"(bar)
      : "memory");
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// L2 prefetch of one box (no shared memory, no barrier): the long-K kernels run all CTAs in step along K, so the first CTA to
// ask for a K slice misses to DRAM and its neighbours queue behind the same miss - every stage then waits a DRAM round trip
// (~2 us) with only 128-144 KB in flight per SM.  Prefetching a few K blocks ahead turns those loads into L2 hits.
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// Multicast TMA load (cluster of single CTAs): the box lands at the same shared-memory offset in every CTA of `mask` and
// each destination CTA's own mbarrier (same offset) receives the complete_tx bytes.
__device__ __forceinline__ void tma_load_3d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "h"(mask)
      : "memory");
}
// tcgen05.commit arriving on the barrier at this offset in every CTA of `mask` (the CTAs whose TMA multicasts write this
// CTA's stage wait for its MMAs to have read it)
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

#define BIXCOR_TC_MMA(GROUP, KINDSTR)                                                                  \
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"                                       \
               "tcgen05.mma.cta_group::" GROUP ".kind::" KINDSTR " [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), \
               "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)                                   \
               : "memory")
template <int KIND, int CTAS = 1>
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  if (CTAS == 2) {
    if (KIND == KIND_TF32) BIXCOR_TC_MMA("2", "tf32");
    else if (KIND == KIND_F16) BIXCOR_TC_MMA("2", "f16");
    else BIXCOR_TC_MMA("2", "i8");
  } else {
    if (KIND == KIND_TF32) BIXCOR_TC_MMA("1", "tf32");
    else if (KIND == KIND_F16) BIXCOR_TC_MMA("1", "f16");
    else BIXCOR_TC_MMA("1", "i8");
  }
}
#undef BIXCOR_TC_MMA

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
// 16 lanes x 256 bits, 4 repeats = 16 lanes x 32 columns.  Thread t receives, for repeat n, r[4n+0..1] =
// (lane t/4, columns 8n + 2(t%4) + {0,1}) and r[4n+2..3] = (lane t/4 + 8, same columns): the accumulator
// fragment layout of mma, i.e. every quad holds 8 consecutive columns of a row.
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// two consecutive doubles of one packed row: one 16-byte store when the pair is 16-byte aligned, two 8-byte
// stores otherwise (predicated, no branch: rows of one warp instruction have mixed parity)
__device__ __forceinline__ void store_pair(char* addr, double v0, double v1, uint32_t odd) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %3, 0;\n\t"
      "@!p st.global.v2.f64 [%0], {%1, %2};\n\t"
      "@p st.global.f64 [%0], %1;\n\t"
      "@p st.global.f64 [%0+8], %2;\n\t}"
      ::"l"(addr), "d"(v0), "d"(v1), "r"(odd)
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

// instruction descriptor: dense, K-major A and B, M = 128, N = 128 (256 for the wide int8 form)
template <int KIND, int CTAS = 1, int N = BN>
__host__ __device__ constexpr uint32_t make_idesc() {
  return (KIND == KIND_TF32 ? (1u << 4) | (2u << 7) | (2u << 10)    // D = F32, A = B = TF32
          : KIND == KIND_F16 ? (1u << 4)                            // D = F32, A = B = F16 (format code 0)
                             : (2u << 4) | (1u << 7) | (1u << 10))  // D = S32, A = B = signed int8
         | ((uint32_t)(N >> 3) << 17) | ((uint32_t)((BM * CTAS) >> 4) << 24);  // M = 256 across the CTA pair
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
// FP32 promotion interval of the f16 kind in K blocks of 64 samples.  Measured on the 20k TOM (run 40): 2 -> 29.9 ms,
// max deviation 3.2e-6; 4 -> 26.0 ms, 6.6e-6.  3 gives the accuracy of the 3xTF32 path (4.7e-6) that this kind replaces.
#ifndef BIXCOR_F16_CHUNK_KB
#define BIXCOR_F16_CHUNK_KB 3
#endif
template <>
struct KindTraits<KIND_F16> {
  static constexpr int NACC = 1;
  static constexpr int ACC_STAGES = 2;
  static constexpr int CHUNK_KB = BIXCOR_F16_CHUNK_KB;
};
template <>
struct KindTraits<KIND_I8> {
  static constexpr int NACC = kI8Wide ? 4 : 3;  // hh, hl+lh, ll (exact int32); wide form: hh, hl, lh, ll
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
      if (kProfile && KIND == KIND_I8 && (dbg == 7 || dbg == 10)) {  // profiling: no FP64 instruction at all (values are garbage)
        v0 = __hiloint2double(0x43300000, (int)racc[c + q]);
        v1 = __hiloint2double(0x43300000, (int)racc[c + q + 1]);
      } else if (KIND == KIND_I8) {
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
    if (kProfile && dbg == 5) {  // profiling: no phase B at all
      __syncwarp();
      continue;
    }
    if (kProfile && dbg == 9) {  // profiling: upper bound of a 128-bit phase B (LDS.128 + STG.128, addresses forced to
                                 // 16-byte alignment, so the VALUES LAND IN THE WRONG PLACE for odd rows: timing only)
      const int r4 = lane >> 3, u = lane & 7;
#pragma unroll
      for (int k4 = 0; k4 < 8; ++k4) {
        const int m = 4 * k4 + r4;
        double v0, v1;
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v0), "=d"(v1)
                     : "r"(stg_s + (uint32_t)(m * 128) + ((uint32_t)(u ^ (m & 7)) << 4)));
        if (KIND == KIND_I8) {
          v0 *= cf;
          v1 *= cf;
        }
        if (PW6) {
          const double s0 = v0 * v0, s1 = v1 * v1;
          v0 = s0 * s0 * s0;
          v1 = s1 * s1 * s1;
        }
        const uint64_t ga = ((uint64_t)(wbase + ((uint32_t)(m * a - (m * (m - 1)) / 2 + 2 * u) << 3) + (uint32_t)(c * 8))) & ~15ull;
        asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(ga), "d"(v0), "d"(v1) : "memory");
      }
      __syncwarp();
      continue;
    }
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) {  // phase B (lane = 2 rows x 16 columns): two 128-byte runs per store
      double v = *reinterpret_cast<const double*>(sb + ld_off[k2 & 3] + k2 * (2 * EPI_CHUNK * 8));
      if (KIND == KIND_I8 && !(kProfile && (dbg == 7 || dbg == 10))) v *= cf;
      if (PW6 && !(kProfile && (dbg == 7 || dbg == 10))) {
        const double sq = v * v;
        v = sq * sq * sq;
      }
      if (kProfile && (dbg == 4 || dbg == 10)) {  // profiling: everything but the global store (10: and no FP64)
        asm volatile("" ::"d"(v));
        continue;
      }
      *reinterpret_cast<double*>(wbase + roff[k2] + (uint32_t)(c * 8)) = v;
    }
    __syncwarp();
  }
}

// Interior tile of the packed correlation in the int32 transport form (EpiParams::out_s32): the exact integer Gram values
// of this warp's 32 rows x EPI_COLS columns go out as they are.  Lane = row while the accumulators sit in registers; each
// 32 x 32 block is transposed through the warp's staging (16-byte units XOR-swizzled by the row, conflict free) so that a
// store instruction writes one 128-byte run of one packed row.  wbase32: element (row_base, col_base) of the int32 vector;
// a = p - row_base - shift - 1 (packed length step between consecutive rows).
__device__ __forceinline__ void packed_interior_s32(const uint32_t (&racc)[EPI_COLS], double* stg, int lane, int* wbase32, int a) {
  const uint32_t stg_s = smem_u32(stg);
#pragma unroll
  for (int c0 = 0; c0 < EPI_COLS; c0 += 32) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {  // phase A: my row's 32 values as eight 16-byte units
      const uint32_t addr = stg_s + (uint32_t)lane * 128u + ((((uint32_t)u) ^ ((uint32_t)lane & 7u)) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(racc[c0 + 4 * u]), "r"(racc[c0 + 4 * u + 1]),
                   "r"(racc[c0 + 4 * u + 2]), "r"(racc[c0 + 4 * u + 3])
                   : "memory");
    }
    __syncwarp();
    uint32_t off = 0;  // element offset of row m relative to row 0 of the tile (running sum: a, a - 1, ...)
#pragma unroll
    for (int m = 0; m < 32; ++m) {  // phase B: lane = column; one row per store instruction
      uint32_t v;
      const uint32_t addr = stg_s + (uint32_t)m * 128u + (((((uint32_t)lane >> 2) ^ ((uint32_t)m & 7u)) << 4) | (((uint32_t)lane & 3u) << 2));
      asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
      wbase32[off + (uint32_t)(c0 + lane)] = (int)v;
      off += (uint32_t)(a - m);
    }
    __syncwarp();
  }
}

// Lean epilogue of one transposed (TILE_SWAPPED) interior tile for one epilogue warp: lane = output column, racc[q] =
// output row rb + q.  rowv: 1/|d| of rows rb .. rb + EPI_COLS (KIND_I8), read as warp-uniform 16-byte loads; colf: the
// lane's column factor; gp: the lane's element of row rb; a = packed length step ro(rb + 1) - ro(rb).
template <int KIND, bool PW6>
__device__ __forceinline__ void packed_swapped(const uint32_t (&racc)[EPI_COLS], double colf, double scale, bool scaled,
                                               const double* rowv, char* gp, int a) {
  uint32_t off = 0;  // byte offset of row rb + q relative to row rb (warp uniform, < 2^32: 64 rows of at most 2^24 columns)
#pragma unroll
  for (int q = 0; q < EPI_COLS; q += 2) {
    double v0, v1;
    uint32_t x0 = racc[q], x1 = racc[q + 1];
    asm volatile("" : "+r"(x0), "+r"(x1));  // ordered against the offset chain below: conversions are not hoisted
    if (KIND == KIND_I8) {
      const double2 rf = __ldg(reinterpret_cast<const double2*>(rowv) + (q >> 1));
      v0 = (double)(int)x0 * (scale * rf.x) * colf;  // same association as the untransposed paths: bit-identical
      v1 = (double)(int)x1 * (scale * rf.y) * colf;
    } else {
      v0 = (double)__uint_as_float(x0);
      v1 = (double)__uint_as_float(x1);
      if (scaled) {
        v0 *= scale;
        v1 *= scale;
      }
    }
    if (PW6) {
      const double s0 = v0 * v0, s1 = v1 * v1;
      v0 = s0 * s0 * s0;
      v1 = s1 * s1 * s1;
    }
    *reinterpret_cast<double*>(gp + off) = v0;
    off += (uint32_t)(a - q) << 3;
    asm volatile("" : "+r"(off));  // keep the offsets a running sum: ptxas otherwise precomputes all 64 of them
    *reinterpret_cast<double*>(gp + off) = v1;
    off += (uint32_t)(a - q - 1) << 3;
    asm volatile("" : "+r"(off));
  }
}

__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// Lean epilogue of one interior tile with bulk-async stores (see BULK_COLS above): lane = row for the whole tile.
// stg_s: this warp's staging (BULK_NBUF row buffers, then EPI_COLS column factors); grow: this lane's packed row at
// the warp's first column; buf: round-robin buffer index, kept across tiles.
template <int KIND, bool PW6>
__device__ __forceinline__ void packed_interior_bulk(const uint32_t (&racc)[EPI_COLS], uint32_t stg_s, int lane, double rowf,
                                                     bool tf32_scaled, const double* colv, char* grow, uint32_t& buf) {
  constexpr int BC = BULK_COLS ? BULK_COLS : 16;
  const uint32_t colf_s = stg_s + (uint32_t)(BULK_NBUF * BULK_BUF_BYTES);
  if (KIND == KIND_I8) {  // 1/|d_j| of the warp's 64 columns: two per lane, read back as broadcast LDS.128
    const double2 cf = __ldg(reinterpret_cast<const double2*>(colv) + lane);
    __syncwarp();
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(colf_s + (uint32_t)lane * 16u), "d"(cf.x), "d"(cf.y) : "memory");
    __syncwarp();
  }
  const uint32_t odd = ((uint32_t)(uintptr_t)grow >> 3) & 1u;
  const uint32_t row_s = stg_s + (uint32_t)lane * BULK_ROW_BYTES;
#pragma unroll
  for (int c = 0; c < EPI_COLS; c += BC) {
    const uint32_t bs = row_s + buf * (uint32_t)BULK_BUF_BYTES;
    // the copy that last read this buffer is done with it
    if (BULK_NBUF <= 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    else asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(BULK_NBUF > 1 ? BULK_NBUF - 1 : 0) : "memory");
    const uint32_t st0 = bs + (odd ? 0u : 16u);
    double carry = 0.0, first = 0.0;
#pragma unroll
    for (int q = 0; q < BC; q += 2) {
      double v0, v1;
      if (KIND == KIND_I8) {
        double c0, c1;
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(c0), "=d"(c1) : "r"(colf_s + (uint32_t)((c + q) * 8)));
        v0 = (double)(int)racc[c + q] * rowf * c0;
        v1 = (double)(int)racc[c + q + 1] * rowf * c1;
      } else {
        v0 = (double)__uint_as_float(racc[c + q]);
        v1 = (double)__uint_as_float(racc[c + q + 1]);
        if (tf32_scaled) {
          v0 *= rowf;
          v1 *= rowf;
        }
      }
      if (PW6) {
        const double s0 = v0 * v0, s1 = v1 * v1;
        v0 = s0 * s0 * s0;
        v1 = s1 * s1 * s1;
      }
      if (q == 0) first = v0;
      // even row: (v[q], v[q+1]) at 16 + 8q; odd row: (v[q-1], v[q]) at 8q (the q = 0 pair fills the lead padding)
      const double a = odd ? carry : v0, b = odd ? v0 : v1;
      asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(st0 + (uint32_t)(q * 8)), "d"(a), "d"(b) : "memory");
      carry = v1;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    char* const g = grow + c * 8;
    if (odd) {
      *reinterpret_cast<double*>(g) = first;
      *reinterpret_cast<double*>(g + (BC - 1) * 8) = carry;
    }
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g + (odd ? 8 : 0)), "r"(bs + 16u),
                 "r"(odd ? (uint32_t)((BC - 2) * 8) : (uint32_t)(BC * 8))
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    if (BULK_NBUF > 1) buf = (buf + 1 == (uint32_t)BULK_NBUF) ? 0u : buf + 1;
  }
}

// map0 / map1: 128-row boxes (A operand of pass 0 / 1); mapb0 / mapb1: the B operand boxes (128 rows for CTAS = 1,
// 64 rows = this CTA's half of the B tile for CTAS = 2).
// MC = 1 (CTAS must be 1; ablation, off by default - env BIXCOR_UMMA_MC=1): clusters of 2 x 2 CTAs work on 2 x 2 neighbouring tiles.  The A operand of a block row is needed
// by the two CTAs of that cluster row, the B operand of a block column by the two CTAs of that cluster column: every CTA
// loads HALF of its A tile and HALF of its B tile and multicasts each half to the two CTAs that need it, so a CTA pulls 32 KB
// per K block out of L2 instead of 64 KB (a CTA pair: 48 KB).  ncu on the 20k TOM: 195 GB (f16, pairs) / 511 GB (TF32) of
// L2 -> SM reads for a 1.6 / 3.2 GB operand, 8.9 TB/s sustained: the long-K kernels are bound by that path.  Measured: correct,
// but no faster (profiles/README_r02.md) - the L2 already merges concurrent requests of a few SMs for the same lines, so a 2-way
// multicast removes no traffic from the limiting L2 -> SM path; only more flops per operand byte (256-wide tiles) would.
// WIDE = 1 (CTA pairs, split-float kinds, EPI_TOM): 256 x 256 tiles per pair.  Each CTA stages its 128 rows of A and 128 of
// the 256 B rows, one MMA covers M = 256 x N = 256: operand bytes per flop drop by a third against the 256 x 128 pair tile
// (the long-K kernels sit on the L2 -> SM operand path, profiles/README_r02.md).  The two accumulator stages of the chunked
// promotion take all 512 TMEM columns; 16 epilogue warps (640 threads, 96 registers) keep the 64 running sums per thread;
// the epilogue stores straight from the lane = row layout (no shared-memory staging: the smem holds 3 stages of 64 KB).
constexpr int WIDE_EPI_WARPS = 16;
constexpr int WIDE_THREADS = 128 + 32 * WIDE_EPI_WARPS;
constexpr int WIDE_STAGES = 3;
constexpr int WIDE_STAGE_BYTES = 4 * PLANE_TILE_BYTES;
constexpr int wide_smem_bytes() { return WIDE_STAGES * WIDE_STAGE_BYTES + 1024 + 256; }

// S32 = 1 (KIND_I8, EPI_PACKED): the int32 transport form of the exact Spearman Gram (EpiParams::out_s32) as its own
// instantiation, so that the f64 epilogue keeps its register allocation (sharing one kernel cost it 11 %: 0.81 -> 0.89 ms).
template <int KIND, int EPI, int CTAS, int MC = 0, int WIDE = 0, int S32 = 0>
__global__ void __launch_bounds__(WIDE ? WIDE_THREADS : THREADS, 1)
gram_umma_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1,
                 const __grid_constant__ CUtensorMap mapb0, const __grid_constant__ CUtensorMap mapb1, const Params P) {
  using T = KindTraits<KIND>;
  static_assert(!WIDE || (CTAS == 2 && !MC && KIND != KIND_I8 && EPI == EPI_TOM), "wide tiles: CTA pairs, split-float kinds, TOM epilogue");
  static_assert(!S32 || (KIND == KIND_I8 && EPI == EPI_PACKED && !MC && !WIDE), "int32 transport: exact int8 kind, packed epilogue");
  constexpr int BNT = WIDE ? 2 * BN : BN;                      // tile width (accumulator columns)
  constexpr int EPIW = WIDE ? WIDE_EPI_WARPS : EPI_WARPS;      // epilogue warps: (TMEM lane quarter) x (64-column slice)
  constexpr int TMEM_COLS = (T::NACC * T::ACC_STAGES * BNT <= 256) ? 256 : 512;
  static_assert(T::NACC * T::ACC_STAGES * BNT <= 512, "TMEM has 512 columns");
  constexpr int STAGES = WIDE ? WIDE_STAGES : Cfg<CTAS>::STAGES;
  constexpr int STAGE_BYTES = WIDE ? WIDE_STAGE_BYTES : Cfg<CTAS>::STAGE_BYTES;
  constexpr int B_TILE_BYTES = WIDE ? PLANE_TILE_BYTES : Cfg<CTAS>::B_TILE_BYTES;
  static_assert(!MC || CTAS == 1, "the multicast clusters are made of single CTAs");
  constexpr int CL = MC ? 4 : CTAS;  // CTAs per scheduling unit (cluster size)
  uint32_t cta_rank = 0;
  if (CL > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
  const int mc_r = MC ? (int)(cta_rank >> 1) : 0, mc_c = MC ? (int)(cta_rank & 1u) : 0;  // position in the 2 x 2 cluster
  const uint16_t mc_mask_a = (uint16_t)(0x3u << (2 * mc_r));                  // the CTAs of my cluster row (same block row)
  const uint16_t mc_mask_b = (uint16_t)((1u << mc_c) | (1u << (mc_c + 2)));   // the CTAs of my cluster column (same block column)
  const int unit0 = (int)blockIdx.x / CL, unit_stride = (int)gridDim.x / CL;  // tile (pair / cluster) scheduling units
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
      mbar_init(full_bar(s), CTAS);  // one expect_tx arrive per CTA of the pair (the leader's barrier collects both)
      mbar_init(empty_bar(s), MC ? 3 : 1);  // MC: my MMAs and those of my row / column mates, whose stages my multicasts fill
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), EPIW * CTAS);  // one arrive per epilogue warp (of both CTAs, on the leader)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    if (CTAS == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {  // one warp of EACH CTA of the pair takes part in the collective allocation
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (CL > 1) cluster_sync_all();  // peer barriers are initialised before any remote arrive / TMA signal
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int n_iter = (P.ntiles + unit_stride - 1) / unit_stride;  // the same for every CTA (barrier participation)
      bool sync_alive = P.tile_sync != nullptr;
      unsigned epoch = 0;  // barriers passed; every producer of the grid executes the same sequence of them
      // The barrier is a performance hint, never a correctness dependency: it assumes the grid's CTAs are co-resident
      // (one CTA per SM, grid <= SM count, checked by the host against the occupancy API).  If foreign work holds SMs
      // (another stream, MPS) a producer gives up after ~5 ms (a tile takes < 1 ms), counts the bail-out in
      // tile_sync[1] and every producer of the launch then runs un-synchronised.
      auto grid_barrier = [&]() {
        if (!sync_alive) return;
        atomicAdd(P.tile_sync, 1u);
        const unsigned target = ++epoch * gridDim.x;
        unsigned seen, spins = 0;
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(P.tile_sync) : "memory");
          if (seen >= target) break;
          __nanosleep(40);
        } while (++spins < (1u << 17));
        if (seen < target) {
          sync_alive = false;
          atomicAdd(P.tile_sync + 1, 1u);
        }
      };
      for (int it = 0; it < n_iter; ++it) {
        const int t = unit0 + it * unit_stride;
        grid_barrier();
        if (t >= P.ntiles) {  // no tile in the last round: still take part in the optional mid-tile barriers
          if (sync_alive && P.tile_sync_kb > 0)
            for (int pass = 0; pass < P.npass; ++pass)
              for (int kb = P.tile_sync_kb; kb < P.kblocks[pass]; kb += P.tile_sync_kb) grid_barrier();
          continue;
        }
        const int2 tile = P.tiles[t];
        const int ty = tile.y & ~TILE_FLAGS;
        // M operand rows (this CTA's 128) and N operand rows (the whole tile, or this CTA's half of it)
        const int a_row = MC ? (tile.x + mc_r) * BM : ((tile.y & TILE_SWAPPED) ? ty + (int)cta_rank : tile.x + (int)cta_rank) * BM;
        const int b_row = MC ? (ty + mc_c) * BN : ((tile.y & TILE_SWAPPED) ? tile.x : ty) * BNT + (CTAS == 2 ? (int)cta_rank * (BNT / 2) : 0);
        for (int pass = 0; pass < P.npass; ++pass) {
          const CUtensorMap* map = pass ? &map1 : &map0;
          const CUtensorMap* mapb = pass ? &mapb1 : &mapb0;
          const int kb_n = P.kblocks[pass];
          const int kelems = (KIND == KIND_TF32) ? 32 : (KIND == KIND_F16 ? 64 : 128);
          // rows / maps of the boxes this CTA loads itself (they are also what it prefetches)
          const CUtensorMap* pmap_a = (MC ? mapb : map);
          const int pa_row = MC ? a_row + 64 * mc_c : a_row, pb_row = MC ? b_row + 64 * mc_r : b_row;
          auto prefetch_kb = [&](int kq) {
            tma_prefetch_3d(pmap_a, kq * kelems, P.plane_a0, pa_row);
            tma_prefetch_3d(pmap_a, kq * kelems, P.plane_a0 + 1, pa_row);
            tma_prefetch_3d(mapb, kq * kelems, P.plane_b0, pb_row);
            tma_prefetch_3d(mapb, kq * kelems, P.plane_b0 + 1, pb_row);
          };
          const int pf = P.prefetch_kb;
          if (pf > 0)
            for (int kq = 0; kq < pf && kq < kb_n; ++kq) prefetch_kb(kq);
          for (int kb = 0; kb < kb_n; ++kb) {
            if (P.tile_sync_kb > 0 && kb > 0 && kb % P.tile_sync_kb == 0) grid_barrier();  // (experiment: re-align mid tile)
            if (pf > 0 && kb + pf < kb_n) prefetch_kb(kb + pf);
            mbar_wait(empty_bar(stage), phase ^ 1u);
            const uint32_t s0 = base + stage * STAGE_BYTES;
            if (CTAS == 2) {
              // the leader's full barrier counts the bytes of both CTAs; B rows: this CTA's half of the tile
              const uint32_t fb = mapa_rank(full_bar(stage), 0);
              // profiling (dbg 11 / 12): leave out one / two of the four operand boxes - the results are garbage, the time
              // tells how much of the K loop is operand traffic (profiles/README_r02.md)
              const int skip = (kProfile && (P.dbg == 11 || P.dbg == 12)) ? P.dbg - 10 : 0;
              mbar_expect_tx_cluster(fb, STAGE_BYTES - (skip >= 1 ? B_TILE_BYTES : 0) - (skip >= 2 ? PLANE_TILE_BYTES : 0));
              tma_load_3d_pair(s0 + 0 * PLANE_TILE_BYTES, map, fb, kb * kelems, P.plane_a0, a_row);
              if (skip < 2) tma_load_3d_pair(s0 + 1 * PLANE_TILE_BYTES, map, fb, kb * kelems, P.plane_a0 + 1, a_row);
              tma_load_3d_pair(s0 + 2 * PLANE_TILE_BYTES, mapb, fb, kb * kelems, P.plane_b0, b_row);
              if (skip < 1) tma_load_3d_pair(s0 + 2 * PLANE_TILE_BYTES + B_TILE_BYTES, mapb, fb, kb * kelems, P.plane_b0 + 1, b_row);
            } else if (MC) {
              // my halves of the A / B tiles (64 rows each), multicast to the row / column mates; the other halves arrive from them
              mbar_expect_tx(full_bar(stage), STAGE_BYTES);
              const uint32_t ha = (uint32_t)mc_c * (PLANE_TILE_BYTES / 2), hb = (uint32_t)mc_r * (PLANE_TILE_BYTES / 2);
              tma_load_3d_mc(s0 + 0 * PLANE_TILE_BYTES + ha, mapb, full_bar(stage), kb * kelems, P.plane_a0, a_row + 64 * mc_c, mc_mask_a);
              tma_load_3d_mc(s0 + 1 * PLANE_TILE_BYTES + ha, mapb, full_bar(stage), kb * kelems, P.plane_a0 + 1, a_row + 64 * mc_c, mc_mask_a);
              tma_load_3d_mc(s0 + 2 * PLANE_TILE_BYTES + hb, mapb, full_bar(stage), kb * kelems, P.plane_b0, b_row + 64 * mc_r, mc_mask_b);
              tma_load_3d_mc(s0 + 3 * PLANE_TILE_BYTES + hb, mapb, full_bar(stage), kb * kelems, P.plane_b0 + 1, b_row + 64 * mc_r, mc_mask_b);
            } else if (kProfile && P.dbg == 6) {  // profiling: no operand traffic at all
              mbar_arrive(full_bar(stage));
            } else {
              mbar_expect_tx(full_bar(stage), STAGE_BYTES);
              tma_load_3d(s0 + 0 * PLANE_TILE_BYTES, map, full_bar(stage), kb * kelems, P.plane_a0, a_row);
              tma_load_3d(s0 + 1 * PLANE_TILE_BYTES, map, full_bar(stage), kb * kelems, P.plane_a0 + 1, a_row);
              tma_load_3d(s0 + 2 * PLANE_TILE_BYTES, mapb, full_bar(stage), kb * kelems, P.plane_b0, b_row);
              tma_load_3d(s0 + 2 * PLANE_TILE_BYTES + B_TILE_BYTES, mapb, full_bar(stage), kb * kelems, P.plane_b0 + 1, b_row);
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
    // ============================ MMA issuer (leader CTA of a pair only) ===
    if (lane == 0 && (MC || cta_rank == 0)) {  // MC: every CTA of the cluster issues its own cta_group::1 MMAs
      constexpr uint32_t idesc = make_idesc<KIND, CTAS, BNT>();
      constexpr uint32_t idesc_wide = make_idesc<KIND, CTAS, 2 * BN>();
      int stage = 0, acc_stage = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int t = unit0; t < P.ntiles; t += unit_stride) {
        for (int pass = 0; pass < P.npass; ++pass) {
          const int kb_n = P.kblocks[pass];
          const int chunk_kb = (kProfile && P.dbg == 8) ? (1 << 30) : T::CHUNK_KB;  // dbg 8: profiling, no chunked promotion
          for (int kb0 = 0; kb0 < kb_n; kb0 += chunk_kb) {
            const int kb1 = (kb_n - kb0 > chunk_kb) ? kb0 + chunk_kb : kb_n;
            // epilogue (of both CTAs) has drained this accumulator set
            if (CTAS == 2) mbar_wait_cluster(tempty_bar(acc_stage), acc_phase ^ 1u);
            else mbar_wait(tempty_bar(acc_stage), acc_phase ^ 1u);
            tc_fence_after();
            const uint32_t acc0 = tmem_base + (uint32_t)(acc_stage * T::NACC * BNT);
            for (int kb = kb0; kb < kb1; ++kb) {
              if (CTAS == 2) mbar_wait_cluster(full_bar(stage), phase);
              else mbar_wait(full_bar(stage), phase);
              tc_fence_after();
              const uint32_t s0 = base + stage * STAGE_BYTES;
              const uint64_t dA0 = make_smem_desc(s0 + 0 * PLANE_TILE_BYTES);
              const uint64_t dA1 = make_smem_desc(s0 + 1 * PLANE_TILE_BYTES);
              const uint64_t dB0 = make_smem_desc(s0 + 2 * PLANE_TILE_BYTES);
              const uint64_t dB1 = make_smem_desc(s0 + 2 * PLANE_TILE_BYTES + B_TILE_BYTES);
#pragma unroll
              for (int k = 0; k < 4; ++k) {             // 4 x 32-byte K steps inside the 128-byte swizzle row
                const uint64_t ko = (uint64_t)(k * 2);   // +32 bytes, in 16-byte descriptor units
                const uint32_t first = (kb == kb0 && k == 0) ? 0u : 1u;
                if (kProfile && (P.dbg == 3 || P.dbg == 6)) continue;
                if (KIND != KIND_I8) {
                  tc_mma<KIND, CTAS>(acc0, dA0 + ko, dB0 + ko, idesc, first);  // hi*hi
                  tc_mma<KIND, CTAS>(acc0, dA0 + ko, dB1 + ko, idesc, 1u);     // hi*lo
                  tc_mma<KIND, CTAS>(acc0, dA1 + ko, dB0 + ko, idesc, 1u);     // lo*hi
                } else if (kI8Wide) {  // B = [hi plane | lo plane] of the column genes in one N = 256 operand
                  tc_mma<KIND, CTAS>(acc0, dA0 + ko, dB0 + ko, idesc_wide, first);           // h * [h | l]
                  tc_mma<KIND, CTAS>(acc0 + 2 * BN, dA1 + ko, dB0 + ko, idesc_wide, first);  // l * [h | l]
                } else {
                  tc_mma<KIND, CTAS>(acc0 + 0 * BN, dA0 + ko, dB0 + ko, idesc, first);  // h*h
                  tc_mma<KIND, CTAS>(acc0 + 1 * BN, dA0 + ko, dB1 + ko, idesc, first);  // h*l
                  tc_mma<KIND, CTAS>(acc0 + 1 * BN, dA1 + ko, dB0 + ko, idesc, 1u);     // l*h
                  tc_mma<KIND, CTAS>(acc0 + 2 * BN, dA1 + ko, dB1 + ko, idesc, first);  // l*l
                }
              }
              // smem stage reusable (in both CTAs) once these MMAs have read it
              if (CTAS == 2) tc_commit_pair(empty_bar(stage));
              else if (MC) tc_commit_mc(empty_bar(stage), (uint16_t)(mc_mask_a | mc_mask_b));
              else tc_commit(empty_bar(stage));
              if (++stage == STAGES) {
                stage = 0;
                phase ^= 1u;
              }
            }
            // chunk accumulators complete -> epilogue warps (of both CTAs)
            if (CTAS == 2) tc_commit_pair(tfull_bar(acc_stage));
            else tc_commit(tfull_bar(acc_stage));
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
    // Ozaki pass: int8 digit-pair Gram accumulated into an FP64 matrix, G += w * (128^2 hh + 128 (hl + lh) + ll)
    constexpr bool OZAKI = (KIND == KIND_I8 && EPI == EPI_ACCUM);
    constexpr bool DIRECT = (KIND == KIND_I8 && (EPI == EPI_PACKED || EPI == EPI_ACCUM));
    static_assert(EPI_CHUNK == 16, "the direct path loads 16 accumulator columns per step");
    const int ew = (warp - 4) & 3;
    const int ch = (warp - 4) >> 2;
    // TMEM columns (inside an accumulator stage) of this warp's hh / hl / lh / ll sums.  Three accumulators: hl and
    // lh share one.  Wide form: the N = 256 result of h x [h|l] sits in columns [0, 256), l x [h|l] in [256, 512); a
    // CTA pair splits B by column halves, so its result reads [hh(0:64) | hl(0:64) | hh(64:128) | hl(64:128)].
    static_assert(!kI8Wide || EPI_COLS == 64, "wide int8 column mapping assumes two column halves");
    const uint32_t o_hh = (KIND == KIND_I8 && kI8Wide && CTAS == 2) ? (uint32_t)(ch * 128) : (uint32_t)(ch * EPI_COLS);
    const uint32_t o_hl = o_hh + ((KIND == KIND_I8 && kI8Wide) ? (CTAS == 2 ? 64u : (uint32_t)BN) : (uint32_t)BN);
    const uint32_t o_lh = o_hh + 2u * BN;  // wide form only
    const uint32_t o_ll = (KIND == KIND_I8 && kI8Wide) ? o_hl + 2u * BN : o_hh + 2u * BN;
    int acc_stage = 0;
    uint32_t acc_phase = 0;
    const EpiParams& E = P.E;
    double* stg = epi_stage_all + (warp - 4) * (EPI_WARP_STAGE_BYTES / 8);
    uint32_t bulk_buf = 0;
    const bool want_mirror = epi_wants_mirror<EPI>(E);
    const int pp = (int)E.p;
    const int dbg = kProfile ? P.dbg : 0;
    // loop invariants pulled out of the constant bank once
    double* const out = E.out;
    const int64_t ldo = E.ldo;
    const int shift = E.shift;
    const int diag_one = E.diag_one;
    const int keep_sign = E.keep_sign;
    const double scale = E.scale * (KIND == KIND_F16 ? kF16Unscale : 1.0);  // f16 planes carry 2^10 each
    // 1 identity, 6 = WGCNA default, 0 generic (any other exponent, or an RBF transform)
    const int pw_mode = (E.rbf_mode || E.abs_out) ? 0 : ((E.beta == 1.0) ? 1 : (E.beta_int == 6 ? 6 : 0));
    long long prof_acc[4] = {0, 0, 0, 0};
    // tempty arrives go to the leader CTA's barrier (its MMA warp is the only consumer)
    auto arrive_tempty = [&](int st) {
      if (CTAS == 2) mbar_arrive_cluster(mapa_rank(tempty_bar(st), 0));
      else mbar_arrive(tempty_bar(st));
    };
    for (int t = unit0; t < P.ntiles; t += unit_stride) {
      int2 tile = P.tiles[t];
      bool tile_valid = !(CTAS == 2 && cta_rank == 1 && (tile.y & TILE_SECOND_INVALID));
      const bool swapped = kSwapInterior && (EPI == EPI_PACKED) && (tile.y & TILE_SWAPPED);
      tile.y &= ~TILE_FLAGS;
      if (CTAS == 2) (swapped ? tile.y : tile.x) += (int)cta_rank;  // this CTA's block column / block row of the pair
      if (MC) {  // this CTA's tile of the 2 x 2 super-tile; tiles outside the launch's rows / the matrix / the upper triangle
                 // still run the protocol (their operand halves are needed by the mates) but store nothing
        tile.x += mc_r;
        tile.y += mc_c;
        tile_valid = tile.x >= P.mc_rb0 && tile.x < P.mc_rb1 && tile.y < P.mc_nb && (!P.mc_upper || tile.x <= tile.y);
      }
      const int row_base = tile.x * BM + ew * 32;
      const int i = row_base + lane;
      const int col_base = tile.y * BNT + ch * EPI_COLS;
      for (int pass = 0; pass < P.npass; ++pass) {
        // Interior tiles of the packed correlation (97% of the tiles at p = 20000) lie strictly above the
        // diagonal and fully inside the matrix: nothing is predicated and they take the lean path below.
        const bool interior = !kSwapInterior && tile_valid && (EPI == EPI_PACKED) && tile.x < tile.y && tile.y * BN + BN <= pp && pw_mode != 0 && !keep_sign && (dbg == 0 || dbg >= 3);
        // DIRECT kernels keep the accumulators of the few non-interior tiles in TMEM and finish them chunk
        // by chunk (the MMA warp waits a little longer on 3% of the tiles); in exchange the 64-register
        // drain array only lives on the lean path and the kernel stays free of local-memory spills.
        const int kb_n = P.kblocks[pass];
        if (kFragEpilogue && EPI_COLS == 64 && interior) {
          // ---- (ablation) lean path without a shared-memory transpose: the accumulators are read in the
          //      16x256b fragment layout, a quad holds 8 consecutive columns of a row and stores them as
          //      16-byte pairs straight to the packed row.
          uint32_t sf[EPI_COLS];  // [(h*2 + g)*16 + 4n + 2k + e]: row 16h + 8k + lane/4, column 32g + 8n + 2(lane%4) + e
          for (int kb0 = 0; kb0 < kb_n; kb0 += T::CHUNK_KB) {
            mbar_wait(tfull_bar(acc_stage), acc_phase);
            tc_fence_after();
            const uint32_t tcol = (uint32_t)(acc_stage * T::NACC * BN);
            uint32_t dep = 0;
#pragma unroll
            for (int hg = 0; hg < 4; ++hg) {
              const int h = hg >> 1, g = hg & 1;
              const uint32_t ta = tmem_base + ((uint32_t)(ew * 32 + h * 16) << 16) + tcol + (uint32_t)(g * 32) + dep;
              if (KIND == KIND_I8) {
                uint32_t r0[16], r1[16], r2[16];
                tmem_ld_16x256b_x4(ta + o_hh, r0);
                tmem_ld_16x256b_x4(ta + o_hl, r1);
                if (kI8Wide) {
                  tmem_ld_16x256b_x4(ta + o_lh, r2);
                  tmem_ld_wait();
#pragma unroll
                  for (int q = 0; q < 16; ++q) r1[q] += r2[q];
                }
                tmem_ld_16x256b_x4(ta + o_ll, r2);
                tmem_ld_wait();
                uint32_t any = 0;
#pragma unroll
                for (int q = 0; q < 16; ++q) {  // S = 256*hh + 16*(hl+lh) + ll, exact in int32 for n <= 1860
                  sf[hg * 16 + q] = (uint32_t)(256 * (int)r0[q] + 16 * (int)r1[q] + (int)r2[q]);
                  any |= sf[hg * 16 + q];
                }
                dep = any & (uint32_t)P.zero;  // (scheduling fence, see the generic drain below)
              } else {
                uint32_t r0[16];
                tmem_ld_16x256b_x4(ta + o_hh, r0);
                tmem_ld_wait();
                uint32_t any = 0;
                if (kb0 == 0) {
#pragma unroll
                  for (int q = 0; q < 16; ++q) sf[hg * 16 + q] = r0[q];
                } else {
#pragma unroll
                  for (int q = 0; q < 16; ++q)
                    sf[hg * 16 + q] = __float_as_uint(__uint_as_float(sf[hg * 16 + q]) + __uint_as_float(r0[q]));
                }
#pragma unroll
                for (int q = 0; q < 16; ++q) any |= sf[hg * 16 + q];
                dep = any & (uint32_t)P.zero;
              }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive_tempty(acc_stage);
            if (++acc_stage == T::ACC_STAGES) {
              acc_stage = 0;
              acc_phase ^= 1u;
            }
          }
          const int q4 = lane & 3, rl = lane >> 2;
          const double* inv = (KIND == KIND_I8) ? P.inv_norm[pass] : nullptr;
          double rowf[4];
          char* rptr[4];
          uint32_t odd[4];
#pragma unroll
          for (int hk = 0; hk < 4; ++hk) {
            const int r = row_base + 16 * (hk >> 1) + 8 * (hk & 1) + rl;
            const int64_t idx0 = packed_row_offset(r, E.p, shift) - E.out_base - r - shift + col_base + 2 * q4;
            rptr[hk] = reinterpret_cast<char*>(out + idx0);
            odd[hk] = (uint32_t)(idx0 & 1);
            rowf[hk] = (KIND == KIND_I8) ? scale * __ldg(inv + r) : scale;
          }
#pragma unroll
          for (int g = 0; g < 2; ++g) {
#pragma unroll
            for (int n = 0; n < 4; ++n) {
              double c0 = 1.0, c1 = 1.0;
              if (KIND == KIND_I8) {
                const double2 cf = __ldg(reinterpret_cast<const double2*>(inv + col_base + 32 * g + 8 * n + 2 * q4));
                c0 = cf.x;
                c1 = cf.y;
              }
#pragma unroll
              for (int hk = 0; hk < 4; ++hk) {
                const int h = hk >> 1, k = hk & 1;
                const uint32_t s0 = sf[(h * 2 + g) * 16 + 4 * n + 2 * k], s1 = sf[(h * 2 + g) * 16 + 4 * n + 2 * k + 1];
                double v0, v1;
                if (KIND == KIND_I8) {
                  v0 = (double)(int)s0 * rowf[hk] * c0;
                  v1 = (double)(int)s1 * rowf[hk] * c1;
                } else {
                  v0 = (double)__uint_as_float(s0);
                  v1 = (double)__uint_as_float(s1);
                  if (scale != 1.0) {
                    v0 *= scale;
                    v1 *= scale;
                  }
                }
                if (pw_mode == 6) {
                  const double a0 = v0 * v0, a1 = v1 * v1;
                  v0 = a0 * a0 * a0;
                  v1 = a1 * a1 * a1;
                }
                store_pair(rptr[hk] + (32 * g + 8 * n) * 8, v0, v1, odd[hk]);
              }
            }
          }
          continue;
        }
        const bool direct = DIRECT && !interior && !(swapped && tile_valid);
        uint32_t racc[EPI_COLS];
        uint32_t tacc_direct = 0;
        const long long pt0 = (kProfile && P.prof) ? clock64() : 0;
        long long pt1 = pt0;
        if (direct) {
          mbar_wait(tfull_bar(acc_stage), acc_phase);
          tc_fence_after();
          tacc_direct = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc_stage * T::NACC * BN);
        } else {
          // ---- drain: TMEM -> registers, chunk by chunk (FP32 round-to-nearest / exact int32 sums)
          const int chunk_kb = (kProfile && dbg == 8) ? (1 << 30) : T::CHUNK_KB;
          for (int kb0 = 0; kb0 < kb_n; kb0 += chunk_kb) {
            mbar_wait(tfull_bar(acc_stage), acc_phase);
            if (kProfile && P.prof && kb0 == 0) pt1 = clock64();
            tc_fence_after();
            const uint32_t tacc = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc_stage * T::NACC * BNT);
            if (KIND == KIND_I8) {
              // `dep` is always 0 (P.zero is a runtime 0): it chains each batch of loads to the previous
              // batch's sums (all 16 of them, or ptxas only hurries the one it needs), otherwise ptxas hoists all 12 tcgen05.ld (192 registers) and spills.
              uint32_t dep = 0;
#pragma unroll
              for (int c = 0; c < EPI_COLS; c += 16) {
                uint32_t r0[16], r1[16], r2[16];
                tmem_ld16(tacc + o_hl + c + dep, r1);
                if (kI8Wide) {  // hl + lh first, so that only three batches are ever live
                  tmem_ld16(tacc + o_lh + c + dep, r2);
                  tmem_ld_wait();
#pragma unroll
                  for (int q = 0; q < 16; ++q) r1[q] += r2[q];
                }
                tmem_ld16(tacc + o_hh + c + dep, r0);
                tmem_ld16(tacc + o_ll + c + dep, r2);
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
              for (int c = 0; c < EPI_COLS; c += 16) {
                uint32_t r0[16];
                tmem_ld16(tacc + o_hh + c, r0);
                tmem_ld_wait();
                if (kb0 == 0) {
#pragma unroll
                  for (int q = 0; q < 16; ++q) racc[c + q] = r0[q];
                } else {
#pragma unroll
                  for (int q = 0; q < 16; ++q)
                    racc[c + q] = __float_as_uint(__uint_as_float(racc[c + q]) + __uint_as_float(r0[q]));
                }
              }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive_tempty(acc_stage);  // TMEM buffer free: the MMA warp moves on
            if (++acc_stage == T::ACC_STAGES) {
              acc_stage = 0;
              acc_phase ^= 1u;
            }
          }
        }
        auto release_direct = [&]() {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) arrive_tempty(acc_stage);
          if (++acc_stage == T::ACC_STAGES) {
            acc_stage = 0;
            acc_phase ^= 1u;
          }
        };
        if (dbg == 1 || !tile_valid) {  // (a pair's second block row that holds nothing: protocol only)
          if (direct) release_direct();
          continue;
        }
        // ---- wide tiles: straight from the lane = row layout (every lane finishes 64 consecutive entries of its own row)
        if (WIDE) {
          if (i < pp) {
            const EpiRow R = epi_row<EPI>(E, i);
#pragma unroll
            for (int q = 0; q < EPI_COLS; ++q) {
              const int j = col_base + q;
              double v = (double)__uint_as_float(racc[q]);
              if (scale != 1.0) v *= scale;
              if (j < pp) epi_store<EPI, true>(E, R, j, v, pass);
            }
          }
          continue;
        }
        // ---- transposed interior tile: lane = output column, register = output row; coalesced stores from registers
        if (swapped) {
          const int rb = tile.x * BN + ch * EPI_COLS;      // first output row of this warp (TMEM columns = row genes)
          const int cb = tile.y * BM + ew * 32 + lane;     // this lane's output column (TMEM lanes = column genes)
          const EpiRow R0 = epi_row<EPI>(E, rb);           // warp uniform
          const double colf = (KIND == KIND_I8) ? P.inv_norm[pass][cb] : 1.0;
          const double* rowv = (KIND == KIND_I8) ? P.inv_norm[pass] + rb : nullptr;
          char* const gp = reinterpret_cast<char*>(out + (R0.ro + cb));
          const int a = pp - rb - shift - 1;  // ro(i + 1) - ro(i) = p - i - shift - 1
          if (pw_mode == 6) packed_swapped<KIND, true>(racc, colf, scale, scale != 1.0, rowv, gp, a);
          else packed_swapped<KIND, false>(racc, colf, scale, scale != 1.0, rowv, gp, a);
          continue;
        }
        // ---- int32 transport of the exact Spearman Gram (host staging threads finish r and the power)
        if (S32) {
          int* const out32 = reinterpret_cast<int*>(out);
          if (interior) {
            const EpiRow R0 = epi_row<EPI>(E, row_base);  // warp uniform
            packed_interior_s32(racc, stg, lane, out32 + (R0.ro + col_base), pp - row_base - shift - 1);
          } else {  // diagonal / ragged tiles (3 % of them): straight from TMEM, lane = row, predicated scalar stores
            const bool row_ok = i < pp;
            const int64_t ro_i = row_ok ? epi_row<EPI>(E, i).ro : 0;
#pragma unroll
            for (int c = 0; c < EPI_COLS; c += 16) {
              uint32_t r0[16], r1[16], r2[16];
              tmem_ld16(tacc_direct + o_hl + c, r1);
              tmem_ld16(tacc_direct + o_hh + c, r0);
              tmem_ld16(tacc_direct + o_ll + c, r2);
              tmem_ld_wait();
#pragma unroll
              for (int q = 0; q < 16; ++q) {
                const int j = col_base + c + q;
                if (row_ok && j < pp && j >= i + shift) out32[ro_i + j] = 256 * (int)r0[q] + 16 * (int)r1[q] + (int)r2[q];
              }
            }
            release_direct();
          }
          continue;
        }
        // ---- lean path: row-side factors are applied with lane = row, column-side factors and the power
        //      after the transpose with lane = column; the 16 row offsets of a lane are 32-bit and computed
        //      once per tile, so a store costs one address add.
        if (!kFragEpilogue && interior) {
          const EpiRow R0 = epi_row<EPI>(E, row_base);  // warp uniform
          double rowf = scale;
          if (KIND == KIND_I8) rowf *= P.inv_norm[pass][i];
          const double* colv = (KIND == KIND_I8) ? P.inv_norm[pass] + col_base : nullptr;
          char* const wbase = reinterpret_cast<char*>(out + (R0.ro + col_base));
          const int a = pp - row_base - shift - 1;  // ro(i + 1) - ro(i) = p - i - shift - 1
          const long long pt2 = (kProfile && P.prof) ? clock64() : 0;
          if (BULK_COLS) {
            char* const grow = wbase + ((int64_t)(lane * a - (lane * (lane - 1)) / 2) << 3);
            if (pw_mode == 6) packed_interior_bulk<KIND, true>(racc, smem_u32(stg), lane, rowf, scale != 1.0, colv, grow, bulk_buf);
            else packed_interior_bulk<KIND, false>(racc, smem_u32(stg), lane, rowf, scale != 1.0, colv, grow, bulk_buf);
          } else if (pw_mode == 6) packed_interior<KIND, true>(racc, stg, lane, rowf, scale != 1.0, colv, wbase, a, dbg);
          else packed_interior<KIND, false>(racc, stg, lane, rowf, scale != 1.0, colv, wbase, a, dbg);
          if (kProfile && P.prof) {
            const long long pt3 = clock64();
            prof_acc[0] += pt1 - pt0;
            prof_acc[1] += pt2 - pt1;
            prof_acc[2] += pt3 - pt2;
            prof_acc[3] += 1;
          }
          continue;
        }
        if (BULK_COLS) bulk_wait_read_all();  // the staging below is reused with plain stores
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
          if (KIND == KIND_I8 && !OZAKI) inv_i = P.inv_norm[pass][i];
        }
        const int rr = lane >> 4, cc = lane & 15;
        const double* inv_col = (KIND == KIND_I8 && !OZAKI) ? P.inv_norm[pass] : nullptr;
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
            tmem_ld16(tacc_direct + o_hl + c, r1);
            if (kI8Wide) {
              tmem_ld16(tacc_direct + o_lh + c, r2);
              tmem_ld_wait();
#pragma unroll
              for (int q = 0; q < 16; ++q) r1[q] += r2[q];
            }
            tmem_ld16(tacc_direct + o_hh + c, r0);
            tmem_ld16(tacc_direct + o_ll + c, r2);
            tmem_ld_wait();
            if (OZAKI) {
#pragma unroll
              for (int q = 0; q < EPI_CHUNK; ++q)  // exact: every term is an integer below 2^53
                vals[q] = (16384.0 * (double)(int)r0[q] + 128.0 * (double)(int)r1[q] + (double)(int)r2[q]) * scale;
            } else {
#pragma unroll
              for (int q = 0; q < EPI_CHUNK; ++q)
                vals[q] = (double)(256 * (int)r0[q] + 16 * (int)r1[q] + (int)r2[q]) * inv_i;
            }
          } else {
#pragma unroll
            for (int q = 0; q < EPI_CHUNK; ++q)
              vals[q] = (KIND == KIND_I8) ? (double)(int)racc[c + q] * inv_i : (double)__uint_as_float(racc[c + q]);
          }
          if (KIND == KIND_I8 && !OZAKI) {
#pragma unroll
            for (int q = 0; q < EPI_CHUNK; ++q) vals[q] *= (jb + q < pp) ? __ldg(inv_col + jb + q) : 0.0;
          }
          if ((EPI == EPI_PACKED || EPI == EPI_DENSE || KIND == KIND_F16) && scale != 1.0) {
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
    if (BULK_COLS) bulk_wait_read_all();  // shared memory stays valid until the copy engine has read it
    if (kProfile && P.prof && lane == 0) {
      unsigned long long* dst = P.prof + ((size_t)blockIdx.x * EPI_WARPS + (warp - 4)) * 4;
      for (int q = 0; q < 4; ++q) dst[q] = (unsigned long long)prof_acc[q];
    }
  }

  tc_fence_before();
  if (CL > 1) cluster_sync_all();  // the peers' shared memory / TMEM stay alive until all CTAs of the cluster are done
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (CTAS == 1)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline char* err_buf() {
  static thread_local char buf[512] = "";  // one per host thread: the multi-device entry points launch from a thread per device
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
inline int make_map(CUtensorMap* map, int kind, const void* ptr, int64_t rows, int K, int box_rows = BM, int nplanes = 2) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    snprintf(err_buf(), 512, "cuTensorMapEncodeTiled entry point not available");
    return -1;
  }
  const int es = kind == KIND_TF32 ? 4 : (kind == KIND_F16 ? 2 : 1);
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)nplanes, (cuuint64_t)rows};
  cuuint64_t strides[2] = {(cuuint64_t)K * es, (cuuint64_t)K * es * nplanes};
  cuuint32_t box[3] = {(cuuint32_t)(ROW_BYTES / es), 1, (cuuint32_t)box_rows};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, kind == KIND_TF32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : (kind == KIND_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8), 3,
                  const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(err_buf(), 512, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld K=%d es=%d)", (int)r,
             (long long)rows, K, es);
    return -1;
  }
  return 0;
}

template <int KIND, int EPI, int CTAS, int MC = 0, int WIDE = 0, int S32 = 0>
inline int launch_one(cudaStream_t stream, int sm_count, const CUtensorMap& m0, const CUtensorMap& m1, const CUtensorMap& b0,
                      const CUtensorMap& b1, const Params& P) {
  auto kern = gram_umma_kernel<KIND, EPI, CTAS, MC, WIDE, S32>;
  constexpr int smem = WIDE ? wide_smem_bytes() : smem_bytes_of<CTAS>();
  constexpr int CL = MC ? 4 : CTAS;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) {
    snprintf(err_buf(), 512, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return -1;
  }
  int units = sm_count / CL;  // persistent: one CTA (pair / cluster) per SM (TPC / 4 SMs of a GPC)
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(WIDE ? WIDE_THREADS : THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (MC) {  // clusters of 4 must fit a GPC: ask how many can be resident at once (the long-K barrier needs all of them)
    cfg.gridDim = dim3((unsigned)(units * CL));
    int max_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) == cudaSuccess && max_clusters > 0 && max_clusters < units) units = max_clusters;
    else cudaGetLastError();
  }
  const int grid = CL * (P.ntiles < units ? P.ntiles : units);
  cfg.gridDim = dim3((unsigned)grid);
  e = cudaLaunchKernelEx(&cfg, kern, m0, m1, b0, b1, P);
  if (e != cudaSuccess) {
    snprintf(err_buf(), 512, "gram_umma_kernel launch (CTAS = %d, MC = %d): %s", CTAS, MC, cudaGetErrorString(e));
    return -1;
  }
  return 0;
}

// operand set s: gene-major digit planes ([row][plane][K]), K elements (padded), inv_norm (KIND_I8).
// ctas = 2: `tiles` is a pair list (first block row of the pair, block col | TILE_SECOND_INVALID).
inline int launch_gram_umma(cudaStream_t stream, int sm_count, int kind, int epi, int ctas, int64_t rows,
                            const void* planes0, int K0, const double* inv0, const void* planes1, int K1,
                            const double* inv1, const int2* tiles, int ntiles, const EpiParams& E, int nplanes = 2,
                            int plane_a0 = 0, int plane_b0 = 0, unsigned* tile_sync_counter = nullptr, int mc = 0,
                            int mc_rb0 = 0, int mc_rb1 = 0, int mc_upper = 0) {
  if (ntiles <= 0) return 0;
  const int kelems = kind == KIND_TF32 ? 32 : (kind == KIND_F16 ? 64 : 128);
  if (K0 % kelems != 0 || (planes1 && K1 % kelems != 0)) {
    snprintf(err_buf(), 512, "operand K must be padded to %d elements", kelems);
    return -1;
  }
  CUtensorMap m0, m1, b0, b1;
  if (make_map(&m0, kind, planes0, rows, K0, BM, nplanes)) return -1;
  const int wide = (mc == 2);  // mc: 0 plain, 1 = 2 x 2 multicast clusters, 2 = 256 x 256 tiles per CTA pair
  if (wide) mc = 0;
  if (mc && (ctas != 1 || planes1)) {
    snprintf(err_buf(), 512, "multicast clusters are made of single CTAs and take one operand set");
    return -1;
  }
  if (wide && (ctas != 2 || planes1)) {
    snprintf(err_buf(), 512, "wide tiles run on CTA pairs and take one operand set");
    return -1;
  }
  // B boxes: MC 64-row halves of both tiles; wide: this CTA's 128 of the 256 tile columns; pairs: 64; single: 128
  if (make_map(&b0, kind, planes0, rows, K0, mc ? BN / 2 : (wide ? BN : BN / ctas), nplanes)) return -1;
  if (planes1) {
    if (make_map(&m1, kind, planes1, rows, K1)) return -1;
    if (make_map(&b1, kind, planes1, rows, K1, BN / ctas)) return -1;
  } else {
    m1 = m0;
    b1 = b0;
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
  P.tile_sync = nullptr;
  P.tile_sync_kb = 0;
  if (const char* v = getenv("BIXCOR_UMMA_TILE_SYNC_KB")) P.tile_sync_kb = atoi(v) > 0 ? atoi(v) : 0;
  {
    // Long-K Grams (TOM, Ozaki passes, sparse panels): ncu shows the operand panels being re-read from DRAM (f16 TOM: 92 GB,
    // L2 hit rate 50 %) because the persistent CTAs drift apart along K.  A per-tile arrival barrier of the producers
    // keeps them in step; it costs a few microseconds per tile, so it is only armed when a tile's K loop is long.
    // tile_sync_counter: two words owned by the caller's per-device, per-stream context (so concurrent launches on
    // other devices / contexts never share it); [0] arrivals, [1] bail-outs of this launch.
    static const int sync_mode = [] {  // env BIXCOR_UMMA_TILE_SYNC: 0 = off, 1 = on for long K (default), 2 = always
      const char* v = getenv("BIXCOR_UMMA_TILE_SYNC");
      return v ? atoi(v) : BIXCOR_UMMA_TILE_SYNC_DEFAULT;
    }();
    if (tile_sync_counter && (sync_mode == 2 || (sync_mode == 1 && P.kblocks[0] >= 96))) {
      // armed only when the persistent grid fits the device at once (>= 1 CTA of this kernel per SM)
      const int cl = mc ? 4 : ctas;
      const int units = sm_count / cl;
      const bool grid_fits = cl * (ntiles < units ? ntiles : units) <= sm_count;
      if (grid_fits && cudaMemsetAsync(tile_sync_counter, 0, 2 * sizeof(unsigned), stream) == cudaSuccess) P.tile_sync = tile_sync_counter;
    }
  }
  {
    // env BIXCOR_UMMA_PREFETCH_KB: L2 prefetch distance of the long-K launches in K blocks (0 = off)
    static const int pf_env = [] {
      const char* v = getenv("BIXCOR_UMMA_PREFETCH_KB");
      return v ? atoi(v) : 0;  // measured slower (20k TOM f16 23.2 -> 24.5 ms, TF32 42.1 -> 54.4 ms): kept as an ablation switch
    }();
    P.prefetch_kb = (P.kblocks[0] >= 96 && pf_env > 0) ? pf_env : 0;
  }
  P.plane_a0 = plane_a0;
  P.plane_b0 = plane_b0;
  P.mc_nb = (int)((rows + BM - 1) / BM);
  P.mc_rb0 = mc_rb0;
  P.mc_rb1 = mc_rb1;
  P.mc_upper = mc_upper;
  P.prof = nullptr;
  {
    const char* d = getenv("BIXCOR_UMMA_DEBUG");
    P.dbg = d ? atoi(d) : 0;
  }
  static unsigned long long* d_prof = nullptr;
  const bool want_prof = getenv("BIXCOR_UMMA_PROF") != nullptr;
  if (want_prof) {  // developer aid: per-warp cycle counters of the lean epilogue, printed after the launch
    if (!d_prof) cudaMalloc(&d_prof, sizeof(unsigned long long) * 4 * EPI_WARPS * 1024);
    cudaMemsetAsync(d_prof, 0, sizeof(unsigned long long) * 4 * EPI_WARPS * 1024, stream);
    P.prof = d_prof;
  }
  struct ProfDump {
    cudaStream_t st;
    unsigned long long* d;
    int nblk;
    ~ProfDump() {
      if (!d) return;
      cudaStreamSynchronize(st);
      std::vector<unsigned long long> h((size_t)4 * EPI_WARPS * nblk);
      cudaMemcpy(h.data(), d, h.size() * 8, cudaMemcpyDeviceToHost);
      double w = 0, dr = 0, ep = 0, nt = 0;
      for (size_t i = 0; i < h.size(); i += 4) {
        w += (double)h[i];
        dr += (double)h[i + 1];
        ep += (double)h[i + 2];
        nt += (double)h[i + 3];
      }
      if (nt > 0)
        fprintf(stderr, "[umma prof] per warp-tile cycles: wait %.0f  drain %.0f  lean epilogue %.0f  (warp-tiles %.0f)\n", w / nt,
                dr / nt, ep / nt, nt);
    }
  } prof_dump{stream, want_prof ? d_prof : nullptr, sm_count};
  if (wide) {
    if (kind == KIND_F16 && epi == EPI_TOM) return launch_one<KIND_F16, EPI_TOM, 2, 0, 1>(stream, sm_count, m0, m1, b0, b1, P);
    if (kind == KIND_TF32 && epi == EPI_TOM) return launch_one<KIND_TF32, EPI_TOM, 2, 0, 1>(stream, sm_count, m0, m1, b0, b1, P);
    snprintf(err_buf(), 512, "no wide-tile tcgen05 kernel for kind %d / epilogue %d", kind, epi);
    return -1;
  }
#define BIXCOR_UMMA_MC_CASE(KIND, EPI) \
  if (mc && kind == KIND && epi == EPI) return launch_one<KIND, EPI, 1, 1>(stream, sm_count, m0, m1, b0, b1, P);
  BIXCOR_UMMA_MC_CASE(KIND_F16, EPI_TOM)
  BIXCOR_UMMA_MC_CASE(KIND_TF32, EPI_TOM)
  BIXCOR_UMMA_MC_CASE(KIND_F16, EPI_ACCUM)
  BIXCOR_UMMA_MC_CASE(KIND_TF32, EPI_ACCUM)
  BIXCOR_UMMA_MC_CASE(KIND_I8, EPI_ACCUM)
#undef BIXCOR_UMMA_MC_CASE
  if (mc) {
    snprintf(err_buf(), 512, "no multicast tcgen05 kernel for kind %d / epilogue %d", kind, epi);
    return -1;
  }
  if (E.out_s32) {  // int32 transport of the exact Spearman Gram
    if (kind != KIND_I8 || epi != EPI_PACKED || planes1) {
      snprintf(err_buf(), 512, "the int32 transport exists for the exact int8 packed correlation only");
      return -1;
    }
    return ctas == 2 ? launch_one<KIND_I8, EPI_PACKED, 2, 0, 0, 1>(stream, sm_count, m0, m1, b0, b1, P)
                     : launch_one<KIND_I8, EPI_PACKED, 1, 0, 0, 1>(stream, sm_count, m0, m1, b0, b1, P);
  }
#define BIXCOR_UMMA_CASE(KIND, EPI)                                                          \
  if (kind == KIND && epi == EPI)                                                            \
    return ctas == 2 ? launch_one<KIND, EPI, 2>(stream, sm_count, m0, m1, b0, b1, P)          \
                     : launch_one<KIND, EPI, 1>(stream, sm_count, m0, m1, b0, b1, P);
  BIXCOR_UMMA_CASE(KIND_TF32, EPI_PACKED)
  BIXCOR_UMMA_CASE(KIND_TF32, EPI_DENSE)
  BIXCOR_UMMA_CASE(KIND_TF32, EPI_DIFFCOR)
  BIXCOR_UMMA_CASE(KIND_TF32, EPI_TOM)
  BIXCOR_UMMA_CASE(KIND_TF32, EPI_ACCUM)
  BIXCOR_UMMA_CASE(KIND_F16, EPI_PACKED)
  BIXCOR_UMMA_CASE(KIND_F16, EPI_DENSE)
  BIXCOR_UMMA_CASE(KIND_F16, EPI_TOM)
  BIXCOR_UMMA_CASE(KIND_F16, EPI_ACCUM)
  BIXCOR_UMMA_CASE(KIND_I8, EPI_PACKED)
  BIXCOR_UMMA_CASE(KIND_I8, EPI_DENSE)
  BIXCOR_UMMA_CASE(KIND_I8, EPI_DIFFCOR)
  BIXCOR_UMMA_CASE(KIND_I8, EPI_ACCUM)
#undef BIXCOR_UMMA_CASE
  snprintf(err_buf(), 512, "no tcgen05 kernel for kind %d / epilogue %d", kind, epi);
  return -1;
}

}  // namespace umma
}  // namespace bixcor
