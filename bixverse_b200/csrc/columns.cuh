// Per-gene column kernels: Spearman average-tie rank transform and Pearson
// standardisation.  Both write the Gram operand directly in the layout the
// GEMM kernels consume (K-major: the samples of one gene are contiguous and
// zero padded to `kpad`), so the standardised matrix never makes an extra
// round trip through HBM.
//
// Reference semantics restated (see oracle/oracle.py): crate
// core::math::matrix_helpers::rank_matrix_col (exposed at
// src/rust/src/enrichment/r_singscore.rs:89-95) == base R
// rank(ties.method = "average"); column_pairwise_cor
// (src/rust/src/base/r_cors_similarity.rs:70-77) == base R cor().
#pragma once

#include "common.cuh"
#include "epilogue.cuh"

namespace bixcor {

enum ColumnMode : int {
  COL_RANKS = 0,  // f64 average ranks (1-based), ld = n
  COL_Z64 = 1,    // f64 unit-norm centred column, zero padded to kpad
  COL_I8 = 2,     // centred integer ranks d = 2*rank-(n+1) as base-16 digits (h, l) + 1/|d|
  COL_TF32 = 3,   // unit-norm centred column split into TF32 hi / lo planes
};

struct ColumnOut {
  int mode;
  int kpad;            // padded K (>= n); elements [n, kpad) are zero filled
  double* f64;         // COL_RANKS / COL_Z64
  int64_t ld64;
  int8_t* i8;          // COL_I8: 2 planes (h, l), plane stride `plane8` bytes
  int64_t ld8;
  int64_t plane8;
  float* f32;          // COL_TF32: 2 planes, plane stride `plane32` floats
  int64_t ld32;
  int64_t plane32;
  int half_planes;     // COL_TF32 with binary16 planes instead (BIXCOR_PREC_F16X3): f32 / ld32 / plane32 then address
                       // __half elements and the values are scaled by kF16Scale before the split

  double* inv_norm;    // COL_I8: 1/sqrt(sum d^2) per gene
};

// Fused operand exchange (one process per GPU, bixcor_xchg_* in csrc/bixcor.cu): the column kernel stores every COL_I8 record
// it finishes not only into its own operand buffer but straight into the operand buffer of every peer GPU, through
// P2P-mapped (cudaIpc) pointers over NVLink / NVSwitch, so that the separate all-gather collective disappears and the
// transfer overlaps the sort.  Pointers are laid out like ColumnOut::i8 / inv_norm (already offset to the launch's first gene).
constexpr int kMaxPeers = 8;
struct PeerOut {
  int n;                       // peer destinations (the local buffer is ColumnOut's and not listed here)
  int8_t* i8[kMaxPeers];
  double* inv_norm[kMaxPeers];
};

__device__ __forceinline__ uint64_t sortable_key(double v) {
  v = v + 0.0;  // -0.0 -> +0.0 so that signed zeros tie (base-R semantics)
  uint64_t b = (uint64_t)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

// exclusive prefix-max over the threads of the block (identity -1)
__device__ __forceinline__ int block_excl_prefix_max(int v, int* scr) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc = max(inc, t);
  }
  __syncthreads();
  if (lane == 31) scr[warp] = inc;
  __syncthreads();
  int wp = -1;
  for (int w = 0; w < warp; ++w) wp = max(wp, scr[w]);
  int ex = __shfl_up_sync(0xffffffffu, inc, 1);
  if (lane == 0) ex = -1;
  return max(wp, ex);
}

// exclusive suffix-min over the threads of the block (identity INT_MAX)
__device__ __forceinline__ int block_excl_suffix_min(int v, int* scr) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_down_sync(0xffffffffu, inc, o);
    if (lane + o < 32) inc = min(inc, t);
  }
  __syncthreads();
  if (lane == 0) scr[warp] = inc;
  __syncthreads();
  int ws = 0x7fffffff;
  for (int w = warp + 1; w < nwarp; ++w) ws = min(ws, scr[w]);
  int ex = __shfl_down_sync(0xffffffffu, inc, 1);
  if (lane == 31) ex = 0x7fffffff;
  return min(ws, ex);
}

// Write one standardised column in the requested operand layout.
// `val(i)` yields the centred value of sample i (double), `scale_div` the norm.
__device__ __forceinline__ void emit_tf32(const ColumnOut& out, int64_t j, int i, double z) {
  if (out.half_planes) {
    __half hi, lo;
    split_f16(z, hi, lo);
    __half* hp = reinterpret_cast<__half*>(out.f32);
    hp[j * out.ld32 + i] = hi;
    hp[out.plane32 + j * out.ld32 + i] = lo;
    return;
  }
  float hi = to_tf32((float)z);
  float lo = to_tf32((float)(z - (double)hi));
  out.f32[j * out.ld32 + i] = hi;
  out.f32[out.plane32 + j * out.ld32 + i] = lo;
}

// A column holding NaN: base-R cor() gives NA against everything.  One CTA writes the operand record of gene j.
__device__ __forceinline__ void emit_nan_column_block(const ColumnOut& out, int64_t j, int n) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const double qnan = __longlong_as_double(0x7ff8000000000000ll);
  if (out.mode == COL_RANKS) {
    for (int i = tid; i < n; i += nt) out.f64[j * out.ld64 + i] = qnan;
  } else if (out.mode == COL_Z64) {
    for (int i = tid; i < out.kpad; i += nt) out.f64[j * out.ld64 + i] = (i < n) ? qnan : 0.0;
  } else if (out.mode == COL_I8) {
    for (int i = tid; i < out.kpad; i += nt) {
      out.i8[j * out.ld8 + i] = 0;
      out.i8[out.plane8 + j * out.ld8 + i] = 0;
    }
    if (tid == 0) out.inv_norm[j] = qnan;
  } else {
    for (int i = tid; i < out.kpad; i += nt) {
      if (out.half_planes) {
        emit_tf32(out, j, i, (i < n) ? qnan : 0.0);
        continue;
      }
      float f = (i < n) ? __int_as_float(0x7fc00000) : 0.f;
      out.f32[j * out.ld32 + i] = f;
      out.f32[out.plane32 + j * out.ld32 + i] = 0.f;
    }
  }
}

// Operand record of gene j from rk2[i] = 2 * (average 1-based rank of sample i) (shared or global memory), one CTA.
// The centred integer rank is d = 2*rank - (n+1); the rank mean is (n+1)/2 with or without ties.
__device__ __forceinline__ void emit_ranked_column_block(const ColumnOut& out, int64_t j, int n, const uint32_t* rk2,
                                                         long long* scr_ll) {
  const int tid = threadIdx.x, nt = blockDim.x;
  if (out.mode == COL_RANKS) {
    for (int i = tid; i < n; i += nt) out.f64[j * out.ld64 + i] = 0.5 * (double)rk2[i];
    return;
  }
  long long ss = 0;
  for (int i = tid; i < n; i += nt) {
    const long long d = (long long)rk2[i] - (long long)(n + 1);
    ss += d * d;
  }
  ss = block_sum_ll(ss, scr_ll);
  const double norm = sqrt((double)ss);
  if (out.mode == COL_Z64) {
    for (int i = tid; i < out.kpad; i += nt) {
      double z = 0.0;
      if (i < n) z = (double)((long long)rk2[i] - (long long)(n + 1)) / norm;
      out.f64[j * out.ld64 + i] = z;
    }
  } else if (out.mode == COL_I8) {
    for (int i = tid; i < out.kpad; i += nt) {
      int d = (i < n) ? (int)rk2[i] - (n + 1) : 0;
      const int l = ((d + 8) & 15) - 8;  // low digit in [-8, 7]
      const int h = (d - l) >> 4;        // |h| <= 127 for n <= 2025
      out.i8[j * out.ld8 + i] = (int8_t)h;
      out.i8[out.plane8 + j * out.ld8 + i] = (int8_t)l;
    }
    if (tid == 0) out.inv_norm[j] = 1.0 / norm;
  } else {
    for (int i = tid; i < out.kpad; i += nt) {
      double z = 0.0;
      if (i < n) z = (double)((long long)rk2[i] - (long long)(n + 1)) / norm;
      emit_tf32(out, j, i, z);
    }
  }
}

// Average-tie ranks of a mostly-zero column held dense (sparse single-cell genes, any number of cells): one CTA per
// gene compacts the non-zero entries into shared memory (at most `cap`, a power of two; dynamic smem cap * 12
// bytes), sorts them, and ranks the zeros analytically: with `neg` negative and `zc` zero entries the zeros share
// rank neg + (zc + 1) / 2, the s-th smallest non-zero sits at full position s (negative) or s + zc (positive).
// Ranks overwrite the column in place.  A column with more than `cap` non-zeros sets overflow[gene] and is left
// untouched: the host ranks those columns with the chunked dense kernels below (zeros tie like any other value).
constexpr int kSparseRankThreads = 512;
__global__ void __launch_bounds__(kSparseRankThreads)
rank_sparse_columns_kernel(double* __restrict__ x, int64_t ldx, int64_t n, int cap, int* __restrict__ overflow) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* key = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t* idx = reinterpret_cast<uint32_t*>(key + cap);
  __shared__ int s_cnt, s_neg, s_nan;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
  double* col = x + (int64_t)blockIdx.x * ldx;
  if (tid == 0) s_cnt = s_neg = s_nan = 0;
  __syncthreads();
  // ---- pass 1: compact the non-zeros (order is irrelevant, they are sorted next)
  for (int64_t i0 = 0; i0 < n; i0 += nt) {
    const int64_t i = i0 + tid;
    const double v = (i < n) ? col[i] : 0.0;
    const bool nz = (v != 0.0);  // NaN counts as non-zero here and poisons the column below
    if (v != v) s_nan = 1;
    const unsigned m = __ballot_sync(0xffffffffu, nz);
    const unsigned mneg = __ballot_sync(0xffffffffu, v < 0.0);
    int base = 0;
    if (lane == 0 && m) {
      base = atomicAdd(&s_cnt, __popc(m));
      if (mneg) atomicAdd(&s_neg, __popc(mneg));
    }
    base = __shfl_sync(0xffffffffu, base, 0);
    if (nz) {
      const int pos = base + __popc(m & ((1u << lane) - 1u));
      if (pos < cap) {
        key[pos] = sortable_key(v);
        idx[pos] = (uint32_t)i;
      }
    }
  }
  __syncthreads();
  const int cnt = s_cnt, neg = s_neg;
  if (s_nan) {  // base-R cor(): a column holding NA correlates to NA with everything
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);
    for (int64_t i = tid; i < n; i += nt) col[i] = qnan;
    return;
  }
  if (cnt > cap) {
    if (tid == 0) overflow[blockIdx.x] = 1;
    return;
  }
  int npow2 = 2;
  while (npow2 < cnt) npow2 <<= 1;
  for (int i = cnt + tid; i < npow2; i += nt) {
    key[i] = ~0ull;
    idx[i] = 0xffffffffu;
  }
  __syncthreads();
  const int half = npow2 >> 1;
  for (int k = 2; k <= npow2; k <<= 1) {
    for (int jj = k >> 1; jj > 0; jj >>= 1) {
      for (int t = tid; t < half; t += nt) {
        const int i = ((t & ~(jj - 1)) << 1) | (t & (jj - 1));
        const int l = i | jj;
        const bool up = (i & k) == 0;
        const uint64_t ki = key[i], kl = key[l];
        if ((ki > kl) == up && ki != kl) {
          key[i] = kl;
          key[l] = ki;
          const uint32_t ti = idx[i];
          idx[i] = idx[l];
          idx[l] = ti;
        }
      }
      __syncthreads();
    }
  }
  // ---- zeros first (the sorted list lives in shared memory, the column can be overwritten)
  const int64_t zc = n - cnt;
  const double zero_rank = (double)neg + 0.5 * (double)(zc + 1);
  for (int64_t i = tid; i < n; i += nt)
    if (col[i] == 0.0) col[i] = zero_rank;
  __syncthreads();
  // ---- non-zeros: tie run [lo, hi] of equal keys by binary search, rank = mean of the full positions + 1
  for (int s = tid; s < cnt; s += nt) {
    const uint64_t k = key[s];
    int a = 0, b = s;  // lower bound
    while (a < b) {
      const int mid = (a + b) >> 1;
      if (key[mid] < k) a = mid + 1; else b = mid;
    }
    const int lo = a;
    a = s;
    b = cnt - 1;  // upper bound (last index with key == k)
    while (a < b) {
      const int mid = (a + b + 1) >> 1;
      if (key[mid] > k) b = mid - 1; else a = mid;
    }
    const int hi = a;
    const double shift = (s < neg) ? 0.0 : (double)zc;
    col[idx[s]] = 0.5 * (double)(lo + hi) + 1.0 + shift;
  }
}

// One CTA per gene.  Dynamic smem: npow2 * 12 bytes (u64 key + u32 index).
// blockDim.x must be a power of two, a multiple of 32 and <= npow2.
__global__ void __launch_bounds__(256)
rank_columns_kernel(const double* __restrict__ x, int64_t ldx, int n, int npow2, ColumnOut out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* key = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t* idx = reinterpret_cast<uint32_t*>(key + npow2);
  __shared__ long long scr_ll[32];
  __shared__ int scr_i[32];
  __shared__ int s_nan;

  const int tid = threadIdx.x, nt = blockDim.x;
  const int64_t j = blockIdx.x;
  const double* col = x + j * ldx;

  if (tid == 0) s_nan = 0;
  __syncthreads();
  bool has_nan = false;
  for (int i = tid; i < npow2; i += nt) {
    uint64_t k = ~0ull;
    if (i < n) {
      double v = col[i];
      if (v != v) has_nan = true;
      k = sortable_key(v);
    }
    key[i] = k;
    idx[i] = (uint32_t)i;
  }
  if (has_nan) s_nan = 1;
  __syncthreads();

  if (s_nan) {  // base-R cor(): a column holding NA correlates to NA with everything
    emit_nan_column_block(out, j, n);
    return;
  }

  // ---- bitonic sort of (key, idx) in shared memory -------------------------
  const int half = npow2 >> 1;
  for (int k = 2; k <= npow2; k <<= 1) {
    for (int jj = k >> 1; jj > 0; jj >>= 1) {
      for (int t = tid; t < half; t += nt) {
        const int i = ((t & ~(jj - 1)) << 1) | (t & (jj - 1));
        const int l = i | jj;
        const bool up = (i & k) == 0;
        const uint64_t ki = key[i], kl = key[l];
        if ((ki > kl) == up && ki != kl) {
          key[i] = kl;
          key[l] = ki;
          const uint32_t ti = idx[i];
          idx[i] = idx[l];
          idx[l] = ti;
        }
      }
      __syncthreads();
    }
  }

  // ---- tie groups: flag the first position of every run of equal keys -------
  for (int s = tid; s < npow2; s += nt) {
    const bool start = (s >= n) || (s == 0) || (key[s] != key[s - 1]);
    if (start) idx[s] |= 0x80000000u;
  }
  __syncthreads();  // keys are dead from here on; their storage is reused

  uint32_t* val = reinterpret_cast<uint32_t*>(key);          // lo+hi+2 per sorted position
  uint32_t* rk2 = reinterpret_cast<uint32_t*>(key) + npow2;  // 2*rank per original position
  const int c = npow2 / nt;
  const int s0 = tid * c, s1 = s0 + c;

  int last = -1;
  for (int s = s0; s < s1; ++s)
    if (idx[s] & 0x80000000u) last = s;
  int cur = block_excl_prefix_max(last, scr_i);
  for (int s = s0; s < s1; ++s) {
    if (idx[s] & 0x80000000u) cur = s;
    val[s] = (uint32_t)cur;  // lo: first position of the run
  }
  int first = 0x7fffffff;  // smallest run-end position inside this thread's chunk
  for (int s = s0; s < s1; ++s) {
    const bool end = (s + 1 >= npow2) || (idx[s + 1] & 0x80000000u);
    if (end) { first = s; break; }
  }
  cur = block_excl_suffix_min(first, scr_i);
  for (int s = s1 - 1; s >= s0; --s) {
    const bool end = (s + 1 >= npow2) || (idx[s + 1] & 0x80000000u);
    if (end) cur = s;
    val[s] += (uint32_t)cur + 2u;  // lo + hi + 2 == 2 * average 1-based rank
  }
  __syncthreads();
  for (int s = tid; s < n; s += nt) rk2[idx[s] & 0x7fffffffu] = val[s];
  __syncthreads();

  emit_ranked_column_block(out, j, n, rk2, scr_ll);
}

// ----------------------------------------------------------------------------
// Spearman for n > 16384 samples (no upper limit in the reference: rank_matrix_col sorts any column).  A column no
// longer fits one CTA's shared memory, so it is cut into chunks of kBigChunk samples:
//   rank_big_sort_kernel   CTA (chunk, gene): bitonic sort of the chunk in shared memory -> sorted keys + sample indices
//   rank_big_count_kernel  CTA (chunk, gene): every element counts, over ALL chunks of its gene, the keys below it (lo) and
//                          the keys not above it (hi) by binary search: its tie run occupies sorted positions [lo, hi - 1],
//                          so 2 * average rank = lo + hi + 1.  No merge passes; the sorted chunks of a gene (<= a few MB) stay in L2.
//   rank_big_emit_kernel   CTA per gene: operand record from the 2 * rank values
// ----------------------------------------------------------------------------
constexpr int kBigChunk = 16384;
constexpr int kBigThreads = 1024;

__global__ void __launch_bounds__(kBigThreads)
rank_big_sort_kernel(const double* __restrict__ x, int64_t ldx, int n, int nchunks, uint64_t* __restrict__ keys,
                     uint32_t* __restrict__ idx_out, int* __restrict__ nan_flag) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* key = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t* idx = reinterpret_cast<uint32_t*>(key + kBigChunk);
  const int tid = threadIdx.x, nt = blockDim.x;
  const int c = blockIdx.x;
  const int64_t g = blockIdx.y;
  const double* col = x + g * ldx;
  const int base = c * kBigChunk;
  bool has_nan = false;
  for (int i = tid; i < kBigChunk; i += nt) {
    uint64_t k = ~0ull;
    if (base + i < n) {
      const double v = col[base + i];
      if (v != v) has_nan = true;
      k = sortable_key(v);
    }
    key[i] = k;
    idx[i] = (uint32_t)(base + i);
  }
  if (has_nan) nan_flag[g] = 1;
  __syncthreads();
  constexpr int half = kBigChunk >> 1;
  for (int k = 2; k <= kBigChunk; k <<= 1) {
    for (int jj = k >> 1; jj > 0; jj >>= 1) {
      for (int t = tid; t < half; t += nt) {
        const int i = ((t & ~(jj - 1)) << 1) | (t & (jj - 1));
        const int l = i | jj;
        const bool up = (i & k) == 0;
        const uint64_t ki = key[i], kl = key[l];
        if ((ki > kl) == up && ki != kl) {
          key[i] = kl;
          key[l] = ki;
          const uint32_t ti = idx[i];
          idx[i] = idx[l];
          idx[l] = ti;
        }
      }
      __syncthreads();
    }
  }
  uint64_t* ko = keys + ((size_t)g * nchunks + c) * kBigChunk;
  uint32_t* io = idx_out + ((size_t)g * nchunks + c) * kBigChunk;
  for (int i = tid; i < kBigChunk; i += nt) {
    ko[i] = key[i];
    io[i] = idx[i];
  }
}

__global__ void __launch_bounds__(256)
rank_big_count_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ idx, int n, int nchunks,
                      uint32_t* __restrict__ rk2) {
  const int c = blockIdx.x;
  const int64_t g = blockIdx.y;
  const uint64_t* gk = keys + (size_t)g * nchunks * kBigChunk;
  const int valid = min(kBigChunk, n - c * kBigChunk);
  for (int s = threadIdx.x; s < valid; s += blockDim.x) {
    const uint64_t k = gk[(size_t)c * kBigChunk + s];
    uint32_t lo = 0, hi = 0;
    for (int cc = 0; cc < nchunks; ++cc) {
      const uint64_t* b = gk + (size_t)cc * kBigChunk;
      const int v = min(kBigChunk, n - cc * kBigChunk);
      int a0 = 0, a1 = v;  // first position with key >= k
      while (a0 < a1) {
        const int mid = (a0 + a1) >> 1;
        if (__ldg(b + mid) < k) a0 = mid + 1; else a1 = mid;
      }
      lo += (uint32_t)a0;
      a1 = v;  // first position with key > k (continue from the lower bound)
      while (a0 < a1) {
        const int mid = (a0 + a1) >> 1;
        if (__ldg(b + mid) <= k) a0 = mid + 1; else a1 = mid;
      }
      hi += (uint32_t)a0;
    }
    rk2[(size_t)g * n + idx[((size_t)g * nchunks + c) * kBigChunk + s]] = lo + hi + 1u;  // lo + (hi - 1) + 2
  }
}

__global__ void __launch_bounds__(256)
rank_big_emit_kernel(const uint32_t* __restrict__ rk2, const int* __restrict__ nan_flag, int n, ColumnOut out) {
  __shared__ long long scr_ll[32];
  const int64_t j = blockIdx.x;
  if (nan_flag[j]) {
    emit_nan_column_block(out, j, n);
    return;
  }
  emit_ranked_column_block(out, j, n, rk2 + (size_t)j * n, scr_ll);
}

// ----------------------------------------------------------------------------
// Warp-per-gene rank kernel for n <= 1024: the whole column lives in registers
// (E = npow2/32 elements per lane, sorted position s = lane*E + r).  Bitonic
// stages whose partner distance is below E are compare-exchanges between
// registers of one thread; the others are lane exchanges by shuffle.  No
// shared-memory traffic and no block barrier in the sort.
// ----------------------------------------------------------------------------
template <typename RK, bool PEERS = false>
__device__ __forceinline__ void emit_ranked_column_warp(const ColumnOut& out, int64_t j, int n, const RK* rk2, int lane,
                                                        const PeerOut* po = nullptr) {
  if (out.mode == COL_RANKS) {
    for (int i = lane; i < n; i += 32) out.f64[j * out.ld64 + i] = 0.5 * (double)rk2[i];
    return;
  }
  long long ss = 0;
  for (int i = lane; i < n; i += 32) {
    const long long d = (long long)rk2[i] - (long long)(n + 1);
    ss += d * d;
  }
  ss = warp_sum_ll(ss);
  const double norm = sqrt((double)ss);
  if (out.mode == COL_Z64) {
    for (int i = lane; i < out.kpad; i += 32) {
      double z = 0.0;
      if (i < n) z = (double)((int)rk2[i] - (n + 1)) / norm;
      out.f64[j * out.ld64 + i] = z;
    }
  } else if (out.mode == COL_I8) {
    // four samples per lane and store: one 32-bit word of each digit plane
    for (int i4 = lane * 4; i4 < out.kpad; i4 += 128) {
      uint32_t wh = 0, wl = 0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int i = i4 + q;
        const int d = (i < n) ? (int)rk2[i] - (n + 1) : 0;
        const int l = ((d + 8) & 15) - 8;
        const int h = (d - l) >> 4;
        wh |= (uint32_t)(h & 0xff) << (8 * q);
        wl |= (uint32_t)(l & 0xff) << (8 * q);
      }
      *reinterpret_cast<uint32_t*>(out.i8 + j * out.ld8 + i4) = wh;
      *reinterpret_cast<uint32_t*>(out.i8 + out.plane8 + j * out.ld8 + i4) = wl;
      if (PEERS) {  // the same two words into every peer's operand buffer (128 contiguous bytes per warp store, over NVLink)
        for (int r = 0; r < po->n; ++r) {
          *reinterpret_cast<uint32_t*>(po->i8[r] + j * out.ld8 + i4) = wh;
          *reinterpret_cast<uint32_t*>(po->i8[r] + out.plane8 + j * out.ld8 + i4) = wl;
        }
      }
    }
    if (lane == 0) {
      out.inv_norm[j] = 1.0 / norm;
      if (PEERS)
        for (int r = 0; r < po->n; ++r) po->inv_norm[r][j] = 1.0 / norm;
    }
  } else {
    for (int i = lane; i < out.kpad; i += 32) {
      double z = 0.0;
      if (i < n) z = (double)((int)rk2[i] - (n + 1)) / norm;
      emit_tf32(out, j, i, z);
    }
  }
}

// Compare-exchange primitives of the register sort.  Keys are the raw doubles (no NaN reaches the
// sort; -0.0 == +0.0 under DSETP, which is the base-R tie semantics), so the comparison runs on the
// FP64 pipe and only the selects occupy the integer ALU pipe, which is what bounds this kernel.
// Inline setp/selp: ptxas otherwise turns the conditional swaps into branches.
__device__ __forceinline__ void ce_asc(double& ka, double& kb, uint32_t& ia, uint32_t& ib) {
  // afterwards ka <= kb; equal keys never swap
  const double a = ka, b = kb;
  const uint32_t x = ia, y = ib;
  asm("{\n\t.reg .pred p;\n\t"
      "setp.lt.f64 p, %5, %4;\n\t"
      "selp.f64 %0, %5, %4, p;\n\t"
      "selp.f64 %1, %4, %5, p;\n\t"
      "selp.b32 %2, %7, %6, p;\n\t"
      "selp.b32 %3, %6, %7, p;\n\t}"
      : "=&d"(ka), "=&d"(kb), "=&r"(ia), "=&r"(ib)
      : "d"(a), "d"(b), "r"(x), "r"(y));
}
// Lane exchange: this lane keeps the smaller (keep_min != 0) or the larger of (own, partner's);
// on equal keys both lanes keep their own element, so no index is lost or duplicated.
__device__ __forceinline__ void ce_lane(double& k, uint32_t& i, double ok, uint32_t oi, int keep_min) {
  // take = keep_min ? (partner < own) : (partner > own) = lt XOR (ne AND !keep_min): two DSETP, no predicate logic op
  asm("{\n\t.reg .pred m, q, t;\n\t"
      "setp.eq.s32 m, %4, 0;\n\t"
      "setp.neu.and.f64 q, %2, %0, m;\n\t"
      "setp.lt.xor.f64 t, %2, %0, q;\n\t"
      "selp.f64 %0, %2, %0, t;\n\t"
      "selp.b32 %1, %3, %1, t;\n\t}"
      : "+d"(k), "+r"(i)
      : "d"(ok), "r"(oi), "r"(keep_min));
}

// ---- the sort network, generic over the element type ------------------------------------------------------------------
// SortKI: (full key, sample index) - the exact form.  SortPK: ONE double whose low 10 bits hold the sample index (see
// rank_columns_warp_body): 2 instead of 3 registers per element, a compare-exchange moves 2 instead of 3 words, a lane
// exchange shuffles 2 instead of 3, and - all elements being distinct - needs one comparison instead of two.
struct SortKI {
  double k;
  uint32_t i;
};
struct SortPK {
  double k;
};
__device__ __forceinline__ void ce_asc(SortKI& a, SortKI& b) { ce_asc(a.k, b.k, a.i, b.i); }
__device__ __forceinline__ void ce_asc(SortPK& a, SortPK& b) {  // afterwards a.k <= b.k
  const double x = a.k, y = b.k;
  asm("{\n\t.reg .pred p;\n\t"
      "setp.lt.f64 p, %3, %2;\n\t"
      "selp.f64 %0, %3, %2, p;\n\t"
      "selp.f64 %1, %2, %3, p;\n\t}"
      : "=&d"(a.k), "=&d"(b.k)
      : "d"(x), "d"(y));
}
__device__ __forceinline__ void ce_lane(SortKI& own, int lm, int keep_min, const SortKI& give) {
  const double ok = __shfl_xor_sync(0xffffffffu, give.k, lm);
  const uint32_t oi = __shfl_xor_sync(0xffffffffu, give.i, lm);
  ce_lane(own.k, own.i, ok, oi, keep_min);
}
__device__ __forceinline__ void ce_lane(SortPK& own, int lm, int keep_min, const SortPK& give) {
  const double ok = __shfl_xor_sync(0xffffffffu, give.k, lm);
  // distinct keys: take = keep_min ? (partner < own) : (partner > own) = (partner < own) XOR (keep_min == 0)
  asm("{\n\t.reg .pred m, t;\n\t"
      "setp.eq.s32 m, %2, 0;\n\t"
      "setp.lt.xor.f64 t, %1, %0, m;\n\t"
      "selp.f64 %0, %1, %0, t;\n\t}"
      : "+d"(own.k)
      : "d"(ok), "r"(keep_min));
}

// Bitonic sort of the warp's 32 * E elements, position s = lane*E + r, ascending everywhere.
// Each merge of size k opens with the MIRROR stage (s <-> s ^ (k-1)), after which every
// half-cleaner stage (s <-> s ^ jj) sorts upwards: no direction flags, no key flipping.
// Merges inside one lane (k <= E) are fully unrolled register networks; merges across lanes
// run in a RUNTIME loop over k (the fully unrolled network is instruction-fetch bound).
template <int E, class T>
__device__ __forceinline__ void bitonic_sort_warp(T (&a)[E], int lane) {
  constexpr int NP = 32 * E;
#pragma unroll
  for (int k = 2; k <= E; k <<= 1) {
#pragma unroll
    for (int r = 0; r < E; ++r) {
      const int l = r ^ (k - 1);
      if (r < l) ce_asc(a[r], a[l]);
    }
#pragma unroll
    for (int jj = k >> 2; jj > 0; jj >>= 1) {
#pragma unroll
      for (int r = 0; r < E; ++r)
        if ((r & jj) == 0) ce_asc(a[r], a[r | jj]);
    }
  }
#pragma unroll 1
  for (int k = 2 * E; k <= NP; k <<= 1) {
    {  // mirror stage: partner lane = lane ^ (k/E - 1), partner register = E-1-r
      const int lm = k / E - 1;
      const int keep_min = ((lane & (k / (2 * E))) == 0) ? 1 : 0;
      if (E == 1) ce_lane(a[0], lm, keep_min, a[0]);
#pragma unroll
      for (int r = 0; r < E / 2; ++r) {  // registers r and E-1-r trade places across the lane pair
        const int m = E - 1 - r;
        const T gm = a[m], gr = a[r];
        ce_lane(a[r], lm, keep_min, gm);
        ce_lane(a[m], lm, keep_min, gr);
      }
    }
#pragma unroll 1
    for (int jj = k >> 2; jj >= E; jj >>= 1) {
      const int lm = jj / E;  // partner lane = lane ^ lm, same register
      const int keep_min = ((lane & lm) == 0) ? 1 : 0;
#pragma unroll
      for (int r = 0; r < E; ++r) ce_lane(a[r], lm, keep_min, a[r]);
    }
#pragma unroll
    for (int jj = E / 2; jj > 0; jj >>= 1) {
#pragma unroll
      for (int r = 0; r < E; ++r)
        if ((r & jj) == 0) ce_asc(a[r], a[r | jj]);
    }
  }
}

// -DBIXCOR_RANK_PACKED=0 builds the kernel without the packed fast path (ablation: every column takes the exact sort)
#ifndef BIXCOR_RANK_PACKED
#define BIXCOR_RANK_PACKED 1
#endif

// Tie runs -> 2 * average rank of every element, written to rk2[sample index].  start_mask bit r: position lane*E + r
// starts a run of equal values; last: the lane's last run start (-1 if none); idx[r]: sample index at position lane*E + r.
template <int E>
__device__ __forceinline__ void assign_ranks_warp(uint32_t start_mask, int last, const uint32_t (&idx)[E], int n, int lane, uint16_t* rk2) {
  // exclusive prefix max of `last` over lanes
  int inc = last;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc = max(inc, t);
  }
  int cur = __shfl_up_sync(0xffffffffu, inc, 1);
  if (lane == 0) cur = -1;
  uint32_t lo[E];
#pragma unroll
  for (int r = 0; r < E; ++r) {
    if (start_mask & (1u << r)) cur = lane * E + r;
    lo[r] = (uint32_t)cur;
  }
  // run ends: position s ends a run iff s+1 starts one (or s is the last slot)
  uint32_t next_first = __shfl_down_sync(0xffffffffu, start_mask & 1u, 1);
  if (lane == 31) next_first = 1u;
  const uint32_t end_mask = (E == 32) ? ((start_mask >> 1) | (next_first << 31))
                                      : (((start_mask >> 1) | (next_first << (E - 1))) & ((E == 32) ? 0xffffffffu : ((1u << E) - 1u)));
  int first = 0x7fffffff;
  if (end_mask) first = lane * E + (__ffs(end_mask) - 1);
  int dec = first;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_down_sync(0xffffffffu, dec, o);
    if (lane + o < 32) dec = min(dec, t);
  }
  cur = __shfl_down_sync(0xffffffffu, dec, 1);
  if (lane == 31) cur = 0x7fffffff;
#pragma unroll
  for (int r = E - 1; r >= 0; --r) {
    const int s = lane * E + r;
    if (end_mask & (1u << r)) cur = s;
    // lo + hi + 2 == 2 * average 1-based rank; hi is clamped to the last real slot (+inf ties with padding)
    if (idx[r] < (uint32_t)n) rk2[idx[r]] = (uint16_t)(lo[r] + (uint32_t)min(cur, n - 1) + 2u);
  }
}

// Exact form: sort on (full key, sample index).  Out of line: it only runs for columns the packed fast path hands back
// (values closer than 1024 ulp that are not equal), and keeping it out of the hot code keeps that within the instruction cache.
template <int E>
__device__ __noinline__ void rank_exact_warp(const double* __restrict__ col, int n, int lane, uint16_t* rk2) {
  SortKI ki[E];
#pragma unroll
  for (int r = 0; r < E; ++r) {
    const int i = r * 32 + lane;
    ki[r].k = (i < n) ? col[i] : __longlong_as_double(0x7ff0000000000000ll);
    ki[r].i = (uint32_t)i;
  }
  bitonic_sort_warp<E>(ki, lane);
  const double prev_last = __shfl_up_sync(0xffffffffu, ki[E - 1].k, 1);
  uint32_t start_mask = 0, idx[E];
  int last = -1;
#pragma unroll
  for (int r = 0; r < E; ++r) {
    const int s = lane * E + r;
    const double prev = (r == 0) ? prev_last : ki[r - 1].k;
    idx[r] = ki[r].i;
    if ((s == 0) || (ki[r].k != prev)) {
      start_mask |= 1u << r;
      last = s;
    }
  }
  assign_ranks_warp<E>(start_mask, last, idx, n, lane, rk2);
}

template <int E, bool PEERS>
__device__ __forceinline__ void rank_columns_warp_body(const double* __restrict__ x, int64_t ldx, int n, int64_t p, const ColumnOut& out,
                                                       const PeerOut* po) {
  constexpr int NP = 32 * E;
  __shared__ uint16_t rk2s[4][NP];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t j = (int64_t)blockIdx.x * 4 + warp;
  if (j >= p) return;  // whole warp leaves; no block-level barrier below
  const double* col = x + j * ldx;
  uint16_t* rk2 = rk2s[warp];

  // Padding slots hold +inf with an index >= n.  A real +inf ties with the padding, which is
  // harmless: the tie run then starts at the first real +inf and its end is clamped to n - 1.
  //
  // Fast path (packed keys).  The sort only needs the ORDER of the values; ties are resolved afterwards.  Every element
  // becomes one double: the value with -0 folded into +0, +-inf replaced by +-DBL_MAX, its low 10 mantissa bits replaced by
  // the sample index (< 1024).  All packed keys of a column are distinct, they order like the values wherever the values
  // differ above those 10 bits, and elements that agree above them (true ties, or values closer than 1024 ulp) end up
  // adjacent.  The replaced bits (plus an "was infinite" flag) are parked in shared memory by sample index; after the sort
  // every neighbour pair that agrees above the index bits is checked there: equal parked bits = the same value = a true tie
  // (the only case in practice: tied data is tied exactly).  If any such pair differs, the column is redone with the exact
  // (key, index) sort.  Bit-exact either way.
  SortPK pk[E];
  bool has_nan = false;
#pragma unroll
  for (int r = 0; r < E; ++r) {
    const int i = r * 32 + lane;  // coalesced load; the initial order is irrelevant to a sort
    double v = __longlong_as_double(0x7ff0000000000000ll);
    if (i < n) {
      v = col[i];
      if (v != v) has_nan = true;
    }
    unsigned long long b = (unsigned long long)__double_as_longlong(v + 0.0);
    uint32_t parked = (uint32_t)b & 1023u;
    if ((b & 0x7ff0000000000000ull) == 0x7ff0000000000000ull) {
      b = (b & 0x8000000000000000ull) | 0x7fefffffffffffffull;
      parked = 1024u | 1023u;
    }
    if (BIXCOR_RANK_PACKED) rk2[i] = (uint16_t)parked;
    pk[r].k = __longlong_as_double((long long)((b & ~0x3ffull) | (unsigned long long)i));
  }
  if (__any_sync(0xffffffffu, has_nan)) {
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);
    if (out.mode == COL_RANKS) {
      for (int i = lane; i < n; i += 32) out.f64[j * out.ld64 + i] = qnan;
    } else if (out.mode == COL_Z64) {
      for (int i = lane; i < out.kpad; i += 32) out.f64[j * out.ld64 + i] = (i < n) ? qnan : 0.0;
    } else if (out.mode == COL_I8) {
      for (int i = lane; i < out.kpad; i += 32) {
        out.i8[j * out.ld8 + i] = 0;
        out.i8[out.plane8 + j * out.ld8 + i] = 0;
        if (PEERS)
          for (int r = 0; r < po->n; ++r) {
            po->i8[r][j * out.ld8 + i] = 0;
            po->i8[r][out.plane8 + j * out.ld8 + i] = 0;
          }
      }
      if (lane == 0) {
        out.inv_norm[j] = qnan;
        if (PEERS)
          for (int r = 0; r < po->n; ++r) po->inv_norm[r][j] = qnan;
      }
    } else {
      for (int i = lane; i < out.kpad; i += 32) {
        if (out.half_planes) {
          emit_tf32(out, j, i, (i < n) ? qnan : 0.0);
          continue;
        }
        out.f32[j * out.ld32 + i] = (i < n) ? __int_as_float(0x7fc00000) : 0.f;
        out.f32[out.plane32 + j * out.ld32 + i] = 0.f;
      }
    }
    return;
  }
  bool redo = !BIXCOR_RANK_PACKED;
  if (BIXCOR_RANK_PACKED) {
    __syncwarp();  // parked bits visible to every lane
    bitonic_sort_warp<E>(pk, lane);
    const unsigned long long prev_last = (unsigned long long)__double_as_longlong(__shfl_up_sync(0xffffffffu, pk[E - 1].k, 1));
    uint32_t start_mask = 0, idx[E];  // bit r: position lane*E + r starts a run of equal values
    int last = -1;
    uint32_t prev_parked = rk2[(uint32_t)prev_last & 1023u];
    bool mixed = false;
#pragma unroll
    for (int r = 0; r < E; ++r) {
      const int s = lane * E + r;
      const unsigned long long cur = (unsigned long long)__double_as_longlong(pk[r].k);
      const unsigned long long prv = (r == 0) ? prev_last : (unsigned long long)__double_as_longlong(pk[r - 1].k);
      idx[r] = (uint32_t)cur & 1023u;
      const uint32_t parked = rk2[idx[r]];
      const bool same = (s != 0) && ((cur ^ prv) < 1024ull);  // agree above the index bits
      if (same && parked != prev_parked) mixed = true;        // ... but are different values: needs the exact sort
      if (!same) {
        start_mask |= 1u << r;
        last = s;
      }
      prev_parked = parked;
    }
    redo = __any_sync(0xffffffffu, mixed);  // (also orders the reads of the parked bits before the rank writes below)
    if (!redo) assign_ranks_warp<E>(start_mask, last, idx, n, lane, rk2);
  }
  if (redo) rank_exact_warp<E>(col, n, lane, rk2);
  __syncwarp();
  emit_ranked_column_warp<uint16_t, PEERS>(out, j, n, rk2, lane, po);
}

// Measured on 20 000 genes x 1 000 samples (scripts/rank_probe.py, profiles/rank_probe_r02.txt): 3 CTAs per SM (167 registers, no
// spills) 0.247 ms, 4 CTAs per SM (capped at 128 registers, 12 bytes of spills) 0.260 ms; the (key, index) sort of round 1: 0.303 ms
#ifndef BIXCOR_RANK_MIN_CTAS
#define BIXCOR_RANK_MIN_CTAS 3
#endif
template <int E>
__global__ void __launch_bounds__(128, BIXCOR_RANK_MIN_CTAS)
rank_columns_warp_kernel(const double* __restrict__ x, int64_t ldx, int n, int64_t p, ColumnOut out) {
  rank_columns_warp_body<E, false>(x, ldx, n, p, out, nullptr);
}
// the same kernel with the peer stores of the fused operand exchange (COL_I8 only)
template <int E>
__global__ void __launch_bounds__(128, BIXCOR_RANK_MIN_CTAS)
rank_columns_warp_peers_kernel(const double* __restrict__ x, int64_t ldx, int n, int64_t p, ColumnOut out,
                               const __grid_constant__ PeerOut peers) {
  rank_columns_warp_body<E, true>(x, ldx, n, p, out, &peers);
}

// Generic form of the exchange for the column kernels without fused peer stores: copies `bytes` (a multiple of 16) from the
// local buffer to the same offset of every peer buffer with 128-bit stores.
struct PeerCopy {
  int n;
  char* dst[kMaxPeers];
};
__global__ void __launch_bounds__(256)
peer_broadcast_kernel(const char* __restrict__ src, const __grid_constant__ PeerCopy peers, int64_t bytes) {
  const int64_t n16 = bytes >> 4, stride = (int64_t)gridDim.x * blockDim.x;
  const int4* s16 = reinterpret_cast<const int4*>(src);
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n16; t += stride) {
    const int4 v = s16[t];
    for (int r = 0; r < peers.n; ++r) reinterpret_cast<int4*>(peers.dst[r])[t] = v;
  }
}

// Cross-GPU barrier of the exchange (one CTA): thread r tells peer r that this rank's records of step `epoch` are complete
// (system-scope release store into the peer's flag array; the column kernel's peer stores were performed before this
// kernel started) and waits until peer r has said the same here.  A peer that never arrives (a dead process) traps after ~2 minutes
// instead of hanging the GPU for good; ordinary skew between ranks (first-call set-up, a slow host) stays far below that.
struct PeerFlags {
  int n;                        // world size (own rank included)
  uint32_t* flags[kMaxPeers];   // flag array of every rank (16 words = 64 bytes apart per writer)
};
__global__ void __launch_bounds__(32)
peer_barrier_kernel(const __grid_constant__ PeerFlags pf, int rank, uint32_t epoch) {
  const int r = threadIdx.x;
  if (r >= pf.n) return;
  __threadfence_system();
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(pf.flags[r] + rank * 16), "r"(epoch) : "memory");
  const uint32_t* mine = pf.flags[rank] + r * 16;
  const long long t0 = clock64();
  for (;;) {
    uint32_t seen;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(mine) : "memory");
    if ((int32_t)(seen - epoch) >= 0) break;
    if (clock64() - t0 > 240000000000ll) __trap();  // ~2 minutes at 2 GHz
    __nanosleep(100);
  }
}

// Pearson: centre and scale one gene per CTA to unit Euclidean norm, so that
// Z'Z is the correlation matrix (sd == 0 -> 0/0 = NaN column, IEEE propagation).
// norm_mode: 0 = centre + unit norm (correlation), 1 = centre only (covariance),
//            2 = unit norm without centring (cosine)
__global__ void __launch_bounds__(128)
standardise_columns_kernel(const double* __restrict__ x, int64_t ldx, int n, ColumnOut out, int norm_mode) {
  __shared__ double scr[32];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int64_t j = blockIdx.x;
  const double* col = x + j * ldx;
  double s = 0.0;
  for (int i = tid; i < n; i += nt) s += col[i];
  const double mean = (norm_mode == 2) ? 0.0 : block_sum(s, scr) / (double)n;
  double q = 0.0;
  for (int i = tid; i < n; i += nt) {
    const double c = col[i] - mean;
    q += c * c;
  }
  const double norm = (norm_mode == 1) ? 1.0 : sqrt(block_sum(q, scr));
  for (int i = tid; i < out.kpad; i += nt) {
    double z = 0.0;
    if (i < n) z = (col[i] - mean) / norm;
    if (out.mode == COL_Z64) out.f64[j * out.ld64 + i] = z;
    else emit_tf32(out, j, i, z);
  }
}

// Same for n <= 2048 with one trip through HBM: the column is read ONCE with 128-bit loads (16 bytes per lane,
// coalesced), kept in registers (up to 8 double2 per thread), reduced by warp shuffles + a block step, and the
// operand is written with 128-bit stores.  Requires n even and 16-byte aligned columns (host checks).
__global__ void __launch_bounds__(128)
standardise_columns_reg_kernel(const double* __restrict__ x, int64_t ldx, int n, ColumnOut out, int norm_mode) {
  __shared__ double scr[32];
  const int tid = threadIdx.x;
  const int64_t j = blockIdx.x;
  const double2* col = reinterpret_cast<const double2*>(x + j * ldx);
  const int n2 = n >> 1;
  double2 v[8];
  double s = 0.0;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int i2 = r * 128 + tid;
    v[r] = (i2 < n2) ? __ldg(col + i2) : make_double2(0.0, 0.0);
    s += v[r].x + v[r].y;
  }
  const double mean = (norm_mode == 2) ? 0.0 : block_sum(s, scr) / (double)n;
  double q = 0.0;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int i2 = r * 128 + tid;
    if (i2 < n2) {
      v[r].x -= mean;
      v[r].y -= mean;
      q += v[r].x * v[r].x + v[r].y * v[r].y;
    }
  }
  const double norm = (norm_mode == 1) ? 1.0 : sqrt(block_sum(q, scr));
  const int k2 = out.kpad >> 1;
#pragma unroll
  for (int r = 0; r < 9; ++r) {  // kpad <= n + 31: one extra trip covers the zero padding
    const int i2 = r * 128 + tid;
    if (i2 >= k2) break;
    double2 z = make_double2(0.0, 0.0);
    if (r < 8 && i2 < n2) z = make_double2(v[r].x / norm, v[r].y / norm);
    if (out.mode == COL_Z64) {
      reinterpret_cast<double2*>(out.f64 + j * out.ld64)[i2] = z;
    } else if (out.half_planes) {
      __half h0, l0, h1, l1;
      split_f16(z.x, h0, l0);
      split_f16(z.y, h1, l1);
      __half* hp = reinterpret_cast<__half*>(out.f32);
      reinterpret_cast<__half2*>(hp + j * out.ld32)[i2] = __halves2half2(h0, h1);
      reinterpret_cast<__half2*>(hp + out.plane32 + j * out.ld32)[i2] = __halves2half2(l0, l1);
    } else {
      const float h0 = to_tf32((float)z.x), h1 = to_tf32((float)z.y);
      const float l0 = to_tf32((float)(z.x - (double)h0)), l1 = to_tf32((float)(z.y - (double)h1));
      reinterpret_cast<float2*>(out.f32 + j * out.ld32)[i2] = make_float2(h0, h1);
      reinterpret_cast<float2*>(out.f32 + out.plane32 + j * out.ld32)[i2] = make_float2(l0, l1);
    }
  }
}

// ----------------------------------------------------------------------------
// Ozaki-style operand for the FP64-accurate Gram on the int8 tensor cores (BIXCOR_PREC_OZAKI).
// Row g of an f64 matrix becomes scale[g] = 2^e with |z / scale| < 0.5 and OZ_SLICES balanced base-128 digits per
// element, z / scale = sum_t d_t 128^-(t+1) + O(128^-6), d_t in [-64, 64] (int8).  Record layout: [OZ_SLICES][kpad]
// int8 per row, so planes (2a, 2a+1) form the digit pair "group a" that one pass of the int8 kernel multiplies.
// All arithmetic here is exact in FP64 (scaling by powers of two, rint, subtraction of the rounded part).
// ----------------------------------------------------------------------------
constexpr int OZ_SLICES = 6;

__global__ void __launch_bounds__(128)
ozaki_slice_kernel(const double* __restrict__ z, int64_t ld, int k, int kpad, int8_t* __restrict__ planes,
                   double* __restrict__ scale) {
  __shared__ double scr[4];
  __shared__ int s_bad;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t g = blockIdx.x;
  const double* src = z + g * ld;
  if (tid == 0) s_bad = 0;
  __syncthreads();
  double amax = 0.0;
  bool bad = false;
  for (int i = tid; i < k; i += 128) {
    const double v = src[i];
    if (!(fabs(v) <= 1.7976931348623157e308)) bad = true;  // NaN or inf
    amax = fmax(amax, fabs(v));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  if (lane == 0) scr[warp] = amax;
  if (bad) s_bad = 1;
  __syncthreads();
  amax = fmax(fmax(scr[0], scr[1]), fmax(scr[2], scr[3]));
  double s = 0.0, inv_s = 0.0;
  if (amax > 0.0 && !s_bad) {
    int ex;
    frexp(amax, &ex);        // amax = m * 2^ex, m in [0.5, 1)
    s = ldexp(1.0, ex + 1);  // |z| / s < 0.5
    inv_s = ldexp(1.0, -(ex + 1));
  }
  if (tid == 0) scale[g] = s_bad ? __longlong_as_double(0x7ff8000000000000ll) : s;
  int8_t* rec = planes + g * (int64_t)OZ_SLICES * kpad;
  for (int i4 = tid * 4; i4 < kpad; i4 += 512) {  // four elements per thread: one 32-bit store per plane
    uint32_t w[OZ_SLICES];
#pragma unroll
    for (int t = 0; t < OZ_SLICES; ++t) w[t] = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = i4 + q;
      double v = (i < k && inv_s != 0.0) ? src[i] * inv_s : 0.0;
#pragma unroll
      for (int t = 0; t < OZ_SLICES; ++t) {
        const double tt = v * 128.0;
        const double d = rint(tt);
        v = tt - d;
        w[t] |= ((uint32_t)(int)d & 0xffu) << (8 * q);
      }
    }
#pragma unroll
    for (int t = 0; t < OZ_SLICES; ++t) *reinterpret_cast<uint32_t*>(rec + (int64_t)t * kpad + i4) = w[t];
  }
}

// Finishers of the Ozaki path: G holds sum_ab w_ab * (digit-pair Gram) for the upper triangle (dense, ld = p);
// the true Gram value is scale[i] * scale[j] * G[i][j].
__global__ void __launch_bounds__(256)
ozaki_finish_packed_kernel(const double* __restrict__ G, const double* __restrict__ scale, EpiParams E, int64_t row0,
                           int64_t row1) {
  const int64_t i = row0 + blockIdx.y;
  if (i >= row1) return;
  const double si = scale[i];
  const int64_t ro = packed_row_offset(i, E.p, E.shift) - E.out_base - i - E.shift;
  for (int64_t j = i + E.shift + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < E.p; j += (int64_t)gridDim.x * blockDim.x)
    E.out[ro + j] = epi_power(E, G[i * E.p + j] * si * scale[j] * E.scale);
}

__global__ void __launch_bounds__(256)
ozaki_finish_dense_kernel(const double* __restrict__ G, const double* __restrict__ scale, EpiParams E, int64_t row0,
                          int64_t row1) {
  const int64_t i = row0 + blockIdx.y;
  if (i >= row1) return;
  const double si = scale[i];
  const int64_t jstart = E.mirror ? i : 0;
  for (int64_t j = jstart + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < E.p; j += (int64_t)gridDim.x * blockDim.x) {
    double w = epi_power(E, G[i * E.p + j] * si * scale[j] * E.scale);
    if (i == j && E.diag_one) w = 1.0;
    E.out[i * E.ldo + j] = w;
    if (E.mirror && j != i) E.out[j * E.ldo + i] = w;
  }
}

__global__ void __launch_bounds__(256)
ozaki_finish_tom_kernel(const double* __restrict__ G, const double* __restrict__ scale, EpiParams E, int64_t row0,
                        int64_t row1) {
  const int64_t i = row0 + blockIdx.y;
  if (i >= row1) return;
  EpiRow R;
  R.i = (int)i;
  R.ro = packed_row_offset(i, E.p, 1) - E.out_base - i - 1;
  R.k_i = E.kvec[i];
  R.d_i = E.dvec[i];
  const double si = scale[i];
  const int64_t jstart = (E.tom_packed || E.mirror) ? i : 0;
  for (int64_t j = jstart + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < E.p; j += (int64_t)gridDim.x * blockDim.x)
    epi_store<EPI_TOM>(E, R, (int)j, G[i * E.p + j] * si * scale[j], 0);
}

// Split an existing f64 matrix (ld = ldin, `rows` x `k`) into TF32 hi/lo planes
// (used for the TOM A*A fast path, where the operand is the affinity matrix itself).
// binary16 planes: an entry with |z| >= 63 does not fit (2^10 z overflows) and raises *overflow.
__global__ void split_tf32_kernel(const double* __restrict__ in, int64_t ldin, int k, ColumnOut out, int* overflow) {
  const int64_t j = blockIdx.x;
  for (int i = threadIdx.x; i < out.kpad; i += blockDim.x) {
    double z = (i < k) ? in[j * ldin + i] : 0.0;
    if (out.half_planes && fabs(z) >= 63.0) atomicExch(overflow, 1);
    emit_tf32(out, j, i, z);
  }
}

}  // namespace bixcor
