// libbixcor: C-ABI + host runtime of the B200-native co-expression correlation path.
// See include/bixcor.h for the boundary contract and the reference call sites replaced.
#include "../../include/bixcor.h"

#include <cuda_runtime.h>
#include <emmintrin.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <stdexcept>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

#include "aux_kernels.cuh"
#include "columns.cuh"
#include "common.cuh"
#include "gram_dmma.cuh"
#include "gram_umma.cuh"
#include "nccl_dyn.h"

namespace bixcor {

// ----------------------------------------------------------------------------
// errors
// ----------------------------------------------------------------------------
static std::mutex g_err_mu;
static char g_err[1024] = "";  // fixed storage: reporting an error never allocates (it may be reporting std::bad_alloc)
static std::atomic<int64_t> g_launches{0};
static std::atomic<int64_t> g_h2d_bytes{0}, g_d2h_bytes{0};  // bytes that crossed PCIe through the host entry points (bixcor_transfer_counters)
static double g_fieller = 1.06;

static int fail(int code, const char* fmt, ...) noexcept {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  try {
    std::lock_guard<std::mutex> lk(g_err_mu);
    memcpy(g_err, buf, sizeof g_err);
  } catch (...) {
  }
  return code;
}

// ---- exception barrier of the C ABI (include/bixcor.h: "No C++ exception crosses the ABI") -------------------
// Every extern "C" body runs between API_BEGIN and API_END: std::bad_alloc -> BIXCOR_ERR_OOM, anything else ->
// BIXCOR_ERR_INTERNAL, message in bixcor_last_error().  g_inject is the test hook behind bixcor_debug_inject().
static std::atomic<int> g_inject{0};
static void inject_point() {
  const int what = g_inject.exchange(0);
  if (what == 1) throw std::bad_alloc();
  if (what == 2) throw std::runtime_error("injected failure (bixcor_debug_inject)");
}
static int api_exception(const char* fn) noexcept {
  try {
    throw;
  } catch (const std::bad_alloc&) {
    return fail(BIXCOR_ERR_OOM, "%s: host allocation failed (std::bad_alloc)", fn);
  } catch (const std::exception& e) {
    return fail(BIXCOR_ERR_INTERNAL, "%s: C++ exception stopped at the ABI: %s", fn, e.what());
  } catch (...) {
    return fail(BIXCOR_ERR_INTERNAL, "%s: unknown C++ exception stopped at the ABI", fn);
  }
}
#define API_BEGIN \
  try {           \
    inject_point();
#define API_END                       \
  }                                   \
  catch (...) {                       \
    return api_exception(__func__);   \
  }
#define API_END_VAL(v)                \
  }                                   \
  catch (...) {                       \
    api_exception(__func__);          \
    return (v);                       \
  }

// Host worker threads that are always joined, also when a later std::thread constructor or the body between
// creation and join throws (a joinable std::thread destroyed un-joined calls std::terminate).
struct ThreadGroup {
  std::vector<std::thread> th;
  template <class F>
  void run(F&& f) {
    th.emplace_back(std::forward<F>(f));
  }
  void join() {
    for (auto& t : th)
      if (t.joinable()) t.join();
    th.clear();
  }
  ~ThreadGroup() { join(); }
};

#define CU_TRY(expr)                                                                                    \
  do {                                                                                                  \
    cudaError_t e__ = (expr);                                                                           \
    if (e__ != cudaSuccess)                                                                             \
      return fail(e__ == cudaErrorMemoryAllocation ? BIXCOR_ERR_OOM : BIXCOR_ERR_CUDA, "%s:%d %s: %s",  \
                  __FILE__, __LINE__, #expr, cudaGetErrorString(e__));                                  \
  } while (0)
#define RC_TRY(expr)        \
  do {                      \
    int rc__ = (expr);      \
    if (rc__ != BIXCOR_OK) return rc__; \
  } while (0)

// ----------------------------------------------------------------------------
// per-device context
// ----------------------------------------------------------------------------
struct Buffer {
  void* ptr = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return BIXCOR_OK;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
    size_t want = (bytes + (size_t(1) << 20) - 1) & ~((size_t(1) << 20) - 1);
    CU_TRY(cudaMalloc(&ptr, want));
    cap = want;
    return BIXCOR_OK;
  }
  void release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
  }
};

struct PinnedBuffer {
  void* ptr = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return BIXCOR_OK;
    if (ptr) cudaFreeHost(ptr);
    ptr = nullptr;
    cap = 0;
    CU_TRY(cudaMallocHost(&ptr, bytes));
    cap = bytes;
    return BIXCOR_OK;
  }
  void release() {
    if (ptr) cudaFreeHost(ptr);
    ptr = nullptr;
    cap = 0;
  }
};

struct TileList {
  int2* d_tiles = nullptr;
  int ntiles = 0;
  // group g covers tiles [group_tile[g], group_tile[g+1]) and rows [group_row[g], group_row[g+1])
  std::vector<int> group_tile;
  std::vector<int64_t> group_row;
};

enum TimeCat : int { T_COLS = 0, T_GEMM = 1, T_OTHER = 2 };

struct HostStager;  // pinned ring + copier threads for pageable host buffers (below)

// Fused operand exchange between the GPUs of one box, one process per GPU (bixcor_xchg_*): one cudaMalloc block per rank,
// exported with cudaIpc and mapped by every peer -
//   [operand buffer 0][operand buffer 1][1/|d| buffer 0][1/|d| buffer 1][flags: world x 64 bytes]
// The column kernel of rank r stores the records of ITS gene slab into buffer (epoch & 1) of every rank; a one-CTA barrier
// kernel (system-scope flags in the same block) tells every peer that the slab is complete and waits for theirs; the Gram
// kernel then reads the local buffer.  Two buffers: a peer may already be writing step t + 1 while this rank's Gram of step t
// is still reading; it cannot reach step t + 2 before this rank has signalled step t + 1, i.e. finished the Gram of step t.
struct Xchg {
  int world = 0, rank = 0, precision = 0, opened = 0;
  int64_t n = 0, p = 0, slab = 0;
  size_t rb = 0, op_off[2] = {0, 0}, inv_off[2] = {0, 0}, flag_off = 0, total = 0;
  char* local = nullptr;
  char* peer[kMaxPeers] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // peer[rank] == local
  uint32_t epoch = 0;
};

struct Ctx {
  int device = -1;
  int sm_count = 148;
  cudaStream_t compute = nullptr, copy_in = nullptr, copy_out = nullptr;
  bool owns_compute = true;
  Buffer x[2];          // staged expression matrices
  Buffer op[2];         // Gram operands (Z / digit planes / TF32 planes)
  Buffer aux[2];        // per-gene vectors (inv norm)
  Buffer oz[2];         // Ozaki digit planes (+ per-row scale at the end)
  Buffer ozg;           // Ozaki FP64 accumulation matrix (p x p)
  Buffer out;           // packed / dense result
  Buffer out2;          // affinity matrix for TOM
  Buffer out32;         // binary32 copy of finished bands (f32 transport of the <= 1e-5-class results)
  Buffer vec;           // k, diag, colsum
  Buffer panel;         // sparse dense panel
  Buffer sparse_in;     // staged CSR
  Buffer flag;          // one int: device-side overflow / range flags read back by the host
  Buffer rank_scratch;  // sorted chunks / ranks of the n > 16384 Spearman path
  Buffer tile_sync;     // two words: arrival / bail-out counters of the long-K producer barrier (gram_umma.cuh)
  PinnedBuffer inv_host;            // host copy of the per-gene 1/|d| (int32 transport of the exact Spearman Gram)
  Xchg xchg;                        // fused operand exchange over P2P-mapped peer memory (one process per GPU)
  HostStager* stager = nullptr;     // device -> pageable host (results)
  HostStager* stager_in = nullptr;  // pageable host -> device (inputs): its own small ring, so uploads never queue behind result slots
  std::map<std::tuple<int64_t, int64_t, int64_t, int, int>, TileList> tile_cache;
  // timing
  std::vector<std::tuple<int, cudaEvent_t, cudaEvent_t>> timed;
  std::vector<cudaEvent_t> event_pool;
  size_t events_used = 0;
  double last_ms[3] = {0, 0, 0};
  int last_gemm_launches = 0;
};

static std::vector<Ctx*> g_ctx;
// NCCL communicators over the devices of this process (bixcor_init(n), n > 1): g_comms[d] belongs to g_ctx[d].
// Empty when there is one device, when libnccl cannot be opened or with env BIXCOR_NCCL=0 (the multi-device entry
// points then exchange with peer copies / through the host, as in round 1).
static NcclApi g_nccl;
static std::vector<ncclComm_t> g_comms;
#define NCCL_TRY(expr)                                                                                        \
  do {                                                                                                        \
    ncclResult_t r__ = (expr);                                                                                \
    if (r__ != ncclSuccess)                                                                                   \
      return fail(BIXCOR_ERR_CUDA, "%s:%d %s: NCCL %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(r__)); \
  } while (0)
static void nccl_teardown() {
  for (auto cm : g_comms)
    if (cm) g_nccl.CommDestroy(cm);
  g_comms.clear();
}
static std::mutex g_init_mu;
static int g_dmma_variant = 1;  // dmma::Shape<CFG> (1 = 128x64 tiles, 2 CTAs/SM: measured fastest), env BIXCOR_DMMA_VARIANT
static int g_umma_mc = 0;       // tcgen05 long-K kernels in 2 x 2 multicast clusters (env BIXCOR_UMMA_MC = 1): verified, measured, NOT
                                // faster (20k TOM f16 25.2 vs 23.2 ms, TF32 44.4 vs 42.1 ms, profiles/README_r02.md): L2 already merges the
                                // requests of up to ~4 SMs for the same lines, so 2-way multicast saves no L2 -> SM bandwidth
static int g_umma_wide = 0;     // TOM on 256 x 256 pair tiles (env BIXCOR_UMMA_WIDE): 0 never (default), 1 = the TF32 kind (20k TOM: 40.3-41.9 ms
                                // against 41.2-42.1, inside the run-to-run spread), 2 = the f16 kind as well (slower: 25.2 vs 23.2 ms)
static int g_umma_ctas = 2;     // tcgen05 kernels: CTA pairs (cta_group::2) 1 = never, 2 = auto, 3 = always; env BIXCOR_UMMA_CTAS
static bool g_timing = false;   // bixcor_timing_enable
static bool g_async = false;    // device-pointer entry points return without synchronising

static void stager_destroy(Ctx* c);
static void xchg_free(Ctx* c);

static int ctx_create(int device, Ctx** out) {
  Ctx* c = new Ctx();
  c->device = device;
  CU_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    delete c;
    return fail(BIXCOR_ERR_NO_GPU, "device %d is sm_%d%d; libbixcor is built for sm_100a only", device, prop.major,
                prop.minor);
  }
  c->sm_count = prop.multiProcessorCount;
  CU_TRY(cudaStreamCreateWithFlags(&c->compute, cudaStreamNonBlocking));
  CU_TRY(cudaStreamCreateWithFlags(&c->copy_in, cudaStreamNonBlocking));
  CU_TRY(cudaStreamCreateWithFlags(&c->copy_out, cudaStreamNonBlocking));
  *out = c;
  return BIXCOR_OK;
}

static void ctx_destroy(Ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (auto& b : c->x) b.release();
  for (auto& b : c->op) b.release();
  for (auto& b : c->aux) b.release();
  for (auto& b : c->oz) b.release();
  c->ozg.release();
  c->out.release();
  c->out2.release();
  c->out32.release();
  c->vec.release();
  c->panel.release();
  c->sparse_in.release();
  c->flag.release();
  c->tile_sync.release();
  c->rank_scratch.release();
  stager_destroy(c);
  xchg_free(c);
  c->inv_host.release();
  for (auto& kv : c->tile_cache)
    if (kv.second.d_tiles) cudaFree(kv.second.d_tiles);
  for (auto e : c->event_pool) cudaEventDestroy(e);
  if (c->compute && c->owns_compute) cudaStreamDestroy(c->compute);
  if (c->copy_in) cudaStreamDestroy(c->copy_in);
  if (c->copy_out) cudaStreamDestroy(c->copy_out);
  delete c;
}

static int ensure_init() {
  {
    std::lock_guard<std::mutex> lk(g_init_mu);
    if (!g_ctx.empty()) return BIXCOR_OK;
  }
  return bixcor_init_devices(0, 1);
}

// ---- timing helpers ---------------------------------------------------------
static cudaEvent_t ctx_event(Ctx* c) {
  if (c->events_used == c->event_pool.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    c->event_pool.push_back(e);
  }
  return c->event_pool[c->events_used++];
}
static void timing_reset(Ctx* c) {
  c->timed.clear();
  c->events_used = 0;
  c->last_ms[0] = c->last_ms[1] = c->last_ms[2] = 0;
  c->last_gemm_launches = 0;
}
// per-call housekeeping: the event pool is recycled unless the caller is accumulating timings
static void call_begin(Ctx* c) {
  if (!g_timing) c->events_used = 0;
}
struct Timed {
  Ctx* c;
  int cat;
  cudaEvent_t e0, e1;
  Timed(Ctx* c_, int cat_) : c(c_), cat(cat_), e0(nullptr), e1(nullptr) {
    if (!g_timing || c->events_used > 16384) return;
    e0 = ctx_event(c);
    e1 = ctx_event(c);
    cudaEventRecord(e0, c->compute);
  }
  ~Timed() {
    if (cat == T_GEMM) c->last_gemm_launches++;
    if (!e0) return;
    cudaEventRecord(e1, c->compute);
    try {
      c->timed.emplace_back(cat, e0, e1);
    } catch (...) {  // (a destructor must not throw: the sample is dropped)
    }
  }
};
static void timing_collect(Ctx* c) {
  for (auto& t : c->timed) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, std::get<1>(t), std::get<2>(t)) == cudaSuccess) c->last_ms[std::get<0>(t)] += ms;
  }
  c->timed.clear();
}

// ----------------------------------------------------------------------------
// tile lists: upper-triangle (or full strip) tiles of the block rows covering
// [row_begin,row_end), ordered in bands of `band` block rows so that the CTAs of
// one wave share operand rows in L2 and every band completes one contiguous
// packed range (the unit of the D2H pipeline).
// ----------------------------------------------------------------------------
// pairs != 0: entries are CTA-pair super-tiles of the 2-CTA tcgen05 kernels, (first block row, block col) with the
// second block row = first + 1 and umma::TILE_SECOND_INVALID set on the column when that row holds nothing to
// compute (outside [row_begin,row_end) or entirely below the diagonal).
// swap (packed correlation on the tcgen05 kernels, lean-eligible epilogue): interior tiles - strictly above the diagonal
// and fully inside the matrix - are emitted transposed (umma::TILE_SWAPPED: one block row x one / two block columns),
// the diagonal and ragged-edge tiles in the plain geometry.
static int get_tiles(Ctx* c, int64_t p, int64_t row_begin, int64_t row_end, int upper_only, int band, int pairs,
                     TileList** out, int swap = 0) {
  if (!upper_only) swap = 0;
  auto key = std::make_tuple(p, row_begin, row_end, upper_only * 2 + (pairs == 1 ? 1 : 0) + (swap ? 4 : 0) + (pairs == 2 ? 8 : 0) + (pairs == 3 ? 16 : 0), band);
  auto it = c->tile_cache.find(key);
  if (it != c->tile_cache.end()) {
    *out = &it->second;
    return BIXCOR_OK;
  }
  if (c->tile_cache.size() > 64) {
    for (auto& kv : c->tile_cache)
      if (kv.second.d_tiles) cudaFree(kv.second.d_tiles);
    c->tile_cache.clear();
  }
  const int nb = (int)((p + kTile - 1) / kTile);
  const int b0 = (int)(row_begin / kTile), b1 = (int)((row_end + kTile - 1) / kTile);
  std::vector<int2> tiles;
  TileList tl;
  for (int bb = b0; bb < b1; bb += band) {
    const int be = std::min(b1, bb + band);
    tl.group_tile.push_back((int)tiles.size());
    tl.group_row.push_back(std::max<int64_t>(row_begin, (int64_t)bb * kTile));
    const int j0 = upper_only ? bb : 0;
    if (swap) {
      const int nfull = (int)(p / kTile);  // block columns [0, nfull) are complete
      for (int bi = bb; bi < be; ++bi) {
        // plain geometry: the diagonal tile and, where the last block column is ragged, that column
        tiles.push_back(make_int2(bi, pairs ? (bi | umma::TILE_SECOND_INVALID) : bi));
        if (nfull < nb && bi < nb - 1) tiles.push_back(make_int2(bi, pairs ? ((nb - 1) | umma::TILE_SECOND_INVALID) : nb - 1));
        // transposed geometry: block columns bi + 1 .. nfull - 1, two per CTA pair
        for (int bj = bi + 1; bj < nfull; bj += pairs ? 2 : 1) {
          int y = bj | umma::TILE_SWAPPED;
          if (pairs && bj + 1 >= nfull) y |= umma::TILE_SECOND_INVALID;
          tiles.push_back(make_int2(bi, y));
        }
      }
      continue;
    }
    if (pairs == 3) {  // wide pair tiles: (first block row of the pair, 256-column block | flags); the epilogue predicates j > i
      const int nw = (nb + 1) / 2;
      for (int wj = upper_only ? bb / 2 : 0; wj < nw; ++wj)
        for (int bi = bb; bi < be; bi += 2) {
          if (upper_only && wj < bi / 2) continue;
          const bool second_ok = (bi + 1 < b1) && (!upper_only || wj >= (bi + 1) / 2);
          tiles.push_back(make_int2(bi, second_ok ? wj : (wj | umma::TILE_SECOND_INVALID)));
        }
      continue;
    }
    if (pairs == 2) {  // 2 x 2 multicast super-tiles (first block row, first block column); validity is decided per CTA in the kernel
      for (int bj = upper_only ? bb : 0; bj < nb; bj += 2)
        for (int bi = bb; bi < be; bi += 2)
          if (!upper_only || bi <= bj) tiles.push_back(make_int2(bi, bj));
      continue;
    }
    for (int bj = j0; bj < nb; ++bj) {
      if (!pairs) {
        for (int bi = bb; bi < be; ++bi)
          if (!upper_only || bi <= bj) tiles.push_back(make_int2(bi, bj));
      } else {
        for (int bi = bb; bi < be; bi += 2) {  // bands start at b0 and hold an even number of block rows
          if (upper_only && bi > bj) continue;
          const bool second_ok = (bi + 1 < b1) && (!upper_only || bi + 1 <= bj);
          tiles.push_back(make_int2(bi, second_ok ? bj : (bj | umma::TILE_SECOND_INVALID)));
        }
      }
    }
  }
  tl.group_tile.push_back((int)tiles.size());
  tl.group_row.push_back(row_end);
  tl.ntiles = (int)tiles.size();
  if (tl.ntiles > 0) {
    CU_TRY(cudaMalloc(&tl.d_tiles, sizeof(int2) * tiles.size()));
    CU_TRY(cudaMemcpyAsync(tl.d_tiles, tiles.data(), sizeof(int2) * tiles.size(), cudaMemcpyHostToDevice, c->compute));
    CU_TRY(cudaStreamSynchronize(c->compute));
  }
  auto ins = c->tile_cache.emplace(key, std::move(tl));
  *out = &ins.first->second;
  return BIXCOR_OK;
}

// ----------------------------------------------------------------------------
// Gram operand preparation
// ----------------------------------------------------------------------------
struct Operand {
  int precision = BIXCOR_PREC_F64;
  int K = 0;  // padded contraction length
  int64_t p = 0;
  const double* z64 = nullptr;  // F64: p x K
  const int8_t* i8 = nullptr;   // I8EXACT: 2 planes (h, l) p x K
  const float* f32 = nullptr;   // TF32X3: 2 planes p x K (F16X3: the same slot holds 2 binary16 planes)
  const double* inv_norm = nullptr;
  const int8_t* oz = nullptr;        // OZAKI: OZ_SLICES digit planes per row, p x OZ_SLICES x ozK
  const double* oz_scale = nullptr;  // OZAKI: power-of-two row scales
  int ozK = 0;
};

static int round_up(int v, int m) { return (v + m - 1) / m * m; }
static bool is_ozaki(int precision) { return precision == BIXCOR_PREC_OZAKI || precision == BIXCOR_PREC_OZAKI4; }
static bool is_split_float(int precision) { return precision == BIXCOR_PREC_TF32X3 || precision == BIXCOR_PREC_F16X3; }

// Does this (precision, epilogue) run on the CTA-pair tcgen05 kernels (=> pair tile lists)?  g_umma_ctas: 1 = never,
// 3 = always, 2 (default) = where it measured faster on B200: the packed correlation (short K; -3 %).  With the
// chunked FP32 promotion of the long-K TOM / sparse Grams the cross-CTA drain handshake costs more than the
// halved B traffic saves (TOM 20k: 54 ms vs 43 ms), see profiles/umma_probe_r01.txt.
// Returns the tile geometry: 0 = one CTA per 128 x 128 tile, 1 = CTA pair (256 x 128), 2 = 2 x 2 multicast cluster of single
// CTAs (gram_umma.cuh, MC): the long-K A * A of the split-float TOM, which ncu shows bound by L2 -> SM operand traffic.
static int umma_pairs(int precision, int epi) {
  if (precision == BIXCOR_PREC_F64) return 0;
  if (epi == EPI_TOM && ((g_umma_wide == 2 && is_split_float(precision)) || (g_umma_wide == 1 && precision == BIXCOR_PREC_TF32X3)))
    return 3;  // 256 x 256 tiles per CTA pair
  if (g_umma_mc && epi == EPI_TOM && is_split_float(precision)) return 2;
  if (g_umma_ctas == 1) return 0;
  if (is_ozaki(precision)) return 1;  // long-K digit passes stream their operands from HBM / L2: half the B traffic pays
  if (precision == BIXCOR_PREC_F16X3) return 1;  // f16 kind: pairs measured faster on the long-K TOM as well (29.9 -> 28.6 ms)
  return (g_umma_ctas == 3 || epi == EPI_PACKED) ? 1 : 0;
}

static int k_padding(int precision) {
  if (precision == BIXCOR_PREC_F16X3) return 64;
  return (precision == BIXCOR_PREC_F64 || is_ozaki(precision)) ? dmma::BK : (precision == BIXCOR_PREC_I8EXACT ? 128 : 32);
}
// precisions whose operand is ONE record per gene that ranks can exchange (F64, 3xTF32, int8 exact, 3xF16)
static bool exchangeable(int precision) { return (precision >= 0 && precision <= 2) || precision == BIXCOR_PREC_F16X3; }

// Slice an f64 operand (rows x K, ld = ldz) into the Ozaki digit planes of ctx slot `slot`.
static int ozaki_slice(Ctx* c, const double* d_z, int64_t ldz, int64_t k, int64_t rows, int slot, Operand* op) {
  const int ozK = round_up((int)k, 128);
  const size_t plane_bytes = (size_t)OZ_SLICES * ozK * rows;
  RC_TRY(c->oz[slot].reserve(plane_bytes + 256 + sizeof(double) * (size_t)rows));
  int8_t* planes = (int8_t*)c->oz[slot].ptr;
  double* scale = (double*)((char*)c->oz[slot].ptr + ((plane_bytes + 255) & ~size_t(255)));
  {
    Timed tm(c, T_COLS);
    ozaki_slice_kernel<<<(unsigned)rows, 128, 0, c->compute>>>(d_z, ldz, (int)k, ozK, planes, scale);
    g_launches++;
  }
  CU_TRY(cudaGetLastError());
  op->oz = planes;
  op->oz_scale = scale;
  op->ozK = ozK;
  return BIXCOR_OK;
}

static int check_precision(int precision, int spearman, int64_t n) {
  if (precision == BIXCOR_PREC_F64 || is_split_float(precision) || is_ozaki(precision)) return BIXCOR_OK;
  if (precision == BIXCOR_PREC_I8EXACT) {
    if (!spearman) return fail(BIXCOR_ERR_INVALID, "BIXCOR_PREC_I8EXACT is defined for Spearman only");
    // sum d_i d_j <= n(n^2-1)/3 must fit the int32 recombination of the three digit accumulators
    if (n > 1860) return fail(BIXCOR_ERR_UNSUPPORTED, "BIXCOR_PREC_I8EXACT supports n <= 1860 samples (got %lld)", (long long)n);
    return BIXCOR_OK;
  }
  return fail(BIXCOR_ERR_INVALID, "unknown precision %d", precision);
}

// BIXCOR_PREC_AUTO (what the drop-in patch passes).  The reference's API has no precision argument
// (rs_cor_upper_triangle(x, spearman, shift), src/rust/src/base/r_cors_similarity.rs:251-268), so the library picks the
// fastest path that provably keeps the FP64 contract: exact-integer tcgen05 Gram for Spearman with n <= 1860, FP64
// DMMA otherwise; env BIXCOR_PRECISION replaces the rule process-wide (opt-in to the <= 1e-5 fast paths from R).
static int env_precision() {
  const char* v = getenv("BIXCOR_PRECISION");
  if (!v || !*v) return BIXCOR_PREC_AUTO;
  if (!strcmp(v, "f64")) return BIXCOR_PREC_F64;
  if (!strcmp(v, "tf32")) return BIXCOR_PREC_TF32X3;
  if (!strcmp(v, "f16")) return BIXCOR_PREC_F16X3;
  if (!strcmp(v, "i8")) return BIXCOR_PREC_I8EXACT;
  if (!strcmp(v, "ozaki")) return BIXCOR_PREC_OZAKI;
  if (!strcmp(v, "ozaki4")) return BIXCOR_PREC_OZAKI4;
  return BIXCOR_PREC_AUTO;
}
static int resolve_precision(int precision, int spearman, int64_t n) {
  if (precision != BIXCOR_PREC_AUTO) return precision;
  int want = env_precision();
  const bool i8_ok = spearman && n <= 1860;
  if (want == BIXCOR_PREC_I8EXACT && !i8_ok) want = BIXCOR_PREC_AUTO;
  if (want == BIXCOR_PREC_AUTO) return i8_ok ? BIXCOR_PREC_I8EXACT : BIXCOR_PREC_F64;
  return want;
}
// precision of an A * A product (TOM) or a sparse Gram: no integer-exact form exists for those
static int resolve_precision_float(int precision) {
  if (precision != BIXCOR_PREC_AUTO) return precision;
  const int want = env_precision();
  return (want == BIXCOR_PREC_AUTO || want == BIXCOR_PREC_I8EXACT) ? BIXCOR_PREC_F64 : want;
}
// precision of the correlation that feeds a TOM of precision `tom_precision` (as given by the caller)
static int tom_cor_precision(int tom_precision, int spearman, int64_t n) {
  if (tom_precision == BIXCOR_PREC_AUTO) return resolve_precision(BIXCOR_PREC_AUTO, spearman, n);
  return (is_split_float(tom_precision) || is_ozaki(tom_precision)) ? tom_precision : BIXCOR_PREC_F64;
}

// ColumnOut addressed from gene `g0` on (the column kernels index their outputs by the gene number inside the launch)
static ColumnOut column_out_from(ColumnOut co, int64_t g0) {
  if (co.f64) co.f64 += g0 * co.ld64;
  if (co.i8) co.i8 += g0 * co.ld8;
  if (co.f32) {
    if (co.half_planes) co.f32 = (float*)((__half*)co.f32 + g0 * co.ld32);
    else co.f32 += g0 * co.ld32;
  }
  if (co.inv_norm) co.inv_norm += g0;
  return co;
}

// Spearman for n > 16384 samples: chunked shared-memory sorts + rank by counting (columns.cuh), genes in batches so that
// the scratch (12 bytes per padded sample for the sorted chunks, 4 per sample for the ranks) stays below ~2 GB.
// ld = column stride of d_x in elements.
static int launch_rank_big(Ctx* c, const double* d_x, int64_t ld, int n, int64_t p, const ColumnOut& co) {
  if ((int64_t)n > (int64_t(1) << 21))
    return fail(BIXCOR_ERR_UNSUPPORTED, "Spearman rank kernel supports n <= 2097152 samples (got %d)", n);
  const int nchunks = (n + kBigChunk - 1) / kBigChunk;
  const size_t per_gene = (size_t)nchunks * kBigChunk * 12 + (size_t)n * 4 + 16;
  const int64_t batch = std::max<int64_t>(1, std::min<int64_t>(p, (int64_t)((size_t(2) << 30) / per_gene)));
  RC_TRY(c->rank_scratch.reserve(per_gene * (size_t)batch + 1024));
  char* base = (char*)c->rank_scratch.ptr;
  uint64_t* keys = (uint64_t*)base;
  uint32_t* idx = (uint32_t*)(base + (size_t)batch * nchunks * kBigChunk * 8);
  uint32_t* rk2 = idx + (size_t)batch * nchunks * kBigChunk;
  int* nan_flag = (int*)(rk2 + (size_t)batch * n);
  constexpr int smem = kBigChunk * 12;
  CU_TRY(cudaFuncSetAttribute(rank_big_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int64_t g0 = 0; g0 < p; g0 += batch) {
    const int64_t ng = std::min(batch, p - g0);
    CU_TRY(cudaMemsetAsync(nan_flag, 0, sizeof(int) * (size_t)ng, c->compute));
    const dim3 grid((unsigned)nchunks, (unsigned)ng);
    rank_big_sort_kernel<<<grid, kBigThreads, smem, c->compute>>>(d_x + g0 * ld, ld, n, nchunks, keys, idx, nan_flag);
    rank_big_count_kernel<<<grid, 256, 0, c->compute>>>(keys, idx, n, nchunks, rk2);
    rank_big_emit_kernel<<<(unsigned)ng, 256, 0, c->compute>>>(rk2, nan_flag, n, column_out_from(co, g0));
    g_launches += 3;
  }
  CU_TRY(cudaGetLastError());
  return BIXCOR_OK;
}

// Spearman column kernel dispatch: warp-per-gene register sort for n <= 1024, CTA shared-memory sort above.
static int launch_rank_kernel(Ctx* c, const double* d_x, int64_t n, int64_t p, const ColumnOut& co) {
  if (n <= 1024) {
    int e = 1;
    while (32 * e < n) e <<= 1;
    const unsigned grid = (unsigned)((p + 3) / 4);
    switch (e) {
      case 1: rank_columns_warp_kernel<1><<<grid, 128, 0, c->compute>>>(d_x, n, (int)n, p, co); break;
      case 2: rank_columns_warp_kernel<2><<<grid, 128, 0, c->compute>>>(d_x, n, (int)n, p, co); break;
      case 4: rank_columns_warp_kernel<4><<<grid, 128, 0, c->compute>>>(d_x, n, (int)n, p, co); break;
      case 8: rank_columns_warp_kernel<8><<<grid, 128, 0, c->compute>>>(d_x, n, (int)n, p, co); break;
      case 16: rank_columns_warp_kernel<16><<<grid, 128, 0, c->compute>>>(d_x, n, (int)n, p, co); break;
      default: rank_columns_warp_kernel<32><<<grid, 128, 0, c->compute>>>(d_x, n, (int)n, p, co); break;
    }
  } else if (n <= 16384) {
    int npow2 = 2048;
    while (npow2 < n) npow2 <<= 1;
    const size_t smem = (size_t)npow2 * 12;
    CU_TRY(cudaFuncSetAttribute(rank_columns_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rank_columns_kernel<<<(unsigned)p, 256, smem, c->compute>>>(d_x, n, (int)n, npow2, co);
  } else {
    return launch_rank_big(c, d_x, n, (int)n, p, co);
  }
  g_launches++;
  CU_TRY(cudaGetLastError());
  return BIXCOR_OK;
}

static int64_t operand_row_bytes(int64_t n, int precision) {
  const int64_t K = round_up((int)n, k_padding(precision));
  return precision == BIXCOR_PREC_I8EXACT ? 2 * K : (precision == BIXCOR_PREC_F16X3 ? 4 * K : 8 * K);
}

// Gene-major operand layout (one contiguous record per gene, so gene slabs built on different
// ranks can be all-gathered):  F64 [K f64] | I8EXACT [K int8 h][K int8 l] | TF32X3 [K f32 hi][K f32 lo].
static void bind_operand(Operand* op, int64_t n, int64_t p, int precision, const void* base, const double* inv) {
  op->precision = precision;
  op->K = round_up((int)n, k_padding(precision));
  op->p = p;
  op->z64 = nullptr;
  op->i8 = nullptr;
  op->f32 = nullptr;
  op->inv_norm = inv;
  if (precision == BIXCOR_PREC_F64 || is_ozaki(precision)) op->z64 = (const double*)base;
  else if (precision == BIXCOR_PREC_I8EXACT) op->i8 = (const int8_t*)base;
  else op->f32 = (const float*)base;
}

// d_x: device n x p column-major.  Writes the operand records of genes [g0, g1) into `dst`
// (base pointer of the full p-gene operand) and their 1/|d| into inv (I8EXACT).
static int build_operand(Ctx* c, const double* d_x, int64_t n, int64_t p, int spearman, int precision, int64_t g0,
                         int64_t g1, void* dst, double* inv, int norm_mode = 0) {
  if (n < 2) return fail(BIXCOR_ERR_INVALID, "need at least 2 samples (n = %lld)", (long long)n);
  if (n > (spearman ? (1 << 21) : (1 << 30)))
    return fail(BIXCOR_ERR_UNSUPPORTED, "supported sample counts: n <= 2097152 (Spearman), n <= 2^30 (Pearson); got %lld", (long long)n);
  RC_TRY(check_precision(precision, spearman, n));
  if (g0 < 0 || g1 > p || g0 > g1) return fail(BIXCOR_ERR_INVALID, "bad gene range");
  if (g0 == g1) return BIXCOR_OK;
  const int K = round_up((int)n, k_padding(precision));
  ColumnOut co;
  memset(&co, 0, sizeof co);
  co.kpad = K;
  if (precision == BIXCOR_PREC_F64 || is_ozaki(precision)) {
    co.mode = COL_Z64;
    co.f64 = (double*)dst + g0 * K;
    co.ld64 = K;
  } else if (precision == BIXCOR_PREC_I8EXACT) {
    co.mode = COL_I8;
    co.i8 = (int8_t*)dst + g0 * 2 * K;
    co.ld8 = 2 * K;
    co.plane8 = K;
    co.inv_norm = inv + g0;
  } else if (precision == BIXCOR_PREC_F16X3) {
    co.mode = COL_TF32;
    co.half_planes = 1;
    co.f32 = (float*)((__half*)dst + g0 * 2 * K);
    co.ld32 = 2 * K;
    co.plane32 = K;
  } else {
    co.mode = COL_TF32;
    co.f32 = (float*)dst + g0 * 2 * K;
    co.ld32 = 2 * K;
    co.plane32 = K;
  }
  const double* xs = d_x + g0 * n;
  const int64_t cnt = g1 - g0;
  Timed tm(c, T_COLS);
  if (spearman) return launch_rank_kernel(c, xs, n, cnt, co);
  // register-resident single-pass kernel when the column fits (n <= 2048, even, 16-byte aligned columns / records)
  const bool reg_ok = n <= 2048 && (n % 2 == 0) && ((uintptr_t)xs % 16 == 0) && (K % 2 == 0) && (K - n <= 256);
  if (reg_ok) standardise_columns_reg_kernel<<<(unsigned)cnt, 128, 0, c->compute>>>(xs, n, (int)n, co, norm_mode);
  else standardise_columns_kernel<<<(unsigned)cnt, 128, 0, c->compute>>>(xs, n, (int)n, co, norm_mode);
  g_launches++;
  CU_TRY(cudaGetLastError());
  return BIXCOR_OK;
}

// Builds the full operand in ctx slot `slot`.
static int prepare_operand(Ctx* c, const double* d_x, int64_t n, int64_t p, int spearman, int precision, int slot,
                           Operand* op, int norm_mode = 0) {
  RC_TRY(check_precision(precision, spearman, n));
  RC_TRY(c->op[slot].reserve((size_t)operand_row_bytes(n, precision) * p));
  RC_TRY(c->aux[slot].reserve(sizeof(double) * p));
  RC_TRY(build_operand(c, d_x, n, p, spearman, precision, 0, p, c->op[slot].ptr, (double*)c->aux[slot].ptr, norm_mode));
  bind_operand(op, n, p, precision, c->op[slot].ptr, (const double*)c->aux[slot].ptr);
  if (is_ozaki(precision)) RC_TRY(ozaki_slice(c, op->z64, op->K, n, p, slot, op));
  return BIXCOR_OK;
}

// ----------------------------------------------------------------------------
// Gram launch (dispatch over precision / epilogue)
// ----------------------------------------------------------------------------
template <int EPI>
static int launch_dmma(Ctx* c, const dmma::Params& P, int ntiles) {
  if (ntiles <= 0) return BIXCOR_OK;
  Timed tm(c, T_GEMM);
#define BIXCOR_DMMA_CASE(V)                                                                                  \
  case V: {                                                                                                  \
    auto kern = dmma::gram_dmma_kernel<EPI, V>;                                                              \
    CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dmma::smem_bytes_of<V>())); \
    kern<<<ntiles * (128 / dmma::Shape<V>::BN), dmma::threads_of<V>(), dmma::smem_bytes_of<V>(), c->compute>>>(P); \
  } break;
  switch (g_dmma_variant) {
    BIXCOR_DMMA_CASE(0)
    BIXCOR_DMMA_CASE(1)
    BIXCOR_DMMA_CASE(2)
    default:
      return fail(BIXCOR_ERR_INVALID, "bad DMMA kernel shape %d", g_dmma_variant);
  }
#undef BIXCOR_DMMA_CASE
  g_launches++;
  CU_TRY(cudaGetLastError());
  return BIXCOR_OK;
}

// per-context counters of the long-K producer barrier (nullptr when the allocation fails: the barrier is then off)
static unsigned* tile_sync_words(Ctx* c) {
  if (c->tile_sync.reserve(256) != BIXCOR_OK) return nullptr;
  return (unsigned*)c->tile_sync.ptr;
}

static EpiParams make_epi(int64_t p) {
  EpiParams e;
  memset(&e, 0, sizeof e);
  e.p = p;
  e.shift = 1;
  e.beta = 1.0;
  e.scale = 1.0;
  return e;
}

static void set_beta(EpiParams& e, double beta, int keep_sign) {
  e.beta = beta;
  e.keep_sign = keep_sign;
  e.beta_int = 0;
  if (beta > 1.0 && beta <= 64.0 && beta == std::floor(beta)) e.beta_int = (int)beta;
}

// Transposed interior tiles need the lean epilogue (identity or ^6, unsigned, no RBF) and a tcgen05 kind.
static int swap_tiles(int precision, const EpiParams& e) {
  if (!umma::kSwapInterior || (!is_split_float(precision) && precision != BIXCOR_PREC_I8EXACT)) return 0;
  return (!e.rbf_mode && !e.keep_sign && (e.beta == 1.0 || e.beta_int == 6)) ? 1 : 0;
}

// FP64-class Gram on the int8 tensor cores.  G (p x p f64, ctx->ozg) accumulates the digit-pair Grams of the
// six (group a, group b) combinations with s + t <= 5; each pass is the exact int8 kernel with EPI_ACCUM and the
// weight 128^-(2a + 2b + 4); a finisher applies the row scales and the requested epilogue for rows [row0,row1).
static int launch_gram_ozaki(Ctx* c, const Operand& a, int64_t p, const int2* tiles, int ntiles, int epi,
                             const EpiParams& e, int64_t row0, int64_t row1) {
  if (!a.oz || row0 < 0) return fail(BIXCOR_ERR_INVALID, "Ozaki path: operand not sliced / row range missing");
  if (epi != EPI_PACKED && epi != EPI_DENSE && epi != EPI_TOM)
    return fail(BIXCOR_ERR_UNSUPPORTED, "BIXCOR_PREC_OZAKI supports the correlation and TOM entry points");
  if ((int64_t)a.ozK > 60000) return fail(BIXCOR_ERR_UNSUPPORTED, "BIXCOR_PREC_OZAKI supports K <= 60000 (int32 digit sums)");
  RC_TRY(c->ozg.reserve(sizeof(double) * (size_t)p * p));
  double* G = (double*)c->ozg.ptr;
  CU_TRY(cudaMemsetAsync(G + row0 * p, 0, sizeof(double) * (size_t)(row1 - row0) * p, c->compute));
  // six slices: all digit products with s + t <= 5; four slices (BIXCOR_PREC_OZAKI4): s + t <= 3, the first three passes
  static const int kPass[6][2] = {{0, 0}, {0, 1}, {1, 0}, {1, 1}, {0, 2}, {2, 0}};
  const int npass = a.precision == BIXCOR_PREC_OZAKI4 ? 3 : 6;
  for (int ps = 0; ps < npass; ++ps) {
    const int* ab = kPass[ps];
    EpiParams ea = make_epi(p);
    ea.out = G;
    ea.ldo = p;
    ea.scale = std::ldexp(1.0, -7 * (2 * ab[0] + 2 * ab[1] + 4));
    Timed tm(c, T_GEMM);
    if (umma::launch_gram_umma(c->compute, c->sm_count, umma::KIND_I8, EPI_ACCUM, umma_pairs(a.precision, epi) == 1 ? 2 : 1, p, a.oz, a.ozK, nullptr, nullptr, 0,
                               nullptr, tiles, ntiles, ea, OZ_SLICES, 2 * ab[0], 2 * ab[1], tile_sync_words(c)) != 0)
      return fail(BIXCOR_ERR_UNSUPPORTED, "tcgen05 path: %s", umma::last_error());
    g_launches++;
  }
  {
    Timed tm(c, T_OTHER);
    const dim3 grid(8, (unsigned)(row1 - row0));
    if (epi == EPI_PACKED) ozaki_finish_packed_kernel<<<grid, 256, 0, c->compute>>>(G, a.oz_scale, e, row0, row1);
    else if (epi == EPI_DENSE) ozaki_finish_dense_kernel<<<grid, 256, 0, c->compute>>>(G, a.oz_scale, e, row0, row1);
    else ozaki_finish_tom_kernel<<<grid, 256, 0, c->compute>>>(G, a.oz_scale, e, row0, row1);
    g_launches++;
  }
  CU_TRY(cudaGetLastError());
  return BIXCOR_OK;
}

static int launch_gram(Ctx* c, const Operand& a, const Operand* b, int64_t p, const int2* tiles, int ntiles, int epi,
                       const EpiParams& e, const Operand* cols = nullptr, int64_t row0 = -1, int64_t row1 = -1) {
  if (ntiles <= 0) return BIXCOR_OK;
  if (is_ozaki(a.precision)) {
    if (b || cols) return fail(BIXCOR_ERR_UNSUPPORTED, "BIXCOR_PREC_OZAKI supports the correlation and TOM entry points");
    return launch_gram_ozaki(c, a, p, tiles, ntiles, epi, e, row0, row1);
  }
  if (cols && a.precision != BIXCOR_PREC_F64)
    return fail(BIXCOR_ERR_UNSUPPORTED, "two-matrix correlation runs on the FP64 path only");
  if (a.precision == BIXCOR_PREC_F64) {
    dmma::Params P;
    memset(&P, 0, sizeof P);
    P.R0 = P.C0 = a.z64;
    P.pc = (int)p;
    if (cols) {
      P.C0 = cols->z64;
      P.pc = (int)cols->p;
    }
    P.ld0 = a.K;
    P.K0 = a.K;
    if (b) {
      P.R1 = P.C1 = b->z64;
      P.ld1 = b->K;
      P.K1 = b->K;
    }
    P.p = (int)p;
    P.tiles = tiles;
    P.E = e;
    switch (epi) {
      case EPI_PACKED: return launch_dmma<EPI_PACKED>(c, P, ntiles);
      case EPI_DENSE: return launch_dmma<EPI_DENSE>(c, P, ntiles);
      case EPI_DIFFCOR: return launch_dmma<EPI_DIFFCOR>(c, P, ntiles);
      case EPI_TOM: return launch_dmma<EPI_TOM>(c, P, ntiles);
      case EPI_ACCUM: return launch_dmma<EPI_ACCUM>(c, P, ntiles);
    }
    return fail(BIXCOR_ERR_INVALID, "bad epilogue %d", epi);
  }
  const int kind = a.precision == BIXCOR_PREC_I8EXACT ? umma::KIND_I8 : (a.precision == BIXCOR_PREC_F16X3 ? umma::KIND_F16 : umma::KIND_TF32);
  const void* pl0 = kind == umma::KIND_I8 ? (const void*)a.i8 : (const void*)a.f32;
  const void* pl1 = b ? (kind == umma::KIND_I8 ? (const void*)b->i8 : (const void*)b->f32) : nullptr;
  Timed tm(c, T_GEMM);
  const int geom = umma_pairs(a.precision, epi);
  const int nbk = (int)((p + kTile - 1) / kTile);
  const int rb0 = row0 >= 0 ? (int)(row0 / kTile) : 0, rb1 = row1 >= 0 ? (int)((row1 + kTile - 1) / kTile) : nbk;
  if (umma::launch_gram_umma(c->compute, c->sm_count, kind, epi, (geom == 1 || geom == 3) ? 2 : 1, p, pl0, a.K, a.inv_norm, pl1, b ? b->K : 0,
                             b ? b->inv_norm : nullptr, tiles, ntiles, e, 2, 0, 0, tile_sync_words(c), geom == 2 ? 1 : (geom == 3 ? 2 : 0),
                             rb0, rb1, e.upper_tiles) != 0)
    return fail(BIXCOR_ERR_UNSUPPORTED, "tcgen05 path: %s", umma::last_error());
  g_launches++;
  return BIXCOR_OK;
}

// ----------------------------------------------------------------------------
// host <-> device staging
// ----------------------------------------------------------------------------
static bool is_pinned(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost;
}

// ----------------------------------------------------------------------------
// Pageable host buffers (what an R session owns) cannot be DMA'd directly, pinning them in place costs more than
// the transfer (cudaHostRegister of 1.6 GB: 224 ms against 29 ms of PCIe, scripts/micro/d2h_floor.cu), and the
// driver's own staging reaches 17 GB/s.  The library stages through a ring of pinned slots and keeps the two
// halves of the job on different threads: the calling thread only enqueues DMAs, a persistent pool of copier threads
// moves finished slots into (out of) the caller's memory in 4 MB pieces (16 threads: 75 GB/s, above PCIe's 55 GB/s).
// ----------------------------------------------------------------------------
// Streaming copy for the copier threads: non-temporal stores do not read the destination lines first (a plain store
// to memory that is not in cache costs a read-for-ownership, i.e. 3 bytes of DRAM traffic per byte copied instead of 2),
// and glibc only switches to them far above the 4 MB pieces used here.
static void stream_copy(char* dst, const char* src, size_t len) {
  if (len < 4096) {
    memcpy(dst, src, len);
    return;
  }
  const size_t head = (16 - ((uintptr_t)dst & 15)) & 15;
  if (head) {
    memcpy(dst, src, head);
    dst += head;
    src += head;
    len -= head;
  }
  const size_t n = len / 64;
  for (size_t i = 0; i < n; ++i) {
    const __m128i a = _mm_loadu_si128((const __m128i*)(src + 64 * i));
    const __m128i b = _mm_loadu_si128((const __m128i*)(src + 64 * i + 16));
    const __m128i c2 = _mm_loadu_si128((const __m128i*)(src + 64 * i + 32));
    const __m128i d = _mm_loadu_si128((const __m128i*)(src + 64 * i + 48));
    _mm_stream_si128((__m128i*)(dst + 64 * i), a);
    _mm_stream_si128((__m128i*)(dst + 64 * i + 16), b);
    _mm_stream_si128((__m128i*)(dst + 64 * i + 32), c2);
    _mm_stream_si128((__m128i*)(dst + 64 * i + 48), d);
  }
  _mm_sfence();
  if (len > 64 * n) memcpy(dst + 64 * n, src + 64 * n, len - 64 * n);
}

// int32 transport of the exact Spearman Gram (EpiParams::out_s32): the Gram kernel leaves the exact integers S_ij in
// the packed layout (4 bytes per pair), half the bytes cross PCIe, and the copier threads - which touch every element
// anyway - finish r_ij = S_ij / (|d_i| |d_j|) (and |r|^6) while they move the piece into the caller's pageable buffer.
// The arithmetic is the device epilogue's, operation for operation (int -> double, times 1/|d_i|, times 1/|d_j|, then
// sq = v v, sq sq sq): IEEE multiplications only, so host and device results are bit-identical.
struct ExpandS32 {
  const double* inv = nullptr;  // host copy of 1/|d| per gene
  int64_t p = 0;
  int shift = 1;
  int pow6 = 0;
};

static inline int64_t packed_row_of(int64_t g, int64_t p, int shift) {  // largest i with packed_row_offset(i) <= g
  const double a = shift ? 2.0 * (double)p - 1.0 : 2.0 * (double)p + 1.0;
  int64_t i = (int64_t)((a - std::sqrt(std::max(0.0, a * a - 8.0 * (double)g))) / 2.0);
  i = std::max<int64_t>(0, std::min<int64_t>(p - 1, i));
  while (i > 0 && packed_row_offset(i, p, shift) > g) --i;
  while (i + 1 < p && packed_row_offset(i + 1, p, shift) <= g) ++i;
  return i;
}

// one run of a packed row (host_simd.cpp): AVX2 when the host has it, scalar otherwise - bit-identical forms
extern "C" __attribute__((visibility("hidden"))) int bixcor_host_has_avx2(void);
extern "C" __attribute__((visibility("hidden"))) void bixcor_expand_run_scalar(double* dst, const int32_t* src, int64_t n, double rf, const double* cf, int pow6);
extern "C" __attribute__((visibility("hidden"))) void bixcor_expand_run_avx2(double* dst, const int32_t* src, int64_t n, double rf, const double* cf, int pow6);
extern "C" __attribute__((visibility("hidden"))) void bixcor_widen_f32_scalar(double* dst, const float* src, int64_t n);
extern "C" __attribute__((visibility("hidden"))) void bixcor_widen_f32_avx2(double* dst, const float* src, int64_t n);
// binary32 transport of the <= 1e-5-class results: the copier threads widen float -> double into the caller's buffer
static inline void widen_f32(double* dst, const float* src, int64_t n) {
  static const bool have_avx2 = bixcor_host_has_avx2() != 0;
  if (have_avx2) bixcor_widen_f32_avx2(dst, src, n);
  else bixcor_widen_f32_scalar(dst, src, n);
  _mm_sfence();
}
static inline void expand_run(double* dst, const int32_t* src, int64_t n, double rf, const double* cf, int pow6) {
  static const bool have_avx2 = bixcor_host_has_avx2() != 0;
  if (have_avx2) bixcor_expand_run_avx2(dst, src, n, rf, cf, pow6);
  else bixcor_expand_run_scalar(dst, src, n, rf, cf, pow6);
}

// dst[t] for the packed elements [g0, g0 + count) from their integer Gram values
static void expand_packed_s32(double* dst, const int32_t* src, int64_t count, int64_t g0, const ExpandS32& X) {
  int64_t i = packed_row_of(g0, X.p, X.shift), pos = g0;
  while (count > 0) {
    const int64_t row_start = packed_row_offset(i, X.p, X.shift), row_len = X.p - i - X.shift;
    if (pos >= row_start + row_len) {  // (rows without stored elements)
      ++i;
      continue;
    }
    const int64_t j0 = i + X.shift + (pos - row_start), nrun = std::min(count, row_start + row_len - pos);
    expand_run(dst, src, nrun, X.inv[i], X.inv + j0, X.pow6);
    dst += nrun;
    src += nrun;
    pos += nrun;
    count -= nrun;
    ++i;
  }
  _mm_sfence();
}

// Geometry measured on the B200 box (scripts/e2e_sweep.py, profiles/e2e_sweep_r02.json; 1.6 GB result, 16 host threads):
// 8 x 32 MB slots / 4 MB pieces / blocking events 43.5 ms; 16 x 4 MB / 512 KB / spinning 35.7 ms (pinned buffers: 32 ms).
// A 64 MB ring stays in the last-level cache the DMA engine writes into, so the copier threads read the data from
// cache instead of DRAM; small slots need low-latency wake-ups, hence the spin on the DMA events.
static size_t g_stage_slot_bytes = size_t(4) << 20;   // env BIXCOR_STAGE_SLOT_MB
static int g_stage_slots = 16;                         // env BIXCOR_STAGE_SLOTS (<= kMaxSlots)
static size_t g_stage_piece = size_t(512) << 10;       // env BIXCOR_STAGE_PIECE_KB
static int g_stage_nt = 1;                             // env BIXCOR_STAGE_NT: streaming stores into the caller's buffer
static int g_stage_spin = 1;                           // env BIXCOR_STAGE_SPIN: copier threads spin on the DMA events instead of blocking
// The int32 transport has its own ring geometry: its copier threads write two bytes for every byte they read, the DMA of a
// slot is short (1 MB: 19 us) against its expansion, and a finer grain keeps all threads busy (scripts/e2e_sweep.py --round2b /
// --round2c, profiles/e2e_sweep_s32_r02.json: 32 x 1 MB / 256 KB pieces 22-23 ms, 16 x 4 MB / 512 KB 29-31 ms at config 2;
// below 1 MB the cudaMemcpyAsync + event-record calls of the launching thread dominate).  The same pinned slots are used.
static size_t g_s32_slot_bytes = size_t(1) << 20;      // env BIXCOR_S32_SLOT_KB
static int g_s32_slots = 32;                           // env BIXCOR_S32_SLOTS
static size_t g_s32_piece = size_t(256) << 10;         // env BIXCOR_S32_PIECE_KB
static int g_s32_transport = 1;                        // env BIXCOR_S32_TRANSPORT: int32 transport of the exact Spearman Gram to pageable hosts
static int g_f32_transport = 1;                        // env BIXCOR_F32_TRANSPORT: binary32 transport of the split-float (<= 1e-5 class) correlation / TOM results

// How a copier thread learns that the DMA of a slot has landed: by default it waits on the slot's CUDA event.  The
// alternative kept here (env BIXCOR_STAGE_SEQ=1) keeps the copier threads out of the driver altogether: the copy stream writes
// a sequence number into one pinned host word after every device-to-host DMA (cuStreamWriteValue32, stream ordered) and the
// threads spin on that word with plain loads.  Measured at config 2 (scripts/e2e_sweep.py --round2e, three alternating
// repetitions, profiles/e2e_sweep_s32_r02.json): 26.8-28.9 ms against 23.3-23.9 ms with events - a stream memory operation
// between two copies costs the copy engine more than an event record does - so it stays off.
typedef CUresult (*StreamWriteValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
static int g_stage_seq = 0;
static StreamWriteValue32Fn stream_write_value_fn() {
  static StreamWriteValue32Fn fn = nullptr;
  static bool tried = false;
  if (tried) return fn;
  tried = true;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
    fn = (StreamWriteValue32Fn)p;
  else cudaGetLastError();
  return fn;
}

struct HostStager {
  static constexpr int kMaxSlots = 64;
  size_t kSlotBytes = g_stage_slot_bytes;
  size_t kPiece = g_stage_piece;
  int kSlots = g_stage_slots;   // slots allocated
  int kCycle = g_stage_slots;   // slots a transfer cycles through unless it says otherwise (the int32 transport: g_s32_slots)
  struct Slot {
    void* ptr = nullptr;
    cudaEvent_t dma_done = nullptr;  // D2H: data has landed in the slot; H2D: the DMA has read the slot
    int pieces_left = 0;             // host copies still running on this slot
    bool dma_pending = false;        // H2D: dma_done must be waited for before the slot is refilled
    uint32_t seq = 0;                // D2H: sequence number the copy stream writes to *seq_host once this slot's DMA has landed
  };
  struct Task {
    int slot;
    char* dst;
    const char* src;
    size_t len;
    bool wait_dma;  // D2H piece: wait for the slot's DMA first
    const ExpandS32* ex = nullptr;  // D2H piece of an int32 Gram: src holds len / 4 integers, dst receives as many doubles
    int64_t g0 = 0;                 // packed index of the piece's first element
    bool widen = false;             // D2H piece of a binary32 result: src holds len / 4 floats, dst receives as many doubles
    uint32_t wait_seq = 0;          // D2H piece, sequence mode: the slot has landed once *seq_host has reached this number
  };
  volatile uint32_t* seq_host = nullptr;  // pinned word written by the copy stream after every D2H DMA (nullptr: event mode)
  CUdeviceptr seq_dev = 0;
  uint32_t seq_issued = 0;
  StreamWriteValue32Fn write_fn = nullptr;
  int device = 0;
  Slot slots[kMaxSlots];
  int next_slot = 0;
  std::mutex m;
  std::condition_variable cv_task, cv_slot;
  std::vector<Task> queue;  // LIFO is fine: pieces are independent
  size_t queue_head = 0;
  bool stop = false;
  int error = 0;  // first cudaError_t seen by a worker
  std::vector<std::thread> workers;

  void worker() {
    cudaSetDevice(device);
    for (;;) {
      Task t;
      {
        std::unique_lock<std::mutex> lk(m);
        cv_task.wait(lk, [&] { return stop || queue_head < queue.size(); });
        if (queue_head >= queue.size()) return;  // stop requested and nothing left
        t = queue[queue_head++];
        if (queue_head == queue.size()) {
          queue.clear();
          queue_head = 0;
        }
      }
      cudaError_t e = cudaSuccess;
      if (t.wait_dma && seq_host) {
        uint64_t spins = 0;
        const auto t_start = std::chrono::steady_clock::now();
        while ((int32_t)(*seq_host - t.wait_seq) < 0) {
          _mm_pause();
          if ((++spins & 0xFFFFFu) == 0 && std::chrono::steady_clock::now() - t_start > std::chrono::seconds(120)) {
            e = cudaErrorLaunchTimeout;  // (the stream died: the launching thread reports its error, nobody hangs)
            break;
          }
        }
        std::atomic_thread_fence(std::memory_order_acquire);
      } else if (t.wait_dma) e = cudaEventSynchronize(slots[t.slot].dma_done);
      if (e == cudaSuccess) {
        if (t.ex) expand_packed_s32((double*)t.dst, (const int32_t*)t.src, (int64_t)(t.len / 4), t.g0, *t.ex);
        else if (t.widen) widen_f32((double*)t.dst, (const float*)t.src, (int64_t)(t.len / 4));
        else if (t.wait_dma && g_stage_nt) stream_copy(t.dst, t.src, t.len);  // into the caller's (cold) buffer
        else memcpy(t.dst, t.src, t.len);
      }
      {
        std::lock_guard<std::mutex> lk(m);
        if (e != cudaSuccess && !error) error = (int)e;
        if (--slots[t.slot].pieces_left == 0) cv_slot.notify_all();
      }
    }
  }
  // the calling thread, right after it enqueued the D2H DMA into slot sidx on stream s: how the copier threads will see it land
  int mark_landing(int sidx, cudaStream_t s) {
    if (seq_host) {
      ++seq_issued;
      const CUresult r = write_fn((CUstream)s, seq_dev, seq_issued, 0);
      if (r != CUDA_SUCCESS) return fail(BIXCOR_ERR_CUDA, "cuStreamWriteValue32 failed with CUresult %d", (int)r);
      slots[sidx].seq = seq_issued;
      return BIXCOR_OK;
    }
    CU_TRY(cudaEventRecord(slots[sidx].dma_done, s));
    return BIXCOR_OK;
  }
  // the calling thread: next slot of the ring, once its previous host copies (and H2D DMA) are done
  // nslots: how many of the ring's slots this transfer cycles through (0 = all of them)
  int acquire(int* out_slot, int nslots = 0) {
    if (nslots <= 0 || nslots > kSlots) nslots = kCycle;
    const int sidx = next_slot % nslots;
    next_slot = (sidx + 1) % nslots;
    {
      std::unique_lock<std::mutex> lk(m);
      cv_slot.wait(lk, [&] { return slots[sidx].pieces_left == 0; });
      if (error) return fail(BIXCOR_ERR_CUDA, "staging copy: %s", cudaGetErrorString((cudaError_t)error));
    }
    if (slots[sidx].dma_pending) {
      CU_TRY(cudaEventSynchronize(slots[sidx].dma_done));
      slots[sidx].dma_pending = false;
    }
    *out_slot = sidx;
    return BIXCOR_OK;
  }
  // split [0, len) of a slot into pieces for the pool; host_ptr is the caller's memory
  // ex != nullptr (to_host only): the slot holds int32 Gram values, host_ptr is the double destination of its first one
  // widen (to_host only): the slot holds floats, host_ptr is the double destination of its first one
  void post(int sidx, char* host_ptr, size_t len, bool to_host, const ExpandS32* ex = nullptr, int64_t g0 = 0, bool widen = false,
            size_t piece = 0) {
    if (piece == 0) piece = kPiece;
    std::lock_guard<std::mutex> lk(m);
    for (size_t o = 0; o < len; o += piece) {
      const size_t l = std::min(piece, len - o);
      Task t;
      t.slot = sidx;
      t.len = l;
      t.wait_dma = to_host;
      t.ex = ex;
      t.widen = widen;
      t.wait_seq = slots[sidx].seq;
      t.g0 = g0 + (int64_t)(o / 4);
      t.dst = to_host ? host_ptr + ((ex || widen) ? 2 * o : o) : (char*)slots[sidx].ptr + o;
      t.src = to_host ? (const char*)slots[sidx].ptr + o : host_ptr + o;
      queue.push_back(t);
      slots[sidx].pieces_left++;
    }
    cv_task.notify_all();
  }
  int wait_slot(int sidx) {
    std::unique_lock<std::mutex> lk(m);
    cv_slot.wait(lk, [&] { return slots[sidx].pieces_left == 0; });
    if (error) return fail(BIXCOR_ERR_CUDA, "staging copy: %s", cudaGetErrorString((cudaError_t)error));
    return BIXCOR_OK;
  }
  int drain() {
    for (int i = 0; i < kSlots; ++i) RC_TRY(wait_slot(i));
    return BIXCOR_OK;
  }
};

static int g_copy_threads = 0;  // env BIXCOR_COPY_THREADS; 0 = from the host's core count

static void stager_free(HostStager* st);

static int stager_get(Ctx* c, HostStager** out, bool inbound = false) {
  HostStager*& slot_ref = inbound ? c->stager_in : c->stager;
  if (slot_ref) {
    *out = slot_ref;
    return BIXCOR_OK;
  }
  HostStager* st = new HostStager();
  st->device = c->device;
  if (inbound) st->kSlots = st->kCycle = std::min(st->kSlots, 6);
  else st->kSlots = std::max(st->kSlots, std::min((int)HostStager::kMaxSlots, g_s32_slots));  // (the int32 transport cycles through more, smaller slots)
  st->kSlotBytes = std::max(st->kSlotBytes, inbound ? size_t(0) : g_s32_slot_bytes);
  for (int si = 0; si < st->kSlots; ++si) {
    auto& sl = st->slots[si];
    // slots beyond the default cycle only ever serve the int32 transport's (smaller) slot size
    const size_t cap = (inbound || si < st->kCycle) ? st->kSlotBytes : std::min(st->kSlotBytes, g_s32_slot_bytes);
    if (cudaMallocHost(&sl.ptr, cap) != cudaSuccess ||
        cudaEventCreateWithFlags(&sl.dma_done, cudaEventDisableTiming | (g_stage_spin ? 0 : cudaEventBlockingSync)) != cudaSuccess) {
      cudaGetLastError();
      stager_free(st);
      return fail(BIXCOR_ERR_OOM, "cannot allocate the pinned staging ring (%d x %zu MB)", g_stage_slots, g_stage_slot_bytes >> 20);
    }
  }
  if (!inbound && g_stage_seq && g_stage_spin && stream_write_value_fn()) {
    void* hp = nullptr;
    void* dp = nullptr;
    if (cudaHostAlloc(&hp, 64, cudaHostAllocDefault) == cudaSuccess && cudaHostGetDevicePointer(&dp, hp, 0) == cudaSuccess) {
      memset(hp, 0, 64);
      st->seq_host = (volatile uint32_t*)hp;
      st->seq_dev = (CUdeviceptr)(uintptr_t)dp;
      st->write_fn = stream_write_value_fn();
    } else {
      cudaGetLastError();
      if (hp) cudaFreeHost(hp);
    }
  }
  unsigned hw = std::thread::hardware_concurrency();
  if (hw == 0) hw = 4;
  int nd = std::max<int>(1, (int)g_ctx.size());
  if (const char* lw = getenv("LOCAL_WORLD_SIZE")) nd = std::max(nd, atoi(lw));  // one process per GPU (torchrun): share the cores
  int nthr = g_copy_threads > 0 ? g_copy_threads : (int)std::max(2u, std::min(14u, (hw > 2 ? hw - 2 : 1u) / (unsigned)nd));
  if (inbound) nthr = std::max(2, std::min(4, nthr / 3));  // inputs are a tenth of the traffic
  slot_ref = st;
  try {
    for (int t = 0; t < nthr; ++t) st->workers.emplace_back([st] {
      try {
        st->worker();
      } catch (...) {  // (mutex / condition variable failure: the pool stops, callers see the error flag)
        std::lock_guard<std::mutex> lk(st->m);
        if (!st->error) st->error = (int)cudaErrorUnknown;
        st->cv_slot.notify_all();
      }
    });
  } catch (...) {
    if (st->workers.empty()) {
      stager_free(st);
      slot_ref = nullptr;
      throw;
    }
  }
  *out = st;
  return BIXCOR_OK;
}

static void stager_destroy(Ctx* c) {
  stager_free(c->stager);
  stager_free(c->stager_in);
  c->stager = c->stager_in = nullptr;
}

static void stager_free(HostStager* st) {
  if (!st) return;
  {
    std::lock_guard<std::mutex> lk(st->m);
    st->stop = true;
    st->cv_task.notify_all();
  }
  for (auto& t : st->workers)
    if (t.joinable()) t.join();
  for (auto& sl : st->slots) {
    if (sl.dma_done) cudaEventDestroy(sl.dma_done);
    if (sl.ptr) cudaFreeHost(sl.ptr);
  }
  if (st->seq_host) cudaFreeHost((void*)st->seq_host);
  delete st;
}

// Host -> device.  Pinned sources are DMA'd directly; pageable ones go through the pinned ring: the pool copies
// the caller's memory into slot k + 1 while slot k is on the wire.
static int h2d(Ctx* c, void* d_dst, const void* h_src, size_t bytes, cudaStream_t s) {
  if (bytes == 0) return BIXCOR_OK;
  g_h2d_bytes += (int64_t)bytes;
  if (is_pinned(h_src)) {
    CU_TRY(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, s));
    return BIXCOR_OK;
  }
  HostStager* st;
  RC_TRY(stager_get(c, &st, true));
  int prev = -1;
  size_t prev_off = 0, prev_len = 0;
  auto issue = [&](int sidx, size_t off, size_t len) -> int {  // host copies of the slot are done: put it on the wire
    RC_TRY(st->wait_slot(sidx));
    CU_TRY(cudaMemcpyAsync((char*)d_dst + off, st->slots[sidx].ptr, len, cudaMemcpyHostToDevice, s));
    CU_TRY(cudaEventRecord(st->slots[sidx].dma_done, s));
    st->slots[sidx].dma_pending = true;
    return BIXCOR_OK;
  };
  for (size_t o = 0; o < bytes; o += st->kSlotBytes) {
    const size_t len = std::min(st->kSlotBytes, bytes - o);
    int sidx = 0;
    RC_TRY(st->acquire(&sidx));
    st->post(sidx, (char*)const_cast<void*>(h_src) + o, len, false);
    if (prev >= 0) RC_TRY(issue(prev, prev_off, prev_len));
    prev = sidx;
    prev_off = o;
    prev_len = len;
  }
  if (prev >= 0) RC_TRY(issue(prev, prev_off, prev_len));
  CU_TRY(cudaStreamSynchronize(s));  // the caller's buffer may be reused, the ring is free again
  for (auto& sl : st->slots) sl.dma_pending = false;
  return BIXCOR_OK;
}

// Device -> host of finished ranges; `ready` (may be null) gates the copy.  Pinned destinations: plain async DMA.
// Pageable destinations: DMA into the pinned ring on the copy stream, the pool moves every landed slot into the
// caller's memory; the calling thread never copies and only blocks when all slots are in flight.
struct D2HPipe {
  Ctx* c;
  HostStager* st = nullptr;
  explicit D2HPipe(Ctx* c_, const void* = nullptr) : c(c_) {}
  int push(void* h_dst, const void* d_src, size_t bytes, cudaEvent_t ready) {
    if (bytes == 0) return BIXCOR_OK;
    g_d2h_bytes += (int64_t)bytes;
    if (ready) CU_TRY(cudaStreamWaitEvent(c->copy_out, ready, 0));
    if (is_pinned(h_dst)) {
      CU_TRY(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, c->copy_out));
      return BIXCOR_OK;
    }
    if (!st) RC_TRY(stager_get(c, &st));
    for (size_t o = 0; o < bytes; o += st->kSlotBytes) {
      const size_t len = std::min(st->kSlotBytes, bytes - o);
      int sidx = 0;
      RC_TRY(st->acquire(&sidx));
      CU_TRY(cudaMemcpyAsync(st->slots[sidx].ptr, (const char*)d_src + o, len, cudaMemcpyDeviceToHost, c->copy_out));
      RC_TRY(st->mark_landing(sidx, c->copy_out));
      st->post(sidx, (char*)h_dst + o, len, true);
    }
    return BIXCOR_OK;
  }
  // int32 transport: `count` exact Gram integers at d_src become doubles at h_dst (pageable), g0 = packed index of the first
  ExpandS32 ex;
  int push_s32(double* h_dst, const int32_t* d_src, int64_t count, int64_t g0, cudaEvent_t ready) {
    if (count <= 0) return BIXCOR_OK;
    if (ready) CU_TRY(cudaStreamWaitEvent(c->copy_out, ready, 0));
    if (!st) RC_TRY(stager_get(c, &st));
    const size_t bytes = sizeof(int32_t) * (size_t)count;
    g_d2h_bytes += (int64_t)bytes;
    const size_t slot_bytes = std::min(st->kSlotBytes, g_s32_slot_bytes);
    for (size_t o = 0; o < bytes; o += slot_bytes) {
      const size_t len = std::min(slot_bytes, bytes - o);
      int sidx = 0;
      RC_TRY(st->acquire(&sidx, g_s32_slots));
      CU_TRY(cudaMemcpyAsync(st->slots[sidx].ptr, (const char*)d_src + o, len, cudaMemcpyDeviceToHost, c->copy_out));
      RC_TRY(st->mark_landing(sidx, c->copy_out));
      st->post(sidx, (char*)(h_dst + o / 4), len, true, &ex, g0 + (int64_t)(o / 4), false, g_s32_piece);
    }
    return BIXCOR_OK;
  }
  // binary32 transport: `count` floats at d_src become doubles at h_dst (pageable or page-locked: the widening needs a host pass)
  int push_f32(double* h_dst, const float* d_src, int64_t count, cudaEvent_t ready) {
    if (count <= 0) return BIXCOR_OK;
    if (ready) CU_TRY(cudaStreamWaitEvent(c->copy_out, ready, 0));
    if (!st) RC_TRY(stager_get(c, &st));
    const size_t bytes = sizeof(float) * (size_t)count;
    g_d2h_bytes += (int64_t)bytes;
    for (size_t o = 0; o < bytes; o += st->kSlotBytes) {
      const size_t len = std::min(st->kSlotBytes, bytes - o);
      int sidx = 0;
      RC_TRY(st->acquire(&sidx));
      CU_TRY(cudaMemcpyAsync(st->slots[sidx].ptr, (const char*)d_src + o, len, cudaMemcpyDeviceToHost, c->copy_out));
      RC_TRY(st->mark_landing(sidx, c->copy_out));
      st->post(sidx, (char*)(h_dst + o / 4), len, true, nullptr, 0, true);
    }
    return BIXCOR_OK;
  }
  int finish() {
    if (st) RC_TRY(st->drain());
    CU_TRY(cudaStreamSynchronize(c->copy_out));
    return BIXCOR_OK;
  }
};

// Binary32 transport of a finished band of a split-float (<= 1e-5 class: 3xTF32 / 3xF16) result.  Those values carry ~22
// significant bits, so rounding them to float before they cross PCIe changes nothing inside their accuracy class
// (relative 6e-8) and halves the device-to-host bytes - the floor of every host -> host call.  The band [d_src, +count)
// of the f64 result is narrowed into the same element range of c->out32 on the compute stream (a streaming pass of 12
// bytes per element: ~0.4 ms for the 1.6 GB of a 20 000-gene result) and the copier threads widen it into h_dst.
// d_base: element 0 of the result buffer the band belongs to (out32 mirrors its indexing).
static int push_band_f32(Ctx* c, D2HPipe* pipe, double* h_dst, const double* d_base, const double* d_src, int64_t count) {
  if (count <= 0) return BIXCOR_OK;
  float* d_n = (float*)c->out32.ptr + (d_src - d_base);
  {
    Timed tm(c, T_OTHER);
    const unsigned blocks = (unsigned)std::min<int64_t>((count + 1023) / 1024, (int64_t)c->sm_count * 16);
    narrow_f64_f32_kernel<<<blocks, 256, 0, c->compute>>>(d_src, d_n, count);
    g_launches++;
  }
  CU_TRY(cudaGetLastError());
  cudaEvent_t ev = ctx_event(c);
  CU_TRY(cudaEventRecord(ev, c->compute));
  return pipe->push_f32(h_dst, d_n, count, ev);
}
static bool f32_transport(int precision, const EpiParams& e) {
  return g_f32_transport && is_split_float(precision) && e.scale == 1.0;  // (correlations / affinities / TOM: magnitudes <= 1)
}

// ----------------------------------------------------------------------------
// device-level building blocks (all pointers device, ctx current)
// ----------------------------------------------------------------------------
static int validate_rows(int64_t p, int64_t row_begin, int64_t row_end) {
  if (p < 1) return fail(BIXCOR_ERR_INVALID, "p must be >= 1");
  if (row_begin < 0 || row_end > p || row_begin > row_end)
    return fail(BIXCOR_ERR_INVALID, "bad row range [%lld,%lld) for p = %lld", (long long)row_begin, (long long)row_end,
                (long long)p);
  if (row_begin == row_end) return BIXCOR_OK;  // an empty shard (small p, many ranks) is a no-op wherever it starts
  if (row_begin % kTile != 0)
    return fail(BIXCOR_ERR_INVALID, "row_begin (%lld) must be a multiple of %d", (long long)row_begin, kTile);
  // tiles are 128-row blocks and the epilogues guard rows with i < p only: a ragged row_end inside the matrix would
  // write rows [row_end, next block boundary) past the caller's slice
  if (row_end % kTile != 0 && row_end != p)
    return fail(BIXCOR_ERR_INVALID, "row_end (%lld) must be a multiple of %d or p (%lld)", (long long)row_end, kTile,
                (long long)p);
  return BIXCOR_OK;
}

// Packed correlation of rows [row_begin,row_end).  If `pipe` is given, every
// finished band is copied to h_out (host slice) while later bands compute.
// optional RBF transform of the correlation in the epilogue
struct RbfSpec {
  int mode = 0;
  double eps = 1.0;
  int from_cor = 1;
  int complement = 0;
};

static int dev_cor_upper_rows_op(Ctx* c, const Operand& op, int64_t p, int shift, double beta, int64_t row_begin,
                                 int64_t row_end, double* d_out, D2HPipe* pipe, double* h_out,
                                 const RbfSpec* rbf = nullptr) {
  EpiParams e = make_epi(p);
  e.out = d_out;
  e.out_base = packed_row_offset(row_begin, p, shift);
  e.shift = shift;
  set_beta(e, beta, 0);
  if (rbf && rbf->mode) {
    e.rbf_mode = rbf->mode;
    e.rbf_eps = rbf->eps;
    e.rbf_from_cor = rbf->from_cor;
    e.rbf_complement = rbf->complement;
  }
  // exact Spearman Gram into a pageable buffer: int32 transport (see host_cor_upper_pipelined); the norms go over first
  const bool s32 = pipe && g_s32_transport && op.precision == BIXCOR_PREC_I8EXACT && op.inv_norm && !e.rbf_mode && !e.keep_sign &&
                   (e.beta == 1.0 || e.beta_int == 6) && !umma::kI8Wide;
  if (s32) {
    RC_TRY(c->inv_host.reserve(sizeof(double) * (size_t)p));
    cudaEvent_t here = ctx_event(c);
    CU_TRY(cudaEventRecord(here, c->compute));
    CU_TRY(cudaStreamWaitEvent(c->copy_out, here, 0));
    CU_TRY(cudaMemcpyAsync(c->inv_host.ptr, op.inv_norm, sizeof(double) * (size_t)p, cudaMemcpyDeviceToHost, c->copy_out));
    e.out_s32 = 1;
    pipe->ex.inv = (const double*)c->inv_host.ptr;
    pipe->ex.p = p;
    pipe->ex.shift = shift;
    pipe->ex.pow6 = e.beta_int == 6;
  }
  const bool f32t = pipe && f32_transport(op.precision, e);
  if (f32t) RC_TRY(c->out32.reserve(sizeof(float) * (size_t)std::max<int64_t>(1, packed_row_offset(row_end, p, shift) - e.out_base)));
  TileList* tl;
  RC_TRY(get_tiles(c, p, row_begin, row_end, 1, 8, umma_pairs(op.precision, EPI_PACKED), &tl, swap_tiles(op.precision, e)));
  if (!pipe)  // results stay in HBM: one launch over all tiles (no per-band tail)
    return launch_gram(c, op, nullptr, p, tl->d_tiles, tl->ntiles, EPI_PACKED, e, nullptr, row_begin, row_end);
  const int ngroups = (int)tl->group_tile.size() - 1;
  for (int g = 0; g < ngroups; ++g) {
    const int t0 = tl->group_tile[g], t1 = tl->group_tile[g + 1];
    RC_TRY(launch_gram(c, op, nullptr, p, tl->d_tiles + t0, t1 - t0, EPI_PACKED, e, nullptr, tl->group_row[g],
                       tl->group_row[g + 1]));
    const int64_t o0 = packed_row_offset(tl->group_row[g], p, shift) - e.out_base;
    const int64_t o1 = packed_row_offset(tl->group_row[g + 1], p, shift) - e.out_base;
    if (f32t) {
      RC_TRY(push_band_f32(c, pipe, h_out + o0, d_out, d_out + o0, o1 - o0));
      continue;
    }
    cudaEvent_t ev = ctx_event(c);
    CU_TRY(cudaEventRecord(ev, c->compute));
    if (s32) RC_TRY(pipe->push_s32(h_out + o0, (const int32_t*)d_out + o0, o1 - o0, e.out_base + o0, ev));
    else RC_TRY(pipe->push(h_out + o0, d_out + o0, sizeof(double) * (size_t)(o1 - o0), ev));
  }
  return BIXCOR_OK;
}

static int dev_cor_upper_rows(Ctx* c, const double* d_x, int64_t n, int64_t p, int spearman, int shift, double beta,
                              int precision, int64_t row_begin, int64_t row_end, double* d_out, D2HPipe* pipe,
                              double* h_out, const RbfSpec* rbf = nullptr) {
  RC_TRY(validate_rows(p, row_begin, row_end));
  if (row_begin == row_end) return BIXCOR_OK;
  precision = resolve_precision(precision, spearman, n);
  Operand op;
  RC_TRY(prepare_operand(c, d_x, n, p, spearman, precision, 0, &op));
  return dev_cor_upper_rows_op(c, op, p, shift, beta, row_begin, row_end, d_out, pipe, h_out, rbf);
}

static int finish(Ctx* c);

// ---- fused operand exchange over peer memory (struct Xchg) -------------------------------------------------------------
// unmap the peers' blocks (this rank's own block stays: peers may still have it mapped)
static void xchg_release_peers(Ctx* c) {
  Xchg& X = c->xchg;
  if (!X.local) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < X.world; ++r)
    if (r != X.rank && X.peer[r]) {
      cudaIpcCloseMemHandle(X.peer[r]);
      X.peer[r] = nullptr;
    }
  X.opened = 0;
  cudaGetLastError();
}
static void xchg_free(Ctx* c) {
  Xchg& X = c->xchg;
  if (!X.local) return;
  xchg_release_peers(c);
  cudaFree(X.local);
  cudaGetLastError();
  X = Xchg();
}

// Column kernel of this rank's gene slab with the records stored into every rank's buffer `b`, then the cross-GPU barrier.
// d_x: the full n x p matrix (x_is_slab == 0) or only the slab's genes (x_is_slab != 0).
static int xchg_build_and_sync(Ctx* c, const double* d_x, int x_is_slab, int spearman) {
  Xchg& X = c->xchg;
  const uint32_t epoch = ++X.epoch;
  const int b = (int)(epoch & 1u);
  const int64_t n = X.n, p = X.p;
  const int64_t g0 = std::min(p, (int64_t)X.rank * X.slab), g1 = std::min(p, g0 + X.slab);
  const int K = round_up((int)n, k_padding(X.precision));
  if (g1 > g0) {
    const double* xs = x_is_slab ? d_x : d_x + g0 * n;
    const bool fused = spearman && X.precision == BIXCOR_PREC_I8EXACT && n <= 1024;
    if (fused) {  // rank kernel with peer stores: the NVLink traffic overlaps the sort
      ColumnOut co;
      memset(&co, 0, sizeof co);
      co.kpad = K;
      co.mode = COL_I8;
      co.i8 = (int8_t*)(X.local + X.op_off[b]) + g0 * 2 * K;
      co.ld8 = 2 * K;
      co.plane8 = K;
      co.inv_norm = (double*)(X.local + X.inv_off[b]) + g0;
      PeerOut po;
      memset(&po, 0, sizeof po);
      for (int r = 0; r < X.world; ++r) {
        if (r == X.rank) continue;
        po.i8[po.n] = (int8_t*)(X.peer[r] + X.op_off[b]) + g0 * 2 * K;
        po.inv_norm[po.n] = (double*)(X.peer[r] + X.inv_off[b]) + g0;
        po.n++;
      }
      int e = 1;
      while (32 * e < n) e <<= 1;
      const int64_t cnt = g1 - g0;
      const unsigned grid = (unsigned)((cnt + 3) / 4);
      Timed tm(c, T_COLS);
      switch (e) {
        case 1: rank_columns_warp_peers_kernel<1><<<grid, 128, 0, c->compute>>>(xs, n, (int)n, cnt, co, po); break;
        case 2: rank_columns_warp_peers_kernel<2><<<grid, 128, 0, c->compute>>>(xs, n, (int)n, cnt, co, po); break;
        case 4: rank_columns_warp_peers_kernel<4><<<grid, 128, 0, c->compute>>>(xs, n, (int)n, cnt, co, po); break;
        case 8: rank_columns_warp_peers_kernel<8><<<grid, 128, 0, c->compute>>>(xs, n, (int)n, cnt, co, po); break;
        case 16: rank_columns_warp_peers_kernel<16><<<grid, 128, 0, c->compute>>>(xs, n, (int)n, cnt, co, po); break;
        default: rank_columns_warp_peers_kernel<32><<<grid, 128, 0, c->compute>>>(xs, n, (int)n, cnt, co, po); break;
      }
      g_launches++;
      CU_TRY(cudaGetLastError());
    } else {  // any other column kernel: build the slab locally, then copy its records (and norms) to the peers
      RC_TRY(build_operand(c, xs, n, g1 - g0, spearman, X.precision, 0, g1 - g0, X.local + X.op_off[b] + (size_t)g0 * X.rb,
                           (double*)(X.local + X.inv_off[b]) + g0));
      Timed tm(c, T_OTHER);
      PeerCopy pc;
      memset(&pc, 0, sizeof pc);
      for (int r = 0; r < X.world; ++r)
        if (r != X.rank) pc.dst[pc.n++] = X.peer[r] + X.op_off[b] + (size_t)g0 * X.rb;
      const int64_t bytes = (int64_t)(g1 - g0) * (int64_t)X.rb;  // records are multiples of 16 bytes, cudaMalloc-aligned
      peer_broadcast_kernel<<<(unsigned)std::min<int64_t>((bytes / 16 + 255) / 256, (int64_t)c->sm_count * 8), 256, 0, c->compute>>>(
          X.local + X.op_off[b] + (size_t)g0 * X.rb, pc, bytes);
      g_launches++;
      if (X.precision == BIXCOR_PREC_I8EXACT) {
        for (int r = 0, q = 0; r < X.world; ++r)
          if (r != X.rank) pc.dst[q++] = X.peer[r] + X.inv_off[b] + sizeof(double) * (size_t)g0;
        const int64_t ib = (int64_t)sizeof(double) * (X.slab);  // whole (128-aligned) slab: a multiple of 16 bytes
        peer_broadcast_kernel<<<(unsigned)((ib / 16 + 255) / 256), 256, 0, c->compute>>>(X.local + X.inv_off[b] + sizeof(double) * (size_t)g0, pc, ib);
        g_launches++;
      }
      CU_TRY(cudaGetLastError());
    }
  }
  {
    Timed tm(c, T_OTHER);
    PeerFlags pf;
    memset(&pf, 0, sizeof pf);
    pf.n = X.world;
    for (int r = 0; r < X.world; ++r) pf.flags[r] = (uint32_t*)(X.peer[r] + X.flag_off);
    peer_barrier_kernel<<<1, 32, 0, c->compute>>>(pf, X.rank, epoch);
    g_launches++;
  }
  CU_TRY(cudaGetLastError());
  return BIXCOR_OK;
}

// Host -> host packed correlation of rows [row_begin,row_end) with both PCIe directions busy at once.  The packed
// result (8 bytes per pair) is ten times the input, so the device-to-host stream is the floor of the call; it should
// start at once and never wait.  Tile (i, j), i <= j, only needs the genes of block rows i and j, so the bands of the
// triangle are processed BOTTOM-UP: the last band needs the last genes only - their columns are uploaded, ranked /
// standardised and multiplied first, and its packed range is already on its way back while the earlier gene slabs are
// still being uploaded (host-to-device on its own stream and DMA engine).  Per band: upload the band's gene columns
// (8 MB), column kernel for those genes, Gram of the band, device-to-host of its packed range.
static int host_cor_upper_pipelined(Ctx* c, const double* h_x, int64_t n, int64_t p, int spearman, int shift, double beta,
                                    int precision, int64_t row_begin, int64_t row_end, double* h_out, const RbfSpec* rbf) {
  RC_TRY(validate_rows(p, row_begin, row_end));
  if (row_begin == row_end) return BIXCOR_OK;
  precision = resolve_precision(precision, spearman, n);
  RC_TRY(check_precision(precision, spearman, n));
  const int64_t o_base = packed_row_offset(row_begin, p, shift), o_end = packed_row_offset(row_end, p, shift);
  RC_TRY(c->x[0].reserve(sizeof(double) * (size_t)n * p));
  RC_TRY(c->out.reserve(sizeof(double) * (size_t)std::max<int64_t>(1, o_end - o_base)));
  double* d_x = (double*)c->x[0].ptr;
  double* d_out = (double*)c->out.ptr;
  D2HPipe pipe(c);
  if (!exchangeable(precision)) {  // Ozaki: the digit slicing runs over the whole operand, no per-slab form
    RC_TRY(h2d(c, d_x, h_x, sizeof(double) * (size_t)n * p, c->compute));
    RC_TRY(dev_cor_upper_rows(c, d_x, n, p, spearman, shift, beta, precision, row_begin, row_end, d_out, &pipe, h_out, rbf));
    RC_TRY(finish(c));
    return pipe.finish();
  }
  RC_TRY(c->op[0].reserve((size_t)operand_row_bytes(n, precision) * p));
  RC_TRY(c->aux[0].reserve(sizeof(double) * p));
  Operand op;
  bind_operand(&op, n, p, precision, c->op[0].ptr, (const double*)c->aux[0].ptr);
  EpiParams e = make_epi(p);
  e.out = d_out;
  e.out_base = o_base;
  e.shift = shift;
  set_beta(e, beta, 0);
  if (rbf && rbf->mode) {
    e.rbf_mode = rbf->mode;
    e.rbf_eps = rbf->eps;
    e.rbf_from_cor = rbf->from_cor;
    e.rbf_complement = rbf->complement;
  }
  // exact Spearman Gram into a pageable buffer: int32 transport (half the PCIe bytes; the copier threads finish r)
  const bool s32 = g_s32_transport && precision == BIXCOR_PREC_I8EXACT && !e.rbf_mode && !e.keep_sign &&
                   (e.beta == 1.0 || e.beta_int == 6) && !umma::kI8Wide;  // (page-locked destinations too: 4 instead of 8 bytes per pair on the wire)
  if (s32) {
    RC_TRY(c->inv_host.reserve(sizeof(double) * (size_t)p));
    e.out_s32 = 1;
    pipe.ex.inv = (const double*)c->inv_host.ptr;
    pipe.ex.p = p;
    pipe.ex.shift = shift;
    pipe.ex.pow6 = e.beta_int == 6;
  }
  const bool f32t = f32_transport(precision, e);  // split-float kinds: binary32 on the wire
  if (f32t) RC_TRY(c->out32.reserve(sizeof(float) * (size_t)std::max<int64_t>(1, o_end - o_base)));
  TileList* tl;
  RC_TRY(get_tiles(c, p, row_begin, row_end, 1, 8, umma_pairs(precision, EPI_PACKED), &tl, swap_tiles(precision, e)));
  const int ngroups = (int)tl->group_tile.size() - 1;
  const bool pinned_src = is_pinned(h_x);
  // Band g misses the genes [lo(g), hi(g)): its own block rows; the bottom band also brings everything to its right.
  auto lo = [&](int g) { return tl->group_row[g]; };
  auto hi = [&](int g) { return g == ngroups - 1 ? p : tl->group_row[g + 1]; };
  std::vector<cudaEvent_t> up((size_t)ngroups);
  for (auto& ev : up) ev = ctx_event(c);
  auto upload = [&](int g) -> int {  // enqueue the upload of band g's genes and record up[g] behind it
    const size_t off = (size_t)lo(g) * n, bytes = sizeof(double) * (size_t)(hi(g) - lo(g)) * n;
    if (pinned_src) {
      g_h2d_bytes += (int64_t)bytes;
      CU_TRY(cudaMemcpyAsync(d_x + off, h_x + off, bytes, cudaMemcpyHostToDevice, c->copy_in));
    } else RC_TRY(h2d(c, d_x + off, h_x + off, bytes, c->copy_in));  // staged through the inbound pinned ring (blocks this thread)
    CU_TRY(cudaEventRecord(up[(size_t)g], c->copy_in));
    return BIXCOR_OK;
  };
  // Pinned source: the uploads are asynchronous DMAs, all of them are queued here and now.  Pageable source: staging
  // needs a host thread for the whole upload, so a helper thread runs it bottom-up while this thread launches the bands
  // and feeds the device-to-host stream; `enqueued` is the lowest band whose event has been recorded.
  std::mutex um;
  std::condition_variable ucv;
  int enqueued = ngroups, up_rc = BIXCOR_OK;
  ThreadGroup uploader;
  if (pinned_src) {
    for (int g = ngroups - 1; g >= 0; --g) RC_TRY(upload(g));
    enqueued = 0;
  } else {
    uploader.run([&] {
      int rc = BIXCOR_OK;
      try {
        if (cudaSetDevice(c->device) != cudaSuccess) rc = fail(BIXCOR_ERR_CUDA, "cudaSetDevice(%d) failed", c->device);
        for (int g = ngroups - 1; g >= 0 && rc == BIXCOR_OK; --g) {
          rc = upload(g);
          std::lock_guard<std::mutex> lk(um);
          if (rc == BIXCOR_OK) enqueued = g;
          else up_rc = rc;
          ucv.notify_all();
        }
      } catch (...) {
        rc = api_exception("upload thread");
      }
      std::lock_guard<std::mutex> lk(um);
      if (rc != BIXCOR_OK) up_rc = rc;
      ucv.notify_all();
    });
  }
  int loop_rc = BIXCOR_OK;
  for (int g = ngroups - 1; g >= 0 && loop_rc == BIXCOR_OK; --g) {
    loop_rc = [&]() -> int {
      {
        std::unique_lock<std::mutex> lk(um);
        ucv.wait(lk, [&] { return enqueued <= g || up_rc != BIXCOR_OK; });
        if (up_rc != BIXCOR_OK) return up_rc;
      }
      CU_TRY(cudaStreamWaitEvent(c->compute, up[(size_t)g], 0));
      RC_TRY(build_operand(c, d_x, n, p, spearman, precision, lo(g), hi(g), c->op[0].ptr, (double*)c->aux[0].ptr));
      if (s32) {  // the norms of the new genes go to the host ahead of the band's data (same stream: ordered before its DMAs)
        cudaEvent_t cols = ctx_event(c);
        CU_TRY(cudaEventRecord(cols, c->compute));
        CU_TRY(cudaStreamWaitEvent(c->copy_out, cols, 0));
        CU_TRY(cudaMemcpyAsync((double*)c->inv_host.ptr + lo(g), (const double*)c->aux[0].ptr + lo(g), sizeof(double) * (size_t)(hi(g) - lo(g)),
                               cudaMemcpyDeviceToHost, c->copy_out));
      }
      const int t0 = tl->group_tile[g], t1 = tl->group_tile[g + 1];
      RC_TRY(launch_gram(c, op, nullptr, p, tl->d_tiles + t0, t1 - t0, EPI_PACKED, e, nullptr, tl->group_row[g], tl->group_row[g + 1]));
      const int64_t o0 = packed_row_offset(tl->group_row[g], p, shift) - o_base;
      const int64_t o1 = packed_row_offset(tl->group_row[g + 1], p, shift) - o_base;
      if (f32t) return push_band_f32(c, &pipe, h_out + o0, d_out, d_out + o0, o1 - o0);
      cudaEvent_t ev = ctx_event(c);
      CU_TRY(cudaEventRecord(ev, c->compute));
      if (s32) return pipe.push_s32(h_out + o0, (const int32_t*)d_out + o0, o1 - o0, o_base + o0, ev);
      return pipe.push(h_out + o0, d_out + o0, sizeof(double) * (size_t)(o1 - o0), ev);
    }();
  }
  uploader.join();  // (also on the error path: the thread references this frame)
  if (loop_rc != BIXCOR_OK) return loop_rc;
  RC_TRY(finish(c));
  return pipe.finish();
}

static int check_rbf(int rbf_type, double epsilon) {
  if (rbf_type != BIXCOR_RBF_GAUSSIAN && rbf_type != BIXCOR_RBF_BUMP && rbf_type != BIXCOR_RBF_INVERSE_QUADRATIC)
    return fail(BIXCOR_ERR_INVALID, "Invalid RBF function: %d", rbf_type);
  if (!(epsilon > 0.0)) return fail(BIXCOR_ERR_INVALID, "epsilon must be positive");
  return BIXCOR_OK;
}

static int dev_diffcor_rows(Ctx* c, const double* d_xa, int64_t na, const double* d_xb, int64_t nb, int64_t p,
                            int spearman, int precision, int64_t row_begin, int64_t row_end, double* d_ra,
                            double* d_rb, double* d_z, double* d_pv, D2HPipe* pipe, double* h_ra, double* h_rb,
                            double* h_z, double* h_pv) {
  RC_TRY(validate_rows(p, row_begin, row_end));
  if (na < 4 || nb < 4) return fail(BIXCOR_ERR_INVALID, "differential correlation needs n_a, n_b >= 4");
  if (row_begin == row_end) return BIXCOR_OK;
  precision = resolve_precision(precision, spearman, std::max(na, nb));
  Operand a, b;
  RC_TRY(prepare_operand(c, d_xa, na, p, spearman, precision, 0, &a));
  RC_TRY(prepare_operand(c, d_xb, nb, p, spearman, precision, 1, &b));
  const double cf = spearman ? g_fieller : 1.0;
  const double inv_se = 1.0 / std::sqrt(cf / (double)(na - 3) + cf / (double)(nb - 3));
  // Two packed Grams through the lean epilogue, then the Fisher-z step as a full-occupancy streaming kernel
  // (fused into the Gram epilogue it is latency bound on 8 warps per SM: 10-18 ms against ~4 ms at 20000 genes).
  EpiParams ea = make_epi(p), eb = make_epi(p);
  ea.out_base = eb.out_base = packed_row_offset(row_begin, p, 1);
  ea.shift = eb.shift = 1;
  ea.out = d_ra;
  eb.out = d_rb;
  TileList* tl;
  RC_TRY(get_tiles(c, p, row_begin, row_end, 1, 8, umma_pairs(precision, EPI_PACKED), &tl, swap_tiles(precision, ea)));
  const int ngroups = pipe ? (int)tl->group_tile.size() - 1 : 1;
  for (int g = 0; g < ngroups; ++g) {
    const int t0 = pipe ? tl->group_tile[g] : 0, t1 = pipe ? tl->group_tile[g + 1] : tl->ntiles;
    const int64_t g0 = pipe ? tl->group_row[g] : row_begin, g1 = pipe ? tl->group_row[g + 1] : row_end;
    RC_TRY(launch_gram(c, a, nullptr, p, tl->d_tiles + t0, t1 - t0, EPI_PACKED, ea, nullptr, g0, g1));
    RC_TRY(launch_gram(c, b, nullptr, p, tl->d_tiles + t0, t1 - t0, EPI_PACKED, eb, nullptr, g0, g1));
    const int64_t o0 = packed_row_offset(g0, p, 1) - ea.out_base;
    const int64_t o1 = packed_row_offset(g1, p, 1) - ea.out_base;
    if (o1 > o0) {
      Timed tm(c, T_OTHER);
      const unsigned grid = (unsigned)std::min<int64_t>((o1 - o0 + 255) / 256, (int64_t)c->sm_count * 32);
      diffcor_zp_kernel<<<grid, 256, 0, c->compute>>>(d_ra + o0, d_rb + o0, o1 - o0, inv_se, d_z + o0, d_pv + o0);
      g_launches++;
    }
    if (pipe) {
      cudaEvent_t ev = ctx_event(c);
      CU_TRY(cudaEventRecord(ev, c->compute));
      const size_t bytes = sizeof(double) * (size_t)(o1 - o0);
      RC_TRY(pipe->push(h_ra + o0, d_ra + o0, bytes, ev));
      RC_TRY(pipe->push(h_rb + o0, d_rb + o0, bytes, nullptr));
      RC_TRY(pipe->push(h_z + o0, d_z + o0, bytes, nullptr));
      RC_TRY(pipe->push(h_pv + o0, d_pv + o0, bytes, nullptr));
    }
  }
  CU_TRY(cudaGetLastError());
  return BIXCOR_OK;
}

// Dense correlation / affinity.  Full range + mirror on one GPU; a strict sub-range
// computes the full strip (all block columns) of the requested genes' columns.
static int dev_affinity(Ctx* c, const double* d_x, int64_t n, int64_t p, int spearman, double beta, int keep_sign,
                        int precision, int64_t row_begin, int64_t row_end, double* d_app) {
  RC_TRY(validate_rows(p, row_begin, row_end));
  if (row_begin == row_end) return BIXCOR_OK;
  precision = resolve_precision(precision, spearman, n);
  Operand op;
  RC_TRY(prepare_operand(c, d_x, n, p, spearman, precision, 0, &op));
  const bool full = (row_begin == 0 && row_end == p);
  TileList* tl;
  RC_TRY(get_tiles(c, p, row_begin, row_end, full ? 1 : 0, 8, umma_pairs(precision, EPI_DENSE), &tl));
  EpiParams e = make_epi(p);
  e.out = d_app;
  e.ldo = p;
  e.mirror = full ? 1 : 0;
  e.diag_one = 1;
  set_beta(e, beta, keep_sign);
  e.abs_out = (!keep_sign && beta == 1.0) ? 1 : 0;  // unsigned affinity without a power: |r| (for beta != 1, |r|^beta already is)
  return launch_gram(c, op, nullptr, p, tl->d_tiles, tl->ntiles, EPI_DENSE, e, nullptr, row_begin, row_end);
}

static int dev_tom_connectivity(Ctx* c, const double* d_a, int64_t p, int signed_, int64_t row_begin, int64_t row_end,
                                double* d_k) {
  if (row_begin < 0 || row_end > p || row_begin > row_end) return fail(BIXCOR_ERR_INVALID, "bad row range");
  if (row_begin == row_end) return BIXCOR_OK;
  Timed tm(c, T_OTHER);
  tom_connectivity_kernel<<<(unsigned)(row_end - row_begin), 256, 0, c->compute>>>(d_a, p, row_begin, signed_ ? 1 : 0, d_k);
  g_launches++;
  CU_TRY(cudaGetLastError());
  return BIXCOR_OK;
}

// Split rows [row_begin,row_end) of the dense affinity matrix (column i == row i) into the hi / lo operand planes of a
// split-float kind: records [K hi][K lo] per gene in `d_planes` (base pointer of the full p-gene buffer).
static int dev_split_planes(Ctx* c, const double* d_a, int64_t p, int precision, int64_t row_begin, int64_t row_end,
                            void* d_planes) {
  if (!is_split_float(precision)) return fail(BIXCOR_ERR_INVALID, "operand planes exist for BIXCOR_PREC_TF32X3 / BIXCOR_PREC_F16X3");
  if (row_begin < 0 || row_end > p || row_begin > row_end) return fail(BIXCOR_ERR_INVALID, "bad row range");
  if (row_begin == row_end) return BIXCOR_OK;
  const int K = round_up((int)p, k_padding(precision));
  const bool half = precision == BIXCOR_PREC_F16X3;
  ColumnOut co;
  memset(&co, 0, sizeof co);
  co.mode = COL_TF32;
  co.half_planes = half;
  co.kpad = K;
  co.f32 = half ? (float*)((__half*)d_planes + row_begin * 2 * K) : (float*)d_planes + row_begin * 2 * K;
  co.ld32 = 2 * K;
  co.plane32 = K;
  RC_TRY(c->flag.reserve(256));
  int* d_flag = (int*)c->flag.ptr;
  CU_TRY(cudaMemsetAsync(d_flag, 0, sizeof(int), c->compute));
  {
    Timed tm(c, T_COLS);
    split_tf32_kernel<<<(unsigned)(row_end - row_begin), 256, 0, c->compute>>>(d_a + row_begin * p, p, (int)p, co, d_flag);
    g_launches++;
  }
  CU_TRY(cudaGetLastError());
  if (half) {
    int overflow = 0;
    CU_TRY(cudaMemcpyAsync(&overflow, d_flag, sizeof(int), cudaMemcpyDeviceToHost, c->compute));
    CU_TRY(cudaStreamSynchronize(c->compute));
    if (overflow)
      return fail(BIXCOR_ERR_UNSUPPORTED, "BIXCOR_PREC_F16X3 needs affinities below 63 in magnitude; use BIXCOR_PREC_TF32X3");
  }
  return BIXCOR_OK;
}

// The TOM GEMM operand is the affinity matrix itself (A symmetric: column i == row i).
// d_planes != nullptr (split-float kinds): the hi / lo planes of ALL genes are given (built per slab by their owners and
// exchanged); d_a may then be nullptr - the epilogue rebuilds a_ij and the diagonal from the planes.
// pipe != nullptr (host entry points): the rows are launched in bands of block rows and every finished band goes to the
// host (h_out addresses the same layout as d_out) while the next bands compute.  Top-down: the A * A produces its output
// (1.6 GB in 25 ms with the f16 kind) faster than PCIe takes it, so the largest bands go first and the device-to-host
// stream is busy from the first band on (bottom-up measured 52 ms host -> host for the f16 kind: the stream idles on the
// small bands and the top band's 320 MB leave after the last kernel); a dense mirrored row i is complete once the bands
// down to its own have run (its lower part is the mirror of earlier rows).
static int dev_tom_rows(Ctx* c, const double* d_a, const double* d_k, int64_t p, int signed_, int version,
                        int precision, int packed, int64_t row_begin, int64_t row_end, double* d_out,
                        const void* d_planes = nullptr, D2HPipe* pipe = nullptr, double* h_out = nullptr) {
  RC_TRY(validate_rows(p, row_begin, row_end));
  if (version != BIXCOR_TOM_V1 && version != BIXCOR_TOM_V2)
    return fail(BIXCOR_ERR_INVALID, "Invalid TOM version type: %d", version);
  if (row_begin == row_end) return BIXCOR_OK;
  precision = resolve_precision_float(precision);
  if (d_planes && !is_split_float(precision)) return fail(BIXCOR_ERR_INVALID, "operand planes exist for the split-float kinds only");
  if (!d_a && !d_planes) return fail(BIXCOR_ERR_INVALID, "null pointer");
  RC_TRY(c->vec.reserve(sizeof(double) * (size_t)p * 4));
  double* d_diag = (double*)c->vec.ptr;  // first p entries: diagonal
  Operand op;
  op.precision = precision;
  op.p = p;
  if (d_planes) {
    const int K = round_up((int)p, k_padding(precision));
    op.K = K;
    op.f32 = (const float*)d_planes;
    Timed tm(c, T_OTHER);
    extract_diag_planes_kernel<<<(unsigned)((p + 255) / 256), 256, 0, c->compute>>>(d_planes, 2 * (int64_t)K, K, precision == BIXCOR_PREC_F16X3, p,
                                                                                   d_diag);
    g_launches++;
  } else {
    {
      Timed tm(c, T_OTHER);
      extract_diag_kernel<<<(unsigned)((p + 255) / 256), 256, 0, c->compute>>>(d_a, p, d_diag);
      g_launches++;
    }
    if (precision == BIXCOR_PREC_F64) {
      if (p % dmma::BK != 0) {
        // pad K: copy A into a K-padded operand buffer
        const int K = round_up((int)p, dmma::BK);
        RC_TRY(c->op[1].reserve(sizeof(double) * (size_t)K * p));
        CU_TRY(cudaMemsetAsync(c->op[1].ptr, 0, sizeof(double) * (size_t)K * p, c->compute));
        CU_TRY(cudaMemcpy2DAsync(c->op[1].ptr, sizeof(double) * K, d_a, sizeof(double) * p, sizeof(double) * p, p,
                                 cudaMemcpyDeviceToDevice, c->compute));
        op.K = K;
        op.z64 = (const double*)c->op[1].ptr;
      } else {
        op.K = (int)p;
        op.z64 = d_a;
      }
    } else if (is_split_float(precision)) {
      const int K = round_up((int)p, k_padding(precision));
      RC_TRY(c->op[1].reserve(sizeof(float) * 2 * (size_t)K * p));
      RC_TRY(dev_split_planes(c, d_a, p, precision, 0, p, c->op[1].ptr));
      op.K = K;
      op.f32 = (const float*)c->op[1].ptr;
    } else if (is_ozaki(precision)) {
      op.K = (int)p;
      op.z64 = d_a;
      RC_TRY(ozaki_slice(c, d_a, p, p, p, 1, &op));  // rows of A are its columns (symmetric): digit planes of every gene
    } else {
      return fail(BIXCOR_ERR_INVALID, "TOM supports BIXCOR_PREC_F64, BIXCOR_PREC_TF32X3, BIXCOR_PREC_F16X3 and BIXCOR_PREC_OZAKI");
    }
  }
  const bool full = (row_begin == 0 && row_end == p);
  const int upper = (packed || full) ? 1 : 0;
  TileList* tl;
  RC_TRY(get_tiles(c, p, row_begin, row_end, upper, 8, umma_pairs(precision, EPI_TOM), &tl));
  EpiParams e = make_epi(p);
  e.out = d_out;
  e.ldo = p;
  e.shift = 1;
  e.out_base = packed ? packed_row_offset(row_begin, p, 1) : 0;
  e.mirror = (!packed && full) ? 1 : 0;
  e.A = d_a;
  if (!d_a) {
    e.A_planes = d_planes;
    e.planes_K = op.K;
    e.planes_ld = 2 * (int64_t)op.K;
    e.planes_half = precision == BIXCOR_PREC_F16X3;
  }
  e.kvec = d_k;
  e.dvec = d_diag;
  e.tom_v2 = (version == BIXCOR_TOM_V2);
  e.tom_signed = signed_ ? 1 : 0;
  e.tom_packed = packed ? 1 : 0;
  e.upper_tiles = upper;
  // non-packed strip mode writes into out + 0 with absolute (i*p + j) addressing
  if (!pipe) return launch_gram(c, op, nullptr, p, tl->d_tiles, tl->ntiles, EPI_TOM, e, nullptr, row_begin, row_end);
  const bool f32t = f32_transport(precision, e);  // split-float kinds: binary32 on the wire (out32 mirrors d_out's indexing)
  if (f32t)
    RC_TRY(c->out32.reserve(sizeof(float) * (size_t)(packed ? std::max<int64_t>(1, packed_row_offset(row_end, p, 1) - e.out_base) : row_end * p)));
  const int ngroups = (int)tl->group_tile.size() - 1;
  for (int t = 0; t < ngroups; ++t) {
    const int g = t;
    const int t0 = tl->group_tile[g], t1 = tl->group_tile[g + 1];
    const int64_t g0 = tl->group_row[g], g1 = tl->group_row[g + 1];
    RC_TRY(launch_gram(c, op, nullptr, p, tl->d_tiles + t0, t1 - t0, EPI_TOM, e, nullptr, g0, g1));
    if (f32t) {
      if (packed) {
        const int64_t o0 = packed_row_offset(g0, p, 1) - e.out_base, o1 = packed_row_offset(g1, p, 1) - e.out_base;
        RC_TRY(push_band_f32(c, pipe, h_out + o0, d_out, d_out + o0, o1 - o0));
      } else {
        RC_TRY(push_band_f32(c, pipe, h_out + g0 * p, d_out, d_out + g0 * p, (g1 - g0) * p));
      }
      continue;
    }
    cudaEvent_t ev = ctx_event(c);
    CU_TRY(cudaEventRecord(ev, c->compute));
    if (packed) {
      const int64_t o0 = packed_row_offset(g0, p, 1) - e.out_base, o1 = packed_row_offset(g1, p, 1) - e.out_base;
      RC_TRY(pipe->push(h_out + o0, d_out + o0, sizeof(double) * (size_t)(o1 - o0), ev));
    } else {
      RC_TRY(pipe->push(h_out + g0 * p, d_out + g0 * p, sizeof(double) * (size_t)(g1 - g0) * p, ev));
    }
  }
  return BIXCOR_OK;
}

// binary16 planes: values of magnitude >= 63 do not fit; the densify kernels raise a device flag that the host reads once
static int sparse_overflow_check(Ctx* c) {
  int overflow = 0;
  CU_TRY(cudaMemcpyAsync(&overflow, c->flag.ptr, sizeof(int), cudaMemcpyDeviceToHost, c->compute));
  CU_TRY(cudaStreamSynchronize(c->compute));
  if (overflow)
    return fail(BIXCOR_ERR_UNSUPPORTED, "BIXCOR_PREC_F16X3 needs sparse values below 63 in magnitude; use BIXCOR_PREC_TF32X3");
  return BIXCOR_OK;
}

// streamed != 0: one slab of a longer sequence - the caller has cleared the overflow flag and checks it at the end (the
// check synchronises the stream, which would serialise the upload of the next slab behind this one's kernels)
static int dev_sparse_accumulate(Ctx* c, const int64_t* d_indptr, const int32_t* d_indices, const float* d_data,
                                 int64_t cells, int64_t genes, int precision, double* d_gram, double* d_colsum,
                                 int streamed = 0) {
  precision = resolve_precision_float(precision);
  if (precision != BIXCOR_PREC_F64 && !is_split_float(precision))
    return fail(BIXCOR_ERR_UNSUPPORTED, "sparse Gram runs on BIXCOR_PREC_F64, BIXCOR_PREC_TF32X3 or BIXCOR_PREC_F16X3");
  const bool fast = is_split_float(precision), half = precision == BIXCOR_PREC_F16X3;
  int* d_flag = nullptr;
  if (half) {
    RC_TRY(c->flag.reserve(256));
    d_flag = (int*)c->flag.ptr;
    if (!streamed) CU_TRY(cudaMemsetAsync(d_flag, 0, sizeof(int), c->compute));
  }
  const int chunk = 8192;  // cells per dense panel (K of one Gram accumulation pass)
  RC_TRY(c->panel.reserve(sizeof(double) * (size_t)chunk * genes));  // f64 panel == two f32 planes
  TileList* tl;
  RC_TRY(get_tiles(c, genes, 0, genes, 1, 8, umma_pairs(precision, EPI_ACCUM), &tl));
  for (int64_t c0 = 0; c0 < cells; c0 += chunk) {
    const int nc = (int)std::min<int64_t>(chunk, cells - c0);
    const int K = round_up(nc, k_padding(precision));
    Operand op;
    op.precision = precision;
    op.K = K;
    op.p = genes;
    {
      Timed tm(c, T_OTHER);
      CU_TRY(cudaMemsetAsync(c->panel.ptr, 0, sizeof(double) * (size_t)K * genes, c->compute));
      if (half) {
        csr_densify_f16_kernel<<<(nc * 32 + 255) / 256, 256, 0, c->compute>>>(d_indptr, d_indices, d_data, c0, nc, genes,
                                                                               (__half*)c->panel.ptr, K, d_flag);
        panel_rowsum_f16_kernel<<<(unsigned)genes, 128, 0, c->compute>>>((const __half*)c->panel.ptr, K, nc, d_colsum);
        op.f32 = (const float*)c->panel.ptr;
      } else if (fast) {
        csr_densify_tf32_kernel<<<(nc * 32 + 255) / 256, 256, 0, c->compute>>>(d_indptr, d_indices, d_data, c0, nc, genes,
                                                                                (float*)c->panel.ptr, K);
        panel_rowsum_tf32_kernel<<<(unsigned)genes, 128, 0, c->compute>>>((const float*)c->panel.ptr, K, nc, d_colsum);
        op.f32 = (const float*)c->panel.ptr;
      } else {
        csr_densify_kernel<<<(nc * 32 + 255) / 256, 256, 0, c->compute>>>(d_indptr, d_indices, d_data, c0, nc, genes,
                                                                           (double*)c->panel.ptr, K);
        panel_rowsum_kernel<<<(unsigned)genes, 128, 0, c->compute>>>((const double*)c->panel.ptr, K, nc, d_colsum);
        op.z64 = (const double*)c->panel.ptr;
      }
      g_launches += 2;
    }
    EpiParams e = make_epi(genes);
    e.out = d_gram;
    e.ldo = genes;
    RC_TRY(launch_gram(c, op, nullptr, genes, tl->d_tiles, tl->ntiles, EPI_ACCUM, e));
  }
  CU_TRY(cudaGetLastError());
  if (half && !streamed) RC_TRY(sparse_overflow_check(c));
  return BIXCOR_OK;
}

static int dev_sparse_finalise(Ctx* c, const double* d_gram, const double* d_colsum, int64_t total_cells,
                               int64_t genes, double* d_out) {
  Timed tm(c, T_OTHER);
  dim3 grid((unsigned)((genes + 255) / 256), (unsigned)genes);
  sparse_finalise_kernel<<<grid, 256, 0, c->compute>>>(d_gram, d_colsum, (double)total_cells, genes, d_out);
  g_launches++;
  CU_TRY(cudaGetLastError());
  return BIXCOR_OK;
}

static int finish(Ctx* c) {
  CU_TRY(cudaStreamSynchronize(c->compute));
  return BIXCOR_OK;
}
static int finish_dev(Ctx* c) { return g_async ? BIXCOR_OK : finish(c); }

// run fn(ctx, rank) on every device of the process (one host thread per device)
template <class F>
static int for_each_device(F fn) {
  const int nd = (int)g_ctx.size();
  if (nd == 1) {
    CU_TRY(cudaSetDevice(g_ctx[0]->device));
    call_begin(g_ctx[0]);
    return fn(g_ctx[0], 0, 1);
  }
  std::vector<int> rcs(nd, BIXCOR_OK);
  ThreadGroup tg;
  for (int d = 0; d < nd; ++d)
    tg.run([&, d] {
      try {  // an exception must not leave a worker thread (std::terminate); it becomes this device's return code
        if (cudaSetDevice(g_ctx[d]->device) != cudaSuccess) {
          rcs[d] = fail(BIXCOR_ERR_CUDA, "cudaSetDevice(%d) failed", g_ctx[d]->device);
          return;
        }
        call_begin(g_ctx[d]);
        rcs[d] = fn(g_ctx[d], d, nd);
      } catch (...) {
        rcs[d] = api_exception("device worker");
      }
    });
  tg.join();
  for (int rc : rcs)
    if (rc != BIXCOR_OK) return rc;
  return BIXCOR_OK;
}

static void shard_rows_impl(int64_t p, int shift, int world, int rank, int64_t* r0, int64_t* r1) {
  const double total = shift ? 0.5 * (double)p * (double)(p - 1) : 0.5 * (double)p * (double)(p + 1);
  auto boundary = [&](int q) -> int64_t {
    if (q <= 0) return 0;
    if (q >= world) return p;
    const double target = total * (double)q / (double)world;
    const double a = shift ? 2.0 * p - 1.0 : 2.0 * p + 1.0;
    const double disc = std::max(0.0, a * a - 8.0 * target);
    double r = (a - std::sqrt(disc)) / 2.0;
    int64_t rr = (int64_t)std::nearbyint(r / kTile) * kTile;
    return std::max<int64_t>(0, std::min<int64_t>(p, rr));
  };
  *r0 = boundary(rank);
  *r1 = boundary(rank + 1);
}

}  // namespace bixcor

// ============================================================================
// C ABI
// ============================================================================
using namespace bixcor;

// rows [0,p) split over host threads in blocks balanced by pair count (pure index logic, no GPU involved)
template <class F>
static void parallel_rows(int64_t p, F fn) {
  const unsigned hw = std::thread::hardware_concurrency();
  const int nthr = (p < 2048) ? 1 : (int)std::max(1u, std::min(hw ? hw : 4u, 16u));
  if (nthr == 1) {
    fn(0, p);
    return;
  }
  ThreadGroup tg;
  for (int t = 0; t < nthr; ++t) {
    int64_t r0, r1;
    shard_rows_impl(p, 1, nthr, t, &r0, &r1);
    if (r1 > r0) tg.run([=] { fn(r0, r1); });
  }
  tg.join();
}

// Column c of the symmetric matrix behind a packed vector, visited in row order: rows above the diagonal come from
// element (r, c) of packed row r, the diagonal is 1 (shift) or stored, rows below are the contiguous run of packed row c.
template <class F>
static inline void sym_column_nonzeros(const double* packed, int s, int64_t p, int64_t c, F emit) {
  for (int64_t r = 0; r < c; ++r) {
    const double v = packed[packed_row_offset(r, p, s) + (c - r - s)];
    if (v != 0.0) emit(r, v);
  }
  const double d = s ? 1.0 : packed[packed_row_offset(c, p, 0)];
  if (d != 0.0) emit(c, d);
  const double* row = packed + packed_row_offset(c, p, s);
  for (int64_t r = c + 1; r < p; ++r) {
    const double v = row[r - c - s];
    if (v != 0.0) emit(r, v);
  }
}

// columns [0,p) split evenly over host threads (every column of a symmetric matrix costs the same)
template <class F>
static void parallel_cols(int64_t p, F fn) {
  const unsigned hw = std::thread::hardware_concurrency();
  const int nthr = (p < 2048) ? 1 : (int)std::max(1u, std::min(hw ? hw : 4u, 16u));
  if (nthr == 1) {
    fn(0, p);
    return;
  }
  ThreadGroup tg;
  for (int t = 0; t < nthr; ++t) {
    const int64_t c0 = p * t / nthr, c1 = p * (t + 1) / nthr;
    if (c1 > c0) tg.run([=] { fn(c0, c1); });
  }
  tg.join();
}

// Reusable barrier of the per-device host threads; a failed thread makes every thread leave with its code.
struct PhaseBarrier {
  std::mutex m;
  std::condition_variable cv;
  int n, count = 0, gen = 0, rc = BIXCOR_OK;
  explicit PhaseBarrier(int n_) : n(n_) {}
  int arrive(int my_rc) {
    std::unique_lock<std::mutex> lk(m);
    if (my_rc != BIXCOR_OK && rc == BIXCOR_OK) rc = my_rc;
    const int g = gen;
    if (++count == n) {
      count = 0;
      ++gen;
      cv.notify_all();
    } else {
      cv.wait(lk, [&] { return gen != g; });
    }
    return rc;
  }
};

static void even_slab(int64_t p, int nd, int d, int64_t* c0, int64_t* c1) {
  const int64_t nb = (p + kTile - 1) / kTile;
  *c0 = std::min<int64_t>(p, nb * d / nd * kTile);
  *c1 = std::min<int64_t>(p, nb * (d + 1) / nd * kTile);
}

extern "C" {

const char* bixcor_last_error(void) {
  static thread_local char copy[sizeof g_err];
  try {
    std::lock_guard<std::mutex> lk(g_err_mu);
    memcpy(copy, g_err, sizeof copy);
  } catch (...) {
    copy[0] = 0;
  }
  copy[sizeof copy - 1] = 0;
  return copy;
}

/* test hook: the next API call throws from inside its body (1 = std::bad_alloc, 2 = std::runtime_error) */
int bixcor_debug_inject(int what) {
  g_inject.store(what);
  return BIXCOR_OK;
}

int64_t bixcor_launch_count(void) { return g_launches.load(); }

/* bytes moved host -> device and device -> host by the host entry points since the last reset */
int bixcor_transfer_counters(int64_t* h2d_bytes, int64_t* d2h_bytes, int reset) {
  if (h2d_bytes) *h2d_bytes = g_h2d_bytes.load();
  if (d2h_bytes) *d2h_bytes = g_d2h_bytes.load();
  if (reset) {
    g_h2d_bytes = 0;
    g_d2h_bytes = 0;
  }
  return BIXCOR_OK;
}

// Host halves of the two transports as plain functions (no GPU involved), so that the CPU-only test suite can check what the
// staging threads do to a piece: the int32 -> r (-> |r|^6) expansion of packed elements [g0, g0 + count) and the float -> double
// widening.
int bixcor_host_expand_s32(const int32_t* src, int64_t count, int64_t g0, const double* inv_norm, int64_t p, int shift, int pow6,
                           double* dst) {
  API_BEGIN
  if (!src || !dst || !inv_norm) return fail(BIXCOR_ERR_INVALID, "null pointer");
  shift = shift ? 1 : 0;
  if (p < 1 || count < 0 || g0 < 0 || g0 + count > packed_row_offset(p, p, shift)) return fail(BIXCOR_ERR_INVALID, "bad packed range");
  ExpandS32 X;
  X.inv = inv_norm;
  X.p = p;
  X.shift = shift;
  X.pow6 = pow6 ? 1 : 0;
  expand_packed_s32(dst, src, count, g0, X);
  return BIXCOR_OK;
  API_END
}

int bixcor_host_widen_f32(const float* src, int64_t count, double* dst) {
  API_BEGIN
  if (!src || !dst || count < 0) return fail(BIXCOR_ERR_INVALID, "bad arguments");
  widen_f32(dst, src, count);
  return BIXCOR_OK;
  API_END
}

int bixcor_set_transport(int int32_exact_spearman, int binary32_split_float) {
  API_BEGIN
  if (int32_exact_spearman >= 0) g_s32_transport = int32_exact_spearman != 0;
  if (binary32_split_float >= 0) g_f32_transport = binary32_split_float != 0;
  return BIXCOR_OK;
  API_END
}

int bixcor_set_fieller(double c) {
  API_BEGIN
  if (!(c > 0.0)) return fail(BIXCOR_ERR_INVALID, "Fieller factor must be positive");
  g_fieller = c;
  return BIXCOR_OK;
  API_END
}

int bixcor_init(int n_gpus) { return bixcor_init_devices(0, n_gpus); }

int bixcor_init_devices(int first_device, int n_gpus) {
  API_BEGIN
  std::lock_guard<std::mutex> lk(g_init_mu);
  if (!g_ctx.empty()) {
    if ((int)g_ctx.size() == n_gpus && g_ctx[0]->device == first_device) return BIXCOR_OK;
    nccl_teardown();
    for (auto* c : g_ctx) ctx_destroy(c);
    g_ctx.clear();
  }
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    return fail(BIXCOR_ERR_NO_GPU, "no CUDA device visible (%s); libbixcor has no CPU fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  }
  if (n_gpus < 1 || first_device < 0 || first_device + n_gpus > count)
    return fail(BIXCOR_ERR_INVALID, "requested devices [%d,%d) but %d visible", first_device, first_device + n_gpus,
                count);
  if (const char* v = getenv("BIXCOR_DMMA_VARIANT")) g_dmma_variant = atoi(v);
  if (const char* v = getenv("BIXCOR_UMMA_CTAS")) g_umma_ctas = (atoi(v) >= 1 && atoi(v) <= 3) ? atoi(v) : 2;
  if (const char* v = getenv("BIXCOR_UMMA_MC")) g_umma_mc = atoi(v) != 0;
  if (const char* v = getenv("BIXCOR_UMMA_WIDE")) g_umma_wide = std::max(0, std::min(2, atoi(v)));
  if (const char* v = getenv("BIXCOR_COPY_THREADS")) g_copy_threads = std::max(0, std::min(64, atoi(v)));
  if (const char* v = getenv("BIXCOR_STAGE_SLOT_MB")) g_stage_slot_bytes = (size_t)std::max(1, std::min(256, atoi(v))) << 20;
  if (const char* v = getenv("BIXCOR_STAGE_SLOT_KB")) g_stage_slot_bytes = (size_t)std::max(64, std::min(262144, atoi(v))) << 10;
  if (const char* v = getenv("BIXCOR_STAGE_SLOTS")) g_stage_slots = std::max(2, std::min((int)HostStager::kMaxSlots, atoi(v)));
  if (const char* v = getenv("BIXCOR_STAGE_PIECE_KB")) g_stage_piece = (size_t)std::max(64, std::min(65536, atoi(v))) << 10;
  if (const char* v = getenv("BIXCOR_STAGE_NT")) g_stage_nt = atoi(v) != 0;
  if (const char* v = getenv("BIXCOR_STAGE_SPIN")) g_stage_spin = atoi(v) != 0;
  if (const char* v = getenv("BIXCOR_STAGE_SEQ")) g_stage_seq = atoi(v) != 0;
  if (const char* v = getenv("BIXCOR_S32_TRANSPORT")) g_s32_transport = atoi(v) != 0;
  if (const char* v = getenv("BIXCOR_S32_SLOT_KB")) g_s32_slot_bytes = (size_t)std::max(64, std::min(262144, atoi(v))) << 10;
  if (const char* v = getenv("BIXCOR_S32_SLOTS")) g_s32_slots = std::max(2, std::min((int)HostStager::kMaxSlots, atoi(v)));
  if (const char* v = getenv("BIXCOR_S32_PIECE_KB")) g_s32_piece = (size_t)std::max(64, std::min(65536, atoi(v))) << 10;
  if (const char* v = getenv("BIXCOR_F32_TRANSPORT")) g_f32_transport = atoi(v) != 0;
  for (int d = 0; d < n_gpus; ++d) {
    Ctx* c = nullptr;
    int rc = ctx_create(first_device + d, &c);
    if (rc != BIXCOR_OK) {
      for (auto* cc : g_ctx) ctx_destroy(cc);
      g_ctx.clear();
      return rc;
    }
    g_ctx.push_back(c);
  }
  // NVLink peer access between the devices of the process (slab exchange of the multi-device TOM)
  for (auto* a : g_ctx)
    for (auto* b : g_ctx) {
      if (a == b) continue;
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, a->device, b->device) == cudaSuccess && can) {
        cudaSetDevice(a->device);
        cudaDeviceEnablePeerAccess(b->device, 0);
      }
      cudaGetLastError();  // "already enabled" is fine
    }
  // collectives between the devices of this process (north star: k and the partial Grams are all-reduced with NCCL)
  const char* nv = getenv("BIXCOR_NCCL");
  if (n_gpus > 1 && !(nv && atoi(nv) == 0)) {
    if (g_nccl.load()) {
      std::vector<int> devs;
      for (auto* c : g_ctx) devs.push_back(c->device);
      g_comms.assign((size_t)n_gpus, nullptr);
      ncclResult_t r = g_nccl.CommInitAll(g_comms.data(), n_gpus, devs.data());
      if (r != ncclSuccess) {
        fail(BIXCOR_ERR_CUDA, "ncclCommInitAll: %s (falling back to peer copies)", g_nccl.GetErrorString(r));
        g_comms.clear();
      }
    } else {
      fail(BIXCOR_ERR_CUDA, "NCCL unavailable: %s (falling back to peer copies)", g_nccl.error);
    }
  }
  return BIXCOR_OK;
  API_END
}

/* NCCL_VERSION_CODE of the library serving the in-process collectives, 0 when they are not active */
int bixcor_nccl_version(void) { return g_comms.empty() ? 0 : g_nccl.version; }

void bixcor_shutdown(void) {
  try {
    std::lock_guard<std::mutex> lk(g_init_mu);
    nccl_teardown();
    for (auto* c : g_ctx) ctx_destroy(c);
    g_ctx.clear();
  } catch (...) {
  }
}

int bixcor_timing_enable(int on) {
  API_BEGIN
  g_timing = on != 0;
  for (auto* c : g_ctx) timing_reset(c);
  return BIXCOR_OK;
  API_END
}

int bixcor_dev_set_async(int on) {
  API_BEGIN
  g_async = on != 0;
  return BIXCOR_OK;
  API_END
}

int bixcor_dev_set_stream(void* stream, int external) {
  API_BEGIN
  RC_TRY(ensure_init());
  Ctx* c = g_ctx[0];
  CU_TRY(cudaSetDevice(c->device));
  CU_TRY(cudaStreamSynchronize(c->compute));
  if (c->owns_compute && c->compute) cudaStreamDestroy(c->compute);
  if (external) {  // a NULL handle is the legacy default stream
    c->compute = (cudaStream_t)stream;
    c->owns_compute = false;
  } else {
    CU_TRY(cudaStreamCreateWithFlags(&c->compute, cudaStreamNonBlocking));
    c->owns_compute = true;
  }
  return BIXCOR_OK;
  API_END
}

int bixcor_dev_synchronize(void) {
  API_BEGIN
  RC_TRY(ensure_init());
  for (auto* c : g_ctx) {
    CU_TRY(cudaSetDevice(c->device));
    CU_TRY(cudaStreamSynchronize(c->compute));
  }
  return BIXCOR_OK;
  API_END
}

int bixcor_last_timing(double* what4) {
  API_BEGIN
  if (g_ctx.empty()) return fail(BIXCOR_ERR_INVALID, "not initialised");
  for (auto* c : g_ctx) {
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->compute);
    timing_collect(c);
  }
  // max over devices (they run concurrently)
  for (int i = 0; i < 4; ++i) what4[i] = 0;
  for (auto* c : g_ctx) {
    for (int i = 0; i < 3; ++i) what4[i] = std::max(what4[i], c->last_ms[i]);
    what4[3] = std::max(what4[3], (double)c->last_gemm_launches);
  }
  return BIXCOR_OK;
  API_END
}

// ---- packed helpers ---------------------------------------------------------
int64_t bixcor_packed_len(int64_t p, int shift) { return shift ? p * (p - 1) / 2 : p * (p + 1) / 2; }
int64_t bixcor_packed_offset(int64_t i, int64_t p, int shift) { return packed_row_offset(i, p, shift ? 1 : 0); }

int bixcor_upper_to_dense(const double* packed, int shift, int64_t p, double* out) {
  API_BEGIN
  if (!packed || !out || p < 0) return fail(BIXCOR_ERR_INVALID, "null pointer / bad p");
  const int s = shift ? 1 : 0;
  parallel_rows(p, [=](int64_t r0, int64_t r1) {
    for (int64_t i = r0; i < r1; ++i) {
      const double* row = packed + packed_row_offset(i, p, s);
      if (s) out[i + i * p] = 1.0;
      for (int64_t j = i + s; j < p; ++j) {
        const double v = row[j - i - s];
        out[i + j * p] = v;
        out[j + i * p] = v;
      }
    }
  });
  return BIXCOR_OK;
  API_END
}

int bixcor_dense_to_upper(const double* dense, int shift, int64_t p, double* out) {
  API_BEGIN
  if (!dense || !out || p < 0) return fail(BIXCOR_ERR_INVALID, "null pointer / bad p");
  const int s = shift ? 1 : 0;
  parallel_rows(p, [=](int64_t r0, int64_t r1) {
    for (int64_t i = r0; i < r1; ++i) {
      double* row = out + packed_row_offset(i, p, s);
      for (int64_t j = i + s; j < p; ++j) row[j - i - s] = dense[i + j * p];
    }
  });
  return BIXCOR_OK;
  API_END
}

int64_t bixcor_upper_to_sparse_nnz(const double* packed, int shift, int64_t p) {
  API_BEGIN
  if (!packed || p < 0) return -1;
  const int s = shift ? 1 : 0;
  std::vector<int64_t> part(64, 0);
  std::atomic<int> slot{0};
  parallel_rows(p, [&](int64_t r0, int64_t r1) {  // rows of the packed vector: every stored off-diagonal non-zero counts twice
    int64_t nnz = 0;
    for (int64_t i = r0; i < r1; ++i) {
      const double* row = packed + packed_row_offset(i, p, s);
      if (s) nnz += 1;  // unit diagonal
      for (int64_t j = i + s; j < p; ++j)
        if (row[j - i - s] != 0.0) nnz += (i == j) ? 1 : 2;
    }
    part[(size_t)slot.fetch_add(1) % part.size()] += nnz;  // (at most 16 workers: one slot each)
  });
  int64_t total = 0;
  for (int64_t v : part) total += v;
  return total;
  API_END_VAL(-1)
}
int bixcor_upper_to_sparse(const double* packed, int shift, int64_t p, int64_t* indptr, int32_t* indices, double* data) {
  API_BEGIN
  if (!packed || !indptr || p < 0) return fail(BIXCOR_ERR_INVALID, "null pointer / bad p");
  const int s = shift ? 1 : 0;
  // pass 1: non-zeros per column (threads over column blocks), prefix sum, pass 2: every column fills its own range
  std::vector<int64_t> cnt((size_t)p + 1, 0);
  parallel_cols(p, [&](int64_t c0, int64_t c1) {
    for (int64_t c = c0; c < c1; ++c) {
      int64_t n = 0;
      sym_column_nonzeros(packed, s, p, c, [&](int64_t, double) { ++n; });
      cnt[(size_t)c] = n;
    }
  });
  int64_t pos = 0;
  for (int64_t c = 0; c < p; ++c) {
    indptr[c] = pos;
    pos += cnt[(size_t)c];
  }
  indptr[p] = pos;
  if (pos > 0 && (!indices || !data)) return fail(BIXCOR_ERR_INVALID, "null pointer");
  parallel_cols(p, [&](int64_t c0, int64_t c1) {
    for (int64_t c = c0; c < c1; ++c) {
      int64_t w = indptr[c];
      sym_column_nonzeros(packed, s, p, c, [&](int64_t r, double v) {
        indices[w] = (int32_t)r;
        data[w++] = v;
      });
    }
  });
  return BIXCOR_OK;
  API_END
}

int bixcor_shard_rows(int64_t p, int shift, int world, int rank, int64_t* row_begin, int64_t* row_end) {
  API_BEGIN
  if (world < 1 || rank < 0 || rank >= world || p < 0 || !row_begin || !row_end)
    return fail(BIXCOR_ERR_INVALID, "bad shard arguments");
  shard_rows_impl(p, shift ? 1 : 0, world, rank, row_begin, row_end);
  return BIXCOR_OK;
  API_END
}

// ---- device-pointer entry points -------------------------------------------
#define DEV_PROLOGUE()                      \
  RC_TRY(ensure_init());                    \
  Ctx* c = g_ctx[0];                        \
  CU_TRY(cudaSetDevice(c->device));         \
  call_begin(c);

int bixcor_dev_rank_columns(const double* d_x, int64_t n, int64_t p, double* d_out) {
  API_BEGIN
  DEV_PROLOGUE();
  if (n < 1 || p < 1) return fail(BIXCOR_ERR_INVALID, "bad shape");
  ColumnOut co;
  memset(&co, 0, sizeof co);
  co.mode = COL_RANKS;
  co.f64 = d_out;
  co.ld64 = n;
  co.kpad = (int)n;
  {
    Timed tm(c, T_COLS);
    RC_TRY(launch_rank_kernel(c, d_x, n, p, co));
  }
  return finish_dev(c);
  API_END
}

int bixcor_dev_cor_upper_rows(const double* d_x, int64_t n, int64_t p, int spearman, int shift, double beta,
                              int precision, int64_t row_begin, int64_t row_end, double* d_out) {
  API_BEGIN
  DEV_PROLOGUE();
  RC_TRY(dev_cor_upper_rows(c, d_x, n, p, spearman, shift ? 1 : 0, beta, precision, row_begin, row_end, d_out, nullptr,
                            nullptr));
  return finish_dev(c);
  API_END
}

int64_t bixcor_dev_operand_row_bytes(int64_t n, int precision) {
  if (n < 1 || !exchangeable(precision)) return -1;
  return operand_row_bytes(n, precision);
}

int bixcor_dev_build_operand(const double* d_x, int64_t n, int64_t p, int spearman, int precision, int64_t gene_begin,
                             int64_t gene_end, void* d_operand, double* d_inv_norm) {
  API_BEGIN
  DEV_PROLOGUE();
  RC_TRY(build_operand(c, d_x, n, p, spearman, precision, gene_begin, gene_end, d_operand, d_inv_norm));
  return finish_dev(c);
  API_END
}

int bixcor_dev_operand_inv_norm(const void* d_operand, int64_t n, int64_t p, int precision, double* d_inv_norm) {
  API_BEGIN
  DEV_PROLOGUE();
  if (!d_operand || !d_inv_norm) return fail(BIXCOR_ERR_INVALID, "null pointer");
  if (precision != BIXCOR_PREC_I8EXACT) return finish_dev(c);  // the other operands are unit-norm already
  const int K = round_up((int)n, k_padding(precision));
  {
    Timed tm(c, T_COLS);
    inv_norm_from_digits_kernel<<<(unsigned)((p * 32 + 255) / 256), 256, 0, c->compute>>>((const int8_t*)d_operand, K, p,
                                                                                       d_inv_norm);
    g_launches++;
  }
  CU_TRY(cudaGetLastError());
  return finish_dev(c);
  API_END
}

int bixcor_dev_cor_upper_rows_op(const void* d_operand, const double* d_inv_norm, int64_t n, int64_t p, int precision,
                                 int shift, double beta, int64_t row_begin, int64_t row_end, double* d_out) {
  API_BEGIN
  DEV_PROLOGUE();
  RC_TRY(validate_rows(p, row_begin, row_end));
  if (!exchangeable(precision)) return fail(BIXCOR_ERR_INVALID, "operand exchange supports precisions 0..2 and 5 (got %d)", precision);
  if (row_begin < row_end) {
    Operand op;
    bind_operand(&op, n, p, precision, d_operand, d_inv_norm);
    RC_TRY(dev_cor_upper_rows_op(c, op, p, shift ? 1 : 0, beta, row_begin, row_end, d_out, nullptr, nullptr));
  }
  return finish_dev(c);
  API_END
}

int bixcor_dev_cor_upper_rows_op_to_host(const void* d_operand, const double* d_inv_norm, int64_t n, int64_t p,
                                         int precision, int shift, double beta, int64_t row_begin,
                                         int64_t row_end, double* h_out_slice) {
  API_BEGIN
  DEV_PROLOGUE();
  RC_TRY(validate_rows(p, row_begin, row_end));
  if (!d_operand || !h_out_slice) return fail(BIXCOR_ERR_INVALID, "null pointer");
  if (!exchangeable(precision)) return fail(BIXCOR_ERR_INVALID, "unknown precision %d", precision);
  if (row_begin == row_end) return BIXCOR_OK;
  shift = shift ? 1 : 0;
  Operand op;
  bind_operand(&op, n, p, precision, d_operand, d_inv_norm);
  const int64_t o0 = packed_row_offset(row_begin, p, shift), o1 = packed_row_offset(row_end, p, shift);
  RC_TRY(c->out.reserve(sizeof(double) * (size_t)std::max<int64_t>(1, o1 - o0)));
  D2HPipe pipe(c, h_out_slice);
  RC_TRY(dev_cor_upper_rows_op(c, op, p, shift, beta, row_begin, row_end, (double*)c->out.ptr, &pipe, h_out_slice));
  RC_TRY(finish(c));
  return pipe.finish();
  API_END
}

int bixcor_xchg_create(int64_t n, int64_t p, int precision, int world, int rank, void* handle_out) {
  API_BEGIN
  DEV_PROLOGUE();
  if (!handle_out) return fail(BIXCOR_ERR_INVALID, "null pointer");
  if (n < 2 || p < 1) return fail(BIXCOR_ERR_INVALID, "bad shape");
  if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world)
    return fail(BIXCOR_ERR_INVALID, "the operand exchange serves 1..%d ranks of one box (world %d, rank %d)", kMaxPeers, world, rank);
  precision = resolve_precision(precision, 1, n);  // AUTO: what a Spearman call resolves to (pass the precision explicitly for Pearson)
  if (!exchangeable(precision)) return fail(BIXCOR_ERR_INVALID, "precision %d has no per-gene operand record", precision);
  static_assert(sizeof(cudaIpcMemHandle_t) == BIXCOR_XCHG_HANDLE_BYTES, "handle size");
  xchg_free(c);
  Xchg& X = c->xchg;
  X.world = world;
  X.rank = rank;
  X.precision = precision;
  X.n = n;
  X.p = p;
  X.slab = ((p + world - 1) / world + kTile - 1) / kTile * kTile;  // equal 128-aligned gene slabs, the last one ragged
  X.rb = (size_t)operand_row_bytes(n, precision);
  auto up = [](size_t v) { return (v + 1023) & ~size_t(1023); };
  const size_t opb = up(X.rb * (size_t)X.slab * (size_t)world), invb = up(sizeof(double) * (size_t)X.slab * (size_t)world);
  X.op_off[0] = 0;
  X.op_off[1] = opb;
  X.inv_off[0] = 2 * opb;
  X.inv_off[1] = 2 * opb + invb;
  X.flag_off = 2 * opb + 2 * invb;
  X.total = X.flag_off + 64 * (size_t)kMaxPeers;
  void* ptr = nullptr;
  if (cudaMalloc(&ptr, X.total) != cudaSuccess) {
    cudaGetLastError();
    const size_t mb = X.total >> 20;
    X = Xchg();
    return fail(BIXCOR_ERR_OOM, "cannot allocate the %zu MB exchange block", mb);
  }
  X.local = (char*)ptr;
  X.peer[rank] = X.local;
  CU_TRY(cudaMemset(X.local, 0, X.total));
  CU_TRY(cudaDeviceSynchronize());
  CU_TRY(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)handle_out, X.local));
  return BIXCOR_OK;
  API_END
}

int bixcor_xchg_open(const void* handles) {
  API_BEGIN
  DEV_PROLOGUE();
  Xchg& X = c->xchg;
  if (!X.local) return fail(BIXCOR_ERR_INVALID, "bixcor_xchg_create comes first");
  if (!handles) return fail(BIXCOR_ERR_INVALID, "null pointer");
  if (X.opened) return BIXCOR_OK;
  for (int r = 0; r < X.world; ++r) {
    if (r == X.rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + (size_t)r * sizeof h, sizeof h);
    void* ptr = nullptr;
    CU_TRY(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    X.peer[r] = (char*)ptr;
  }
  X.opened = 1;
  return BIXCOR_OK;
  API_END
}

int bixcor_xchg_release_peers(void) {
  API_BEGIN
  if (g_ctx.empty()) return BIXCOR_OK;
  xchg_release_peers(g_ctx[0]);
  return BIXCOR_OK;
  API_END
}

int bixcor_xchg_destroy(void) {
  API_BEGIN
  if (g_ctx.empty()) return BIXCOR_OK;
  xchg_free(g_ctx[0]);
  return BIXCOR_OK;
  API_END
}

// d_out_slice (device) or h_out_slice (host) receives the packed rows; exactly one of them is non-null
static int cor_upper_rows_xchg(Ctx* c, const double* d_x, int x_is_slab, int64_t n, int64_t p, int spearman, int shift, double beta,
                               int64_t row_begin, int64_t row_end, double* d_out_slice, double* h_out_slice) {
  Xchg& X = c->xchg;
  if (!X.opened && X.world > 1) return fail(BIXCOR_ERR_INVALID, "bixcor_xchg_create / bixcor_xchg_open come first");
  if (!X.local || n != X.n || p != X.p) return fail(BIXCOR_ERR_INVALID, "shape differs from the one the exchange was created for");
  RC_TRY(validate_rows(p, row_begin, row_end));
  if (!d_x || (!d_out_slice && !h_out_slice)) return fail(BIXCOR_ERR_INVALID, "null pointer");
  RC_TRY(check_precision(X.precision, spearman, n));
  shift = shift ? 1 : 0;
  RC_TRY(xchg_build_and_sync(c, d_x, x_is_slab, spearman));  // every rank takes part, also one with an empty row range
  if (row_begin == row_end) return d_out_slice ? finish_dev(c) : finish(c);
  const int b = (int)(X.epoch & 1u);
  Operand op;
  bind_operand(&op, n, p, X.precision, X.local + X.op_off[b], (const double*)(X.local + X.inv_off[b]));
  if (d_out_slice) {
    RC_TRY(dev_cor_upper_rows_op(c, op, p, shift, beta, row_begin, row_end, d_out_slice, nullptr, nullptr));
    return finish_dev(c);
  }
  const int64_t o0 = packed_row_offset(row_begin, p, shift), o1 = packed_row_offset(row_end, p, shift);
  RC_TRY(c->out.reserve(sizeof(double) * (size_t)std::max<int64_t>(1, o1 - o0)));
  D2HPipe pipe(c, h_out_slice);
  RC_TRY(dev_cor_upper_rows_op(c, op, p, shift, beta, row_begin, row_end, (double*)c->out.ptr, &pipe, h_out_slice));
  RC_TRY(finish(c));
  return pipe.finish();
}

int bixcor_dev_cor_upper_rows_xchg(const double* d_x, int64_t n, int64_t p, int spearman, int shift, double beta,
                                   int64_t row_begin, int64_t row_end, double* d_out_slice) {
  API_BEGIN
  DEV_PROLOGUE();
  return cor_upper_rows_xchg(c, d_x, 0, n, p, spearman, shift, beta, row_begin, row_end, d_out_slice, nullptr);
  API_END
}

int bixcor_dev_cor_upper_rows_xchg_to_host(const double* d_x_slab, int64_t n, int64_t p, int spearman, int shift, double beta,
                                           int64_t row_begin, int64_t row_end, double* h_out_slice) {
  API_BEGIN
  DEV_PROLOGUE();
  return cor_upper_rows_xchg(c, d_x_slab, 1, n, p, spearman, shift, beta, row_begin, row_end, nullptr, h_out_slice);
  API_END
}

int bixcor_dev_cor_dense(const double* d_x, int64_t n, int64_t p, int spearman, int precision, double* d_out) {
  API_BEGIN
  DEV_PROLOGUE();
  RC_TRY(dev_affinity(c, d_x, n, p, spearman, 1.0, 1, precision, 0, p, d_out));
  return finish_dev(c);
  API_END
}

int bixcor_dev_diffcor_rows(const double* d_xa, int64_t na, const double* d_xb, int64_t nb, int64_t p, int spearman,
                            int precision, int64_t row_begin, int64_t row_end, double* d_r_a, double* d_r_b,
                            double* d_z, double* d_pval) {
  API_BEGIN
  DEV_PROLOGUE();
  RC_TRY(dev_diffcor_rows(c, d_xa, na, d_xb, nb, p, spearman, precision, row_begin, row_end, d_r_a, d_r_b, d_z, d_pval,
                          nullptr, nullptr, nullptr, nullptr, nullptr));
  return finish_dev(c);
  API_END
}

int bixcor_dev_affinity_cols(const double* d_x, int64_t n, int64_t p, int spearman, int signed_, double beta,
                             int precision, int64_t row_begin, int64_t row_end, double* d_a_pp) {
  API_BEGIN
  DEV_PROLOGUE();
  RC_TRY(dev_affinity(c, d_x, n, p, spearman, beta, signed_ ? 1 : 0, precision, row_begin, row_end, d_a_pp));
  return finish_dev(c);
  API_END
}

int bixcor_dev_tom_connectivity(const double* d_a_pp, int64_t p, int signed_, int64_t row_begin, int64_t row_end,
                                double* d_k) {
  API_BEGIN
  DEV_PROLOGUE();
  RC_TRY(dev_tom_connectivity(c, d_a_pp, p, signed_, row_begin, row_end, d_k));
  return finish_dev(c);
  API_END
}

int bixcor_dev_tom_rows(const double* d_a_pp, const double* d_k, int64_t p, int signed_, int version, int precision,
                        int packed, int64_t row_begin, int64_t row_end, double* d_out) {
  API_BEGIN
  DEV_PROLOGUE();
  RC_TRY(dev_tom_rows(c, d_a_pp, d_k, p, signed_, version, precision, packed, row_begin, row_end, d_out));
  return finish_dev(c);
  API_END
}

int64_t bixcor_dev_tom_plane_row_bytes(int64_t p, int precision) {
  if (p < 1 || !is_split_float(precision)) return -1;
  return (int64_t)round_up((int)p, k_padding(precision)) * (precision == BIXCOR_PREC_F16X3 ? 4 : 8);
}

int bixcor_dev_tom_split_planes(const double* d_a_pp, int64_t p, int precision, int64_t row_begin, int64_t row_end,
                                void* d_planes) {
  API_BEGIN
  DEV_PROLOGUE();
  if (!d_a_pp || !d_planes) return fail(BIXCOR_ERR_INVALID, "null pointer");
  RC_TRY(dev_split_planes(c, d_a_pp, p, precision, row_begin, row_end, d_planes));
  return finish_dev(c);
  API_END
}

int bixcor_dev_tom_rows_planes(const void* d_planes, const double* d_k, int64_t p, int signed_, int version, int precision,
                               int packed, int64_t row_begin, int64_t row_end, double* d_out) {
  API_BEGIN
  DEV_PROLOGUE();
  if (!d_planes || !d_k || !d_out) return fail(BIXCOR_ERR_INVALID, "null pointer");
  RC_TRY(dev_tom_rows(c, nullptr, d_k, p, signed_, version, precision, packed, row_begin, row_end, d_out, d_planes));
  return finish_dev(c);
  API_END
}

int bixcor_dev_sparse_gram_accumulate(const int64_t* d_indptr, const int32_t* d_indices, const float* d_data,
                                      int64_t cells, int64_t genes, int precision, double* d_gram_pp,
                                      double* d_colsum) {
  API_BEGIN
  DEV_PROLOGUE();
  RC_TRY(dev_sparse_accumulate(c, d_indptr, d_indices, d_data, cells, genes, precision, d_gram_pp, d_colsum));
  return finish_dev(c);
  API_END
}

int bixcor_dev_sparse_gram_finalise(const double* d_gram_pp, const double* d_colsum, int64_t total_cells,
                                    int64_t genes, double* d_out_packed) {
  API_BEGIN
  DEV_PROLOGUE();
  RC_TRY(dev_sparse_finalise(c, d_gram_pp, d_colsum, total_cells, genes, d_out_packed));
  return finish_dev(c);
  API_END
}

// ---- host-pointer entry points ---------------------------------------------
int bixcor_rank_columns(const double* x, int64_t n, int64_t p, double* out) {
  API_BEGIN
  RC_TRY(ensure_init());
  if (!x || !out) return fail(BIXCOR_ERR_INVALID, "null pointer");
  if (n == 0 || p == 0) return BIXCOR_OK;
  Ctx* c = g_ctx[0];
  CU_TRY(cudaSetDevice(c->device));
  const size_t bytes = sizeof(double) * (size_t)n * p;
  RC_TRY(c->x[0].reserve(bytes));
  RC_TRY(c->out.reserve(bytes));
  RC_TRY(h2d(c, c->x[0].ptr, x, bytes, c->compute));
  RC_TRY(bixcor_dev_rank_columns((const double*)c->x[0].ptr, n, p, (double*)c->out.ptr));
  RC_TRY(finish(c));
  D2HPipe pipe(c, out);
  RC_TRY(pipe.push(out, c->out.ptr, bytes, nullptr));
  return pipe.finish();
  API_END
}

static int host_shape_check(const void* x, int64_t n, int64_t p, const void* out) {
  if (!x || !out) return fail(BIXCOR_ERR_INVALID, "null pointer");
  if (n < 2) return fail(BIXCOR_ERR_INVALID, "need at least 2 samples (n = %lld)", (long long)n);
  if (p < 1) return fail(BIXCOR_ERR_INVALID, "need at least 1 gene (p = %lld)", (long long)p);
  if (p > 2000000) return fail(BIXCOR_ERR_UNSUPPORTED, "p too large");
  return BIXCOR_OK;
}

int bixcor_cor_upper_rows(const double* x, int64_t n, int64_t p, int spearman, int shift, double beta, int precision,
                          int64_t row_begin, int64_t row_end, double* out_slice) {
  API_BEGIN
  RC_TRY(ensure_init());
  RC_TRY(host_shape_check(x, n, p, out_slice));
  RC_TRY(validate_rows(p, row_begin, row_end));
  shift = shift ? 1 : 0;
  // split the requested rows over the process's devices (contiguous packed ranges)
  return for_each_device([&](Ctx* c, int d, int nd) -> int {
    // sub-shard [row_begin,row_end) by pair count: reuse the global partition on the remaining triangle
    int64_t r0 = row_begin, r1 = row_end;
    if (nd > 1) {
      int64_t q0, q1;
      shard_rows_impl(p - row_begin, shift, nd, d, &q0, &q1);
      r0 = row_begin + q0;
      r1 = std::min(row_end, row_begin + q1);
      if (d == nd - 1) r1 = row_end;
      if (r0 > r1) r0 = r1;
    }
    const int64_t base = packed_row_offset(row_begin, p, shift);
    const int64_t o0 = packed_row_offset(r0, p, shift);
    return host_cor_upper_pipelined(c, x, n, p, spearman, shift, beta, precision, r0, r1, out_slice + (o0 - base), nullptr);
  });
  API_END
}

int bixcor_cor_upper(const double* x, int64_t n, int64_t p, int spearman, int shift, double beta, int precision,
                     double* out_packed) {
  API_BEGIN
  return bixcor_cor_upper_rows(x, n, p, spearman, shift, beta, precision, 0, p, out_packed);
  API_END
}

int bixcor_cor_dense(const double* x, int64_t n, int64_t p, int spearman, int precision, double* out_pp) {
  API_BEGIN
  RC_TRY(ensure_init());
  RC_TRY(host_shape_check(x, n, p, out_pp));
  const size_t in_bytes = sizeof(double) * (size_t)n * p;
  if (g_ctx.size() > 1 && p >= 2 * kTile) {
    // several devices: every device ranks / standardises all genes (cheap) and computes the full-width strip of an equal
    // gene slab; column i of the column-major result is row i of the strip, so a slab is one contiguous host range
    return for_each_device([&](Ctx* c, int d, int nd) -> int {
      int64_t c0, c1;
      even_slab(p, nd, d, &c0, &c1);
      if (c1 <= c0) return BIXCOR_OK;
      RC_TRY(c->x[0].reserve(in_bytes));
      RC_TRY(c->out.reserve(sizeof(double) * (size_t)p * p));  // strip mode addresses the full matrix
      RC_TRY(h2d(c, c->x[0].ptr, x, in_bytes, c->compute));
      RC_TRY(dev_affinity(c, (const double*)c->x[0].ptr, n, p, spearman, 1.0, 1, precision, c0, c1, (double*)c->out.ptr));
      RC_TRY(finish(c));
      D2HPipe pipe(c);
      RC_TRY(pipe.push(out_pp + c0 * p, (double*)c->out.ptr + c0 * p, sizeof(double) * (size_t)(c1 - c0) * p, nullptr));
      return pipe.finish();
    });
  }
  Ctx* c = g_ctx[0];
  CU_TRY(cudaSetDevice(c->device));
  call_begin(c);
  const size_t out_bytes = sizeof(double) * (size_t)p * p;
  RC_TRY(c->x[0].reserve(in_bytes));
  RC_TRY(c->out.reserve(out_bytes));
  RC_TRY(h2d(c, c->x[0].ptr, x, in_bytes, c->compute));
  RC_TRY(dev_affinity(c, (const double*)c->x[0].ptr, n, p, spearman, 1.0, 1, precision, 0, p, (double*)c->out.ptr));
  RC_TRY(finish(c));
  D2HPipe pipe(c, out_pp);
  RC_TRY(pipe.push(out_pp, c->out.ptr, out_bytes, nullptr));
  return pipe.finish();
  API_END
}

int bixcor_diffcor_rows(const double* xa, int64_t na, const double* xb, int64_t nb, int64_t p, int spearman,
                        int precision, int64_t row_begin, int64_t row_end, double* r_a, double* r_b, double* z,
                        double* pval) {
  API_BEGIN
  RC_TRY(ensure_init());
  RC_TRY(host_shape_check(xa, na, p, r_a));
  RC_TRY(host_shape_check(xb, nb, p, r_b));
  if (!z || !pval) return fail(BIXCOR_ERR_INVALID, "null pointer");
  RC_TRY(validate_rows(p, row_begin, row_end));
  return for_each_device([&](Ctx* c, int d, int nd) -> int {
    int64_t r0 = row_begin, r1 = row_end;
    if (nd > 1) {
      int64_t q0, q1;
      shard_rows_impl(p - row_begin, 1, nd, d, &q0, &q1);
      r0 = row_begin + q0;
      r1 = std::min(row_end, row_begin + q1);
      if (d == nd - 1) r1 = row_end;
      if (r0 > r1) r0 = r1;
    }
    const int64_t base = packed_row_offset(row_begin, p, 1);
    const int64_t o0 = packed_row_offset(r0, p, 1), o1 = packed_row_offset(r1, p, 1);
    const size_t len = (size_t)std::max<int64_t>(1, o1 - o0);
    RC_TRY(c->x[0].reserve(sizeof(double) * (size_t)na * p));
    RC_TRY(c->x[1].reserve(sizeof(double) * (size_t)nb * p));
    RC_TRY(c->out.reserve(sizeof(double) * 4 * len));
    RC_TRY(h2d(c, c->x[0].ptr, xa, sizeof(double) * (size_t)na * p, c->compute));
    RC_TRY(h2d(c, c->x[1].ptr, xb, sizeof(double) * (size_t)nb * p, c->compute));
    double* d_out = (double*)c->out.ptr;
    D2HPipe pipe(c, r_a);
    const int64_t ho = o0 - base;
    RC_TRY(dev_diffcor_rows(c, (const double*)c->x[0].ptr, na, (const double*)c->x[1].ptr, nb, p, spearman, precision,
                            r0, r1, d_out, d_out + len, d_out + 2 * len, d_out + 3 * len, &pipe, r_a + ho, r_b + ho,
                            z + ho, pval + ho));
    RC_TRY(finish(c));
    return pipe.finish();
  });
  API_END
}

int bixcor_diffcor(const double* xa, int64_t na, const double* xb, int64_t nb, int64_t p, int spearman, int precision,
                   double* r_a, double* r_b, double* z, double* pval) {
  API_BEGIN
  return bixcor_diffcor_rows(xa, na, xb, nb, p, spearman, precision, 0, p, r_a, r_b, z, pval);
  API_END
}

static int tom_common(Ctx* c, double* d_a, int64_t p, int signed_, int version, int precision, int packed,
                      double* out) {
  RC_TRY(c->vec.reserve(sizeof(double) * (size_t)p * 4));
  double* d_k = (double*)c->vec.ptr + p;  // [0,p) is the diagonal scratch of dev_tom_rows
  RC_TRY(dev_tom_connectivity(c, d_a, p, signed_, 0, p, d_k));
  const size_t out_elems = packed ? (size_t)std::max<int64_t>(1, p * (p - 1) / 2) : (size_t)p * p;
  RC_TRY(c->out.reserve(sizeof(double) * out_elems));
  D2HPipe pipe(c, out);
  RC_TRY(dev_tom_rows(c, d_a, d_k, p, signed_, version, precision, packed, 0, p, (double*)c->out.ptr, nullptr, &pipe, out));
  RC_TRY(finish(c));
  return pipe.finish();
}

// TOM on every device of the process (bixcor_init(n), n > 1): the same three phases as the one-process-per-GPU
// driver (bixverse_b200/distributed.py), with cudaMemcpyPeer over NVLink for the slab exchange and the host for k.
//   phase 1  device d: affinity columns of gene slab d (from x, or uploaded from a_pp) + their connectivity
//   exchange NCCL (communicators of bixcor_init): grouped ncclBroadcast of every slab + ncclAllReduce(sum) of k in one
//            fused launch per device; without libnccl: peer copies of the slabs, k through the host
//   phase 3  device d: TOM rows of its share (pair-balanced when packed, equal strips when dense) -> caller's buffer
static int tom_multi_device(const double* x, int64_t n, const double* a_pp, int64_t p, int spearman, int signed_,
                            int version, double beta, int precision, int packed, double* out) {
  const int nd_all = (int)g_ctx.size();
  PhaseBarrier bar(nd_all);
  const bool use_nccl = !g_comms.empty();
  std::vector<double> k_host(use_nccl ? 0 : (size_t)p, 0.0);
  const int cor_prec = tom_cor_precision(precision, spearman, n);
  const int tom_prec = resolve_precision_float(precision);
  // split-float kinds: every device splits its own slab into the hi / lo operand planes and the PLANES are exchanged
  // (f16: half the bytes of the f64 slabs, and no device converts the whole matrix); phase 3 runs from the planes alone
  const bool planes_mode = use_nccl && is_split_float(tom_prec);
  const size_t plane_rb = planes_mode ? (size_t)round_up((int)p, k_padding(tom_prec)) * (tom_prec == BIXCOR_PREC_F16X3 ? 4 : 8) : 0;
  return for_each_device([&](Ctx* c, int d, int nd) -> int {
    int64_t c0, c1;
    even_slab(p, nd, d, &c0, &c1);
    double* d_a = nullptr;
    double* d_k = nullptr;
    int rc = [&]() -> int {
      RC_TRY(c->out2.reserve(sizeof(double) * (size_t)p * p));
      RC_TRY(c->vec.reserve(sizeof(double) * (size_t)p * 4));
      d_a = (double*)c->out2.ptr;
      d_k = (double*)c->vec.ptr + p;
      if (use_nccl) CU_TRY(cudaMemsetAsync(d_k, 0, sizeof(double) * (size_t)p, c->compute));  // all-reduce(sum) of slab-wise k
      if (c1 > c0) {
        if (x) {
          RC_TRY(c->x[0].reserve(sizeof(double) * (size_t)n * p));
          RC_TRY(h2d(c, c->x[0].ptr, x, sizeof(double) * (size_t)n * p, c->compute));
          RC_TRY(dev_affinity(c, (const double*)c->x[0].ptr, n, p, spearman, beta, signed_ ? 1 : 0, cor_prec, c0, c1, d_a));
        } else {
          RC_TRY(h2d(c, d_a + c0 * p, a_pp + c0 * p, sizeof(double) * (size_t)(c1 - c0) * p, c->compute));
        }
        RC_TRY(dev_tom_connectivity(c, d_a, p, signed_, c0, c1, d_k));
        if (!use_nccl) {
          RC_TRY(finish(c));
          CU_TRY(cudaMemcpy(k_host.data() + c0, d_k + c0, sizeof(double) * (size_t)(c1 - c0), cudaMemcpyDeviceToHost));
        }
      }
      if (planes_mode) {
        RC_TRY(c->op[1].reserve(plane_rb * (size_t)p));
        RC_TRY(dev_split_planes(c, d_a, p, tom_prec, c0, c1, c->op[1].ptr));
      }
      return use_nccl ? BIXCOR_OK : finish(c);
    }();
    rc = bar.arrive(rc);  // every device reaches the exchange or none does (a collective with a missing rank would hang)
    if (rc != BIXCOR_OK) return rc;
    rc = [&]() -> int {
      if (use_nccl) {
        // one fused NCCL launch per device: every slab is broadcast from its owner over NVLink / NVSwitch and the
        // connectivity vector is all-reduced (sum of slab-wise contributions, zero elsewhere)
        NCCL_TRY(g_nccl.GroupStart());
        for (int s_ = 0; s_ < nd; ++s_) {
          int64_t s0, s1;
          even_slab(p, nd, s_, &s0, &s1);
          if (s1 <= s0) continue;
          if (planes_mode) {
            char* pl = (char*)c->op[1].ptr + (size_t)s0 * plane_rb;
            NCCL_TRY(g_nccl.Broadcast(pl, pl, (size_t)(s1 - s0) * plane_rb, ncclChar, s_, g_comms[d], c->compute));
          } else {
            NCCL_TRY(g_nccl.Broadcast(d_a + s0 * p, d_a + s0 * p, (size_t)(s1 - s0) * p, ncclDouble, s_, g_comms[d], c->compute));
          }
        }
        NCCL_TRY(g_nccl.AllReduce(d_k, d_k, (size_t)p, ncclDouble, ncclSum, g_comms[d], c->compute));
        NCCL_TRY(g_nccl.GroupEnd());
        return finish(c);
      }
      for (int s_ = 0; s_ < nd; ++s_) {
        if (s_ == d) continue;
        int64_t s0, s1;
        even_slab(p, nd, s_, &s0, &s1);
        if (s1 > s0)
          CU_TRY(cudaMemcpyPeerAsync(d_a + s0 * p, c->device, (const double*)g_ctx[s_]->out2.ptr + s0 * p, g_ctx[s_]->device,
                                     sizeof(double) * (size_t)(s1 - s0) * p, c->compute));
      }
      CU_TRY(cudaMemcpyAsync(d_k, k_host.data(), sizeof(double) * (size_t)p, cudaMemcpyHostToDevice, c->compute));
      return finish(c);
    }();
    rc = bar.arrive(rc);  // nobody reuses a slab buffer while a peer may still be reading it
    if (rc != BIXCOR_OK) return rc;
    int64_t r0, r1;
    if (packed) shard_rows_impl(p, 1, nd, d, &r0, &r1);
    else even_slab(p, nd, d, &r0, &r1);
    if (r1 <= r0) return BIXCOR_OK;
    if (packed) {
      const int64_t o0 = packed_row_offset(r0, p, 1), o1 = packed_row_offset(r1, p, 1);
      RC_TRY(c->out.reserve(sizeof(double) * (size_t)std::max<int64_t>(1, o1 - o0)));
      D2HPipe pipe(c, out);
      RC_TRY(dev_tom_rows(c, planes_mode ? nullptr : d_a, d_k, p, signed_, version, precision, 1, r0, r1, (double*)c->out.ptr,
                          planes_mode ? c->op[1].ptr : nullptr, &pipe, out + o0));
      RC_TRY(finish(c));
      return pipe.finish();
    }
    RC_TRY(c->out.reserve(sizeof(double) * (size_t)p * p));  // strip mode addresses the full matrix
    D2HPipe pipe(c, out);
    RC_TRY(dev_tom_rows(c, planes_mode ? nullptr : d_a, d_k, p, signed_, version, precision, 0, r0, r1, (double*)c->out.ptr,
                        planes_mode ? c->op[1].ptr : nullptr, &pipe, out));
    RC_TRY(finish(c));
    return pipe.finish();
  });
}

int bixcor_tom(const double* a_pp, int64_t p, int signed_, int version, int precision, double* out_pp) {
  API_BEGIN
  RC_TRY(ensure_init());
  if (!a_pp || !out_pp) return fail(BIXCOR_ERR_INVALID, "null pointer");
  if (p < 1) return fail(BIXCOR_ERR_INVALID, "p must be >= 1");
  if (version != BIXCOR_TOM_V1 && version != BIXCOR_TOM_V2)
    return fail(BIXCOR_ERR_INVALID, "Invalid TOM version type: %d", version);
  if (g_ctx.size() > 1 && p >= 2 * kTile)
    return tom_multi_device(nullptr, 0, a_pp, p, 0, signed_, version, 1.0, precision, 0, out_pp);
  Ctx* c = g_ctx[0];
  CU_TRY(cudaSetDevice(c->device));
  call_begin(c);
  const size_t bytes = sizeof(double) * (size_t)p * p;
  RC_TRY(c->out2.reserve(bytes));
  RC_TRY(h2d(c, c->out2.ptr, a_pp, bytes, c->compute));
  return tom_common(c, (double*)c->out2.ptr, p, signed_, version, precision, 0, out_pp);
  API_END
}

// cor_module_tom in one call (R/methods_coexp_cor.R:118-164: get_sym_matrix -> rs_tom -> rs_dense_to_upper_triangle):
// packed correlation in, packed (shift = 1) TOM out; the dense matrices only ever exist in HBM.
int bixcor_tom_packed(const double* packed_in, int shift_in, int64_t p, int signed_, int version, int precision,
                      double* out_packed) {
  API_BEGIN
  RC_TRY(ensure_init());
  if (!packed_in || !out_packed) return fail(BIXCOR_ERR_INVALID, "null pointer");
  if (p < 2) return fail(BIXCOR_ERR_INVALID, "p must be >= 2");
  if (version != BIXCOR_TOM_V1 && version != BIXCOR_TOM_V2)
    return fail(BIXCOR_ERR_INVALID, "Invalid TOM version type: %d", version);
  Ctx* c = g_ctx[0];
  CU_TRY(cudaSetDevice(c->device));
  call_begin(c);
  const int s = shift_in ? 1 : 0;
  const size_t in_bytes = sizeof(double) * (size_t)(s ? p * (p - 1) / 2 : p * (p + 1) / 2);
  RC_TRY(c->out.reserve(std::max(in_bytes, sizeof(double) * (size_t)(p * (p - 1) / 2))));
  RC_TRY(c->out2.reserve(sizeof(double) * (size_t)p * p));
  RC_TRY(h2d(c, c->out.ptr, packed_in, in_bytes, c->compute));
  {
    Timed tm(c, T_OTHER);
    unpack_upper_rows_kernel<<<(unsigned)p, 256, 0, c->compute>>>((const double*)c->out.ptr, p, s, (double*)c->out2.ptr);
    const unsigned nb = (unsigned)((p + 31) / 32);
    mirror_upper_kernel<<<dim3(nb, nb), 256, 0, c->compute>>>((double*)c->out2.ptr, p);
    g_launches += 2;
  }
  CU_TRY(cudaGetLastError());
  return tom_common(c, (double*)c->out2.ptr, p, signed_, version, precision, 1, out_packed);
  API_END
}

int bixcor_tom_from_exp(const double* x, int64_t n, int64_t p, int spearman, int signed_, int version, double beta,
                        int precision, int packed, double* out) {
  API_BEGIN
  RC_TRY(ensure_init());
  RC_TRY(host_shape_check(x, n, p, out));
  if (version != BIXCOR_TOM_V1 && version != BIXCOR_TOM_V2)
    return fail(BIXCOR_ERR_INVALID, "Invalid TOM version type: %d", version);
  if (g_ctx.size() > 1 && p >= 2 * kTile)
    return tom_multi_device(x, n, nullptr, p, spearman, signed_, version, beta, precision, packed, out);
  Ctx* c = g_ctx[0];
  CU_TRY(cudaSetDevice(c->device));
  call_begin(c);
  const size_t in_bytes = sizeof(double) * (size_t)n * p;
  RC_TRY(c->x[0].reserve(in_bytes));
  RC_TRY(c->out2.reserve(sizeof(double) * (size_t)p * p));
  RC_TRY(h2d(c, c->x[0].ptr, x, in_bytes, c->compute));
  // the correlation feeding the TOM is always the exact / requested-precision Gram
  const int cor_prec = tom_cor_precision(precision, spearman, n);
  RC_TRY(dev_affinity(c, (const double*)c->x[0].ptr, n, p, spearman, beta, signed_ ? 1 : 0, cor_prec, 0, p,
                      (double*)c->out2.ptr));
  return tom_common(c, (double*)c->out2.ptr, p, signed_, version, precision, packed, out);
  API_END
}

// ---- sibling Gram kernels ("next" rows of SURVEY 8f) ---------------------------------------------
static int dense_symmetric(const double* x, int64_t n, int64_t p, int norm_mode, double scale, int diag_one,
                           double* out_pp) {
  RC_TRY(ensure_init());
  RC_TRY(host_shape_check(x, n, p, out_pp));
  Ctx* c = g_ctx[0];
  CU_TRY(cudaSetDevice(c->device));
  call_begin(c);
  const size_t in_bytes = sizeof(double) * (size_t)n * p, out_bytes = sizeof(double) * (size_t)p * p;
  RC_TRY(c->x[0].reserve(in_bytes));
  RC_TRY(c->out.reserve(out_bytes));
  RC_TRY(h2d(c, c->x[0].ptr, x, in_bytes, c->compute));
  Operand op;
  RC_TRY(prepare_operand(c, (const double*)c->x[0].ptr, n, p, 0, BIXCOR_PREC_F64, 0, &op, norm_mode));
  TileList* tl;
  RC_TRY(get_tiles(c, p, 0, p, 1, 8, 0, &tl));
  EpiParams e = make_epi(p);
  e.out = (double*)c->out.ptr;
  e.ldo = p;
  e.mirror = 1;
  e.diag_one = diag_one;
  e.scale = scale;
  RC_TRY(launch_gram(c, op, nullptr, p, tl->d_tiles, tl->ntiles, EPI_DENSE, e));
  RC_TRY(finish(c));
  D2HPipe pipe(c, out_pp);
  RC_TRY(pipe.push(out_pp, c->out.ptr, out_bytes, nullptr));
  return pipe.finish();
}

int bixcor_covariance(const double* x, int64_t n, int64_t p, double* out_pp) {
  API_BEGIN
  return dense_symmetric(x, n, p, 1, 1.0 / (double)(n - 1), 0, out_pp);
  API_END
}

int bixcor_cos(const double* x, int64_t n, int64_t p, double* out_pp) {
  API_BEGIN
  return dense_symmetric(x, n, p, 2, 1.0, 0, out_pp);
  API_END
}

// cor2 on the device: leaves the column-major p x q matrix cor(x_i, y_j) in c->out
static int dev_cor2(Ctx* c, const double* x, int64_t n, int64_t p, const double* y, int64_t q, int spearman) {
  RC_TRY(c->x[0].reserve(sizeof(double) * (size_t)n * p));
  RC_TRY(c->x[1].reserve(sizeof(double) * (size_t)n * q));
  RC_TRY(c->out.reserve(sizeof(double) * (size_t)p * q));
  RC_TRY(h2d(c, c->x[0].ptr, x, sizeof(double) * (size_t)n * p, c->compute));
  RC_TRY(h2d(c, c->x[1].ptr, y, sizeof(double) * (size_t)n * q, c->compute));
  Operand ox, oy;
  RC_TRY(prepare_operand(c, (const double*)c->x[0].ptr, n, p, spearman, BIXCOR_PREC_F64, 0, &ox));
  RC_TRY(prepare_operand(c, (const double*)c->x[1].ptr, n, q, spearman, BIXCOR_PREC_F64, 1, &oy));
  // rows of the kernel = genes of y, cols = genes of x: out[iy * p + jx] is the column-major p x q result
  std::vector<int2> tiles;
  for (int bi = 0; bi < (int)((q + kTile - 1) / kTile); ++bi)
    for (int bj = 0; bj < (int)((p + kTile - 1) / kTile); ++bj) tiles.push_back(make_int2(bi, bj));
  RC_TRY(c->vec.reserve(sizeof(int2) * tiles.size() + 64));
  CU_TRY(cudaMemcpyAsync(c->vec.ptr, tiles.data(), sizeof(int2) * tiles.size(), cudaMemcpyHostToDevice, c->compute));
  CU_TRY(cudaStreamSynchronize(c->compute));  // `tiles` is a stack-owned staging vector
  EpiParams e = make_epi(q);
  e.out = (double*)c->out.ptr;
  e.ldo = p;
  return launch_gram(c, oy, nullptr, q, (const int2*)c->vec.ptr, (int)tiles.size(), EPI_DENSE, e, &ox);
}

int bixcor_cor2(const double* x, int64_t n, int64_t p, const double* y, int64_t q, int spearman, double* out_pq) {
  API_BEGIN
  RC_TRY(ensure_init());
  RC_TRY(host_shape_check(x, n, p, out_pq));
  RC_TRY(host_shape_check(y, n, q, out_pq));
  Ctx* c = g_ctx[0];
  CU_TRY(cudaSetDevice(c->device));
  call_begin(c);
  RC_TRY(dev_cor2(c, x, n, p, y, q, spearman));
  RC_TRY(finish(c));
  D2HPipe pipe(c, out_pq);
  RC_TRY(pipe.push(out_pq, c->out.ptr, sizeof(double) * (size_t)p * q, nullptr));
  return pipe.finish();
  API_END
}

// rs_rbh_cor for one (origin, target) pair of module matrices already restricted to their shared features
// (src/rust/src/methods/r_rbh.rs:187-272; crate calculate_rbh_cor): |cor2| on the device, k-th best value of every
// row and column, reciprocal hits above min_similarity compacted origin-major into the caller's arrays (0-based).
int bixcor_rbh_cor(const double* x, int64_t n, int64_t p, const double* y, int64_t q, int k_best, int spearman,
                   double min_similarity, int32_t* out_origin, int32_t* out_target, double* out_sim, int64_t capacity,
                   int64_t* n_hits) {
  API_BEGIN
  RC_TRY(ensure_init());
  if (!n_hits) return fail(BIXCOR_ERR_INVALID, "null pointer");
  *n_hits = 0;
  RC_TRY(host_shape_check(x, n, p, n_hits));
  RC_TRY(host_shape_check(y, n, q, n_hits));
  if (k_best < 1) return fail(BIXCOR_ERR_INVALID, "k_best must be >= 1");
  if (capacity > 0 && (!out_origin || !out_target || !out_sim)) return fail(BIXCOR_ERR_INVALID, "null pointer");
  Ctx* c = g_ctx[0];
  CU_TRY(cudaSetDevice(c->device));
  call_begin(c);
  RC_TRY(dev_cor2(c, x, n, p, y, q, spearman));
  const size_t pq = (size_t)p * q;
  RC_TRY(c->out2.reserve(sizeof(double) * (size_t)(p + q) + pq));
  double* thr = (double*)c->out2.ptr;
  unsigned char* hit = (unsigned char*)(thr + p + q);
  {
    Timed tm(c, T_OTHER);
    rbh_kth_best_kernel<<<(unsigned)((p + q + 7) / 8), 256, 0, c->compute>>>((const double*)c->out.ptr, p, q, k_best, thr, thr + p);
    rbh_mark_kernel<<<(unsigned)((pq + 255) / 256), 256, 0, c->compute>>>((double*)c->out.ptr, p, q, thr, thr + p, min_similarity, hit);
    g_launches += 2;
  }
  CU_TRY(cudaGetLastError());
  std::vector<double> sim(pq);
  std::vector<unsigned char> mask(pq);
  CU_TRY(cudaMemcpyAsync(sim.data(), c->out.ptr, sizeof(double) * pq, cudaMemcpyDeviceToHost, c->compute));
  CU_TRY(cudaMemcpyAsync(mask.data(), hit, pq, cudaMemcpyDeviceToHost, c->compute));
  RC_TRY(finish(c));
  int64_t cnt = 0;
  for (int64_t i = 0; i < p; ++i)
    for (int64_t j = 0; j < q; ++j)
      if (mask[(size_t)(i + j * p)]) {
        if (cnt < capacity) {
          out_origin[cnt] = (int32_t)i;
          out_target[cnt] = (int32_t)j;
          out_sim[cnt] = sim[(size_t)(i + j * p)];
        }
        ++cnt;
      }
  *n_hits = cnt;
  if (cnt > capacity) return fail(BIXCOR_ERR_INVALID, "rbh: %lld hits exceed the capacity %lld", (long long)cnt, (long long)capacity);
  return BIXCOR_OK;
  API_END
}

int bixcor_cov2cor(const double* cov_pp, int64_t p, double* out_pp) {
  API_BEGIN
  RC_TRY(ensure_init());
  if (!cov_pp || !out_pp) return fail(BIXCOR_ERR_INVALID, "null pointer");
  if (p < 1) return fail(BIXCOR_ERR_INVALID, "p must be >= 1");
  Ctx* c = g_ctx[0];
  CU_TRY(cudaSetDevice(c->device));
  call_begin(c);
  const size_t bytes = sizeof(double) * (size_t)p * p;
  RC_TRY(c->out2.reserve(bytes));
  RC_TRY(c->out.reserve(bytes));
  RC_TRY(h2d(c, c->out2.ptr, cov_pp, bytes, c->compute));
  {
    Timed tm(c, T_OTHER);
    dim3 grid((unsigned)((p + 255) / 256), (unsigned)p);
    cov2cor_kernel<<<grid, 256, 0, c->compute>>>((const double*)c->out2.ptr, p, (double*)c->out.ptr);
    g_launches++;
  }
  CU_TRY(cudaGetLastError());
  RC_TRY(finish(c));
  D2HPipe pipe(c, out_pp);
  RC_TRY(pipe.push(out_pp, c->out.ptr, bytes, nullptr));
  return pipe.finish();
  API_END
}

// ---- RBF affinities and CoReMo stability ("next" rows 1-2 of SURVEY 8f) --------------------------
int bixcor_rbf(const double* x, int64_t len, double epsilon, int rbf_type, double* out) {
  API_BEGIN
  RC_TRY(ensure_init());
  if (len < 0 || (len > 0 && (!x || !out))) return fail(BIXCOR_ERR_INVALID, "null pointer / bad length");
  RC_TRY(check_rbf(rbf_type, epsilon));
  if (len == 0) return BIXCOR_OK;
  Ctx* c = g_ctx[0];
  CU_TRY(cudaSetDevice(c->device));
  call_begin(c);
  const size_t bytes = sizeof(double) * (size_t)len;
  RC_TRY(c->out2.reserve(bytes));
  RC_TRY(c->out.reserve(bytes));
  RC_TRY(h2d(c, c->out2.ptr, x, bytes, c->compute));
  {
    Timed tm(c, T_OTHER);
    const unsigned grid = (unsigned)std::min<int64_t>((len + 255) / 256, (int64_t)c->sm_count * 16);
    rbf_elementwise_kernel<<<grid, 256, 0, c->compute>>>((const double*)c->out2.ptr, len, rbf_type, epsilon,
                                                          (double*)c->out.ptr);
    g_launches++;
  }
  CU_TRY(cudaGetLastError());
  RC_TRY(finish(c));
  D2HPipe pipe(c, out);
  RC_TRY(pipe.push(out, c->out.ptr, bytes, nullptr));
  return pipe.finish();
  API_END
}

int bixcor_rbf_iterate_epsilons(const double* dist_packed, int64_t p, int shift, const double* epsilons, int n_eps,
                                int rbf_type, double* out_p_x_neps) {
  API_BEGIN
  RC_TRY(ensure_init());
  if (!dist_packed || !epsilons || !out_p_x_neps) return fail(BIXCOR_ERR_INVALID, "null pointer");
  if (p < 1 || n_eps < 1) return fail(BIXCOR_ERR_INVALID, "bad shape");
  for (int e = 0; e < n_eps; ++e) RC_TRY(check_rbf(rbf_type, epsilons[e]));
  shift = shift ? 1 : 0;
  Ctx* c = g_ctx[0];
  CU_TRY(cudaSetDevice(c->device));
  call_begin(c);
  const int64_t len = shift ? p * (p - 1) / 2 : p * (p + 1) / 2;
  const int rows_per_chunk = 512;
  const int nchunks = (int)((p + rows_per_chunk - 1) / rows_per_chunk);
  RC_TRY(c->out2.reserve(sizeof(double) * (size_t)std::max<int64_t>(1, len)));
  RC_TRY(c->out.reserve(sizeof(double) * (size_t)p * n_eps));
  RC_TRY(c->panel.reserve(sizeof(double) * (size_t)p * kRbfMaxEps * nchunks));
  RC_TRY(h2d(c, c->out2.ptr, dist_packed, sizeof(double) * (size_t)len, c->compute));
  for (int e0 = 0; e0 < n_eps; e0 += kRbfMaxEps) {
    RbfEps E;
    E.n = std::min(kRbfMaxEps, n_eps - e0);
    for (int e = 0; e < kRbfMaxEps; ++e) E.eps[e] = e < E.n ? epsilons[e0 + e] : 1.0;
    double* k = (double*)c->out.ptr + (size_t)e0 * p;
    Timed tm(c, T_OTHER);
    rbf_rowsum_kernel<<<(unsigned)((p + 7) / 8), 256, 0, c->compute>>>((const double*)c->out2.ptr, p, shift, rbf_type, E, k);
    dim3 gc((unsigned)((p + 127) / 128), (unsigned)nchunks);
    rbf_colsum_kernel<<<gc, 128, 0, c->compute>>>((const double*)c->out2.ptr, p, shift, rbf_type, E, rows_per_chunk,
                                                  (double*)c->panel.ptr);
    dim3 gr((unsigned)((p + 255) / 256), (unsigned)E.n);
    rbf_reduce_kernel<<<gr, 256, 0, c->compute>>>((const double*)c->panel.ptr, p, E.n, nchunks, k);
    g_launches += 3;
  }
  CU_TRY(cudaGetLastError());
  RC_TRY(finish(c));
  D2HPipe pipe(c, out_p_x_neps);
  RC_TRY(pipe.push(out_p_x_neps, c->out.ptr, sizeof(double) * (size_t)p * n_eps, nullptr));
  return pipe.finish();
  API_END
}

int bixcor_cor_upper_rbf(const double* x, int64_t n, int64_t p, int spearman, int shift, double epsilon, int rbf_type,
                         int complement, int precision, double* out_packed) {
  API_BEGIN
  RC_TRY(ensure_init());
  RC_TRY(host_shape_check(x, n, p, out_packed));
  RC_TRY(check_rbf(rbf_type, epsilon));
  Ctx* c = g_ctx[0];
  CU_TRY(cudaSetDevice(c->device));
  call_begin(c);
  shift = shift ? 1 : 0;
  RbfSpec rbf;
  rbf.mode = rbf_type;
  rbf.eps = epsilon;
  rbf.complement = complement ? 1 : 0;
  return host_cor_upper_pipelined(c, x, n, p, spearman, shift, 1.0, precision, 0, p, out_packed, &rbf);
  API_END
}

int bixcor_coremo_stability(const double* x, int64_t n, int64_t p, const int32_t* indices_1based, int n_idx,
                            double epsilon, int rbf_type, int spearman, int precision, double* const* outs) {
  API_BEGIN
  RC_TRY(ensure_init());
  if (!x || !indices_1based || !outs) return fail(BIXCOR_ERR_INVALID, "null pointer");
  if (n < 3) return fail(BIXCOR_ERR_INVALID, "need at least 3 samples to leave one out (n = %lld)", (long long)n);
  if (p < 1 || n_idx < 0) return fail(BIXCOR_ERR_INVALID, "bad shape");
  RC_TRY(check_rbf(rbf_type, epsilon));
  for (int k = 0; k < n_idx; ++k) {
    if (indices_1based[k] < 1 || indices_1based[k] > n)
      return fail(BIXCOR_ERR_INVALID, "sample index %d out of range 1..%lld", indices_1based[k], (long long)n);
    if (!outs[k]) return fail(BIXCOR_ERR_INVALID, "null output vector %d", k);
  }
  const int64_t len = p * (p - 1) / 2;
  // the left-out samples are independent jobs: with several devices, device d takes samples d, d + nd, ...
  // (the reference runs them on a rayon pool, r_coremo.rs:204-226)
  return for_each_device([&](Ctx* c, int d, int nd) -> int {
    if (d >= n_idx) return BIXCOR_OK;
    RC_TRY(c->x[0].reserve(sizeof(double) * (size_t)n * p));
    RC_TRY(c->x[1].reserve(sizeof(double) * (size_t)(n - 1) * p));
    RC_TRY(h2d(c, c->x[0].ptr, x, sizeof(double) * (size_t)n * p, c->compute));
    RbfSpec rbf;
    rbf.mode = rbf_type;
    rbf.eps = epsilon;
    rbf.complement = 1;  // 1 - phi(1 - |r|)
    // two result buffers: the D2H of one sample overlaps the kernels of the next
    Buffer* res[2] = {&c->out, &c->out2};
    for (auto* b : res) RC_TRY(b->reserve(sizeof(double) * (size_t)std::max<int64_t>(1, len)));
    D2HPipe pipe(c);
    cudaEvent_t reuse[2] = {nullptr, nullptr};
    int b = 0;
    for (int k = d; k < n_idx; k += nd, b ^= 1) {
      if (reuse[b]) CU_TRY(cudaStreamWaitEvent(c->compute, reuse[b], 0));  // its previous D2H has left the buffer
      {
        Timed tm(c, T_OTHER);
        remove_row_kernel<<<(unsigned)p, 128, 0, c->compute>>>((const double*)c->x[0].ptr, n, indices_1based[k] - 1,
                                                               (double*)c->x[1].ptr);
        g_launches++;
      }
      RC_TRY(dev_cor_upper_rows(c, (const double*)c->x[1].ptr, n - 1, p, spearman, 1, 1.0, precision, 0, p,
                                (double*)res[b]->ptr, nullptr, nullptr, &rbf));
      cudaEvent_t ev = ctx_event(c);
      CU_TRY(cudaEventRecord(ev, c->compute));
      RC_TRY(pipe.push(outs[k], res[b]->ptr, sizeof(double) * (size_t)len, ev));
      reuse[b] = ctx_event(c);
      CU_TRY(cudaEventRecord(reuse[b], c->copy_out));  // pinned: the DMA itself; pageable: the DMAs into the staging ring
    }
    RC_TRY(finish(c));
    return pipe.finish();
  });
  API_END
}

// rs_pairwise_gene_cors_mc (src/rust/src/meta_cell/r_mc_processing.rs:258-283)
// cells_to_keep == nullptr: every cell; else the correlation runs over the n_keep listed cells (0-based, unique)
static int sparse_pairwise_impl(const int64_t* indptr, const int32_t* indices, const float* data, int64_t cells,
                                int64_t genes, int is_csr, const int32_t* cells_to_keep, int64_t n_keep,
                                const int32_t* gene_idx1, const int32_t* gene_idx2, int64_t n_pairs, int spearman,
                                double* out) {
  RC_TRY(ensure_init());
  if (!indptr || n_pairs < 0 || (n_pairs > 0 && (!gene_idx1 || !gene_idx2 || !out)))
    return fail(BIXCOR_ERR_INVALID, "null pointer / bad pair count");
  if (cells < 2 || genes < 1) return fail(BIXCOR_ERR_INVALID, "bad shape");
  if (n_pairs == 0) return BIXCOR_OK;
  std::vector<int32_t> cell_pos;
  const int64_t all_cells = cells;  // extent of the sparse arrays; `cells` below = cells the correlation runs over
  if (cells_to_keep) {
    if (n_keep < 2) return fail(BIXCOR_ERR_INVALID, "cells_to_keep needs at least 2 cells");
    cell_pos.assign((size_t)all_cells, -1);
    for (int64_t k = 0; k < n_keep; ++k) {
      const int32_t ci = cells_to_keep[k];
      if (ci < 0 || ci >= all_cells) return fail(BIXCOR_ERR_INVALID, "cells_to_keep[%lld] out of range (0-based)", (long long)k);
      if (cell_pos[ci] >= 0) return fail(BIXCOR_ERR_INVALID, "cells_to_keep[%lld] is a duplicate", (long long)k);
      cell_pos[ci] = (int32_t)k;
    }
    cells = n_keep;
  }
  const int64_t nptr = (is_csr ? all_cells : genes) + 1;
  const int64_t nnz = indptr[nptr - 1];
  if (nnz < 0 || (nnz > 0 && (!indices || !data))) return fail(BIXCOR_ERR_INVALID, "bad sparse arrays");
  // Spearman beyond the 16384-cell limit of the dense rank kernel: ranks of the non-zeros only (rank_sparse_columns_kernel)
  const bool sparse_ranks = spearman && cells > 16384;
  // dense slots for the genes that actually occur
  std::vector<int32_t> slot_of((size_t)genes, -1), gene_of_slot, s1((size_t)n_pairs), s2((size_t)n_pairs);
  for (int64_t k = 0; k < n_pairs; ++k) {
    const int32_t ga = gene_idx1[k], gb = gene_idx2[k];
    if (ga < 0 || ga >= genes || gb < 0 || gb >= genes)
      return fail(BIXCOR_ERR_INVALID, "gene index out of range at pair %lld (0-based indices expected)", (long long)k);
    for (int32_t g : {ga, gb})
      if (slot_of[g] < 0) {
        slot_of[g] = (int32_t)gene_of_slot.size();
        gene_of_slot.push_back(g);
      }
    s1[k] = slot_of[ga];
    s2[k] = slot_of[gb];
  }
  const int64_t u = (int64_t)gene_of_slot.size();
  const int K = round_up((int)cells, dmma::BK);
  if ((double)u * (double)(cells + K) * 8.0 > 120e9)
    return fail(BIXCOR_ERR_UNSUPPORTED, "pair list touches %lld genes x %lld cells: split the pair list", (long long)u,
                (long long)cells);
  Ctx* c = g_ctx[0];
  CU_TRY(cudaSetDevice(c->device));
  call_begin(c);
  const size_t b_ptr = sizeof(int64_t) * (size_t)nptr, b_idx = sizeof(int32_t) * (size_t)nnz,
               b_dat = sizeof(float) * (size_t)nnz, b_map = sizeof(int32_t) * (size_t)genes,
               b_gos = sizeof(int32_t) * (size_t)u, b_pair = sizeof(int32_t) * (size_t)n_pairs,
               b_cpos = sizeof(int32_t) * cell_pos.size();
  auto al = [](size_t v) { return (v + 255) & ~size_t(255); };
  RC_TRY(c->sparse_in.reserve(al(b_ptr) + al(b_idx) + al(b_dat) + al(b_map) + al(b_gos) + 2 * al(b_pair) + al(b_cpos) + 256));
  char* base = (char*)c->sparse_in.ptr;
  int64_t* d_ptr = (int64_t*)base;
  int32_t* d_idx = (int32_t*)(base + al(b_ptr));
  float* d_dat = (float*)((char*)d_idx + al(b_idx));
  int32_t* d_map = (int32_t*)((char*)d_dat + al(b_dat));
  int32_t* d_gos = (int32_t*)((char*)d_map + al(b_map));
  int32_t* d_s1 = (int32_t*)((char*)d_gos + al(b_gos));
  int32_t* d_s2 = (int32_t*)((char*)d_s1 + al(b_pair));
  int32_t* d_cpos = cell_pos.empty() ? nullptr : (int32_t*)((char*)d_s2 + al(b_pair));
  RC_TRY(h2d(c, d_ptr, indptr, b_ptr, c->compute));
  RC_TRY(h2d(c, d_idx, indices, b_idx, c->compute));
  RC_TRY(h2d(c, d_dat, data, b_dat, c->compute));
  RC_TRY(h2d(c, d_map, slot_of.data(), b_map, c->compute));
  RC_TRY(h2d(c, d_gos, gene_of_slot.data(), b_gos, c->compute));
  RC_TRY(h2d(c, d_s1, s1.data(), b_pair, c->compute));
  RC_TRY(h2d(c, d_s2, s2.data(), b_pair, c->compute));
  if (d_cpos) RC_TRY(h2d(c, d_cpos, cell_pos.data(), b_cpos, c->compute));
  CU_TRY(cudaStreamSynchronize(c->compute));  // the index vectors above are stack/heap temporaries
  RC_TRY(c->x[0].reserve(sizeof(double) * (size_t)u * cells));
  RC_TRY(c->out.reserve(sizeof(double) * (size_t)n_pairs));
  CU_TRY(cudaMemsetAsync(c->x[0].ptr, 0, sizeof(double) * (size_t)u * cells, c->compute));
  {
    Timed tm(c, T_OTHER);
    if (is_csr)
      csr_gather_genes_kernel<<<(unsigned)((all_cells * 32 + 255) / 256), 256, 0, c->compute>>>(
          d_ptr, d_idx, d_dat, all_cells, genes, d_map, d_cpos, (double*)c->x[0].ptr, cells);
    else
      csc_gather_genes_kernel<<<(unsigned)u, 256, 0, c->compute>>>(d_ptr, d_idx, d_dat, all_cells, d_gos, d_cpos,
                                                                   (double*)c->x[0].ptr, cells);
    g_launches++;
  }
  if (sparse_ranks) {
    constexpr int kCap = 16384;
    constexpr int kSmem = kCap * 12;
    RC_TRY(c->vec.reserve(sizeof(int) * (size_t)u + 256));
    int* d_over = (int*)c->vec.ptr;  // one flag per densified gene: more than kCap non-zero cells
    CU_TRY(cudaMemsetAsync(d_over, 0, sizeof(int) * (size_t)u, c->compute));
    CU_TRY(cudaFuncSetAttribute(rank_sparse_columns_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    {
      Timed tm(c, T_COLS);
      rank_sparse_columns_kernel<<<(unsigned)u, kSparseRankThreads, kSmem, c->compute>>>((double*)c->x[0].ptr, cells, cells, kCap, d_over);
      g_launches++;
    }
    std::vector<int> over((size_t)u, 0);
    CU_TRY(cudaMemcpyAsync(over.data(), d_over, sizeof(int) * (size_t)u, cudaMemcpyDeviceToHost, c->compute));
    CU_TRY(cudaStreamSynchronize(c->compute));
    // dense genes (more non-zero cells than one CTA sorts in shared memory): the chunked dense rank kernels rank the whole
    // column in place - the zeros are one more tie run
    for (int64_t g = 0; g < u;) {
      if (!over[(size_t)g]) {
        ++g;
        continue;
      }
      int64_t g1 = g;
      while (g1 < u && over[(size_t)g1]) ++g1;
      ColumnOut co;
      memset(&co, 0, sizeof co);
      co.mode = COL_RANKS;
      co.f64 = (double*)c->x[0].ptr + g * cells;
      co.ld64 = cells;
      co.kpad = (int)cells;
      Timed tm(c, T_COLS);
      RC_TRY(launch_rank_big(c, (const double*)c->x[0].ptr + g * cells, cells, (int)cells, g1 - g, co));
      g = g1;
    }
  }
  Operand op;
  RC_TRY(prepare_operand(c, (const double*)c->x[0].ptr, cells, u, sparse_ranks ? 0 : spearman, BIXCOR_PREC_F64, 0, &op));
  {
    Timed tm(c, T_OTHER);
    pair_dot_kernel<<<(unsigned)((n_pairs * 32 + 255) / 256), 256, 0, c->compute>>>(op.z64, op.K, d_s1, d_s2, n_pairs,
                                                                                   (double*)c->out.ptr);
    g_launches++;
  }
  CU_TRY(cudaGetLastError());
  RC_TRY(finish(c));
  D2HPipe pipe(c, out);
  RC_TRY(pipe.push(out, c->out.ptr, sizeof(double) * (size_t)n_pairs, nullptr));
  return pipe.finish();
}

int bixcor_sparse_pairwise_cor(const int64_t* indptr, const int32_t* indices, const float* data, int64_t cells,
                               int64_t genes, int is_csr, const int32_t* gene_idx1, const int32_t* gene_idx2,
                               int64_t n_pairs, int spearman, double* out) {
  API_BEGIN
  return sparse_pairwise_impl(indptr, indices, data, cells, genes, is_csr, nullptr, 0, gene_idx1, gene_idx2, n_pairs, spearman,
                              out);
  API_END
}

// rs_pairwise_gene_cors (src/rust/src/single_cell/r_sc_processing.rs:474-499): as above over the listed cells only
int bixcor_sparse_pairwise_cor_cells(const int64_t* indptr, const int32_t* indices, const float* data, int64_t cells,
                                     int64_t genes, int is_csr, const int32_t* cells_to_keep, int64_t n_keep,
                                     const int32_t* gene_idx1, const int32_t* gene_idx2, int64_t n_pairs, int spearman,
                                     double* out) {
  API_BEGIN
  if (!cells_to_keep) return fail(BIXCOR_ERR_INVALID, "null pointer");
  return sparse_pairwise_impl(indptr, indices, data, cells, genes, is_csr, cells_to_keep, n_keep, gene_idx1, gene_idx2, n_pairs,
                              spearman, out);
  API_END
}

// One device's share of the sparse Gram: upload cells [cs0, cs1) of the CSR (row pointers rebased to the shard),
// accumulate the partial X'X and the column sums into d_gram / d_colsum (zeroed here).
static int sparse_shard_accumulate(Ctx* c, const int64_t* indptr, const int32_t* indices, const float* data, int64_t cs0,
                                   int64_t cs1, int64_t genes, int precision, double** d_gram, double** d_colsum) {
  const int64_t nc = cs1 - cs0;
  precision = resolve_precision_float(precision);
  RC_TRY(c->out2.reserve(sizeof(double) * (size_t)genes * genes));
  RC_TRY(c->vec.reserve(sizeof(double) * (size_t)genes * 4));
  RC_TRY(c->flag.reserve(256));
  CU_TRY(cudaMemsetAsync(c->out2.ptr, 0, sizeof(double) * (size_t)genes * genes, c->compute));
  CU_TRY(cudaMemsetAsync(c->vec.ptr, 0, sizeof(double) * (size_t)genes, c->compute));
  CU_TRY(cudaMemsetAsync(c->flag.ptr, 0, sizeof(int), c->compute));
  *d_gram = (double*)c->out2.ptr;
  *d_colsum = (double*)c->vec.ptr;
  if (nc <= 0) return BIXCOR_OK;
  // The CSR arrays (2 GB at 1M cells x 250 non-zeros) go up in slabs of 8 dense panels; the upload of slab s + 1 (copy-in
  // stream, two device buffers) runs while the densify + Gram kernels of slab s are busy, so the host -> device transfer is
  // hidden behind the math instead of preceding it.
  const int64_t kSlab = 8 * 8192;
  const int64_t nslab = (nc + kSlab - 1) / kSlab;
  int64_t max_nnz = 0;
  for (int64_t sl = 0; sl < nslab; ++sl) {
    const int64_t b = cs0 + sl * kSlab, e = std::min(cs1, b + kSlab);
    max_nnz = std::max(max_nnz, indptr[e] - indptr[b]);
  }
  auto al = [](size_t v) { return (v + 255) & ~size_t(255); };
  const size_t b_ptr = al(sizeof(int64_t) * (size_t)(std::min(nc, kSlab) + 1)), b_idx = al(sizeof(int32_t) * (size_t)max_nnz),
               b_dat = al(sizeof(float) * (size_t)max_nnz), half_bytes = b_ptr + b_idx + b_dat;
  RC_TRY(c->sparse_in.reserve(2 * half_bytes + 256));
  cudaEvent_t done[2] = {nullptr, nullptr};
  std::vector<int64_t> local;
  for (int64_t sl = 0; sl < nslab; ++sl) {
    const int h = (int)(sl & 1);
    const int64_t b = cs0 + sl * kSlab, e = std::min(cs1, b + kSlab), ncell = e - b, off = indptr[b], nnz = indptr[e] - off;
    char* base = (char*)c->sparse_in.ptr + (size_t)h * half_bytes;
    int64_t* d_ptr = (int64_t*)base;
    int32_t* d_idx = (int32_t*)(base + b_ptr);
    float* d_dat = (float*)(base + b_ptr + b_idx);
    if (done[h]) CU_TRY(cudaStreamWaitEvent(c->copy_in, done[h], 0));  // the kernels that read this buffer two slabs ago
    local.resize((size_t)ncell + 1);
    for (int64_t i = 0; i <= ncell; ++i) local[(size_t)i] = indptr[b + i] - off;  // row pointers relative to the slab
    RC_TRY(h2d(c, d_ptr, local.data(), sizeof(int64_t) * (size_t)(ncell + 1), c->copy_in));
    CU_TRY(cudaStreamSynchronize(c->copy_in));  // `local` is reused by the next slab (pinned callers: plain async DMA above)
    RC_TRY(h2d(c, d_idx, indices + off, sizeof(int32_t) * (size_t)nnz, c->copy_in));
    RC_TRY(h2d(c, d_dat, data + off, sizeof(float) * (size_t)nnz, c->copy_in));
    cudaEvent_t up = ctx_event(c);
    CU_TRY(cudaEventRecord(up, c->copy_in));
    CU_TRY(cudaStreamWaitEvent(c->compute, up, 0));
    RC_TRY(dev_sparse_accumulate(c, d_ptr, d_idx, d_dat, ncell, genes, precision, *d_gram, *d_colsum, 1));
    done[h] = ctx_event(c);
    CU_TRY(cudaEventRecord(done[h], c->compute));
  }
  if (precision == BIXCOR_PREC_F16X3) RC_TRY(sparse_overflow_check(c));
  return BIXCOR_OK;
}

int bixcor_sparse_gram_cor(const int64_t* indptr, const int32_t* indices, const float* data, int64_t cells,
                           int64_t genes, int precision, double* out_packed) {
  API_BEGIN
  RC_TRY(ensure_init());
  if (!indptr || !out_packed) return fail(BIXCOR_ERR_INVALID, "null pointer");
  if (cells < 2 || genes < 1) return fail(BIXCOR_ERR_INVALID, "bad shape");
  const int64_t nnz = indptr[cells];
  if (nnz < 0 || (nnz > 0 && (!indices || !data))) return fail(BIXCOR_ERR_INVALID, "bad CSR arrays");
  const size_t out_bytes = sizeof(double) * (size_t)(genes * (genes - 1) / 2);
  if (g_ctx.size() > 1 && !g_comms.empty() && cells >= 8192 * (int64_t)g_ctx.size()) {
    // contiguous cell shards, one per device; the partial Gram matrices and column sums are summed with
    // ncclAllReduce over NVLink; device 0 turns the total into Pearson correlations and returns the packed vector
    PhaseBarrier bar((int)g_ctx.size());
    return for_each_device([&](Ctx* c, int d, int nd) -> int {
      double *d_gram = nullptr, *d_colsum = nullptr;
      const int64_t cs0 = cells * d / nd, cs1 = cells * (d + 1) / nd;
      int rc = sparse_shard_accumulate(c, indptr, indices, data, cs0, cs1, genes, precision, &d_gram, &d_colsum);
      rc = bar.arrive(rc);  // all devices reach the collective or none does
      if (rc != BIXCOR_OK) return rc;
      NCCL_TRY(g_nccl.GroupStart());
      NCCL_TRY(g_nccl.AllReduce(d_gram, d_gram, (size_t)genes * genes, ncclDouble, ncclSum, g_comms[d], c->compute));
      NCCL_TRY(g_nccl.AllReduce(d_colsum, d_colsum, (size_t)genes, ncclDouble, ncclSum, g_comms[d], c->compute));
      NCCL_TRY(g_nccl.GroupEnd());
      if (d != 0) return finish(c);
      RC_TRY(c->out.reserve(std::max<size_t>(out_bytes, 8)));
      RC_TRY(dev_sparse_finalise(c, d_gram, d_colsum, cells, genes, (double*)c->out.ptr));
      RC_TRY(finish(c));
      D2HPipe pipe(c);
      RC_TRY(pipe.push(out_packed, c->out.ptr, out_bytes, nullptr));
      return pipe.finish();
    });
  }
  Ctx* c = g_ctx[0];
  CU_TRY(cudaSetDevice(c->device));
  call_begin(c);
  double *d_gram = nullptr, *d_colsum = nullptr;
  RC_TRY(sparse_shard_accumulate(c, indptr, indices, data, 0, cells, genes, precision, &d_gram, &d_colsum));
  RC_TRY(c->out.reserve(std::max<size_t>(out_bytes, 8)));
  RC_TRY(dev_sparse_finalise(c, d_gram, d_colsum, cells, genes, (double*)c->out.ptr));
  RC_TRY(finish(c));
  D2HPipe pipe(c, out_packed);
  RC_TRY(pipe.push(out_packed, c->out.ptr, out_bytes, nullptr));
  return pipe.finish();
  API_END
}

}  // extern "C"
