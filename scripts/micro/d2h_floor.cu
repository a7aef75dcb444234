// D2H / host-copy micro-benchmark behind the e2e floor of bench.py (VERDICT r1 "prove or fix the e2e floor").
//   nvcc -O3 -std=c++17 -arch=sm_100a -Xcompiler -O3,-march=x86-64-v3,-pthread -o d2h_floor d2h_floor.cu
//   ./d2h_floor [n_gpus]
// Measures, on the GPU box's own host:
//   1. plain cudaMemcpyAsync device -> PINNED host, 200 MB .. 1.6 GB per GPU, 1 .. n_gpus GPUs at once (one host
//      thread per GPU): the PCIe floor of the 1.6 GB packed result;
//   2. cudaMemcpy device -> PAGEABLE host (the driver's own staging);
//   3. cudaHostRegister / cudaHostUnregister of a 1.6 GB pageable buffer (pinning the caller's buffer in place);
//   4. host memcpy pinned -> pageable, 1.6 GB: glibc memcpy and a non-temporal (streaming-store) copy, 1 .. T threads.
#include <cuda_runtime.h>
#include <immintrin.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e = (x);                                                                   \
    if (e != cudaSuccess) {                                                                \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);      \
      exit(1);                                                                             \
    }                                                                                      \
  } while (0)

static void nt_copy(void* dst, const void* src, size_t bytes) {  // 32-byte aligned chunks, streaming stores
  const __m256i* s = (const __m256i*)src;
  __m256i* d = (__m256i*)dst;
  const size_t n = bytes / 32;
  for (size_t i = 0; i + 4 <= n; i += 4) {
    __m256i a = _mm256_load_si256(s + i), b = _mm256_load_si256(s + i + 1), c = _mm256_load_si256(s + i + 2), e = _mm256_load_si256(s + i + 3);
    _mm256_stream_si256(d + i, a);
    _mm256_stream_si256(d + i + 1, b);
    _mm256_stream_si256(d + i + 2, c);
    _mm256_stream_si256(d + i + 3, e);
  }
  _mm_sfence();
}

template <class F>
static double par(int nthr, size_t bytes, F f) {
  std::vector<std::thread> th;
  const size_t chunk = ((bytes / nthr) + 4095) & ~size_t(4095);
  const double t0 = now();
  for (int t = 0; t < nthr; ++t) {
    const size_t o = (size_t)t * chunk;
    if (o >= bytes) break;
    const size_t len = std::min(chunk, bytes - o);
    th.emplace_back([=] { f(o, len); });
  }
  for (auto& t : th) t.join();
  return now() - t0;
}

int main(int argc, char** argv) {
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  int want = argc > 1 ? atoi(argv[1]) : ndev;
  if (want < ndev) ndev = want;
  const size_t GB16 = 1599920000;  // the packed result of config 2
  printf("{\"devices\": %d, \"host_threads\": %u,\n", ndev, std::thread::hardware_concurrency());
  // ---- 1. pinned D2H floor
  printf(" \"d2h_pinned\": [\n");
  std::vector<void*> dbuf(ndev), hbuf(ndev);
  std::vector<cudaStream_t> st(ndev);
  for (int d = 0; d < ndev; ++d) {
    CK(cudaSetDevice(d));
    CK(cudaMalloc(&dbuf[d], GB16));
    CK(cudaMemset(dbuf[d], 1, GB16));
    CK(cudaMallocHost(&hbuf[d], GB16));
    CK(cudaStreamCreate(&st[d]));
  }
  bool first = true;
  for (int g = 1; g <= ndev; g *= 2) {
    for (size_t bytes : {size_t(200) << 20, size_t(800) << 20, GB16}) {
      double best = 1e9;
      for (int rep = 0; rep < 4; ++rep) {
        const double t0 = now();
        std::vector<std::thread> th;
        for (int d = 0; d < g; ++d)
          th.emplace_back([&, d] {
            cudaSetDevice(d);
            cudaMemcpyAsync(hbuf[d], dbuf[d], bytes, cudaMemcpyDeviceToHost, st[d]);
            cudaStreamSynchronize(st[d]);
          });
        for (auto& t : th) t.join();
        const double dt = now() - t0;
        if (rep && dt < best) best = dt;
      }
      printf("%s  {\"gpus\": %d, \"mb_per_gpu\": %.0f, \"ms\": %.3f, \"gbs_per_gpu\": %.2f, \"gbs_total\": %.2f}", first ? "" : ",\n", g,
             bytes / 1048576.0, best * 1e3, bytes / best / 1e9, g * (double)bytes / best / 1e9);
      first = false;
    }
  }
  printf("\n ],\n");
  CK(cudaSetDevice(0));
  // ---- 2. driver-staged D2H into pageable memory
  void* page = aligned_alloc(4096, GB16 + 4096);
  memset(page, 0, GB16);  // touched: page faults are not part of this number
  {
    double best = 1e9;
    for (int rep = 0; rep < 3; ++rep) {
      const double t0 = now();
      CK(cudaMemcpy(page, dbuf[0], GB16, cudaMemcpyDeviceToHost));
      const double dt = now() - t0;
      if (dt < best) best = dt;
    }
    printf(" \"d2h_pageable_driver\": {\"ms\": %.2f, \"gbs\": %.2f},\n", best * 1e3, GB16 / best / 1e9);
  }
  // ---- 3. pinning the caller's buffer in place
  {
    double t0 = now();
    cudaError_t e = cudaHostRegister(page, GB16, cudaHostRegisterDefault);
    double t1 = now();
    if (e == cudaSuccess) {
      double tc0 = now();
      CK(cudaMemcpy(page, dbuf[0], GB16, cudaMemcpyDeviceToHost));
      double tc1 = now();
      double t2 = now();
      cudaHostUnregister(page);
      double t3 = now();
      printf(" \"host_register_1p6gb\": {\"register_ms\": %.2f, \"copy_ms\": %.2f, \"unregister_ms\": %.2f},\n", (t1 - t0) * 1e3, (tc1 - tc0) * 1e3,
             (t3 - t2) * 1e3);
      // chunked registration (64 MB) to see whether the cost is linear
      t0 = now();
      const size_t ck = size_t(64) << 20;
      int nreg = 0;
      for (size_t o = 0; o + ck <= GB16; o += ck, ++nreg) cudaHostRegister((char*)page + o, ck, cudaHostRegisterDefault);
      t1 = now();
      for (size_t o = 0; o + ck <= GB16; o += ck) cudaHostUnregister((char*)page + o);
      t2 = now();
      printf(" \"host_register_64mb_chunks\": {\"chunks\": %d, \"register_ms_each\": %.3f, \"unregister_ms_each\": %.3f},\n", nreg,
             (t1 - t0) * 1e3 / nreg, (t2 - t1) * 1e3 / nreg);
    } else {
      printf(" \"host_register_1p6gb\": {\"error\": \"%s\"},\n", cudaGetErrorString(e));
      cudaGetLastError();
    }
  }
  // ---- 4. host copies pinned -> pageable
  printf(" \"host_copy_1p6gb\": [\n");
  first = true;
  const unsigned hw = std::thread::hardware_concurrency();
  for (int nthr : {1, 2, 4, 8, 16, 32}) {
    if ((unsigned)nthr > hw) break;
    double bm = 1e9, bn = 1e9;
    for (int rep = 0; rep < 3; ++rep) {
      double dt = par(nthr, GB16, [&](size_t o, size_t len) { memcpy((char*)page + o, (char*)hbuf[0] + o, len); });
      if (dt < bm) bm = dt;
      dt = par(nthr, GB16, [&](size_t o, size_t len) { nt_copy((char*)page + o, (char*)hbuf[0] + o, len); });
      if (dt < bn) bn = dt;
    }
    printf("%s  {\"threads\": %d, \"memcpy_ms\": %.2f, \"memcpy_gbs\": %.2f, \"nt_ms\": %.2f, \"nt_gbs\": %.2f}", first ? "" : ",\n", nthr, bm * 1e3,
           GB16 / bm / 1e9, bn * 1e3, GB16 / bn / 1e9);
    first = false;
  }
  printf("\n ]\n}\n");
  return 0;
}
