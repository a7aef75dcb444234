// Measured tcgen05.mma issue peaks on this GPU for the kinds libbixcor uses ("measure int8/tf32 MMA peak
// instead of assuming 2x / 0.5x bf16").  One CTA per SM issues back-to-back MMAs on operands that are already in
// shared memory (no TMA, no epilogue): the number is the tensor pipe's rate for that instruction shape, the
// denominator of roofline.frac for the Gram kernels.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o mma_peak mma_peak.cu -lcuda
//   ./mma_peak > profiles/mma_peaks_r02.json
#include <cstdio>
#include <vector>

#include "../../bixverse_b200/csrc/gram_umma.cuh"

using namespace bixcor;
using namespace bixcor::umma;

// KIND: umma::Kind, N: MMA N (128 or 256), CTAS: 1 or 2 (cta_group::2, M = 256 over the pair)
template <int KIND, int N, int CTAS>
__global__ void __launch_bounds__(128, 1) mma_peak_kernel(int iters, unsigned long long* cycles) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = base + 96 * 1024;
  const uint32_t slot = bar + 16;
  uint32_t cta_rank = 0;
  if (CTAS == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
  if (warp == 1 && lane == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    if (CTAS == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "n"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "n"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
  if (warp == 1 && lane == 0 && cta_rank == 0) {
    constexpr uint32_t idesc = make_idesc<KIND, CTAS, N>();
    // A: 128 rows x 128 B (16 KB), B: N / CTAS rows x 128 B per CTA; 4 K steps of 32 B inside the swizzle row, two operand
    // tiles alternate so that consecutive MMAs do not read the very same lines
    const uint64_t dA[2] = {make_smem_desc(base), make_smem_desc(base + 16 * 1024)};
    const uint64_t dB[2] = {make_smem_desc(base + 32 * 1024), make_smem_desc(base + 64 * 1024)};
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        tc_mma<KIND, CTAS>(tmem, dA[it & 1] + 2 * k, dB[it & 1] + 2 * k, idesc, (it | k) ? 1u : 0u);
        tc_mma<KIND, CTAS>(tmem + N, dA[(it + 1) & 1] + 2 * k, dB[it & 1] + 2 * k, idesc, (it | k) ? 1u : 0u);
      }
    }
    if (CTAS == 2) tc_commit_pair(bar); else tc_commit(bar);
    mbar_wait(bar, 0);
    cycles[blockIdx.x] = (unsigned long long)(clock64() - t0);
  }
  tc_fence_before();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (CTAS == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
  }
}

template <int KIND, int N, int CTAS>
static double run(int sms, int iters, const char* name, bool last) {
  auto kern = mma_peak_kernel<KIND, N, CTAS>;
  const int smem = 100 * 1024;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  unsigned long long* d_cyc;
  cudaMalloc(&d_cyc, sizeof(unsigned long long) * 256);
  cudaLaunchConfig_t cfg = {};
  const int grid = (sms / CTAS) * CTAS;
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTAS;
  attr[0].val.clusterDim.y = attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e9f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    cudaLaunchKernelEx(&cfg, kern, iters, d_cyc);
    cudaEventRecord(e1);
    cudaError_t e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) {
      printf("  {\"kind\": \"%s\", \"error\": \"%s\"}%s\n", name, cudaGetErrorString(e), last ? "" : ",");
      return 0;
    }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  const int kelem = KIND == KIND_TF32 ? 8 : (KIND == KIND_F16 ? 16 : 32);  // K elements of one 32-byte MMA step
  // per iteration: 8 MMAs of (128 * CTAS) x N x kelem on each CTA group
  const double flops = 2.0 * (128.0 * CTAS) * N * kelem * 8.0 * iters * (grid / CTAS);
  const double tf = flops / (best * 1e-3) / 1e12;
  printf("  {\"kind\": \"%s\", \"mma_n\": %d, \"cta_group\": %d, \"ctas\": %d, \"ms\": %.4f, \"tflops\": %.1f}%s\n", name, N, CTAS, grid, best, tf,
         last ? "" : ",");
  cudaFree(d_cyc);
  return tf;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount, iters = 20000;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"how\": \"back-to-back tcgen05.mma from resident shared-memory operands, one CTA per SM, CUDA events, best of 4\",\n \"rows\": [\n",
         prop.name, sms);
  run<KIND_I8, 128, 1>(sms, iters, "i8", false);
  run<KIND_I8, 256, 1>(sms, iters, "i8", false);
  run<KIND_I8, 128, 2>(sms, iters, "i8", false);
  run<KIND_I8, 256, 2>(sms, iters, "i8", false);
  run<KIND_F16, 128, 1>(sms, iters, "f16", false);
  run<KIND_F16, 256, 1>(sms, iters, "f16", false);
  run<KIND_F16, 128, 2>(sms, iters, "f16", false);
  run<KIND_F16, 256, 2>(sms, iters, "f16", false);
  run<KIND_TF32, 128, 1>(sms, iters, "tf32", false);
  run<KIND_TF32, 256, 1>(sms, iters, "tf32", false);
  run<KIND_TF32, 128, 2>(sms, iters, "tf32", false);
  run<KIND_TF32, 256, 2>(sms, iters, "tf32", true);
  printf(" ]\n}\n");
  return 0;
}
