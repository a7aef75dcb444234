// Micro-benchmark: vector FP64 (DMUL/DFMA), I2F.F64 and DSETP issue rates per SM on B200.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(double* out, int iters, double seed) {
  double a[16];
  int ii[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { a[i] = seed + i * 1e-3 + threadIdx.x * 1e-6; ii[i] = threadIdx.x + i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (OP == 0) a[i] = a[i] * 1.0000001;                 // DMUL
      else if (OP == 1) a[i] = fma(a[i], 1.0000001, 1e-9);   // DFMA
      else if (OP == 2) { a[i] += (double)ii[i]; ii[i] += it; }  // I2F.F64 + DADD + IADD
      else if (OP == 3) { ii[i] += (a[i] < seed + ii[i]) ? 1 : 2; }  // I2F + DSETP + SEL
      else if (OP == 4) { float f = (float)a[i]; f = f * 1.0000001f; a[i] = (double)f; }  // F2F both ways + FMUL
    }
  }
  double s = 0; int t = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) { s += a[i]; t += ii[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + t;
}
template <int OP>
void run(const char* name, int warps_per_sm, int ops_per_iter) {
  double* out; cudaMalloc(&out, 148 * 1024 * 8);
  const int iters = 4096;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<OP><<<148, warps_per_sm * 32>>>(out, 16, 1.0);
  cudaEventRecord(e0);
  k<OP><<<148, warps_per_sm * 32>>>(out, iters, 1.0);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double warp_instr = (double)warps_per_sm * iters * 16 * ops_per_iter;  // per SM
  double clk = ms * 1e-3 * 1.965e9;
  printf("%-28s warps/SM %2d  %.3f ms  %.3f warp-instr/clk/SM (counted ops)  -> %.1f lanes/clk/SM\n", name, warps_per_sm, ms, warp_instr / clk, 32 * warp_instr / clk);
  cudaFree(out);
}
int main() {
  for (int w : {4, 8, 16, 32}) run<0>("DMUL", w, 1);
  for (int w : {4, 8, 16, 32}) run<1>("DFMA", w, 1);
  for (int w : {4, 8, 16}) run<2>("I2F.F64+DADD(+IADD)", w, 2);
  for (int w : {4, 8, 16}) run<3>("I2F.F64+DADD+DSETP", w, 3);
  for (int w : {4, 8, 16}) run<4>("F2F.F32.F64+FMUL+F2F.F64.F32", w, 3);
  return 0;
}
