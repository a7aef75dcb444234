// tcgen05 / TMEM Gram kernels (fast 3xTF32 path and exact int8 Spearman path).
#pragma once
#include "common.cuh"
namespace bixcor {
namespace umma {
inline const char* last_error() { return "tcgen05 Gram path not built yet"; }
inline int launch_gram_umma(cudaStream_t, int, int, int, const int8_t*, const float*, const double*, int,
                            const int8_t*, const float*, const double*, int64_t, const int2*, int, int, double*,
                            int64_t, int64_t, int, int, int, double, int, double*, double*, double*, double*, double,
                            const double*, const double*, const double*, int, int) {
  return -1;
}
}  // namespace umma
}  // namespace bixcor
