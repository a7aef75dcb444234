// Host-side SIMD helpers of libbixcor (compiled by the host compiler alone: the AVX2 intrinsic headers stay out of the
// CUDA translation unit).  See expand_packed_s32 in bixcor.cu: the copier threads turn the exact integer Spearman Gram
// (int32, half the PCIe bytes) into r = S / (|d_i| |d_j|) (and |r|^6) while they move it into the caller's buffer.
#include <immintrin.h>
#include <stdint.h>
#include <string.h>

static inline void expand_one(double* dst, int32_t sv, double rf, double cfv, int pow6) {
  double v = (double)sv * rf;
  v *= cfv;
  if (pow6) {
    const double sq = v * v;
    v = sq * sq * sq;
  }
  long long bits;
  memcpy(&bits, &v, sizeof bits);
  _mm_stream_si64(reinterpret_cast<long long*>(dst), bits);  // streaming store: the destination is not read first
}

extern "C" __attribute__((visibility("hidden"))) int bixcor_host_has_avx2(void) { return __builtin_cpu_supports("avx2") ? 1 : 0; }

// dst[t] = pow(double(src[t]) * rf * cf[t]) for one run of a packed row, scalar form
extern "C" __attribute__((visibility("hidden"))) void bixcor_expand_run_scalar(double* dst, const int32_t* src, int64_t n, double rf, const double* cf, int pow6) {
  for (int64_t t = 0; t < n; ++t) expand_one(dst + t, src[t], rf, cf[t], pow6);
}

// the same with 4 elements per step: identical IEEE multiplications in the same order, so both forms agree bit for bit
extern "C" __attribute__((visibility("hidden"), target("avx2"))) void bixcor_expand_run_avx2(double* dst, const int32_t* src, int64_t n, double rf,
                                                                      const double* cf, int pow6) {
  int64_t t = 0;
  while (t < n && ((uintptr_t)(dst + t) & 31)) {
    expand_one(dst + t, src[t], rf, cf[t], pow6);
    ++t;
  }
  const __m256d vrf = _mm256_set1_pd(rf);
  for (; t + 4 <= n; t += 4) {
    __m256d v = _mm256_cvtepi32_pd(_mm_loadu_si128(reinterpret_cast<const __m128i*>(src + t)));
    v = _mm256_mul_pd(v, vrf);
    v = _mm256_mul_pd(v, _mm256_loadu_pd(cf + t));
    if (pow6) {
      const __m256d sq = _mm256_mul_pd(v, v);
      v = _mm256_mul_pd(_mm256_mul_pd(sq, sq), sq);
    }
    _mm256_stream_pd(dst + t, v);
  }
  for (; t < n; ++t) expand_one(dst + t, src[t], rf, cf[t], pow6);
}

// binary32 transport of the <= 1e-5-class results (bixcor.cu, D2HPipe::push_f32): dst[t] = (double)src[t], streaming stores
extern "C" __attribute__((visibility("hidden"))) void bixcor_widen_f32_scalar(double* dst, const float* src, int64_t n) {
  for (int64_t t = 0; t < n; ++t) {
    const double v = (double)src[t];
    long long bits;
    memcpy(&bits, &v, sizeof bits);
    _mm_stream_si64(reinterpret_cast<long long*>(dst + t), bits);
  }
}

extern "C" __attribute__((visibility("hidden"), target("avx2"))) void bixcor_widen_f32_avx2(double* dst, const float* src, int64_t n) {
  int64_t t = 0;
  while (t < n && ((uintptr_t)(dst + t) & 31)) {
    bixcor_widen_f32_scalar(dst + t, src + t, 1);
    ++t;
  }
  for (; t + 8 <= n; t += 8) {
    const __m256 f = _mm256_loadu_ps(src + t);
    _mm256_stream_pd(dst + t, _mm256_cvtps_pd(_mm256_castps256_ps128(f)));
    _mm256_stream_pd(dst + t + 4, _mm256_cvtps_pd(_mm256_extractf128_ps(f, 1)));
  }
  if (t < n) bixcor_widen_f32_scalar(dst + t, src + t, n - t);
}
