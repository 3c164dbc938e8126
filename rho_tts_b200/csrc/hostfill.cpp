// Host-side constant fill of feature-row tails (host_api.cu): plain C++ (g++), no CUDA.  Non-temporal stores, widest
// the CPU has (AVX-512 / AVX2 / SSE2 picked at run time): the rows are written once into the caller's pinned buffer
// and read by somebody else later, they should not travel through the cache, and two fill threads have to keep up
// with 0.64 GB of constants per 1.28 GB that crosses the link.
#include <cstddef>
#include <cstdint>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace rho {

#if defined(__x86_64__)
__attribute__((target("avx512f"))) static void fill_avx512(float* p, int64_t n, float v) {
  while (n > 0 && (reinterpret_cast<uintptr_t>(p) & 63u)) { *p++ = v; --n; }
  const __m512 vv = _mm512_set1_ps(v);
  const int64_t q = n >> 4;
  for (int64_t i = 0; i < q; ++i) _mm512_stream_ps(p + 16 * i, vv);
  p += 16 * q; n -= 16 * q;
  while (n > 0) { *p++ = v; --n; }
  _mm_sfence();
}
__attribute__((target("avx2"))) static void fill_avx2(float* p, int64_t n, float v) {
  while (n > 0 && (reinterpret_cast<uintptr_t>(p) & 31u)) { *p++ = v; --n; }
  const __m256 vv = _mm256_set1_ps(v);
  const int64_t q = n >> 3;
  for (int64_t i = 0; i < q; ++i) _mm256_stream_ps(p + 8 * i, vv);
  p += 8 * q; n -= 8 * q;
  while (n > 0) { *p++ = v; --n; }
  _mm_sfence();
}
static void fill_sse2(float* p, int64_t n, float v) {
  while (n > 0 && (reinterpret_cast<uintptr_t>(p) & 15u)) { *p++ = v; --n; }
  const __m128 vv = _mm_set1_ps(v);
  const int64_t q = n >> 2;
  for (int64_t i = 0; i < q; ++i) _mm_stream_ps(p + 4 * i, vv);
  p += 4 * q; n -= 4 * q;
  while (n > 0) { *p++ = v; --n; }
  _mm_sfence();
}
#endif

void host_fill_f32(float* p, int64_t n, float v) {
#if defined(__x86_64__)
  static const int level = [] {
    __builtin_cpu_init();
    return __builtin_cpu_supports("avx512f") ? 2 : (__builtin_cpu_supports("avx2") ? 1 : 0);
  }();
  if (level == 2) fill_avx512(p, n, v);
  else if (level == 1) fill_avx2(p, n, v);
  else fill_sse2(p, n, v);
#else
  for (int64_t i = 0; i < n; ++i) p[i] = v;
#endif
}

}  // namespace rho
