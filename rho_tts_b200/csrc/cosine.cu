// Speaker-embedding cosine similarity: dot(ref, e) / (|ref| * |e|), one warp per embedding.
// Reference: src/rho_tts/base_tts.py:341-344 (numpy fp32, no epsilon guard: a zero vector
// gives nan, exactly like the reference).
#include "common.cuh"
#include "kernels.h"

namespace rho {

__global__ void __launch_bounds__(256)
k_cosine(const float* __restrict__ emb, const float* __restrict__ ref, int n, int dim,
         char* __restrict__ out, int out_stride) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n) return;
  const float* __restrict__ e = emb + (size_t)w * dim;
  float dot = 0.f, ne = 0.f, nr = 0.f;
  if ((dim & 3) == 0 && ((((uintptr_t)e) | ((uintptr_t)ref)) & 15u) == 0) {
    for (int i = 4 * lane; i < dim; i += 128) {
      const float4 a = *reinterpret_cast<const float4*>(e + i);
      const float4 r = __ldg(reinterpret_cast<const float4*>(ref + i));
      dot += a.x * r.x + a.y * r.y + a.z * r.z + a.w * r.w;
      ne += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
      nr += r.x * r.x + r.y * r.y + r.z * r.z + r.w * r.w;
    }
  } else {
    for (int i = lane; i < dim; i += 32) {
      const float a = e[i], r = __ldg(ref + i);
      dot += a * r; ne += a * a; nr += r * r;
    }
  }
  dot = warp_sum(dot); ne = warp_sum(ne); nr = warp_sum(nr);
  if (lane == 0) {
    const float denom = __fmul_rn(__fsqrt_rn(nr), __fsqrt_rn(ne));
    *reinterpret_cast<float*>(out + (size_t)w * out_stride) = __fdiv_rn(dot, denom);
  }
}

cudaError_t launch_cosine(const float* emb, const float* ref, int n, int dim, float* out, int out_stride_bytes,
                          cudaStream_t st, LaunchCtx* lc) {
  if (n <= 0) return cudaSuccess;
  const int warps_per_cta = 8;
  lc->begin(KID_COSINE, st);
  k_cosine<<<(n + warps_per_cta - 1) / warps_per_cta, 32 * warps_per_cta, 0, st>>>(
      emb, ref, n, dim, reinterpret_cast<char*>(out), out_stride_bytes ? out_stride_bytes : (int)sizeof(float));
  lc->end(st);
  return cudaGetLastError();
}

}  // namespace rho
