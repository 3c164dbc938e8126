// 24 kHz -> 16 kHz polyphase windowed-sinc resampler (3 -> 2).
//
// Reference: torchaudio.functional.resample(x, 24000, 16000)
//   torchaudio/functional/functional.py:1305-1405 (_get_sinc_resample_kernel: sinc_interp_hann,
//   lowpass_filter_width 6, rolloff 0.99 -> width 10, 2 phases x 23 taps) and :1408-1432
//   (_apply_sinc_resample_kernel: pad (10, 13), conv1d stride 3, truncate to ceil(2L/3)).
//   y[2m+p] = sum_i xpad[3m+i] * k[p][i],  xpad[j] = x[j-10] (0 outside [0, L)).
// Reference call sites: base_tts.py:632 and the 16 kHz loaders behind base_tts.py:338 /
// stt_validator.py:78-107.
//
// HBM-bound: 4 B in + 2.67 B out per input sample.  Each CTA stages 3*M+24 input samples in
// shared memory with 128-bit loads and every thread produces 4 consecutive outputs (2 pairs)
// with a single 128-bit store; taps sit in __constant__ (warp-uniform operand of the FFMA).
#include "common.cuh"
#include "kernels.h"

namespace rho {

__constant__ float c_taps[2][RS_TAPS];

cudaError_t upload_resample_taps(const float* taps) {
  return cudaMemcpyToSymbol(c_taps, taps, sizeof(float) * 2 * RS_TAPS);
}

constexpr int RSM_THREADS = 256;
constexpr int RSM_SUB = RSM_THREADS * 2;       // output pairs per sub-tile (2 per thread)
constexpr int RSM_ITERS = 4;                   // sub-tiles per CTA
constexpr int RSM_PAIRS = RSM_SUB * RSM_ITERS; // 2048 pairs = 4096 outputs = 6144 inputs per CTA
constexpr int RSM_SMEM = 3 * RSM_SUB + 32;     // floats staged per sub-tile (halo 12 left, 14+ right)

__global__ void __launch_bounds__(RSM_THREADS)
k_resample3to2(const float* __restrict__ x, const int64_t* __restrict__ off, const char* __restrict__ len_base,
               int len_stride, float* __restrict__ y, const int64_t* __restrict__ y_off, int32_t* __restrict__ y_len) {
  __shared__ __align__(16) float sm[2][RSM_SMEM];
  const int c = blockIdx.x;
  const int L = *reinterpret_cast<const int32_t*>(len_base + (size_t)c * len_stride);
  const int target = L <= 0 ? 0 : (int)((2LL * L + 2) / 3);     // ceil(2L/3)
  if (blockIdx.y == 0 && threadIdx.x == 0 && y_len) y_len[c] = target;
  const int n_pairs = (target + 1) >> 1;
  const int tile_m0 = blockIdx.y * RSM_PAIRS;
  if (tile_m0 >= n_pairs) return;
  const float* __restrict__ xs = x + off[c];
  float* __restrict__ ys = y + y_off[c];

  for (int it = 0; it < RSM_ITERS; ++it) {
    const int m0 = tile_m0 + it * RSM_SUB;
    if (m0 >= n_pairs) break;
    float* s = sm[it & 1];
    // stage x[a0 .. a0 + 3*SUB + 24), a0 = 3*m0 - 12 (multiple of 4)
    const long long a0 = 3LL * m0 - 12;
    for (int q = threadIdx.x; q < RSM_SMEM / 4; q += RSM_THREADS) {
      const long long g = a0 + 4 * q;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (g >= 0 && g + 3 < L) v = ldg_stream4(xs + g);
      else if (g + 3 >= 0 && g < L) {
        if (g + 0 >= 0 && g + 0 < L) v.x = xs[g + 0];
        if (g + 1 >= 0 && g + 1 < L) v.y = xs[g + 1];
        if (g + 2 >= 0 && g + 2 < L) v.z = xs[g + 2];
        if (g + 3 >= 0 && g + 3 < L) v.w = xs[g + 3];
      }
      *reinterpret_cast<float4*>(s + 4 * q) = v;
    }
    __syncthreads();   // double-buffered: the next iteration writes the other buffer
    // thread u -> pairs m0+2u, m0+2u+1; needs s[6u+2 .. 6u+28)
    const int u = threadIdx.x;
    float v[26];
    const float2* sp = reinterpret_cast<const float2*>(s + 6 * u + 2);
#pragma unroll
    for (int k = 0; k < 13; ++k) { const float2 t = sp[k]; v[2 * k] = t.x; v[2 * k + 1] = t.y; }
    float o00 = 0.f, o01 = 0.f, o10 = 0.f, o11 = 0.f;
    // taps 0 and 20..22 of phase 0, and 0..2, 21..22 of phase 1, sit on the clamp of the
    // window (|k| ~ 3e-24): they are included anyway, 46 FMAs per pair.
#pragma unroll
    for (int i = 0; i < RS_TAPS; ++i) {
      o00 = fmaf(v[i], c_taps[0][i], o00);
      o01 = fmaf(v[i], c_taps[1][i], o01);
      o10 = fmaf(v[i + 3], c_taps[0][i], o10);
      o11 = fmaf(v[i + 3], c_taps[1][i], o11);
    }
    const int o = 2 * (m0 + 2 * u);
    if (o + 3 < target) {
      stg_stream4(ys + o, make_float4(o00, o01, o10, o11));
    } else {
      if (o + 0 < target) ys[o + 0] = o00;
      if (o + 1 < target) ys[o + 1] = o01;
      if (o + 2 < target) ys[o + 2] = o10;
      if (o + 3 < target) ys[o + 3] = o11;
    }
  }
}

cudaError_t launch_resample3to2(const float* x, const int64_t* off, const int32_t* len, int len_stride_bytes,
                                int n, int64_t max_len, float* y, const int64_t* y_off, int32_t* y_len,
                                cudaStream_t st, LaunchCtx* lc) {
  if (n <= 0) return cudaSuccess;
  const int64_t max_pairs = ((2 * max_len + 2) / 3 + 1) / 2;
  unsigned tiles = (unsigned)((max_pairs + RSM_PAIRS - 1) / RSM_PAIRS);
  if (tiles == 0) tiles = 1;
  dim3 grid((unsigned)n, tiles);
  lc->begin(KID_RESAMPLE, st);
  k_resample3to2<<<grid, RSM_THREADS, 0, st>>>(x, off, reinterpret_cast<const char*>(len),
                                               len_stride_bytes ? len_stride_bytes : (int)sizeof(int32_t),
                                               y, y_off, y_len);
  lc->end(st);
  return cudaGetLastError();
}

}  // namespace rho
