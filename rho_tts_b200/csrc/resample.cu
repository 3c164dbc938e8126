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
// HBM-bound: 4 B in + 2.67 B out per input sample.  Each CTA stages the 6144 + 24 input samples of its 4096
// outputs in shared memory with zero-filling LDGSTS -- all 24 KB in flight at once, no registers (the stream is
// bound by the bytes in flight per SM) -- and every thread then produces 4 x 4 consecutive outputs, each quad
// as four packed fp32x2 dot products (39 FFMA2, taps in uniform registers) and one 128-bit store.
#include "common.cuh"
#include "kernels.h"

namespace rho {

// Taps as float2 pairs for the packed dot products.  With V[k] = (v[2k], v[2k+1]) the four outputs of a quad are
//   o00 = sum_k V[k] . A0[k]  (k = 0..11)   A0[k] = (k0[2k],   k0[2k+1])
//   o10 = sum_k V[k] . B0[k]  (k = 1..12)   B0[k] = (k0[2k-3], k0[2k-2])
//   o01 = sum_k V[k] . A1[k]  (k = 0..11)   A1[k] = (k1[2k],   k1[2k+1])
//   o11 = sum_k V[k] . B1[k]  (k = 1..12)   B1[k] = (k1[2k-3], k1[2k-2])        (k[p][i] = 0 outside 0..22)
// All 23 taps of each phase are used, as torchaudio's conv1d does.
__constant__ float2 c_rA0[13], c_rB0[13], c_rA1[13], c_rB1[13];

cudaError_t upload_resample_taps(const float* taps) {
  const float* k0 = taps;
  const float* k1 = taps + RS_TAPS;
  auto t0 = [&](int i) { return (i >= 0 && i < RS_TAPS) ? k0[i] : 0.f; };
  auto t1 = [&](int i) { return (i >= 0 && i < RS_TAPS) ? k1[i] : 0.f; };
  float2 A0[13], B0[13], A1[13], B1[13];
  for (int k = 0; k < 13; ++k) {
    A0[k] = make_float2(t0(2 * k), t0(2 * k + 1));
    B0[k] = make_float2(t0(2 * k - 3), t0(2 * k - 2));
    A1[k] = make_float2(t1(2 * k), t1(2 * k + 1));
    B1[k] = make_float2(t1(2 * k - 3), t1(2 * k - 2));
  }
  cudaError_t e;
  if ((e = cudaMemcpyToSymbol(c_rA0, A0, sizeof(A0))) != cudaSuccess) return e;
  if ((e = cudaMemcpyToSymbol(c_rB0, B0, sizeof(B0))) != cudaSuccess) return e;
  if ((e = cudaMemcpyToSymbol(c_rA1, A1, sizeof(A1))) != cudaSuccess) return e;
  return cudaMemcpyToSymbol(c_rB1, B1, sizeof(B1));
}

constexpr int RSM_THREADS = 256;
constexpr int RSM_QUADS = 4;                                  // output quads per thread
constexpr int RSM_PAIRS = RSM_THREADS * 2 * RSM_QUADS;        // 2048 pairs = 4096 outputs = 6144 inputs per CTA
constexpr int RSM_SMEM = 3 * RSM_PAIRS + 32;                  // floats staged per CTA (halo 12 left, 14+ right)

__global__ void __launch_bounds__(RSM_THREADS)
k_resample3to2(const float* __restrict__ x, const int64_t* __restrict__ off, const char* __restrict__ len_base,
               int len_stride, float* __restrict__ y, const int64_t* __restrict__ y_off, int32_t* __restrict__ y_len) {
  __shared__ __align__(16) float sm[RSM_SMEM];
  const int c = blockIdx.x;
  const int L = *reinterpret_cast<const int32_t*>(len_base + (size_t)c * len_stride);
  const int target = L <= 0 ? 0 : (int)((2LL * L + 2) / 3);     // ceil(2L/3)
  if (blockIdx.y == 0 && threadIdx.x == 0 && y_len) y_len[c] = target;
  const int n_pairs = (target + 1) >> 1;
  const int m0 = blockIdx.y * RSM_PAIRS;
  if (m0 >= n_pairs) return;
  const float* __restrict__ xs = x + off[c];
  float* __restrict__ ys = y + y_off[c];
  // stage x[a0 .. a0 + 3*PAIRS + 32), a0 = 3*m0 - 12 (multiple of 4); bytes outside [0, L) are zero-filled
  const long long a0 = 3LL * m0 - 12;
  const bool al = (reinterpret_cast<uintptr_t>(xs) & 15u) == 0;
  for (int q = threadIdx.x; q < RSM_SMEM / 4; q += RSM_THREADS) {
    const long long g = a0 + 4 * q;
    if (al && g >= 0 && g < L) {
      cp_async16_zfill(sm + 4 * q, xs + g, (L - g >= 4) ? 16 : 4 * (int)(L - g));
    } else {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (g + 0 >= 0 && g + 0 < L) v.x = xs[g + 0];
      if (g + 1 >= 0 && g + 1 < L) v.y = xs[g + 1];
      if (g + 2 >= 0 && g + 2 < L) v.z = xs[g + 2];
      if (g + 3 >= 0 && g + 3 < L) v.w = xs[g + 3];
      *reinterpret_cast<float4*>(sm + 4 * q) = v;
    }
  }
  cp_async_commit();
  cp_async_wait_all();
  __syncthreads();
#pragma unroll
  for (int it = 0; it < RSM_QUADS; ++it) {
    // quad u -> pairs m0+2u, m0+2u+1 -> outputs 4u .. 4u+3 of the tile; needs sm[6u+2 .. 6u+28)
    const int u = threadIdx.x + it * RSM_THREADS;
    if (m0 + 2 * u >= n_pairs) break;
    const float2* sp = reinterpret_cast<const float2*>(sm + 6 * u + 2);
    float2 V[13];
#pragma unroll
    for (int k = 0; k < 13; ++k) V[k] = sp[k];
    float2 o00 = make_float2(0.f, 0.f), o10 = o00, o01 = o00, o11 = o00;
#pragma unroll
    for (int k = 0; k < 12; ++k) { o00 = __ffma2_rn(V[k], c_rA0[k], o00); o01 = __ffma2_rn(V[k], c_rA1[k], o01); }
#pragma unroll
    for (int k = 1; k < 13; ++k) { o10 = __ffma2_rn(V[k], c_rB0[k], o10); o11 = __ffma2_rn(V[k], c_rB1[k], o11); }
    const float r0 = o00.x + o00.y, r1 = o01.x + o01.y, r2 = o10.x + o10.y, r3 = o11.x + o11.y;
    const int o = 2 * (m0 + 2 * u);
    if (o + 3 < target) {
      stg_stream4(ys + o, make_float4(r0, r1, r2, r3));
    } else {
      if (o + 0 < target) ys[o + 0] = r0;
      if (o + 1 < target) ys[o + 1] = r1;
      if (o + 2 < target) ys[o + 2] = r2;
      if (o + 3 < target) ys[o + 3] = r3;
    }
  }
}

// ---------------------------------------------------------------------------------------------- any ratio
// torchaudio.functional.resample(x, orig_freq, new_freq) for a reduced ratio orig:new -- the speed control of
// BaseTTS._apply_speed_pitch (base_tts.py:631-637: resample(audio, int(sr * speed), sr)).
//   y[new*m + p] = sum_i xpad[orig*m + i] * k[p][i],  xpad[j] = x[j - width] (0 outside [0, L)),  K = 2*width + orig,
//   truncated to ceil(new * L / orig) samples.
// One CTA per M_tile input blocks: the orig*M_tile + K samples it needs are staged in shared memory (zero fill at the
// clip ends), taps come from L1/L2 (the table is [new][K] fp32, 47 KB for the 103:100 of speed 1.03).
constexpr int RSG_THREADS = 256;
constexpr int RSG_SMEM_FLOATS = 8192;                         // staged inputs per CTA (32 KB)

__global__ void __launch_bounds__(RSG_THREADS)
k_resample_general(const float* __restrict__ x, const int64_t* __restrict__ off, const char* __restrict__ len_base,
                   int len_stride, float* __restrict__ y, const int64_t* __restrict__ y_off,
                   int32_t* __restrict__ y_len, int orig, int nw, int width, int K, const float* __restrict__ taps,
                   int m_tile) {
  __shared__ float sm[RSG_SMEM_FLOATS];
  const int c = blockIdx.x;
  const long long L = *reinterpret_cast<const int32_t*>(len_base + (size_t)c * len_stride);
  const long long target = L <= 0 ? 0 : ((long long)nw * L + orig - 1) / orig;     // ceil(new * L / orig)
  if (blockIdx.y == 0 && threadIdx.x == 0 && y_len) y_len[c] = (int32_t)target;
  const long long n_blocks = (target + nw - 1) / nw;
  const long long m0 = (long long)blockIdx.y * m_tile;
  if (m0 >= n_blocks) return;
  const float* __restrict__ xs = x + off[c];
  float* __restrict__ ys = y + y_off[c];
  const int mt = (int)min((long long)m_tile, n_blocks - m0);
  const int count = orig * (mt - 1) + K;
  const long long a0 = (long long)orig * m0 - width;
  for (int i = threadIdx.x; i < count; i += RSG_THREADS) {
    const long long g = a0 + i;
    sm[i] = (g >= 0 && g < L) ? xs[g] : 0.f;
  }
  __syncthreads();
  const int n_out = mt * nw;
  for (int ol = threadIdx.x; ol < n_out; ol += RSG_THREADS) {
    const int m = ol / nw, p = ol - m * nw;
    const long long o = (long long)nw * m0 + ol;
    if (o >= target) continue;
    const float* __restrict__ s = sm + orig * m;
    const float* __restrict__ k = taps + (size_t)p * K;
    float a0f = 0.f, a1f = 0.f;
    int i = 0;
    for (; i + 1 < K; i += 2) { a0f = fmaf(s[i], __ldg(k + i), a0f); a1f = fmaf(s[i + 1], __ldg(k + i + 1), a1f); }
    if (i < K) a0f = fmaf(s[i], __ldg(k + i), a0f);
    ys[o] = a0f + a1f;
  }
}

cudaError_t launch_resample_general(const float* x, const int64_t* off, const int32_t* len, int len_stride_bytes,
                                    int n, int64_t max_len, int orig, int nw, int width, const float* taps,
                                    float* y, const int64_t* y_off, int32_t* y_len, cudaStream_t st, LaunchCtx* lc) {
  if (n <= 0) return cudaSuccess;
  const int K = 2 * width + orig;
  if (K + orig > RSG_SMEM_FLOATS) return cudaErrorInvalidValue;
  const int m_tile = (RSG_SMEM_FLOATS - K) / orig + 1;
  const int64_t target = ((int64_t)nw * max_len + orig - 1) / orig;
  const int64_t n_blocks = (target + nw - 1) / nw;
  unsigned tiles = (unsigned)((n_blocks + m_tile - 1) / m_tile);
  if (tiles == 0) tiles = 1;
  if (tiles > 65535u) return cudaErrorInvalidValue;
  lc->begin(KID_RESAMPLE_GENERAL, st);
  k_resample_general<<<dim3((unsigned)n, tiles), RSG_THREADS, 0, st>>>(
      x, off, reinterpret_cast<const char*>(len), len_stride_bytes ? len_stride_bytes : (int)sizeof(int32_t), y, y_off,
      y_len, orig, nw, width, K, taps, m_tile);
  lc->end(st);
  return cudaGetLastError();
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
