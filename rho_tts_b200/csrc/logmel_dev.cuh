// Device-side building blocks of the log-mel front end, shared by the standalone kernel (logmel.cu)
// and the fused post-process -> resample -> log-mel kernel (fused.cu).
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace rho {

constexpr int LM_GROUPS = 16;                       // FFTs per batch
constexpr int LM_LANES = 20;                        // threads per FFT
constexpr int LM_THREADS = LM_GROUPS * LM_LANES;    // 320
constexpr int LM_BF = 2 * LM_GROUPS;                // frames per batch (32)
#ifndef RHO_LM_BATCHES
#define RHO_LM_BATCHES 8
#endif
constexpr int LM_BATCHES = RHO_LM_BATCHES;          // batches per CTA
constexpr int LM_TILE = LM_BF * LM_BATCHES;         // frames per CTA of k_logmel_frames (256; 4 / 8 / 16 batches: 3.90 / 3.80 / 3.70 ms on C3v)
constexpr int LM_SLAB = HOP16 * LM_BF + (N_FFT - HOP16);  // 5360 samples cover 32 frames
// The slab is stored in blocks of 320 samples (one frame pair's hop) at a stride of 340 floats and the
// FFT buffers at a stride of 420 float2, so that the shared-memory bank of every stage-1/stage-2 access
// is (thread id + const) mod 32: conflict-free although a 20-thread FFT group straddles warps.
constexpr int LM_SLAB_BLK = 2 * HOP16;              // 320
constexpr int LM_SLAB_STRIDE = LM_SLAB_BLK + 20;    // 340
constexpr int LM_SLAB_SM = ((LM_SLAB + LM_SLAB_BLK - 1) / LM_SLAB_BLK) * LM_SLAB_STRIDE;   // 17 * 340
constexpr int LM_FB = 420;                          // float2 per FFT buffer (20 rows x 21)
constexpr int LM_PS = 201;                          // power row stride (odd: conflict-free across frames)

#include "mel_sparse_gen.inc"
static_assert(MEL_PARTS == LM_THREADS / 32, "one mel part per warp");


// Complex arithmetic on the sm_100 packed-fp32 pipe: one FADD2 / FMUL2 / FFMA2 handles the real and the
// imaginary part together (halves the issue slots of the butterflies; same IEEE roundings as scalar code).
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.f, -1.f), a); }
__device__ __forceinline__ float2 cscale(float s, float2 a) { return __fmul2_rn(make_float2(s, s), a); }
__device__ __forceinline__ float2 cfma(float s, float2 a, float2 c) { return __ffma2_rn(make_float2(s, s), a, c); }  // s*a + c
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// forward 5-point DFT, in place
__device__ __forceinline__ void dft5(float2& x0, float2& x1, float2& x2, float2& x3, float2& x4) {
  const float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f;   // cos(2pi/5), cos(4pi/5)
  const float s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;    // sin(2pi/5), sin(4pi/5)
  const float2 t1 = cadd(x1, x4), t2 = cadd(x2, x3), t3 = csub(x1, x4), t4 = csub(x2, x3);
  const float2 m1 = cfma(c2, t2, cfma(c1, t1, x0));
  const float2 m2 = cfma(c1, t2, cfma(c2, t1, x0));
  const float2 u1 = cfma(s2, t4, cscale(s1, t3));
  const float2 u2 = cfma(-s1, t4, cscale(s2, t3));
  x0 = cadd(x0, cadd(t1, t2));
  // y1 = m1 - i*u1, y4 = m1 + i*u1, y2 = m2 - i*u2, y3 = m2 + i*u2   ( -i*(a,b) = (b,-a) )
  x1 = make_float2(m1.x + u1.y, m1.y - u1.x);
  x4 = make_float2(m1.x - u1.y, m1.y + u1.x);
  x2 = make_float2(m2.x + u2.y, m2.y - u2.x);
  x3 = make_float2(m2.x - u2.y, m2.y + u2.x);
}

// forward 4-point DFT, in place
__device__ __forceinline__ void dft4(float2& x0, float2& x1, float2& x2, float2& x3) {
  const float2 a = cadd(x0, x2), b = csub(x0, x2), c = cadd(x1, x3), d = csub(x1, x3);
  x0 = cadd(a, c);
  x2 = csub(a, c);
  x1 = make_float2(b.x + d.y, b.y - d.x);   // b - i*d
  x3 = make_float2(b.x - d.y, b.y + d.x);   // b + i*d
}

// forward 20-point DFT, in place, Good-Thomas 4x5 (no internal twiddles):
//   input index n = (5a + 4b) mod 20, output index k = (5*k1 + 16*k2) mod 20.
__device__ __forceinline__ void dft20(float2 (&v)[20]) {
#ifdef FZ_PROBE_CHEAP_DFT       // timing probe only (WRONG results): four 5-point DFTs instead of the 20-point one
  dft5(v[0], v[4], v[8], v[12], v[16]); dft5(v[1], v[5], v[9], v[13], v[17]);
  dft5(v[2], v[6], v[10], v[14], v[18]); dft5(v[3], v[7], v[11], v[15], v[19]);
  return;
#endif
  float2 T[4][5];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
#pragma unroll
    for (int b = 0; b < 5; ++b) T[a][b] = v[(5 * a + 4 * b) % 20];
    dft5(T[a][0], T[a][1], T[a][2], T[a][3], T[a][4]);
  }
#pragma unroll
  for (int k2 = 0; k2 < 5; ++k2) {
    dft4(T[0][k2], T[1][k2], T[2][k2], T[3][k2]);
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) v[(5 * k1 + 16 * k2) % 20] = T[k1][k2];
  }
}

// sample i of the (zero-padded to N, reflect-extended) 16 kHz clip
__device__ __forceinline__ float lm_sample(const float* __restrict__ xs, long long i, int n_valid, int N) {
  if (i < 0) i = -i;
  if (i >= N) i = 2LL * (N - 1) - i;
  return (i >= 0 && i < n_valid) ? xs[i] : 0.f;
}

// number of frames of clip with n16 samples, and how many of them see any signal
// (x + 4) / 4 of the Whisper normalisation (x / 4 == x * 0.25 bit for bit).  The frame kernels store log-mel values
// already scaled: the scaling is monotone, so max(x, m - 8) followed by it equals max(scaled x, scaled (m - 8)) bit for
// bit, and k_logmel_norm only has to WRITE the values the clamp changes.
__device__ __forceinline__ float lm_scaled(float x) { return __fmul_rn(__fadd_rn(x, 4.0f), 0.25f); }

__device__ __forceinline__ void lm_frame_counts(int n16, int pad_frames, int* T, int* T_real, int* N, int* n_valid) {
  if (pad_frames > 0) {
    *N = pad_frames * HOP16;
    *n_valid = n16 < *N ? (n16 > 0 ? n16 : 0) : *N;
    *T = pad_frames;
    int tz = (*n_valid + (N_FFT / 2) + HOP16 - 1) / HOP16;   // first frame whose window starts past the signal
    if (tz < 2) tz = 2;
    *T_real = tz < *T ? tz : *T;
  } else {
    *N = n16; *n_valid = n16;
    *T = (n16 > N_FFT / 2) ? n16 / HOP16 : 0;               // reflect padding needs > n_fft/2 samples
    *T_real = *T;
  }
}


}  // namespace rho
