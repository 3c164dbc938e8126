// MFCC statistics of the drift classifier's feature vector (SURVEY.md 8f NEXT-3; validation/classifier/trainer.py:50-52):
//   mfcc = librosa.feature.mfcc(y=y, sr=16000, n_mfcc=13);  mean(mfcc, axis=1), std(mfcc, axis=1)
// for a ragged batch of 16 kHz clips, librosa >= 0.10 semantics (oracle/mfcc.py states them and how they are pinned):
//   k_mfcc_frames : STFT 2048 / hop 512 (periodic hann, centred, ZERO padding) -> |X|^2 -> 128 slaney mel bands ->
//                   10 log10(max(1e-10, .)) to HBM, and the clip maximum (ordered-int atomicMax)
//   k_mfcc_stats  : max(., clip max - 80) -> orthonormal DCT-II, first 13 -> mean and population std over the frames
// One warp per frame.  The 2048-point real transform is a 1024-point complex one of the even / odd packed samples, done
// as four 256-point warp FFTs (warp_fft.cuh: butterflies in registers) of the decimated sequences z[4m + r] and one
// in-place radix-4 combine with W1024 twiddles; the real spectrum is unpacked in (k, 1024 - k) pairs straight into
// power values.
#include <cmath>
#include "kernels.h"
#include "warp_fft.cuh"

namespace rho {

constexpr int MF_NFFT = 2048;
constexpr int MF_HOP = 512;
constexpr int MF_BINS = 1025;
constexpr int MF_MELS = 128;
constexpr int MF_NCOEF = 13;
#ifndef RHO_MF_WARPS
#define RHO_MF_WARPS 8
#endif
constexpr int MF_WARPS = RHO_MF_WARPS;
constexpr int MF_FPW = 2;                                 // frames per warp
constexpr int MF_Z = 4 * PV_E_SIZE;                       // four natural-order (padded) 256-point spectra
constexpr int MF_WARP_FLOATS = 2 * (PV_E_SIZE + MF_Z) + 4;      // exchange buffer, spectra (the power values replace them in place), P[1024]
constexpr int MF_SMEM = (int)(sizeof(float2) * PV_TW_SIZE + sizeof(float) * MF_WARPS * MF_WARP_FLOATS);   // + the filterbank's non-zeros

__device__ __forceinline__ int mf_frames(long long n) { return (int)(1 + n / MF_HOP); }

__global__ void __launch_bounds__(32 * MF_WARPS)
k_mfcc_frames(const float* __restrict__ x, const int64_t* __restrict__ off, const char* __restrict__ len_base,
              int len_stride, MfccTables tb, float* __restrict__ db, long long db_stride, int* __restrict__ clip_max) {
  extern __shared__ __align__(16) unsigned char mf_smem[];
  float2* tw = reinterpret_cast<float2*>(mf_smem);
  float* warp_base = reinterpret_cast<float*>(tw + PV_TW_SIZE);
  const int c = blockIdx.y;
  const long long n = *reinterpret_cast<const int32_t*>(len_base + (size_t)c * len_stride);
  const int T = mf_frames(n > 0 ? n : 0);
  const int f0 = blockIdx.x * (MF_WARPS * MF_FPW);
  if (f0 >= T) return;
  pv_fill_twiddles<false>(tb.w256, tw);
  // the filterbank's non-zero weights, once per CTA: every lane walks a different band, from global memory that was
  // one 32-byte sector per lane and step (half of the kernel's L1 traffic)
  float* __restrict__ s_melw = warp_base + (size_t)MF_WARPS * MF_WARP_FLOATS;
  for (int i = threadIdx.x; i < tb.mel_nnz; i += 32 * MF_WARPS) s_melw[i] = __ldg(tb.mel_w + i);
  __syncthreads();
  const float2* __restrict__ tw1 = tw;
  const float2* __restrict__ tw2 = tw + 256;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float2* __restrict__ E = reinterpret_cast<float2*>(warp_base + (size_t)w * MF_WARP_FLOATS);
  float2* __restrict__ Z = E + PV_E_SIZE;                 // Z[r * PV_E_SIZE + pv_nat(k)]
  float* __restrict__ P = reinterpret_cast<float*>(Z + MF_Z);
  const float* __restrict__ xs = x + off[c];
  const bool al16 = (((uintptr_t)xs) & 15u) == 0;
  const int k1 = lane >> 2, q = lane & 3, bq = ((q & 1) << 1) | (q >> 1);
  float lmax = -INFINITY;
  for (int it = 0; it < MF_FPW; ++it) {
    const int f = f0 + it * MF_WARPS + w;
    if (f >= T) break;                                    // warp-uniform
    const long long base = (long long)f * MF_HOP - MF_NFFT / 2;     // a multiple of 4
    // ---- the windowed frame, read once in 128-bit pieces: z[j] = (x[2 j], x[2 j + 1]) * hann goes to the slot of the
    // decimated sequence it belongs to, Z[(j & 3)][j >> 2] -- pass r below reads its 256 inputs from the region it later
    // writes its spectrum to (all reads of a pass are in registers before its first write)
#pragma unroll 4
    for (int t = 0; t < MF_NFFT / 128; ++t) {
      const int s0 = 4 * (lane + 32 * t);
      const long long i0 = base + s0;
      float4 xv;
      if (al16 && i0 >= 0 && i0 + 3 < n) {
        xv = *reinterpret_cast<const float4*>(xs + i0);
      } else {
        xv.x = (i0 >= 0 && i0 < n) ? xs[i0] : 0.f;             xv.y = (i0 + 1 >= 0 && i0 + 1 < n) ? xs[i0 + 1] : 0.f;
        xv.z = (i0 + 2 >= 0 && i0 + 2 < n) ? xs[i0 + 2] : 0.f; xv.w = (i0 + 3 >= 0 && i0 + 3 < n) ? xs[i0 + 3] : 0.f;
      }
      const float4 h = __ldg(reinterpret_cast<const float4*>(tb.hann) + (lane + 32 * t));
      const int j = s0 >> 1;                                  // even: slots j & 3 in {0, 2} and {1, 3}, same j >> 2
      Z[(j & 3) * PV_E_SIZE + (j >> 2)] = make_float2(xv.x * h.x, xv.y * h.y);
      Z[((j + 1) & 3) * PV_E_SIZE + (j >> 2)] = make_float2(xv.z * h.z, xv.w * h.w);
    }
    __syncwarp();
    // ---- four 256-point FFTs of z_r[m] = z[4 m + r]
    for (int r = 0; r < 4; ++r) {
      float2* __restrict__ Zr = Z + r * PV_E_SIZE;
      float2 v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = Zr[lane + 32 * j];
      warp_fft256<false>(v, E, tw1, tw2, lane);
#pragma unroll
      for (int a = 0; a < 8; ++a) Zr[k1 + 8 * a + 72 * bq] = v[a];      // pv_nat(k1 + 8 a + 64 b)
      __syncwarp();
    }
    // ---- Z[k + 256 q] = sum_r W1024^(r k) Z_r[k] W4^(r q): in place on the four slots of k
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int k = lane + 32 * t, pk = pv_nat(k);
      const float2 a0 = Z[pk];
      const float2 a1 = cmul(Z[PV_E_SIZE + pk], __ldg(tb.w1024 + k));
      const float2 a2 = cmul(Z[2 * PV_E_SIZE + pk], __ldg(tb.w1024 + 2 * k));
      const float2 a3 = cmul(Z[3 * PV_E_SIZE + pk], __ldg(tb.w1024 + 3 * k));
      radix4<false>(a0, a1, a2, a3, Z[pk], Z[PV_E_SIZE + pk], Z[2 * PV_E_SIZE + pk], Z[3 * PV_E_SIZE + pk]);
    }
    __syncwarp();
    // ---- real spectrum from the packed one, as powers:  X[K] = E + W2048^K O,  X[1024 - K] = conj(E - W2048^K O)
    // The power of bin K replaces the real part of the slot Z[K] lived in: a lane reads the pair (K, 1024 - K) and writes
    // the same two slots, so no second buffer is needed; only P[1024] has no slot of its own.
    auto zat = [&](int K) { return Z[K + 8 * (K >> 6)]; };            // (K >> 8) * PV_E_SIZE + pv_nat(K & 255)
    float* __restrict__ Pf = reinterpret_cast<float*>(Z);
    auto pslot = [&](int K) -> float& { return K == 1024 ? P[0] : Pf[2 * (K + 8 * (K >> 6))]; };
#pragma unroll 4
    for (int t = 0; t < 16; ++t) {
      const int K = lane + 32 * t;                        // 0..511, partner 1024 - K
      const float2 zk = zat(K), zc = zat((1024 - K) & 1023);
      const float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y - zc.y));
      const float2 o = make_float2(0.5f * (zk.y + zc.y), -0.5f * (zk.x - zc.x));
      const float2 wo = cmul(__ldg(tb.w2048 + K), o);
      const float ar = e.x + wo.x, ai = K ? e.y + wo.y : 0.f;
      const float br = e.x - wo.x, bi = K ? e.y - wo.y : 0.f;
      pslot(K) = ar * ar + ai * ai;
      pslot(1024 - K) = br * br + bi * bi;
    }
    if (lane == 0) {                                      // K = 512 is its own partner: X = Re z - i Im z ... |X|^2 = |z|^2
      const float2 z = zat(512);
      const float2 wo = cmul(__ldg(tb.w2048 + 512), make_float2(z.y, 0.f));
      const float ar = z.x + wo.x, ai = wo.y;
      pslot(512) = ar * ar + ai * ai;
    }
    __syncwarp();
    // ---- 128 mel bands (sparse triangles), dB
    float* __restrict__ row = db + (size_t)c * db_stride + (size_t)f * MF_MELS;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = lane + 32 * j;
      const int lo = tb.mel_lo[m], cnt = tb.mel_cnt[m];
      const float* __restrict__ wv = s_melw + tb.mel_wofs[m];
      float acc = 0.f;
      for (int i = 0; i < cnt; ++i) acc = fmaf(wv[i], pslot(lo + i), acc);
      const float d = 10.0f * log10f(fmaxf(1e-10f, acc));
      row[m] = d;
      lmax = fmaxf(lmax, d);
    }
    __syncwarp();
  }
  lmax = warp_max(lmax);
  if (lane == 0 && lmax > -INFINITY) atomicMax(&clip_max[c], float_to_ordered(lmax));
}

__global__ void k_mfcc_init(int* __restrict__ clip_max, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) clip_max[i] = float_to_ordered(-INFINITY);
}

// One CTA per clip: max(dB, clip max - 80) -> DCT-II (13 x 128, orthonormal) -> sum and sum of squares per
// coefficient in double -> mean, population standard deviation.
__global__ void __launch_bounds__(128)
k_mfcc_stats(const float* __restrict__ db, long long db_stride, const char* __restrict__ len_base, int len_stride,
             const int* __restrict__ clip_max, const float* __restrict__ g_dct, float* __restrict__ out) {
  __shared__ float D[MF_NCOEF * MF_MELS];
  __shared__ double red[2][4][MF_NCOEF];
  const int c = blockIdx.x;
  const long long n = *reinterpret_cast<const int32_t*>(len_base + (size_t)c * len_stride);
  const int T = mf_frames(n > 0 ? n : 0);
  for (int i = threadIdx.x; i < MF_NCOEF * MF_MELS; i += 128) D[i] = g_dct[i];
  __syncthreads();
  const float floor_db = ordered_to_float(clip_max[c]) - 80.0f;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* __restrict__ base = db + (size_t)c * db_stride;
  double s1 = 0.0, s2 = 0.0;                              // lane k < 13 accumulates coefficient k
  for (int f = w; f < T; f += 4) {
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = fmaxf(base[(size_t)f * MF_MELS + lane + 32 * j], floor_db);
#pragma unroll
    for (int k = 0; k < MF_NCOEF; ++k) {
      float p = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) p = fmaf(D[k * MF_MELS + lane + 32 * j], v[j], p);
      p = warp_sum(p);
      if (lane == k) { s1 += (double)p; s2 += (double)p * (double)p; }
    }
  }
  if (lane < MF_NCOEF) { red[0][w][lane] = s1; red[1][w][lane] = s2; }
  __syncthreads();
  if (threadIdx.x < MF_NCOEF) {
    const int k = threadIdx.x;
    const double a = red[0][0][k] + red[0][1][k] + red[0][2][k] + red[0][3][k];
    const double b = red[1][0][k] + red[1][1][k] + red[1][2][k] + red[1][3][k];
    const double mean = a / T;
    const double var = fmax(0.0, b / T - mean * mean);
    out[(size_t)c * 2 * MF_NCOEF + k] = (float)mean;
    out[(size_t)c * 2 * MF_NCOEF + MF_NCOEF + k] = (float)sqrt(var);
  }
}

size_t mfcc_workspace_bytes(int n, int64_t max_len) {
  if (n <= 0) return 0;
  const int64_t T = 1 + (max_len > 0 ? max_len : 0) / MF_HOP;
  return align_up((size_t)n * (size_t)T * MF_MELS * sizeof(float), 256) + align_up((size_t)n * sizeof(int), 256);
}

cudaError_t launch_mfcc_stats(const MfccTables& tb, const float* x, const int64_t* off, const int32_t* len,
                              int len_stride_bytes, int n, int64_t max_len, float* out, void* workspace,
                              cudaStream_t st, LaunchCtx* lc) {
  if (n <= 0) return cudaSuccess;
  if (n > 65535) return cudaErrorInvalidValue;
  const int64_t T = 1 + (max_len > 0 ? max_len : 0) / MF_HOP;
  float* db = (float*)workspace;
  int* clip_max = (int*)((char*)workspace + align_up((size_t)n * (size_t)T * MF_MELS * sizeof(float), 256));
  const char* lb = reinterpret_cast<const char*>(len);
  const int ls = len_stride_bytes ? len_stride_bytes : (int)sizeof(int32_t);
  const int smem = MF_SMEM + (int)align_up((size_t)tb.mel_nnz * sizeof(float), 16);
  cudaError_t e = cudaFuncSetAttribute(k_mfcc_frames, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  lc->begin(KID_MFCC_FRAMES, st);
  k_mfcc_init<<<(n + 255) / 256, 256, 0, st>>>(clip_max, n);
  const unsigned gx = (unsigned)((T + MF_WARPS * MF_FPW - 1) / (MF_WARPS * MF_FPW));
  k_mfcc_frames<<<dim3(gx, (unsigned)n), 32 * MF_WARPS, smem, st>>>(x, off, lb, ls, tb, db, (long long)T * MF_MELS,
                                                                     clip_max);
  lc->end(st);
  lc->begin(KID_MFCC_STATS, st);
  k_mfcc_stats<<<(unsigned)n, 128, 0, st>>>(db, (long long)T * MF_MELS, lb, ls, clip_max, tb.dct, out);
  lc->end(st);
  return cudaGetLastError();
}

}  // namespace rho
