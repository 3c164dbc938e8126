// Fused "apply + features" kernel for one-segment items (the post-process of single clips):
//   y = (x[start:end] - dc) * fades                      base_tts.py:394-433 (after the trim scan, :348-392)
//   decay sums over the first / last third of y          base_tts.py:297-323
//   w16 = resample(y, 24000, 16000)                      torchaudio functional.py:1305-1432
//   log10(mel(|STFT(w16)|^2)) + per-clip max             transformers feature_extraction_whisper.py:135-164
// in ONE pass over the clip: x is read once, y is written once, the 16 kHz signal never leaves shared
// memory, and only the raw log-mel frames that see signal are written (k_logmel_norm finishes them).
// Against the unfused chain (k_gather -> k_resample3to2 -> k_logmel_frames) this removes the re-read of y,
// and the write + read of the 16 kHz intermediate: 4.05 GB -> 2.0 GB of HBM traffic per 1000 x 10 s clips.
//
// One CTA per SM, two independent 320-thread halves (named barriers), each half owns a stream of
// 32-frame batches of one clip.  Per batch and half:
//   span  : x[start + 240*t0 - 312 .. +8068) staged with LDGSTS (prefetched under the previous batch's FFT)
//   apply : in place (x-dc)*fade -> y; the batch's own 7680 samples go to HBM (128-bit stores), decay sums
//   FIR   : 2 phases x 23 taps from the span -> 5360 samples of the 16 kHz slab (reflection fixed up in smem)
//   FFT   : as k_logmel_frames (two real frames per 400-point complex FFT, 20 x 20, packed fp32x2 butterflies)
//   mel   : immediate-weight FFMAs, SFU log2, ordered-int atomicMax of the clip maximum
#include <algorithm>
#include "logmel_dev.cuh"

namespace rho {

__constant__ float c_ftaps[2][RS_TAPS];

cudaError_t upload_fused_taps(const float* taps) {
  return cudaMemcpyToSymbol(c_ftaps, taps, sizeof(float) * 2 * RS_TAPS);
}

#ifndef FZ_CLIP_MAJOR
#define FZ_CLIP_MAJOR 0
#endif
constexpr int FZ_HALVES = 2;
constexpr int FZ_THREADS = LM_THREADS * FZ_HALVES;          // 640
constexpr int FZ_LEAD = 312;                                // span starts 312 samples before the batch's own range
constexpr int FZ_SPAN = 8068;                               // 24 kHz samples staged per batch (multiple of 4)
constexpr int FZ_OWN = 240 * LM_BF;                         // 7680 output samples owned by a batch
constexpr int FZ_DPAIRS = LM_SLAB / 4;                      // 1340 groups of 4 consecutive 16 kHz samples

struct alignas(16) FzHalf {
  float span[FZ_SPAN];
  float2 fb[LM_GROUPS * LM_FB];
  float pw[LM_BF * LM_PS];          // 16 kHz slab between FIR and FFT stage 1, power spectra afterwards
  double redd[2][LM_THREADS / 32];
  float redf[LM_THREADS / 32];
};
static_assert(LM_SLAB_SM <= LM_BF * LM_PS, "the slab must fit in the power buffer");

struct FzSmem {
  FzHalf h[FZ_HALVES];
  float hann[N_FFT];
  float2 tw[N_FFT];
};
static_assert(sizeof(FzSmem) <= 232448, "shared memory budget (227 KB)");

__device__ __forceinline__ void half_sync(int half) {
  asm volatile("bar.sync %0, %1;" :: "r"(half + 1), "r"(LM_THREADS) : "memory");
}

__device__ __forceinline__ float fz_fade_gain(int o, int n, int fade) {
  if (o < fade) return 0.5f * (1.f - cosf(linspace32(0.f, RHO_PI_F, fade, o)));
  if (o >= n - fade) return 0.5f * (1.f + cosf(linspace32(0.f, RHO_PI_F, fade, o - (n - fade))));
  return 1.f;
}

// Slow path of the apply step for the (few) 128-bit pieces that touch a fade, a clip edge or a third
// boundary.  Kept out of line: cosf's argument-reduction code would otherwise be inlined eight times
// into the hot loop and push it out of the instruction cache.
__device__ __noinline__ void fz_apply_edge(float* e, int o, int n, int fade, bool need_fade, float dc, int third,
                                           bool owned, float* __restrict__ ys, float* a_first, float* a_last) {
  for (int k = 0; k < 4; ++k) {
    const int oo = o + k;
    float val = 0.f;
    if (oo >= 0 && oo < n) {
      val = __fsub_rn(e[k], dc);
      if (need_fade) val = __fmul_rn(val, fz_fade_gain(oo, n, fade));
      if (owned) {
        ys[oo] = val;
        if (oo < third) *a_first += val * val;
        if (oo >= n - third) *a_last += val * val;
      }
    }
    e[k] = val;
  }
}

template <int NM>
__global__ void __launch_bounds__(FZ_THREADS, 1)
k_fused_features(const float* __restrict__ x, const int64_t* __restrict__ seg_off, const SegState* __restrict__ seg,
                 ItemState* __restrict__ item, const int32_t* __restrict__ item_first_seg,
                 float* __restrict__ y, const int64_t* __restrict__ y_off, int fade,
                 const float* __restrict__ g_hann, const float2* __restrict__ g_tw, int pad_frames,
                 float* __restrict__ mel, long long mel_stride, int* __restrict__ clip_max,
                 int32_t* __restrict__ len16_out, int tile_pairs, int n_items) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FzSmem& S = *reinterpret_cast<FzSmem*>(smem_raw);
  for (int i = threadIdx.x; i < N_FFT; i += FZ_THREADS) { S.hann[i] = g_hann[i]; S.tw[i] = g_tw[i]; }
  __syncthreads();                                   // the only CTA-wide barrier: the halves run independently from here

  // clip-major 1-D grid: the CTAs of one clip are scheduled back to back (L2 locality of its samples)
#if FZ_CLIP_MAJOR
  const int c = blockIdx.x / tile_pairs;
  const int pair = blockIdx.x - c * tile_pairs;
#else
  const int c = blockIdx.x % n_items;
  const int pair = blockIdx.x / n_items;
#endif
  const int half = threadIdx.x / LM_THREADS;
  const int tid = threadIdx.x - half * LM_THREADS;
  FzHalf& H = S.h[half];

  const int s = item_first_seg[c];
  const SegState st = seg[s];
  const int n = st.end - st.start;                   // samples of y
  const float dc = st.dc;
  const int n16 = n > 0 ? (int)((2LL * n + 2) / 3) : 0;
  int T, T_real, N, n_valid;
  lm_frame_counts(n16, pad_frames, &T, &T_real, &N, &n_valid);
  if (pair == 0 && threadIdx.x == 0) len16_out[c] = n16;
  const int t_cover = max(T_real, (n + 239) / 240);  // batches needed for the features AND to write all of y
  const int tile_t0 = (pair * FZ_HALVES + half) * LM_TILE;
  if (tile_t0 >= t_cover) return;

  const float* __restrict__ xs = x + seg_off[s] + st.start;
  float* __restrict__ ys = y + y_off[c];
  float* __restrict__ out = mel + (long long)c * NM * mel_stride;
  const int third = n / 3;
  const bool need_fade = fade > 0 && n >= 2 * fade;
  const int g = tid / LM_LANES, lane = tid - g * LM_LANES;
  float2* fb = H.fb + g * LM_FB;
  float lmax = -INFINITY;
  float a_first = 0.f, a_last = 0.f;

  auto stage_span = [&](int t0) {
    const long long j0 = 240LL * t0 - FZ_LEAD;
    for (int q = tid; q < FZ_SPAN / 4; q += LM_THREADS) {
      const long long j = j0 + 4 * q;
      int nb = 0;
      if (j >= 0 && j < n) nb = (n - j >= 4) ? 16 : 4 * (int)(n - j);
      cp_async16_zfill(H.span + 4 * q, nb ? xs + j : xs, nb);
    }
    cp_async_commit();
  };
  stage_span(tile_t0);

  for (int b = 0; b < LM_BATCHES; ++b) {
    const int t0 = tile_t0 + b * LM_BF;
    if (t0 >= t_cover) break;
    const bool next = (b + 1 < LM_BATCHES) && (t0 + LM_BF < t_cover);
    cp_async_wait_all();
    half_sync(half);
    // ---- apply: span <- (x - dc) * fade; the batch's own samples go to HBM; decay sums
    {
      const int j0 = 240 * t0 - FZ_LEAD;
      const int own_lo = 240 * t0, own_hi = own_lo + FZ_OWN;
      // thirds as 128-bit-piece ranges: pieces entirely inside the first / last third take the fast path
      for (int q = tid; q < FZ_SPAN / 4; q += LM_THREADS) {
        const int o = j0 + 4 * q;
        if (o + 3 < 0 || o >= n) continue;                    // zero-filled: stays zero
        float4 v = *reinterpret_cast<float4*>(H.span + 4 * q);
        const bool owned = o >= own_lo && o < own_hi;         // own_lo, own_hi and o are multiples of 4
        const bool in_first = o + 3 < third, in_last = o >= n - third;
        const bool clean = o >= 0 && o + 3 < n && (!need_fade || (o >= fade && o + 3 < n - fade)) &&
                           (in_first || o >= third) && (in_last || o + 3 < n - third);
        if (clean) {
          v.x = __fsub_rn(v.x, dc); v.y = __fsub_rn(v.y, dc); v.z = __fsub_rn(v.z, dc); v.w = __fsub_rn(v.w, dc);
          if (owned) {
            stg_stream4(ys + o, v);
            const float ss = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
            if (in_first) a_first += ss;
            if (in_last) a_last += ss;
          }
        } else {
          float e[4] = {v.x, v.y, v.z, v.w};
          fz_apply_edge(e, o, n, fade, need_fade, dc, third, owned, ys, &a_first, &a_last);
          v = make_float4(e[0], e[1], e[2], e[3]);
        }
        *reinterpret_cast<float4*>(H.span + 4 * q) = v;
      }
    }
    half_sync(half);
    if (t0 >= T_real) {                                      // a batch that only had output samples left to write
      if (next) stage_span(t0 + LM_BF);
      continue;
    }
    // ---- FIR 24k -> 16k: slab position i holds w16[w0 + i], w0 = 160*t0 - 200;
    //      w16[2m+p] = sum_t y[3m - 10 + t] * k[p][t]  ->  span offset 2 + 3*(i/2) + t
    const int w0 = HOP16 * t0 - N_FFT / 2;
    float* slab = H.pw;
    for (int d = tid; d < FZ_DPAIRS; d += LM_THREADS) {
      const float2* sp = reinterpret_cast<const float2*>(H.span + 2 + 6 * d);
      float v[26];
#pragma unroll
      for (int k = 0; k < 13; ++k) { const float2 t2 = sp[k]; v[2 * k] = t2.x; v[2 * k + 1] = t2.y; }
      float o00 = 0.f, o01 = 0.f, o10 = 0.f, o11 = 0.f;
      // taps 0, 20..22 of phase 0 and 0..2, 21..22 of phase 1 sit on the window clamp (|k| ~ 3e-24): skipped
#pragma unroll
      for (int i = 1; i < 20; ++i) { o00 = fmaf(v[i], c_ftaps[0][i], o00); o10 = fmaf(v[i + 3], c_ftaps[0][i], o10); }
#pragma unroll
      for (int i = 3; i < 21; ++i) { o01 = fmaf(v[i], c_ftaps[1][i], o01); o11 = fmaf(v[i + 3], c_ftaps[1][i], o11); }
      const int wi = w0 + 4 * d;
      float4 r;
      r.x = (wi + 0 < n_valid) ? o00 : 0.f;
      r.y = (wi + 1 < n_valid) ? o01 : 0.f;
      r.z = (wi + 2 < n_valid) ? o10 : 0.f;
      r.w = (wi + 3 < n_valid) ? o11 : 0.f;
      *reinterpret_cast<float4*>(slab + 4 * d + 20 * (d / (LM_SLAB_BLK / 4))) = r;
    }
    half_sync(half);
    // ---- reflect padding of torch.stft(center=True): indices < 0 and >= N mirror the computed ones
    const bool left = w0 < 0, right = (long long)w0 + LM_SLAB > N;
    if (left || right) {
      if (left) {
        for (int i = tid; i < -w0; i += LM_THREADS) {
          const int src = -(w0 + i) - w0;                    // slab position of w16[-(w0+i)]
          slab[i + 20 * (i / LM_SLAB_BLK)] = slab[src + 20 * (src / LM_SLAB_BLK)];
        }
      }
      if (right) {
        const int i_lo = max(0, N - w0);
        for (int i = i_lo + tid; i < LM_SLAB; i += LM_THREADS) {
          const long long wsrc = 2LL * (N - 1) - (w0 + i);
          const int src = (int)(wsrc - w0);
          float val = 0.f;
          if (wsrc >= 0 && wsrc < n_valid && src >= 0 && src < LM_SLAB) val = slab[src + 20 * (src / LM_SLAB_BLK)];
          slab[i + 20 * (i / LM_SLAB_BLK)] = val;
        }
      }
      half_sync(half);
    }
    // ---- FFT stage 1: lane = n2, 20-point DFT over n1 of z[20*n1 + n2], then twiddle W400^(n2*k1)
    float2 v[20];
    {
      const float* fa = slab + g * LM_SLAB_STRIDE + lane;
      const float* fbm = fa + HOP16;
#pragma unroll
      for (int n1 = 0; n1 < 20; ++n1) {
        const float h = S.hann[20 * n1 + lane];
        v[n1] = make_float2(fa[20 * n1 + (n1 >= 16 ? 20 : 0)] * h, fbm[20 * n1 + (n1 >= 8 ? 20 : 0)] * h);
      }
    }
    dft20(v);
#pragma unroll
    for (int k1 = 0; k1 < 20; ++k1) fb[k1 * 21 + lane] = (k1 == 0) ? v[0] : cmul(v[k1], S.tw[k1 * 20 + lane]);
    half_sync(half);
    if (next) stage_span(t0 + LM_BF);                        // span and slab are dead: prefetch under stage 2 / mel
    // ---- FFT stage 2: lane = k1, 20-point DFT over n2 -> Z[k1 + 20*k2]
#pragma unroll
    for (int n2 = 0; n2 < 20; ++n2) v[n2] = fb[lane * 21 + n2];
    dft20(v);
    half_sync(half);
#pragma unroll
    for (int k2 = 0; k2 < 20; ++k2) fb[lane + 20 * k2] = v[k2];
    half_sync(half);
    // ---- split the two real spectra and take |.|^2
    {
      float* pa = H.pw + (2 * g) * LM_PS;
      float* pb = pa + LM_PS;
#pragma unroll
      for (int j = 0; j < 11; ++j) {
        const int k = lane + 20 * j;
        if (k <= N_FFT / 2) {
          const float2 z = fb[k];
          const float2 w = fb[k == 0 ? 0 : N_FFT - k];
          const float ar = z.x + w.x, ai = z.y - w.y, br = z.x - w.x, bi = z.y + w.y;
          pa[k] = 0.25f * (ar * ar + ai * ai);
          pb[k] = 0.25f * (br * br + bi * bi);
        }
      }
    }
    half_sync(half);
    // ---- mel projection + log10
    {
      const int f = tid & 31, part = tid >> 5;
      const int t = t0 + f;
      const bool live = t < T_real;
      const float* p = H.pw + f * LM_PS;
      float* __restrict__ o = out + t;
      auto emit = [&](int m, float acc) {
        const float ls = __log2f(fmaxf(acc, 1e-10f)) * 0.30102999566398120f;
        if (live) {
          o[(long long)m * mel_stride] = ls;
          lmax = fmaxf(lmax, ls);
        }
      };
      if (NM == 80) mel_sparse_80(part, p, emit); else mel_sparse_128(part, p, emit);
    }
    // the next iteration's first barrier orders these reads before pw is overwritten
  }
  // ---- per-half reductions: clip max (ordered-int atomicMax), decay sums (double atomics)
  lmax = warp_max(lmax);
  const double df = warp_sum((double)a_first), dl = warp_sum((double)a_last);
  half_sync(half);
  if ((tid & 31) == 0) { H.redf[tid >> 5] = lmax; H.redd[0][tid >> 5] = df; H.redd[1][tid >> 5] = dl; }
  half_sync(half);
  if (tid < 32) {
    constexpr int NW = LM_THREADS / 32;
    float m = tid < NW ? H.redf[tid] : -INFINITY;
    double a = tid < NW ? H.redd[0][tid] : 0.0, bq = tid < NW ? H.redd[1][tid] : 0.0;
    m = warp_max(m); a = warp_sum(a); bq = warp_sum(bq);
    if (tid == 0) {
      if (m > -INFINITY) atomicMax(&clip_max[c], float_to_ordered(m));
      if (a != 0.0) atomicAdd(&item[c].s_first, a);
      if (bq != 0.0) atomicAdd(&item[c].s_last, bq);
    }
  }
}

cudaError_t launch_fused_features(const Tables& tb, const float* x, const int64_t* seg_off, const Workspace& ws,
                                  const int32_t* item_first_seg, int n_items, int64_t max_len,
                                  const Derived& d, float* y, const int64_t* y_off, int n_mels, int pad_frames,
                                  float* mel, int64_t mel_stride_frames, cudaStream_t st, LaunchCtx* lc) {
  if (n_items <= 0) return cudaSuccess;
  // frames to cover: the features' frames and, for clips the 30 s window truncates, the rest of y
  const int64_t max16 = (2 * max_len + 2) / 3;
  int64_t max_real;
  if (pad_frames > 0) {
    const int64_t nv = max16 < (int64_t)pad_frames * HOP16 ? max16 : (int64_t)pad_frames * HOP16;
    max_real = (nv + N_FFT / 2 + HOP16 - 1) / HOP16;
    if (max_real < 2) max_real = 2;
    if (max_real > pad_frames) max_real = pad_frames;
  } else {
    max_real = max16 / HOP16;
  }
  const int64_t cover = std::max<int64_t>(max_real, (max_len + 239) / 240);
  unsigned tiles = (unsigned)((cover + LM_TILE - 1) / LM_TILE);
  if (tiles == 0) tiles = 1;
  const unsigned gy = (tiles + FZ_HALVES - 1) / FZ_HALVES;
  const size_t smem = sizeof(FzSmem);
  auto kern = (n_mels == 80) ? k_fused_features<80> : k_fused_features<128>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if ((uint64_t)n_items * gy > 0x7fffffffull) return cudaErrorInvalidValue;
  dim3 grid((unsigned)n_items * gy);
  lc->begin(KID_FUSED, st);
  kern<<<grid, FZ_THREADS, smem, st>>>(x, seg_off, ws.seg, ws.item, item_first_seg, y, y_off, d.fade, tb.hann,
                                       tb.twiddle, pad_frames, mel, mel_stride_frames, ws.clip_max, ws.len16,
                                       (int)gy, n_items);
  lc->end(st);
  return cudaGetLastError();
}

}  // namespace rho
