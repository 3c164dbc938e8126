// Whisper-style log-mel front end: framed STFT -> power -> mel -> log10 -> clip-max clamp -> scale.
//
// Reference: transformers WhisperFeatureExtractor (reached from the reference through
// src/rho_tts/validation/stt/stt_validator.py:78-107):
//   transformers/models/whisper/feature_extraction_whisper.py:135-164 (_torch_extract_fbank_features),
//   :296-303 (pad / truncate to 480 000 samples), transformers/audio_utils.py:453-544 (filterbank).
//
// Kernel 1 (k_logmel_frames): one CTA = 16 groups of 20 threads.  Each group runs one 400-point
// complex FFT that carries TWO real frames (frame A in the real part, frame B in the imaginary
// part); 400 = 20 x 20, each 20-point DFT is a twiddle-free 4x5 prime-factor transform held in
// registers, so a frame pair makes exactly two trips through shared memory.  The power spectra
// of the 32 frames of a batch are then projected on the (97.5 % sparse) mel filterbank, log10'd,
// written, and the per-clip maximum is folded into an ordered-int atomicMax.
// Kernel 2 (k_logmel_norm): max(x, clipmax-8), (x+4)/4 -- on values kernel 1 stored already scaled, so only the
// clamped ones are written back -- and the constant fill of the frames that only see zero padding.
#include "logmel_dev.cuh"
#include "records_dev.cuh"

namespace rho {

constexpr int LM_TWS = 22;                        // float2 per twiddle row (conflict-free 128-bit reads)

struct alignas(16) LmSmem {
  float2 fb[LM_GROUPS * LM_FB];                   // FFT buffers; afterwards the power spectra of the batch
  float slab[LM_SLAB_SM];                         // 16 kHz samples of the batch (prefetched under stage 2 / mel)
  float pw[LM_BF * LM_PS];                        // spectrum exchange between partner lanes
  float hannT[N_FFT];                             // [n2][n1] = hann[20*n1 + n2]: a thread's 20 window values are contiguous
  float2 twT[20 * LM_TWS];                        // [n2][k1] = W400^(n2*k1)
  float red[LM_THREADS / 32];
};
static_assert(LM_GROUPS * 200 * 2 <= LM_BF * LM_PS, "the spectrum exchange must fit in the pw region");
static_assert(LM_BF * LM_PS * 4 <= LM_GROUPS * LM_FB * 8, "the power spectra must fit in the fb region");

__global__ void k_logmel_init(int* __restrict__ clip_max, int* __restrict__ tiles_done, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { clip_max[i] = INT_MIN; if (tiles_done) tiles_done[i] = 0; }
}

// Same FFT / split / mel structure as k_fused_features (fused.cu), fed from a 16 kHz signal in HBM.
template <int NM>
__global__ void __launch_bounds__(LM_THREADS, 2)
k_logmel_frames(const float* __restrict__ x16, const int64_t* __restrict__ off, const int32_t* __restrict__ len16,
                const float* __restrict__ g_hann, const float2* __restrict__ g_tw, int n_mels, int pad_frames,
                float* __restrict__ mel, long long mel_stride, int* __restrict__ clip_max,
                int32_t* __restrict__ n_frames_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  LmSmem& S = *reinterpret_cast<LmSmem*>(smem_raw);
  const int c = blockIdx.x;
  int T, T_real, N, n_valid;
  lm_frame_counts(len16[c], pad_frames, &T, &T_real, &N, &n_valid);
  if (blockIdx.y == 0 && threadIdx.x == 0 && n_frames_out) n_frames_out[c] = T;
  const int tile_t0 = blockIdx.y * LM_TILE;
  if (tile_t0 >= T_real) return;

  const int tid = threadIdx.x;
  for (int i = tid; i < N_FFT; i += LM_THREADS) {
    const int r = i / 20, c20 = i - 20 * r;
    S.hannT[c20 * 20 + r] = g_hann[i];            // transpose: [n2][n1]
    S.twT[r * LM_TWS + c20] = g_tw[i];            // the table is symmetric in (k1, n2): re-stride only
  }

  const float* __restrict__ xs = x16 + off[c];
  float* __restrict__ out = mel + (long long)c * n_mels * mel_stride;
  const int g = tid / LM_LANES, lane = tid - g * LM_LANES;
  float2* fb = S.fb + g * LM_FB;
  float lmax = -INFINITY;

  // Stages the samples of frames [t0, t0+32): clip indices [160*t0-200, 160*t0-200+5360).  Interior
  // slabs go global -> shared with LDGSTS (no registers, all pieces in flight); slabs that touch the
  // reflected edges or the zero padding are gathered sample by sample.
  auto stage_slab = [&](int t0) {
    const long long i0 = (long long)HOP16 * t0 - N_FFT / 2;
    for (int q = tid; q < LM_SLAB / 4; q += LM_THREADS) {
      float* dst = S.slab + 4 * q + 20 * (q / (LM_SLAB_BLK / 4));
      const long long i = i0 + 4 * q;
      if (i >= 0 && i + 3 < n_valid) {
        cp_async16_zfill(dst, xs + i, 16);
      } else {                                    // reflected edge / zero padding: a few pieces per clip
        dst[0] = lm_sample(xs, i + 0, n_valid, N); dst[1] = lm_sample(xs, i + 1, n_valid, N);
        dst[2] = lm_sample(xs, i + 2, n_valid, N); dst[3] = lm_sample(xs, i + 3, n_valid, N);
      }
    }
    cp_async_commit();
  };
  stage_slab(tile_t0);

  for (int b = 0; b < LM_BATCHES; ++b) {
    const int t0 = tile_t0 + b * LM_BF;
    if (t0 >= T_real) break;
    cp_async_wait_all();
    __syncthreads();                              // slab (and, first time, the tables) visible; last batch's mel reads done
    // ---- FFT stage 1: lane = n2, 20-point DFT over n1 of z[20*n1 + n2], then twiddle W400^(n2*k1)
    float2 v[20];
    {
      // frame A = samples [320g, 320g+400), frame B = [320g+160, 320g+560) of the slab; sample s lives
      // at s + 20*(s/320), and 20*n1+lane crosses a block boundary at n1 = 16 (A) / n1 = 8 (B).
      const float* fa = S.slab + g * LM_SLAB_STRIDE + lane;
      const float* fbm = fa + HOP16;
      const float4* hq = reinterpret_cast<const float4*>(S.hannT + 20 * lane);
#pragma unroll
      for (int q = 0; q < 5; ++q) {
        const float4 h4 = hq[q];
        const float hh[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int n1 = 4 * q + e;
          v[n1] = __fmul2_rn(make_float2(fa[20 * n1 + (n1 >= 16 ? 20 : 0)], fbm[20 * n1 + (n1 >= 8 ? 20 : 0)]),
                             make_float2(hh[e], hh[e]));
        }
      }
    }
    dft20(v);
    {
      const float4* tq = reinterpret_cast<const float4*>(S.twT + LM_TWS * lane);
      fb[lane] = v[0];
#pragma unroll
      for (int q = 0; q < 10; ++q) {
        const float4 t4 = tq[q];                  // twiddles of k1 = 2q, 2q+1
        if (q > 0) fb[(2 * q) * 21 + lane] = cmul(v[2 * q], make_float2(t4.x, t4.y));
        fb[(2 * q + 1) * 21 + lane] = cmul(v[2 * q + 1], make_float2(t4.z, t4.w));
      }
    }
    __syncthreads();
    // the slab is dead now: prefetch the next batch's samples under stage 2 / power / mel
    if (b + 1 < LM_BATCHES && t0 + LM_BF < T_real) stage_slab(t0 + LM_BF);
    // ---- FFT stage 2: lane = k1, 20-point DFT over n2 -> Z[k1 + 20*k2] in v[k2]
#pragma unroll
    for (int n2 = 0; n2 < 20; ++n2) v[n2] = fb[lane * 21 + n2];
    dft20(v);
    // ---- split the two real spectra and take |.|^2:  A = (Z[k]+conj Z[400-k])/2, B = (Z[k]-conj Z[400-k])/(2i).
    // Bin k = k1 + 20*k2 <= 200 needs Z[400-k]: the upper half (k2 >= 10) of lane 20-k1.  Every lane publishes its
    // upper half in the pw region and reads its partner's; the power spectra go to the fb region.
    float2* pub = reinterpret_cast<float2*>(S.pw) + g * 200;
#pragma unroll
    for (int k2 = 10; k2 < 20; ++k2) pub[lane + 20 * (k2 - 10)] = v[k2];
    __syncthreads();
    float* power = reinterpret_cast<float*>(S.fb);
    {
      float* pa = power + (2 * g) * LM_PS + lane;
      float* pb = pa + LM_PS;
      const float2* part = pub + (20 - lane);     // lane 0: reads stay inside the pw region (unused)
#pragma unroll
      for (int k2 = 0; k2 < 10; ++k2) {
        float2 w = part[20 * (9 - k2)];
        if (lane == 0) w = (k2 == 0) ? v[0] : v[20 - k2];
        const float2 z = v[k2];
        const float ar = z.x + w.x, ai = z.y - w.y, br = z.x - w.x, bi = z.y + w.y;
        pa[20 * k2] = 0.25f * (ar * ar + ai * ai);
        pb[20 * k2] = 0.25f * (br * br + bi * bi);
      }
      if (lane == 0) {                            // k = 200: Z[200] pairs with itself
        const float2 z = v[10];
        pa[200] = z.x * z.x;
        pb[200] = z.y * z.y;
      }
    }
    __syncthreads();
    // ---- mel projection + log10: lane-of-warp = frame, warp = one part of the mel rows.  The filterbank
    // is straight-line FFMA code with immediate weights (mel_sparse_gen.inc); log10 = log2 * log10(2)
    // on the SFU (|err| < 3e-6 on a value that is divided by 4 afterwards).
    {
      const int f = tid & 31, part = tid >> 5;
      const int t = t0 + f;
      const bool live = t < T_real;
      const float* p = power + f * LM_PS;
      float* __restrict__ o = out + t;
      auto emit = [&](int m, float acc) {
        const float ls = __log2f(fmaxf(acc, 1e-10f)) * 0.30102999566398120f;
        if (live) {
          o[(long long)m * mel_stride] = lm_scaled(ls);
          lmax = fmaxf(lmax, ls);
        }
      };
      if (NM == 80) mel_sparse_80(part, p, emit); else mel_sparse_128(part, p, emit);
    }
    // the next batch's first barrier orders these reads before stage 1 overwrites the fb region
  }
  lmax = warp_max(lmax);
  if ((tid & 31) == 0) S.red[tid >> 5] = lmax;
  __syncthreads();
  if (tid < 32) {
    float m = tid < LM_THREADS / 32 ? S.red[tid] : -INFINITY;
    m = warp_max(m);
    if (tid == 0 && m > -INFINITY) atomicMax(&clip_max[c], float_to_ordered(m));
  }
}

// grid (n clips, groups of 8 rows); normalises the frames with signal and fills the zero-padding frames.
// One 128-bit column piece per thread and pass, all 8 rows of the group: the 8 loads of a piece with signal
// are issued before the first store (the kernel used to wait for every load: long-scoreboard bound).
#ifndef RHO_NORM_ROWS
#define RHO_NORM_ROWS 8
#endif
constexpr int NORM_ROWS = RHO_NORM_ROWS;   // mel rows per CTA
constexpr int NORM_BATCH = 4;     // rows in flight per thread: 48 registers -> 5 CTAs per SM (8 in flight needed 78 -> 3 CTAs)
__global__ void __launch_bounds__(256, 5)
k_logmel_norm(const int32_t* __restrict__ len16, int n_mels, int pad_frames, float* __restrict__ mel,
              long long mel_stride, const int* __restrict__ clip_max, int fill_done, FinalizeArgs fin, int fill_to,
              float* __restrict__ pad_value) {
  const int c = blockIdx.x;
  // fused path: the first warp of the clip's first CTA also assembles the clip's record (decay decision, cosine)
  if (fin.rec && blockIdx.y == 0 && threadIdx.x < 32)
    finalize_item(fin.seg, fin.item, fin.item_first_seg, c, threadIdx.x, fin.decay_thr, fin.rec, fin.emb, fin.ref, fin.dim,
                  fin.peers);
  int T, T_real, N, n_valid;
  lm_frame_counts(len16[c], pad_frames, &T, &T_real, &N, &n_valid);
  const bool publisher = fin.rec && blockIdx.y == 0 && threadIdx.x == 0;      // the thread that stored this clip's record
  if (T <= 0) { if (publisher) publish_records(fin.peers); return; }
  if (pad_frames > 0) T = min(T, max(fill_to, T_real));   // compact rows: frames >= fill_to are not materialised
  // fill_done: the fused kernel has already written the constant for the columns >= round_up(T_real, 4)
  if (fill_done && T > T_real) T = min(T, (T_real + 3) & ~3);
  const float mx = ordered_to_float(clip_max[c]);
  const float floor_v = __fsub_rn(mx, 8.0f);
  const float fill = __fmul_rn(__fadd_rn(fmaxf(-10.0f, floor_v), 4.0f), 0.25f);   // x / 4 == x * 0.25 bit for bit
  if (pad_value && blockIdx.y == 0 && threadIdx.x == 32) pad_value[c] = fill;
  // the stored values are lm_scaled(x): the reference's (max(x, floor) + 4) / 4 is max(stored, lm_scaled(floor)) bit for bit
  const float floor_s = lm_scaled(floor_v);
  auto nrm = [&](float v) { return fmaxf(v, floor_s); };
  const int m0 = blockIdx.y * NORM_ROWS;
  const int rows = min(NORM_ROWS, n_mels - m0);
  float* __restrict__ base = mel + ((long long)c * n_mels + m0) * mel_stride;
  const bool vec = (mel_stride % 4 == 0) && ((((uintptr_t)base) & 15u) == 0);
  if (vec) {
    const int T4 = T >> 2;
    for (int cp = threadIdx.x; cp < T4; cp += 256) {
      const int t = 4 * cp;
      float* col = base + t;
      if (t >= T_real) {
        const float4 f4 = make_float4(fill, fill, fill, fill);
#pragma unroll
        for (int m = 0; m < NORM_ROWS; ++m) if (m < rows) stg_stream4(col + (long long)m * mel_stride, f4);
      } else {
        // the piece that straddles T_real takes the same path: its tail is inside the row, just not computed
        const bool k1 = t + 1 < T_real, k2 = t + 2 < T_real, k3 = t + 3 < T_real;
#pragma unroll
        for (int mb = 0; mb < NORM_ROWS; mb += NORM_BATCH) {
          float4 r4[NORM_BATCH];
#pragma unroll
          for (int j = 0; j < NORM_BATCH; ++j)
            if (mb + j < rows) r4[j] = *reinterpret_cast<const float4*>(col + (long long)(mb + j) * mel_stride);
#pragma unroll
          for (int j = 0; j < NORM_BATCH; ++j)
            if (mb + j < rows) {
              // written only if the clamp changes something (7 % of the values on speech-like clips, contiguous in
              // time) or the piece straddles T_real
              const float lo4 = fminf(fminf(r4[j].x, k1 ? r4[j].y : r4[j].x), fminf(k2 ? r4[j].z : r4[j].x, k3 ? r4[j].w : r4[j].x));
              if (!k3 || lo4 < floor_s)
                stg_stream4(col + (long long)(mb + j) * mel_stride,
                            make_float4(nrm(r4[j].x), k1 ? nrm(r4[j].y) : fill, k2 ? nrm(r4[j].z) : fill,
                                        k3 ? nrm(r4[j].w) : fill));
            }
        }
      }
    }
    const int rem = T - 4 * T4;
    for (int i = threadIdx.x; i < rows * rem; i += 256) {
      const int m = i / rem, t = 4 * T4 + (i - m * rem);
      float* row = base + (long long)m * mel_stride;
      const float v = row[t];
      if (t >= T_real) row[t] = fill; else if (v < floor_s) row[t] = floor_s;
    }
  } else {
    for (int m = 0; m < rows; ++m) {
      float* row = base + (long long)m * mel_stride;
      for (int t = threadIdx.x; t < T; t += 256) {
        const float v = row[t];
        if (t >= T_real) row[t] = fill; else if (v < floor_s) row[t] = floor_s;
      }
    }
  }
  if (publisher) publish_records(fin.peers);
}

cudaError_t launch_logmel(const Tables& tb, const float* x16, const int64_t* off, const int32_t* len16,
                          int n, int64_t max_len16, int n_mels, int pad_frames, float* mel,
                          int64_t mel_stride_frames, int32_t* n_frames, int* clip_max,
                          cudaStream_t st, LaunchCtx* lc, int fill_to, float* pad_value) {
  if (n <= 0) return cudaSuccess;
  cudaError_t e0 = launch_logmel_init(clip_max, n, st, lc);
  if (e0 != cudaSuccess) return e0;
  int64_t max_real;
  if (pad_frames > 0) {
    int64_t nv = max_len16 < (int64_t)pad_frames * HOP16 ? max_len16 : (int64_t)pad_frames * HOP16;
    max_real = (nv + N_FFT / 2 + HOP16 - 1) / HOP16;
    if (max_real < 2) max_real = 2;
    if (max_real > pad_frames) max_real = pad_frames;
  } else {
    max_real = max_len16 / HOP16;
  }
  unsigned tiles = (unsigned)((max_real + LM_TILE - 1) / LM_TILE);
  if (tiles == 0) tiles = 1;
  const size_t smem = sizeof(LmSmem);
  auto kern = (n_mels == 80) ? k_logmel_frames<80> : k_logmel_frames<128>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  dim3 grid((unsigned)n, tiles);
  lc->begin(KID_LOGMEL_FRAMES, st);
  kern<<<grid, LM_THREADS, smem, st>>>(x16, off, len16, tb.hann, tb.twiddle, n_mels, pad_frames, mel,
                                       mel_stride_frames, clip_max, n_frames);
  lc->end(st);
  return launch_logmel_norm(len16, n, n_mels, pad_frames, mel, mel_stride_frames, clip_max, st, lc, false, nullptr,
                            fill_to, pad_value);
}

cudaError_t launch_logmel_init(int* clip_max, int n, cudaStream_t st, LaunchCtx* lc, int* tiles_done) {
  if (n <= 0) return cudaSuccess;
  lc->begin(KID_LOGMEL_INIT, st);
  k_logmel_init<<<(n + 255) / 256, 256, 0, st>>>(clip_max, tiles_done, n);
  lc->end(st);
  return cudaGetLastError();
}

cudaError_t launch_logmel_norm(const int32_t* len16, int n, int n_mels, int pad_frames, float* mel,
                               int64_t mel_stride_frames, const int* clip_max, cudaStream_t st, LaunchCtx* lc,
                               bool fill_done, const FinalizeArgs* fin, int fill_to, float* pad_value) {
  if (n <= 0) return cudaSuccess;
  if (fill_to <= 0) fill_to = pad_frames;
  dim3 g2((unsigned)n, (unsigned)((n_mels + NORM_ROWS - 1) / NORM_ROWS));
  FinalizeArgs fa{};
  if (fin) fa = *fin;
  lc->begin(KID_LOGMEL_NORM, st);
  k_logmel_norm<<<g2, 256, 0, st>>>(len16, n_mels, pad_frames, mel, mel_stride_frames, clip_max, fill_done ? 1 : 0, fa,
                                    fill_to, pad_value);
  lc->end(st);
  return cudaGetLastError();
}

}  // namespace rho
