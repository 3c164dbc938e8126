// Speaker-encoder front end (SURVEY.md 8f NEXT-3): what resemblyzer runs on the CPU between the 16 kHz waveform and its
// LSTM, for a ragged batch.  The reference reaches it through BaseTTS._compute_speaker_similarity (base_tts.py:326-347),
// QwenTTS._initialize_reference_embedding (providers/qwen.py:199-216) and the drift classifier (validation/classifier/
// trainer.py:41-47): preprocess_wav -> VoiceEncoder.embed_utterance.  resemblyzer is a third-party dependency that is
// not under /root/reference; oracle/speaker.py restates its published algorithm:
//   normalize_volume(wav, -30 dBFS, increase_only)          k_spk_sumsq + k_spk_scale (or the gain folded into k_spk_mel)
//   compute_partial_slices(len, rate 1.3, coverage 0.75)    spk_slices() below, host and device
//   wav_to_mel_spectrogram: librosa melspectrogram(sr 16000, n_fft 400, hop 160, n_mels 40): periodic hann, centred with
//     ZERO padding, |X|^2, slaney mel bands, no log; [T, 40] with T = 1 + len / 160                       k_spk_mel
//   np.array([mel[s] for s in mel_slices]) -> [P, 160, 40]                                                k_spk_mel
//   mean of the partial embeddings, L2 normalisation                                                     k_spk_pool
// Not here: librosa.resample (soxr; the 16 kHz signal is an input), webrtcvad (trim_long_silences), the LSTM.
// k_spk_mel is the Whisper front end's FFT (logmel.cu: two real frames per 400-point complex FFT, 400 = 20 x 20) with
// zero instead of reflect padding, 40 baked mel rows (mel_sparse_gen.inc), a gain per clip, and a [32 frames][40] tile in
// shared memory so that the frame-major rows -- and the up to three partial utterances a frame belongs to -- are written
// as contiguous runs.
#include "logmel_dev.cuh"

namespace rho {

constexpr int SPK_TWS = 22;
constexpr int SPK_TILE_STRIDE = SPK_MELS + 1;       // 41: conflict-free when lane = frame

struct alignas(16) SpkSmem {
  float2 fb[LM_GROUPS * LM_FB];
  float slab[LM_SLAB_SM];
  float pw[LM_BF * LM_PS];
  float hannT[N_FFT];
  float2 twT[20 * SPK_TWS];
  float tile[LM_BF * SPK_TILE_STRIDE];
};

// ---- volume normalisation -------------------------------------------------------------------------------------------
__global__ void k_spk_zero(double* __restrict__ sums, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) sums[i] = 0.0;
}

// grid (clips, chunks): sum of squares of the clip in double
__global__ void __launch_bounds__(256)
k_spk_sumsq(const float* __restrict__ x, const int64_t* __restrict__ off, const char* __restrict__ len_base,
            int len_stride, double* __restrict__ sums) {
  const int c = blockIdx.x;
  const long long n = *reinterpret_cast<const int32_t*>(len_base + (size_t)c * len_stride);
  const long long per = (long long)SPK_SUM_CHUNK;
  const long long lo = (long long)blockIdx.y * per, hi = min(n, lo + per);
  if (lo >= hi) return;
  const float* __restrict__ xs = x + off[c];
  double acc = 0.0;
  float part = 0.f;                                 // fp32 over 16 values, then folded into the double
  int cnt = 0;
  for (long long i = lo + threadIdx.x; i < hi; i += 256) {
    const float v = xs[i];
    part = fmaf(v, v, part);
    if (++cnt == 16) { acc += (double)part; part = 0.f; cnt = 0; }
  }
  acc += (double)part;
  acc = warp_sum(acc);
  __shared__ double red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(&sums[c], t);
  }
}

// resemblyzer/audio.py normalize_volume, in the float32 arithmetic numpy uses for a float32 waveform:
//   rms = sqrt(mean((wav * 32767)^2)); dBFS = 20 log10(rms / 32767); change = target - dBFS; wav * 10^(change / 20)
// mode: 0 = always, 1 = increase only, 2 = decrease only.  An all-zero clip has rms 0: the gain is +inf like the
// reference's (0 * inf = NaN samples); an EMPTY clip gets gain 1.
__device__ __forceinline__ float spk_gain(double sumsq, long long n, float target_dbfs, int mode) {
  if (n <= 0) return 1.f;
  const float ms = (float)(sumsq / (double)n * (32767.0 * 32767.0));
  const float rms = sqrtf(ms);
  if (rms == 0.f) return mode == 2 ? 1.f : INFINITY;
  const float dbfs = 20.f * log10f(rms / 32767.f);
  const float change = target_dbfs - dbfs;
  if ((change < 0.f && mode == 1) || (change > 0.f && mode == 2)) return 1.f;
  return powf(10.f, change / 20.f);
}

__global__ void k_spk_gain(const double* __restrict__ sums, const char* __restrict__ len_base, int len_stride, int n,
                           float target_dbfs, int mode, float* __restrict__ gain) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const long long len = *reinterpret_cast<const int32_t*>(len_base + (size_t)c * len_stride);
  gain[c] = spk_gain(sums[c], len, target_dbfs, mode);
}

__global__ void __launch_bounds__(256)
k_spk_scale(const float* __restrict__ x, const int64_t* __restrict__ off, const char* __restrict__ len_base,
            int len_stride, const float* __restrict__ gain, float* __restrict__ y, const int64_t* __restrict__ y_off) {
  const int c = blockIdx.x;
  const long long n = *reinterpret_cast<const int32_t*>(len_base + (size_t)c * len_stride);
  const long long lo = (long long)blockIdx.y * SPK_SUM_CHUNK, hi = min(n, lo + (long long)SPK_SUM_CHUNK);
  if (lo >= hi) return;
  const float g = gain[c];
  const float* __restrict__ xs = x + off[c];
  float* __restrict__ ys = y + y_off[c];
  if (g == 1.f && xs == ys) return;                 // the reference returns the waveform itself
  const bool vec = (((uintptr_t)xs | (uintptr_t)ys) & 15u) == 0;      // lo is a multiple of 4
  if (vec) {
    const long long n4 = (hi - lo) >> 2;
    for (long long q = threadIdx.x; q < n4; q += 256) {
      float4 v = *reinterpret_cast<const float4*>(xs + lo + 4 * q);
      v.x = __fmul_rn(v.x, g); v.y = __fmul_rn(v.y, g); v.z = __fmul_rn(v.z, g); v.w = __fmul_rn(v.w, g);
      *reinterpret_cast<float4*>(ys + lo + 4 * q) = v;
    }
    for (long long i = lo + 4 * n4 + threadIdx.x; i < hi; i += 256) ys[i] = __fmul_rn(xs[i], g);
  } else {
    for (long long i = lo + threadIdx.x; i < hi; i += 256) ys[i] = __fmul_rn(xs[i], g);
  }
}

// ---- mel spectrogram + partial utterances -----------------------------------------------------------------------------
// grid (clips, tiles of 256 frames), 320 threads.  Frame t of clip c covers samples [160 t - 200, 160 t + 200) of the
// gain-scaled clip, zeros outside [0, len).
__global__ void __launch_bounds__(LM_THREADS, 2)
k_spk_mel(const float* __restrict__ x16, const int64_t* __restrict__ off, const char* __restrict__ len_base, int len_stride,
          const float* __restrict__ g_hann, const float2* __restrict__ g_tw, const float* __restrict__ gain,
          int frame_step, double min_coverage, int pad_to_slices, float* __restrict__ mel,
          const int64_t* __restrict__ frame_off, float* __restrict__ partials, const int32_t* __restrict__ part_off) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SpkSmem& S = *reinterpret_cast<SpkSmem*>(smem_raw);
  const int c = blockIdx.x;
  const long long n = max(0, *reinterpret_cast<const int32_t*>(len_base + (size_t)c * len_stride));
  long long n_pad = n;
  int n_part = 0;
  if (pad_to_slices || partials) {
    long long padded;
    n_part = spk_slices(n, frame_step, min_coverage, &padded);
    if (padded > n_pad) n_pad = padded;
  }
  const int T = (int)(1 + n_pad / HOP16);
  const int tile_t0 = blockIdx.y * LM_TILE;
  if (tile_t0 >= T) return;

  const int tid = threadIdx.x;
  for (int i = tid; i < N_FFT; i += LM_THREADS) {
    const int r = i / 20, c20 = i - 20 * r;
    S.hannT[c20 * 20 + r] = g_hann[i];
    S.twT[r * SPK_TWS + c20] = g_tw[i];
  }
  // a non-finite gain (all-zero clip: 0 * inf = NaN in the reference too) must not touch the zero padding: such clips take
  // the element-wise staging path, which scales the samples inside the clip only
  const float g_clip = gain ? gain[c] : 1.f;
  const bool odd_gain = !isfinite(g_clip);
  const float gn = odd_gain ? 1.f : g_clip;
  const float* __restrict__ xs = x16 + off[c];
  const bool al16 = (((uintptr_t)xs) & 15u) == 0;
  float* __restrict__ mrow = mel ? mel + frame_off[c] * SPK_MELS : nullptr;
  float* __restrict__ prow = partials ? partials + (long long)part_off[c] * SPK_PART_FRAMES * SPK_MELS : nullptr;
  const int g = tid / LM_LANES, lane = tid - g * LM_LANES;
  float2* fb = S.fb + g * LM_FB;

  auto stage_slab = [&](int t0) {
    const long long i0 = (long long)HOP16 * t0 - N_FFT / 2;
    for (int q = tid; q < LM_SLAB / 4; q += LM_THREADS) {
      float* dst = S.slab + 4 * q + 20 * (q / (LM_SLAB_BLK / 4));
      const long long i = i0 + 4 * q;
      if (al16 && !odd_gain && i >= 0 && i + 3 < n) {
        cp_async16_zfill(dst, xs + i, 16);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) dst[k] = (i + k >= 0 && i + k < n) ? (odd_gain ? xs[i + k] * g_clip : xs[i + k]) : 0.f;
      }
    }
    cp_async_commit();
  };
  stage_slab(tile_t0);

  for (int b = 0; b < LM_BATCHES; ++b) {
    const int t0 = tile_t0 + b * LM_BF;
    if (t0 >= T) break;
    cp_async_wait_all();
    __syncthreads();
    // ---- FFT stage 1 (see logmel.cu): the sample is scaled by the clip's gain first, like the reference's waveform
    float2 v[20];
    {
      const float* fa = S.slab + g * LM_SLAB_STRIDE + lane;
      const float* fbm = fa + HOP16;
      const float* hq = S.hannT + 20 * lane;
#pragma unroll
      for (int n1 = 0; n1 < 20; ++n1) {
        const float a = __fmul_rn(fa[20 * n1 + (n1 >= 16 ? 20 : 0)], gn), bb = __fmul_rn(fbm[20 * n1 + (n1 >= 8 ? 20 : 0)], gn);
        v[n1] = make_float2(a * hq[n1], bb * hq[n1]);
      }
    }
    dft20(v);
    {
      const float4* tq = reinterpret_cast<const float4*>(S.twT + SPK_TWS * lane);
      fb[lane] = v[0];
#pragma unroll
      for (int q = 0; q < 10; ++q) {
        const float4 t4 = tq[q];
        if (q > 0) fb[(2 * q) * 21 + lane] = cmul(v[2 * q], make_float2(t4.x, t4.y));
        fb[(2 * q + 1) * 21 + lane] = cmul(v[2 * q + 1], make_float2(t4.z, t4.w));
      }
    }
    __syncthreads();
    if (b + 1 < LM_BATCHES && t0 + LM_BF < T) stage_slab(t0 + LM_BF);
    // ---- FFT stage 2, split of the two real spectra, |.|^2
#pragma unroll
    for (int n2 = 0; n2 < 20; ++n2) v[n2] = fb[lane * 21 + n2];
    dft20(v);
    float2* pub = reinterpret_cast<float2*>(S.pw) + g * 200;
#pragma unroll
    for (int k2 = 10; k2 < 20; ++k2) pub[lane + 20 * (k2 - 10)] = v[k2];
    __syncthreads();
    float* power = reinterpret_cast<float*>(S.fb);
    {
      float* pa = power + (2 * g) * LM_PS + lane;
      float* pb = pa + LM_PS;
      const float2* part = pub + (20 - lane);
#pragma unroll
      for (int k2 = 0; k2 < 10; ++k2) {
        float2 w = part[20 * (9 - k2)];
        if (lane == 0) w = (k2 == 0) ? v[0] : v[20 - k2];
        const float2 z = v[k2];
        const float ar = z.x + w.x, ai = z.y - w.y, br = z.x - w.x, bi = z.y + w.y;
        pa[20 * k2] = 0.25f * (ar * ar + ai * ai);
        pb[20 * k2] = 0.25f * (br * br + bi * bi);
      }
      if (lane == 0) {
        const float2 z = v[10];
        pa[200] = z.x * z.x;
        pb[200] = z.y * z.y;
      }
    }
    __syncthreads();
    // ---- 40 mel bands, no log: lane of warp = frame, warp = a part of the rows; into the [frame][band] tile
    {
      const int f = tid & 31, part = tid >> 5;
      const float* p = power + f * LM_PS;
      float* trow = S.tile + f * SPK_TILE_STRIDE;
      auto emit = [&](int m, float acc) { trow[m] = acc; };
      mel_sparse_40(part, p, emit);
    }
    __syncthreads();
    // ---- the tile is 32 consecutive rows of the clip's [T, 40] spectrogram: one contiguous run, and one per partial
    const int nf = min(LM_BF, T - t0);
    for (int i = tid; i < nf * SPK_MELS; i += LM_THREADS) {
      const int f = i / SPK_MELS, m = i - f * SPK_MELS, t = t0 + f;
      const float val = S.tile[f * SPK_TILE_STRIDE + m];
      if (mrow) mrow[(long long)t * SPK_MELS + m] = val;
      if (prow) {
        // partial j holds frames [step j, step j + 160)
        int j_hi = t / frame_step;
        if (j_hi > n_part - 1) j_hi = n_part - 1;
        int j_lo = t - (SPK_PART_FRAMES - 1);
        j_lo = j_lo > 0 ? (j_lo + frame_step - 1) / frame_step : 0;
        for (int j = j_lo; j <= j_hi; ++j)
          prow[((long long)j * SPK_PART_FRAMES + (t - frame_step * j)) * SPK_MELS + m] = val;
      }
    }
    // the next batch's first barrier orders the tile reads before the next mel phase writes it
  }
}

// One CTA per clip: mean of its partial embeddings (added in order, fp32, like numpy's axis-0 reduction), then divided by
// its L2 norm (resemblyzer/voice_encoder.py embed_utterance).
__global__ void __launch_bounds__(256)
k_spk_pool(const float* __restrict__ pe, const int32_t* __restrict__ part_off, int dim, float* __restrict__ out) {
  const int c = blockIdx.x;
  const int p0 = part_off[c], p1 = part_off[c + 1];
  const int cnt = p1 - p0;
  __shared__ double red[8];
  __shared__ float s_norm;
  double ss = 0.0;
  for (int d = threadIdx.x; d < dim; d += 256) {
    float acc = 0.f;
    for (int j = p0; j < p1; ++j) acc = __fadd_rn(acc, pe[(long long)j * dim + d]);
    const float mean = cnt > 0 ? __fdiv_rn(acc, (float)cnt) : nanf("");
    out[(long long)c * dim + d] = mean;
    ss += (double)mean * (double)mean;
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    s_norm = (float)sqrt(t);
  }
  __syncthreads();
  const float nrm = s_norm;
  for (int d = threadIdx.x; d < dim; d += 256) out[(long long)c * dim + d] = __fdiv_rn(out[(long long)c * dim + d], nrm);
}

// ---- launchers --------------------------------------------------------------------------------------------------------
cudaError_t launch_spk_normalize(const float* x, const int64_t* off, const int32_t* len, int len_stride_bytes, int n,
                                 int64_t max_len, float target_dbfs, int mode, float* y, const int64_t* y_off,
                                 float* gain, double* sums, cudaStream_t st, LaunchCtx* lc) {
  if (n <= 0) return cudaSuccess;
  const char* lb = reinterpret_cast<const char*>(len);
  const int ls = len_stride_bytes ? len_stride_bytes : (int)sizeof(int32_t);
  unsigned chunks = (unsigned)((max_len + SPK_SUM_CHUNK - 1) / SPK_SUM_CHUNK);
  if (chunks == 0) chunks = 1;
  if (chunks > 65535) return cudaErrorInvalidValue;
  lc->begin(KID_SPK_SUMSQ, st);
  k_spk_zero<<<(n + 255) / 256, 256, 0, st>>>(sums, n);
  k_spk_sumsq<<<dim3((unsigned)n, chunks), 256, 0, st>>>(x, off, lb, ls, sums);
  k_spk_gain<<<(n + 255) / 256, 256, 0, st>>>(sums, lb, ls, n, target_dbfs, mode, gain);
  lc->end(st);
  if (y) {
    lc->begin(KID_SPK_SCALE, st);
    k_spk_scale<<<dim3((unsigned)n, chunks), 256, 0, st>>>(x, off, lb, ls, gain, y, y_off);
    lc->end(st);
  }
  return cudaGetLastError();
}

cudaError_t launch_spk_mel(const Tables& tb, const float* x16, const int64_t* off, const int32_t* len,
                           int len_stride_bytes, int n, int64_t max_len, int frame_step, double min_coverage,
                           bool pad_to_slices, const float* gain, float* mel, const int64_t* frame_off, float* partials,
                           const int32_t* part_off, cudaStream_t st, LaunchCtx* lc) {
  if (n <= 0) return cudaSuccess;
  const char* lb = reinterpret_cast<const char*>(len);
  const int ls = len_stride_bytes ? len_stride_bytes : (int)sizeof(int32_t);
  int64_t longest = max_len > 0 ? max_len : 0;
  // a clip's last partial ends at most frame_step + 1 <= 161 hops past its own end
  if (pad_to_slices || partials) longest += (int64_t)(SPK_PART_FRAMES + 1) * HOP16;
  const int64_t T = 1 + longest / HOP16;
  const unsigned tiles = (unsigned)((T + LM_TILE - 1) / LM_TILE);
  if (tiles > 65535) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(k_spk_mel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SpkSmem));
  if (e != cudaSuccess) return e;
  lc->begin(KID_SPK_MEL, st);
  k_spk_mel<<<dim3((unsigned)n, tiles), LM_THREADS, sizeof(SpkSmem), st>>>(
      x16, off, lb, ls, tb.hann, tb.twiddle, gain, frame_step, min_coverage, pad_to_slices ? 1 : 0, mel, frame_off,
      partials, part_off);
  lc->end(st);
  return cudaGetLastError();
}

cudaError_t launch_spk_pool(const float* partial_embeds, const int32_t* part_off, int n, int dim, float* out,
                            cudaStream_t st, LaunchCtx* lc) {
  if (n <= 0 || dim <= 0) return cudaSuccess;
  lc->begin(KID_SPK_POOL, st);
  k_spk_pool<<<(unsigned)n, 256, 0, st>>>(partial_embeds, part_off, dim, out);
  lc->end(st);
  return cudaGetLastError();
}

}  // namespace rho
