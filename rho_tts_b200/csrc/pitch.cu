// Pitch shift of a ragged batch of clips: the pitch half of BaseTTS._apply_speed_pitch (base_tts.py:639-648), i.e.
// torchaudio.functional.pitch_shift (functional.py:1596-1641):
//   stft(512, hop 128, periodic hann, centre, reflect)        k_pv_stft      real FFT as a 256-point complex
//   phase_vocoder(rate = 2^(-n_steps/12))  (:723-803)         k_pv_phase     wrapped phase increments + magnitudes
//                                                             k_pv_cumsum    phase accumulator (double, rounded to fp32)
//   istft(length = round(L / rate))                           k_pv_istft     polar -> inverse FFT -> window -> overlap-add
//   resample(int(sr / rate) -> sr), crop / zero-pad to L      k_resample_windowed
//
// The fp32 reference is sensitive to its own rounding: the phase accumulator reaches ~1e6 rad on the top bins, where
// one fp32 ulp is 0.06 rad (fp32 pitch_shift differs from the same algorithm in fp64 by 1e-4 .. 5e-4).  So the phase
// path repeats torch's fp32 operations one by one, in torch's order, with explicitly rounded intrinsics (no FMA
// contraction): the time steps as the vectorised arange kernel makes them, phase_advance as linspace makes it (table
// from the host), the wrap as a true division + round-half-even, the accumulator in double rounded to fp32 per element
// (torch.cumsum on CPU), and the magnitude interpolation as two products and a sum.  oracle/pitch.py documents how
// each of these was pinned against torch.
#include <cmath>
#include "kernels.h"
#include "warp_fft.cuh"

namespace rho {

constexpr int PV_NFFT = 512;
constexpr int PV_HOP = 128;
constexpr int PV_NFREQ = 257;
constexpr int PV_LD = 264;               // row stride of the spectrogram planes (elements)
constexpr int PV_OUT_HOPS = 37;          // output hops (of 128 samples) per CTA of k_pv_istft

__device__ __forceinline__ int pv_frames(long long L) { return (int)(1 + L / PV_HOP); }
__device__ __forceinline__ int pv_out_frames(int T, double rate) { return (int)ceil((double)T / rate); }
__device__ __forceinline__ long long pv_stretch_len(long long L, double rate) { return (long long)rint((double)L / rate); }

// torch.arange(0, T, rate, dtype=float32)[j] as the CPU range kernel evaluates it (oracle/pitch.py arange_f32)
__device__ __forceinline__ float pv_time_step(int j, int J, double rate, int vec) {
  if (vec > 0) {
    const int full = (J / (2 * vec)) * (2 * vec);
    if (j < full) {
      const int j0 = (j / vec) * vec;
      const float base = (float)__dmul_rn((double)j0, rate);
      return (float)__dadd_rn((double)base, __dmul_rn((double)(j - j0), rate));
    }
  }
  return (float)__dmul_rn((double)j, rate);
}

// ------------------------------------------------------------------ STFT
// One WARP per frame (8 frames in flight per CTA, 4 frames per warp): the 512 windowed samples are packed into a
// 256-point complex sequence in the warp's own shared-memory buffers, each lane does two radix-4 butterflies per pass
// (only __syncwarp between passes), and the real spectrum is unpacked in (k, 256 - k) pairs:
//   E = (Z[k] + conj(Z[256-k])) / 2,  O = -i/2 (Z[k] - conj(Z[256-k])),  X[k] = E + W512^k O,  X[256-k] = conj(E - W512^k O)
// Stored as (|X|, angle(X)): the vocoder needs exactly these two of every frame (functional.py:783-787) and each
// frame is used by ~2 / rate output frames.
constexpr int PV_STFT_WARPS = 8;
constexpr int PV_STFT_FPW = 4;           // frames per warp

// |X| as sqrt(re^2 + im^2) (one fma + IEEE sqrt, <= 1 ulp like hypotf; audio spectra are far from overflow / underflow)
__device__ __forceinline__ float2 pv_abs_angle(float re, float im) { return make_float2(__fsqrt_rn(fmaf(re, re, im * im)), atan2f(im, re)); }

__global__ void __launch_bounds__(32 * PV_STFT_WARPS)
k_pv_stft(const float* __restrict__ x, const int64_t* __restrict__ off, const char* __restrict__ len_base, int len_stride,
          const float2* __restrict__ g_w256, const float2* __restrict__ g_w512, const float* __restrict__ g_hann,
          float2* __restrict__ spec, long long spec_stride) {
  __shared__ float2 Eall[PV_STFT_WARPS][PV_E_SIZE];
  __shared__ float2 tw[PV_TW_SIZE];
  const int c = blockIdx.y;
  const long long L = *reinterpret_cast<const int32_t*>(len_base + (size_t)c * len_stride);
  if (L <= PV_NFFT / 2) return;                      // reflect padding needs L > 256 (the host refuses such calls)
  const int T = pv_frames(L);
  const int f0 = blockIdx.x * (PV_STFT_WARPS * PV_STFT_FPW);
  if (f0 >= T) return;
  const float* __restrict__ xs = x + off[c];
  float2* __restrict__ sp = spec + (size_t)c * spec_stride;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float2* __restrict__ E = Eall[w];
  pv_fill_twiddles<false>(g_w256, tw);
  __syncthreads();
  const float2* __restrict__ tw1 = tw;
  const float2* __restrict__ tw2 = tw + 256;
  const int k1 = lane >> 2, q = lane & 3, bq = ((q & 1) << 1) | (q >> 1);
  for (int it = 0; it < PV_STFT_FPW; ++it) {
    const int f = f0 + it * PV_STFT_WARPS + w;
    if (f >= T) break;                               // warp-uniform
    const long long base = (long long)f * PV_HOP - PV_NFFT / 2;
    const bool interior = base >= 0 && base + PV_NFFT <= L;
    const bool vec2 = interior && ((reinterpret_cast<uintptr_t>(xs + base) & 7u) == 0);
    float2 v[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int n = lane + 32 * r;
      const float2 h = __ldg(reinterpret_cast<const float2*>(g_hann) + n);
      float2 xv;
      if (vec2) {
        xv = *reinterpret_cast<const float2*>(xs + base + 2 * n);
      } else {
        long long i0 = base + 2 * n, i1 = i0 + 1;
        if (!interior) {
          i0 = i0 < 0 ? -i0 : (i0 >= L ? 2 * (L - 1) - i0 : i0);
          i1 = i1 < 0 ? -i1 : (i1 >= L ? 2 * (L - 1) - i1 : i1);
        }
        xv = make_float2(xs[i0], xs[i1]);
      }
      v[r] = make_float2(xv.x * h.x, xv.y * h.y);
    }
    warp_fft256<false>(v, E, tw1, tw2, lane);
    __syncwarp();
#pragma unroll
    for (int a = 0; a < 8; ++a) E[k1 + 8 * a + 72 * bq] = v[a];       // pv_nat(k1 + 8 a + 64 b)
    __syncwarp();
    float2* __restrict__ row = sp + (size_t)f * PV_LD;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int k = lane + 32 * m;                   // 0..127, partner 256 - k
      const float2 zk = E[pv_nat(k)], zc = E[pv_nat((256 - k) & 255)];
      const float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y - zc.y));
      const float2 o = make_float2(0.5f * (zk.y + zc.y), -0.5f * (zk.x - zc.x));
      const float2 wo = cmul(__ldg(g_w512 + k), o);
      // DC and Nyquist of a real transform have an exactly zero (+0) imaginary part: angle is 0 or +pi
      row[k] = pv_abs_angle(e.x + wo.x, k ? e.y + wo.y : 0.f);
      row[256 - k] = pv_abs_angle(e.x - wo.x, k ? -(e.y - wo.y) : 0.f);
    }
    if (lane == 0) {                                 // k = 128 is its own partner
      const float2 z = E[pv_nat(128)];
      const float2 e = make_float2(z.x, 0.f), o = make_float2(z.y, 0.f);
      const float2 wo = cmul(__ldg(g_w512 + 128), o);
      row[128] = pv_abs_angle(e.x + wo.x, e.y + wo.y);
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------ phase vocoder: increments and magnitudes
__global__ void __launch_bounds__(256)
k_pv_phase(const float2* __restrict__ spec, long long spec_stride, const char* __restrict__ len_base, int len_stride,
           const float* __restrict__ g_padv, double rate, int vec, float* __restrict__ mag, float* __restrict__ ph,
           long long plane_stride) {
  const int c = blockIdx.y;
  const long long L = *reinterpret_cast<const int32_t*>(len_base + (size_t)c * len_stride);
  if (L <= PV_NFFT / 2) return;
  const int T = pv_frames(L);
  const int J = pv_out_frames(T, rate);
  const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (j >= J) return;
  const int lane = threadIdx.x & 31;
  const float ts = pv_time_step(j, J, rate, vec);
  const float alpha = ts - truncf(ts);                           // ts % 1.0
  const long long i0 = (long long)ts, i1 = (long long)__fadd_rn(ts, 1.0f);
  const float2* __restrict__ sp = spec + (size_t)c * spec_stride;
  float* __restrict__ mg = mag + (size_t)c * plane_stride + (size_t)j * PV_LD;
  float* __restrict__ pc = ph + (size_t)c * plane_stride;
  const float two_pi = 6.283185307179586f;
  const float one_m = __fsub_rn(1.0f, alpha);
  for (int k = lane; k < PV_NFREQ; k += 32) {
    const float2 s0 = i0 < T ? sp[(size_t)i0 * PV_LD + k] : make_float2(0.f, 0.f);     // two zero frames of padding:
    const float2 s1 = i1 < T ? sp[(size_t)i1 * PV_LD + k] : make_float2(0.f, 0.f);     // |0| = 0, angle(0) = 0
    const float n0 = s0.x, a0 = s0.y, n1 = s1.x, a1 = s1.y;
    const float adv = g_padv[k];
    float p = __fsub_rn(__fsub_rn(a1, a0), adv);
    p = __fsub_rn(p, __fmul_rn(two_pi, rintf(__fdiv_rn(p, two_pi))));
    p = __fadd_rn(p, adv);
    mg[k] = __fadd_rn(__fmul_rn(alpha, n1), __fmul_rn(one_m, n0));
    if (j == 0) pc[k] = a0;                                      // phase_0
    if (j + 1 < J) pc[(size_t)(j + 1) * PV_LD + k] = p;          // cat([phase_0, phase[..., :-1]])
  }
}

// phase_acc = cumsum(phase): double accumulator, every element rounded to fp32 (torch.cumsum on CPU)
__global__ void __launch_bounds__(64)
k_pv_cumsum(const char* __restrict__ len_base, int len_stride, double rate, float* __restrict__ ph,
            long long plane_stride) {
  const int c = blockIdx.y;
  const int k = blockIdx.x * 64 + threadIdx.x;
  const long long L = *reinterpret_cast<const int32_t*>(len_base + (size_t)c * len_stride);
  if (L <= PV_NFFT / 2 || k >= PV_NFREQ) return;
  const int J = pv_out_frames(pv_frames(L), rate);
  float* __restrict__ p = ph + (size_t)c * plane_stride + k;
  double acc = 0.0;
  int j = 0;
  for (; j + 8 <= J; j += 8) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = p[(size_t)(j + u) * PV_LD];
#pragma unroll
    for (int u = 0; u < 8; ++u) { acc += (double)v[u]; p[(size_t)(j + u) * PV_LD] = (float)acc; }
  }
  for (; j < J; ++j) { acc += (double)p[(size_t)j * PV_LD]; p[(size_t)j * PV_LD] = (float)acc; }
}

// ------------------------------------------------------------------ inverse STFT
// One CTA makes PV_OUT_HOPS = 37 hops (4736 samples) of the stretched waveform, one WARP per frame.  Full-signal
// sample nf = n + 256 receives frames nf/128 - 3 .. nf/128, so the tile needs 40 frames: 5 rounds of 8 frames that
// are 5 hops apart (the frames of a round never touch the same output sample, so the overlap-add needs no atomics and
// its order is fixed), one CTA barrier per round.  Per frame: polar(mag, phase_acc) in (k, 256 - k) pairs ->
//   Z[k] = e + i o,  Z[256-k] = conj(e) + i conj(o),  e = X[k] + conj(X[256-k]),  o = (X[k] - conj(X[256-k])) conj(W512^k)
// -> inverse 256-point complex FFT (x[2n] = Re z[n] / 512, x[2n+1] = Im z[n] / 512) -> window -> add.
constexpr int PV_IS_WARPS = 8;
constexpr int PV_IS_ROUNDS = 5;
constexpr int PV_ACC_F2 = PV_OUT_HOPS * 72;      // the tile as float2 pairs, 64 per hop padded to 72 (pv_nat): conflict-free adds
static_assert(PV_OUT_HOPS == PV_IS_WARPS * PV_IS_ROUNDS - 3, "tile = frames - 3 hops");

// torch.polar(mag, phase_acc)[k]: the fp32 phase is reduced in double (exact input, error ~1e-10 rad)
__device__ __forceinline__ float2 pv_polar(const float* __restrict__ mg, const float* __restrict__ pc, int k) {
  const float m = mg[k];
  const double p = (double)pc[k];
  const double q = rint(p * 0.15915494309189535);
  const double red = fma(-q, 2.4492935982947064e-16, fma(-q, 6.283185307179586, p));
  float sn, cs;
  __sincosf((float)red, &sn, &cs);                                        // |red| <= pi: abs error < 5e-7
  return make_float2(m * cs, (k == 0 || k == 256) ? 0.f : m * sn);        // c2r ignores Im of DC and Nyquist
}

__global__ void __launch_bounds__(32 * PV_IS_WARPS)
k_pv_istft(const float* __restrict__ mag, const float* __restrict__ ph, long long plane_stride,
           const char* __restrict__ len_base, int len_stride, const float2* __restrict__ g_w256,
           const float2* __restrict__ g_w512, const float* __restrict__ g_hann, double rate,
           float* __restrict__ wave, long long wave_stride) {
  __shared__ float2 Eall[PV_IS_WARPS][PV_E_SIZE];
  __shared__ float2 tw[PV_TW_SIZE];
  __shared__ float2 acc[PV_ACC_F2];
  const int c = blockIdx.y;
  const long long L = *reinterpret_cast<const int32_t*>(len_base + (size_t)c * len_stride);
  if (L <= PV_NFFT / 2) return;
  const int J = pv_out_frames(pv_frames(L), rate);
  const long long LS = pv_stretch_len(L, rate);
  const int h0 = blockIdx.x * PV_OUT_HOPS;
  if ((long long)h0 * PV_HOP >= LS) return;
  for (int i = threadIdx.x; i < PV_ACC_F2; i += 32 * PV_IS_WARPS) acc[i] = make_float2(0.f, 0.f);
  pv_fill_twiddles<true>(g_w256, tw);
  __syncthreads();
  const float2* __restrict__ tw1 = tw;
  const float2* __restrict__ tw2 = tw + 256;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float2* __restrict__ E = Eall[w];
  const int k1 = lane >> 2, q = lane & 3, bq = ((q & 1) << 1) | (q >> 1);
  const int src = (32 - lane) & 31;
  const long long F0 = (long long)(h0 + 2) * PV_HOP;            // first full-signal sample of the tile
  const float* __restrict__ mgc = mag + (size_t)c * plane_stride;
  const float* __restrict__ pcc = ph + (size_t)c * plane_stride;
  for (int r = 0; r < PV_IS_ROUNDS; ++r) {
    const int j = h0 - 1 + r + PV_IS_ROUNDS * w;
    if (j >= 0 && j < J) {                                       // warp-uniform
      const float* __restrict__ mg = mgc + (size_t)j * PV_LD;
      const float* __restrict__ pc = pcc + (size_t)j * PV_LD;
      // Z[k] and Z[256 - k] of the packed transform from the pair X[k], X[256 - k] (k = lane + 32 m)
      float2 v[8], zp[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int k = lane + 32 * m;
        const float2 xa = pv_polar(mg, pc, k), xb = pv_polar(mg, pc, 256 - k);
        const float2 e = make_float2(xa.x + xb.x, xa.y - xb.y);
        const float2 d = make_float2(xa.x - xb.x, xa.y + xb.y);
        const float2 wk = __ldg(g_w512 + k);
        const float2 o = cmul(d, make_float2(wk.x, -wk.y));
        v[m] = make_float2(e.x - o.y, e.y + o.x);
        zp[m] = make_float2(e.x + o.y, -e.y + o.x);              // conj(e) + i conj(o) = Z[256 - k]
      }
      // lane l needs Z[l + 32 r], r = 4..7 = Z[256 - (32 - l) - 32 (7 - r)]: pair 7 - r of lane 32 - l; lane 0 holds its
      // own (Z[32 r] = pair 8 - r) and Z[128], which is its own partner
#pragma unroll
      for (int rr = 4; rr < 8; ++rr)
        v[rr] = make_float2(__shfl_sync(0xffffffffu, zp[7 - rr].x, src), __shfl_sync(0xffffffffu, zp[7 - rr].y, src));
      if (lane == 0) {
        const float2 xa = pv_polar(mg, pc, 128);
        const float2 wk = __ldg(g_w512 + 128);
        const float2 o = cmul(make_float2(0.f, 2.f * xa.y), make_float2(wk.x, -wk.y));
        v[4] = make_float2(2.f * xa.x - o.y, o.x);
        v[5] = zp[3]; v[6] = zp[2]; v[7] = zp[1];
      }
      warp_fft256<true>(v, E, tw1, tw2, lane);
      __syncwarp();
      // v[a] = z[n], n = k1 + 8 a + 64 b: samples 2n, 2n + 1 of frame j
      const int hop = j - (h0 + 2);                              // -3 .. 36
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        const int n = k1 + 8 * a + 64 * bq;
        const int i2 = 64 * hop + n;
        if (i2 >= 0 && i2 < PV_OUT_HOPS * 64) {
          const float2 h = __ldg(reinterpret_cast<const float2*>(g_hann) + n);
          float2 t = acc[pv_nat(i2)];
          t.x += (v[a].x * (1.0f / 512.0f)) * h.x;
          t.y += (v[a].y * (1.0f / 512.0f)) * h.y;
          acc[pv_nat(i2)] = t;
        }
      }
    }
    __syncthreads();
  }
  const long long full = PV_NFFT + (long long)PV_HOP * (J - 1);
  float* __restrict__ wv = wave + (size_t)c * wave_stride;
  const float* __restrict__ accf = reinterpret_cast<const float*>(acc);
  for (int i = threadIdx.x; i < PV_OUT_HOPS * PV_HOP; i += 32 * PV_IS_WARPS) {
    const long long nf = F0 + i, n = nf - PV_NFFT / 2;
    if (n >= LS) break;
    float out = 0.f;
    if (nf < full) {
      long long jlo = (nf - (PV_NFFT - 1) + PV_HOP - 1) / PV_HOP;
      if (nf - (PV_NFFT - 1) < 0) jlo = 0;
      const long long jhi = min((long long)J - 1, nf / PV_HOP);
      float env = 0.f;
      for (long long jj = jlo; jj <= jhi; ++jj) { const float hw = __ldg(g_hann + (nf - jj * PV_HOP)); env = __fadd_rn(env, __fmul_rn(hw, hw)); }
      out = __fdiv_rn(accf[2 * pv_nat(i >> 1) + (i & 1)], env);
    }
    wv[n] = out;
  }
}

// ------------------------------------------------------------------ windowed-sinc resample, any ratio, then crop / pad
// out[o] = sum_w taps[p][w] * in[q * orig + ilo[p] + w - width],  p = o mod new, q = o / new: only the taps inside the
// Hann window (2 * width + 2 per phase) are kept -- outside it torchaudio's fp32 taps are ~1e-23.  Output o >= the
// resampled length is zero (functional.py:1708-1713 pads), o >= L is not produced (crop).
constexpr int RW_QC = 8;                  // outputs per thread: the same phase in 8 consecutive blocks of `nw` outputs

// Outputs o and o + nw share their phase, i.e. their taps: a thread keeps the W taps of ONE phase in registers and
// makes that phase's output in RW_QC consecutive blocks (o = q * nw + p), so the tap table is read once per 8 outputs
// instead of once per output (64 B per output from L2 otherwise: that was the bound).  Consecutive threads hold
// consecutive phases: input reads (stride orig / nw samples across lanes) and output writes are coalesced.
template <int WT>
__global__ void __launch_bounds__(256)
k_resample_windowed(const float* __restrict__ wave, long long wave_stride, const char* __restrict__ len_base,
                    int len_stride, double rate, int orig, int nw, int width, int W, const float* __restrict__ taps,
                    const int* __restrict__ ilo, float* __restrict__ y, const int64_t* __restrict__ y_off, int qc_len) {
  const int c = blockIdx.y;
  const long long L = *reinterpret_cast<const int32_t*>(len_base + (size_t)c * len_stride);
  if (L <= PV_NFFT / 2) return;
  // qc_len = RW_QC: phase-major (thread = one phase, 8 blocks); qc_len = 1: thread = one output (ratios with few
  // phases, e.g. the octaves: the table is a few rows and stays in L1, and phase-major writes would not coalesce)
  const long long v = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long qc = v / nw;
  const int p = (int)(v - qc * nw);
  const long long q_total = (L + nw - 1) / nw;
  const long long q0 = qc * qc_len;
  if (q0 >= q_total) return;
  const long long LS = pv_stretch_len(L, rate);
  const long long target = ((long long)nw * LS + orig - 1) / orig;
  const float* __restrict__ in = wave + (size_t)c * wave_stride;
  float* __restrict__ out = y + y_off[c];
  const float* __restrict__ k = taps + (size_t)p * W;
  const long long s_base = (long long)ilo[p] - width;
  float kr[WT > 0 ? WT : 1];
  if (WT > 0) {
#pragma unroll
    for (int w = 0; w < WT; w += 4) {
      const float4 kk = __ldg(reinterpret_cast<const float4*>(k + w));
      kr[w] = kk.x; kr[w + 1] = kk.y; kr[w + 2] = kk.z; kr[w + 3] = kk.w;
    }
  }
  const long long q1 = min(q0 + qc_len, q_total);
  for (long long q = q0; q < q1; ++q) {
    const long long o = q * nw + p;
    if (o >= L) break;
    float a = 0.f;
    if (o < target) {
      const long long s0 = q * orig + s_base;
      const float* __restrict__ src = in + s0;
      if (WT > 0 && s0 >= 0 && s0 + WT <= LS) {
        float a1 = 0.f;
#pragma unroll
        for (int w = 0; w < WT; w += 2) { a = fmaf(src[w], kr[w], a); a1 = fmaf(src[w + 1], kr[w + 1], a1); }
        a += a1;
      } else if (WT > 0) {
#pragma unroll
        for (int w = 0; w < WT; ++w) { const long long si = s0 + w; if (si >= 0 && si < LS) a = fmaf(in[si], kr[w], a); }
      } else {
        for (int w = 0; w < W; ++w) { const long long si = s0 + w; if (si >= 0 && si < LS) a = fmaf(in[si], __ldg(k + w), a); }
      }
    }
    out[o] = a;
  }
}

// ------------------------------------------------------------------ host side
PitchPlan pitch_plan(int n, int64_t max_len, double rate) {
  PitchPlan p;
  p.T_max = 1 + max_len / PV_HOP;
  p.J_max = (int64_t)std::ceil((double)p.T_max / rate);
  p.LS_max = ((int64_t)std::nearbyint((double)max_len / rate) + 3) / 4 * 4;
  const size_t nn = (size_t)(n > 0 ? n : 1);
  size_t o = 0;
  p.spec = o; o += align_up(nn * (size_t)p.T_max * PV_LD * sizeof(float2), 256);
  p.mag = o;  o += align_up(nn * (size_t)p.J_max * PV_LD * sizeof(float), 256);
  p.ph = o;   o += align_up(nn * (size_t)p.J_max * PV_LD * sizeof(float), 256);
  p.wave = o; o += align_up(nn * (size_t)p.LS_max * sizeof(float), 256);
  p.total = o;
  return p;
}

cudaError_t launch_pitch_shift(const PitchTables& tb, const float* x, const int64_t* off, const int32_t* len,
                               int len_stride_bytes, int n, int64_t max_len, double rate, int arange_vec, int orig,
                               int nw, int width, int W, const float* taps, const int* ilo, float* y,
                               const int64_t* y_off, void* workspace, cudaStream_t st, LaunchCtx* lc) {
  if (n <= 0 || max_len <= 0) return cudaSuccess;
  const PitchPlan p = pitch_plan(n, max_len, rate);
  char* ws = (char*)workspace;
  float2* spec = (float2*)(ws + p.spec);
  float* mag = (float*)(ws + p.mag);
  float* ph = (float*)(ws + p.ph);
  float* wave = (float*)(ws + p.wave);
  const char* lb = reinterpret_cast<const char*>(len);
  const int ls = len_stride_bytes ? len_stride_bytes : (int)sizeof(int32_t);
  const long long spec_stride = (long long)p.T_max * PV_LD, plane_stride = (long long)p.J_max * PV_LD;
  const int stft_fr = PV_STFT_WARPS * PV_STFT_FPW;
  const unsigned gx_stft = (unsigned)((p.T_max + stft_fr - 1) / stft_fr);
  const unsigned gx_ph = (unsigned)((p.J_max + 7) / 8);
  const unsigned gx_is = (unsigned)((p.LS_max + PV_OUT_HOPS * PV_HOP - 1) / (PV_OUT_HOPS * PV_HOP));
  const int qc_len = nw >= 256 ? RW_QC : 1;
  const int64_t q_chunks = ((max_len + nw - 1) / nw + qc_len - 1) / qc_len;
  const unsigned gx_rs = (unsigned)((q_chunks * nw + 255) / 256);
  if (gx_stft > 0x7fffffffu || n > 65535) return cudaErrorInvalidValue;
  lc->begin(KID_PV_STFT, st);
  k_pv_stft<<<dim3(gx_stft, (unsigned)n), 32 * PV_STFT_WARPS, 0, st>>>(x, off, lb, ls, tb.w256, tb.w512, tb.hann, spec, spec_stride);
  lc->end(st);
  lc->begin(KID_PV_PHASE, st);
  k_pv_phase<<<dim3(gx_ph, (unsigned)n), 256, 0, st>>>(spec, spec_stride, lb, ls, tb.padv, rate, arange_vec, mag, ph,
                                                       plane_stride);
  lc->end(st);
  lc->begin(KID_PV_CUMSUM, st);
  k_pv_cumsum<<<dim3((PV_NFREQ + 63) / 64, (unsigned)n), 64, 0, st>>>(lb, ls, rate, ph, plane_stride);
  lc->end(st);
  lc->begin(KID_PV_ISTFT, st);
  k_pv_istft<<<dim3(gx_is > 0 ? gx_is : 1, (unsigned)n), 32 * PV_IS_WARPS, 0, st>>>(mag, ph, plane_stride, lb, ls, tb.w256, tb.w512,
                                                                      tb.hann, rate, wave, (long long)p.LS_max);
  lc->end(st);
  lc->begin(KID_PV_RESAMPLE, st);
  if (W == 16)
    k_resample_windowed<16><<<dim3(gx_rs, (unsigned)n), 256, 0, st>>>(wave, (long long)p.LS_max, lb, ls, rate, orig, nw,
                                                                      width, W, taps, ilo, y, y_off, qc_len);
  else
    k_resample_windowed<0><<<dim3(gx_rs, (unsigned)n), 256, 0, st>>>(wave, (long long)p.LS_max, lb, ls, rate, orig, nw,
                                                                     width, W, taps, ilo, y, y_off, qc_len);
  lc->end(st);
  return cudaGetLastError();
}

}  // namespace rho
