// QwenTTS._post_process_audio for a ragged batch of clips (SURVEY.md 8f NEXT-1).
//
// Reference: src/rho_tts/providers/qwen.py
//   :268-313  overall RMS gate (1e-8) -> windowed decay correction when n > 2 windows -> global gain to
//             -23 dBFS -> tanh(x / 0.95) * 0.95
//   :315-378  per-2-s-window RMS, gain = first / rms capped at +18 dB, skipped when max - min < 0.05,
//             two 3-tap smoothing passes with fixed end points, np.interp (float64) between window centres
//
// The reference makes five full passes over the clip (overall RMS, window RMS, envelope multiply, RMS again,
// gain + tanh) plus a float64 envelope as long as the clip.  Here: ONE read pass and ONE read + write pass.
//   k_qwen_moments : per half window (sr samples: the pieces between window edges and window centres)
//                    S0 = sum x^2, S1 = sum u x^2, S2 = sum u^2 x^2 (u = sample index inside the piece), double.
//                    Window RMS needs S0 only; the RMS *after* the envelope is applied needs no second pass,
//                    because the envelope is linear inside a piece: sum (x (A + B u))^2 = A^2 S0 + 2AB S1 + B^2 S2.
//   k_qwen_plan    : one thread per clip: the gate, the gains, the smoothing, the envelope knots, the global
//                    gain -- in double / fp32 exactly where the reference uses python floats / fp32 tensors.
//   k_qwen_apply   : y = tanh(fl(fl(x * env) * g) / 0.95) * 0.95, env = float(np.interp) evaluated in double.
// 12 bytes of HBM traffic per sample instead of the reference's 40+.
#include "common.cuh"
#include "kernels.h"

namespace rho {

constexpr int QW_THREADS = 256;
constexpr int QW_MAX_WINDOWS = 64;          // 128 s at any sample rate; longer clips are refused by the launcher

// per-clip plan written by k_qwen_plan (fixed size)
struct QwenPlan {
  int mode;                                  // 0: copy unchanged, 1: global gain + tanh, 2: envelope + gain + tanh
  int n_windows;
  float gain;                                // global gain as the fp32 scalar the reference multiplies with
  float pad;
  double knots[QW_MAX_WINDOWS];              // smoothed per-window gains (np.interp's fp)
  double slopes[QW_MAX_WINDOWS];             // (fp[j+1] - fp[j]) / (xp[j+1] - xp[j]), as numpy precomputes them
};

__global__ void __launch_bounds__(QW_THREADS)
k_qwen_moments(const float* __restrict__ x, const int64_t* __restrict__ off, const char* __restrict__ len_base,
               int len_stride, int half, int pieces_per_clip, double* __restrict__ mom) {
  const int c = blockIdx.x, h = blockIdx.y;
  const int n = *reinterpret_cast<const int32_t*>(len_base + (size_t)c * len_stride);
  const long long t0 = (long long)h * half;
  if (t0 >= n) return;
  const int cnt = (int)min((long long)half, n - t0);
  const float* __restrict__ p = x + off[c] + t0;
  // fp32 per thread (<= ~100 terms each), double across threads
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
  const bool al = (reinterpret_cast<uintptr_t>(p) & 15u) == 0;
  const int n4 = al ? (cnt >> 2) : 0;
  for (int q = threadIdx.x; q < n4; q += QW_THREADS) {
    const float4 v = ldg_stream4(p + 4 * q);
    const float u = (float)(4 * q);
    const float a = v.x * v.x, b = v.y * v.y, cc = v.z * v.z, d = v.w * v.w;
    s0 += (a + b) + (cc + d);
    s1 += (u * a + (u + 1.f) * b) + ((u + 2.f) * cc + (u + 3.f) * d);
    s2 += (u * u * a + (u + 1.f) * (u + 1.f) * b) + ((u + 2.f) * (u + 2.f) * cc + (u + 3.f) * (u + 3.f) * d);
  }
  for (int i = 4 * n4 + threadIdx.x; i < cnt; i += QW_THREADS) {
    const float v = p[i], a = v * v, u = (float)i;
    s0 += a; s1 += u * a; s2 += u * u * a;
  }
  __shared__ double red[3][QW_THREADS / 32];
  double d0 = warp_sum((double)s0), d1 = warp_sum((double)s1), d2 = warp_sum((double)s2);
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][w] = d0; red[1][w] = d1; red[2][w] = d2; }
  __syncthreads();
  if (w == 0) {
    d0 = lane < QW_THREADS / 32 ? red[0][lane] : 0.0;
    d1 = lane < QW_THREADS / 32 ? red[1][lane] : 0.0;
    d2 = lane < QW_THREADS / 32 ? red[2][lane] : 0.0;
    d0 = warp_sum(d0); d1 = warp_sum(d1); d2 = warp_sum(d2);
    if (lane == 0) {
      double* m = mom + ((size_t)c * pieces_per_clip + h) * 3;
      m[0] = d0; m[1] = d1; m[2] = d2;
    }
  }
}

// fp32 RMS of `cnt` samples whose squares sum to s (the reference: torch.sqrt(torch.mean(chunk ** 2)) in fp32)
__device__ __forceinline__ float qw_rms32(double s, double cnt) { return sqrtf((float)(s / cnt)); }

__global__ void k_qwen_plan(const char* __restrict__ len_base, int len_stride, int n_clips, int half,
                            int pieces_per_clip, const double* __restrict__ mom, QwenPlan* __restrict__ plan) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_clips) return;
  const int n = *reinterpret_cast<const int32_t*>(len_base + (size_t)c * len_stride);
  QwenPlan& P = plan[c];
  P.mode = 0; P.n_windows = 0; P.gain = 1.f; P.pad = 0.f;
  if (n <= 0) return;
  const double* m = mom + (size_t)c * pieces_per_clip * 3;
  const int n_pieces = (n + half - 1) / half;
  double total = 0.0;
  for (int h = 0; h < n_pieces; ++h) total += m[3 * h];
  if (qw_rms32(total, (double)n) < 1e-8f) return;                        // qwen.py:288-289: unchanged
  const long long window = 2LL * half;                                   // int(sr * 2.0)
  int mode = 1;
  double energy = total;                                                 // sum of squares after the envelope
  if (n > 2 * window) {                                                  // :296
    const int nw = (int)(n / window);                                    // >= 2
    double g[QW_MAX_WINDOWS], sm[QW_MAX_WINDOWS];
    const double ref = (double)qw_rms32(m[0] + m[3], (double)window);
    if (nw <= QW_MAX_WINDOWS && ref >= 1e-8) {                           // :343-345
      const double cap = pow(10.0, 18.0 / 20.0);
      double gmax = -1e300, gmin = 1e300;
      for (int i = 0; i < nw; ++i) {
        const double r = (double)qw_rms32(m[3 * (2 * i)] + m[3 * (2 * i + 1)], (double)window);
        g[i] = (r < 1e-8) ? 1.0 : fmin(ref / r, cap);                    // :351-357
        gmax = fmax(gmax, g[i]); gmin = fmin(gmin, g[i]);
      }
      if (gmax - gmin >= 0.05) {                                         // :361-363
        for (int pass = 0; pass < 2; ++pass) {                           // :366-371
          for (int i = 0; i < nw; ++i) sm[i] = g[i];
          for (int i = 1; i < nw - 1; ++i) sm[i] = (g[i - 1] + g[i] + g[i + 1]) / 3.0;
          for (int i = 0; i < nw; ++i) g[i] = sm[i];
        }
        mode = 2;
        P.n_windows = nw;
        for (int i = 0; i < nw; ++i) P.knots[i] = g[i];
        for (int i = 0; i + 1 < nw; ++i) P.slopes[i] = (g[i + 1] - g[i]) / (double)window;
        // energy after the envelope, piece by piece: piece h = [h*half, (h+1)*half)
        //   h == 0                     : env = g[0]                      (np.interp clamps left of the first centre)
        //   h >= 2*nw - 1              : env = g[nw-1]                   (right of the last centre, incl. the ragged tail)
        //   h = 2k+1, 2k+2 (k < nw-1)  : env = g[k] + (g[k+1] - g[k]) * (t - centre_k) / window
        energy = 0.0;
        for (int h = 0; h < n_pieces; ++h) {
          const double S0 = m[3 * h], S1 = m[3 * h + 1], S2 = m[3 * h + 2];
          double A, B = 0.0;
          if (h == 0) A = g[0];
          else if (h >= 2 * nw - 1) A = g[nw - 1];
          else {
            const int k = (h - 1) >> 1;
            B = (g[k + 1] - g[k]) / (double)window;
            A = g[k] + (((h - 1) & 1) ? B * (double)half : 0.0);
          }
          energy += A * A * S0 + 2.0 * A * B * S1 + B * B * S2;
        }
      }
    }
  }
  const float rms = qw_rms32(energy, (double)n);                         // :302
  float gain = 1.f;
  if (rms > 1e-8f) {
    const double current_db = (double)(20.0f * log10f(rms));             // fp32 tensor ops, then .item()
    gain = (float)pow(10.0, (-23.0 - current_db) / 20.0);                // python floats, then fp32 scalar multiply
  }
  P.gain = gain;
  P.mode = mode;
}

__global__ void __launch_bounds__(QW_THREADS)
k_qwen_apply(const float* __restrict__ x, const int64_t* __restrict__ off, const char* __restrict__ len_base,
             int len_stride, int half, const QwenPlan* __restrict__ plan, float* __restrict__ y,
             const int64_t* __restrict__ y_off, int tile) {
  const int c = blockIdx.x;
  const int n = *reinterpret_cast<const int32_t*>(len_base + (size_t)c * len_stride);
  const long long t0 = (long long)blockIdx.y * tile;
  if (t0 >= n) return;
  const int cnt = (int)min((long long)tile, n - t0);
  const float* __restrict__ xs = x + off[c] + t0;
  float* __restrict__ ys = y + y_off[c] + t0;
  const QwenPlan& P = plan[c];
  const int mode = P.mode;
  const float g = P.gain;
  const int nw = P.n_windows;
  const double window = 2.0 * (double)half;
  const double c0 = 0.5 * window, cl = ((double)nw - 0.5) * window;
  const long long W = 2LL * half;
  const long long kt = (t0 >= half) ? (t0 - half) / W : -1;    // segment of the tile's first sample (-1: left of the first centre)
  const long long b1 = half + (kt + 1) * W;                    // the next window centre
  auto env = [&](long long t) -> float {                                  // float(np.interp(t, centres, knots))
    const double td = (double)t;
    if (td <= c0) return (float)P.knots[0];
    if (td >= cl) return (float)P.knots[nw - 1];
    int k = (int)((td - c0) / window);
    if (k > nw - 2) k = nw - 2;
    return (float)(P.slopes[k] * (td - ((double)k + 0.5) * window) + P.knots[k]);   // slope * (x - xp[j]) + fp[j]
  };
  // tanh(z) = 1 - 2 / (exp(2z) + 1) on the SFU (ex2.approx + rcp.approx): absolute error < 3e-7 on outputs bounded
  // by 1 (libm's tanhf costs ~30 instructions a sample and made this kernel issue-bound); saturates correctly.
  auto tail = [&](float v) -> float {
    v = __fmul_rn(v, g);
    const float e2 = __expf(v * (2.0f / 0.95f));            // exp(2 * (v / 0.95)); the quotient's last ulp is immaterial here
    return __fmul_rn(1.0f - __fdividef(2.0f, e2 + 1.0f), 0.95f);
  };
  auto fin = [&](float v, long long t) -> float {
    if (mode == 0) return v;
    if (mode == 2) v = __fmul_rn(v, env(t));
    return tail(v);
  };
  const bool al = ((reinterpret_cast<uintptr_t>(xs) | reinterpret_cast<uintptr_t>(ys)) & 15u) == 0;
  const int n4 = al ? (cnt >> 2) : 0;
  // four 128-bit pieces per thread in flight (the stream is bound by the bytes in flight per SM)
  constexpr int QU = 4;
  for (int q0 = threadIdx.x; q0 < n4; q0 += QU * QW_THREADS) {
    float4 vv[QU];
#pragma unroll
    for (int u = 0; u < QU; ++u) {
      const int q = q0 + u * QW_THREADS;
      if (q < n4) vv[u] = ldg_stream4(xs + 4 * q);
    }
#pragma unroll
    for (int u = 0; u < QU; ++u) {
      const int q = q0 + u * QW_THREADS;
      if (q >= n4) break;
      float4 v = vv[u];
      const long long t = t0 + 4 * q;
      if (mode == 2) {
        // four consecutive samples almost always share an interpolation segment, and the tile spans at most two
        // (kt, kt + 1: found once per CTA, no per-piece division); k = -1 / >= nw-1 are np.interp's clamped ends
        long long k = -2;
        if (t + 3 < b1) k = kt; else if (t >= b1 && t + 3 < b1 + W) k = kt + 1;
        if (k == -1 || (k >= nw - 1 && k != -2)) {
          const float e = (float)P.knots[k < 0 ? 0 : nw - 1];
          v.x = __fmul_rn(v.x, e); v.y = __fmul_rn(v.y, e); v.z = __fmul_rn(v.z, e); v.w = __fmul_rn(v.w, e);
        } else if (k >= 0) {
          const double sl = P.slopes[k], fk = P.knots[k], u0 = (double)(t - (half + k * W));
          v.x = __fmul_rn(v.x, (float)(sl * u0 + fk));
          v.y = __fmul_rn(v.y, (float)(sl * (u0 + 1.0) + fk));
          v.z = __fmul_rn(v.z, (float)(sl * (u0 + 2.0) + fk));
          v.w = __fmul_rn(v.w, (float)(sl * (u0 + 3.0) + fk));
        } else {
          v.x = __fmul_rn(v.x, env(t)); v.y = __fmul_rn(v.y, env(t + 1));
          v.z = __fmul_rn(v.z, env(t + 2)); v.w = __fmul_rn(v.w, env(t + 3));
        }
      }
      if (mode != 0) { v.x = tail(v.x); v.y = tail(v.y); v.z = tail(v.z); v.w = tail(v.w); }
      stg_stream4(ys + 4 * q, v);
    }
  }
  for (int i = 4 * n4 + threadIdx.x; i < cnt; i += QW_THREADS) ys[i] = fin(xs[i], t0 + i);
}

size_t qwen_workspace_bytes(int n, int64_t max_len, int sr) {
  const int half = sr > 0 ? sr : 1;
  const int64_t pieces = (max_len + half - 1) / half + 1;
  return align_up(sizeof(QwenPlan) * (size_t)(n > 0 ? n : 1), 256) +
         align_up(sizeof(double) * 3 * (size_t)(n > 0 ? n : 1) * (size_t)pieces, 256);
}

cudaError_t launch_qwen_postprocess(const float* x, const int64_t* off, const int32_t* len, int len_stride_bytes,
                                    int n, int64_t max_len, int sr, float* y, const int64_t* y_off,
                                    void* workspace, cudaStream_t st, LaunchCtx* lc) {
  if (n <= 0 || max_len <= 0) return cudaSuccess;
  const int half = sr;                                                     // window = int(sr * 2.0) = 2 * sr
  const int64_t pieces = (max_len + half - 1) / half + 1;
  if (max_len / (2LL * half) > QW_MAX_WINDOWS || pieces > 65535) return cudaErrorInvalidValue;
  char* wsb = (char*)workspace;
  QwenPlan* plan = (QwenPlan*)wsb;
  double* mom = (double*)(wsb + align_up(sizeof(QwenPlan) * (size_t)n, 256));
  const char* lb = reinterpret_cast<const char*>(len);
  const int ls = len_stride_bytes ? len_stride_bytes : (int)sizeof(int32_t);
  lc->begin(KID_QWEN_MOMENTS, st);
  k_qwen_moments<<<dim3((unsigned)n, (unsigned)pieces), QW_THREADS, 0, st>>>(x, off, lb, ls, half, (int)pieces, mom);
  lc->end(st);
  lc->begin(KID_QWEN_PLAN, st);
  k_qwen_plan<<<(n + 127) / 128, 128, 0, st>>>(lb, ls, n, half, (int)pieces, mom, plan);
  lc->end(st);
  const int tile = 16384;
  const unsigned tiles = (unsigned)((max_len + tile - 1) / tile);
  lc->begin(KID_QWEN_APPLY, st);
  k_qwen_apply<<<dim3((unsigned)n, tiles), QW_THREADS, 0, st>>>(x, off, lb, ls, half, plan, y, y_off, tile);
  lc->end(st);
  return cudaGetLastError();
}

}  // namespace rho
