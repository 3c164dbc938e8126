// Host-side constant tables (no CUDA): resample taps, Hann window, slaney mel filterbank,
// FFT twiddles.  Built once per handle and uploaded; also exported through
// rho_b200_host_table so the CPU test-suite can compare them with torchaudio/transformers.
#include <cmath>
#include <vector>
#include <algorithm>
#include "kernels.h"

namespace rho {

// torchaudio/functional/functional.py:1305-1405 for orig=3, new=2 (24 kHz -> 16 kHz), built in
// fp32 like torchaudio does for an fp32 waveform (dtype=waveform.dtype, :1487).
void host_resample_taps(float* out) {
  const int orig = 3, nw = 2, lpw = 6;
  const double base = std::min(orig, nw) * 0.99;           // python double
  const int width = (int)std::ceil(lpw * orig / base);    // 10
  const float basef = (float)base;
  const float scale = (float)(base / orig);
  const float pif = (float)M_PI;
  for (int p = 0; p < nw; ++p) {
    const float ph = (float)(-p) / (float)nw;
    for (int i = 0; i < 2 * width + orig; ++i) {
      const float idx = (float)(i - width) / (float)orig;
      float t = ph + idx;
      t = t * basef;
      t = std::min(std::max(t, (float)-lpw), (float)lpw);
      const float wa = ((t * pif) / (float)lpw) / 2.0f;
      const float c = cosf(wa);
      const float window = c * c;
      t = t * pif;
      const float sinc = (t == 0.0f) ? 1.0f : sinf(t) / t;
      out[p * RS_TAPS + i] = sinc * (window * scale);
    }
  }
}

// The same for any reduced ratio orig:new (torchaudio/functional/functional.py:1305-1405): width and the
// [new][2*width + orig] tap table, every step in fp32 as torchaudio does for an fp32 waveform.
int host_resample_width(int orig, int nw) {
  const double base = std::min(orig, nw) * 0.99;
  return (int)std::ceil(6 * orig / base);
}
void host_resample_taps_general(int orig, int nw, float* out) {
  const int lpw = 6;
  const double base = std::min(orig, nw) * 0.99;
  const int width = (int)std::ceil(lpw * orig / base);
  const int K = 2 * width + orig;
  const float basef = (float)base;
  const float scale = (float)(base / orig);
  const float pif = (float)M_PI;
  for (int p = 0; p < nw; ++p) {
    const float ph = (float)(-p) / (float)nw;
    for (int i = 0; i < K; ++i) {
      const float idx = (float)(i - width) / (float)orig;
      float t = ph + idx;
      t = t * basef;
      t = std::min(std::max(t, (float)-lpw), (float)lpw);
      const float wa = ((t * pif) / (float)lpw) / 2.0f;
      const float c = cosf(wa);
      const float window = c * c;
      t = t * pif;
      const float sinc = (t == 0.0f) ? 1.0f : sinf(t) / t;
      out[(size_t)p * K + i] = sinc * (window * scale);
    }
  }
}

// The same taps, but only those inside the Hann window of each phase: tap i of phase p sits at
// t = (-p / nw + (i - width) / orig) * base and the window clamps |t| to 6, where the fp32 tap is ~1e-23.  The W =
// 2 * width + 2 taps from ilo[p] = floor(p * orig / nw + width - 6 * orig / base) cover every tap with |t| < 6.
void host_resample_taps_windowed(int orig, int nw, int width, int W, float* taps, int* ilo) {
  const int lpw = 6;
  const double base = std::min(orig, nw) * 0.99;
  const int K = 2 * width + orig;
  const float basef = (float)base;
  const float scale = (float)(base / orig);
  const float pif = (float)M_PI;
  for (int p = 0; p < nw; ++p) {
    const double centre = (double)p * orig / nw + width;
    int lo = (int)std::floor(centre - lpw * orig / base);
    if (lo < 0) lo = 0;
    ilo[p] = lo;
    const float ph = (float)(-p) / (float)nw;
    for (int w = 0; w < W; ++w) {
      const int i = lo + w;
      float v = 0.f;
      if (i < K) {
        const float idx = (float)(i - width) / (float)orig;
        float t = ph + idx;
        t = t * basef;
        t = std::min(std::max(t, (float)-lpw), (float)lpw);
        const float wa = ((t * pif) / (float)lpw) / 2.0f;
        const float c = cosf(wa);
        const float window = c * c;
        t = t * pif;
        const float sinc = (t == 0.0f) ? 1.0f : sinf(t) / t;
        v = sinc * (window * scale);
      }
      taps[(size_t)p * W + w] = v;
    }
  }
}

// Tables of the pitch shifter: FFT twiddles, torch.hann_window(512), and phase_advance =
// torch.linspace(0, pi * 128, 257) as torch's fp32 CPU kernel makes it: step = float(end) / 256 in fp32, the first
// half float(step * i), the second half end - step * (256 - i) as one fused multiply-subtract (oracle/pitch.py
// linspace_f32 pins this element by element against torch).
void host_pitch_tables(float* w256, float* w512, float* hann512, float* padv) {
  for (int k = 0; k < 256; ++k) {
    w256[2 * k] = (float)std::cos(2.0 * M_PI * k / 256);
    w256[2 * k + 1] = (float)(-std::sin(2.0 * M_PI * k / 256));
  }
  for (int k = 0; k <= 256; ++k) {
    w512[2 * k] = (float)std::cos(2.0 * M_PI * k / 512);
    w512[2 * k + 1] = (float)(-std::sin(2.0 * M_PI * k / 512));
  }
  for (int n = 0; n < 512; ++n) hann512[n] = (float)(0.5 - 0.5 * std::cos(2.0 * M_PI * n / 512));
  const int n = 257;
  const float end = (float)(M_PI * 128);
  const float step = end / (float)(n - 1);
  for (int i = 0; i < n; ++i) {
    volatile float prod = step * (float)i;      // volatile: keep the product a separately rounded fp32 value
    padv[i] = i < n / 2 ? (float)prod : std::fmaf(-step, (float)(n - 1 - i), end);
  }
}

// torch.hann_window(400) (periodic)
void host_hann(float* out) {
  for (int n = 0; n < N_FFT; ++n) out[n] = (float)(0.5 - 0.5 * std::cos(2.0 * M_PI * n / N_FFT));
}

// Stage-1 twiddles of the 20 x 20 FFT, laid out [k1][n2] = exp(-2*pi*i*n2*k1/400) so that the 20 threads
// of an FFT group read consecutive entries (conflict-free, and the same address in every group).
void host_twiddles(float* out) {
  for (int k1 = 0; k1 < 20; ++k1)
    for (int n2 = 0; n2 < 20; ++n2) {
      const int k = k1 * n2;
      out[2 * (k1 * 20 + n2)] = (float)std::cos(2.0 * M_PI * k / N_FFT);
      out[2 * (k1 * 20 + n2) + 1] = (float)(-std::sin(2.0 * M_PI * k / N_FFT));
    }
}

static double hz_to_mel(double f) {
  if (f >= 1000.0) return 15.0 + std::log(f / 1000.0) * (27.0 / std::log(6.4));
  return 3.0 * f / 200.0;
}
static double mel_to_hz(double m) {
  if (m >= 15.0) return 1000.0 * std::exp((std::log(6.4) / 27.0) * (m - 15.0));
  return 200.0 * m / 3.0;
}

// transformers/audio_utils.py:453-544 with mel_scale="slaney", norm="slaney", 0..8000 Hz, 201 bins
// of a 16 kHz / 400-point FFT; float64 then cast to fp32 (feature_extraction_whisper.py:152).
void host_mel_filterbank(int n_mels, float* out) { host_mel_filterbank_bins(n_mels, N_BINS, out); }

// the same triangles over n_bins = n_fft / 2 + 1 bins of a 16 kHz signal (librosa.filters.mel(sr=16000, n_fft, n_mels))
void host_mel_filterbank_bins(int n_mels, int n_bins, float* out) {
  const double m_lo = hz_to_mel(0.0), m_hi = hz_to_mel(8000.0);
  std::vector<double> centres(n_mels + 2);
  for (int i = 0; i < n_mels + 2; ++i) {
    // numpy.linspace: start + i*step with step = (stop-start)/div; last point pinned to stop
    const double step = (m_hi - m_lo) / (n_mels + 1);
    const double m = (i == n_mels + 1) ? m_hi : m_lo + i * step;
    centres[i] = mel_to_hz(m);
  }
  for (int k = 0; k < n_bins; ++k) {
    const double step = 8000.0 / (n_bins - 1);
    const double f = (k == n_bins - 1) ? 8000.0 : 0.0 + k * step;
    for (int m = 0; m < n_mels; ++m) {
      const double down = -(centres[m] - f) / (centres[m + 1] - centres[m]);
      const double up = (centres[m + 2] - f) / (centres[m + 2] - centres[m + 1]);
      double v = std::max(0.0, std::min(down, up));
      v *= 2.0 / (centres[m + 2] - centres[m]);
      out[m * n_bins + k] = (float)v;
    }
  }
}

// Tables of the MFCC front end: torch / scipy periodic hann(2048), W1024 and W2048 twiddles, and the first 13 rows of the
// orthonormal DCT-II over 128 inputs (scipy.fftpack.dct(type=2, norm="ortho")).
void host_mfcc_tables(float* hann2048, float* w1024, float* w2048, float* dct) {
  for (int n = 0; n < 2048; ++n) hann2048[n] = (float)(0.5 - 0.5 * std::cos(2.0 * M_PI * n / 2048));
  for (int k = 0; k < 1024; ++k) {
    w1024[2 * k] = (float)std::cos(2.0 * M_PI * k / 1024);
    w1024[2 * k + 1] = (float)(-std::sin(2.0 * M_PI * k / 1024));
  }
  for (int k = 0; k <= 1024; ++k) {
    w2048[2 * k] = (float)std::cos(2.0 * M_PI * k / 2048);
    w2048[2 * k + 1] = (float)(-std::sin(2.0 * M_PI * k / 2048));
  }
  for (int k = 0; k < 13; ++k)
    for (int m = 0; m < 128; ++m) {
      double v = std::cos(M_PI * k * (2.0 * m + 1.0) / 256.0) * std::sqrt(2.0 / 128.0);
      if (k == 0) v *= std::sqrt(0.5);
      dct[k * 128 + m] = (float)v;
    }
}

}  // namespace rho
