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
//   span  : x[start + 240*t0 - 312 .. +8068) staged by ONE TMA bulk copy (32 KB, mbarrier completion) when the
//           span lies inside the clip, else zero-filling LDGSTS; prefetched under the previous batch's FFT -- across
//           tiles too: thread 0 resolves the half's next unit one tile ahead (FzNext)
//   apply : interior batches (no fade, no clip edge, one decay zone): y = x - dc straight from registers to
//           HBM, the span keeps the raw x and the FIR folds dc in (FIR(x - dc) = FIR(x) - dc * sum(taps));
//           edge batches: in place (x-dc)*fade -> y
//   FIR   : 2 phases x 23 taps as packed fp32x2 dot products (FFMA2, taps in uniform registers) from the
//           span -> 5360 samples of the 16 kHz slab (reflection fixed up in smem)
//   FFT   : as k_logmel_frames (two real frames per 400-point complex FFT, 20 x 20, packed fp32x2 butterflies)
//   mel   : the sparse filterbank as a table in constant memory walked by a small loop (FZ_MEL_TABLE), SFU log2,
//           ordered-int atomicMax of the clip maximum
// MODE 1 / 2 of the template: features of finished audio in y / items joined from several segments (see the kernel)
#include <algorithm>
#include <vector>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include "logmel_dev.cuh"

namespace rho {

// Taps as float2 pairs for the packed dot products.  With V[k] = (v[2k], v[2k+1]) the four outputs are
//   o00 = sum_k V[k] . A0[k]  (k = 0..9)    A0[k] = (k0[2k],   k0[2k+1])       k0[0] := 0
//   o10 = sum_k V[k] . B0[k]  (k = 2..11)   B0[k] = (k0[2k-3], k0[2k-2])       k0[20] := 0
//   o01 = sum_k V[k] . A1[k]  (k = 1..10)   A1[k] = (k1[2k],   k1[2k+1])       k1[2], k1[21] := 0
//   o11 = sum_k V[k] . B1[k]  (k = 3..11)   B1[k] = (k1[2k-3], k1[2k-2])
// (taps 0, 20..22 of phase 0 and 0..2, 21..22 of phase 1 sit on the window clamp, |k| ~ 3e-24: dropped).
// c_fsum[p] = sum of the taps of phase p that are used, for folding the DC offset into the FIR.
__constant__ float2 c_fA0[12], c_fB0[12], c_fA1[12], c_fB1[12];
__constant__ float c_fsum[2];
__constant__ float c_ftaps[2][RS_TAPS];

cudaError_t upload_fused_taps(const float* taps) {
  const float* k0 = taps;
  const float* k1 = taps + RS_TAPS;
  auto t0 = [&](int i) { return (i >= 1 && i <= 19) ? k0[i] : 0.f; };
  auto t1 = [&](int i) { return (i >= 3 && i <= 20) ? k1[i] : 0.f; };
  float2 A0[12], B0[12], A1[12], B1[12];
  for (int k = 0; k < 12; ++k) {
    A0[k] = make_float2(t0(2 * k), t0(2 * k + 1));
    B0[k] = make_float2(t0(2 * k - 3), t0(2 * k - 2));
    A1[k] = make_float2(t1(2 * k), t1(2 * k + 1));
    B1[k] = make_float2(t1(2 * k - 3), t1(2 * k - 2));
  }
  float sum[2] = {0.f, 0.f};
  double s0 = 0.0, s1 = 0.0;
  for (int i = 1; i <= 19; ++i) s0 += (double)k0[i];
  for (int i = 3; i <= 20; ++i) s1 += (double)k1[i];
  sum[0] = (float)s0; sum[1] = (float)s1;
  cudaError_t e;
  if ((e = cudaMemcpyToSymbol(c_ftaps, taps, sizeof(float) * 2 * RS_TAPS)) != cudaSuccess) return e;
  if ((e = cudaMemcpyToSymbol(c_fA0, A0, sizeof(A0))) != cudaSuccess) return e;
  if ((e = cudaMemcpyToSymbol(c_fB0, B0, sizeof(B0))) != cudaSuccess) return e;
  if ((e = cudaMemcpyToSymbol(c_fA1, A1, sizeof(A1))) != cudaSuccess) return e;
  if ((e = cudaMemcpyToSymbol(c_fB1, B1, sizeof(B1))) != cudaSuccess) return e;
  return cudaMemcpyToSymbol(c_fsum, sum, sizeof(sum));
}

// FZ_MEL_TABLE: the sparse filterbank as DATA in constant memory (every lane of a warp works on the same (row, bin), so
// the weight is one broadcast constant load) walked by a small loop, instead of 1 900 instructions of straight-line FFMA
// with immediate weights that each of the ten warps of a half streams a tenth of, once per batch: the immediates are one
// instruction per weight cheaper, but their code is 30 KB of a ~75 KB loop body against a 32 KB instruction cache.
// Layout: ONE stream of float4 per bank, walked front to back by the ten warps of a half (warp w starts at part4[w] and
// owns rows [part[w], part[w + 1])).  Rows go in bundles of FZ_MEL_ROWS consecutive rows that are walked together -- so
// many independent fmaf chains per lane instead of one: a header {byte offset of each row's first bin in a power row}
// + {bytes of power every row of the bundle covers}, then the weights, group of four by group of four, the rows of the
// bundle interleaved.  Rows are padded with zeros to the bundle's longest row and to a multiple of four
// (fmaf(0, p, a) == a for the finite, non-negative powers: the sums keep their bits); the padding stays inside the 201
// bins; a bundle that sticks out of its part is filled with all-zero rows whose results are not stored.
// Rows per bundle: 2 for the 80-band bank, 4 for the 128-band one (measured: C2 0.883 / 0.888 ms with 2 / 4, C4 7.82 / 7.72)
#ifndef FZ_MEL_ROWS_80
#define FZ_MEL_ROWS_80 2
#endif
#ifndef FZ_MEL_ROWS_128
#define FZ_MEL_ROWS_128 4
#endif
__host__ __device__ constexpr int fz_mel_rows(int which) { return which == 0 ? FZ_MEL_ROWS_80 : FZ_MEL_ROWS_128; }
__host__ __device__ constexpr int fz_mel_hdr4(int which) { return (fz_mel_rows(which) + 1 + 3) / 4; }   // float4 per bundle header
constexpr int FZ_MEL_TAB4 = FUSED_MEL_STREAM_FLOAT4; // float4 per bank (checked when the stream is built)
__constant__ float4 c_mel_tab[2][FZ_MEL_TAB4];
__constant__ int c_mel_part[2][12];                // rows [part[w], part[w + 1]) belong to warp w of a half
__constant__ int c_mel_part4[2][12];               // ... and start at this float4 of the stream

// Host only: builds the stream of one bank (which = 0: 80 bands, 1: 128) into tab[FZ_MEL_TAB4] and the warps' row / stream
// starts into part[11] / part4[11]; rows_per_bundle (optional) receives the bundle size.  Returns the float4 used, or -1.
// Also behind rho_b200_host_mel_stream, so that the packing can be checked without a GPU (tests/test_cabi.py).
int build_fused_mel_stream(int which, int n_mels, const int* lo, const int* cnt, const int* wofs, const float* w, int nnz,
                           float4* tab, int* part, int* part4, int* rows_per_bundle) {
  if (which < 0 || which > 1 || n_mels > 128 || nnz > 416) return -1;
  const int R = fz_mel_rows(which), HDR4 = fz_mel_hdr4(which);
  if (rows_per_bundle) *rows_per_bundle = R;
  constexpr int PARTS = LM_THREADS / 32;
  for (int i = 0; i < FZ_MEL_TAB4; ++i) tab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = 0; i <= PARTS; ++i) { part[i] = 0; part4[i] = 0; }
  // Contiguous groups of rows, one per warp, chosen so that the SLOWEST warp is as fast as possible (the mel phase ends
  // at a barrier): a small dynamic program over the cost the loop below has per bundle -- ~14 instructions of set-up,
  // 12.5 per row and group of four weights (every row of a bundle walks the bundle's longest row), 10 per row for the
  // logarithm and the store.  (The running-sum split this replaces left the slowest warp of the 128-band bank 35 %
  // above the mean; this one 19 %, the 80-band bank 12 % instead of 15 %.)
  auto part_cost = [&](int a, int b) {
    double c = 0.0;
    for (int m0 = a; m0 < b; m0 += R) {
      int g = 0;
      for (int r = 0; r < R && m0 + r < b; ++r) g = std::max(g, (cnt[m0 + r] + 3) / 4);
      c += 14.0 + 12.5 * R * g + 10.0 * R;
    }
    return c;
  };
  // Measured, same box: the 128-band bank gains 4 % from it (C4: 7.70 -> 7.40 ms), the 80-band bank loses 1 % against the
  // running-sum split (C2: 0.878 -> 0.887 ms; the model is off for bundles of two), so each bank keeps what is faster.
#ifndef FZ_MEL_PART_DP
#define FZ_MEL_PART_DP 2                             // bit `which` set: dynamic program for that bank
#endif
  if (!((FZ_MEL_PART_DP >> which) & 1)) {           // running-sum split (groups + 2 per row)
    long long total = 0, acc = 0;
    for (int m = 0; m < n_mels; ++m) total += (cnt[m] + 3) / 4 + 2;
    int nxt = 1;
    for (int m = 0; m < n_mels; ++m) {
      acc += (cnt[m] + 3) / 4 + 2;
      if (nxt < PARTS && acc * PARTS >= total * nxt) part[nxt++] = m + 1;
    }
    while (nxt <= PARTS) part[nxt++] = n_mels;
  } else {
    std::vector<std::vector<double>> best(PARTS + 1, std::vector<double>(n_mels + 1, 1e300));
    std::vector<std::vector<int>> from(PARTS + 1, std::vector<int>(n_mels + 1, 0));
    best[0][0] = 0.0;
    for (int pt = 1; pt <= PARTS; ++pt)
      for (int i = 0; i <= n_mels; ++i)
        for (int j = 0; j <= i; ++j) {
          if (best[pt - 1][j] >= 1e300) continue;
          const double v = std::max(best[pt - 1][j], part_cost(j, i));
          if (v < best[pt][i]) { best[pt][i] = v; from[pt][i] = j; }
        }
    int i = n_mels;
    part[PARTS] = n_mels;
    for (int pt = PARTS; pt >= 1; --pt) { i = from[pt][i]; part[pt - 1] = i; }
  }
  if (const char* ov = getenv(which == 0 ? "RHO_FUSED_MEL_PARTS_80" : "RHO_FUSED_MEL_PARTS_128")) {   // A/B tool: "0,10,20,...,80"
    std::vector<int> v;
    for (const char* q = ov; *q;) { v.push_back(atoi(q)); while (*q && *q != ',') ++q; if (*q) ++q; }
    bool ok = (int)v.size() == PARTS + 1 && v.front() == 0 && v.back() == n_mels;
    for (size_t i = 1; ok && i < v.size(); ++i) ok = v[i] >= v[i - 1];
    if (ok) for (int i = 0; i <= PARTS; ++i) part[i] = v[i];
  }
  int n4 = 0;
  for (int pt = 0; pt < PARTS; ++pt) {
    part4[pt] = n4;
    for (int m0 = part[pt]; m0 < part[pt + 1]; m0 += R) {
      int groups = 0;
      for (int r = 0; r < R; ++r) if (m0 + r < part[pt + 1]) groups = std::max(groups, (cnt[m0 + r] + 3) / 4);
      if (n4 + HDR4 + R * groups > FZ_MEL_TAB4 || 4 * groups > N_BINS) return -1;
      unsigned hdr[8] = {};
      for (int r = 0; r < R; ++r) {
        int first = 0;
        if (m0 + r < part[pt + 1]) {
          const int m = m0 + r;
          first = lo[m];
          if (first + 4 * groups > N_BINS) first = N_BINS - 4 * groups;      // leading zeros instead of trailing ones
          float* base = reinterpret_cast<float*>(tab + n4 + HDR4);
          for (int j = 0; j < cnt[m]; ++j) {
            const int k = lo[m] - first + j;                                   // position in the padded row
            base[((k / 4) * R + r) * 4 + (k % 4)] = 0.25f * w[wofs[m] + j];   // the spectra are 4 |.|^2: exact scaling
          }
        }
        hdr[r] = 4u * (unsigned)first;
      }
      hdr[R] = 16u * (unsigned)groups;
      memcpy(&tab[n4], hdr, sizeof(float4) * (size_t)HDR4);
      n4 += HDR4 + R * groups;
    }
  }
  part4[PARTS] = n4;
  return n4;
}

cudaError_t upload_fused_mel(int which, int n_mels, const int* lo, const int* cnt, const int* wofs, const float* w, int nnz) {
  std::vector<float4> tabv((size_t)FZ_MEL_TAB4);          // (local: handles may be created concurrently)
  float4* tab = tabv.data();
  int part[12] = {}, part4[12] = {};
  if (build_fused_mel_stream(which, n_mels, lo, cnt, wofs, w, nnz, tab, part, part4, nullptr) < 0) return cudaErrorInvalidValue;
  cudaError_t e;
  if ((e = cudaMemcpyToSymbol(c_mel_tab, tab, sizeof(float4) * FZ_MEL_TAB4, sizeof(float4) * FZ_MEL_TAB4 * which)) != cudaSuccess) return e;
  if ((e = cudaMemcpyToSymbol(c_mel_part4, part4, sizeof(part4), sizeof(part4) * which)) != cudaSuccess) return e;
  return cudaMemcpyToSymbol(c_mel_part, part, sizeof(part), sizeof(part) * which);
}

#ifndef FZ_MEL_TABLE
#define FZ_MEL_TABLE 1
#endif
#ifndef FZ_CLIP_MAJOR
#define FZ_CLIP_MAJOR 0
#endif
// A/B switches (tools/ab_fused.sh builds the variants; the defaults are the product configuration)
#ifndef FZ_BULK
#define FZ_BULK 1          // span staged by one TMA bulk copy when it lies inside the clip
#endif
#ifndef FZ_FAST_APPLY
#define FZ_FAST_APPLY 1    // interior batches: y from registers, dc folded into the FIR
#endif
#ifndef FZ_FFMA2_FIR
#define FZ_FFMA2_FIR 1     // FIR as packed fp32x2 dot products
#endif
#ifndef FZ_FIR_QPT
#define FZ_FIR_QPT 5       // consecutive output quads per FIR thread (1: one quad per thread and round, strided)
#endif
#ifndef FZ_SPLIT_APPLY
#define FZ_SPLIT_APPLY 1   // interior batches: warps 8-9 write y while warps 0-7 run the FIR (instead of every thread doing both in turn)
#endif
#ifndef FZ_INLINE_NORM
#define FZ_INLINE_NORM 1   // the half that finishes a clip last writes the constant fill of its zero-padding frames at once: k_logmel_norm 0.249 -> 0.113 ms, this kernel 0.866 -> 0.975 ms, step -2 % (0: k_logmel_norm writes the fill; sliced / dedicated-warp variants: profiles/ncu_r01_v7_summary.md)
#endif
#ifndef FZ_TC_MEL
// 0: mel projection on the CUDA cores (the bank is 97.5 % zeros; FZ_MEL_TABLE)        -> 0.864 ms   <- product
// 1: on the tensor cores inside this kernel (tcgen05, 3xTF32, filterbank in TMEM)     -> 1.033 ms
// 2: same with two products, P rounded to TF32 (accuracy 4.2e-5 -> 6.6e-5 vs float64) -> 0.981 ms
// Both tensor-core variants pass every parity test; they lose because the power spectra have to be re-laid out
// as hi / lo UMMA operands in shared memory (44 scattered stores per thread and batch instead of 20, plus 80 KB of
// operand reads by the tensor core) on the resource this kernel is shortest of, and 78 M128 N32 K8 MMAs per
// 32 frames keep the accumulator busy well into the next batch (profiles/ab_r01_tc_mel.log).
#define FZ_TC_MEL 0
#endif
#ifndef FZ_DEFER_TILE_END
#define FZ_DEFER_TILE_END 1   // the tile-end atomics are issued and NOT waited for: their results (am I the half that finished
                              // this clip last?) are consumed in the first batch of the half's next tile, where the fill of the
                              // padding constant then happens; two barriers and the atomic round trip leave every tile end
#endif
bool fused_inline_norm() { return FZ_INLINE_NORM != 0; }
#ifndef FZ_HALVES_N
#define FZ_HALVES_N 2              // independent 320-thread instruction streams per SM (1: A/B experiment)
#endif
constexpr int FZ_HALVES = FZ_HALVES_N;
constexpr int FZ_THREADS = LM_THREADS * FZ_HALVES;          // 640
constexpr int FZ_LEAD = 312;                                // span starts 312 samples before the batch's own range
constexpr int FZ_SPAN = 8068;                               // 24 kHz samples staged per batch (multiple of 4)
constexpr int FZ_OWN = 240 * LM_BF;                         // 7680 output samples owned by a batch
constexpr int FZ_DPAIRS = LM_SLAB / 4;                      // 1340 groups of 4 consecutive 16 kHz samples

constexpr int FZ_FIR_MAIN = 256;                            // split mode: threads (8 warps) that run the FIR five quads at a time
constexpr int FZ_TWS = 22;                                  // float2 per twiddle row (conflict-free 128-bit reads)

// How the frames of a clip are cut into tiles: tile k covers the 32-frame batches [start[k], start[k+1])
// (launch_fused_features picks the sizes per call).  Units are handed out tile-major: the last tiles of the clips,
// which are the short ones, come at the end and level the halves.
constexpr int FZ_MAX_TILES = 64;
struct FzSched {
  int n_tiles;
  int start[FZ_MAX_TILES + 1];
};

// The unit (clip c, tile) a half works on next, resolved by its thread 0 one tile ahead so that a tile
// starts from shared memory instead of a chain of dependent global loads, with its first span already in flight.
struct FzNext {
  long long x_off;                  // seg_off[s] + start: first sample of y in x
  long long y_off;
  int u, c, start, end;
  float dc;
  int tile;
  int prefetched;                   // the unit's first span is a bulk copy already issued on FzHalf::bar
  int s0, ns;                       // MODE 2 (joined items): the item's first segment and its number of segments
};

// MODE 2: what a half keeps of a segment's span record (SegSpan, plan_items) while it works on the item
constexpr int FZ_MAX_JSEG = 8;      // cached spans per item; longer items read the rest from global memory
struct JSeg {
  long long x_base, prev_x;         // offsets in x of the segment's processed sample 0 / of the crossfade tail before it
  int dst, ov, body, pause;
  float dc, dcp;
};

struct alignas(1024) FzHalf {
  float2 fb[LM_GROUPS * LM_FB];     // FFT buffers; afterwards the power spectra (FZ_TC_MEL: as UMMA operands, 512-byte atoms)
  float span[FZ_SPAN];
  float pw[LM_BF * LM_PS];          // 16 kHz slab between FIR and FFT stage 1, then the spectrum exchange
  double redd[2][LM_THREADS / 32];
  float redf[LM_THREADS / 32];
  uint64_t bar;                     // mbarrier the TMA copy of the span completes on
  uint64_t mma_bar;                 // mbarrier the tensor-core mel projection of a batch commits to
  int last;                         // "this half finished its clip last" broadcast
  int fill_max;                     // ... and the clip maximum it read (ordered-int form)
  int next_unit;                    // the tile this half works on next (claimed one tile ahead)
  int pend_c, pend_tiles, pend_T_real, pend_T;   // FZ_DEFER_TILE_END: the clip whose tile-end count is in flight
  int fill_c;                       // ... and the clip whose padding constant this half writes next (-1: none)
  FzNext nd;
  JSeg js[FZ_MAX_JSEG];
};
static_assert(LM_SLAB_SM <= LM_BF * LM_PS, "the slab must fit in the pw region");
static_assert(LM_GROUPS * 201 * 2 <= LM_BF * LM_PS, "the spectrum exchange must fit in the pw region");
static_assert(LM_BF * LM_PS * 4 <= LM_GROUPS * LM_FB * 8, "the power spectra must fit in the fb region");
static_assert((FZ_SPAN * 4) % 16 == 0, "bulk copy size");

struct FzSmem {
  FzHalf h[FZ_HALVES];
  float hannT[N_FFT];               // [n2][n1] = hann[20*n1 + n2]: a thread's 20 window values are contiguous
  float2 twT[20 * FZ_TWS];          // [n2][k1] = W400^(n2*k1), row stride 22
  uint32_t tmem_base;
};
static_assert(sizeof(FzSmem) + 1024 <= 232448, "shared memory budget (227 KB, incl. the slack to align to 1024)");

// ---- tensor-core mel projection (FZ_TC_MEL) ---------------------------------------------------------------
// mel[m][f] = sum_k F[m][k] * P[f][k] for the 32 frames of a batch, as tcgen05.mma.kind::tf32 M128 N32 K8:
//   A = the filterbank (hi | lo, 2 x 208 TMEM columns, rows >= n_mels and bins >= 201 are zero), written once per
//       CTA with tcgen05.st; the kernel is persistent, so once per SM and launch
//   B = the batch's power spectra, K-major with the 64-byte swizzle, written by the spectrum-split step straight
//       into the (dead) FFT buffers as hi (top 19 bits) and lo (remainder): 13 blocks of 16 bins x 32 frames x 64 B
//   D = 128 x 32 fp32 in TMEM (columns 416 + 32 * half), read back by the first four warps of the half one batch
//       later (after the next FIR: the MMAs have long retired) -> log10, clip maximum, 128-byte row stores
// 3xTF32 (Fhi*Phi + Flo*Phi + Fhi*Plo): fp32-class accuracy, < 2e-6 relative (tests/test_gpu_parity.py).
constexpr int TC_KSTEPS = 26;                                 // 26 x 8 = 208 >= 201 bins
constexpr int TC_OPERAND_BYTES = 13 * LM_BF * 64;             // 26624 per hi / lo buffer
constexpr uint32_t TC_COL_WHI = 0, TC_COL_WLO = 208, TC_COL_D = 416, TC_TMEM_COLS = 512;
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(LM_BF >> 3) << 17) | ((128u >> 4) << 24);
static_assert(2 * TC_OPERAND_BYTES <= LM_GROUPS * LM_FB * 8, "hi + lo operands must fit in the fb region");
static_assert(LM_BF == 32, "the MMA shape and the epilogue assume 32 frames per batch");

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ float tc_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }
__device__ __forceinline__ uint32_t tc_elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint32_t a_tmem, uint32_t desc_lo, uint32_t desc_hi,
                                       uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 bd;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b64 bd, {%2, %3};\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bd, %4, p;\n\t}"
               :: "r"(d_tmem), "r"(a_tmem), "r"(desc_lo), "r"(desc_hi), "r"(TC_IDESC), "r"(accumulate) : "memory");
}
// byte offset of power[frame f][bin k] inside an operand buffer: block of 16 bins, row of 64 B, 16-byte chunks
// XOR-swizzled with bits [7,9) of the address (= (f >> 1) & 3)
__device__ __forceinline__ uint32_t tc_off(int f, int k) {
  return ((uint32_t)(k >> 4) << 11) | ((uint32_t)f << 6) | ((((uint32_t)(k >> 2) ^ (uint32_t)(f >> 1)) & 3u) << 4) |
         ((uint32_t)(k & 3) << 2);
}

// Drains the accumulator of the batch that started at frame t0p: the first four warps of the half, one mel row
// per thread.  Waits for the commit of that batch's MMAs first (they were issued about one FIR ago).
template <int NM>
__device__ __noinline__ void fz_tc_epilogue(uint64_t* mma_bar, unsigned parity, uint32_t d_tmem, int tid, int t0p,
                                            int T_real, float* __restrict__ out, long long mel_stride, float* lmax) {
  mbar_wait(mma_bar, parity);
  tc_fence_after();
  const int q = (threadIdx.x >> 5) & 3;                        // the TMEM lane quarter of this warp
  const int m = 32 * q + (tid & 31);
  uint32_t r[32];
  const uint32_t a = d_tmem + ((uint32_t)(32 * q) << 16);
#pragma unroll
  for (int j = 0; j < 2; ++j)
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[16 * j + 0]), "=r"(r[16 * j + 1]), "=r"(r[16 * j + 2]), "=r"(r[16 * j + 3]),
                   "=r"(r[16 * j + 4]), "=r"(r[16 * j + 5]), "=r"(r[16 * j + 6]), "=r"(r[16 * j + 7]),
                   "=r"(r[16 * j + 8]), "=r"(r[16 * j + 9]), "=r"(r[16 * j + 10]), "=r"(r[16 * j + 11]),
                   "=r"(r[16 * j + 12]), "=r"(r[16 * j + 13]), "=r"(r[16 * j + 14]), "=r"(r[16 * j + 15])
                 : "r"(a + 16 * j) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  tc_fence_before();
  if (m >= NM) return;
  float* __restrict__ row = out + (long long)m * mel_stride + t0p;
  float mx = *lmax;
  const bool full = t0p + 32 <= T_real && (mel_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(row) & 15u) == 0);
  if (full) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 v;
      v.x = __log2f(fmaxf(__uint_as_float(r[4 * j + 0]), 1e-10f)) * 0.30102999566398120f;
      v.y = __log2f(fmaxf(__uint_as_float(r[4 * j + 1]), 1e-10f)) * 0.30102999566398120f;
      v.z = __log2f(fmaxf(__uint_as_float(r[4 * j + 2]), 1e-10f)) * 0.30102999566398120f;
      v.w = __log2f(fmaxf(__uint_as_float(r[4 * j + 3]), 1e-10f)) * 0.30102999566398120f;
      mx = fmaxf(fmaxf(mx, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
      *reinterpret_cast<float4*>(row + 4 * j) = make_float4(lm_scaled(v.x), lm_scaled(v.y), lm_scaled(v.z), lm_scaled(v.w));
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (t0p + j < T_real) {
        const float ls = __log2f(fmaxf(__uint_as_float(r[j]), 1e-10f)) * 0.30102999566398120f;
        row[j] = lm_scaled(ls);
        mx = fmaxf(mx, ls);
      }
    }
  }
  *lmax = mx;
}

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

__device__ __forceinline__ void fz_cp_async4(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void fz_cp_async8(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}

__device__ __noinline__ JSeg fz_jseg_global(const SegSpan* __restrict__ gs, int k) {
  const SegSpan sp = gs[k];
  JSeg j;
  j.x_base = sp.x_base; j.prev_x = sp.prev_x; j.dst = sp.dst; j.ov = sp.ov; j.body = sp.body; j.pause = sp.pause;
  j.dc = sp.dc; j.dcp = sp.dcp;
  return j;
}
__device__ __forceinline__ JSeg fz_jseg(const JSeg* __restrict__ cached, const SegSpan* __restrict__ gs, int k) {
  return k < FZ_MAX_JSEG ? cached[k] : fz_jseg_global(gs, k);
}

// MODE 2: a window inside one segment whose samples are not 16-byte aligned at the window start (the position of a
// segment inside its item is arbitrary): 8- or 4-byte LDGSTS pieces.  Out of line: a fifth of the batches, three call sites.
__device__ __noinline__ void fz_stage_pieces(float* __restrict__ span, const float* __restrict__ src, int tid) {
  if ((reinterpret_cast<uintptr_t>(src) & 7u) == 0) {
    for (int q = tid; q < FZ_SPAN / 2; q += LM_THREADS) fz_cp_async8(span + 2 * q, src + 2 * q);
  } else {
    for (int q = tid; q < FZ_SPAN; q += LM_THREADS) fz_cp_async4(span + q, src + q);
  }
  cp_async_commit();
}

// MODE 2, batches that touch a crossfade, a pause, a segment boundary or an end of the item: four samples of the JOINED
// audio at item positions o .. o + 3, computed from the segments exactly as k_gather does (join.cu: equal-power crossfade
// of the DC-free tails, body, pause zeros, then the fades of the item), written to y when the batch owns them, with the
// decay sums.  Out of line: it runs on a few percent of the batches.
__device__ __noinline__ void fz_join_piece(float* e, int o, int n, const JSeg* __restrict__ cached,
                                           const SegSpan* __restrict__ gs, int ns, const float* __restrict__ x, int fade,
                                           bool need_fade, int third, int own_lo, int own_hi, float* __restrict__ ys,
                                           float* a_first, float* a_last) {
  int k = 0;
  JSeg s = fz_jseg(cached, gs, 0);
  if (o >= 0 && o + 3 < n) {
    // most pieces of such a window are still plain body samples of one segment, clear of the fades: no per-sample search
    while (k + 1 < ns && o >= s.dst + s.ov + s.body + s.pause) s = fz_jseg(cached, gs, ++k);
    const int jj = o - s.dst;
    if (jj >= s.ov && jj + 3 < s.ov + s.body && (!need_fade || (o >= fade && o + 3 < n - fade))) {
      const float* __restrict__ xc = x + s.x_base + jj;
#pragma unroll
      for (int i = 0; i < 4; ++i) e[i] = __fsub_rn(xc[i], s.dc);
      if (o >= own_lo && o < own_hi) {                   // own_lo, own_hi and o are multiples of 4
        *reinterpret_cast<float4*>(ys + o) = make_float4(e[0], e[1], e[2], e[3]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (o + i < third) *a_first += e[i] * e[i];
          if (o + i >= n - third) *a_last += e[i] * e[i];
        }
      }
      return;
    }
  }
  for (int i = 0; i < 4; ++i) {
    const int oo = o + i;
    float val = 0.f;
    if (oo >= 0 && oo < n) {
      while (k + 1 < ns && oo >= s.dst + s.ov + s.body + s.pause) s = fz_jseg(cached, gs, ++k);
      const int jj = oo - s.dst;
      if (jj < s.ov) {
        const float fo = cosf(linspace32(0.f, RHO_HALF_PI_F, s.ov, jj));
        const float fi = cosf(linspace32(RHO_HALF_PI_F, 0.f, s.ov, jj));
        const float a = __fmul_rn(__fsub_rn(x[s.prev_x + jj], s.dcp), fo);
        const float b = __fmul_rn(__fsub_rn(x[s.x_base + jj], s.dc), fi);
        val = __fadd_rn(a, b);
      } else if (jj < s.ov + s.body) {
        val = __fsub_rn(x[s.x_base + jj], s.dc);
      }
      if (need_fade) val = __fmul_rn(val, fz_fade_gain(oo, n, fade));
      if (oo >= own_lo && oo < own_hi) {
        ys[oo] = val;
        if (oo < third) *a_first += val * val;
        if (oo >= n - third) *a_last += val * val;
      }
    }
    e[i] = val;
  }
}

// MODE 0: one-segment items, everything above in one pass over x.
// MODE 2: items JOINED from several segments (base_tts.py:912-926), in the same single pass: a batch whose window lies
//         inside one segment's body is the one-segment case with that segment's x and DC; the others (crossfades,
//         pauses, segment boundaries: a few percent) compute their window of the joined audio sample by sample
//         (fz_join_piece).  y is written here: no k_gather pass, no second read of the audio.
// FROM_Y = false: one-segment items, everything above in one pass over x.
// FROM_Y = true : the features of FINISHED items (the output of k_gather: joined, faded audio at y + y_off[c], length
//                 item[c].out_len): the span is read from y, nothing is applied or written back, the 16 kHz signal still
//                 never leaves shared memory -- the path of base_tts.py:912-926 items that have several segments.
// fill_to: the constant of the zero-padding frames is written for frames < fill_to only (3000, or the row length of a
//          compact feature tensor, RHO_V_COMPACT_PAD).
// The constant of the frames that only see zero padding (2/3 of a 10 s clip's features): pure stores by the 320 threads of
// a half.  Out of line: three call sites, none of them on the per-batch path.
template <int NM>
__device__ __noinline__ void fz_fill_padding(float* __restrict__ fout, long long mel_stride, int fT_real, int fT,
                                             int fmax_ordered, int tid) {
  const float mx = ordered_to_float(fmax_ordered);
  const float fill = __fmul_rn(__fadd_rn(fmaxf(-10.0f, __fsub_rn(mx, 8.0f)), 4.0f), 0.25f);
  const int t_lo = (fT_real + 3) & ~3;               // whole 128-bit pieces from here; [T_real, t_lo) is k_logmel_norm's
  const bool vec = (mel_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(fout) & 15u) == 0);
  if (vec) {
    const int T4 = (fT - t_lo) >> 2;
    const float4 f4 = make_float4(fill, fill, fill, fill);
    for (int cp = tid; cp < T4; cp += LM_THREADS) {
      float* col = fout + t_lo + 4 * cp;
#pragma unroll 8
      for (int m = 0; m < NM; ++m) stg_stream4(col + (long long)m * mel_stride, f4);
    }
    for (int i = tid; i < NM * ((fT - t_lo) & 3); i += LM_THREADS) {
      const int rem = (fT - t_lo) & 3, m = i / rem;
      fout[(long long)m * mel_stride + t_lo + 4 * T4 + (i - m * rem)] = fill;
    }
  } else {
    for (int m = 0; m < NM; ++m)
      for (int t = t_lo + tid; t < fT; t += LM_THREADS) fout[(long long)m * mel_stride + t] = fill;
  }
}

template <int NM, int MODE>
__global__ void __launch_bounds__(FZ_THREADS, 1)
k_fused_features(const float* __restrict__ x, const int64_t* __restrict__ seg_off, const SegState* __restrict__ seg,
                 ItemState* __restrict__ item, const int32_t* __restrict__ item_first_seg,
                 float* __restrict__ y, const int64_t* __restrict__ y_off, int fade,
                 const float* __restrict__ g_hann, const float2* __restrict__ g_tw, int pad_frames,
                 float* __restrict__ mel, long long mel_stride, int* __restrict__ clip_max,
                 int32_t* __restrict__ len16_out, int* __restrict__ tiles_done, int* __restrict__ work_counter,
                 const __grid_constant__ FzSched sched, int n_items, const float* __restrict__ mel_dense,
                 int fill_to, const SegSpan* __restrict__ gspan) {
  constexpr bool FROM_Y = MODE == 1;                 // finished audio in y: features only
  constexpr bool JOIN = MODE == 2;                   // items joined from several segments, y written here
  extern __shared__ unsigned char smem_raw[];
  FzSmem& S = *reinterpret_cast<FzSmem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  for (int i = threadIdx.x; i < N_FFT; i += FZ_THREADS) {
    const int r = i / 20, c20 = i - 20 * r;
    S.hannT[c20 * 20 + r] = g_hann[i];               // transpose: [n2][n1]
    S.twT[r * FZ_TWS + c20] = g_tw[i];               // the table is symmetric in (k1, n2): re-stride only
  }
  if (threadIdx.x < FZ_HALVES) {
    mbar_init(&S.h[threadIdx.x].bar, 1); mbar_init(&S.h[threadIdx.x].mma_bar, 1); mbar_fence_init();
  }
#if FZ_TC_MEL
  if (threadIdx.x < 32) {                            // warp 0 allocates all 512 TMEM columns (one CTA per SM)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&S.tmem_base)), "r"(TC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = S.tmem_base;
  if (threadIdx.x < 128) {                           // filterbank -> TMEM: lane = mel row, column = bin, hi | lo
    const int q = threadIdx.x >> 5, m = threadIdx.x;
    const uint32_t rowa = tmem + ((uint32_t)(32 * q) << 16);
    for (int kc = 0; kc < TC_KSTEPS; ++kc) {
      uint32_t hh[8], ll[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = 8 * kc + j;
        const float wv = (m < NM && k < N_BINS) ? mel_dense[m * N_BINS + k] : 0.f;
        const float wh = tc_hi(wv);
        hh[j] = __float_as_uint(wh);
        ll[j] = __float_as_uint(tc_hi(wv - wh));
      }
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                   :: "r"(rowa + TC_COL_WHI + 8 * kc), "r"(hh[0]), "r"(hh[1]), "r"(hh[2]), "r"(hh[3]), "r"(hh[4]), "r"(hh[5]), "r"(hh[6]), "r"(hh[7]) : "memory");
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                   :: "r"(rowa + TC_COL_WLO + 8 * kc), "r"(ll[0]), "r"(ll[1]), "r"(ll[2]), "r"(ll[3]), "r"(ll[4]), "r"(ll[5]), "r"(ll[6]), "r"(ll[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
#endif
  __syncthreads();                                   // the last CTA-wide barrier before the end: the halves run independently
#if FZ_TC_MEL
  tc_fence_after();
#endif

  const int half = threadIdx.x / LM_THREADS;
  const int tid = threadIdx.x - half * LM_THREADS;
  FzHalf& H = S.h[half];
  unsigned parity = 0;                               // phase of H.bar, carried across tiles
#if FZ_TC_MEL
  unsigned mma_parity = 0;                           // phase of H.mma_bar the next epilogue waits for
  const uint32_t d_tmem = tmem + TC_COL_D + (uint32_t)(LM_BF * half);
  const bool epi_warp = tid < 128 && 32 * ((threadIdx.x >> 5) & 3) < NM;   // first four warps of the half: one TMEM lane quarter each
#endif

  // Persistent CTA: one per SM.  Every half claims its own tiles (sched: length chosen per call) from a global counter (tile-major
  // order: the short last tiles of the clips come at the end) and runs at its own pace -- a half never waits
  // for its partner, and no SM is left with one tile more than the others.  The next claim is issued at the
  // start of a tile and consumed at its end, so the atomic's round trip is never exposed.
  const int n_units = n_items * sched.n_tiles;       // unit u = (clip u % n_items, tile u / n_items): big tiles first
  // thread 0 of the half only: resolve unit `un` into H.nd
  auto fill_desc = [&](int un) {
    FzNext d;
    d.u = un; d.c = 0; d.start = 0; d.end = 0; d.dc = 0.f; d.tile = 0; d.prefetched = 0; d.x_off = 0; d.y_off = 0;
    d.s0 = 0; d.ns = 0;
    if (un < n_units) {
      d.tile = un / n_items;                             // tile-major: the last (short) tiles of the clips come at the end
      d.c = un - d.tile * n_items;
      d.y_off = y_off[d.c];
      if (FROM_Y) {
        d.start = 0; d.end = item[d.c].out_len; d.dc = 0.f;
        d.x_off = d.y_off;
      } else if (JOIN) {
        d.start = 0; d.end = item[d.c].out_len; d.dc = 0.f;
        d.s0 = item_first_seg[d.c]; d.ns = item_first_seg[d.c + 1] - d.s0;
      } else {
        const int sn = item_first_seg[d.c];
        const SegState sn_st = seg[sn];
        d.start = sn_st.start; d.end = sn_st.end; d.dc = sn_st.dc;
        d.x_off = seg_off[sn] + sn_st.start;
      }
    }
    H.nd = d;
  };
  // thread 0 only, once the span buffer is dead: issue the bulk copy of the next unit's first span
  auto prefetch_next_tile = [&]() {
    const FzNext d = H.nd;
    if (!FZ_BULK || JOIN || d.u >= n_units) return;
    const int nn = d.end - d.start;
    const int nn16 = nn > 0 ? (int)((2LL * nn + 2) / 3) : 0;
    int Tn, Tn_real, Nn, nvn;
    lm_frame_counts(nn16, pad_frames, &Tn, &Tn_real, &Nn, &nvn);
    const int tt0 = sched.start[d.tile] * LM_BF;
    if (tt0 >= (FROM_Y ? Tn_real : max(Tn_real, (nn + 239) / 240))) return;
    const float* xsn = (FROM_Y ? y : x) + d.x_off;
    const long long j0 = 240LL * tt0 - FZ_LEAD;
    if ((reinterpret_cast<uintptr_t>(xsn) & 15u) != 0 || j0 < 0 || j0 + FZ_SPAN > nn) return;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_arrive_expect_tx(&H.bar, FZ_SPAN * 4);
    bulk_g2s(H.span, xsn + j0, FZ_SPAN * 4, &H.bar);
    H.nd.prefetched = 1;
  };
  auto do_fill = [&](int fc, int fT_real, int fT, int fmax_ordered) {
    fz_fill_padding<NM>(mel + (long long)fc * NM * mel_stride, mel_stride, fT_real, fT, fmax_ordered, tid);
  };
  int pend_before = 0;                               // thread 0: the in-flight result of the tile-end count (FZ_DEFER_TILE_END)
  bool pend = false;
  // thread 0, at a point where the previous tile end is at least one FIR old: was this half the last finisher of that clip?
  auto resolve_pending = [&]() {
    int fc = -1;
    if (pend) {
      if (pend_before == H.pend_tiles - 1) { H.fill_max = atomicMax(&clip_max[H.pend_c], (int)0x80000000); fc = H.pend_c; }
      pend = false;
    }
    H.fill_c = fc;
  };
  if (tid == 0) { H.fill_c = -1; fill_desc(atomicAdd(work_counter, 1)); }
  for (;;) {
  half_sync(half);                                   // the descriptor written by tid 0 is visible
  const FzNext nd = H.nd;
  if (JOIN && nd.u < n_units && tid < min(nd.ns, FZ_MAX_JSEG)) {     // the item's span records, next to the arithmetic
    const SegSpan sp = gspan[nd.s0 + tid];
    JSeg j;
    j.x_base = sp.x_base; j.prev_x = sp.prev_x; j.dst = sp.dst; j.ov = sp.ov; j.body = sp.body; j.pause = sp.pause;
    j.dc = sp.dc; j.dcp = sp.dcp;
    H.js[tid] = j;
  }
  half_sync(half);                                   // ... and read by everybody before it is replaced
  const int u = nd.u;
  if (u >= n_units) break;
  if (tid == 0) H.next_unit = atomicAdd(work_counter, 1);
  const int c = nd.c;
  const int tile = nd.tile;
  SegState st;
  st.start = nd.start; st.end = nd.end; st.dc = nd.dc;
  const int n = st.end - st.start;                   // samples of y
  float dc = st.dc;                                  // MODE 2: of the segment a batch lies in
  const int n16 = n > 0 ? (int)((2LL * n + 2) / 3) : 0;
  int T, T_real, N, n_valid;
  lm_frame_counts(n16, pad_frames, &T, &T_real, &N, &n_valid);
  if (tile == 0 && tid == 0) len16_out[c] = n16;
  const int t_cover = FROM_Y ? T_real : max(T_real, (n + 239) / 240);  // batches needed for the features AND to write all of y
  const int tile_t0 = sched.start[tile] * LM_BF;
  const int tile_batches = sched.start[tile + 1] - sched.start[tile];
  if (tile_t0 >= t_cover) { if (tid == 0) fill_desc(H.next_unit); continue; }

  const float* __restrict__ xs = (FROM_Y ? y : x) + nd.x_off;
  float* __restrict__ ys = y + nd.y_off;
  float* __restrict__ out = mel + (long long)c * NM * mel_stride;
  const int third = n / 3;
  const bool need_fade = fade > 0 && n >= 2 * fade;
  const int edge_lo = need_fade ? fade : 0, edge_hi = need_fade ? n - fade : n;   // [edge_lo, edge_hi): y = x - dc
  const bool xs_al16 = ((reinterpret_cast<uintptr_t>(xs) & 15u) == 0);
  const int g = tid / LM_LANES, lane = tid - g * LM_LANES;
  float2* fb = H.fb + g * LM_FB;
  float lmax = -INFINITY;
  float a_first = 0.f, a_last = 0.f;
#if FZ_TC_MEL
  bool pending = false;                              // a batch's mel projection is in flight on the tensor cores
  int pend_t0 = 0;
#endif
  const SegSpan* __restrict__ gs = JOIN ? gspan + nd.s0 : nullptr;
  const int ns = nd.ns;
  // MODE 2: does the window of the batch at frame t0 lie inside the body of ONE segment (and clear of the item's fades)?
  // Then it is staged from that segment's x like a one-segment clip: *xe = pointer to the sample at item position 0.
  // The windows of a tile move forward only: jk remembers the segment the last one started in.  Evaluated ONCE per
  // batch, when its span is staged (one batch ahead); the batch itself uses the saved answer (nx_*).
  int jk = 0;
  bool nx_staged = false, nx_bulk = false;
  float nx_dc = 0.f;
  auto jclass = [&](int t0, const float** xe, float* dce) {
    const long long j0 = 240LL * t0 - FZ_LEAD;
    if (j0 < edge_lo || j0 + FZ_SPAN > edge_hi) return false;
    JSeg s = fz_jseg(H.js, gs, jk);
    while (jk + 1 < ns && j0 >= (long long)s.dst + s.ov + s.body + s.pause) s = fz_jseg(H.js, gs, ++jk);
    const long long lo = (long long)s.dst + s.ov;
    if (j0 >= lo && j0 + FZ_SPAN <= lo + s.body) { *xe = x + (s.x_base - s.dst); *dce = s.dc; return true; }
    return false;
  };

  // A span that lies inside the clip is ONE bulk copy issued by one thread; spans that stick out
  // (clip edges: zero fill) are staged in 16-byte zero-filling LDGSTS pieces by everybody.
  auto span_is_bulk = [&](int t0) {
    const long long j0 = 240LL * t0 - FZ_LEAD;
    return FZ_BULK && xs_al16 && j0 >= 0 && j0 + FZ_SPAN <= n;
  };
  auto stage_span = [&](int t0) {
    const long long j0 = 240LL * t0 - FZ_LEAD;
    if (JOIN) {
      // a window inside one segment: a bulk copy if that segment's samples are 16-byte aligned at the window start
      // (the position of a segment inside its item is arbitrary), else 8- or 4-byte LDGSTS pieces; any other window is
      // computed sample by sample when its batch starts (fz_join_piece): nothing to stage
      const float* xe; float dce = 0.f;
      nx_staged = jclass(t0, &xe, &dce);
      nx_dc = dce; nx_bulk = false;
      if (!nx_staged) return;
      const float* src = xe + j0;
      const uintptr_t a = reinterpret_cast<uintptr_t>(src);
      if (FZ_BULK && (a & 15u) == 0) {
        nx_bulk = true;
        if (tid == 0) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_arrive_expect_tx(&H.bar, FZ_SPAN * 4);
          bulk_g2s(H.span, src, FZ_SPAN * 4, &H.bar);
        }
      } else {
        fz_stage_pieces(H.span, src, tid);
      }
      return;
    }
    if (span_is_bulk(t0)) {
      if (tid == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic accesses to the span
        mbar_arrive_expect_tx(&H.bar, FZ_SPAN * 4);
        bulk_g2s(H.span, xs + j0, FZ_SPAN * 4, &H.bar);
      }
    } else {
      for (int q = tid; q < FZ_SPAN / 4; q += LM_THREADS) {
        const long long j = j0 + 4 * q;
        int nb = 0;
        if (j >= 0 && j < n) nb = (n - j >= 4) ? 16 : 4 * (int)(n - j);
        cp_async16_zfill(H.span + 4 * q, nb ? xs + j : xs, nb);
      }
      cp_async_commit();
    }
  };
  if (!nd.prefetched) stage_span(tile_t0);

  for (int b = 0; b < tile_batches; ++b) {
    const int t0 = tile_t0 + b * LM_BF;
    if (t0 >= t_cover) break;
    const bool next = (b + 1 < tile_batches) && (t0 + LM_BF < t_cover);
    // the claim issued at the start of the tile has long returned: resolve it during the second batch, or at the
    // last batch's prefetch point if the tile has only one
    if (b == 1 && tid == 0) fill_desc(H.next_unit);
    // Bulk mode needs no barrier here: every thread waits on the mbarrier itself, and nothing this batch writes
    // before its first barrier (the slab, in the pw region) is still read by the previous batch (its power
    // spectra live in the fb region, its partner exchange finished before its last barrier).
    const bool staged = JOIN ? nx_staged : true;             // MODE 2: the span holds the raw samples of ONE segment
    if (JOIN) dc = nx_dc;
    if (JOIN ? nx_bulk : span_is_bulk(t0)) { mbar_wait(&H.bar, parity); parity ^= 1u; } else { cp_async_wait_all(); half_sync(half); }
    const int j0 = 240 * t0 - FZ_LEAD;
    const int own_lo = 240 * t0, own_hi = own_lo + FZ_OWN;
    // interior batch: every sample of the span is plain x - dc, and the owned range lies in ONE decay zone
    const bool in_first = own_hi <= third, in_last = own_lo >= n - third;
    const bool fast = FROM_Y || (FZ_FAST_APPLY && staged && j0 >= edge_lo && j0 + FZ_SPAN <= edge_hi &&
                                 (in_first || own_lo >= third) && (in_last || own_hi <= n - third));
    const bool split = FZ_SPLIT_APPLY && !FROM_Y && fast && t0 < T_real;
    if (FROM_Y) {
      // finished audio: nothing to apply, nothing to write back; samples outside the item were zero-filled by the staging
    } else if (JOIN && !staged) {
      // ---- apply, a batch at a joint (crossfade, pause, segment boundary, end of the item): the window of the joined
      // audio sample by sample into the span; the batch's own samples go to HBM; decay sums
      for (int q = tid; q < FZ_SPAN / 4; q += LM_THREADS) {
        float e[4];
        fz_join_piece(e, j0 + 4 * q, n, H.js, gs, ns, x, fade, need_fade, third, own_lo, own_hi, ys, &a_first, &a_last);
        *reinterpret_cast<float4*>(H.span + 4 * q) = make_float4(e[0], e[1], e[2], e[3]);
      }
      half_sync(half);
    } else if (fast) {
      // ---- apply, interior: registers -> HBM; the span keeps the raw x.  Split mode (a batch that also runs the FIR):
      // the interior apply and the FIR do not depend on each other, so warps 8-9 write all of y (30 pieces per thread) while
      // warps 0-7 start the FIR at once (5 quads per thread; the last 60 quads go to warps 8-9 as well) -- the two kinds of
      // work sit side by side on the SM instead of one after the other in every thread
      float ss = 0.f;
      const float2 ndc = make_float2(-dc, -dc);
      if (!split) {
#pragma unroll
        for (int k = 0; k < FZ_OWN / 4 / LM_THREADS; ++k) {
          const int q = FZ_LEAD / 4 + tid + k * LM_THREADS;
          const float4 v = *reinterpret_cast<const float4*>(H.span + 4 * q);
          const float2 lo = __fadd2_rn(make_float2(v.x, v.y), ndc), hi = __fadd2_rn(make_float2(v.z, v.w), ndc);
          stg_stream4(ys + own_lo + 4 * (tid + k * LM_THREADS), make_float4(lo.x, lo.y, hi.x, hi.y));
          ss = fmaf(lo.x, lo.x, ss); ss = fmaf(lo.y, lo.y, ss); ss = fmaf(hi.x, hi.x, ss); ss = fmaf(hi.y, hi.y, ss);
        }
      } else if (tid >= FZ_FIR_MAIN) {
        // The two apply warps are what the FIR barrier waits for (probe: without the decay sums the kernel is 5 % faster),
        // so the sums are packed (two partial sums on the fp32x2 pipe: half the instructions) and skipped altogether in
        // the middle third of the clip, where nobody wants them.
        constexpr int AT = LM_THREADS - FZ_FIR_MAIN;         // 64 apply threads
        static_assert((FZ_OWN / 4) % AT == 0, "apply pieces per thread");
        if (in_first || in_last) {
          float2 ss2 = make_float2(0.f, 0.f);
#pragma unroll 6
          for (int k = 0; k < FZ_OWN / 4 / AT; ++k) {
            const int i = (tid - FZ_FIR_MAIN) + k * AT;
            const float4 v = *reinterpret_cast<const float4*>(H.span + FZ_LEAD + 4 * i);
            const float2 lo = __fadd2_rn(make_float2(v.x, v.y), ndc), hi = __fadd2_rn(make_float2(v.z, v.w), ndc);
            stg_stream4(ys + own_lo + 4 * i, make_float4(lo.x, lo.y, hi.x, hi.y));
            ss2 = __ffma2_rn(lo, lo, ss2); ss2 = __ffma2_rn(hi, hi, ss2);
          }
          ss = ss2.x + ss2.y;
        } else {
#pragma unroll 6
          for (int k = 0; k < FZ_OWN / 4 / AT; ++k) {
            const int i = (tid - FZ_FIR_MAIN) + k * AT;
            const float4 v = *reinterpret_cast<const float4*>(H.span + FZ_LEAD + 4 * i);
            const float2 lo = __fadd2_rn(make_float2(v.x, v.y), ndc), hi = __fadd2_rn(make_float2(v.z, v.w), ndc);
            stg_stream4(ys + own_lo + 4 * i, make_float4(lo.x, lo.y, hi.x, hi.y));
          }
        }
      }
      if (in_first) a_first += ss;
      if (in_last) a_last += ss;
    } else {
      // ---- apply, edge batch: span <- (x - dc) * fade; the batch's own samples go to HBM; decay sums
      for (int q = tid; q < FZ_SPAN / 4; q += LM_THREADS) {
        const int o = j0 + 4 * q;
        if (o + 3 < 0 || o >= n) continue;                    // zero-filled: stays zero
        float4 v = *reinterpret_cast<float4*>(H.span + 4 * q);
        const bool owned = o >= own_lo && o < own_hi;         // own_lo, own_hi and o are multiples of 4
        const bool p_first = o + 3 < third, p_last = o >= n - third;
        const bool clean = o >= edge_lo && o + 3 < edge_hi && (p_first || o >= third) && (p_last || o + 3 < n - third);
        if (clean) {
          v.x = __fsub_rn(v.x, dc); v.y = __fsub_rn(v.y, dc); v.z = __fsub_rn(v.z, dc); v.w = __fsub_rn(v.w, dc);
          if (owned) {
            stg_stream4(ys + o, v);
            const float ss = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
            if (p_first) a_first += ss;
            if (p_last) a_last += ss;
          }
        } else {
          float e[4] = {v.x, v.y, v.z, v.w};
          fz_apply_edge(e, o, n, fade, need_fade, dc, third, owned, ys, &a_first, &a_last);
          v = make_float4(e[0], e[1], e[2], e[3]);
        }
        *reinterpret_cast<float4*>(H.span + 4 * q) = v;
      }
      half_sync(half);
    }
    if (t0 >= T_real) {                                      // a batch that only had output samples left to write
      if (FZ_DEFER_TILE_END && FZ_INLINE_NORM && b == 0 && tid == 0) resolve_pending();
      if (next) { half_sync(half); stage_span(t0 + LM_BF); }
      else {
        half_sync(half);
        if (tid == 0) { if (b == 0) fill_desc(H.next_unit); prefetch_next_tile(); }
      }
      if (FZ_DEFER_TILE_END && FZ_INLINE_NORM && b == 0 && H.fill_c >= 0) do_fill(H.fill_c, H.pend_T_real, H.pend_T, H.fill_max);
      continue;
    }
    // ---- FIR 24k -> 16k: slab position i holds w16[w0 + i], w0 = 160*t0 - 200;
    //      w16[2m+p] = sum_t y[3m - 10 + t] * k[p][t]  ->  span offset 2 + 3*(i/2) + t
    const int w0 = HOP16 * t0 - N_FFT / 2;
    float* slab = H.pw;
    {
      // interior batches read the raw x: start the accumulators at -dc * sum(taps)
      const float i0 = fast ? -dc * c_fsum[0] : 0.f, i1 = fast ? -dc * c_fsum[1] : 0.f;
#if FZ_FFMA2_FIR
#if FZ_FIR_QPT > 1
      // FZ_FIR_QPT consecutive quads (4 outputs each) per thread from ONE window of the span held in registers: a quad on
      // its own reads 26 floats for 4 outputs, five in a row read 48 for 20 -- the FIR's loads were 22 % of the kernel's
      // shared-memory traffic.  1340 quads = 268 threads x 5: one balanced pass (the strided loop gave the first two warps
      // five rounds and the others four); in split mode 256 threads x 5 + 60 x 1.
      static_assert(FZ_DPAIRS % FZ_FIR_QPT == 0, "quads per thread must divide the quads per batch");
      auto fir_quads = [&](auto qpt_tag, int d0) {
        constexpr int QPT = decltype(qpt_tag)::value;
        const float2* sp = reinterpret_cast<const float2*>(H.span + 2 + 6 * d0);
        float2 V[3 * QPT + 9];
#pragma unroll
        for (int k = 0; k < 3 * QPT + 9; ++k) V[k] = sp[k];
#pragma unroll
        for (int j = 0; j < QPT; ++j) {
          const int d = d0 + j;
          float2 o00 = make_float2(i0, 0.f), o10 = o00, o01 = make_float2(i1, 0.f), o11 = o01;
#ifdef FZ_PROBE_HALF_FIR        // timing probe only (WRONG results): every other tap pair skipped
#define FZ_KSTEP 2
#else
#define FZ_KSTEP 1
#endif
#pragma unroll
          for (int k = 0; k < 10; k += FZ_KSTEP) o00 = __ffma2_rn(V[3 * j + k], c_fA0[k], o00);
#pragma unroll
          for (int k = 2; k < 12; k += FZ_KSTEP) o10 = __ffma2_rn(V[3 * j + k], c_fB0[k], o10);
#pragma unroll
          for (int k = 1; k < 11; k += FZ_KSTEP) o01 = __ffma2_rn(V[3 * j + k], c_fA1[k], o01);
#pragma unroll
          for (int k = 3; k < 12; k += FZ_KSTEP) o11 = __ffma2_rn(V[3 * j + k], c_fB1[k], o11);
          const int wi = w0 + 4 * d;
          float4 r;
          r.x = (wi + 0 < n_valid) ? o00.x + o00.y : 0.f;
          r.y = (wi + 1 < n_valid) ? o01.x + o01.y : 0.f;
          r.z = (wi + 2 < n_valid) ? o10.x + o10.y : 0.f;
          r.w = (wi + 3 < n_valid) ? o11.x + o11.y : 0.f;
          *reinterpret_cast<float4*>(slab + 4 * d + 20 * (d / (LM_SLAB_BLK / 4))) = r;
        }
      };
      // one call site for the five-quad pass (the code is 450 instructions): split batches give it the first 256 threads
      // and the last 60 quads to warps 8-9 one at a time, the others the first 268 threads
      if (tid < (split ? FZ_FIR_MAIN : FZ_DPAIRS / FZ_FIR_QPT)) fir_quads(std::integral_constant<int, FZ_FIR_QPT>{}, FZ_FIR_QPT * tid);
      else if (split && FZ_FIR_MAIN * FZ_FIR_QPT + (tid - FZ_FIR_MAIN) < FZ_DPAIRS)
        fir_quads(std::integral_constant<int, 1>{}, FZ_FIR_MAIN * FZ_FIR_QPT + (tid - FZ_FIR_MAIN));
#else
      for (int d = tid; d < FZ_DPAIRS; d += LM_THREADS) {
        const float2* sp = reinterpret_cast<const float2*>(H.span + 2 + 6 * d);
        float2 V[13];
#pragma unroll
        for (int k = 0; k < 13; ++k) V[k] = sp[k];
        float2 o00 = make_float2(i0, 0.f), o10 = o00, o01 = make_float2(i1, 0.f), o11 = o01;
#pragma unroll
        for (int k = 0; k < 10; ++k) o00 = __ffma2_rn(V[k], c_fA0[k], o00);
#pragma unroll
        for (int k = 2; k < 12; ++k) o10 = __ffma2_rn(V[k], c_fB0[k], o10);
#pragma unroll
        for (int k = 1; k < 11; ++k) o01 = __ffma2_rn(V[k], c_fA1[k], o01);
#pragma unroll
        for (int k = 3; k < 12; ++k) o11 = __ffma2_rn(V[k], c_fB1[k], o11);
        const int wi = w0 + 4 * d;
        float4 r;
        r.x = (wi + 0 < n_valid) ? o00.x + o00.y : 0.f;
        r.y = (wi + 1 < n_valid) ? o01.x + o01.y : 0.f;
        r.z = (wi + 2 < n_valid) ? o10.x + o10.y : 0.f;
        r.w = (wi + 3 < n_valid) ? o11.x + o11.y : 0.f;
        *reinterpret_cast<float4*>(slab + 4 * d + 20 * (d / (LM_SLAB_BLK / 4))) = r;
      }
#endif
#else
      for (int d = tid; d < FZ_DPAIRS; d += LM_THREADS) {
        const float2* sp = reinterpret_cast<const float2*>(H.span + 2 + 6 * d);
        float v[26];
#pragma unroll
        for (int k = 0; k < 13; ++k) { const float2 t2 = sp[k]; v[2 * k] = t2.x; v[2 * k + 1] = t2.y; }
        float o00 = i0, o01 = i1, o10 = i0, o11 = i1;
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
#endif
    }
#if FZ_TC_MEL
    // the previous batch's mel projection has had a whole FIR to finish: drain it.  The waiting warps also
    // make sure the tensor core is done reading the fb region before stage 1 below overwrites it.
    if (pending) {
      if (epi_warp) fz_tc_epilogue<NM>(&H.mma_bar, mma_parity, d_tmem, tid, pend_t0, T_real, out, mel_stride, &lmax);
      mma_parity ^= 1u;
      pending = false;
    }
#endif
    half_sync(half);
    if (FZ_DEFER_TILE_END && FZ_INLINE_NORM && b == 0 && tid == 0) resolve_pending();   // the previous tile end is a FIR old
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
      const float4* __restrict__ tq = reinterpret_cast<const float4*>(S.twT + FZ_TWS * lane);
      float2* __restrict__ fbw = fb;
      fbw[lane] = v[0];
#pragma unroll
      for (int q5 = 0; q5 < 10; q5 += 5) {
        float4 t4[5];                                         // twiddles of k1 = 2q, 2q+1: five loads ahead of the stores
#pragma unroll
        for (int q = 0; q < 5; ++q) t4[q] = tq[q5 + q];
#pragma unroll
        for (int qq = 0; qq < 5; ++qq) {
          const int q = q5 + qq;
          if (q > 0) fbw[(2 * q) * 21 + lane] = cmul(v[2 * q], make_float2(t4[qq].x, t4[qq].y));
          fbw[(2 * q + 1) * 21 + lane] = cmul(v[2 * q + 1], make_float2(t4[qq].z, t4[qq].w));
        }
      }
    }
    half_sync(half);
    if (next) stage_span(t0 + LM_BF);                        // span and slab are dead: prefetch under stage 2 / mel
    else if (tid == 0) { if (b == 0) fill_desc(H.next_unit); prefetch_next_tile(); }
    // the padding constant of the clip this half finished last a tile ago (decided by thread 0 after the FIR barrier)
    if (FZ_DEFER_TILE_END && FZ_INLINE_NORM && b == 0 && H.fill_c >= 0) do_fill(H.fill_c, H.pend_T_real, H.pend_T, H.fill_max);
    // ---- FFT stage 2: lane = k1, 20-point DFT over n2 -> Z[k1 + 20*k2] in v[k2]
#pragma unroll
    for (int n2 = 0; n2 < 20; ++n2) v[n2] = fb[lane * 21 + n2];
    dft20(v);
    // ---- split the two real spectra and take |.|^2:  A = (Z[k]+conj Z[400-k])/2, B = (Z[k]-conj Z[400-k])/(2i).
    // Bin k = k1 + 20*k2 <= 200 needs Z[400-k] = Z[(20-k1) + 20*(19-k2)]: the upper half (k2 >= 10) of lane
    // 20-k1.  Every lane publishes its upper half, then reads its partner's; lane 0 is its own partner.
    // The exchange goes through the pw region (the slab in it died at the stage-1 barrier) and the power
    // spectra go to the fb region (dead once everybody is past the barrier below): no barrier is needed between
    // stage 2 and the publication, nor between the mel projection and the next batch's FIR.
    // Lane 0 pairs with itself, one k2 later (Z[400 - 20 k2] = Z[20 (20 - k2)]): with part = pub + 20 its reads land on
    // its own published column, and on pub[200] = Z[400] = Z[0] for k2 = 0, which it publishes as well.
    float2* __restrict__ pub = reinterpret_cast<float2*>(H.pw) + g * 201;
#pragma unroll
    for (int k2 = 10; k2 < 20; ++k2) pub[lane + 20 * (k2 - 10)] = v[k2];
    if (lane == 0) pub[200] = v[0];
    half_sync(half);
#if FZ_TC_MEL
    {
      // power spectra as tensor-core operands: hi (top 19 bits) and lo (remainder), K-major, 64-byte swizzle
      unsigned char* opb = reinterpret_cast<unsigned char*>(H.fb);
      const int odd = g & 1;                                 // odd groups store frame B first: the two groups a warp
      const int f0 = 2 * g + odd, f1 = 2 * g + 1 - odd;      // straddles then hit different banks
      const float2* part = pub + (20 - lane);                // lane 0: reads stay inside the pw region (unused)
      auto put = [&](int f, int k, float pv) {
        const uint32_t o = tc_off(f, k);
#if FZ_TC_MEL == 2
        // two products (Fhi + Flo) * P with P rounded to nearest TF32: half the operand traffic, 2^-12 per term
        *reinterpret_cast<float*>(opb + o) = __uint_as_float((__float_as_uint(pv) + 0x1000u) & 0xffffe000u);
#else
        const float h = tc_hi(pv);
        *reinterpret_cast<float*>(opb + o) = h;
        *reinterpret_cast<float*>(opb + TC_OPERAND_BYTES + o) = pv - h;
#endif
      };
#pragma unroll
      for (int k2 = 0; k2 < 10; ++k2) {
        float2 w = part[20 * (9 - k2)];                      // partner's k2' = 19 - k2, stored at 20 * (k2' - 10)
        if (lane == 0) w = (k2 == 0) ? v[0] : v[20 - k2];
        const float2 z = v[k2];
        const float ar = z.x + w.x, ai = z.y - w.y, br = z.x - w.x, bi = z.y + w.y;
        const float pA = 0.25f * (ar * ar + ai * ai), pB = 0.25f * (br * br + bi * bi);
        const int k = lane + 20 * k2;
        put(f0, k, odd ? pB : pA);
        put(f1, k, odd ? pA : pB);
      }
      if (lane == 0) {                                       // k = 200: Z[200] pairs with itself
        const float2 z = v[10];
        put(2 * g, 200, z.x * z.x);
        put(2 * g + 1, 200, z.y * z.y);
      } else if (lane < 8) {                                 // bins 201..207 pad K to 208: must be finite (x 0)
        put(2 * g, 200 + lane, 0.f);
        put(2 * g + 1, 200 + lane, 0.f);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> the tensor core's async reads
    }
    half_sync(half);
    if (tid < 32) {
      tc_fence_after();
      if (tc_elect_one()) {
        const uint32_t b_hi = (smem_u32(H.fb) >> 4) & 0x3FFFu, b_lo = ((smem_u32(H.fb) + TC_OPERAND_BYTES) >> 4) & 0x3FFFu;
        constexpr uint32_t DLO = 1u << 16;                                   // leading byte offset field (unused: 1)
        constexpr uint32_t DHI = (512u >> 4) | (1u << 14) | (4u << 29);      // SBO 512 B, version 1, SWIZZLE_64B
#pragma unroll
        for (int ks = 0; ks < TC_KSTEPS; ++ks) {
          const uint32_t off = (uint32_t)(((ks >> 1) * 2048 + (ks & 1) * 32) >> 4);
          tc_mma(d_tmem, tmem + TC_COL_WHI + 8 * ks, (b_hi + off) | DLO, DHI, ks > 0);
          tc_mma(d_tmem, tmem + TC_COL_WLO + 8 * ks, (b_hi + off) | DLO, DHI, 1);
#if FZ_TC_MEL != 2
          tc_mma(d_tmem, tmem + TC_COL_WHI + 8 * ks, (b_lo + off) | DLO, DHI, 1);
#endif
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&H.mma_bar)) : "memory");
      }
      __syncwarp();
    }
    pending = true;
    pend_t0 = t0;
    // (the accumulator is drained after the next batch's FIR, or at the end of the tile)
#else
    float* power = reinterpret_cast<float*>(H.fb);
    {
      float* __restrict__ pa = power + (2 * g) * LM_PS + lane;
      float* __restrict__ pb = pa + LM_PS;
      const float2* __restrict__ part = pub + (20 - lane);
      float2 wp[10];                                         // all ten partner reads in flight before the first use
#pragma unroll
      for (int k2 = 0; k2 < 10; ++k2) wp[k2] = part[20 * (9 - k2)];   // partner's k2' = 19 - k2, stored at 20 * (k2' - 10)
      // 4 |A|^2 and 4 |B|^2 (the factor 1/4 of A = (Z + conj Z')/2 is applied once per mel row, where it is exact:
      // scaling by a power of two commutes with every rounding of the projection).  Packed: S = Z + W = (ar, bi),
      // D = Z - W = (br, ai); |A|^2 = ar^2 + ai^2 = S.x^2 + D.y^2, |B|^2 = br^2 + bi^2 = D.x^2 + S.y^2.
#pragma unroll
      for (int k2 = 0; k2 < 10; ++k2) {
        const float2 sm = cadd(v[k2], wp[k2]), df = csub(v[k2], wp[k2]);
        const float2 s2 = __fmul2_rn(sm, sm), d2 = __fmul2_rn(df, df);
        pa[20 * k2] = s2.x + d2.y;
        pb[20 * k2] = d2.x + s2.y;
      }
      if (lane == 0) {                                       // k = 200: Z[200] pairs with itself
        const float2 z = v[10];
        pa[200] = 4.f * (z.x * z.x);
        pb[200] = 4.f * (z.y * z.y);
      }
    }
    half_sync(half);
    // ---- mel projection + log10
    {
      const int f = tid & 31, part = tid >> 5;
      const int t = t0 + f;
      const bool live = t < T_real;
      const float* p = power + f * LM_PS;
      float* __restrict__ o = out + t;
      auto emit = [&](int m, float acc) {
        const float ls = __log2f(fmaxf(0.25f * acc, 1e-10f)) * 0.30102999566398120f;   // the spectra are 4 |.|^2
        if (live) {
          o[(long long)m * mel_stride] = lm_scaled(ls);
          lmax = fmaxf(lmax, ls);
        }
      };
#if FZ_MEL_TABLE
      constexpr int which = NM == 80 ? 0 : 1;
      const int m1 = c_mel_part[which][part + 1];
      int m = c_mel_part[which][part];
      float* __restrict__ orow = o + (long long)m * mel_stride;
      unsigned pbase = smem_u32(p);
      asm volatile("" : "+r"(pbase));                        // keep the row address in a register (ptxas re-derives it per row)
      const float4* __restrict__ tp = c_mel_tab[which] + c_mel_part4[which][part];
      constexpr int R = fz_mel_rows(which);
      for (; m < m1; m += R) {
        const unsigned* __restrict__ hd = reinterpret_cast<const unsigned*>(tp);   // {first bin of each row (bytes)}, {bytes}
        unsigned pa[R];
#pragma unroll
        for (int r = 0; r < R; ++r) pa[r] = pbase + hd[r];
        const unsigned pe = pa[0] + hd[R];
        tp += fz_mel_hdr4(which);
        float acc[R];                                        // per row the same left-to-right fmaf chain as the unrolled form
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = 0.f;
#pragma unroll 1
        for (; pa[0] != pe; tp += R) {
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const float4 w = tp[r];
            float p0, p1, p2, p3;
            asm volatile("ld.shared.f32 %0, [%4];\n\tld.shared.f32 %1, [%4+4];\n\tld.shared.f32 %2, [%4+8];\n\tld.shared.f32 %3, [%4+12];"
                         : "=f"(p0), "=f"(p1), "=f"(p2), "=f"(p3) : "r"(pa[r]));
            acc[r] = fmaf(w.x, p0, acc[r]); acc[r] = fmaf(w.y, p1, acc[r]);
            acc[r] = fmaf(w.z, p2, acc[r]); acc[r] = fmaf(w.w, p3, acc[r]);
            pa[r] += 16u;
          }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
          // log2 of a value >= 1e-10: the flush-to-zero form gives the same bits without the denormal pre-scaling
          float l2;
          asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(fmaxf(acc[r], 1e-10f)));   // the 1/4 is in the table
          const float ls = l2 * 0.30102999566398120f;
          if (live && m + r < m1) {
            *orow = lm_scaled(ls);
            lmax = fmaxf(lmax, ls);
          }
          orow += mel_stride;
        }
      }
#else
      if (NM == 80) mel_sparse_80(part, p, emit); else mel_sparse_128(part, p, emit);
#endif
    }
#endif
    // the next batch's barrier after its FIR orders these reads before its stage 1 overwrites the fb region
  }
#if FZ_TC_MEL
  if (pending) {                                     // the last batch of the tile
    if (epi_warp) fz_tc_epilogue<NM>(&H.mma_bar, mma_parity, d_tmem, tid, pend_t0, T_real, out, mel_stride, &lmax);
    mma_parity ^= 1u;
    pending = false;
  }
#endif
  // ---- per-half reductions: clip max (ordered-int atomicMax), decay sums (double atomics)
  lmax = warp_max(lmax);
  const double df = warp_sum((double)a_first), dl = warp_sum((double)a_last);
  if (!FZ_DEFER_TILE_END) half_sync(half);           // (deferred: warp 0 read these right after the barrier below, a tile ago)
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
      // The half that finishes its clip last knows the clip maximum: it writes the constant that fills the
      // frames which only see zero padding (2/3 of a 10 s clip's features).  Pure stores, nothing to wait
      // for; the frames with signal are clamped / scaled by k_logmel_norm, which then moves 1/3 of the bytes.
      // (Normalising them here too was measured and is NOT a gain: the loads expose L2 latency on a half that
      //  has nothing else to do, profiles/README.md.)
      // Ordering: the count is incremented with an acquire-release atomic at GPU scope.  Its release half orders this
      // half's atomicMax (and decay sums) before the increment; its acquire half makes the maxima of the halves that
      // incremented earlier visible to whoever sees the full count, so the last finisher's read of the clip maximum
      // (which decides the constant of 2/3 of the clip's features) is ordered by the PTX memory model, not by where
      // the L2 happens to resolve atomics.  (Round 1 relied on a data dependency between relaxed atomics: 1.7 % faster,
      // not guaranteed.)
      int last = 0;
      if (FZ_INLINE_NORM && tiles_done) {
        int clip_tiles = 0;                          // tiles of the schedule that start inside this clip
        while (clip_tiles < sched.n_tiles && sched.start[clip_tiles] * LM_BF < t_cover) ++clip_tiles;
        if (FZ_DEFER_TILE_END) {
          // issued, not waited for: every tile runs resolve_pending() in its first batch, so no older count is in flight
          asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], 1;" : "=r"(pend_before) : "l"(tiles_done + c) : "memory");
          H.pend_c = c; H.pend_tiles = clip_tiles; H.pend_T_real = T_real; H.pend_T = min(T, fill_to);
          pend = min(T, fill_to) > T_real;           // nothing to fill: nobody needs the answer
        } else {
          int before;
          asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], 1;" : "=r"(before) : "l"(tiles_done + c) : "memory");
          last = (before == clip_tiles - 1);
          if (last) H.fill_max = atomicMax(&clip_max[c], (int)0x80000000);
        }
      }
      H.last = last;
    }
  }
  if (!FZ_DEFER_TILE_END) {
    half_sync(half);
    if (H.last && min(T, fill_to) > T_real) do_fill(c, T_real, min(T, fill_to), H.fill_max);
  }
  }   // tiles of this half
  if (FZ_DEFER_TILE_END && FZ_INLINE_NORM) {         // the count of this half's last tile
    if (tid == 0) resolve_pending();
    half_sync(half);
    if (H.fill_c >= 0) do_fill(H.fill_c, H.pend_T_real, H.pend_T, H.fill_max);
  }
#if FZ_TC_MEL
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(TC_TMEM_COLS) : "memory");
  }
#endif
}

cudaError_t launch_fused_features(const Tables& tb, const float* x, const int64_t* seg_off, const Workspace& ws,
                                  const int32_t* item_first_seg, int n_items, int64_t max_len,
                                  const Derived& d, float* y, const int64_t* y_off, int n_mels, int pad_frames,
                                  float* mel, int64_t mel_stride_frames, int sm_count, cudaStream_t st, LaunchCtx* lc,
                                  int mode, int fill_to) {
  const bool from_y = mode == 1;
  if (n_items <= 0) return cudaSuccess;
  if (fill_to <= 0) fill_to = pad_frames;
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
  const int64_t cover = from_y ? max_real : std::max<int64_t>(max_real, (max_len + 239) / 240);
  cudaError_t e;                                     // sm_count: of the handle's device (a process may drive several)
  // Unit = (clip, tile).  Every tile ends with a reduction, three atomics and a fence and starts with an empty
  // pipeline (~3 us together), so tiles should be long -- but the halves draw units from one counter, and whoever draws
  // a long tile last finishes last.  Measured on 1000 x 10 s (32 batches per clip), k_fused_features in ms:
  //   uniform tiles of 2 / 4 / 8 / 11 / 13 / 16 batches   1.019 / 0.961 / 0.933 / 0.934 / 0.918 / 0.936
  //   13,13,6  14,14,4  16,13,3  15,15,2  22,8,2  30,2     0.919   0.925   0.928   0.944   0.928  1.027
  //   guided (each tile 40 % of the rest: 13,8,5,3,2,1)    0.952   (twice the tiles: the per-tile cost wins)
  // Policy: tiles of 40 % of the longest clip (two long tiles and a half-size one to level the end), capped with few
  // clips so that there are a few units per half.  RHO_FUSED_SCHED / RHO_FUSED_TILE_BATCHES override it for A/B runs.
  const int64_t n_b = std::max<int64_t>(1, (cover + LM_BF - 1) / LM_BF);
  static const int forced_bpt = [] { const char* v = getenv("RHO_FUSED_TILE_BATCHES"); return v ? atoi(v) : 0; }();   // A/B tool: uniform tiles
  const int64_t halves = (int64_t)FZ_HALVES * sm_count;
  int64_t cap = std::max<int64_t>(1, (n_b * n_items) / (4 * halves));      // >= ~4 units per half before the tail
  FzSched sched;
  sched.n_tiles = 0;
  sched.start[0] = 0;
  static const std::vector<int> forced_sched = [] {                        // A/B tool: explicit sizes, last one repeats
    std::vector<int> v;
    if (const char* e = getenv("RHO_FUSED_SCHED")) for (const char* q = e; *q;) { v.push_back(atoi(q)); while (*q && *q != ',') ++q; if (*q) ++q; }
    return v;
  }();
  for (int64_t pos = 0; pos < n_b;) {
    int64_t left = n_b - pos, sz;
    if (!forced_sched.empty()) sz = std::max(1, forced_sched[std::min<size_t>(sched.n_tiles, forced_sched.size() - 1)]);
    else if (forced_bpt > 0) sz = forced_bpt;
    else {
      // ... and of at most 32 batches: with items of very different lengths (joined items of 2..6 ragged segments: the
      // longest is three times the average) a tile of 40 % of the LONGEST item swallows most items whole and the halves
      // end far apart -- C3, join + features, k_fused_features: 7.35 ms with 300-batch tiles, 6.59 / 6.66 / 6.92 ms with
      // uniform 26 / 52 / 104, later 6.40 / 6.36 / 6.33 with 20 / 26 / 32 (profiles/ab_r02_join_tiles.log); the table
      // holds 64 tiles, longer items get longer ones
      sz = std::min<int64_t>(std::min<int64_t>(std::max<int64_t>((n_b * 2 + 4) / 5, 1), cap), 32);
      sz = std::max<int64_t>(sz, (n_b + FZ_MAX_TILES - 2) / (FZ_MAX_TILES - 1));
    }
    if (sched.n_tiles == FZ_MAX_TILES - 1) sz = left;                      // table full: the rest in one tile
    sz = std::min(sz, left);
    pos += sz;
    sched.start[++sched.n_tiles] = (int)pos;
  }
  const int64_t tiles = sched.n_tiles;
  const size_t smem = sizeof(FzSmem) + 1024;      // slack to align the halves to 1024 bytes (swizzle atoms)
  auto kern = mode == 1 ? ((n_mels == 80) ? k_fused_features<80, 1> : k_fused_features<128, 1>)
            : mode == 2 ? ((n_mels == 80) ? k_fused_features<80, 2> : k_fused_features<128, 2>)
                        : ((n_mels == 80) ? k_fused_features<80, 0> : k_fused_features<128, 0>);
  if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
  const uint64_t n_units = (uint64_t)n_items * (uint64_t)tiles;
  if (n_units > 0x7fffffffull) return cudaErrorInvalidValue;
  const uint64_t n_work = (n_units + FZ_HALVES - 1) / FZ_HALVES;
  dim3 grid((unsigned)(n_work < (uint64_t)sm_count ? n_work : (uint64_t)sm_count));   // persistent: one CTA per SM
  lc->begin(KID_FUSED, st);
  kern<<<grid, FZ_THREADS, smem, st>>>(x, seg_off, ws.seg, ws.item, item_first_seg, y, y_off, d.fade, tb.hann,
                                       tb.twiddle, pad_frames, mel, mel_stride_frames, ws.clip_max, ws.len16,
                                       ws.tiles_done, ws.work_counter, sched, n_items,
                                       tb.mel_dense[n_mels == 80 ? 0 : 1], fill_to, ws.span);
  lc->end(st);
  return cudaGetLastError();
}

}  // namespace rho
