// Internal launcher declarations shared by the translation units of librho_b200.
#pragma once
#include <atomic>
#include <vector>
#include <nvtx3/nvToolsExt.h>
#include "common.cuh"

namespace rho {

#ifndef RHO_SCAN_FR
#define RHO_SCAN_FR 32               // measured on B200 (C2 scan, dense rows): 32 -> 0.165 ms, 64 -> 0.170, 96 -> 0.184, 128 -> 0.194
#endif
#ifndef RHO_GATHER_CHUNKS
#define RHO_GATHER_CHUNKS 1          // passes per gather CTA
#endif
constexpr int SCAN_FR = RHO_SCAN_FR;  // energy frames (= threads) per scan CTA
constexpr int GATHER_THREADS = 256;
#ifndef RHO_GATHER_CHUNK
#define RHO_GATHER_CHUNK 8192
#endif
constexpr int GATHER_CHUNK = RHO_GATHER_CHUNK;   // output samples per pass of a gather CTA (8 x 128-bit pieces per thread in flight)
constexpr int GATHER_TILE = GATHER_CHUNK * RHO_GATHER_CHUNKS;   // output samples per gather CTA

constexpr int RS_TAPS = 23;           // 24k->16k polyphase kernel length (width 10, orig 3)
constexpr int N_FFT = 400;
constexpr int HOP16 = 160;
constexpr int N_BINS = 201;
constexpr int MEL_PAD_FRAMES = 3000;
constexpr int MAX_MELS = 128;

// Device-resident constant tables owned by the handle.
struct Tables {
  float* hann;        // [400]
  float2* twiddle;    // [400]  exp(-2*pi*i*k/400)
  // sparse mel filterbanks (slaney): for mel m the non-zero bins are [lo[m], lo[m]+cnt[m]) with
  // weights w[wofs[m] ...]
  int* mel_lo[2];     // index 0: 80 mels, 1: 128 mels
  int* mel_cnt[2];
  int* mel_wofs[2];
  float* mel_w[2];
  int mel_nnz[2];
  float* mel_dense[2];  // [n_mels][201] row-major, fp32 (for the tensor-core path)
};

// Launch bookkeeping: counts kernel launches and, when profiling is on, brackets every kernel with
// CUDA events on the launching stream so per-kernel device time can be read back
// (rho_b200_profile_*; bench.py uses it for the roofline line).
enum KernelId {
  KID_INIT = 0, KID_SCAN, KID_FINALIZE_SEGS, KID_PLAN, KID_GATHER, KID_FINALIZE_ITEMS,
  KID_RESAMPLE, KID_LOGMEL_INIT, KID_LOGMEL_FRAMES, KID_LOGMEL_NORM, KID_COSINE, KID_SINGLE, KID_FUSED, KID_MEL_GEMM, KID_QWEN_MOMENTS, KID_QWEN_PLAN, KID_QWEN_APPLY, KID_RESAMPLE_GENERAL,
  KID_PV_STFT, KID_PV_PHASE, KID_PV_CUMSUM, KID_PV_ISTFT, KID_PV_RESAMPLE, KID_MFCC_FRAMES, KID_MFCC_STATS, KID_XCH_WAIT, KID_STFT_TC, KID_SPK_SUMSQ, KID_SPK_SCALE, KID_SPK_MEL, KID_SPK_POOL, KID_COUNT
};
extern const char* const kKernelNames[KID_COUNT];

struct LaunchCtx {
  std::atomic<int64_t> launches{0};   // a handle may be shared by concurrent sessions (ui/state.py:85-87 of the reference)
  bool profiling = false;
  struct Span { int id; cudaEvent_t a, b; };
  std::vector<Span> spans;
  void begin(int id, cudaStream_t st) {
    nvtxRangePushA(kKernelNames[id]);       // closed by end(): one NVTX range per kernel launch, nested in the entry point's
    if (!profiling) return;
    Span sp; sp.id = id;
    cudaEventCreate(&sp.a); cudaEventCreate(&sp.b);
    cudaEventRecord(sp.a, st);
    spans.push_back(sp);
  }
  void end(cudaStream_t st) {
    nvtxRangePop();
    ++launches;
    if (profiling && !spans.empty()) cudaEventRecord(spans.back().b, st);
  }
};

// join.cu
enum { JOIN_PREPARE = 1, JOIN_GATHER = 2, JOIN_FINISH = 4, JOIN_ALL = 7,     // stages of launch_join
       JOIN_INIT_FEATURES = 8,     // with JOIN_PREPARE: also reset the log-mel clip maxima / tile counters
       JOIN_ONE_SEG_ITEMS = 16 };  // with JOIN_PREPARE: item i is segment i -- k_finalize_segs plans it, no k_plan_items
cudaError_t launch_trim_scan(const float* x, const int64_t* off, const int32_t* len, const uint8_t* trim_flags,
                             int n_seg, int64_t max_len, const Derived& d, const Workspace& ws,
                             rho_seg_info* info, cudaStream_t st, LaunchCtx* lc);
cudaError_t launch_join(const float* x, const int64_t* seg_off, const int32_t* seg_len, int n_seg, int64_t max_seg_len,
                        const int32_t* item_first_seg, int n_items, int64_t max_item_len,
                        const Derived& d, float* y, const int64_t* y_off, rho_record* rec, rho_seg_info* seg_info,
                        const Workspace& ws, cudaStream_t st, LaunchCtx* lc, int stages = JOIN_ALL,
                        const float* emb = nullptr, const float* ref_emb = nullptr, int emb_dim = 0);
cudaError_t launch_remove_dc(float* x, int64_t n, float* dc_out, double* scratch, cudaStream_t st, LaunchCtx* lc);
cudaError_t launch_apply_fades(float* x, int64_t n, int fade, int fade_in, int fade_out, cudaStream_t st, LaunchCtx* lc);
cudaError_t launch_sound_decay(const float* x, int64_t n, double thr, rho_record* rec, double* scratch,
                               cudaStream_t st, LaunchCtx* lc);
cudaError_t launch_sound_decay_batch(const float* y, const int64_t* off, const int32_t* len, int len_stride_bytes,
                                     int n, int64_t max_len, double thr, rho_record* rec, double* sums,
                                     cudaStream_t st, LaunchCtx* lc);

cudaError_t launch_pcm16(const float* y, const int64_t* off, const int32_t* len, int len_stride_bytes, int n,
                         int64_t max_len, int16_t* out, const int64_t* out_off, cudaStream_t st, LaunchCtx* lc);

// resample.cu
cudaError_t upload_resample_taps(const float* taps /* [2][23] */);
cudaError_t launch_resample3to2(const float* x, const int64_t* off, const int32_t* len, int len_stride_bytes,
                                int n, int64_t max_len, float* y, const int64_t* y_off, int32_t* y_len,
                                cudaStream_t st, LaunchCtx* lc);

cudaError_t launch_resample_general(const float* x, const int64_t* off, const int32_t* len, int len_stride_bytes,
                                    int n, int64_t max_len, int orig, int nw, int width, const float* taps,
                                    float* y, const int64_t* y_off, int32_t* y_len, cudaStream_t st, LaunchCtx* lc);

// logmel.cu
cudaError_t launch_logmel(const Tables& tb, const float* x16, const int64_t* off, const int32_t* len16,
                          int n, int64_t max_len16, int n_mels, int pad_frames, float* mel,
                          int64_t mel_stride_frames, int32_t* n_frames, int* clip_max,
                          cudaStream_t st, LaunchCtx* lc, int fill_to = 0, float* pad_value = nullptr);

cudaError_t launch_logmel_init(int* clip_max, int n, cudaStream_t st, LaunchCtx* lc, int* tiles_done = nullptr);
// what k_logmel_norm needs to assemble the records itself (fused path: saves the k_finalize_items launch)
// The record "gather" of the multi-GPU path, fused into the kernel that assembles the records (no collective call on
// the critical path; exchange.cu, rho_b200_exchange_*).  Besides rec[it], the record of item `it` is stored to
// sink[q][slot + it] for every rank q < n: sink[q] is rank q's gathered-record buffer mapped into this process (CUDA IPC
// over NVLink / NVSwitch peer memory; sink[own rank] is the local buffer).  The CTA that finishes last publishes
// flag[q][rank] = epoch with a system-scope release after a system-scope fence: rank q's readers wait for that value.
constexpr int MAX_RECORD_PEERS = 16;
struct RecordPeers {
  int n;                                 // 0: no peer stores
  int rank;
  long long slot;                        // index of this rank's first record in every gathered buffer (this epoch's parity)
  unsigned epoch;
  int n_items;                           // records this call produces (the last of them publishes the flags)
  int* done;                             // local counter of records stored (reset by the publisher)
  rho_record* sink[MAX_RECORD_PEERS];
  unsigned* flag[MAX_RECORD_PEERS];
};
struct FinalizeArgs {
  const SegState* seg;
  const ItemState* item;
  const int32_t* item_first_seg;
  double decay_thr;
  rho_record* rec;
  const float* emb;
  const float* ref;
  int dim;
  RecordPeers peers;
};
// fill_to (0 = pad_frames): frames >= fill_to are not materialised (compact feature rows); pad_value[c] (optional)
// receives the constant every frame >= T_real of clip c holds
cudaError_t launch_logmel_norm(const int32_t* len16, int n, int n_mels, int pad_frames, float* mel,
                               int64_t mel_stride_frames, const int* clip_max, cudaStream_t st, LaunchCtx* lc,
                               bool fill_done = false, const FinalizeArgs* fin = nullptr, int fill_to = 0,
                               float* pad_value = nullptr);

// fused.cu
bool fused_inline_norm();   // the fused kernel writes the constant fill of the zero-padding frames itself
cudaError_t upload_fused_taps(const float* taps /* [2][23] */);
cudaError_t upload_fused_mel(int which, int n_mels, const int* lo, const int* cnt, const int* wofs, const float* w, int nnz);
constexpr int FUSED_MEL_STREAM_FLOAT4 = 640;   // capacity of one bank's stream (fused.cu: FZ_MEL_TAB4)
int build_fused_mel_stream(int which, int n_mels, const int* lo, const int* cnt, const int* wofs, const float* w, int nnz,
                           float4* tab, int* part, int* part4, int* rows_per_bundle);
cudaError_t launch_fused_features(const Tables& tb, const float* x, const int64_t* seg_off, const Workspace& ws,
                                  const int32_t* item_first_seg, int n_items, int64_t max_len,
                                  const Derived& d, float* y, const int64_t* y_off, int n_mels, int pad_frames,
                                  float* mel, int64_t mel_stride_frames, int sm_count, cudaStream_t st, LaunchCtx* lc,
                                  int mode = 0, int fill_to = 0);   // mode 0: one-segment items, 1: features of the
                                                                    // finished audio in y, 2: joined items (y written here)

// mel_gemm.cu: the mel projection as a tcgen05 / TMEM / TMA GEMM (3xTF32), in isolation
cudaError_t launch_mel_gemm(const Tables& tb, const float* power, int64_t n_frames, int64_t ld_power, int n_mels,
                            float* mel, int64_t ld_mel, int64_t frames_per_item, int64_t item_stride, int sm_count,
                            cudaStream_t st, LaunchCtx* lc);

// stft_tc.cu: the windowed DFT as two tcgen05 GEMMs (400 = 25 x 16), in isolation
size_t stft_tc_table_bytes();
void host_stft_tc_tables(unsigned char* out /* stft_tc_table_bytes() */);
cudaError_t launch_stft_tc(const unsigned char* tables, const float* x16, const int64_t* off, const int32_t* len16,
                           int pad_frames, const int32_t* tiles, int n_tiles, float* power, int64_t ld_power, int sm_count,
                           cudaStream_t st, LaunchCtx* lc);

// qwen.cu: QwenTTS._post_process_audio (windowed decay correction, -23 dBFS, tanh soft clip)
size_t qwen_workspace_bytes(int n, int64_t max_len, int sr);
cudaError_t launch_qwen_postprocess(const float* x, const int64_t* off, const int32_t* len, int len_stride_bytes,
                                    int n, int64_t max_len, int sr, float* y, const int64_t* y_off,
                                    void* workspace, cudaStream_t st, LaunchCtx* lc);

// pitch.cu: torchaudio.functional.pitch_shift (stft 512/128 -> phase vocoder -> istft -> resample -> crop / pad)
struct PitchTables {
  float2* w256;    // [256] exp(-2 pi i k / 256)
  float2* w512;    // [257] exp(-2 pi i k / 512)
  float* hann;     // [512] torch.hann_window(512)
  float* padv;     // [257] torch.linspace(0, pi * 128, 257): phase_advance
};
struct PitchPlan {
  int64_t T_max, J_max, LS_max;
  size_t spec, mag, ph, wave, total;
};
PitchPlan pitch_plan(int n, int64_t max_len, double rate);
cudaError_t launch_pitch_shift(const PitchTables& tb, const float* x, const int64_t* off, const int32_t* len,
                               int len_stride_bytes, int n, int64_t max_len, double rate, int arange_vec, int orig,
                               int nw, int width, int W, const float* taps, const int* ilo, float* y,
                               const int64_t* y_off, void* workspace, cudaStream_t st, LaunchCtx* lc);

// mfcc.cu: librosa.feature.mfcc(sr=16000, n_mfcc=13) mean / std per clip (drift-classifier front end)
struct MfccTables {
  float2* w256;    // [256]  exp(-2 pi i k / 256)
  float2* w1024;   // [1024] exp(-2 pi i k / 1024)
  float2* w2048;   // [1025] exp(-2 pi i k / 2048)
  float* hann;     // [2048]
  int* mel_lo;     // [128] sparse slaney filterbank over 1025 bins
  int* mel_cnt;
  int* mel_wofs;
  float* mel_w;
  int mel_nnz;     // floats in mel_w
  float* dct;      // [13][128]
};
size_t mfcc_workspace_bytes(int n, int64_t max_len);
cudaError_t launch_mfcc_stats(const MfccTables& tb, const float* x, const int64_t* off, const int32_t* len,
                              int len_stride_bytes, int n, int64_t max_len, float* out, void* workspace,
                              cudaStream_t st, LaunchCtx* lc);

// spk.cu: resemblyzer's front end (volume normalisation, 40-band mel spectrogram, partial utterances, embedding pooling)
constexpr int SPK_MELS = 40;
constexpr int SPK_PART_FRAMES = 160;          // partials_n_frames
constexpr int SPK_SUM_CHUNK = 65536;          // samples per CTA of the sum-of-squares / scaling kernels
// resemblyzer/voice_encoder.py compute_partial_slices: how many partial utterances a clip of n_samples gives (partial j =
// mel frames [frame_step j, frame_step j + 160)), and the length embed_utterance zero-pads the clip to when the last
// partial sticks out (*padded_len = end of the last partial in samples; smaller than n_samples if nothing sticks out)
__host__ __device__ inline int spk_slices(long long n_samples, int frame_step, double min_coverage, long long* padded_len) {
  const long long n_frames = (n_samples + 1 + HOP16 - 1) / HOP16;                 // ceil((n + 1) / 160)
  long long steps = n_frames - SPK_PART_FRAMES + frame_step + 1;
  if (steps < 1) steps = 1;
  long long count = (steps + frame_step - 1) / frame_step;                        // len(range(0, steps, frame_step))
  const long long last = (count - 1) * frame_step * HOP16;                        // first sample of the last partial
  const double coverage = (double)(n_samples - last) / (double)(SPK_PART_FRAMES * HOP16);
  if (coverage < min_coverage && count > 1) --count;
  if (padded_len) *padded_len = ((count - 1) * frame_step + SPK_PART_FRAMES) * HOP16;
  return (int)count;
}
cudaError_t launch_spk_normalize(const float* x, const int64_t* off, const int32_t* len, int len_stride_bytes, int n,
                                 int64_t max_len, float target_dbfs, int mode, float* y, const int64_t* y_off,
                                 float* gain, double* sums, cudaStream_t st, LaunchCtx* lc);
cudaError_t launch_spk_mel(const Tables& tb, const float* x16, const int64_t* off, const int32_t* len,
                           int len_stride_bytes, int n, int64_t max_len, int frame_step, double min_coverage,
                           bool pad_to_slices, const float* gain, float* mel, const int64_t* frame_off, float* partials,
                           const int32_t* part_off, cudaStream_t st, LaunchCtx* lc);
cudaError_t launch_spk_pool(const float* partial_embeds, const int32_t* part_off, int n, int dim, float* out,
                            cudaStream_t st, LaunchCtx* lc);

// cosine.cu
cudaError_t launch_cosine(const float* emb, const float* ref, int n, int dim, float* out, int out_stride_bytes,
                          cudaStream_t st, LaunchCtx* lc);

// tables.cpp (host only, no CUDA)
void host_resample_taps(float* out /* [2][23] */);
int host_resample_width(int orig, int nw);
void host_resample_taps_general(int orig, int nw, float* out /* [nw][2*width+orig] */);
void host_resample_taps_windowed(int orig, int nw, int width, int W, float* taps /* [nw][W] */, int* ilo /* [nw] */);
void host_pitch_tables(float* w256 /* [256][2] */, float* w512 /* [257][2] */, float* hann512, float* padv /* [257] */);
void host_hann(float* out /* [400] */);
void host_mel_filterbank(int n_mels, float* out /* [n_mels][201] */);
void host_mel_filterbank_bins(int n_mels, int n_bins, float* out /* [n_mels][n_bins] */);
void host_mfcc_tables(float* hann2048, float* w1024 /* [1024][2] */, float* w2048 /* [1025][2] */, float* dct /* [13][128] */);
void host_twiddles(float* out /* [400][2] */);

}  // namespace rho
