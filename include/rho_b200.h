/*
 * rho_b200.h -- C ABI of librho_b200.so: the B200 (sm_100a) implementation of the
 * rho-tts audio post-processing + validation front end.
 *
 * The reference (rhofield/rho-tts) is pure Python and has no FFI for this path; its
 * extension point is "subclass BaseTTS, register with TTSFactory"
 * (src/rho_tts/factory.py:110-122).  The Python shim in rho_tts_b200/ keeps those
 * signatures and calls the entry points below through ctypes.  Each entry point
 * names the reference code it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - Plain C types only.  `stream` is a cudaStream_t passed as void*.
 *   - Unless a function says HOST, every data pointer is a DEVICE pointer owned by
 *     the caller; kernels are enqueued on `stream` and the call returns without
 *     synchronising.  No allocation happens inside those calls: scratch comes from
 *     the caller-provided workspace (size it with rho_b200_workspace_bytes).
 *   - Ragged batches: `x` is one fp32 buffer, clip/segment s occupies
 *     x[off[s] .. off[s]+len[s]).  Every off[s] must be a multiple of 4 (16-byte
 *     aligned) so 128-bit loads are legal.
 *   - Return value: 0 on success, negative rho_status on failure; text via
 *     rho_b200_last_error().  There is no CPU fallback anywhere.
 */
#ifndef RHO_B200_H
#define RHO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RHO_B200_ABI_VERSION 2

typedef struct rho_handle rho_handle;

typedef enum {
  RHO_OK = 0,
  RHO_ERR_INVALID = -1,   /* bad argument */
  RHO_ERR_CUDA = -2,      /* CUDA runtime error */
  RHO_ERR_NOGPU = -3,     /* no sm_100 device */
  RHO_ERR_WORKSPACE = -4, /* workspace too small */
  RHO_ERR_LAYOUT = -5     /* misaligned offsets / inconsistent ragged layout */
} rho_status;

/* Call-time configuration.  Mirrors the BaseTTS attributes read at call time
 * (src/rho_tts/base_tts.py:72-81, 366-367, 420, 455, 519). */
typedef struct {
  int32_t sr;             /* sample_rate */
  int32_t trim_enabled;   /* self.trim_silence */
  double silence_db;      /* self.silence_threshold_db        (-50.0) */
  double fade_sec;        /* self.fade_duration_sec           (0.02)  */
  double xfade_sec;       /* self.crossfade_duration_sec      (0.05)  */
  double pause_sec;       /* self.inter_sentence_pause_sec    (0.1)   */
  double decay_thr;       /* self.sound_decay_threshold       (0.3)   */
} rho_params;

/* flags in rho_seg_info.flags / rho_record.flags */
#define RHO_F_ALL_SILENT 1u /* no frame above threshold: reference returns a 2-D (1, min(window,L)) view */
#define RHO_F_FALLBACK   2u /* join took the reference's except-branch: plain cat of the ORIGINAL segments */
#define RHO_F_TWO_D      4u /* result tensor is (1, n) in the reference, not (n,) */
#define RHO_F_UNTOUCHED  8u /* trimming disabled or empty input */

/* Per-segment trim result (16 B).  base_tts.py:348-399. */
typedef struct {
  int32_t start;  /* first kept sample  */
  int32_t end;    /* one past last kept */
  float dc;       /* mean over [start,end) that was subtracted */
  uint32_t flags;
} rho_seg_info;

/* Per-item score record (48 B): the thing that is all-gathered across GPUs. */
typedef struct {
  int32_t start;        /* segment 0 trim start                                   */
  int32_t end;          /* segment 0 trim end                                     */
  int32_t out_len;      /* samples written for this item                          */
  uint32_t flags;
  float dc;             /* segment 0 DC                                           */
  float first_rms;      /* base_tts.py:314                                        */
  float last_rms;       /* base_tts.py:315                                        */
  float cosine;         /* base_tts.py:341-344 (0 until rho_b200_cosine ran)      */
  double decay_ratio;   /* last/first in double, 1.0 on the early-outs (:304-318) */
  int32_t ok;           /* ratio >= decay_thr (:322)                              */
  int32_t n_segments;
} rho_record;

/* ------------------------------------------------------------------ lifecycle */
int rho_b200_abi_version(void);
/* Creates a handle on CUDA device `device`; builds and uploads the constant tables
 * (resample taps, Hann window, mel filterbanks 80/128, FFT twiddles). */
int rho_b200_create(rho_handle** h, int device);
int rho_b200_destroy(rho_handle* h);
/* Thread-local message of the last failing call on this thread. */
const char* rho_b200_last_error(void);

/* HOST-side tables, computable without a GPU (used by the CPU test-suite to compare
 * the library's constants with torchaudio / transformers):
 *   kind 0: resample taps 24k->16k, out[2*23]      (torchaudio functional.py:1305-1405)
 *   kind 1: periodic Hann(400), out[400]           (feature_extraction_whisper.py:141)
 *   kind 2: mel filterbank, arg = n_mels (80, 128: Whisper; 40: speaker encoder), out[n_mels*201] row-major [mel][bin]
 *                                                   (transformers audio_utils.py:453-544)
 *   kind 3: pitch shift phase_advance = torch.linspace(0, pi*128, 257) as torch's fp32 kernel makes it, out[257]
 *   kind 4: MFCC DCT-II rows (orthonormal), out[13*128]
 *   kind 5: MFCC slaney filterbank, 128 bands over the 1025 bins of a 2048-point FFT at 16 kHz, out[128*1025]
 *   kind 6: windowed resample taps of arg -> 24000 Hz (reduced ratio orig:new), out[new * W], W = 2*width+2 rounded
 *           up to a multiple of 4                  (torchaudio functional.py:1305-1405, taps inside the Hann window)
 * Returns the number of floats written, or a negative status. */
int rho_b200_host_table(int kind, int arg, float* out, size_t out_capacity);

/* Host only: the mel filterbank (n_mels = 80 or 128) as the fused kernel walks it -- one stream of float4 in constant
 * memory (csrc/fused.cu): bundles of *rows_per_bundle rows, each a header {byte offset of every row's first bin in a
 * power row ..., bytes of power each row covers} padded to whole float4, then the rows' weights (x 1/4: the kernel's
 * spectra are 4 |.|^2) group of four by group of four, the rows of the bundle interleaved, zero-padded.  part[11] /
 * part4[11]: the rows / the float4 each of the ten warps of a half starts at.  Returns the floats written (capacity at
 * least 2560), or a negative status.  For tests: the packing is checked against kind 2 of rho_b200_host_table. */
int rho_b200_host_mel_stream(int n_mels, float* stream, size_t out_capacity, int32_t* part, int32_t* part4, int* rows_per_bundle);

/* Bytes of scratch the calls below need for a batch of `n_segments` segments in
 * `n_items` items whose longest segment has `max_seg_len` samples. */
size_t rho_b200_workspace_bytes(int n_segments, int n_items, int64_t max_seg_len);

/* ------------------------------------------------- post-process / join (a1-a5) */
/* Silence-trim bounds only.  Replaces BaseTTS._trim_silence (base_tts.py:348-392).
 * trim_flags[s]: bit0 = from_start, bit1 = from_end (NULL = both).
 * Writes info[s] = {start, end, dc over [start,end), flags}. */
int rho_b200_trim_scan(rho_handle* h, const float* x, const int64_t* off, const int32_t* len,
                       const uint8_t* trim_flags, int n_segments, int64_t max_seg_len,
                       const rho_params* p, rho_seg_info* info,
                       void* workspace, size_t ws_bytes, void* stream);

/* The whole per-item path of BaseTTS._run_pipeline after generation
 * (base_tts.py:912-926): _smooth_segment_join (:435-536: per-segment trim + DC,
 * equal-power crossfades, inter-sentence pauses, fallback on mixed ranks, final fades)
 * followed by _validate_sound_decay (:297-323).  Item i owns segments
 * [item_first_seg[i], item_first_seg[i+1]).  A one-segment item is the plain
 * post-process of one clip (trim both ends -> DC -> fades).
 * y_off[i] (multiple of 4) is where item i is written; the caller reserves
 * sum(len of its segments) + max(0, n-2)*pause samples for it.
 * Outputs: y, rec[i], and (optional, may be NULL) seg_info[s]. */
int rho_b200_join(rho_handle* h, const float* x, const int64_t* seg_off, const int32_t* seg_len,
                  int n_segments, int64_t max_seg_len,
                  const int32_t* item_first_seg, int n_items, int64_t max_item_len,
                  const rho_params* p, float* y, const int64_t* y_off,
                  rho_record* rec, rho_seg_info* seg_info,
                  void* workspace, size_t ws_bytes, void* stream);

/* Element-wise pieces exposed one by one for the BaseTTS method shim.  In place on x. */
/* BaseTTS._remove_dc_offset (base_tts.py:394-399): x -= mean(x).  dc_out (device, 1 float) optional. */
int rho_b200_remove_dc(rho_handle* h, float* x, int64_t n, float* dc_out,
                       void* workspace, size_t ws_bytes, void* stream);
/* BaseTTS._apply_fades (base_tts.py:401-433). */
int rho_b200_apply_fades(rho_handle* h, float* x, int64_t n, int fade_in, int fade_out,
                         const rho_params* p, void* stream);
/* BaseTTS._validate_sound_decay (base_tts.py:297-323) on one clip; rec (device) gets
 * first_rms/last_rms/decay_ratio/ok. */
int rho_b200_sound_decay(rho_handle* h, const float* x, int64_t n, const rho_params* p,
                         rho_record* rec, void* workspace, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------ resample (a7) */
/* torchaudio.functional.resample(x, 24000, 16000) (torchaudio functional.py:1405-1432;
 * reference call sites base_tts.py:632 and the 16 kHz loaders behind :338 / stt_validator).
 * Clip s has len[s] input samples (device array, may be produced by rho_b200_join:
 * pass &rec[0].out_len with len_stride = sizeof(rho_record)); writes
 * ceil(2*len/3) samples at y + y_off[s] and that count to y_len[s]. */
int rho_b200_resample3to2(rho_handle* h, const float* x, const int64_t* off, const int32_t* len,
                          int len_stride_bytes, int n, int64_t max_len,
                          float* y, const int64_t* y_off, int32_t* y_len, void* stream);

/* -------------------------------------------------------------- log-mel (a8) */
/* WhisperFeatureExtractor (transformers feature_extraction_whisper.py:135-164, 296-303):
 * STFT(400, hop 160, periodic Hann, centre/reflect) -> |.|^2 -> mel(n_mels in {80,128})
 * -> log10(clamp 1e-10) -> max(x, clipmax-8) -> (x+4)/4.
 * pad_frames = 3000: pad/truncate every clip to 30 s, output [n][n_mels][3000].
 * pad_frames = 0   : unpadded; clip s gets T_s = len16[s]/160 frames written with row
 *                    stride `mel_stride_frames` at mel + s*n_mels*mel_stride_frames,
 *                    and n_frames[s] = T_s.
 * n_frames may be NULL when pad_frames = 3000. */
int rho_b200_logmel(rho_handle* h, const float* x16, const int64_t* off, const int32_t* len16,
                    int n, int64_t max_len16, int n_mels, int pad_frames,
                    float* mel, int64_t mel_stride_frames, int32_t* n_frames,
                    void* workspace, size_t ws_bytes, void* stream);

/* --------------------------------------------------------------- mel projection as a GEMM (a8) */
/* mel[i*item_stride + m*ld_mel + t] = sum_k filterbank[m][k] * power[(i*frames_per_item + t)*ld_power + k],
 * m < n_mels (80|128), k < 201, i*frames_per_item + t < n_frames  (frames_per_item <= 0: one item, all frames):
 * `mel_filters.T @ magnitudes` of transformers feature_extraction_whisper.py:159, reached by the reference
 * through stt_validator.py:78-107.  Runs on the tensor cores (tcgen05.mma kind::tf32, 3xTF32 split for fp32-class
 * accuracy, filterbank resident in TMEM, power tiles streamed by TMA).  This is the contraction measured on its
 * own for the tensor-pipe number; rho_b200_logmel / rho_b200_validate keep the sparse (97.5 % zeros) FFMA form.
 * power: device, fp32, 16-byte aligned, ld_power % 4 == 0 and >= 201.  Enqueues on `stream`, never syncs. */
int rho_b200_mel_project(rho_handle* h, const float* power, int64_t n_frames, int64_t ld_power, int n_mels,
                         float* mel, int64_t ld_mel, int64_t frames_per_item, int64_t item_stride, void* stream);

/* --------------------------------------------------------------- windowed DFT as tensor-core GEMMs (a8, north_star) */
/* P[row][k] = |STFT|^2 of feature_extraction_whisper.py:149-158 (torch.stft n_fft 400 / hop 160 / periodic Hann /
 * centre + reflect, then magnitude squared), k <= 200, on the tensor cores: the DFT is factored 400 = 25 x 16 into two
 * tcgen05.mma.kind::tf32 GEMMs (3xTF32 split) with the twiddle between them (csrc/stft_tc.cu).  This is the
 * "windowed-DFT contraction" of BASELINE.json's north_star measured on its own, like rho_b200_mel_project; together
 * they are the log-mel front end on tensor cores only.  The product path keeps the shared-memory FFT.
 * x16 / off / len16 / pad_frames: as rho_b200_logmel.  tiles (device, 16-byte aligned, n_tiles x 4 int32): clip,
 * first frame, number of frames (1..8), first output row -- frame t of the tile goes to power + (row + t) * ld_power.
 * Only columns 0..200 of a row are written. */
int rho_b200_stft_power_tc(rho_handle* h, const float* x16, const int64_t* off, const int32_t* len16, int pad_frames,
                           const int32_t* tiles, int n_tiles, float* power, int64_t ld_power, void* stream);

/* --------------------------------------------------------------- 16-bit PCM payload (the caller after the path) */
/* BaseTTS._save_wav's in-tree WAV writer (base_tts.py:661-667; the fallback it takes when torchaudio.save has no
 * backend): (np.clip(audio, -1, 1) * 32767).astype(np.int16), i.e. an fp32 product truncated toward zero.  out[s] has
 * the length of clip s; with it a clip leaves the device as 2 bytes per sample instead of 4. */
int rho_b200_pcm16(rho_handle* h, const float* y, const int64_t* off, const int32_t* len, int len_stride_bytes, int n,
                   int64_t max_len, int16_t* out, const int64_t* out_off, void* stream);

/* --------------------------------------------------------------- pitch shift (NEXT-4, the pitch half) */
/* The pitch branch of BaseTTS._apply_speed_pitch (base_tts.py:639-648): torchaudio.functional.pitch_shift(audio,
 * sample_rate, n_steps) = stft(512, hop 128) -> phase_vocoder(rate = 2^(-n_steps / 12)) -> istft(round(L / rate)) ->
 * resample(int(sample_rate / rate) -> sample_rate) -> crop / zero-pad to L (functional.py:1596-1713, 723-803), for n
 * clips; y[s] has the length of x[s].  The fp32 reference is sensitive to its own rounding (phase accumulator ~1e6
 * rad): arange_vec is the lane count of the vectorised torch.arange kernel of the torch build being mirrored (8 for
 * the 2.11 CPU wheels; 0 = float(rate * j)).  min_len is the caller's statement of the shortest clip: like
 * torch.stft's reflect padding the call refuses clips of <= 256 samples.  n_steps == 0 is refused (the reference
 * never makes that call).  Exact mirror of the time steps up to 32768 output frames per clip (~170 s at 24 kHz).
 * workspace: rho_b200_pitch_workspace_bytes(n, max_len, n_steps), 256-byte aligned. */
size_t rho_b200_pitch_workspace_bytes(int n, int64_t max_len, double n_steps);
int rho_b200_pitch_shift(rho_handle* h, const float* x, const int64_t* off, const int32_t* len, int len_stride_bytes,
                         int n, int64_t min_len, int64_t max_len, int sample_rate, double n_steps, int arange_vec,
                         float* y, const int64_t* y_off, void* workspace, size_t ws_bytes, void* stream);

/* --------------------------------------------------------------- MFCC statistics (NEXT-3, drift-classifier front end) */
/* validation/classifier/trainer.py:50-52: mfcc = librosa.feature.mfcc(y=y, sr=16000, n_mfcc=13); mean and std over the
 * frames, for n clips of 16 kHz audio (e.g. the output of rho_b200_resample3to2).  librosa >= 0.10 semantics: stft 2048 /
 * hop 512, periodic hann, centred with zero padding, |X|^2, 128 slaney mel bands, 10 log10(max(1e-10, .)) clamped at
 * (clip maximum - 80), orthonormal DCT-II, first 13.  out[s] = 13 means then 13 population standard deviations.
 * librosa is absent from the authoring image: the oracle is pinned on transformers' port of these steps and on scipy's
 * DCT (oracle/mfcc.py).  Clips must not be empty.  workspace: rho_b200_mfcc_workspace_bytes(n, max_len), 256-byte
 * aligned. */
size_t rho_b200_mfcc_workspace_bytes(int n, int64_t max_len);
int rho_b200_mfcc_stats(rho_handle* h, const float* x16, const int64_t* off, const int32_t* len, int len_stride_bytes,
                        int n, int64_t max_len, float* out, void* workspace, size_t ws_bytes, void* stream);

/* --------------------------------------------------------------- speaker-encoder front end (NEXT-3, resemblyzer) */
/* What resemblyzer runs on the CPU between a 16 kHz waveform and its LSTM, reached from BaseTTS._compute_speaker_similarity
 * (base_tts.py:326-347: preprocess_wav -> voice_encoder.embed_utterance -> cosine with the reference embedding),
 * QwenTTS._initialize_reference_embedding (providers/qwen.py:199-216) and validation/classifier/trainer.py:41-47.
 * resemblyzer (pyproject.toml: `resemblyzer>=0.1.4`) is not vendored by the reference: oracle/speaker.py restates
 * audio.py / voice_encoder.py of 0.1.4 and is pinned on transformers' port of librosa's spectrogram.  Not replaced:
 * librosa.resample (the 16 kHz signal is the input; rho_b200_resample3to2 is this library's 24 -> 16 kHz), webrtcvad
 * (trim_long_silences: run it on the host between the two calls if wanted) and the LSTM itself. */

/* compute_partial_slices (voice_encoder.py): number of partial utterances of a clip of n_samples -- partial j is mel frames
 * [frame_step*j, frame_step*j + 160) -- and, in *padded_len, the sample the last partial ends at (embed_utterance
 * zero-pads the clip to it when it lies past the end).  frame_step = round(16000 / rate / 160) = 77 for the default rate
 * 1.3; min_coverage 0.75.  Host only, no GPU. */
int rho_b200_spk_slices(int64_t n_samples, int frame_step, double min_coverage, int64_t* padded_len);

/* normalize_volume (audio.py): gain[s] = 10^((target_dbfs - dBFS(clip s)) / 20) in float32 like numpy's, or 1 where the
 * mode forbids the change (mode 0 = always, 1 = increase only -- what preprocess_wav uses with -30 dBFS --, 2 = decrease
 * only).  y (may be NULL: gains only; may alias x) receives the scaled clips.  workspace: 8 bytes per clip, 8-byte
 * aligned. */
int rho_b200_normalize_volume(rho_handle* h, const float* x, const int64_t* off, const int32_t* len, int len_stride_bytes,
                              int n, int64_t max_len, float target_dbfs, int mode, float* y, const int64_t* y_off,
                              float* gain, void* workspace, size_t ws_bytes, void* stream);

#define RHO_SPK_PAD_TO_SLICES 1u   /* zero-pad every clip to the end of its last partial first, as embed_utterance does */
/* wav_to_mel_spectrogram (audio.py: librosa.feature.melspectrogram(sr 16000, n_fft 400, hop 160, n_mels 40), transposed):
 * periodic hann, centred with zero padding, |X|^2, 40 slaney bands, no log.  Clip s gives T_s = 1 + len_s / 160 rows
 * (len_s after the padding of RHO_SPK_PAD_TO_SLICES) of 40 floats.
 *   mel      (may be NULL) [frame_off[s] + t][40]
 *   partials (may be NULL) [part_off[s] + j][160][40] = rows [frame_step*j, frame_step*j + 160) of the padded clip's
 *            spectrogram, j < rho_b200_spk_slices(len_s): the LSTM's input batch (voice_encoder.py embed_utterance);
 *            implies the padding
 *   gain     (may be NULL) per-clip factor applied to the samples first (rho_b200_normalize_volume with y = NULL) */
int rho_b200_spk_mel(rho_handle* h, const float* x16, const int64_t* off, const int32_t* len, int len_stride_bytes, int n,
                     int64_t max_len, int frame_step, double min_coverage, unsigned flags, const float* gain, float* mel,
                     const int64_t* frame_off, float* partials, const int32_t* part_off, void* stream);

/* The tail of embed_utterance: out[s] = mean of partial_embeds[part_off[s] .. part_off[s+1]) (rows of `dim` floats, the
 * LSTM's output), divided by its L2 norm.  part_off has n + 1 entries. */
int rho_b200_spk_pool(rho_handle* h, const float* partial_embeds, const int32_t* part_off, int n, int dim, float* out,
                      void* stream);

/* --------------------------------------------------------------- batched decay check on finished audio (a5) */
/* _validate_sound_decay (base_tts.py:297-323) for n clips that are already final, e.g. after the Qwen loudness hook,
 * which the pipeline runs between the join and the decay check (base_tts.py:911-926).  Rewrites first_rms, last_rms,
 * decay_ratio and ok of rec[s] (everything else is kept); clip s is y + off[s], length read with len_stride_bytes
 * like the other entry points.  workspace: 16 bytes per clip, 8-byte aligned. */
int rho_b200_sound_decay_batch(rho_handle* h, const float* y, const int64_t* off, const int32_t* len,
                               int len_stride_bytes, int n, int64_t max_len, const rho_params* p, rho_record* rec,
                               void* workspace, size_t ws_bytes, void* stream);

/* --------------------------------------------------------------- any-ratio resample (NEXT-4, speed control) */
/* torchaudio.functional.resample(x, orig_freq, new_freq) (sinc_interp_hann, width 6, rolloff 0.99; functional.py:1305-1432)
 * for n clips: what BaseTTS._apply_speed_pitch does for speed != 1 (base_tts.py:631-637: orig = int(sr * speed), new = sr).
 * Clip s is read at x + off[s] and written at y + y_off[s]; y_len[s] (optional) receives
 * rho_b200_resample_out_len(len[s], ...) = ceil(new * len / orig) after reduction by the gcd.  The first call with a
 * new reduced ratio builds its tap table on the host and uploads it (synchronously); later calls only enqueue.
 * orig_freq == new_freq is refused (torchaudio returns its input; so does the host mirror). */
int64_t rho_b200_resample_out_len(int64_t len, int orig_freq, int new_freq);
int rho_b200_resample(rho_handle* h, const float* x, const int64_t* off, const int32_t* len, int len_stride_bytes,
                      int n, int64_t max_len, int orig_freq, int new_freq, float* y, const int64_t* y_off,
                      int32_t* y_len, void* stream);

/* --------------------------------------------------------------- Qwen loudness post-process (a9 / NEXT-1) */
/* QwenTTS._post_process_audio (providers/qwen.py:268-378) for n clips: overall-RMS gate (1e-8, clip copied unchanged),
 * windowed decay correction (2 s windows, gain to the first window's RMS capped at +18 dB, applied only when
 * n > 2 windows and the gain range is >= 0.05; two 3-tap smoothing passes; np.interp between window centres),
 * global gain to -23 dBFS, tanh(x / 0.95) * 0.95.  `sr` is the provider's qwen3_sr (24000 by default, :294).
 * Clip s is read at x + off[s] (length *(int32*)((char*)len + s*len_stride_bytes), len_stride_bytes 0 = 4) and
 * written at y + y_off[s]; y may alias x.  One read pass + one read/write pass.  Never syncs. */
size_t rho_b200_qwen_workspace_bytes(int n, int64_t max_len, int sr);
int rho_b200_qwen_postprocess(rho_handle* h, const float* x, const int64_t* off, const int32_t* len,
                              int len_stride_bytes, int n, int64_t max_len, int sr, float* y, const int64_t* y_off,
                              void* workspace, size_t ws_bytes, void* stream);

/* --------------------------------------------------------------- cosine (a6) */
/* dot(ref, e) / (|ref| * |e|) for n embeddings of dimension dim (base_tts.py:341-344).
 * out_stride_bytes lets the result land in rho_record.cosine. */
int rho_b200_cosine(rho_handle* h, const float* emb, const float* ref, int n, int dim,
                    float* out, int out_stride_bytes, void* stream);

/* ------------------------------------------------ whole validation front end */
/* join/post-process -> resample 24k->16k -> log-mel -> cosine, device resident: what every generated item passes
 * through between generation and the accept / retry decision (base_tts.py:912-926 followed by the validation front
 * end, stt_validator.py:78-107 -> feature_extraction_whisper.py:135-164, and the speaker cosine base_tts.py:341-344).
 *   one-segment items (RHO_V_ONE_SEGMENT_ITEMS: the caller asserts item_first_seg = 0,1,2,...): ONE kernel applies
 *     DC / fades, resamples and computes the log-mel frames in one pass over each clip;
 *   joined items: the same kernel also joins -- batches inside one segment are the one-segment case, batches at a
 *     joint (crossfade, pause, item end) are computed sample by sample -- and writes y; RHO_V_GATHER_FIRST keeps the
 *     two-kernel path (the join writes y, the kernel reads the finished y back) for A/B measurements.
 *   Either way the 16 kHz signal never leaves shared memory; scratch16 is only used with RHO_V_NO_FUSION (the
 *   kernel-per-stage path kept for A/B measurements: same offsets as y, 2/3 the length) and may be NULL otherwise.
 * Features: mel + i*n_mels*mel_stride_frames, rows of mel_stride_frames floats.
 *   pad_frames = 3000, default: mel_stride_frames >= 3000, every row complete (feature_extraction_whisper.py:296-303).
 *   pad_frames = 3000, RHO_V_COMPACT_PAD: rows of mel_stride_frames < 3000 frames (a multiple of 4, at least
 *     rho_b200_compact_frames(max_item_len, 3000)): only the frames that can see signal are materialised; every frame
 *     t >= mel_stride_frames of item i equals pad_value[i] (the 30 s window of a 10 s clip is 2/3 such frames).
 *   pad_value (device, n_items floats, may be NULL): the constant of item i's zero-padding frames, in either layout.
 * Records: rec[i]; with a connected record exchange (rho_b200_exchange_*) the same record is also stored into every
 * rank's gathered buffer by the kernel that assembles it. */
#define RHO_V_ONE_SEGMENT_ITEMS 1u
#define RHO_V_NO_FUSION 2u
#define RHO_V_COMPACT_PAD 4u
#define RHO_V_GATHER_FIRST 8u
int rho_b200_validate(rho_handle* h, const float* x, const int64_t* seg_off, const int32_t* seg_len,
                      int n_segments, int64_t max_seg_len,
                      const int32_t* item_first_seg, int n_items, int64_t max_item_len,
                      const rho_params* p, float* y, const int64_t* y_off,
                      int n_mels, int pad_frames, float* mel, int64_t mel_stride_frames, float* pad_value,
                      const float* emb, const float* ref_emb, int emb_dim,
                      rho_record* rec, float* scratch16, uint32_t flags,
                      void* workspace, size_t ws_bytes, void* stream);
/* Frames per row a compact feature tensor needs for items of at most max_item_len 24 kHz samples (pad_frames = 3000:
 * ceil((min(ceil(2L/3), 480000) + 200) / 160) rounded up to a multiple of 4, at most 3000; pad_frames = 0: the
 * unpadded frame count ceil(2L/3) / 160). */
int64_t rho_b200_compact_frames(int64_t max_item_len, int pad_frames);

/* ------------------------------------------------ multi-GPU record exchange */
/* The one exchange of the path is the gather of the 48-byte records (one process per GPU, clips sharded with no
 * data-path collective).  It is fused into the record assembly instead of being a collective call on the critical path:
 *   create : allocates this rank's gathered buffer ([2 parities][world][n_per_rank] records + one flag word per source
 *            rank) and exports it as a 64-byte CUDA IPC handle (ipc_handle_out).  The caller all-gathers the handles
 *            (any transport: torch.distributed object collective, a file, MPI).
 *   connect: maps every peer's buffer into this process (cudaIpcOpenMemHandle, NVLink / NVSwitch peer memory);
 *            all_handles = world x 64 bytes in rank order.  From then on the kernel of rho_b200_validate that writes
 *            rec[i] also stores the record to EVERY rank's buffer at [epoch & 1][rank][i], and the last record of the
 *            call publishes flag[rank] = epoch on every rank (system-scope fence + release store); epoch counts the
 *            rho_b200_validate calls on this handle since connect (rho_b200_exchange_epoch).  Every rank must make the
 *            same sequence of calls.
 *   wait   : enqueues a one-warp kernel on `stream` that waits until all ranks' flags have reached `epoch` (bounded
 *            spin: after ~2 s it records a time-out instead of hanging the GPU).  Work enqueued behind it may read
 *            the gathered block of that epoch.  Two parities: epoch e's block is overwritten by epoch e + 2, so a rank
 *            waits for epoch e before it starts call e + 2 (rho_tts_b200.dist.RecordExchange does).
 *   read   : copies the gathered block of `epoch` (world * n_per_rank records, rank-major) to dst (device) on `stream`;
 *            with timed_out != NULL it also synchronises the stream and reports whether any wait so far ran into its
 *            time-out (1 + the rank that was missing).
 * Failure to map a peer (no IPC in this container, no P2P path) is an error of connect; the caller then keeps using
 * an NCCL all-gather of the records (rho_tts_b200.dist.gather_records). */
int rho_b200_exchange_create(rho_handle* h, int world, int rank, int64_t n_per_rank, void* ipc_handle_out /* 64 B */);
int rho_b200_exchange_connect(rho_handle* h, const void* all_handles /* world * 64 B */);
int rho_b200_exchange_wait(rho_handle* h, int64_t epoch, void* stream);
int64_t rho_b200_exchange_epoch(rho_handle* h);
int rho_b200_exchange_read(rho_handle* h, int64_t epoch, void* dst, int* timed_out, void* stream);
int rho_b200_exchange_destroy(rho_handle* h);

/* HOST entry points: the calls a non-torch embedder makes.  All pointers are HOST buffers (pinned for full speed).
 * Segments are copied in, rho_b200_validate (or rho_b200_join when mel == NULL: no features) runs in chunks of whole
 * items on internal streams (H2D, kernels and D2H overlapped), processed audio, records and features are copied back;
 * the call returns after the last copy.  Re-entrant: concurrent callers on one handle get their own streams / arena.
 *   x, seg_off, seg_len, item_first_seg: the ragged layout of rho_b200_join, in host memory; segments in increasing,
 *     non-overlapping order.  y, y_off: item i is written at y + y_off[i], the caller reserves sum(len) + pauses as
 *     for rho_b200_join, items in increasing order; host samples BETWEEN the items of a chunk are clobbered.
 *     Offsets that are all multiples of 4 give one copy per chunk and direction (else one per segment / item).
 *   mel (may be NULL), mel_stride_frames, pad_value (may be NULL): as rho_b200_validate with the layouts
 *     mel_stride_frames >= 3000 (complete rows) or < 3000 (compact rows + pad_value).  Only the frames that can see
 *     signal cross PCIe in either case: complete rows get their constant tail written by host threads from pad_value
 *     (RHO_HOST_FILL_THREADS, default 3), overlapped with the copies of the following chunks.
 *     mel may also be DEVICE memory of the handle's GPU (found out with cudaPointerGetAttributes): the rows are then
 *     written in place in HBM -- complete ones including their constant tail -- and nothing of the features crosses
 *     PCIe: the hand-off to a consumer on the device (the Whisper encoder, SURVEY.md 8f NEXT-2).
 *   emb [n_items][emb_dim], ref_emb [emb_dim] (may be NULL): speaker cosine into rec[i].cosine. */
int rho_b200_validate_host_ragged(rho_handle* h, const float* x, const int64_t* seg_off, const int32_t* seg_len,
                                  int n_segments, const int32_t* item_first_seg, int n_items, const rho_params* p,
                                  float* y, const int64_t* y_off, int n_mels, int pad_frames, float* mel,
                                  int64_t mel_stride_frames, float* pad_value, const float* emb,
                                  const float* ref_emb, int emb_dim, rho_record* rec);
/* Fixed-length layout: n clips of clip_len samples each, every item is one clip; complete feature rows
 * [n][n_mels][3000] (mel may be NULL).  A wrapper of rho_b200_validate_host_ragged. */
/* Host threads the next host-entry call on this handle uses to write the constant tails of complete feature rows
 * (default 3, RHO_HOST_FILL_THREADS pins it; raised by two, up to 9, after a call whose fill ended well behind its last
 * copy: a host with slower memory). */
int rho_b200_host_fill_threads(rho_handle* h);

int rho_b200_validate_host(rho_handle* h, const float* x, int n, int32_t clip_len,
                           const rho_params* p, float* y /* n*clip_len */, int n_mels, int pad_frames,
                           float* mel /* n*n_mels*pad_frames or NULL */,
                           const float* emb, const float* ref_emb, int emb_dim,
                           rho_record* rec);

/* Number of kernel launches this handle has issued (for bench.py's gpu_launches). */
int64_t rho_b200_launch_count(rho_handle* h);

/* Per-kernel device time, measured with CUDA events recorded on the launching stream around
 * every kernel issued between profile_begin and profile_end (profile_end synchronises on them).
 * ms_per_kernel[id] / launches_per_kernel[id] are indexed by kernel id; rho_b200_kernel_name(id)
 * names them.  Returns the number of kernel ids, or a negative status.  Not for timed runs. */
int rho_b200_profile_begin(rho_handle* h);
int rho_b200_profile_end(rho_handle* h, double* ms_per_kernel, int64_t* launches_per_kernel, int capacity);
/* How this build splits work between kernels (affects only which kernel the bench charges which bytes to). */
#define RHO_BUILD_FUSED_WRITES_FILL 1   /* k_fused_features writes the constant of the zero-padding frames, not k_logmel_norm */
int rho_b200_build_flags(void);
const char* rho_b200_kernel_name(int id);

#ifdef __cplusplus
}
#endif
#endif /* RHO_B200_H */
