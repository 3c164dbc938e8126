// C ABI of librho_b200.so (see include/rho_b200.h).  Host-side glue only: argument checks,
// constants derived the way the reference derives them, workspace carving, launches.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>
#include <cstdlib>
#include <map>
#include <mutex>
#include "handle.h"

namespace rho {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
int cuda_fail(cudaError_t e, const char* where) {
  return fail(RHO_ERR_CUDA, "%s: %s", where, cudaGetErrorString(e));
}

// Python: int(sr * 0.01), 10 ** (dB / 20), int(sr * sec) -- all in double, truncation toward zero.
Derived derive(const rho_params& p) {
  Derived d;
  d.window = (int)((double)p.sr * 0.01);
  d.hop = d.window / 2;
  d.fade = (int)((double)p.sr * p.fade_sec);
  d.cf = (int)((double)p.sr * p.xfade_sec);
  d.pause = (int)((double)p.sr * p.pause_sec);
  d.pause_on = p.pause_sec > 0.0 ? 1 : 0;
  if (d.pause < 0) d.pause = 0;
  d.trim_enabled = p.trim_enabled ? 1 : 0;
  d.thr = (float)std::pow(10.0, p.silence_db / 20.0);
  d.decay_thr = p.decay_thr;
  return d;
}

}  // namespace rho

using namespace rho;

namespace rho {
const char* const kKernelNames[KID_COUNT] = {
  "k_init", "k_scan", "k_finalize_segs", "k_plan_items", "k_gather", "k_finalize_items",
  "k_resample3to2", "k_logmel_init", "k_logmel_frames", "k_logmel_norm", "k_cosine", "k_single_clip_helpers",
  "k_fused_features", "k_mel_gemm", "k_qwen_moments", "k_qwen_plan", "k_qwen_apply", "k_resample_general",
  "k_pv_stft", "k_pv_phase", "k_pv_cumsum", "k_pv_istft", "k_resample_windowed", "k_mfcc_frames", "k_mfcc_stats",
  "k_wait_flags", "k_stft_tc", "k_spk_sumsq", "k_spk_scale", "k_spk_mel", "k_spk_pool"};
}

namespace {

// Lazily built tables: the copy from the (pageable) host vector is synchronous with respect to the host, but the
// kernel that reads the table is launched on the CALLER's stream, which need not be ordered after the legacy default
// stream (torch side streams are non-blocking).  Waiting for the device here orders the two; it happens once per table.
template <typename T>
cudaError_t dev_upload(rho_handle* h, T** dst, const T* src, size_t n) {
  cudaError_t e = cudaMalloc((void**)dst, n * sizeof(T));
  if (e != cudaSuccess) return e;
  h->allocs.push_back(*dst);
  e = cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return e;
  return cudaDeviceSynchronize();
}

struct WsPlan {
  size_t seg, span, item, block_sum, clip_max, len16, tiles_done, work_counter, scratch, total;
  int blocks_per_seg;
};

WsPlan plan_ws(int n_seg, int n_items, int64_t max_seg_len) {
  WsPlan w;
  // hop >= 1; the smallest hop we size for is 40 samples (sr 8 kHz).  blocks = frames upper bound.
  const int64_t min_hop = 40;
  w.blocks_per_seg = (int)(max_seg_len / min_hop + 2);
  size_t o = 0;
  w.seg = o; o += align_up(sizeof(SegState) * (size_t)(n_seg > 0 ? n_seg : 1), 256);
  w.span = o; o += align_up(sizeof(SegSpan) * (size_t)(n_seg > 0 ? n_seg : 1), 256);
  w.item = o; o += align_up(sizeof(ItemState) * (size_t)(n_items > 0 ? n_items : 1), 256);
  w.clip_max = o; o += align_up(sizeof(int) * (size_t)(n_items > 0 ? n_items : 1), 256);
  w.len16 = o; o += align_up(sizeof(int32_t) * (size_t)(n_items > 0 ? n_items : 1), 256);
  w.tiles_done = o; o += align_up(sizeof(int) * (size_t)(n_items > 0 ? n_items : 1), 256);
  w.work_counter = o; o += 256;
  w.scratch = o; o += 256;
  w.block_sum = o; o += align_up(sizeof(float) * (size_t)(n_seg > 0 ? n_seg : 1) * (size_t)w.blocks_per_seg, 256);
  w.total = o;
  return w;
}

// blocks_per_seg actually used by the kernels for this call (depends on the call's hop)
int blocks_for(const Derived& d, int64_t max_seg_len) {
  if (max_seg_len <= 0) return 1;
  return (int)((max_seg_len + 2 * d.hop - d.window) / d.hop + 1);
}

int carve(void* ws, size_t ws_bytes, int n_seg, int n_items, int64_t max_seg_len, const Derived& d,
          Workspace* out, double** scratch) {
  if (d.hop < 40) return fail(RHO_ERR_INVALID, "sample rate too low: hop %d < 40", d.hop);
  const WsPlan w = plan_ws(n_seg, n_items, max_seg_len);
  if (ws == nullptr || ws_bytes < w.total)
    return fail(RHO_ERR_WORKSPACE, "workspace too small: have %zu, need %zu", ws_bytes, w.total);
  if (((uintptr_t)ws & 255u) != 0) return fail(RHO_ERR_LAYOUT, "workspace must be 256-byte aligned");
  char* b = (char*)ws;
  out->seg = (SegState*)(b + w.seg);
  out->span = (SegSpan*)(b + w.span);
  out->item = (ItemState*)(b + w.item);
  out->clip_max = (int*)(b + w.clip_max);
  out->len16 = (int32_t*)(b + w.len16);
  out->tiles_done = (int*)(b + w.tiles_done);
  out->work_counter = (int*)(b + w.work_counter);
  out->block_sum = (float*)(b + w.block_sum);
  out->blocks_per_seg = blocks_for(d, max_seg_len);
  if (out->blocks_per_seg > w.blocks_per_seg) return fail(RHO_ERR_WORKSPACE, "internal: block_sum sizing");
  if (scratch) *scratch = (double*)(b + w.scratch);
  return RHO_OK;
}

int check_params(const rho_params* p) {
  if (!p) return fail(RHO_ERR_INVALID, "params is NULL");
  if (p->sr < 8000 || p->sr > 192000) return fail(RHO_ERR_INVALID, "sample rate %d out of range", p->sr);
  if (p->fade_sec < 0 || p->xfade_sec < 0) return fail(RHO_ERR_INVALID, "negative fade/crossfade duration");
  return RHO_OK;
}

}  // namespace

extern "C" {

int rho_b200_abi_version(void) { return RHO_B200_ABI_VERSION; }

const char* rho_b200_last_error(void) { return g_err; }

static long long gcd_ll(long long a, long long b);

int rho_b200_host_table(int kind, int arg, float* out, size_t cap) {
  if (!out) return fail(RHO_ERR_INVALID, "out is NULL");
  switch (kind) {
    case 0:
      if (cap < 2 * RS_TAPS) return fail(RHO_ERR_INVALID, "capacity");
      host_resample_taps(out); return 2 * RS_TAPS;
    case 1:
      if (cap < N_FFT) return fail(RHO_ERR_INVALID, "capacity");
      host_hann(out); return N_FFT;
    case 2:
      if (arg != 80 && arg != 128 && arg != SPK_MELS) return fail(RHO_ERR_INVALID, "n_mels must be 80, 128 or 40");
      if (cap < (size_t)arg * N_BINS) return fail(RHO_ERR_INVALID, "capacity");
      host_mel_filterbank(arg, out); return arg * N_BINS;
    case 3: {                                          // pitch shift: phase_advance = torch.linspace(0, pi * 128, 257)
      if (cap < 257) return fail(RHO_ERR_INVALID, "capacity");
      std::vector<float> w256(512), w512(514), hn(512);
      host_pitch_tables(w256.data(), w512.data(), hn.data(), out);
      return 257;
    }
    case 4: {                                          // MFCC: 13 x 128 orthonormal DCT-II rows
      if (cap < 13 * 128) return fail(RHO_ERR_INVALID, "capacity");
      std::vector<float> hn(2048), w1(2048), w2(2050);
      host_mfcc_tables(hn.data(), w1.data(), w2.data(), out);
      return 13 * 128;
    }
    case 5:                                            // MFCC: slaney filterbank, 128 bands over 1025 bins
      if (cap < (size_t)128 * 1025) return fail(RHO_ERR_INVALID, "capacity");
      host_mel_filterbank_bins(128, 1025, out); return 128 * 1025;
    case 6: {                                          // windowed resample taps of ratio arg : 24000 (reduced), [nw][W]
      if (arg <= 0 || arg == 24000) return fail(RHO_ERR_INVALID, "orig_freq");
      const long long g = gcd_ll(arg, 24000);
      const int orig = (int)(arg / g), nw = (int)(24000 / g);
      const int width = host_resample_width(orig, nw);
      const int W = (2 * width + 2 + 3) / 4 * 4;
      if (cap < (size_t)nw * W) return fail(RHO_ERR_INVALID, "capacity: need %zu", (size_t)nw * W);
      std::vector<int> lo((size_t)nw);
      host_resample_taps_windowed(orig, nw, width, W, out, lo.data());
      return nw * W;
    }
    default:
      return fail(RHO_ERR_INVALID, "unknown table kind %d", kind);
  }
}

// the sparse form of a slaney bank over the 201 bins: per row the first non-zero bin, the span to the last one, and the
// weights of that span back to back
static void sparse_mel_bank(int nm, std::vector<float>& dense, std::vector<int>& lo, std::vector<int>& cnt,
                            std::vector<int>& wofs, std::vector<float>& w) {
  dense.assign((size_t)nm * N_BINS, 0.f);
  host_mel_filterbank(nm, dense.data());
  lo.assign(nm, 0); cnt.assign(nm, 0); wofs.assign(nm, 0); w.clear();
  for (int m = 0; m < nm; ++m) {
    int a = -1, b = -1;
    for (int k = 0; k < N_BINS; ++k) if (dense[(size_t)m * N_BINS + k] != 0.f) { if (a < 0) a = k; b = k; }
    lo[m] = a < 0 ? 0 : a; cnt[m] = a < 0 ? 0 : b - a + 1; wofs[m] = (int)w.size();
    for (int k = 0; k < cnt[m]; ++k) w.push_back(dense[(size_t)m * N_BINS + lo[m] + k]);
  }
}

int rho_b200_host_mel_stream(int n_mels, float* stream, size_t cap_floats, int32_t* part, int32_t* part4, int* rows_per_bundle) {
  if (n_mels != 80 && n_mels != 128) return fail(RHO_ERR_INVALID, "n_mels must be 80 or 128");
  if (!stream || !part || !part4 || cap_floats < (size_t)4 * FUSED_MEL_STREAM_FLOAT4) return fail(RHO_ERR_INVALID, "capacity");
  std::vector<float> dense, w;
  std::vector<int> lo, cnt, wofs;
  sparse_mel_bank(n_mels, dense, lo, cnt, wofs, w);
  std::vector<float4> tab((size_t)FUSED_MEL_STREAM_FLOAT4);
  int p[12] = {}, p4[12] = {};
  const int n4 = build_fused_mel_stream(n_mels == 80 ? 0 : 1, n_mels, lo.data(), cnt.data(), wofs.data(), w.data(), (int)w.size(),
                                        tab.data(), p, p4, rows_per_bundle);
  if (n4 < 0) return fail(RHO_ERR_INVALID, "the filterbank does not fit the stream");
  memcpy(stream, tab.data(), sizeof(float4) * (size_t)n4);
  for (int i = 0; i < 11; ++i) { part[i] = p[i]; part4[i] = p4[i]; }
  return 4 * n4;
}

int rho_b200_create(rho_handle** out, int device) {
  if (!out) return fail(RHO_ERR_INVALID, "handle out pointer is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0)
    return fail(RHO_ERR_NOGPU, "no CUDA device visible (%s); librho_b200 has no CPU fallback",
                e == cudaSuccess ? "count=0" : cudaGetErrorString(e));
  if (device < 0 || device >= count) return fail(RHO_ERR_INVALID, "device %d out of range [0,%d)", device, count);
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceProperties");
  if (prop.major != 10)
    return fail(RHO_ERR_NOGPU, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
  DeviceGuard guard(device);                         // the caller's current device is restored on every return path
  if (guard.err != cudaSuccess) return cuda_fail(guard.err, "cudaSetDevice");

  rho_handle* h = new rho_handle();
  h->device = device; h->sm_count = prop.multiProcessorCount;
  memset(&h->tb, 0, sizeof(h->tb));

  std::vector<float> taps(2 * RS_TAPS), hann(N_FFT), tw(2 * N_FFT);
  host_resample_taps(taps.data()); host_hann(hann.data()); host_twiddles(tw.data());
  if ((e = upload_resample_taps(taps.data())) != cudaSuccess) { rho_b200_destroy(h); return cuda_fail(e, "taps"); }
  if ((e = upload_fused_taps(taps.data())) != cudaSuccess) { rho_b200_destroy(h); return cuda_fail(e, "fused taps"); }
  if ((e = dev_upload(h, &h->tb.hann, hann.data(), N_FFT)) != cudaSuccess) { rho_b200_destroy(h); return cuda_fail(e, "hann"); }
  if ((e = dev_upload(h, (float**)&h->tb.twiddle, tw.data(), 2 * N_FFT)) != cudaSuccess) { rho_b200_destroy(h); return cuda_fail(e, "twiddle"); }
  for (int which = 0; which < 2; ++which) {
    const int nm = which == 0 ? 80 : 128;
    std::vector<float> dense, w;
    std::vector<int> lo, cnt, wofs;
    sparse_mel_bank(nm, dense, lo, cnt, wofs, w);
    if (w.size() > 416) { rho_b200_destroy(h); return fail(RHO_ERR_INVALID, "mel filterbank nnz %zu > 416", w.size()); }
    h->tb.mel_nnz[which] = (int)w.size();
    if ((e = upload_fused_mel(which, nm, lo.data(), cnt.data(), wofs.data(), w.data(), (int)w.size())) != cudaSuccess) {
      rho_b200_destroy(h); return cuda_fail(e, "fused mel tables");
    }
    if ((e = dev_upload(h, &h->tb.mel_lo[which], lo.data(), nm)) != cudaSuccess ||
        (e = dev_upload(h, &h->tb.mel_cnt[which], cnt.data(), nm)) != cudaSuccess ||
        (e = dev_upload(h, &h->tb.mel_wofs[which], wofs.data(), nm)) != cudaSuccess ||
        (e = dev_upload(h, &h->tb.mel_w[which], w.data(), w.size())) != cudaSuccess ||
        (e = dev_upload(h, &h->tb.mel_dense[which], dense.data(), dense.size())) != cudaSuccess) {
      rho_b200_destroy(h); return cuda_fail(e, "mel tables");
    }
  }
  *out = h;
  return RHO_OK;
}

int rho_b200_destroy(rho_handle* h) {
  if (!h) return RHO_OK;
  DeviceGuard guard(h->device);
  for (void* p : h->allocs) cudaFree(p);
  for (HostCtx* c : h->host_ctx_free) host_ctx_destroy(c);
  delete h;
  return RHO_OK;
}

int64_t rho_b200_launch_count(rho_handle* h) { return h ? h->lc.launches.load() : (int64_t)0; }

int rho_b200_profile_begin(rho_handle* h) {
  RHO_ON_DEVICE(h);
  for (auto& sp : h->lc.spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
  h->lc.spans.clear();
  h->lc.profiling = true;
  return RHO_OK;
}

int rho_b200_profile_end(rho_handle* h, double* ms_per_kernel, int64_t* launches_per_kernel, int capacity) {
  RHO_ON_DEVICE(h);
  if (capacity < KID_COUNT || !ms_per_kernel || !launches_per_kernel) return fail(RHO_ERR_INVALID, "capacity < %d", (int)KID_COUNT);
  h->lc.profiling = false;
  for (int i = 0; i < KID_COUNT; ++i) { ms_per_kernel[i] = 0.0; launches_per_kernel[i] = 0; }
  int rc = RHO_OK;
  for (auto& sp : h->lc.spans) {
    cudaError_t e = cudaEventSynchronize(sp.b);
    float ms = 0.f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, sp.a, sp.b);
    if (e != cudaSuccess) rc = cuda_fail(e, "profile_end");
    else { ms_per_kernel[sp.id] += ms; launches_per_kernel[sp.id] += 1; }
    cudaEventDestroy(sp.a); cudaEventDestroy(sp.b);
  }
  h->lc.spans.clear();
  return rc == RHO_OK ? KID_COUNT : rc;
}

int rho_b200_build_flags(void) { return fused_inline_norm() ? RHO_BUILD_FUSED_WRITES_FILL : 0; }

const char* rho_b200_kernel_name(int id) { return (id >= 0 && id < KID_COUNT) ? kKernelNames[id] : ""; }

size_t rho_b200_workspace_bytes(int n_segments, int n_items, int64_t max_seg_len) {
  return plan_ws(n_segments, n_items, max_seg_len).total;
}

int rho_b200_trim_scan(rho_handle* h, const float* x, const int64_t* off, const int32_t* len,
                       const uint8_t* trim_flags, int n_segments, int64_t max_seg_len,
                       const rho_params* p, rho_seg_info* info, void* workspace, size_t ws_bytes, void* stream) {
  RHO_ON_DEVICE(h);
  int rc = check_params(p); if (rc) return rc;
  if (n_segments < 0 || max_seg_len < 0) return fail(RHO_ERR_INVALID, "negative size");
  if (n_segments == 0) return RHO_OK;
  if (!x || !off || !len || !info) return fail(RHO_ERR_INVALID, "NULL device pointer");
  const Derived d = derive(*p);
  Workspace ws;
  rc = carve(workspace, ws_bytes, n_segments, n_segments, max_seg_len, d, &ws, nullptr); if (rc) return rc;
  cudaError_t e = launch_trim_scan(x, off, len, trim_flags, n_segments, max_seg_len, d, ws, info,
                                   (cudaStream_t)stream, &h->lc);
  return e == cudaSuccess ? RHO_OK : cuda_fail(e, "trim_scan");
}

int rho_b200_join(rho_handle* h, const float* x, const int64_t* seg_off, const int32_t* seg_len,
                  int n_segments, int64_t max_seg_len, const int32_t* item_first_seg, int n_items,
                  int64_t max_item_len, const rho_params* p, float* y, const int64_t* y_off,
                  rho_record* rec, rho_seg_info* seg_info, void* workspace, size_t ws_bytes, void* stream) {
  RHO_ON_DEVICE(h);
  int rc = check_params(p); if (rc) return rc;
  if (n_segments < 0 || n_items < 0 || max_seg_len < 0) return fail(RHO_ERR_INVALID, "negative size");
  if (n_items == 0) return RHO_OK;
  if (!item_first_seg || !y_off || !rec || (n_segments > 0 && (!x || !seg_off || !seg_len || !y)))
    return fail(RHO_ERR_INVALID, "NULL device pointer");
  const Derived d = derive(*p);
  Workspace ws;
  rc = carve(workspace, ws_bytes, n_segments, n_items, max_seg_len, d, &ws, nullptr); if (rc) return rc;
  cudaError_t e = launch_join(x, seg_off, seg_len, n_segments, max_seg_len, item_first_seg, n_items, max_item_len,
                              d, y, y_off, rec, seg_info, ws, (cudaStream_t)stream, &h->lc);
  return e == cudaSuccess ? RHO_OK : cuda_fail(e, "join");
}

int rho_b200_remove_dc(rho_handle* h, float* x, int64_t n, float* dc_out, void* workspace, size_t ws_bytes, void* stream) {
  RHO_ON_DEVICE(h);
  if (n < 0) return fail(RHO_ERR_INVALID, "negative size");
  if (n == 0) return RHO_OK;                       // base_tts.py:396-397
  if (!x) return fail(RHO_ERR_INVALID, "NULL device pointer");
  if (!workspace || ws_bytes < 256) return fail(RHO_ERR_WORKSPACE, "workspace too small: need 256 bytes");
  cudaError_t e = launch_remove_dc(x, n, dc_out, (double*)workspace, (cudaStream_t)stream, &h->lc);
  return e == cudaSuccess ? RHO_OK : cuda_fail(e, "remove_dc");
}

int rho_b200_apply_fades(rho_handle* h, float* x, int64_t n, int fade_in, int fade_out, const rho_params* p, void* stream) {
  RHO_ON_DEVICE(h);
  int rc = check_params(p); if (rc) return rc;
  if (n < 0) return fail(RHO_ERR_INVALID, "negative size");
  if (n == 0) return RHO_OK;
  if (!x) return fail(RHO_ERR_INVALID, "NULL device pointer");
  const Derived d = derive(*p);
  cudaError_t e = launch_apply_fades(x, n, d.fade, fade_in, fade_out, (cudaStream_t)stream, &h->lc);
  return e == cudaSuccess ? RHO_OK : cuda_fail(e, "apply_fades");
}

int rho_b200_sound_decay(rho_handle* h, const float* x, int64_t n, const rho_params* p, rho_record* rec,
                         void* workspace, size_t ws_bytes, void* stream) {
  RHO_ON_DEVICE(h);
  int rc = check_params(p); if (rc) return rc;
  if (n < 0 || n > INT32_MAX) return fail(RHO_ERR_INVALID, "size out of range");
  if (!rec || (n > 0 && !x)) return fail(RHO_ERR_INVALID, "NULL device pointer");
  if (!workspace || ws_bytes < 256) return fail(RHO_ERR_WORKSPACE, "workspace too small: need 256 bytes");
  cudaError_t e = launch_sound_decay(x, n, p->decay_thr, rec, (double*)workspace, (cudaStream_t)stream, &h->lc);
  return e == cudaSuccess ? RHO_OK : cuda_fail(e, "sound_decay");
}

int rho_b200_resample3to2(rho_handle* h, const float* x, const int64_t* off, const int32_t* len,
                          int len_stride_bytes, int n, int64_t max_len, float* y, const int64_t* y_off,
                          int32_t* y_len, void* stream) {
  RHO_ON_DEVICE(h);
  if (n < 0 || max_len < 0) return fail(RHO_ERR_INVALID, "negative size");
  if (n == 0) return RHO_OK;
  if (!x || !off || !len || !y || !y_off) return fail(RHO_ERR_INVALID, "NULL device pointer");
  cudaError_t e = launch_resample3to2(x, off, len, len_stride_bytes, n, max_len, y, y_off, y_len,
                                      (cudaStream_t)stream, &h->lc);
  return e == cudaSuccess ? RHO_OK : cuda_fail(e, "resample3to2");
}

static long long gcd_ll(long long a, long long b) { while (b) { const long long t = a % b; a = b; b = t; } return a; }

int64_t rho_b200_resample_out_len(int64_t len, int orig_freq, int new_freq) {
  if (len <= 0 || orig_freq <= 0 || new_freq <= 0) return 0;
  const long long g = gcd_ll(orig_freq, new_freq);
  const long long o = orig_freq / g, n = new_freq / g;
  return (int64_t)((n * len + o - 1) / o);
}

int rho_b200_resample(rho_handle* h, const float* x, const int64_t* off, const int32_t* len, int len_stride_bytes,
                      int n, int64_t max_len, int orig_freq, int new_freq, float* y, const int64_t* y_off,
                      int32_t* y_len, void* stream) {
  RHO_ON_DEVICE(h);
  if (n < 0 || max_len < 0) return fail(RHO_ERR_INVALID, "negative size");
  if (orig_freq <= 0 || new_freq <= 0) return fail(RHO_ERR_INVALID, "sample rates must be positive");
  if (orig_freq == new_freq) return fail(RHO_ERR_INVALID, "orig_freq == new_freq: torchaudio returns the input unchanged, so does the caller");
  if (n == 0 || max_len == 0) return RHO_OK;
  if (!x || !off || !len || !y || !y_off) return fail(RHO_ERR_INVALID, "NULL device pointer");
  const long long g = gcd_ll(orig_freq, new_freq);
  const int orig = (int)(orig_freq / g), nw = (int)(new_freq / g);
  const int width = host_resample_width(orig, nw);
  const int K = 2 * width + orig;
  if ((long long)K + orig > 8192) return fail(RHO_ERR_INVALID, "ratio %d:%d needs %d taps per phase: not supported", orig, nw, K);
  float* taps = nullptr;
  {
    // first use of a ratio builds and uploads its tap table (the one place this call allocates)
    std::lock_guard<std::mutex> lock(h->mu);
    const uint64_t key = ((uint64_t)(uint32_t)orig << 32) | (uint32_t)nw;
    auto it = h->resample_taps.find(key);
    if (it == h->resample_taps.end()) {
      std::vector<float> t((size_t)nw * K);
      host_resample_taps_general(orig, nw, t.data());
      cudaError_t e = dev_upload(h, &taps, t.data(), t.size());
      if (e != cudaSuccess) return cuda_fail(e, "resample taps");
      h->resample_taps[key] = taps;
    } else {
      taps = it->second;
    }
  }
  cudaError_t e = launch_resample_general(x, off, len, len_stride_bytes, n, max_len, orig, nw, width, taps, y, y_off,
                                          y_len, (cudaStream_t)stream, &h->lc);
  return e == cudaSuccess ? RHO_OK : cuda_fail(e, "resample");
}

static double pitch_rate(double n_steps) { return std::pow(2.0, -n_steps / 12.0); }     // functional.py:1638

size_t rho_b200_pitch_workspace_bytes(int n, int64_t max_len, double n_steps) {
  if (n <= 0 || max_len <= 0) return 0;
  return pitch_plan(n, max_len, pitch_rate(n_steps)).total;
}

int rho_b200_pitch_shift(rho_handle* h, const float* x, const int64_t* off, const int32_t* len, int len_stride_bytes,
                         int n, int64_t min_len, int64_t max_len, int sample_rate, double n_steps, int arange_vec,
                         float* y, const int64_t* y_off, void* workspace, size_t ws_bytes, void* stream) {
  RHO_ON_DEVICE(h);
  if (n < 0 || max_len < 0 || min_len > max_len) return fail(RHO_ERR_INVALID, "bad sizes");
  if (sample_rate <= 0) return fail(RHO_ERR_INVALID, "sample_rate must be positive");
  if (!(std::fabs(n_steps) <= 48.0)) return fail(RHO_ERR_INVALID, "n_steps must be within +-48 semitones, got %g", n_steps);
  if (n_steps == 0.0) return fail(RHO_ERR_INVALID, "n_steps == 0: the reference skips the call (base_tts.py:639)");
  if (n == 0) return RHO_OK;
  // torch.stft(center=True, pad_mode='reflect') refuses clips of <= n_fft / 2 samples
  if (min_len <= 256) return fail(RHO_ERR_INVALID, "pitch shift needs clips longer than 256 samples (reflect padding of the 512-point STFT), shortest is %lld", (long long)min_len);
  if (n > 65535) return fail(RHO_ERR_INVALID, "at most 65535 clips per call");
  if (!x || !off || !len || !y || !y_off) return fail(RHO_ERR_INVALID, "NULL device pointer");
  const double rate = pitch_rate(n_steps);
  const PitchPlan pl = pitch_plan(n, max_len, rate);
  if (!workspace || ws_bytes < pl.total) return fail(RHO_ERR_WORKSPACE, "workspace too small: have %zu, need %zu", ws_bytes, pl.total);
  if (((uintptr_t)workspace) & 255u) return fail(RHO_ERR_INVALID, "workspace must be 256-byte aligned");
  const int orig_freq = (int)((double)sample_rate / rate);                     // functional.py:1639
  if (orig_freq <= 0) return fail(RHO_ERR_INVALID, "int(sample_rate / rate) is not positive");
  const long long g = gcd_ll(orig_freq, sample_rate);
  const int orig = (int)(orig_freq / g), nw = (int)(sample_rate / g);
  const int width = host_resample_width(orig, nw);
  const int W = (2 * width + 2 + 3) / 4 * 4;          // rows of the tap table stay 16-byte aligned
  rho_handle::WinTaps wt{nullptr, nullptr};
  {
    std::lock_guard<std::mutex> lock(h->mu);
    if (!h->pitch_tb_ok) {
      std::vector<float> w256(512), w512(514), hn(512), pv(257);
      host_pitch_tables(w256.data(), w512.data(), hn.data(), pv.data());
      cudaError_t e;
      if ((e = dev_upload(h, (float**)&h->pitch_tb.w256, w256.data(), w256.size())) != cudaSuccess ||
          (e = dev_upload(h, (float**)&h->pitch_tb.w512, w512.data(), w512.size())) != cudaSuccess ||
          (e = dev_upload(h, &h->pitch_tb.hann, hn.data(), hn.size())) != cudaSuccess ||
          (e = dev_upload(h, &h->pitch_tb.padv, pv.data(), pv.size())) != cudaSuccess)
        return cuda_fail(e, "pitch tables");
      h->pitch_tb_ok = true;
    }
    const uint64_t key = ((uint64_t)(uint32_t)orig << 32) | (uint32_t)nw;
    auto it = h->windowed_taps.find(key);
    if (it == h->windowed_taps.end()) {
      std::vector<float> t((size_t)nw * W);
      std::vector<int> lo((size_t)nw);
      if (orig == nw) {
        // int(sample_rate / rate) == sample_rate (a shift of a few thousandths of a semitone): torchaudio's resample
        // returns its input unchanged (functional.py:1417-1418) -- an identity tap, not the 0.99-rolloff low-pass
        std::fill(t.begin(), t.end(), 0.f);
        t[0] = 1.f;
        lo[0] = width;
      } else {
        host_resample_taps_windowed(orig, nw, width, W, t.data(), lo.data());
      }
      cudaError_t e;
      if ((e = dev_upload(h, &wt.taps, t.data(), t.size())) != cudaSuccess ||
          (e = dev_upload(h, &wt.ilo, lo.data(), lo.size())) != cudaSuccess)
        return cuda_fail(e, "windowed resample taps");
      h->windowed_taps[key] = wt;
    } else {
      wt = it->second;
    }
  }
  cudaError_t e = launch_pitch_shift(h->pitch_tb, x, off, len, len_stride_bytes, n, max_len, rate, arange_vec, orig, nw,
                                     width, W, wt.taps, wt.ilo, y, y_off, workspace, (cudaStream_t)stream, &h->lc);
  return e == cudaSuccess ? RHO_OK : cuda_fail(e, "pitch_shift");
}

size_t rho_b200_mfcc_workspace_bytes(int n, int64_t max_len) { return mfcc_workspace_bytes(n, max_len); }

int rho_b200_mfcc_stats(rho_handle* h, const float* x16, const int64_t* off, const int32_t* len, int len_stride_bytes,
                        int n, int64_t max_len, float* out, void* workspace, size_t ws_bytes, void* stream) {
  RHO_ON_DEVICE(h);
  if (n < 0 || max_len < 0) return fail(RHO_ERR_INVALID, "negative size");
  if (n == 0) return RHO_OK;
  if (n > 65535) return fail(RHO_ERR_INVALID, "at most 65535 clips per call");
  if (!x16 || !off || !len || !out) return fail(RHO_ERR_INVALID, "NULL device pointer");
  const size_t need = mfcc_workspace_bytes(n, max_len);
  if (!workspace || ws_bytes < need) return fail(RHO_ERR_WORKSPACE, "workspace too small: have %zu, need %zu", ws_bytes, need);
  if (((uintptr_t)workspace) & 255u) return fail(RHO_ERR_INVALID, "workspace must be 256-byte aligned");
  {
    std::lock_guard<std::mutex> lock(h->mu);
    if (!h->mfcc_tb_ok) {
      std::vector<float> hn(2048), w1(2048), w2(2050), dct(13 * 128), w256(512), w512(514), hn512(512), pv(257);
      host_mfcc_tables(hn.data(), w1.data(), w2.data(), dct.data());
      host_pitch_tables(w256.data(), w512.data(), hn512.data(), pv.data());
      std::vector<float> dense((size_t)128 * 1025);
      host_mel_filterbank_bins(128, 1025, dense.data());
      std::vector<int> lo(128), cnt(128), wofs(128);
      std::vector<float> w;
      for (int m = 0; m < 128; ++m) {
        int a = 0, b = 1025;
        while (a < 1025 && dense[(size_t)m * 1025 + a] == 0.f) ++a;
        while (b > a && dense[(size_t)m * 1025 + b - 1] == 0.f) --b;
        lo[m] = a; cnt[m] = b - a; wofs[m] = (int)w.size();
        for (int k = a; k < b; ++k) w.push_back(dense[(size_t)m * 1025 + k]);
      }
      if (w.empty()) w.push_back(0.f);
      MfccTables& t = h->mfcc_tb;
      t.mel_nnz = (int)w.size();
      cudaError_t e;
      if ((e = dev_upload(h, (float**)&t.w256, w256.data(), w256.size())) != cudaSuccess ||
          (e = dev_upload(h, (float**)&t.w1024, w1.data(), w1.size())) != cudaSuccess ||
          (e = dev_upload(h, (float**)&t.w2048, w2.data(), w2.size())) != cudaSuccess ||
          (e = dev_upload(h, &t.hann, hn.data(), hn.size())) != cudaSuccess ||
          (e = dev_upload(h, &t.mel_lo, lo.data(), lo.size())) != cudaSuccess ||
          (e = dev_upload(h, &t.mel_cnt, cnt.data(), cnt.size())) != cudaSuccess ||
          (e = dev_upload(h, &t.mel_wofs, wofs.data(), wofs.size())) != cudaSuccess ||
          (e = dev_upload(h, &t.mel_w, w.data(), w.size())) != cudaSuccess ||
          (e = dev_upload(h, &t.dct, dct.data(), dct.size())) != cudaSuccess)
        return cuda_fail(e, "mfcc tables");
      h->mfcc_tb_ok = true;
    }
  }
  cudaError_t e = launch_mfcc_stats(h->mfcc_tb, x16, off, len, len_stride_bytes, n, max_len, out, workspace,
                                    (cudaStream_t)stream, &h->lc);
  return e == cudaSuccess ? RHO_OK : cuda_fail(e, "mfcc_stats");
}

// ---- resemblyzer front end (spk.cu)
int rho_b200_spk_slices(int64_t n_samples, int frame_step, double min_coverage, int64_t* padded_len) {
  if (n_samples < 0 || frame_step <= 0 || frame_step > SPK_PART_FRAMES || !(min_coverage > 0.0 && min_coverage <= 1.0))
    return fail(RHO_ERR_INVALID, "spk_slices: n_samples >= 0, 0 < frame_step <= 160, 0 < min_coverage <= 1");
  long long padded = 0;
  const int count = spk_slices(n_samples, frame_step, min_coverage, &padded);
  if (padded_len) *padded_len = padded;
  return count;
}

int rho_b200_normalize_volume(rho_handle* h, const float* x, const int64_t* off, const int32_t* len, int len_stride_bytes,
                              int n, int64_t max_len, float target_dbfs, int mode, float* y, const int64_t* y_off,
                              float* gain, void* workspace, size_t ws_bytes, void* stream) {
  RHO_ON_DEVICE(h);
  if (n < 0 || max_len < 0) return fail(RHO_ERR_INVALID, "negative size");
  if (mode < 0 || mode > 2) return fail(RHO_ERR_INVALID, "mode: 0 = always, 1 = increase only, 2 = decrease only");
  if (n == 0) return RHO_OK;
  if (!x || !off || !len || !gain || (y && !y_off)) return fail(RHO_ERR_INVALID, "NULL device pointer");
  if (!workspace || ws_bytes < sizeof(double) * (size_t)n) return fail(RHO_ERR_WORKSPACE, "workspace too small: have %zu, need %zu", ws_bytes, sizeof(double) * (size_t)n);
  if (((uintptr_t)workspace) & 7u) return fail(RHO_ERR_INVALID, "workspace must be 8-byte aligned");
  cudaError_t e = launch_spk_normalize(x, off, len, len_stride_bytes, n, max_len, target_dbfs, mode, y, y_off, gain,
                                       (double*)workspace, (cudaStream_t)stream, &h->lc);
  return e == cudaSuccess ? RHO_OK : cuda_fail(e, "normalize_volume");
}

int rho_b200_spk_mel(rho_handle* h, const float* x16, const int64_t* off, const int32_t* len, int len_stride_bytes, int n,
                     int64_t max_len, int frame_step, double min_coverage, unsigned flags, const float* gain, float* mel,
                     const int64_t* frame_off, float* partials, const int32_t* part_off, void* stream) {
  RHO_ON_DEVICE(h);
  if (n < 0 || max_len < 0) return fail(RHO_ERR_INVALID, "negative size");
  if (frame_step <= 0 || frame_step > SPK_PART_FRAMES || !(min_coverage > 0.0 && min_coverage <= 1.0))
    return fail(RHO_ERR_INVALID, "spk_mel: 0 < frame_step <= 160, 0 < min_coverage <= 1");
  if (n == 0) return RHO_OK;
  if (!x16 || !off || !len) return fail(RHO_ERR_INVALID, "NULL device pointer");
  if (!mel && !partials) return fail(RHO_ERR_INVALID, "neither mel nor partials asked for");
  if ((mel && !frame_off) || (partials && !part_off)) return fail(RHO_ERR_INVALID, "an output without its offsets");
  cudaError_t e = launch_spk_mel(h->tb, x16, off, len, len_stride_bytes, n, max_len, frame_step, min_coverage,
                                 (flags & RHO_SPK_PAD_TO_SLICES) != 0, gain, mel, frame_off, partials, part_off,
                                 (cudaStream_t)stream, &h->lc);
  return e == cudaSuccess ? RHO_OK : cuda_fail(e, "spk_mel");
}

int rho_b200_spk_pool(rho_handle* h, const float* partial_embeds, const int32_t* part_off, int n, int dim, float* out,
                      void* stream) {
  RHO_ON_DEVICE(h);
  if (n < 0 || dim <= 0) return fail(RHO_ERR_INVALID, "bad size");
  if (n == 0) return RHO_OK;
  if (!partial_embeds || !part_off || !out) return fail(RHO_ERR_INVALID, "NULL device pointer");
  cudaError_t e = launch_spk_pool(partial_embeds, part_off, n, dim, out, (cudaStream_t)stream, &h->lc);
  return e == cudaSuccess ? RHO_OK : cuda_fail(e, "spk_pool");
}

int rho_b200_logmel(rho_handle* h, const float* x16, const int64_t* off, const int32_t* len16, int n,
                    int64_t max_len16, int n_mels, int pad_frames, float* mel, int64_t mel_stride_frames,
                    int32_t* n_frames, void* workspace, size_t ws_bytes, void* stream) {
  RHO_ON_DEVICE(h);
  if (n_mels != 80 && n_mels != 128) return fail(RHO_ERR_INVALID, "n_mels must be 80 or 128, got %d", n_mels);
  if (pad_frames != 0 && pad_frames != MEL_PAD_FRAMES) return fail(RHO_ERR_INVALID, "pad_frames must be 0 or 3000");
  if (n < 0 || max_len16 < 0) return fail(RHO_ERR_INVALID, "negative size");
  if (n == 0) return RHO_OK;
  if (!x16 || !off || !len16 || !mel) return fail(RHO_ERR_INVALID, "NULL device pointer");
  if (pad_frames == 0 && mel_stride_frames < max_len16 / HOP16) return fail(RHO_ERR_INVALID, "mel_stride_frames too small");
  if (pad_frames > 0 && mel_stride_frames < pad_frames) return fail(RHO_ERR_INVALID, "mel_stride_frames too small");
  const size_t need = align_up(sizeof(int) * (size_t)n, 256);
  if (!workspace || ws_bytes < need) return fail(RHO_ERR_WORKSPACE, "workspace too small: have %zu, need %zu", ws_bytes, need);
  cudaError_t e = launch_logmel(h->tb, x16, off, len16, n, max_len16, n_mels, pad_frames, mel, mel_stride_frames,
                                n_frames, (int*)workspace, (cudaStream_t)stream, &h->lc);
  return e == cudaSuccess ? RHO_OK : cuda_fail(e, "logmel");
}

int rho_b200_mel_project(rho_handle* h, const float* power, int64_t n_frames, int64_t ld_power, int n_mels,
                         float* mel, int64_t ld_mel, int64_t frames_per_item, int64_t item_stride, void* stream) {
  RHO_ON_DEVICE(h);
  if (n_mels != 80 && n_mels != 128) return fail(RHO_ERR_INVALID, "n_mels must be 80 or 128, got %d", n_mels);
  if (n_frames < 0) return fail(RHO_ERR_INVALID, "negative size");
  if (n_frames == 0) return RHO_OK;
  if (!power || !mel) return fail(RHO_ERR_INVALID, "NULL device pointer");
  if (ld_power < N_BINS || ld_power % 4 != 0 || (((uintptr_t)power) & 15u))
    return fail(RHO_ERR_INVALID, "power rows must be 16-byte aligned: ld_power %% 4 == 0, ld_power >= 201 (TMA)");
  if (frames_per_item <= 0) { frames_per_item = n_frames; item_stride = 0; }
  if (ld_mel < frames_per_item) return fail(RHO_ERR_INVALID, "ld_mel too small");
  if (frames_per_item < n_frames && item_stride < (int64_t)n_mels * ld_mel) return fail(RHO_ERR_INVALID, "item_stride too small");
  cudaError_t e = launch_mel_gemm(h->tb, power, n_frames, ld_power, n_mels, mel, ld_mel, frames_per_item, item_stride, h->sm_count,
                      (cudaStream_t)stream, &h->lc);
  return e == cudaSuccess ? RHO_OK : cuda_fail(e, "mel_gemm");
}

int rho_b200_stft_power_tc(rho_handle* h, const float* x16, const int64_t* off, const int32_t* len16, int pad_frames,
                           const int32_t* tiles, int n_tiles, float* power, int64_t ld_power, void* stream) {
  RHO_ON_DEVICE(h);
  if (pad_frames != 0 && pad_frames != MEL_PAD_FRAMES) return fail(RHO_ERR_INVALID, "pad_frames must be 0 or 3000");
  if (n_tiles < 0) return fail(RHO_ERR_INVALID, "negative size");
  if (n_tiles == 0) return RHO_OK;
  if (!x16 || !off || !len16 || !tiles || !power) return fail(RHO_ERR_INVALID, "NULL device pointer");
  if (ld_power < N_BINS) return fail(RHO_ERR_INVALID, "ld_power must be >= 201");
  if (((uintptr_t)tiles) & 15u) return fail(RHO_ERR_INVALID, "tiles must be 16-byte aligned");
  {
    std::lock_guard<std::mutex> lock(h->mu);
    if (!h->stft_tc_tables) {
      std::vector<unsigned char> t(stft_tc_table_bytes());
      host_stft_tc_tables(t.data());
      cudaError_t e = dev_upload(h, &h->stft_tc_tables, t.data(), t.size());
      if (e != cudaSuccess) { h->stft_tc_tables = nullptr; return cuda_fail(e, "stft tables"); }
    }
  }
  cudaError_t e = launch_stft_tc(h->stft_tc_tables, x16, off, len16, pad_frames, tiles, n_tiles, power, ld_power,
                                 h->sm_count, (cudaStream_t)stream, &h->lc);
  return e == cudaSuccess ? RHO_OK : cuda_fail(e, "stft_power_tc");
}

int rho_b200_sound_decay_batch(rho_handle* h, const float* y, const int64_t* off, const int32_t* len,
                               int len_stride_bytes, int n, int64_t max_len, const rho_params* p, rho_record* rec,
                               void* workspace, size_t ws_bytes, void* stream) {
  RHO_ON_DEVICE(h);
  int rc = check_params(p); if (rc) return rc;
  if (n < 0 || max_len < 0) return fail(RHO_ERR_INVALID, "negative size");
  if (n == 0) return RHO_OK;
  if (!y || !off || !len || !rec) return fail(RHO_ERR_INVALID, "NULL device pointer");
  const size_t need = 2 * sizeof(double) * (size_t)n;
  if (!workspace || ws_bytes < need) return fail(RHO_ERR_WORKSPACE, "workspace too small: have %zu, need %zu", ws_bytes, need);
  if (((uintptr_t)workspace) & 7u) return fail(RHO_ERR_INVALID, "workspace must be 8-byte aligned");
  const Derived d = derive(*p);
  cudaError_t e = launch_sound_decay_batch(y, off, len, len_stride_bytes, n, max_len, d.decay_thr, rec,
                                           (double*)workspace, (cudaStream_t)stream, &h->lc);
  return e == cudaSuccess ? RHO_OK : cuda_fail(e, "sound_decay_batch");
}

int rho_b200_pcm16(rho_handle* h, const float* y, const int64_t* off, const int32_t* len, int len_stride_bytes, int n,
                   int64_t max_len, int16_t* out, const int64_t* out_off, void* stream) {
  RHO_ON_DEVICE(h);
  if (n < 0 || max_len < 0) return fail(RHO_ERR_INVALID, "negative size");
  if (n == 0 || max_len == 0) return RHO_OK;
  if (n > 65535) return fail(RHO_ERR_INVALID, "at most 65535 clips per call");
  if (!y || !off || !len || !out || !out_off) return fail(RHO_ERR_INVALID, "NULL device pointer");
  cudaError_t e = launch_pcm16(y, off, len, len_stride_bytes, n, max_len, out, out_off, (cudaStream_t)stream, &h->lc);
  return e == cudaSuccess ? RHO_OK : cuda_fail(e, "pcm16");
}

size_t rho_b200_qwen_workspace_bytes(int n, int64_t max_len, int sr) { return qwen_workspace_bytes(n, max_len, sr); }

int rho_b200_qwen_postprocess(rho_handle* h, const float* x, const int64_t* off, const int32_t* len,
                              int len_stride_bytes, int n, int64_t max_len, int sr, float* y, const int64_t* y_off,
                              void* workspace, size_t ws_bytes, void* stream) {
  RHO_ON_DEVICE(h);
  if (n < 0 || max_len < 0) return fail(RHO_ERR_INVALID, "negative size");
  if (sr <= 0) return fail(RHO_ERR_INVALID, "sample rate must be positive, got %d", sr);
  if (n == 0 || max_len == 0) return RHO_OK;
  if (!x || !off || !len || !y || !y_off) return fail(RHO_ERR_INVALID, "NULL device pointer");
  if (max_len / (2LL * sr) > 64) return fail(RHO_ERR_INVALID, "clips longer than 64 windows (128 s) are not supported");
  const size_t need = qwen_workspace_bytes(n, max_len, sr);
  if (!workspace || ws_bytes < need) return fail(RHO_ERR_WORKSPACE, "workspace too small: have %zu, need %zu", ws_bytes, need);
  cudaError_t e = launch_qwen_postprocess(x, off, len, len_stride_bytes, n, max_len, sr, y, y_off, workspace,
                                          (cudaStream_t)stream, &h->lc);
  return e == cudaSuccess ? RHO_OK : cuda_fail(e, "qwen_postprocess");
}

int rho_b200_cosine(rho_handle* h, const float* emb, const float* ref, int n, int dim, float* out,
                    int out_stride_bytes, void* stream) {
  RHO_ON_DEVICE(h);
  if (n < 0 || dim <= 0) return fail(RHO_ERR_INVALID, "bad size");
  if (n == 0) return RHO_OK;
  if (!emb || !ref || !out) return fail(RHO_ERR_INVALID, "NULL device pointer");
  cudaError_t e = launch_cosine(emb, ref, n, dim, out, out_stride_bytes, (cudaStream_t)stream, &h->lc);
  return e == cudaSuccess ? RHO_OK : cuda_fail(e, "cosine");
}

// frames of the 30 s (pad_frames) window that can see signal for an item of at most max_item_len 24 kHz samples,
// rounded up to whole 128-bit pieces: the shortest row a compact feature tensor may have
static int64_t compact_frames(int64_t max_item_len, int pad_frames) {
  const int64_t n16 = (2 * max_item_len + 2) / 3;
  const int64_t nv = std::min<int64_t>(n16, (int64_t)pad_frames * HOP16);
  int64_t t = (nv + N_FFT / 2 + HOP16 - 1) / HOP16;
  t = std::max<int64_t>(t, 2);
  t = std::min<int64_t>((t + 3) / 4 * 4, pad_frames);
  return t;
}

int64_t rho_b200_compact_frames(int64_t max_item_len, int pad_frames) {
  return pad_frames > 0 ? compact_frames(max_item_len, pad_frames) : ((2 * max_item_len + 2) / 3) / HOP16;
}

int rho_b200_validate(rho_handle* h, const float* x, const int64_t* seg_off, const int32_t* seg_len,
                      int n_segments, int64_t max_seg_len, const int32_t* item_first_seg, int n_items,
                      int64_t max_item_len, const rho_params* p, float* y, const int64_t* y_off,
                      int n_mels, int pad_frames, float* mel, int64_t mel_stride_frames, float* pad_value,
                      const float* emb, const float* ref_emb, int emb_dim, rho_record* rec, float* scratch16,
                      uint32_t flags, void* workspace, size_t ws_bytes, void* stream) {
  RHO_ON_DEVICE(h);
  int rc = check_params(p); if (rc) return rc;
  if (p->sr != 24000) return fail(RHO_ERR_INVALID, "validate needs 24 kHz input (3:2 resampler), got %d", p->sr);
  if (n_mels != 80 && n_mels != 128) return fail(RHO_ERR_INVALID, "n_mels must be 80 or 128, got %d", n_mels);
  if (pad_frames != 0 && pad_frames != MEL_PAD_FRAMES) return fail(RHO_ERR_INVALID, "pad_frames must be 0 or 3000");
  if (n_items <= 0) return n_items == 0 ? RHO_OK : fail(RHO_ERR_INVALID, "negative size");
  if (n_segments < 0 || max_seg_len < 0) return fail(RHO_ERR_INVALID, "negative size");
  if (!mel || !rec || !item_first_seg || !y_off || (n_segments > 0 && (!x || !seg_off || !seg_len || !y)))
    return fail(RHO_ERR_INVALID, "NULL device pointer");
  const int64_t max16 = (2 * max_item_len + 2) / 3;
  const bool compact = (flags & RHO_V_COMPACT_PAD) && pad_frames > 0;
  int fill_to = pad_frames;
  if (pad_frames == 0 && mel_stride_frames < max16 / HOP16) return fail(RHO_ERR_INVALID, "mel_stride_frames too small");
  if (compact) {
    const int64_t need = compact_frames(max_item_len, pad_frames);
    if (mel_stride_frames < need || mel_stride_frames % 4 != 0)
      return fail(RHO_ERR_INVALID, "compact rows need mel_stride_frames >= %lld and a multiple of 4, got %lld",
                  (long long)need, (long long)mel_stride_frames);
    fill_to = (int)std::min<int64_t>(mel_stride_frames, pad_frames);
  } else if (pad_frames > 0 && mel_stride_frames < pad_frames) {
    return fail(RHO_ERR_INVALID, "mel_stride_frames too small");
  }
  const Derived d = derive(*p);
  Workspace ws;
  rc = carve(workspace, ws_bytes, n_segments, n_items, max_seg_len, d, &ws, nullptr); if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e;
  // multi-GPU: this call's records also go to every rank's gathered buffer, parity = epoch & 1 (exchange.cu)
  RecordPeers peers{};
  {
    std::lock_guard<std::mutex> lock(h->mu);
    Exchange& X = h->xch;
    if (X.connected && !(flags & RHO_V_NO_FUSION)) {
      if (n_items > X.n_per_rank) return fail(RHO_ERR_INVALID, "exchange sized for %lld items per rank, call has %d", (long long)X.n_per_rank, n_items);
      ++X.epoch;
      peers.n = X.world; peers.rank = X.rank; peers.epoch = (unsigned)X.epoch; peers.n_items = n_items;
      peers.slot = ((long long)(X.epoch & 1) * X.world + X.rank) * X.n_per_rank;
      char* fl = (char*)X.buf + align_up(X.rec_bytes, 256);
      peers.done = (int*)(fl + 192);
      for (int q = 0; q < X.world; ++q) {
        peers.sink[q] = (rho_record*)X.peer[q];
        peers.flag[q] = (unsigned*)((char*)X.peer[q] + align_up(X.rec_bytes, 256));
      }
    }
  }
  const bool one_seg = (flags & RHO_V_ONE_SEGMENT_ITEMS) && n_segments == n_items;
  if (!(flags & RHO_V_NO_FUSION)) {
    // one-segment items:  init -> scan -> bounds / DC / plan -> ONE kernel for apply + resample + log-mel -> clamp +
    //                     records (decay decision, cosine): five launches per batch
    // joined items:       init -> scan -> bounds / DC -> plan -> the same kernel, which also joins: a batch inside one
    //                     segment is the one-segment case, the joints are computed sample by sample -> clamp + records
    //                     (RHO_V_GATHER_FIRST: the round-2 path -- k_gather writes y, the kernel reads it back)
    const bool gather_first = !one_seg && (flags & RHO_V_GATHER_FIRST);
    e = launch_join(x, seg_off, seg_len, n_segments, max_seg_len, item_first_seg, n_items, max_item_len, d, y, y_off,
                    rec, nullptr, ws, st, &h->lc,
                    one_seg ? (JOIN_PREPARE | JOIN_INIT_FEATURES | JOIN_ONE_SEG_ITEMS)
                            : gather_first ? (JOIN_PREPARE | JOIN_GATHER | JOIN_INIT_FEATURES)
                                           : (JOIN_PREPARE | JOIN_INIT_FEATURES));
    if (e != cudaSuccess) return cuda_fail(e, "join prepare");
    e = launch_fused_features(h->tb, x, seg_off, ws, item_first_seg, n_items, max_item_len, d, y, y_off, n_mels,
                              pad_frames, mel, mel_stride_frames, h->sm_count, st, &h->lc,
                              one_seg ? 0 : gather_first ? 1 : 2, fill_to);
    if (e != cudaSuccess) return cuda_fail(e, "fused features");
    // the Whisper clamp of the frames with signal (and the constant of the zero-padding frames unless the fused kernel
    // wrote it: fill_done); its first warp per clip assembles the record
    const FinalizeArgs fin{ws.seg, ws.item, item_first_seg, d.decay_thr, rec, emb, ref_emb, emb_dim, peers};
    e = launch_logmel_norm(ws.len16, n_items, n_mels, pad_frames, mel, mel_stride_frames, ws.clip_max, st, &h->lc,
                           fused_inline_norm(), &fin, fill_to, pad_value);
    if (e != cudaSuccess) return cuda_fail(e, "logmel norm");
  } else {
    if (!scratch16) return fail(RHO_ERR_INVALID, "scratch16 is NULL (needed by the kernel-per-stage path)");
    e = launch_join(x, seg_off, seg_len, n_segments, max_seg_len, item_first_seg, n_items, max_item_len, d, y, y_off,
                    rec, nullptr, ws, st, &h->lc, JOIN_ALL);
    if (e != cudaSuccess) return cuda_fail(e, "join");
    // 16 kHz intermediate lives at the same offsets as y (it is 2/3 as long)
    e = launch_resample3to2(y, y_off, &rec[0].out_len, (int)sizeof(rho_record), n_items, max_item_len,
                            scratch16, y_off, ws.len16, st, &h->lc);
    if (e != cudaSuccess) return cuda_fail(e, "resample3to2");
    e = launch_logmel(h->tb, scratch16, y_off, ws.len16, n_items, max16, n_mels, pad_frames, mel, mel_stride_frames,
                      nullptr, ws.clip_max, st, &h->lc, fill_to, pad_value);
    if (e != cudaSuccess) return cuda_fail(e, "logmel");
    if (emb && ref_emb) {
      e = launch_cosine(emb, ref_emb, n_items, emb_dim, &rec[0].cosine, (int)sizeof(rho_record), st, &h->lc);
      if (e != cudaSuccess) return cuda_fail(e, "cosine");
    }
  }
  return RHO_OK;
}

}  // extern "C"
