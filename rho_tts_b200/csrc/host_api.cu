// HOST entry points of librho_b200 (include/rho_b200.h): the calls an embedder without torch makes.
//   rho_b200_validate_host_ragged   ragged segments / items in host memory -> join, decay check, features
//   rho_b200_validate_host          fixed-length clips, one clip per item (a wrapper of the above)
// The batch is cut into chunks of whole items (up to ~123 MB of samples, smaller while the pipeline fills and drains); four
// chunk slots in a device arena keep the copy-in of chunks k+1 / k+2, the kernels of chunk k and the copy-out of chunk k-1
// in flight at once, each on its own stream.
//
// What crosses PCIe: every input sample once, every processed sample once, one 48-byte record per item and, of the
// Whisper features, only the frames that can see signal.  The 30 s Whisper window of a 10 s clip is 2/3 zero padding,
// and all of those frames hold ONE value per clip (feature_extraction_whisper.py:296-303: padding, then log10(1e-10)
// clamped at max - 8): the device keeps compact rows (rho_b200_validate with RHO_V_COMPACT_PAD), the rows are
// scattered into the caller's [item][n_mels][3000] layout by a strided copy, and the constant tail of every row is
// written into the caller's buffer by host threads from the 4-byte per-item value, overlapped with the copies of the
// following chunks.  (Round 1 shipped the constants over the link: a third of a saturated link.)
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include "handle.h"

using namespace rho;

namespace rho {

void host_fill_f32(float* p, int64_t n, float v);     // hostfill.cpp: non-temporal fill, widest vectors the CPU has

constexpr int HC_MAX_SLOTS = 8;    // chunk slots in the device arena (RHO_HOST_SLOTS, default 4: copy-in may run two chunks ahead)

struct HostCtx {
  cudaStream_t s_in = nullptr, s_compute = nullptr, s_out = nullptr;
  void* arena = nullptr;      size_t arena_bytes = 0;      // device
  void* pinned = nullptr;     size_t pinned_bytes = 0;     // host staging: per-chunk metadata in, pad values out
  std::vector<cudaEvent_t> ev_out;                          // one per chunk: everything of the chunk is in host memory
  std::vector<cudaEvent_t> ev_small;                        // one per chunk: its records and pad values are (blocking-sync:
                                                            // the fill threads sleep on them)
  cudaEvent_t ev_in[HC_MAX_SLOTS] = {}, ev_done[HC_MAX_SLOTS] = {};
  bool ok = false;
};

void host_ctx_destroy(HostCtx* c) {
  if (!c) return;
  if (c->s_in) cudaStreamDestroy(c->s_in);
  if (c->s_compute) cudaStreamDestroy(c->s_compute);
  if (c->s_out) cudaStreamDestroy(c->s_out);
  if (c->arena) cudaFree(c->arena);
  if (c->pinned) cudaFreeHost(c->pinned);
  for (cudaEvent_t e : c->ev_out) cudaEventDestroy(e);
  for (cudaEvent_t e : c->ev_small) cudaEventDestroy(e);
  for (int s = 0; s < HC_MAX_SLOTS; ++s) {
    if (c->ev_in[s]) cudaEventDestroy(c->ev_in[s]);
    if (c->ev_done[s]) cudaEventDestroy(c->ev_done[s]);
  }
  delete c;
}

namespace {

cudaError_t host_ctx_init(HostCtx* c) {
  cudaError_t e;
  if ((e = cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking)) != cudaSuccess) return e;
  if ((e = cudaStreamCreateWithFlags(&c->s_compute, cudaStreamNonBlocking)) != cudaSuccess) return e;
  if ((e = cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking)) != cudaSuccess) return e;
  for (int s = 0; s < HC_MAX_SLOTS; ++s) {
    if ((e = cudaEventCreateWithFlags(&c->ev_in[s], cudaEventDisableTiming)) != cudaSuccess) return e;
    if ((e = cudaEventCreateWithFlags(&c->ev_done[s], cudaEventDisableTiming)) != cudaSuccess) return e;
  }
  c->ok = true;
  return cudaSuccess;
}

// RAII lease of a context from the handle's pool
struct CtxLease {
  rho_handle* h;
  HostCtx* c = nullptr;
  cudaError_t err = cudaSuccess;
  explicit CtxLease(rho_handle* hh) : h(hh) {
    {
      std::lock_guard<std::mutex> lock(h->mu);
      if (!h->host_ctx_free.empty()) { c = h->host_ctx_free.back(); h->host_ctx_free.pop_back(); }
    }
    if (!c) {
      c = new HostCtx();
      err = host_ctx_init(c);
      if (err != cudaSuccess) { host_ctx_destroy(c); c = nullptr; }
    }
  }
  ~CtxLease() {
    if (!c) return;
    std::lock_guard<std::mutex> lock(h->mu);
    h->host_ctx_free.push_back(c);
  }
};

struct Chunk {
  int i0, i1, s0, s1;                 // items [i0, i1), segments [s0, s1)
  int64_t x_lo, x_hi;                 // host sample range of its segments (contiguous mode)
  int64_t y_lo, y_hi;                 // host sample range of its items' outputs
  int64_t dx_samples, dy_samples;     // device samples it needs (x slot, y slot)
  int64_t max_seg_len, max_item_cap;
  int t_dev;                          // frames per feature row on the device (compact), 0: no features
};

// developer switches for tools/e2e_sweep.py (timing experiments only)
bool debug_skip_fill() { static const bool v = getenv("RHO_HOST_DEBUG_SKIP_FILL") != nullptr; return v; }
// test hook: worker 0 sleeps this long per chunk, as a host with slow memory would (tests/test_gpu_parity.py)
int debug_fill_delay_us() { static const int v = [] { const char* e = getenv("RHO_HOST_DEBUG_FILL_DELAY_US"); return e ? atoi(e) : 0; }(); return v; }

// 0.64 GB of constants per 1.28 GB of copy-out: about 26 GB/s of fill to keep up with a gen-5 link.  Three threads do
// that on most hosts of the pool (tools/e2e_sweep.py, profiles/e2e_sweep_r02*.log) -- but not on all: a call whose fill
// ends well after its last copy (more than 5 % of the call) gives the handle's next calls two more threads, up to 9.
// RHO_HOST_FILL_THREADS pins the count.
constexpr int FILL_THREADS_DEFAULT = 3, FILL_THREADS_MAX = 9;
int fill_threads_env() {
  static const int n = [] { const char* v = getenv("RHO_HOST_FILL_THREADS"); return v ? std::max(1, atoi(v)) : 0; }();
  return n;
}
int fill_threads(rho_handle* h) {
  if (fill_threads_env() > 0) return fill_threads_env();
  const int hint = h->host_fill_threads.load(std::memory_order_relaxed);
  return hint > 0 ? hint : FILL_THREADS_DEFAULT;
}

int host_slots() {
  static const int n = [] {
    const char* v = getenv("RHO_HOST_SLOTS");
    const int k = v ? atoi(v) : 4;
    return std::min(HC_MAX_SLOTS, std::max(2, k));
  }();
  return n;
}

int64_t chunk_budget_samples() {
  static const int64_t n = [] {
    if (const char* v = getenv("RHO_HOST_CHUNK_SAMPLES")) return std::max<int64_t>(1, atoll(v));
    return (int64_t)128 * 240000;             // 123 MB of fp32 samples in the steady state (the ramps start at 1/16 of it)
  }();
  return n;
}

}  // namespace
}  // namespace rho

extern "C" {

int rho_b200_validate_host_ragged(rho_handle* h, const float* x, const int64_t* seg_off, const int32_t* seg_len,
                                  int n_segments, const int32_t* item_first_seg, int n_items, const rho_params* p,
                                  float* y, const int64_t* y_off, int n_mels, int pad_frames, float* mel,
                                  int64_t mel_stride_frames, float* pad_value, const float* emb,
                                  const float* ref_emb, int emb_dim, rho_record* rec) {
  RHO_ON_DEVICE(h);
  if (!p) return fail(RHO_ERR_INVALID, "params is NULL");
  if (p->sr < 8000 || p->sr > 192000) return fail(RHO_ERR_INVALID, "sample rate %d out of range", p->sr);
  if (n_segments < 0 || n_items < 0) return fail(RHO_ERR_INVALID, "negative size");
  if (n_items == 0) return RHO_OK;
  if (!item_first_seg || !y_off || !rec || !y || (n_segments > 0 && (!x || !seg_off || !seg_len)))
    return fail(RHO_ERR_INVALID, "NULL host pointer");
  const bool features = mel != nullptr;
  if (features) {
    if (p->sr != 24000) return fail(RHO_ERR_INVALID, "features need 24 kHz input (3:2 resampler), got %d", p->sr);
    if (n_mels != 80 && n_mels != 128) return fail(RHO_ERR_INVALID, "n_mels must be 80 or 128, got %d", n_mels);
    if (pad_frames != 0 && pad_frames != MEL_PAD_FRAMES) return fail(RHO_ERR_INVALID, "pad_frames must be 0 or 3000");
  }
  if (item_first_seg[0] != 0 || item_first_seg[n_items] != n_segments)
    return fail(RHO_ERR_LAYOUT, "item_first_seg must run from 0 to n_segments");
  const bool have_emb = emb && ref_emb && emb_dim > 0;
  const Derived d = derive(*p);
  // `mel` may be DEVICE memory of this handle's GPU: the features then stay in HBM for a consumer on the device (the
  // Whisper encoder: SURVEY 8f NEXT-2) -- complete rows are written in place, nothing of them crosses the link
  bool mel_on_device = false;
  if (features) {
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, mel) == cudaSuccess) {
      if (pa.type == cudaMemoryTypeDevice) {
        if (pa.device != h->device) return fail(RHO_ERR_INVALID, "mel is memory of device %d, the handle runs on device %d", pa.device, h->device);
        mel_on_device = true;
      }
    } else {
      cudaGetLastError();                                   // plain host memory on drivers that report it as an error
    }
  }

  // ---- layout checks, item capacities, chunk plan
  bool contig = true;                                     // all offsets 16-byte aligned: one copy per chunk and direction
  for (int s = 0; s < n_segments; ++s) {
    if (seg_len[s] < 0 || seg_off[s] < 0) return fail(RHO_ERR_LAYOUT, "segment %d: negative offset / length", s);
    if (s + 1 < n_segments && seg_off[s + 1] < seg_off[s] + seg_len[s])
      return fail(RHO_ERR_LAYOUT, "segments must be laid out in increasing, non-overlapping order (segment %d)", s);
    if (seg_off[s] & 3) contig = false;
  }
  std::vector<int64_t> cap((size_t)n_items);
  for (int i = 0; i < n_items; ++i) {
    const int s0 = item_first_seg[i], s1 = item_first_seg[i + 1];
    if (s1 < s0) return fail(RHO_ERR_LAYOUT, "item_first_seg must be non-decreasing (item %d)", i);
    int64_t c = 0;
    for (int s = s0; s < s1; ++s) c += seg_len[s];
    c += (int64_t)std::max(0, s1 - s0 - 2) * d.pause;
    if (c > INT32_MAX) return fail(RHO_ERR_INVALID, "item %d is longer than 2^31 samples", i);
    cap[i] = c;
    if (y_off[i] < 0) return fail(RHO_ERR_LAYOUT, "item %d: negative output offset", i);
    if (i + 1 < n_items && y_off[i + 1] < y_off[i] + c)
      return fail(RHO_ERR_LAYOUT, "outputs must be laid out in increasing order with room for sum(len) + pauses (item %d)", i);
    if (y_off[i] & 3) contig = false;
  }
  // Chunk sizes ramp up from 1/16 of the budget (doubling) and down again at the end (halving): the pipeline fills
  // and drains with small chunks -- the first copy-out starts after ~0.1 ms instead of after a whole 120 MB chunk has
  // come in and been processed -- while the steady state runs on large ones (fewer launches and copy calls).
  const int64_t budget_max = chunk_budget_samples();
  const int64_t budget_min = std::max<int64_t>(1, budget_max / 16);
  int64_t total_samples = 0, done_samples = 0;
  for (int s = 0; s < n_segments; ++s) total_samples += seg_len[s];
  std::vector<Chunk> chunks;
  for (int i = 0; i < n_items;) {
    Chunk c{};
    c.i0 = i; c.s0 = item_first_seg[i];
    const int64_t budget = std::min(budget_max, std::max(budget_min, std::min(done_samples + budget_min,
                                                                              (total_samples - done_samples) / 2)));
    int64_t samples = 0;
    int j = i;
    while (j < n_items) {
      int64_t it = 0;
      for (int s = item_first_seg[j]; s < item_first_seg[j + 1]; ++s) it += seg_len[s];
      if (j > i && samples + it > budget) break;
      samples += it; ++j;
      if (j - i >= 65535) break;
    }
    done_samples += samples;
    c.i1 = j; c.s1 = item_first_seg[j];
    for (int s = c.s0; s < c.s1; ++s) c.max_seg_len = std::max<int64_t>(c.max_seg_len, seg_len[s]);
    for (int q = c.i0; q < c.i1; ++q) c.max_item_cap = std::max(c.max_item_cap, cap[q]);
    if (contig) {
      c.x_lo = c.s1 > c.s0 ? seg_off[c.s0] : 0;
      c.x_hi = c.s1 > c.s0 ? seg_off[c.s1 - 1] + seg_len[c.s1 - 1] : 0;
      c.y_lo = y_off[c.i0];
      c.y_hi = y_off[c.i1 - 1] + cap[c.i1 - 1];
      c.dx_samples = c.x_hi - c.x_lo;
      c.dy_samples = c.y_hi - c.y_lo;
    } else {
      for (int s = c.s0; s < c.s1; ++s) c.dx_samples += (int64_t)align_up((size_t)seg_len[s], 32);
      for (int q = c.i0; q < c.i1; ++q) c.dy_samples += (int64_t)align_up((size_t)cap[q], 32);
    }
    if (features) {
      c.t_dev = (int)rho_b200_compact_frames(c.max_item_cap, pad_frames);
      if (pad_frames == 0) c.t_dev = std::max(4, (c.t_dev + 3) / 4 * 4);
      if (mel_on_device && mel_stride_frames >= (pad_frames > 0 ? std::min<int64_t>(c.t_dev, pad_frames) : c.t_dev))
        c.t_dev = (int)mel_stride_frames;                  // the caller's rows are the rows the kernels write
      if (mel_stride_frames < (pad_frames > 0 ? std::min<int64_t>(c.t_dev, pad_frames) : rho_b200_compact_frames(c.max_item_cap, 0)))
        return fail(RHO_ERR_INVALID, "mel_stride_frames %lld too small: items of up to %lld samples need %d frames per row",
                    (long long)mel_stride_frames, (long long)c.max_item_cap, c.t_dev);
    }
    chunks.push_back(c);
    i = j;
  }
  const int n_chunks = (int)chunks.size();

  // ---- slot sizes
  size_t b_x = 0, b_y = 0, b_mel = 0, b_rec = 0, b_emb = 0, b_meta = 0, b_ws = 0, b_pad = 0;
  for (const Chunk& c : chunks) {
    const int ni = c.i1 - c.i0, ns = c.s1 - c.s0;
    b_x = std::max(b_x, align_up((size_t)(c.dx_samples + 32) * 4, 256));
    b_y = std::max(b_y, align_up((size_t)(c.dy_samples + 32) * 4, 256));
    if (features && !mel_on_device) b_mel = std::max(b_mel, align_up((size_t)ni * n_mels * c.t_dev * 4, 256));
    b_rec = std::max(b_rec, align_up(sizeof(rho_record) * (size_t)ni, 256));
    if (have_emb) b_emb = std::max(b_emb, align_up(sizeof(float) * (size_t)ni * emb_dim, 256));
    b_meta = std::max(b_meta, align_up((size_t)(ns + 1) * 12 + (size_t)(ni + 1) * 12 + 64, 256));
    b_ws = std::max(b_ws, align_up(rho_b200_workspace_bytes(ns, ni, c.max_seg_len), 256));
    b_pad = std::max(b_pad, align_up(sizeof(float) * (size_t)ni, 256));
  }
  const size_t b_ref = align_up(sizeof(float) * (size_t)std::max(1, emb_dim), 256);
  const size_t per_slot = b_x + b_y + b_mel + b_rec + b_emb + b_meta + b_ws + b_pad;
  const int HC_SLOTS = host_slots();
  const size_t total = (size_t)HC_SLOTS * per_slot + b_ref;

  CtxLease lease(h);
  if (!lease.c) return cuda_fail(lease.err, "host context");
  HostCtx& C = *lease.c;
  cudaError_t e;
  if (C.arena_bytes < total) {
    if (C.arena) cudaFree(C.arena);
    C.arena = nullptr; C.arena_bytes = 0;
    if ((e = cudaMalloc(&C.arena, total)) != cudaSuccess) return cuda_fail(e, "cudaMalloc(arena)");
    C.arena_bytes = total;
  }
  // pinned staging: every chunk's metadata (lives until its copy ran) + the pad values coming back
  const size_t pin_total = (size_t)n_chunks * b_meta + align_up(sizeof(float) * (size_t)n_items, 256);
  if (C.pinned_bytes < pin_total) {
    if (C.pinned) cudaFreeHost(C.pinned);
    C.pinned = nullptr; C.pinned_bytes = 0;
    if ((e = cudaMallocHost(&C.pinned, pin_total)) != cudaSuccess) return cuda_fail(e, "cudaMallocHost(staging)");
    C.pinned_bytes = pin_total;
  }
  while ((int)C.ev_out.size() < n_chunks) {
    cudaEvent_t ev = nullptr, ev2 = nullptr;
    if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming | cudaEventBlockingSync)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&ev2, cudaEventDisableTiming | cudaEventBlockingSync)) != cudaSuccess) {
      if (ev) cudaEventDestroy(ev);
      return cuda_fail(e, "cudaEventCreate");
    }
    C.ev_out.push_back(ev);
    C.ev_small.push_back(ev2);
  }
  char* base = (char*)C.arena;
  float* d_ref = (float*)base;
  char* slots = base + b_ref;
  float* h_pad = (float*)((char*)C.pinned + (size_t)n_chunks * b_meta);
  cudaStream_t sc = C.s_compute;
  if (have_emb && (e = cudaMemcpyAsync(d_ref, ref_emb, sizeof(float) * emb_dim, cudaMemcpyHostToDevice, sc)) != cudaSuccess)
    return cuda_fail(e, "ref upload");

  // ---- host fill of the constant tail of every feature row, by helper threads, chunk by chunk as the copies land
  const bool host_fill = features && pad_frames > 0 && !mel_on_device;
  std::atomic<int> issued{0};            // chunks whose copy-out has been enqueued (ev_out recorded)
  std::atomic<int> abort_fill{0};
  std::atomic<int> fill_err{0};
  std::atomic<int> out_recorded{0};      // chunks whose ev_out has been recorded in THIS call
  auto fill_worker = [&](int w, int W) {
    cudaSetDevice(h->device);
    for (int k = 0; k < n_chunks; ++k) {
      while (issued.load(std::memory_order_acquire) <= k) {
        if (abort_fill.load(std::memory_order_relaxed)) return;
        std::this_thread::yield();
      }
      // the constant tail of a row only needs the chunk's pad values: it is written while the chunk's frames with
      // signal (a disjoint part of the same rows) are still crossing the link
      if (cudaEventSynchronize(C.ev_small[k]) != cudaSuccess) { fill_err.store(1); return; }
      const Chunk& c = chunks[k];
      const int64_t t0 = std::min<int64_t>(c.t_dev, mel_stride_frames), t1 = std::min<int64_t>(mel_stride_frames, pad_frames);
      const int64_t rows = (int64_t)(c.i1 - c.i0) * n_mels;
      if (pad_value && w == 0) memcpy(pad_value + c.i0, h_pad + c.i0, sizeof(float) * (size_t)(c.i1 - c.i0));
      if (w == 0 && debug_fill_delay_us() > 0) std::this_thread::sleep_for(std::chrono::microseconds(debug_fill_delay_us()));
      if (t1 <= t0 || debug_skip_fill()) continue;
      for (int64_t r = w; r < rows; r += W) {
        const int64_t it = c.i0 + r / n_mels;
        host_fill_f32(mel + ((int64_t)c.i0 * n_mels + r) * mel_stride_frames + t0, t1 - t0, h_pad[it]);
      }
    }
  };
  std::vector<std::thread> helpers;
  const int W = host_fill ? fill_threads(h) : 0;
  for (int w = 1; w < W; ++w) helpers.emplace_back(fill_worker, w, W);
  // does the fill keep up with the link on this host?  A watcher sleeps on the last chunk's copy-out event and notes
  // when it completed; the fill's own end is compared with it below.
  using Clock = std::chrono::steady_clock;
  const Clock::time_point t_start = Clock::now();
  Clock::time_point t_copies_done = t_start;
  const bool adapt = host_fill && fill_threads_env() == 0 && W < FILL_THREADS_MAX;
  if (adapt) helpers.emplace_back([&] {
    cudaSetDevice(h->device);
    while (out_recorded.load(std::memory_order_acquire) < n_chunks) {
      if (abort_fill.load(std::memory_order_relaxed)) return;
      std::this_thread::yield();
    }
    if (cudaEventSynchronize(C.ev_out[n_chunks - 1]) == cudaSuccess) t_copies_done = Clock::now();
  });

  int status = RHO_OK;
  for (int k = 0; k < n_chunks && status == RHO_OK; ++k) {
    const Chunk& c = chunks[k];
    const int s = k % HC_SLOTS;
    const int ni = c.i1 - c.i0, ns = c.s1 - c.s0;
    char* sb = slots + (size_t)s * per_slot;
    float* d_x = (float*)sb;
    float* d_y = (float*)(sb + b_x);
    float* d_mel = mel_on_device ? mel + (size_t)c.i0 * n_mels * mel_stride_frames : (float*)(sb + b_x + b_y);
    rho_record* d_rec = (rho_record*)(sb + b_x + b_y + b_mel);
    float* d_emb = (float*)(sb + b_x + b_y + b_mel + b_rec);
    char* d_meta = sb + b_x + b_y + b_mel + b_rec + b_emb;
    void* d_ws = d_meta + b_meta;
    float* d_pad = (float*)((char*)d_ws + b_ws);
    // metadata of this chunk in device coordinates
    char* hm = (char*)C.pinned + (size_t)k * b_meta;
    int64_t* m_soff = (int64_t*)hm;
    int64_t* m_yoff = m_soff + (ns + 1);
    int32_t* m_slen = (int32_t*)(m_yoff + (ni + 1));
    int32_t* m_first = m_slen + (ns + 1);
    {
      int64_t run = 0;
      for (int q = 0; q < ns; ++q) {
        m_soff[q] = contig ? seg_off[c.s0 + q] - c.x_lo : run;
        m_slen[q] = seg_len[c.s0 + q];
        run += (int64_t)align_up((size_t)seg_len[c.s0 + q], 32);
      }
      m_soff[ns] = 0; m_slen[ns] = 0;
      run = 0;
      for (int q = 0; q < ni; ++q) {
        m_yoff[q] = contig ? y_off[c.i0 + q] - c.y_lo : run;
        m_first[q] = item_first_seg[c.i0 + q] - c.s0;
        run += (int64_t)align_up((size_t)cap[c.i0 + q], 32);
      }
      m_yoff[ni] = 0; m_first[ni] = ns;
    }
    const size_t meta_bytes = (size_t)(ns + 1) * 12 + (size_t)(ni + 1) * 12;
    // slot reuse: the copy-in of chunk k overwrites what chunk k - SLOTS read and wrote
    if (k >= HC_SLOTS) cudaStreamWaitEvent(C.s_in, C.ev_out[k - HC_SLOTS], 0);
    e = cudaMemcpyAsync(d_meta, hm, meta_bytes, cudaMemcpyHostToDevice, C.s_in);
    if (e == cudaSuccess && ns > 0) {
      if (contig) {
        e = cudaMemcpyAsync(d_x, x + c.x_lo, sizeof(float) * (size_t)(c.x_hi - c.x_lo), cudaMemcpyHostToDevice, C.s_in);
      } else {
        for (int q = 0; q < ns && e == cudaSuccess; ++q)
          if (m_slen[q] > 0)
            e = cudaMemcpyAsync(d_x + m_soff[q], x + seg_off[c.s0 + q], sizeof(float) * (size_t)m_slen[q],
                                cudaMemcpyHostToDevice, C.s_in);
      }
    }
    if (e == cudaSuccess && have_emb)
      e = cudaMemcpyAsync(d_emb, emb + (size_t)c.i0 * emb_dim, sizeof(float) * (size_t)ni * emb_dim, cudaMemcpyHostToDevice, C.s_in);
    if (e != cudaSuccess) { status = cuda_fail(e, "H2D"); break; }
    cudaEventRecord(C.ev_in[s], C.s_in);
    cudaStreamWaitEvent(sc, C.ev_in[s], 0);
    const int64_t* dd_soff = (const int64_t*)d_meta;
    const int64_t* dd_yoff = dd_soff + (ns + 1);
    const int32_t* dd_slen = (const int32_t*)(dd_yoff + (ni + 1));
    const int32_t* dd_first = dd_slen + (ns + 1);
    bool one_seg = ns == ni;
    for (int q = 0; q < ni && one_seg; ++q) one_seg = m_first[q] == q;
    if (features) {
      status = rho_b200_validate(h, d_x, dd_soff, dd_slen, ns, c.max_seg_len, dd_first, ni, c.max_item_cap, p, d_y, dd_yoff,
                                 n_mels, pad_frames, d_mel, c.t_dev, d_pad, have_emb ? d_emb : nullptr,
                                 have_emb ? d_ref : nullptr, emb_dim, d_rec, nullptr,
                                 (one_seg ? RHO_V_ONE_SEGMENT_ITEMS : 0u) | (pad_frames > 0 ? RHO_V_COMPACT_PAD : 0u),
                                 d_ws, b_ws, sc);
    } else {
      status = rho_b200_join(h, d_x, dd_soff, dd_slen, ns, c.max_seg_len, dd_first, ni, c.max_item_cap, p, d_y, dd_yoff,
                             d_rec, nullptr, d_ws, b_ws, sc);
      if (status == RHO_OK && have_emb)
        status = rho_b200_cosine(h, d_emb, d_ref, ni, emb_dim, &d_rec[0].cosine, (int)sizeof(rho_record), sc);
    }
    if (status != RHO_OK) break;
    cudaEventRecord(C.ev_done[s], sc);
    cudaStreamWaitEvent(C.s_out, C.ev_done[s], 0);
    // small results first: the records and the per-item pad values (the host fill of this chunk starts on them)
    e = cudaMemcpyAsync(rec + c.i0, d_rec, sizeof(rho_record) * (size_t)ni, cudaMemcpyDeviceToHost, C.s_out);
    if (e == cudaSuccess && features && pad_frames > 0)
      e = cudaMemcpyAsync(h_pad + c.i0, d_pad, sizeof(float) * (size_t)ni, cudaMemcpyDeviceToHost, C.s_out);
    if (e == cudaSuccess) e = cudaEventRecord(C.ev_small[k], C.s_out);
    issued.store(k + 1, std::memory_order_release);
    if (e == cudaSuccess) {
      if (contig) {
        e = cudaMemcpyAsync(y + c.y_lo, d_y, sizeof(float) * (size_t)(c.y_hi - c.y_lo), cudaMemcpyDeviceToHost, C.s_out);
      } else {
        for (int q = 0; q < ni && e == cudaSuccess; ++q)
          if (cap[c.i0 + q] > 0)
            e = cudaMemcpyAsync(y + y_off[c.i0 + q], d_y + m_yoff[q], sizeof(float) * (size_t)cap[c.i0 + q],
                                cudaMemcpyDeviceToHost, C.s_out);
      }
    }
    if (e == cudaSuccess && features && !mel_on_device) {
      // compact device rows -> the caller's rows: only the frames that can see signal cross the link
      const int64_t wf = std::min<int64_t>(c.t_dev, mel_stride_frames);
      float* dst = mel + (size_t)c.i0 * n_mels * mel_stride_frames;
      if (wf == c.t_dev && wf == mel_stride_frames)      // the caller's rows are the device's rows: one linear copy
        e = cudaMemcpyAsync(dst, d_mel, sizeof(float) * (size_t)wf * ni * n_mels, cudaMemcpyDeviceToHost, C.s_out);
      else
        e = cudaMemcpy2DAsync(dst, sizeof(float) * (size_t)mel_stride_frames, d_mel, sizeof(float) * (size_t)c.t_dev,
                              sizeof(float) * (size_t)wf, (size_t)ni * n_mels, cudaMemcpyDeviceToHost, C.s_out);
    }
    if (e != cudaSuccess) { status = cuda_fail(e, "D2H"); break; }
    cudaEventRecord(C.ev_out[k], C.s_out);
    out_recorded.store(k + 1, std::memory_order_release);
  }
  if (status != RHO_OK) abort_fill.store(1);
  if (host_fill && status == RHO_OK) fill_worker(0, W);     // the calling thread is worker 0
  const Clock::time_point t_fill_done = Clock::now();
  for (std::thread& t : helpers) t.join();
  if (adapt && status == RHO_OK) {
    // (t_fill_done is worker 0's end; the other workers end within a row of it)
    const double lag = std::chrono::duration<double>(t_fill_done - t_copies_done).count();
    const double call = std::chrono::duration<double>(t_fill_done - t_start).count();
    if (t_copies_done != t_start && lag > 0.05 * call)
      h->host_fill_threads.store(std::min(FILL_THREADS_MAX, W + 2), std::memory_order_relaxed);
  }
  const bool pad_late = status == RHO_OK && !host_fill && features && pad_frames > 0 && pad_value;
  const cudaError_t e1 = cudaStreamSynchronize(C.s_in);
  const cudaError_t e2 = cudaStreamSynchronize(sc);
  const cudaError_t e3 = cudaStreamSynchronize(C.s_out);
  if (pad_late && e3 == cudaSuccess) memcpy(pad_value, h_pad, sizeof(float) * (size_t)n_items);
  if (status != RHO_OK) return status;
  if (e1 != cudaSuccess) return cuda_fail(e1, "sync copy-in");
  if (e2 != cudaSuccess) return cuda_fail(e2, "sync compute");
  if (e3 != cudaSuccess) return cuda_fail(e3, "sync copy-out");
  if (fill_err.load()) return fail(RHO_ERR_CUDA, "host fill: waiting for a copy-out event failed");
  return RHO_OK;
}

int rho_b200_host_fill_threads(rho_handle* h) {
  if (!h) return fail(RHO_ERR_INVALID, "handle is NULL");
  return fill_threads(h);
}

int rho_b200_validate_host(rho_handle* h, const float* x, int n, int32_t clip_len, const rho_params* p,
                           float* y, int n_mels, int pad_frames, float* mel, const float* emb,
                           const float* ref_emb, int emb_dim, rho_record* rec) {
  if (!h) return fail(RHO_ERR_INVALID, "handle is NULL");
  if (n < 0 || clip_len <= 0) return fail(RHO_ERR_INVALID, "bad size");
  if (n == 0) return RHO_OK;
  if (!x || !y || !rec) return fail(RHO_ERR_INVALID, "NULL host pointer");
  if (mel && pad_frames != MEL_PAD_FRAMES) return fail(RHO_ERR_INVALID, "host entry point supports pad_frames=3000 only");
  std::vector<int64_t> off((size_t)n);
  std::vector<int32_t> len((size_t)n), first((size_t)n + 1);
  for (int i = 0; i < n; ++i) { off[i] = (int64_t)i * clip_len; len[i] = clip_len; first[i] = i; }
  first[n] = n;
  return rho_b200_validate_host_ragged(h, x, off.data(), len.data(), n, first.data(), n, p, y, off.data(), n_mels,
                                       pad_frames, mel, pad_frames, nullptr, emb, ref_emb, emb_dim, rec);
}

}  // extern "C"
