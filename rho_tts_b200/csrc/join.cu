// Post-process / join kernels: silence-trim scan, DC, crossfade gather, fades, decay sums.
//
// Reference semantics: src/rho_tts/base_tts.py
//   _trim_silence :348-392, _remove_dc_offset :394-399, _apply_fades :401-433,
//   _smooth_segment_join :435-536, _validate_sound_decay :297-323.
//
// Bit-exactness of the trim decision: torch's CPU avg_pool1d adds the `window`
// squared samples of a frame strictly left to right in fp32 (zero padding included)
// and divides by `window`.  k_scan reproduces that chain with __fmul_rn/__fadd_rn
// (never contracted into an FMA) so `sqrt(sum/window) > thr` is the same predicate.
#include "common.cuh"
#include "kernels.h"
#include "records_dev.cuh"

namespace rho {

// ------------------------------------------------------------------ init
__global__ void k_init_items(SegState* __restrict__ seg, ItemState* __restrict__ item,
                             const int32_t* __restrict__ item_first_seg, int n_items,
                             int* __restrict__ clip_max, int* __restrict__ tiles_done,
                             int* __restrict__ work_counter) {
  const int it = blockIdx.x * blockDim.x + threadIdx.x;
  if (it == 0 && work_counter) *work_counter = 0;
  if (it >= n_items) return;
  if (clip_max) clip_max[it] = INT_MIN;                // log-mel running maximum (what k_logmel_init does)
  if (tiles_done) tiles_done[it] = 0;
  ItemState z;
  z.s_first = 0.0; z.s_last = 0.0; z.out_len = 0; z.flags = 0; z.pad0 = 0; z.pad1 = 0;
  item[it] = z;
  const int s0 = item_first_seg[it], s1 = item_first_seg[it + 1];
  const int n = s1 - s0;
  for (int s = s0; s < s1; ++s) {
    const int pos = s - s0;
    const bool fs = (n == 1) || pos > 0;        // base_tts.py:449, 469-474
    const bool fe = (n == 1) || pos < n - 1;
    SegState st;
    st.first = INT_MAX; st.last = -1; st.start = 0; st.end = 0; st.dc = 0.f;
    st.flags = (fs ? 0x100u : 0u) | (fe ? 0x200u : 0u);
    st.item = it; st.pos = pos;
    seg[s] = st;
  }
}

__global__ void k_init_segs(SegState* __restrict__ seg, const uint8_t* __restrict__ trim_flags, int n_seg) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_seg) return;
  const unsigned tf = trim_flags ? trim_flags[s] : 3u;
  SegState st;
  st.first = INT_MAX; st.last = -1; st.start = 0; st.end = 0; st.dc = 0.f;
  st.flags = ((tf & 1u) ? 0x100u : 0u) | ((tf & 2u) ? 0x200u : 0u);
  st.item = s; st.pos = 0;
  seg[s] = st;
}

// ------------------------------------------------------------------ scan
// One thread per energy frame, FR frames per CTA.  The CTA stages the samples its frames
// cover in shared memory, hop-block by hop-block with a row stride RS chosen so that the
// per-thread walk (thread t starts at row t) is bank-conflict free:
//   VEC path  (hop % 4 == 0, window == 2*hop): RS = 4*odd, rows read with 128-bit LDS;
//   generic   : RS odd, scalar LDS.
// Thread t of the tile owns frame f = frame0 + t, which covers rows t and t+1 (plus one
// sample of row t+2 when `window` is odd), and the DC block sum of row t+1 (= global hop
// block f, samples [f*hop, (f+1)*hop)).
template <bool VEC>
__global__ void __launch_bounds__(SCAN_FR)
k_scan(const float* __restrict__ x, const int64_t* __restrict__ off, const int32_t* __restrict__ len,
       SegState* __restrict__ seg, float* __restrict__ block_sum, int blocks_per_seg,
       int window, int hop, int RS, int skew, float thr) {
  extern __shared__ __align__(16) float sm[];
  const int s = blockIdx.x;
  const int L = len[s];
  const int n_frames = (L <= 0) ? 0 : (L + 2 * hop - window) / hop + 1;
  const int frame0 = blockIdx.y * SCAN_FR;
  if (frame0 >= n_frames) return;
  const float* __restrict__ xs = x + off[s];
  const int rows = SCAN_FR + 2;
  const long long g0 = (long long)(frame0 - 1) * hop;  // global sample of row 0, col 0 (pad == hop)

  if (VEC) {
    // TMA bulk copies: every hop row that lies fully inside the clip is ONE cp.async.bulk (480 B at
    // 24 kHz) issued by lane 0 of a warp and completing on an mbarrier; no registers, no per-element
    // instructions.  The few rows that stick out of the clip (avg_pool's zero padding, the ragged
    // end) are staged with zero-filling LDGSTS.
    __shared__ __align__(8) uint64_t bar;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned row_bytes = (unsigned)hop * 4u;
    // rows [r_lo, r_hi) are interior: g0 + r*hop >= 0 and g0 + (r+1)*hop <= L
    int r_lo = 0;
    if (g0 < 0) r_lo = (int)((-g0 + hop - 1) / hop);
    long long r_hi_ll = (L - g0) / hop;
    int r_hi = r_hi_ll < 0 ? 0 : (r_hi_ll > rows ? rows : (int)r_hi_ll);
    if (r_lo > r_hi) r_lo = r_hi;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    __syncthreads();
    if (threadIdx.x == 0) mbar_arrive_expect_tx(&bar, (unsigned)(r_hi - r_lo) * row_bytes);
    if (RS == hop) {
      // dense rows: the whole interior of the tile is ONE bulk copy (the TMA unit retires ~1 op per 46 cycles,
      // so a copy per 480-byte row is what bounds the padded layout)
      if (threadIdx.x == 0 && r_hi > r_lo)
        bulk_g2s(sm + r_lo * RS, xs + (g0 + (long long)r_lo * hop), (unsigned)(r_hi - r_lo) * row_bytes, &bar);
    } else if (lane == 0) {
      for (int row = r_lo + warp; row < r_hi; row += SCAN_FR / 32)
        bulk_g2s(sm + row * RS, xs + (g0 + (long long)row * hop), row_bytes, &bar);
    }
    const int q_per_row = hop >> 2;
    for (int row = warp; row < rows; row += SCAN_FR / 32) {
      if (row >= r_lo && row < r_hi) continue;
      const long long grow = g0 + (long long)row * hop;
      for (int c = lane; c < q_per_row; c += 32) {
        const long long g = grow + 4 * c;
        int nb = 0;
        if (g >= 0 && g < L) nb = (L - g >= 4) ? 16 : 4 * (int)(L - g);
        cp_async16_zfill(sm + row * RS + 4 * c, nb ? xs + g : xs, nb);
      }
    }
    cp_async_commit();
    cp_async_wait_all();
    mbar_wait(&bar, 0);
  } else {
    const int total = rows * hop;
    for (int r = threadIdx.x; r < total; r += SCAN_FR) {
      const int row = r / hop, c = r - row * hop;
      const long long g = g0 + r;
      sm[row * RS + c] = (g >= 0 && g < L) ? xs[g] : 0.f;
    }
  }
  __syncthreads();

  const int t = threadIdx.x;
  const int f = frame0 + t;
  bool loud = false;
  if (f < n_frames) {
    float acc = 0.f, bs = 0.f;
    if (VEC) {
      // x*x on the packed fp32x2 pipe (same IEEE rounding per component as __fmul_rn), the 240-term
      // sum as the strictly sequential scalar chain torch's avg_pool1d runs; DC block sum packed too.
      const float4* p0 = reinterpret_cast<const float4*>(sm + t * RS);
      const float4* p1 = reinterpret_cast<const float4*>(sm + (t + 1) * RS);
      const int nq = hop >> 2;
      // Dense rows (RS == hop, hop/4 = 2 mod 4) put rows r and r+4 on the same banks.  `skew`: the threads
      // whose row has bit 2 set run ONE 128-bit step behind the others, which moves them to the odd bank
      // groups -- every quarter-warp then covers all 32 banks.  Each thread still adds its own 240 terms
      // strictly left to right, so the fp32 chain (and the trim decision) is unchanged.
      const int d0 = skew ? ((t >> 2) & 1) : 0, d1 = skew ? (((t + 1) >> 2) & 1) : 0;
      const int steps = nq + (skew ? 1 : 0);
#pragma unroll 6
      for (int i = 0; i < steps; ++i) {
        const int j = i - d0;
        if (j >= 0 && j < nq) {
          const float4 v = p0[j];
          const float2 a = __fmul2_rn(make_float2(v.x, v.y), make_float2(v.x, v.y));
          const float2 b = __fmul2_rn(make_float2(v.z, v.w), make_float2(v.z, v.w));
          acc = __fadd_rn(acc, a.x); acc = __fadd_rn(acc, a.y);
          acc = __fadd_rn(acc, b.x); acc = __fadd_rn(acc, b.y);
        }
      }
      float2 bs2 = make_float2(0.f, 0.f);
#pragma unroll 6
      for (int i = 0; i < steps; ++i) {
        const int j = i - d1;
        if (j >= 0 && j < nq) {
          const float4 v = p1[j];
          const float2 a = __fmul2_rn(make_float2(v.x, v.y), make_float2(v.x, v.y));
          const float2 b = __fmul2_rn(make_float2(v.z, v.w), make_float2(v.z, v.w));
          acc = __fadd_rn(acc, a.x); acc = __fadd_rn(acc, a.y);
          acc = __fadd_rn(acc, b.x); acc = __fadd_rn(acc, b.y);
          bs2 = __fadd2_rn(bs2, make_float2(v.x, v.y));
          bs2 = __fadd2_rn(bs2, make_float2(v.z, v.w));
        }
      }
      bs = bs2.x + bs2.y;
    } else {
      int left = window;
      for (int b = 0; left > 0; ++b) {
        const float* p = sm + (t + b) * RS;
        const int cnt = left < hop ? left : hop;
        for (int j = 0; j < cnt; ++j) {
          const float v = p[j];
          acc = __fadd_rn(acc, __fmul_rn(v, v));
        }
        left -= cnt;
      }
      const float* p1 = sm + (t + 1) * RS;
      for (int j = 0; j < hop; ++j) bs += p1[j];
    }
    const float e = __fsqrt_rn(__fdiv_rn(acc, (float)window));
    loud = e > thr;
    if (f < blocks_per_seg) block_sum[(size_t)s * blocks_per_seg + f] = bs;
  }
  const unsigned m = __ballot_sync(0xffffffffu, loud);
  if (m != 0u && (threadIdx.x & 31) == 0) {
    const int wbase = frame0 + (threadIdx.x & ~31);
    atomicMin(&seg[s].first, wbase + (__ffs(m) - 1));
    atomicMax(&seg[s].last, wbase + (31 - __clz(m)));
  }
}

// ------------------------------------------------------------------ per-segment finalize
// One warp per segment: frames -> [start, end), flags, DC = mean over [start, end).
__global__ void k_finalize_segs(const float* __restrict__ x, const int64_t* __restrict__ off,
                                const int32_t* __restrict__ len, SegState* __restrict__ seg,
                                const float* __restrict__ block_sum, int blocks_per_seg,
                                int n_seg, int window, int hop, int trim_enabled,
                                rho_seg_info* __restrict__ info_out, ItemState* __restrict__ one_seg_item) {
  const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (s >= n_seg) return;
  SegState st = seg[s];
  const int L = len[s];
  const bool fs = st.flags & 0x100u, fe = st.flags & 0x200u;
  int start, end;
  uint32_t flags = 0;
  if (!trim_enabled || L <= 0) {                       // base_tts.py:360-361
    start = 0; end = L > 0 ? L : 0; flags = RHO_F_UNTOUCHED;
  } else if (st.last < 0) {                            // :379-380  audio[:, :window_size] (2-D)
    start = 0; end = window < L ? window : L; flags = RHO_F_ALL_SILENT;
  } else {                                             // :386-390
    const long long a = fs ? ((long long)st.first * window) / 2 : 0;
    const long long b = fe ? ((long long)(st.last + 2) * window) / 2 : (long long)L;
    long long sa = a < L ? a : L; if (sa < 0) sa = 0;
    long long eb = b < L ? b : L; if (eb < sa) eb = sa;
    start = (int)sa; end = (int)eb;
  }
  // DC over [start, end): whole hop blocks from block_sum, ragged edges straight from x.
  const float* __restrict__ xs = x + off[s];
  double acc = 0.0;
  const int n = end - start;
  if (n > 0) {
    int kb0 = (start + hop - 1) / hop, kb1 = end / hop;
    if (kb1 > blocks_per_seg) kb1 = blocks_per_seg;
    if (kb0 <= kb1) {
      // eight independent loads per lane in flight (one at a time made this a 19 us chain of L2 round trips)
      const float* __restrict__ bs = block_sum + (size_t)s * blocks_per_seg;
      int k = kb0 + lane;
      for (; k + 7 * 32 < kb1; k += 8 * 32) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __ldg(bs + k + 32 * j);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc += (double)v[j];
      }
      for (; k < kb1; k += 32) acc += (double)__ldg(bs + k);
      for (int g = start + lane; g < kb0 * hop; g += 32) acc += (double)xs[g];
      for (int g = kb1 * hop + lane; g < end; g += 32) acc += (double)xs[g];
    } else {
      for (int g = start + lane; g < end; g += 32) acc += (double)xs[g];
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    st.start = start; st.end = end;
    st.dc = n > 0 ? (float)(acc / (double)n) : 0.f;
    st.flags = (st.flags & 0xffffff00u) | flags;
    seg[s] = st;
    if (one_seg_item) {                                  // item s IS segment s (base_tts.py:447-452): k_plan_items' n == 1 case
      one_seg_item[s].out_len = end - start;
      one_seg_item[s].flags = (flags & RHO_F_ALL_SILENT) ? (RHO_F_ALL_SILENT | RHO_F_TWO_D) : (flags & RHO_F_UNTOUCHED);
    }
    if (info_out) {
      rho_seg_info o; o.start = start; o.end = end; o.dc = st.dc; o.flags = flags;
      info_out[s] = o;
    }
  }
}

// ------------------------------------------------------------------ per-item plan
// One thread per item.  Emits, for every segment, the contiguous span of the joined output it
// produces: [crossfade with previous | body | pause].  Mirrors the emit order of
// base_tts.py:481-523 and the rank bookkeeping that makes torch.cat throw (-> fallback, :530-533).
__global__ void k_plan_items(const SegState* __restrict__ seg, const int32_t* __restrict__ seg_len,
                             SegSpan* __restrict__ span, ItemState* __restrict__ item,
                             const int32_t* __restrict__ item_first_seg, int n_items,
                             int cf, int pause, int pause_on,
                             const int64_t* __restrict__ seg_off, const int64_t* __restrict__ y_off) {
  const int it = blockIdx.x * blockDim.x + threadIdx.x;
  if (it >= n_items) return;
  const int s0 = item_first_seg[it], s1 = item_first_seg[it + 1];
  const int n = s1 - s0;
  if (n <= 0) { item[it].out_len = 0; item[it].flags = 0; return; }
  SegSpan z; z.dst = 0; z.ov = 0; z.body = 0; z.pause = 0; z.prev_tail = 0; z.item = it; z.out_len = 0;
  z.item_flags = 0; z.dc = 0.f; z.dcp = 0.f; z.pad0 = z.pad1 = 0; z.x_base = 0; z.prev_x = 0; z.y_base = 0;
  // the part of the span record that k_gather reads instead of chasing seg -> item -> offsets
  auto describe = [&](int out_len, uint32_t flags) {
    if (!seg_off || !y_off) return;
    const bool fb = (flags & RHO_F_FALLBACK) != 0;
    const long long yb = y_off[it];
    for (int i = 0; i < n; ++i) {
      SegSpan sp = span[s0 + i];
      const SegState st = seg[s0 + i];
      sp.item = it; sp.out_len = out_len; sp.item_flags = flags;
      sp.dc = fb ? 0.f : st.dc;
      sp.x_base = seg_off[s0 + i] + (fb ? 0 : st.start);
      sp.y_base = yb + sp.dst;
      if (sp.ov > 0 && i > 0) {
        const SegState pst = seg[s0 + i - 1];
        sp.dcp = pst.dc;
        sp.prev_x = seg_off[s0 + i - 1] + pst.start + sp.prev_tail;
      }
      span[s0 + i] = sp;
    }
  };
  if (n == 1) {                                        // :447-452
    const SegState st = seg[s0];
    SegSpan sp = z; sp.body = st.end - st.start;
    span[s0] = sp;
    item[it].out_len = sp.body;
    item[it].flags = (st.flags & RHO_F_ALL_SILENT) ? (RHO_F_ALL_SILENT | RHO_F_TWO_D) : (st.flags & RHO_F_UNTOUCHED);
    describe(sp.body, item[it].flags);
    return;
  }
  bool has1 = false, has2 = false;
  auto mark = [&](bool two_d, int cnt) {
    if (cnt == 0 && !two_d) return;                    // torch.cat skips 1-D empties
    if (two_d) has2 = true; else has1 = true;
  };
  int pos = 0;
  int Lprev = 0; bool dprev = false;
  for (int i = 0; i < n; ++i) {
    const SegState st = seg[s0 + i];
    const int Li = st.end - st.start;
    const bool di = (st.flags & RHO_F_ALL_SILENT) != 0;
    SegSpan sp = z; sp.dst = pos;
    if (i == 0) {
      // :484-488 -- with crossfade_samples == 0 the reference's slice [..., :-0] is EMPTY: segment 0 is dropped
      sp.body = (Li > cf) ? (cf > 0 ? Li - cf : 0) : Li;
      mark(di, sp.body);
    } else {
      int ov = cf < Lprev ? cf : Lprev; if (Li < ov) ov = Li;   // :491
      if (ov > 10) {
        mark(dprev || di, ov);
        const int rem = (i < n - 1 && Li > ov + cf) ? Li - ov - cf : Li - ov;   // :507-513
        if (rem > 0) mark(di, rem);
        const int pz = (pause_on && i < n - 1) ? pause : 0;                      // :518-521
        if (pause_on && i < n - 1) mark(false, pz);
        sp.ov = ov; sp.body = rem; sp.pause = pz; sp.prev_tail = Lprev - ov;
      } else {
        sp.body = Li;                                  // :523
        mark(di, Li);
      }
    }
    span[s0 + i] = sp;
    pos += sp.ov + sp.body + sp.pause;
    Lprev = Li; dprev = di;
  }
  if (has1 && has2) {
    // fallback: torch.cat(ORIGINAL segments) -- untrimmed, DC kept, no crossfade, no pause.
    int p = 0;
    for (int i = 0; i < n; ++i) {
      SegSpan sp = z; sp.dst = p; sp.body = seg_len[s0 + i];
      span[s0 + i] = sp; p += sp.body;
    }
    item[it].out_len = p; item[it].flags = RHO_F_FALLBACK;
    describe(p, RHO_F_FALLBACK);
  } else {
    item[it].out_len = pos; item[it].flags = has2 ? RHO_F_TWO_D : 0u;
    describe(pos, has2 ? RHO_F_TWO_D : 0u);
  }
}

// ------------------------------------------------------------------ gather
__device__ __forceinline__ float fade_gain(int o, int out_len, int fade) {
  // _apply_fades :420-431; skipped entirely when the item is shorter than 2*fade.
  float g = 1.f;
  if (fade > 0 && out_len >= 2 * fade) {
    if (o < fade) g = 0.5f * (1.f - cosf(linspace32(0.f, RHO_PI_F, fade, o)));
    else if (o >= out_len - fade) g = 0.5f * (1.f + cosf(linspace32(0.f, RHO_PI_F, fade, o - (out_len - fade))));
  }
  return g;
}

// grid (n_seg, tiles); each CTA writes GATHER_TILE consecutive samples of one segment's span.
// Slow path of k_gather for the (few) 128-bit pieces that touch a crossfade, a pause, a fade, a span edge or an
// unaligned segment.  Out of line on purpose: inlined eight times with its four cosf argument reductions it made
// the kernel 260 KB of code, and the instruction cache (not HBM) was what the interior stream waited for.
__device__ __noinline__ void gather_edge_piece(int sp_dst, int sp_ov, int sp_body, const float* __restrict__ xc,
                                               const float* __restrict__ xp, float* __restrict__ yo, int j, int jend,
                                               float dc, float dcp, bool need_fade, int fade, int out_len, int third,
                                               float* a_first, float* a_last) {
  const int o = sp_dst + j;
  for (int k = 0; k < 4; ++k) {
    const int jj = j + k;
    if (jj >= jend) break;
    float val;
    if (jj < sp_ov) {               // :497-504 equal-power crossfade
      const float fo = cosf(linspace32(0.f, RHO_HALF_PI_F, sp_ov, jj));
      const float fi = cosf(linspace32(RHO_HALF_PI_F, 0.f, sp_ov, jj));
      const float a = __fmul_rn(__fsub_rn(xp[jj], dcp), fo);
      const float b = __fmul_rn(__fsub_rn(xc[jj], dc), fi);
      val = __fadd_rn(a, b);
    } else if (jj < sp_ov + sp_body) {
      val = __fsub_rn(xc[jj], dc);
    } else {
      val = 0.f;                    // inter-sentence pause
    }
    const int oo = o + k;
    if (need_fade) val = __fmul_rn(val, fade_gain(oo, out_len, fade));
    yo[jj] = val;
    if (oo < third) *a_first += val * val;
    if (oo >= out_len - third) *a_last += val * val;
  }
}

#ifndef RHO_GATHER_TILE_MAJOR
#define RHO_GATHER_TILE_MAJOR 1
#endif
#ifndef RHO_GATHER_MINB
#define RHO_GATHER_MINB 5            // CTAs per SM the register budget is sized for; measured on C3: 4 -> 2.41 ms, 5 -> 2.23, 6 -> 2.22, 8 -> 2.41
#endif
__global__ void __launch_bounds__(GATHER_THREADS, RHO_GATHER_MINB)
k_gather(const float* __restrict__ x, const int64_t* __restrict__ seg_off, const SegState* __restrict__ seg,
         const SegSpan* __restrict__ span, ItemState* __restrict__ item,
         float* __restrict__ y, const int64_t* __restrict__ y_off, int fade, int tiles) {
#if RHO_GATHER_TILE_MAJOR
  const int s = (int)(blockIdx.x / (unsigned)tiles);                                              // tile index runs fastest
  const int tile = (int)(blockIdx.x - (unsigned)s * (unsigned)tiles);
#else
  const int s = (int)(blockIdx.x % (unsigned)(gridDim.x / (unsigned)tiles));                       // segment runs fastest
  const int tile = (int)(blockIdx.x / (unsigned)(gridDim.x / (unsigned)tiles));
#endif
  const SegSpan sp = span[s];
  const int span_len = sp.ov + sp.body + sp.pause;
  const int jt0 = tile * GATHER_TILE;
  if (jt0 >= span_len) return;
  const int out_len = sp.out_len;
  const int third = out_len / 3;
  const float dc = sp.dc;
  const float* __restrict__ xc = x + sp.x_base;
  float* __restrict__ yo = y + sp.y_base;
  // previous segment (crossfade tail)
  const float* __restrict__ xp = x + sp.prev_x;
  const float dcp = sp.dcp;
  const bool need_fade = fade > 0 && out_len >= 2 * fade;
  const bool aligned = (((uintptr_t)xc | (uintptr_t)yo) & 15u) == 0;

  float a_first = 0.f, a_last = 0.f;
  constexpr int U = GATHER_CHUNK / (4 * GATHER_THREADS);     // 128-bit pieces per thread and pass
  for (int j0 = jt0; j0 < jt0 + GATHER_TILE && j0 < span_len; j0 += GATHER_CHUNK) {
  const int jend = min(j0 + GATHER_CHUNK, span_len);
  {
    // Interior pass (the bulk of every segment): plain body samples, no fade, one decay zone, aligned.  A lean
    // loop -- 8 independent 128-bit loads per thread in flight, no per-piece predicates -- because this stream
    // is bound by the bytes in flight per SM, not by anything it computes.
    const int o0 = sp.dst + j0;
    const bool zone_first = o0 + GATHER_CHUNK <= third, zone_last = o0 >= out_len - third;
    const bool zone_mid = o0 >= third && o0 + GATHER_CHUNK <= out_len - third;
    if (aligned && jend - j0 == GATHER_CHUNK && j0 >= sp.ov && j0 + GATHER_CHUNK <= sp.ov + sp.body &&
        (!need_fade || (o0 >= fade && o0 + GATHER_CHUNK <= out_len - fade)) && (zone_first || zone_last || zone_mid)) {
      float4 w[U];
#pragma unroll
      for (int u = 0; u < U; ++u) w[u] = ldg_stream4(xc + j0 + 4 * (threadIdx.x + u * GATHER_THREADS));
      float ss = 0.f;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        w[u].x = __fsub_rn(w[u].x, dc); w[u].y = __fsub_rn(w[u].y, dc);
        w[u].z = __fsub_rn(w[u].z, dc); w[u].w = __fsub_rn(w[u].w, dc);
        stg_stream4(yo + j0 + 4 * (threadIdx.x + u * GATHER_THREADS), w[u]);
        ss += w[u].x * w[u].x + w[u].y * w[u].y + w[u].z * w[u].z + w[u].w * w[u].w;
      }
      if (zone_first) a_first += ss;
      if (zone_last) a_last += ss;
      continue;
    }
  }
  // Generic pass (span edges, crossfades, pauses, fades, decay-zone boundaries, unaligned segments): piece by
  // piece, no batching -- it runs on a few passes per segment only.
  for (int u = 0; u < U; ++u) {
    const int j = j0 + 4 * (threadIdx.x + u * GATHER_THREADS);
    if (j >= jend) break;
    const int o = sp.dst + j;
    const bool interior = aligned && j + 3 < jend && j >= sp.ov && j + 3 < sp.ov + sp.body &&
                          (!need_fade || (o >= fade && o + 3 < out_len - fade));
    if (interior) {
      float4 w = ldg_stream4(xc + j);
      w.x = __fsub_rn(w.x, dc); w.y = __fsub_rn(w.y, dc); w.z = __fsub_rn(w.z, dc); w.w = __fsub_rn(w.w, dc);
      stg_stream4(yo + j, w);
      if (o < third || o + 3 >= out_len - third) {
        const float vv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (o + k < third) a_first += vv[k] * vv[k];
          if (o + k >= out_len - third) a_last += vv[k] * vv[k];
        }
      }
    } else {
      gather_edge_piece(sp.dst, sp.ov, sp.body, xc, xp, yo, j, jend, dc, dcp, need_fade, fade, out_len, third,
                        &a_first, &a_last);
    }
  }
  }
  // block reduction of the two decay sums, one double atomic each per CTA
  __shared__ double red[2][GATHER_THREADS / 32];
  double df = warp_sum((double)a_first), dl = warp_sum((double)a_last);
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][w] = df; red[1][w] = dl; }
  __syncthreads();
  if (w == 0) {
    df = lane < GATHER_THREADS / 32 ? red[0][lane] : 0.0;
    dl = lane < GATHER_THREADS / 32 ? red[1][lane] : 0.0;
    df = warp_sum(df); dl = warp_sum(dl);
    if (lane == 0) {
      if (df != 0.0) atomicAdd(&item[sp.item].s_first, df);
      if (dl != 0.0) atomicAdd(&item[sp.item].s_last, dl);
    }
  }
}

// ------------------------------------------------------------------ per-item finalize (decay)
// decay_decide() and finalize_item() live in records_dev.cuh: the fused path runs them inside k_logmel_norm.
__global__ void __launch_bounds__(256)
k_finalize_items(const SegState* __restrict__ seg, const ItemState* __restrict__ item,
                 const int32_t* __restrict__ item_first_seg, int n_items, double decay_thr,
                 rho_record* __restrict__ rec, const float* __restrict__ emb, const float* __restrict__ ref, int dim) {
  const int it = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (it >= n_items) return;
  finalize_item(seg, item, item_first_seg, it, lane, decay_thr, rec, emb, ref, dim, RecordPeers{});
}

// ------------------------------------------------------------------ single-clip helpers (method shim)
__global__ void k_sum_single(const float* __restrict__ x, long long n, double* __restrict__ out) {
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    acc += (double)x[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0 && acc != 0.0) atomicAdd(out, acc);
}
__global__ void k_sub_single(float* __restrict__ x, long long n, const double* __restrict__ sum, float* __restrict__ dc_out) {
  const float dc = (float)(*sum / (double)n);
  if (dc_out && blockIdx.x == 0 && threadIdx.x == 0) *dc_out = dc;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    x[i] = __fsub_rn(x[i], dc);
}
__global__ void k_fade_single(float* __restrict__ x, long long n, int fade, int fade_in, int fade_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= fade) return;
  if (fade_in) x[i] = __fmul_rn(x[i], 0.5f * (1.f - cosf(linspace32(0.f, RHO_PI_F, fade, i))));
  if (fade_out) {
    const long long o = n - fade + i;
    x[o] = __fmul_rn(x[o], 0.5f * (1.f + cosf(linspace32(0.f, RHO_PI_F, fade, i))));
  }
}
__global__ void k_decay_sums_single(const float* __restrict__ x, long long n, double* __restrict__ out2) {
  const long long third = n / 3;
  double a = 0.0, b = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < third; i += (long long)gridDim.x * blockDim.x) {
    const float u = x[i], v = x[n - third + i];
    a += (double)(u * u); b += (double)(v * v);
  }
  a = warp_sum(a); b = warp_sum(b);
  if ((threadIdx.x & 31) == 0) { if (a != 0.0) atomicAdd(out2, a); if (b != 0.0) atomicAdd(out2 + 1, b); }
}
__global__ void k_decay_final_single(const double* __restrict__ sums, long long n, double thr, rho_record* __restrict__ rec) {
  rho_record r = *rec;
  decay_decide(sums[0], sums[1], (int)n, thr, &r.first_rms, &r.last_rms, &r.decay_ratio, &r.ok);
  r.out_len = (int)n;
  *rec = r;
}

// ------------------------------------------------------------------ batched decay check on finished audio
// _validate_sound_decay (base_tts.py:297-323) for n clips that are already final -- e.g. after the Qwen loudness
// hook, which runs between the join and the decay check (base_tts.py:911-926).  Pass 1: sum y^2 over the first and
// the last third (16384 samples per CTA, double atomics per clip); pass 2: the decision, written into the records.
__global__ void __launch_bounds__(256)
k_decay_sums_batch(const float* __restrict__ y, const int64_t* __restrict__ off, const char* __restrict__ len_base,
                   int len_stride, double* __restrict__ sums, int tile) {
  const int c = blockIdx.x;
  const int n = *reinterpret_cast<const int32_t*>(len_base + (size_t)c * len_stride);
  const int third = n / 3;
  const long long t0 = (long long)blockIdx.y * tile;
  if (third < 1 || t0 >= n) return;
  // the two thirds are [0, third) and [n - third, n); a tile in the middle has nothing to do
  const long long t1 = min((long long)n, t0 + tile);
  if (t0 >= third && t1 <= n - third) return;
  const float* __restrict__ p = y + off[c];
  float a = 0.f, b = 0.f;
  for (long long i = t0 + threadIdx.x; i < t1; i += 256) {
    const float v = p[i], q = v * v;
    if (i < third) a += q;
    if (i >= n - third) b += q;
  }
  __shared__ double red[2][8];
  double da = warp_sum((double)a), db = warp_sum((double)b);
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][w] = da; red[1][w] = db; }
  __syncthreads();
  if (w == 0) {
    da = lane < 8 ? red[0][lane] : 0.0; db = lane < 8 ? red[1][lane] : 0.0;
    da = warp_sum(da); db = warp_sum(db);
    if (lane == 0) {
      if (da != 0.0) atomicAdd(&sums[2 * c], da);
      if (db != 0.0) atomicAdd(&sums[2 * c + 1], db);
    }
  }
}

__global__ void k_decay_final_batch(const double* __restrict__ sums, const char* __restrict__ len_base, int len_stride,
                                    int n_clips, double thr, rho_record* __restrict__ rec) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_clips) return;
  const int n = *reinterpret_cast<const int32_t*>(len_base + (size_t)c * len_stride);
  rho_record r = rec[c];
  decay_decide(sums[2 * c], sums[2 * c + 1], n, thr, &r.first_rms, &r.last_rms, &r.decay_ratio, &r.ok);
  rec[c] = r;
}

cudaError_t launch_sound_decay_batch(const float* y, const int64_t* off, const int32_t* len, int len_stride_bytes,
                                     int n, int64_t max_len, double thr, rho_record* rec, double* sums,
                                     cudaStream_t st, LaunchCtx* lc) {
  if (n <= 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(sums, 0, 2 * sizeof(double) * (size_t)n, st);
  if (e != cudaSuccess) return e;
  const char* lb = reinterpret_cast<const char*>(len);
  const int ls = len_stride_bytes ? len_stride_bytes : (int)sizeof(int32_t);
  const int tile = 16384;
  const unsigned tiles = (unsigned)((max_len + tile - 1) / tile);
  if (tiles > 0) {
    if (tiles > 65535u) return cudaErrorInvalidValue;
    lc->begin(KID_SINGLE, st);
    k_decay_sums_batch<<<dim3((unsigned)n, tiles), 256, 0, st>>>(y, off, lb, ls, sums, tile);
    lc->end(st);
  }
  lc->begin(KID_SINGLE, st);
  k_decay_final_batch<<<(n + 127) / 128, 128, 0, st>>>(sums, lb, ls, n, thr, rec);
  lc->end(st);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ 16-bit PCM payload of _save_wav
// The in-tree WAV writer (base_tts.py:661-667): (np.clip(x, -1, 1) * 32767).astype(np.int16) -- an fp32 product,
// truncated toward zero.  8 samples per thread: two 128-bit loads, one 128-bit store.
__global__ void __launch_bounds__(256)
k_pcm16(const float* __restrict__ y, const int64_t* __restrict__ off, const char* __restrict__ len_base, int len_stride,
        int16_t* __restrict__ out, const int64_t* __restrict__ out_off) {
  const int c = blockIdx.y;
  const long long n = *reinterpret_cast<const int32_t*>(len_base + (size_t)c * len_stride);
  const float* __restrict__ src = y + off[c];
  int16_t* __restrict__ dst = out + out_off[c];
  auto q = [](float v) { return (int16_t)__float2int_rz(__fmul_rn(fminf(fmaxf(v, -1.0f), 1.0f), 32767.0f)); };
  const long long i0 = ((long long)blockIdx.x * 256 + threadIdx.x) * 8;
  if (i0 >= n) return;
  const bool vec = i0 + 8 <= n && ((reinterpret_cast<uintptr_t>(src + i0) & 15u) == 0) &&
                   ((reinterpret_cast<uintptr_t>(dst + i0) & 15u) == 0);
  if (vec) {
    const float4 a = *reinterpret_cast<const float4*>(src + i0), b = *reinterpret_cast<const float4*>(src + i0 + 4);
    union { int16_t h[8]; uint4 u; } p;
    p.h[0] = q(a.x); p.h[1] = q(a.y); p.h[2] = q(a.z); p.h[3] = q(a.w);
    p.h[4] = q(b.x); p.h[5] = q(b.y); p.h[6] = q(b.z); p.h[7] = q(b.w);
    *reinterpret_cast<uint4*>(dst + i0) = p.u;
  } else {
    for (long long i = i0; i < n && i < i0 + 8; ++i) dst[i] = q(src[i]);
  }
}

cudaError_t launch_pcm16(const float* y, const int64_t* off, const int32_t* len, int len_stride_bytes, int n,
                         int64_t max_len, int16_t* out, const int64_t* out_off, cudaStream_t st, LaunchCtx* lc) {
  if (n <= 0 || max_len <= 0) return cudaSuccess;
  if (n > 65535) return cudaErrorInvalidValue;
  const unsigned gx = (unsigned)((max_len + 2047) / 2048);
  lc->begin(KID_SINGLE, st);
  k_pcm16<<<dim3(gx, (unsigned)n), 256, 0, st>>>(y, off, reinterpret_cast<const char*>(len),
                                                len_stride_bytes ? len_stride_bytes : (int)sizeof(int32_t), out, out_off);
  lc->end(st);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ host launchers
#ifndef RHO_SCAN_DENSE
#define RHO_SCAN_DENSE 1
#endif
// VEC: rows are read with 128-bit LDS, conflict-free when the row stride is 4 * odd.  A dense stride (== hop)
// lets the whole tile arrive in one bulk copy; it is usable when hop/4 is odd (already conflict-free) or
// hop/4 = 2 mod 4 (conflict-free with the one-step skew of k_scan) -- 24 kHz: hop 120, 120/4 = 30.
static inline int scan_row_stride(int hop, bool vec, int* skew) {
  *skew = 0;
  if (vec) {
    int r4 = hop >> 2;
    if (r4 & 1) return hop;
    if (RHO_SCAN_DENSE && (r4 & 3) == 2) { *skew = 1; return hop; }
    return (r4 + 1) << 2;
  }
  return hop | 1;
}

cudaError_t launch_scan(const float* x, const int64_t* off, const int32_t* len, int n_seg, int64_t max_len,
                        const Derived& d, const Workspace& ws, cudaStream_t st, LaunchCtx* lc) {
  if (n_seg <= 0) return cudaSuccess;
  const bool vec = (d.hop % 4 == 0) && (d.window == 2 * d.hop);
  int skew = 0;
  const int RS = scan_row_stride(d.hop, vec, &skew);
  const size_t smem = (size_t)(SCAN_FR + 2) * RS * sizeof(float);
  const int64_t max_frames = max_len <= 0 ? 1 : (max_len + 2 * d.hop - d.window) / d.hop + 1;
  const unsigned tiles = (unsigned)((max_frames + SCAN_FR - 1) / SCAN_FR);
  dim3 grid((unsigned)n_seg, tiles ? tiles : 1u);
  cudaError_t e;
  if (vec) {
    e = cudaFuncSetAttribute(k_scan<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    lc->begin(KID_SCAN, st);
    k_scan<true><<<grid, SCAN_FR, smem, st>>>(x, off, len, ws.seg, ws.block_sum, ws.blocks_per_seg,
                                             d.window, d.hop, RS, skew, d.thr);
  } else {
    e = cudaFuncSetAttribute(k_scan<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    lc->begin(KID_SCAN, st);
    k_scan<false><<<grid, SCAN_FR, smem, st>>>(x, off, len, ws.seg, ws.block_sum, ws.blocks_per_seg,
                                              d.window, d.hop, RS, skew, d.thr);
  }
  lc->end(st);
  return cudaGetLastError();
}

cudaError_t launch_trim_scan(const float* x, const int64_t* off, const int32_t* len, const uint8_t* trim_flags,
                             int n_seg, int64_t max_len, const Derived& d, const Workspace& ws,
                             rho_seg_info* info, cudaStream_t st, LaunchCtx* lc) {
  if (n_seg <= 0) return cudaSuccess;
  lc->begin(KID_INIT, st);
  k_init_segs<<<(n_seg + 255) / 256, 256, 0, st>>>(ws.seg, trim_flags, n_seg); lc->end(st);
  cudaError_t e = launch_scan(x, off, len, n_seg, max_len, d, ws, st, lc);
  if (e != cudaSuccess) return e;
  lc->begin(KID_FINALIZE_SEGS, st);
  k_finalize_segs<<<(n_seg * 32 + 255) / 256, 256, 0, st>>>(x, off, len, ws.seg, ws.block_sum, ws.blocks_per_seg,
                                                          n_seg, d.window, d.hop, d.trim_enabled, info, nullptr);
  lc->end(st);
  return cudaGetLastError();
}

cudaError_t launch_join(const float* x, const int64_t* seg_off, const int32_t* seg_len, int n_seg, int64_t max_seg_len,
                        const int32_t* item_first_seg, int n_items, int64_t max_item_len,
                        const Derived& d, float* y, const int64_t* y_off, rho_record* rec, rho_seg_info* seg_info,
                        const Workspace& ws, cudaStream_t st, LaunchCtx* lc, int stages,
                        const float* emb, const float* ref_emb, int emb_dim) {
  if (n_items <= 0) return cudaSuccess;
  cudaError_t e = cudaSuccess;
  if (stages & JOIN_PREPARE) {
    lc->begin(KID_INIT, st);
    k_init_items<<<(n_items + 255) / 256, 256, 0, st>>>(ws.seg, ws.item, item_first_seg, n_items,
                                                        (stages & JOIN_INIT_FEATURES) ? ws.clip_max : nullptr,
                                                        (stages & JOIN_INIT_FEATURES) ? ws.tiles_done : nullptr,
                                                        (stages & JOIN_INIT_FEATURES) ? ws.work_counter : nullptr);
    lc->end(st);
    if (n_seg > 0) {
      e = launch_scan(x, seg_off, seg_len, n_seg, max_seg_len, d, ws, st, lc);
      if (e != cudaSuccess) return e;
      lc->begin(KID_FINALIZE_SEGS, st);
      k_finalize_segs<<<(n_seg * 32 + 255) / 256, 256, 0, st>>>(x, seg_off, seg_len, ws.seg, ws.block_sum,
                                                              ws.blocks_per_seg, n_seg, d.window, d.hop,
                                                              d.trim_enabled, seg_info,
                                                              (stages & JOIN_ONE_SEG_ITEMS) ? ws.item : nullptr);
      lc->end(st);
    }
    if (!(stages & JOIN_ONE_SEG_ITEMS)) {                // one-segment items were planned by k_finalize_segs
      lc->begin(KID_PLAN, st);
      k_plan_items<<<(n_items + 127) / 128, 128, 0, st>>>(ws.seg, seg_len, ws.span, ws.item, item_first_seg, n_items,
                                                         d.cf, d.pause, d.pause_on, seg_off, y_off);
      lc->end(st);
    }
  }
  if ((stages & JOIN_GATHER) && n_seg > 0) {
    // a segment's span is at most its own length plus one pause
    const int64_t max_span = max_seg_len + d.pause;
    const unsigned tiles = (unsigned)((max_span + GATHER_TILE - 1) / GATHER_TILE);
    const unsigned tl = tiles ? tiles : 1u;
    if ((uint64_t)n_seg * tl > 0x7fffffffull) return cudaErrorInvalidValue;
    lc->begin(KID_GATHER, st);
    k_gather<<<(unsigned)n_seg * tl, GATHER_THREADS, 0, st>>>(x, seg_off, ws.seg, ws.span, ws.item, y, y_off, d.fade,
                                                              (int)tl);
    lc->end(st);
  }
  (void)max_item_len;
  if (stages & JOIN_FINISH) {
    lc->begin(KID_FINALIZE_ITEMS, st);
    k_finalize_items<<<(n_items * 32 + 255) / 256, 256, 0, st>>>(ws.seg, ws.item, item_first_seg, n_items, d.decay_thr,
                                                            rec, emb, ref_emb, emb_dim);
    lc->end(st);
  }
  return cudaGetLastError();
}

cudaError_t launch_remove_dc(float* x, int64_t n, float* dc_out, double* scratch, cudaStream_t st, LaunchCtx* lc) {
  if (n <= 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(scratch, 0, sizeof(double), st);
  if (e != cudaSuccess) return e;
  const int blocks = (int)((n + 4095) / 4096 < 1184 ? (n + 4095) / 4096 : 1184);
  lc->begin(KID_SINGLE, st);
  k_sum_single<<<blocks, 256, 0, st>>>(x, n, scratch); lc->end(st);
  lc->begin(KID_SINGLE, st);
  k_sub_single<<<blocks, 256, 0, st>>>(x, n, scratch, dc_out); lc->end(st);
  return cudaGetLastError();
}

cudaError_t launch_apply_fades(float* x, int64_t n, int fade, int fade_in, int fade_out, cudaStream_t st, LaunchCtx* lc) {
  if (n <= 0 || fade <= 0 || n < 2 * (int64_t)fade || (!fade_in && !fade_out)) return cudaSuccess;
  lc->begin(KID_SINGLE, st);
  k_fade_single<<<(fade + 127) / 128, 128, 0, st>>>(x, n, fade, fade_in, fade_out); lc->end(st);
  return cudaGetLastError();
}

cudaError_t launch_sound_decay(const float* x, int64_t n, double thr, rho_record* rec, double* scratch,
                               cudaStream_t st, LaunchCtx* lc) {
  cudaError_t e = cudaMemsetAsync(scratch, 0, 2 * sizeof(double), st);
  if (e != cudaSuccess) return e;
  if (n >= 3) {
    const int64_t third = n / 3;
    const int blocks = (int)((third + 2047) / 2048 < 1184 ? (third + 2047) / 2048 : 1184);
    lc->begin(KID_SINGLE, st);
    k_decay_sums_single<<<blocks, 256, 0, st>>>(x, n, scratch); lc->end(st);
  }
  lc->begin(KID_SINGLE, st);
  k_decay_final_single<<<1, 1, 0, st>>>(scratch, n, thr, rec); lc->end(st);
  return cudaGetLastError();
}

}  // namespace rho
