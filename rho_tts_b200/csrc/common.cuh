// Shared device/host helpers for librho_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <limits.h>
#include "../../include/rho_b200.h"

namespace rho {

// ---------------------------------------------------------------- derived constants
// Integer/threshold constants the reference derives at call time
// (base_tts.py:366-367, 420, 455, 519).  Computed on the host in double exactly as
// Python does (int() truncation, 10**(dB/20) via pow) and passed by value to kernels.
struct Derived {
  int window;      // int(sr*0.01)
  int hop;         // window/2  (== avg_pool1d stride == padding)
  int fade;        // int(sr*fade_sec)
  int cf;          // int(sr*xfade_sec)
  int pause;       // int(sr*pause_sec)
  int pause_on;    // pause_sec > 0
  int trim_enabled;
  float thr;       // fp32(10**(dB/20))
  double decay_thr;
};

Derived derive(const rho_params& p);

// ---------------------------------------------------------------- workspace layout
// Per-segment scan state (32 B).
struct SegState {
  int first;       // min loud frame (INT_MAX if none)
  int last;        // max loud frame (-1 if none)
  int start, end;  // trim result
  float dc;
  uint32_t flags;
  int item;        // owning item
  int pos;         // position inside the item
};

// Per-segment output span inside its item, produced by plan_items, plus everything k_gather needs to start
// loading samples after ONE dependent fetch (it used to chase seg -> item -> offsets: three round trips per CTA).
struct SegSpan {
  int dst;         // first output sample of this segment's span (relative to the item)
  int ov;          // crossfade length with the previous segment (0: none)
  int body;        // samples copied after the crossfade (from processed offset `ov`)
  int pause;       // zeros appended after the body
  int prev_tail;   // processed-index in the previous segment where the crossfade tail starts
  int item;        // owning item
  int out_len;     // length of the whole item
  uint32_t item_flags;
  float dc;        // DC of this segment (0 in the fallback)
  float dcp;       // DC of the previous segment (crossfade tail)
  int pad0, pad1;
  long long x_base;   // offset in x of processed sample 0 of this segment (seg_off + trim start; fallback: seg_off)
  long long prev_x;   // offset in x of the first crossfade-tail sample of the previous segment
  long long y_base;   // offset in y of this segment's span (y_off[item] + dst)
};

// Per-item state (32 B).
struct ItemState {
  double s_first;  // sum y^2 over the first third
  double s_last;   // sum y^2 over the last third
  int out_len;
  uint32_t flags;
  int pad0, pad1;
};

struct Workspace {
  SegState* seg;        // [n_segments]
  SegSpan* span;        // [n_segments]
  ItemState* item;      // [n_items]
  float* block_sum;     // [n_segments * blocks_per_seg]  sum of x per hop block
  int blocks_per_seg;
  int* clip_max;        // [n_items] running max of log10(mel) as ordered int
  int32_t* len16;       // [n_items]
  int* tiles_done;      // [n_items] tiles of the fused kernel that finished (last finisher normalises)
  int* work_counter;    // [1] next unclaimed tile of the persistent fused kernel
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_min(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_max(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// streaming 128-bit load: read once, do not keep in L1.  Not volatile: the compiler may batch
// several of these ahead of their uses (memory-level parallelism).
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  float4 r;
  asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
      : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
// 16-byte asynchronous global->shared copy (LDGSTS); bytes beyond `src_bytes` are zero-filled.
__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gmem_src, int src_bytes) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(d), "l"(gmem_src), "r"(src_bytes) : "memory");
}
// ---- bulk asynchronous copies (TMA engine, SASS UBLKCP) completing on an mbarrier ----------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// one arrival that also announces `bytes` of asynchronous-copy traffic for the current phase
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// wait until the phase with the given parity has completed; traps instead of hanging the GPU forever
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  const unsigned a = smem_u32(bar);
  unsigned done = 0;
  for (unsigned spin = 0; !done; ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
    if (spin > (1u << 24)) __trap();
  }
}

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void stg_stream4(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// float <-> monotonically ordered int (for atomicMax on floats, negative values included)
__device__ __forceinline__ int float_to_ordered(float f) {
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int i) {
  return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff);
}

// torch.linspace(a, b, n)[i] in fp32: step in fp32, first half counted from a, second from b.
__device__ __forceinline__ float linspace32(float a, float b, int n, int i) {
  if (n <= 1) return a;
  const float step = __fdiv_rn(__fsub_rn(b, a), (float)(n - 1));
  return (i < n / 2) ? __fadd_rn(a, __fmul_rn(step, (float)i))
                     : __fsub_rn(b, __fmul_rn(step, (float)(n - 1 - i)));
}
#define RHO_PI_F 3.14159274101257324f       /* fp32(pi)   */
#define RHO_HALF_PI_F 1.57079637050628662f  /* fp32(pi/2) */
#endif  // __CUDACC__

}  // namespace rho
