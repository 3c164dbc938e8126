// Record assembly shared by k_finalize_items (join.cu) and k_logmel_norm (logmel.cu, fused path).
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace rho {

__device__ __forceinline__ void decay_decide(double s_first, double s_last, int n, double thr,
                                             float* first_rms, float* last_rms, double* ratio, int* ok) {
  // _validate_sound_decay :304-323
  *first_rms = 0.f; *last_rms = 0.f; *ratio = 1.0; *ok = 1;
  const int third = n / 3;
  if (n <= 0 || third < 1) return;
  const float fr = __fsqrt_rn((float)(s_first / (double)third));
  const float lr = __fsqrt_rn((float)(s_last / (double)third));
  *first_rms = fr; *last_rms = lr;
  if ((double)fr < 1e-8) return;
  const double r = (double)lr / (double)fr;
  *ratio = r; *ok = (r >= thr) ? 1 : 0;
}

// One warp per item: lane 0 assembles the record; when embeddings are given the warp also computes the speaker
// cosine (base_tts.py:341-344) so the validate path needs no separate k_cosine launch.
__device__ __forceinline__ void finalize_item(const SegState* __restrict__ seg, const ItemState* __restrict__ item,
                                              const int32_t* __restrict__ item_first_seg, int it, int lane,
                                              double decay_thr, rho_record* __restrict__ rec,
                                              const float* __restrict__ emb, const float* __restrict__ ref, int dim,
                                              const RecordPeers& peers) {
  float cosv = 0.f;
  if (emb && ref) {
    const float* __restrict__ e = emb + (size_t)it * dim;
    float dot = 0.f, ne = 0.f, nr = 0.f;
    if ((dim & 3) == 0 && ((((uintptr_t)e) | ((uintptr_t)ref)) & 15u) == 0) {
      for (int i = 4 * lane; i < dim; i += 128) {
        const float4 a = *reinterpret_cast<const float4*>(e + i);
        const float4 r = __ldg(reinterpret_cast<const float4*>(ref + i));
        dot += a.x * r.x + a.y * r.y + a.z * r.z + a.w * r.w;
        ne += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
        nr += r.x * r.x + r.y * r.y + r.z * r.z + r.w * r.w;
      }
    } else {
      for (int i = lane; i < dim; i += 32) {
        const float a = e[i], r = __ldg(ref + i);
        dot += a * r; ne += a * a; nr += r * r;
      }
    }
    dot = warp_sum(dot); ne = warp_sum(ne); nr = warp_sum(nr);
    cosv = __fdiv_rn(dot, __fmul_rn(__fsqrt_rn(nr), __fsqrt_rn(ne)));
  }
  if (lane != 0) return;
  const ItemState is = item[it];
  const int s0 = item_first_seg[it], n = item_first_seg[it + 1] - s0;
  rho_record r;
  r.start = 0; r.end = 0; r.dc = 0.f;
  if (n > 0) { const SegState st = seg[s0]; r.start = st.start; r.end = st.end; r.dc = st.dc; }
  r.out_len = is.out_len; r.flags = is.flags; r.cosine = cosv; r.n_segments = n;
  decay_decide(is.s_first, is.s_last, is.out_len, decay_thr, &r.first_rms, &r.last_rms, &r.decay_ratio, &r.ok);
  rec[it] = r;
  // multi-GPU: the same 48 bytes go straight into every rank's gathered buffer (peer-mapped memory, NVLink stores);
  // the stores are ordered for the readers by the end of this kernel + the flag kernel that follows it in the stream
  for (int q = 0; q < peers.n; ++q) peers.sink[q][peers.slot + it] = r;
}

// Called by the thread that stored item records (finalize_item's lane 0) once its CTA has nothing else to do: makes its
// peer stores visible system-wide, counts the record, and -- if it was the last one of this call -- tells every rank
// that this rank's block of epoch `epoch` is complete.
__device__ __forceinline__ void publish_records(const RecordPeers& peers) {
  if (peers.n <= 0) return;
  __threadfence_system();
  if (atomicAdd(peers.done, 1) == peers.n_items - 1) {
    *peers.done = 0;                                     // ready for the next call (stream order)
    __threadfence_system();                              // acquire side of the count: every other CTA's fence happened before
    for (int q = 0; q < peers.n; ++q)
      asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(peers.flag[q] + peers.rank), "r"(peers.epoch) : "memory");
  }
}

}  // namespace rho
