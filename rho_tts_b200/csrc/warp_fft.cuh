// 256-point complex FFT of one warp with the butterflies in registers (shared by pitch.cu and mfcc.cu).
#pragma once
#include "common.cuh"

namespace rho {

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

// 256-point complex FFT of one warp, 8 points per lane, as radix-8 x radix-8 x radix-4 with the butterflies in
// registers: lane l starts with v[r] = x[l + 32 r].
//   1. radix-8 over r, twiddle W256^(l k1)                      -> Y[k1][n2 = l]
//   2. exchange through shared memory (rows of 36 float2: conflict-free both ways); lane (k1 = l >> 2, q = l & 3)
//      takes Y[k1][q + 4 s], radix-8 over s, twiddle W32^(q a)  -> V[k1][a][q]
//   3. radix-4 over q = the four lanes of a quad, two xor-shuffle rounds
// and ends with v[a] = X[k1 + 8 a + 64 b], b = 2 (q & 1) + (q >> 1).  One shared-memory round trip instead of four, no
// strided twiddle reads (the 14 twiddles a lane needs come from two small conflict-free tables, tw1[k1][lane] and
// tw2[a][q]).  INV: conjugated constants and tables.  The caller must __syncwarp() before re-using E.
constexpr int PV_E_LD = 36;                  // row stride of the exchange buffer (float2)
constexpr int PV_E_SIZE = 8 * PV_E_LD;       // 288 float2 per warp; also holds a padded natural-order copy (pv_nat)
__device__ __forceinline__ int pv_nat(int k) { return k + 8 * (k >> 6); }   // natural-order index -> padded position

template <bool INV>
__device__ __forceinline__ void radix4(float2 a0, float2 a1, float2 a2, float2 a3, float2& x0, float2& x1, float2& x2,
                                       float2& x3) {
  const float2 e0 = make_float2(a0.x + a2.x, a0.y + a2.y), e1 = make_float2(a0.x - a2.x, a0.y - a2.y);
  const float2 o0 = make_float2(a1.x + a3.x, a1.y + a3.y), d = make_float2(a1.x - a3.x, a1.y - a3.y);
  const float2 o1 = INV ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);       // d * (+i) / d * (-i)
  x0 = make_float2(e0.x + o0.x, e0.y + o0.y); x1 = make_float2(e1.x + o1.x, e1.y + o1.y);
  x2 = make_float2(e0.x - o0.x, e0.y - o0.y); x3 = make_float2(e1.x - o1.x, e1.y - o1.y);
}

template <bool INV>
__device__ __forceinline__ void radix8(float2 (&v)[8]) {
  const float c = 0.70710678118654752f;
  float2 s[4], d[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    s[i] = make_float2(v[i].x + v[i + 4].x, v[i].y + v[i + 4].y);
    d[i] = make_float2(v[i].x - v[i + 4].x, v[i].y - v[i + 4].y);
  }
  // d[i] *= W8^i
  d[1] = INV ? make_float2(c * (d[1].x - d[1].y), c * (d[1].x + d[1].y)) : make_float2(c * (d[1].x + d[1].y), c * (d[1].y - d[1].x));
  d[2] = INV ? make_float2(-d[2].y, d[2].x) : make_float2(d[2].y, -d[2].x);
  d[3] = INV ? make_float2(-c * (d[3].x + d[3].y), c * (d[3].x - d[3].y)) : make_float2(c * (d[3].y - d[3].x), -c * (d[3].x + d[3].y));
  radix4<INV>(s[0], s[1], s[2], s[3], v[0], v[2], v[4], v[6]);
  radix4<INV>(d[0], d[1], d[2], d[3], v[1], v[3], v[5], v[7]);
}

template <bool INV>
__device__ __forceinline__ void warp_fft256(float2 (&v)[8], float2* __restrict__ E, const float2* __restrict__ tw1,
                                            const float2* __restrict__ tw2, int lane) {
  radix8<INV>(v);
#pragma unroll
  for (int k1 = 1; k1 < 8; ++k1) v[k1] = cmul(v[k1], tw1[k1 * 32 + lane]);
#pragma unroll
  for (int k1 = 0; k1 < 8; ++k1) E[k1 * PV_E_LD + lane] = v[k1];
  __syncwarp();
  const int q = lane & 3;
  const float2* __restrict__ row = E + (lane >> 2) * PV_E_LD + q;
#pragma unroll
  for (int t = 0; t < 8; ++t) v[t] = row[4 * t];
  radix8<INV>(v);
#pragma unroll
  for (int a = 1; a < 8; ++a) v[a] = cmul(v[a], tw2[a * 4 + q]);
  const bool hi = (q & 2) != 0, odd = (q & 1) != 0;
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    float2 own = v[a];
    float2 par = make_float2(__shfl_xor_sync(0xffffffffu, own.x, 2), __shfl_xor_sync(0xffffffffu, own.y, 2));
    own = hi ? make_float2(par.x - own.x, par.y - own.y) : make_float2(own.x + par.x, own.y + par.y);
    par = make_float2(__shfl_xor_sync(0xffffffffu, own.x, 1), __shfl_xor_sync(0xffffffffu, own.y, 1));
    // quad lanes 0 / 1 hold e0 / o0 -> X0 = e0 + o0, X2 = e0 - o0; lanes 2 / 3 hold e1 / o1 -> X1 = e1 + w o1, X3 = e1 - w o1
    const float2 t = odd ? own : par;                                           // the "o" term of this pair
    const float2 wt = !hi ? t : (INV ? make_float2(-t.y, t.x) : make_float2(t.y, -t.x));
    const float2 e = odd ? par : own;                                           // the "e" term
    v[a] = odd ? make_float2(e.x - wt.x, e.y - wt.y) : make_float2(e.x + wt.x, e.y + wt.y);
  }
}

// tw1[k1][lane] = W256^(lane k1), tw2[a][q] = W32^(q a); conjugated for the inverse.  Filled by the whole CTA.
constexpr int PV_TW_SIZE = 8 * 32 + 8 * 4;
template <bool INV>
__device__ __forceinline__ void pv_fill_twiddles(const float2* __restrict__ g_w256, float2* __restrict__ tw) {
  for (int i = threadIdx.x; i < PV_TW_SIZE; i += blockDim.x) {
    const int e = i < 256 ? (i >> 5) * (i & 31) : 8 * ((i - 256) >> 2) * ((i - 256) & 3);
    float2 t = __ldg(g_w256 + e);
    if (INV) t.y = -t.y;
    tw[i] = t;
  }
}

}  // namespace rho
