// Windowed DFT of the Whisper front end as tensor-core GEMMs (tcgen05 / TMEM), in isolation:
//   P[t][k] = | sum_n hann[n] * w16[160 t - 200 + n] * exp(-2 pi i k n / 400) |^2        k <= 200
// = torch.stft(n_fft 400, hop 160, periodic Hann, centre / reflect) followed by |.|^2 of transformers
// feature_extraction_whisper.py:149-158 (the reference reaches it through stt_validator.py:78-107).  BASELINE.json's
// north_star names the windowed-DFT contraction next to the mel projection as the two dense GEMMs of the path; this
// kernel is that contraction on the tensor cores, measured on its own like k_mel_gemm.  The product path keeps the
// shared-memory FFT of fused.cu / logmel.cu (DESIGN.md 3 has the comparison).
//
// The dense 400 x 402 form (321.6 kflop per frame, x3 for fp32-class accuracy) loses to a 17 kflop FFT by an order
// of magnitude, so the DFT is FACTORED 400 = 25 x 16 (n = 16 n1 + n2, k = k1 + 25 k2) into two small GEMMs with a
// twiddle in between -- Cooley-Tukey with each stage as a matrix product:
//   stage 1   Y[k1][n2]  = sum_{n1 < 25} W25^(n1 k1) * (hann * x)[16 n1 + n2]        real input: k1 = 0..12 suffice
//   twiddle   Y'[k1][n2] = Y[k1][n2] * W400^(n2 k1)                                   CUDA cores, between the GEMMs
//   stage 2   X[k1 + 25 k2] = sum_{n2 < 16} W16^(n2 k2) * Y'[k1][n2]                  complex x complex as a real GEMM
// and the bins k1' = 25 - k1 come from X[400 - k] = conj X[k] (only |X|^2 is needed).  Both GEMMs are
// tcgen05.mma.kind::tf32, M128 N32 K8, with the 3xTF32 split (hi = top 19 bits, lo = remainder: Ahi*Bhi + Alo*Bhi +
// Ahi*Blo), A = the data (K-major, 64-byte swizzle, written by the worker threads), B = the constant DFT matrices
// (resident in shared memory), D in TMEM:
//   stage 1: rows (frame, n2) = 16 per frame -> 8 frames per 128-row tile, K = 32 (25 used), N = 32 (k1 re/im, 26 used)
//   stage 2: rows (frame, k1) = 13 per frame -> 104 of 128 rows,            K = 32 (n2 re/im), N = 32 (k2 re/im)
// 24 MMAs per 8 frames (the mel projection: 78 per 32 frames).
//
// One persistent CTA per SM: six groups of four worker warps (one TMEM lane quarter each), every group owns every sixth
// tile -- build A1 -> [MMA stage 1] -> read D1, twiddle, build A2 -> [MMA stage 2] -> read D2, |.|^2, store -- plus one
// MMA-issuing warp; while a group waits for the tensor core (or for its next samples) the others do their CUDA-core part.
#include <cmath>
#include <vector>
#include "logmel_dev.cuh"

namespace rho {

constexpr int ST_FR = 8;                         // frames per tile
constexpr int ST_GROUPS = 6;                     // tiles in flight per SM: the stages are tiny (12 MMAs), what has to be
                                                 // covered is the latency of the two hand-offs per tile (2 groups: 8.2 ms
                                                 // per 1 M frames, tensor pipe 2 % active)
constexpr int ST_WORKERS = 128;                  // threads per group
constexpr int ST_THREADS = ST_GROUPS * ST_WORKERS + 32;
constexpr int ST_K1 = 13;                        // k1 = 0..12
constexpr uint32_t ST_A_BYTES = 128 * 128;       // 128 rows x 32 floats: [K-block of 16 floats][row][64 B]
constexpr uint32_t ST_B_BYTES = 32 * 128;        // 32 rows (N) x 32 floats
constexpr uint32_t ST_TMEM_COLS = 512;           // D1 / D2 of six groups, 32 columns each (384 used)
constexpr uint32_t ST_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t ST_DHI = (512u >> 4) | (1u << 14) | (4u << 29);   // SBO 512 B (8 rows x 64 B), version 1, SWIZZLE_64B

// One operand buffer per group: A1 (stage 1) is dead once stage 1 has been committed, A2 (stage 2) is written into the
// same bytes.  Rows 104..127 keep A1 values during stage 2: finite, and nobody reads those rows of D2.
struct StGroup {
  unsigned char a_hi[ST_A_BYTES], a_lo[ST_A_BYTES];
};
struct alignas(1024) StSmem {
  StGroup g[ST_GROUPS];
  unsigned char b1_hi[ST_B_BYTES], b1_lo[ST_B_BYTES], b2_hi[ST_B_BYTES], b2_lo[ST_B_BYTES];
  float win[N_FFT];                              // [n1][n2] = hann[16 n1 + n2] (the natural order)
  float2 tw[16 * ST_K1];                         // [n2][k1] = W400^(n2 k1)
  uint64_t a1_ready[ST_GROUPS], d1_full[ST_GROUPS], a2_ready[ST_GROUPS], d2_full[ST_GROUPS];
  uint32_t tmem_base;
};
constexpr size_t ST_TABLE_BYTES = 4 * ST_B_BYTES + sizeof(float) * N_FFT + sizeof(float2) * 16 * ST_K1;
static_assert(sizeof(StSmem) + 1024 <= 232448, "shared memory budget");

// byte offset of element (row r, column k) of a K-major operand with `rows` rows, 64-byte swizzle: K-blocks of 16 floats,
// 64-byte rows, the 16-byte chunk index XORed with bits [7, 9) of the address
__host__ __device__ __forceinline__ uint32_t st_off(uint32_t rows, uint32_t r, uint32_t k) {
  return (k >> 4) * rows * 64u + r * 64u + ((((k >> 2) ^ (r >> 1)) & 3u) << 4) + (k & 3u) * 4u;
}
// nearest TF32 value (10 explicit mantissa bits): with hi rounded to nearest, |v - hi| <= 2^-12 |v| and the two halves
// together carry 23 bits (truncation: 22)
__device__ __forceinline__ float st_hi(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ void st_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void st_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void st_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t st_desc(uint32_t smem_addr) {
  return (uint64_t)(((smem_addr >> 4) & 0x3FFFu) | (1u << 16)) | ((uint64_t)ST_DHI << 32);
}
__device__ __forceinline__ void st_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(ST_IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void st_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t st_elect() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, unsigned parity) {
  unsigned done;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return done != 0;
}
__device__ __forceinline__ void st_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(ST_THREADS, 1)
k_stft_tc(const float* __restrict__ x16, const int64_t* __restrict__ off, const int32_t* __restrict__ len16, int pad_frames,
          const int4* __restrict__ tiles, int n_tiles, const unsigned char* __restrict__ tables,
          float* __restrict__ power, long long ld_power) {
  extern __shared__ unsigned char st_raw[];
  StSmem& S = *reinterpret_cast<StSmem*>(st_raw + ((1024u - (smem_u32(st_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5;

  // constant operands and tables: one copy per CTA
  for (uint32_t i = threadIdx.x; i < ST_TABLE_BYTES / 16; i += ST_THREADS)
    reinterpret_cast<uint4*>(S.b1_hi)[i] = reinterpret_cast<const uint4*>(tables)[i];
  if (threadIdx.x == 0) {
    for (int g = 0; g < ST_GROUPS; ++g) {
      mbar_init(&S.a1_ready[g], ST_WORKERS); mbar_init(&S.a2_ready[g], ST_WORKERS);
      mbar_init(&S.d1_full[g], 1); mbar_init(&S.d2_full[g], 1);
    }
    mbar_fence_init();
  }
  constexpr int MMA_WARP = ST_GROUPS * 4;
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&S.tmem_base)), "r"(ST_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  st_fence_before();
  __syncthreads();
  st_fence_after();
  const uint32_t tmem = S.tmem_base;
  const int my_tiles = (n_tiles > (int)blockIdx.x) ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == MMA_WARP) {
    // ================================ MMA issuer
    auto stage = [&](int g, int which) {
      StGroup& G = S.g[g];
      const uint32_t d = tmem + (uint32_t)(64 * g + 32 * which);
      const uint32_t a_hi = smem_u32(G.a_hi), a_lo = smem_u32(G.a_lo);
      const uint32_t b_hi = smem_u32(which ? S.b2_hi : S.b1_hi), b_lo = smem_u32(which ? S.b2_lo : S.b1_lo);
      if (st_elect()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t ao = (uint32_t)((ks >> 1) * 128 * 64 + (ks & 1) * 32), bo = (uint32_t)((ks >> 1) * 32 * 64 + (ks & 1) * 32);
          st_mma(d, st_desc(a_hi + ao), st_desc(b_hi + bo), ks > 0);
          st_mma(d, st_desc(a_lo + ao), st_desc(b_hi + bo), 1);
          st_mma(d, st_desc(a_hi + ao), st_desc(b_lo + bo), 1);
        }
        st_commit(which ? &S.d2_full[g] : &S.d1_full[g]);
      }
      __syncwarp();
    };
    // Out-of-order service: whichever group has an operand ready gets its stage next (polling the groups' barriers in
    // turn); a group alternates stage 1 / stage 2, so `done[g]` stages of it have been issued and the barrier it waits
    // on and its phase follow from that count.
    int done[ST_GROUPS], need[ST_GROUPS], left = 0;
#pragma unroll
    for (int g = 0; g < ST_GROUPS; ++g) {
      done[g] = 0;
      need[g] = 2 * ((my_tiles > g) ? (my_tiles - g + ST_GROUPS - 1) / ST_GROUPS : 0);
      left += need[g];
    }
    while (left > 0) {
#pragma unroll
      for (int g = 0; g < ST_GROUPS; ++g) {
        if (done[g] < need[g]) {
          const int which = done[g] & 1;
          const unsigned parity = (unsigned)(done[g] >> 1) & 1u;
          if (mbar_test(which ? &S.a2_ready[g] : &S.a1_ready[g], parity)) {
            st_fence_after();
            stage(g, which);
            ++done[g]; --left;
          }
        }
      }
    }
  } else {
    // ================================ workers
    const int g = warp >> 2, w = threadIdx.x & (ST_WORKERS - 1), q = warp & 3;
    StGroup& G = S.g[g];
    const uint32_t lane_addr = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(64 * g);
    const int f1 = w >> 4, n2 = w & 15;                       // stage-1 row: (frame, n2)
    const int f2 = w / ST_K1, k1 = w - f2 * ST_K1;            // stage-2 row: (frame, k1), valid for w < 104
    // the samples of this thread's row of a tile (the 25 loads are independent: one round trip to HBM)
    auto load_row = [&](const int4 td, float (&v)[25]) {
      const float* __restrict__ xs = x16 + off[td.x];
      int T, T_real, N, n_valid;
      lm_frame_counts(len16[td.x], pad_frames, &T, &T_real, &N, &n_valid);
      const long long s0 = (long long)HOP16 * (td.y + f1) - N_FFT / 2 + n2;
#pragma unroll
      for (int n1 = 0; n1 < 25; ++n1) v[n1] = (f1 < td.z) ? lm_sample(xs, s0 + 16 * n1, n_valid, N) : 0.f;
    };
    float v[25];
    int4 td = make_int4(0, 0, 0, 0);
    if (g < my_tiles) { td = tiles[blockIdx.x + (long long)g * gridDim.x]; load_row(td, v); }
    for (int it = g; it < my_tiles; it += ST_GROUPS) {
      const unsigned parity = (unsigned)(it / ST_GROUPS) & 1u;
      const int4 cur = td;
      // ---- A1: the windowed samples of row (frame, n2), n1 = 0..24 (+ 7 zero columns), as hi / lo TF32 halves
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float e[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) e[c] = (4 * j + c < 25) ? v[4 * j + c] * S.win[16 * (4 * j + c) + n2] : 0.f;
        const float4 h = make_float4(st_hi(e[0]), st_hi(e[1]), st_hi(e[2]), st_hi(e[3]));
        const uint32_t o = st_off(128, (uint32_t)w, (uint32_t)(4 * j));
        *reinterpret_cast<float4*>(G.a_hi + o) = h;
        *reinterpret_cast<float4*>(G.a_lo + o) = make_float4(st_hi(e[0] - h.x), st_hi(e[1] - h.y), st_hi(e[2] - h.z), st_hi(e[3] - h.w));
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      st_arrive(&S.a1_ready[g]);
      // the next tile's samples travel while the tensor core and the twiddle step work on this one
      if (it + ST_GROUPS < my_tiles) { td = tiles[blockIdx.x + (long long)(it + ST_GROUPS) * gridDim.x]; load_row(td, v); }
      // ---- twiddle: D1 row (frame, n2) holds Y[k1] as (re, im) pairs -> Y' -> A2[(frame, k1)][(n2, re / im)]
      mbar_wait(&S.d1_full[g], parity);
      st_fence_after();
      {
        const float2* __restrict__ tw = S.tw + n2 * ST_K1;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t r[16];
          st_ld16(lane_addr + 16 * half, r);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const int k = 8 * half + kk;
            if (k < ST_K1) {
              const float2 y = make_float2(__uint_as_float(r[2 * kk]), __uint_as_float(r[2 * kk + 1]));
              const float2 t = tw[k];
              const float2 z = make_float2(y.x * t.x - y.y * t.y, y.x * t.y + y.y * t.x);
              const float2 h = make_float2(st_hi(z.x), st_hi(z.y));
              const uint32_t o = st_off(128, (uint32_t)(f1 * ST_K1 + k), (uint32_t)(2 * n2));
              *reinterpret_cast<float2*>(G.a_hi + o) = h;
              *reinterpret_cast<float2*>(G.a_lo + o) = make_float2(st_hi(z.x - h.x), st_hi(z.y - h.y));
            }
          }
        }
      }
      st_fence_before();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      st_arrive(&S.a2_ready[g]);
      // ---- epilogue: D2 row (frame, k1) holds X[k1 + 25 k2] -> |.|^2 -> P[frame][bin]; bins 25 - k1 + ... by symmetry
      mbar_wait(&S.d2_full[g], parity);
      st_fence_after();
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t r[16];
        st_ld16(lane_addr + 32 + 16 * half, r);
        if (w < ST_FR * ST_K1 && f2 < cur.z) {
          float* __restrict__ row = power + (long long)(cur.w + f2) * ld_power;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const int k2 = 8 * half + kk;
            const float re = __uint_as_float(r[2 * kk]), im = __uint_as_float(r[2 * kk + 1]);
            const float p = re * re + im * im;
            const int k = k1 + 25 * k2;
            if (k <= 200) row[k] = p;
            else if (k1 > 0) row[N_FFT - k] = p;               // X[400 - k] = conj X[k]
          }
        }
      }
      st_fence_before();                                     // the next tile's MMAs overwrite D1 / D2 after our next arrive
    }
  }
  st_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    st_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(ST_TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------- host
size_t stft_tc_table_bytes() { return ST_TABLE_BYTES; }

// B1 [n = (k1, re/im)][k = n1], B2 [n = (k2, re/im)][k = (n2, re/im)] as hi / lo TF32 halves in the operand layout,
// then the window and the twiddles.  Angles in double, values rounded to fp32 first (the split is of the fp32 matrix).
void host_stft_tc_tables(unsigned char* out) {
  const double two_pi = 6.283185307179586476925286766559;
  std::vector<float> hann(N_FFT);
  host_hann(hann.data());
  auto put = [&](unsigned char* hi, unsigned char* lo, uint32_t n, uint32_t k, double val) {
    const float v = (float)val;
    uint32_t b;
    memcpy(&b, &v, 4);
    b = (b + 0x1000u) & 0xffffe000u;              // nearest TF32, like st_hi on the device
    float h;
    memcpy(&h, &b, 4);
    const float r = v - h;
    uint32_t bl;
    memcpy(&bl, &r, 4);
    bl = (bl + 0x1000u) & 0xffffe000u;
    float l;
    memcpy(&l, &bl, 4);
    const uint32_t o = st_off(32, n, k);
    memcpy(hi + o, &h, 4);
    memcpy(lo + o, &l, 4);
  };
  unsigned char* b1_hi = out; unsigned char* b1_lo = out + ST_B_BYTES;
  unsigned char* b2_hi = out + 2 * ST_B_BYTES; unsigned char* b2_lo = out + 3 * ST_B_BYTES;
  memset(out, 0, 4 * ST_B_BYTES);
  for (uint32_t n = 0; n < 32; ++n)
    for (uint32_t k = 0; k < 32; ++k) {
      const uint32_t kk1 = n >> 1, c = n & 1;
      double v = 0.0;
      if (kk1 < (uint32_t)ST_K1 && k < 25) {
        const double th = two_pi * (double)((k * kk1) % 25) / 25.0;
        v = c == 0 ? std::cos(th) : -std::sin(th);
      }
      put(b1_hi, b1_lo, n, k, v);
      const uint32_t kk2 = n >> 1, cp = n & 1, nn2 = k >> 1, cc = k & 1;
      const double t2 = two_pi * (double)((nn2 * kk2) % 16) / 16.0;
      const double cr = std::cos(t2), ci = -std::sin(t2);
      put(b2_hi, b2_lo, n, k, cp == 0 ? (cc == 0 ? cr : -ci) : (cc == 0 ? ci : cr));
    }
  float* win = reinterpret_cast<float*>(out + 4 * ST_B_BYTES);
  for (int i = 0; i < N_FFT; ++i) win[i] = hann[i];
  float* tw = win + N_FFT;
  for (int nn2 = 0; nn2 < 16; ++nn2)
    for (int kk1 = 0; kk1 < ST_K1; ++kk1) {
      const double th = two_pi * (double)(nn2 * kk1) / 400.0;
      tw[2 * (nn2 * ST_K1 + kk1)] = (float)std::cos(th);
      tw[2 * (nn2 * ST_K1 + kk1) + 1] = (float)(-std::sin(th));
    }
}

cudaError_t launch_stft_tc(const unsigned char* tables, const float* x16, const int64_t* off, const int32_t* len16,
                           int pad_frames, const int32_t* tiles, int n_tiles, float* power, int64_t ld_power, int sm_count,
                           cudaStream_t st, LaunchCtx* lc) {
  if (n_tiles <= 0) return cudaSuccess;
  const size_t smem = sizeof(StSmem) + 1024;
  cudaError_t e = cudaFuncSetAttribute(k_stft_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int grid = n_tiles < sm_count ? n_tiles : sm_count;
  lc->begin(KID_STFT_TC, st);
  k_stft_tc<<<grid, ST_THREADS, smem, st>>>(x16, off, len16, pad_frames, reinterpret_cast<const int4*>(tiles), n_tiles,
                                            tables, power, (long long)ld_power);
  lc->end(st);
  return cudaGetLastError();
}

}  // namespace rho
