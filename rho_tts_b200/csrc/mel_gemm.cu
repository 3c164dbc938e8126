// Mel-filterbank projection as a tensor-core GEMM (tcgen05 / TMEM / TMA), in isolation:
//   mel[m][t] = sum_k F[m][k] * P[t][k]          m < n_mels (80 | 128), k < 201, t < n_frames
// = `mel_filters.T @ magnitudes` of transformers feature_extraction_whisper.py:159 (the reference reaches it
// through src/rho_tts/validation/stt/stt_validator.py:78-107).  SURVEY.md 8(d) asks for the tensor-pipe
// utilisation of this contraction measured on its own; the product path (fused.cu / logmel.cu) keeps the
// projection as 391 immediate-weight FFMAs per frame because the bank is 97.5 % sparse.
//
// Accuracy: the log-mel contract (1e-4 on values that span 80 dB) needs fp32-class products, so the GEMM is
// the 3xTF32 split  F*P = Fhi*Phi + Flo*Phi + Fhi*Plo  with hi = the top 19 bits, lo = the remainder.
//
// Mapping (one persistent CTA per SM, 192 threads):
//   A = the filterbank, padded to 128 x 208, RESIDENT IN TMEM for the whole kernel (hi: columns 0..207,
//       lo: 208..415), written once with tcgen05.st;  M = 128 TMEM lanes = mel rows
//   B = a tile of 32 frames x 208 bins of P, K-major in shared memory exactly as it lies in HBM: seven TMA
//       boxes (32 bins x 32 frames, 128-byte swizzle) per tile, four stages (the loads of three tiles are in
//       flight while one is multiplied); bins >= 201 and frames >= n_frames are zero-filled by the TMA unit
//   D = 128 x 32 fp32 accumulator in TMEM, two buffers (columns 416..479) so the epilogue of tile i overlaps
//       the MMAs of tile i+1
//   warp 0   : TMA producer (one lane)
//   warps 1-4: split the landed tile into hi / lo in place (layout-agnostic, the swizzle does not matter),
//              then drain the previous tile's accumulator: tcgen05.ld -> 128-bit stores of mel[m][t0..t0+31]
//   warp 5   : MMA issuer (one lane): 26 K-steps x 3 tcgen05.mma.kind::tf32 (M128 N32 K8), tcgen05.commit
#include <cuda.h>
#include <cudaTypedefs.h>
#include "common.cuh"
#include "kernels.h"

namespace rho {

constexpr int MG_NT = 32;                               // frames per tile (UMMA N)
constexpr int MG_NKB = 7;                               // TMA boxes per tile (7 x 32 bins >= 201)
constexpr int MG_KSTEPS = 26;                           // UMMA K = 8 tf32 values; 26 x 8 = 208 >= 201
constexpr int MG_BOX_BYTES = MG_NT * 128;               // 4096: 32 rows of 128 bytes
constexpr int MG_TILE_BYTES = MG_NKB * MG_BOX_BYTES;    // 28672
constexpr int MG_STAGES = 4;                            // 4 x (hi + lo) = 224 KB: three tiles of loads in flight per SM
constexpr int MG_THREADS = 192;
constexpr int MG_WORKERS = 128;                         // warps 1-4
constexpr uint32_t MG_COL_WHI = 0, MG_COL_WLO = 208, MG_COL_D = 416;
constexpr uint32_t MG_TMEM_COLS = 512;
// instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6)=1, A=tf32 [7,10)=2, B=tf32 [10,13)=2,
// A,B K-major, N>>3 at [17,23), M>>4 at [24,29)
constexpr uint32_t MG_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(MG_NT >> 3) << 17) | ((128u >> 4) << 24);

struct MgSmem {
  unsigned char hi[MG_STAGES][MG_TILE_BYTES];           // 1024-byte aligned (28672 = 28 * 1024)
  unsigned char lo[MG_STAGES][MG_TILE_BYTES];
  uint64_t full[MG_STAGES];                             // TMA bytes landed
  uint64_t ready[MG_STAGES];                            // hi / lo written (128 arrivals)
  uint64_t empty[MG_STAGES];                            // the MMAs reading the stage retired
  uint64_t d_full[2];                                   // accumulator complete
  uint64_t d_empty[2];                                  // accumulator drained (128 arrivals)
  uint32_t tmem_base;
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               :: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
}
// shared-memory matrix descriptor, K-major, 128-byte swizzle: 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t mg_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// The descriptor is passed as two 32-bit halves: only the 14-bit start-address field of the low word changes
// between K-steps, so a fully unrolled issue loop costs one uniform add per tcgen05.mma.
__device__ __forceinline__ void mg_mma(uint32_t d_tmem, uint32_t a_tmem, uint32_t desc_lo, uint32_t desc_hi,
                                       uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 bd;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b64 bd, {%2, %3};\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bd, %4, p;\n\t}"
               :: "r"(d_tmem), "r"(a_tmem), "r"(desc_lo), "r"(desc_hi), "r"(MG_IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void mg_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }

__global__ void __launch_bounds__(MG_THREADS, 1)
k_mel_gemm(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ F, int n_mels, long long n_frames,
           float* __restrict__ mel, long long ld_mel, long long frames_per_item, long long item_stride, int n_tiles) {
  extern __shared__ unsigned char mg_raw[];
  MgSmem& S = *reinterpret_cast<MgSmem*>(mg_raw + ((1024u - (smem_u32(mg_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < MG_STAGES; ++s) { mbar_init(&S.full[s], 1); mbar_init(&S.ready[s], MG_WORKERS); mbar_init(&S.empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&S.d_full[b], 1); mbar_init(&S.d_empty[b], MG_WORKERS); }
    mbar_fence_init();
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&S.tmem_base)), "r"(MG_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = S.tmem_base;

  // ---- the filterbank goes to TMEM once: lane = mel row, column = bin, hi and lo halves
  if (warp >= 1 && warp <= 4) {
    const int q = warp & 3;                               // the TMEM lane quarter this warp may touch
    const int m = 32 * q + lane;
    const uint32_t row = tmem + ((uint32_t)(32 * q) << 16);
    for (int kc = 0; kc < MG_KSTEPS; ++kc) {
      uint32_t h[8], l[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = 8 * kc + j;
        const float w = (m < n_mels && k < N_BINS) ? F[m * N_BINS + k] : 0.f;
        const float wh = tf32_hi(w);
        h[j] = __float_as_uint(wh);
        l[j] = __float_as_uint(tf32_hi(w - wh));
      }
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                   :: "r"(row + MG_COL_WHI + 8 * kc), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]), "r"(h[4]), "r"(h[5]), "r"(h[6]), "r"(h[7]) : "memory");
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                   :: "r"(row + MG_COL_WLO + 8 * kc), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]), "r"(l[4]), "r"(l[5]), "r"(l[6]), "r"(l[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 0) {
    // ================================ TMA producer
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int s = it % MG_STAGES;
        const unsigned ph = (unsigned)(it / MG_STAGES) & 1u;
        mbar_wait(&S.empty[s], ph ^ 1u);                  // a fresh barrier passes: the stage starts empty
        mbar_arrive_expect_tx(&S.full[s], MG_TILE_BYTES);
        for (int kb = 0; kb < MG_NKB; ++kb)
          tma_load_2d(S.hi[s] + kb * MG_BOX_BYTES, &tmap, 32 * kb, tile * MG_NT, &S.full[s]);
      }
    }
  } else if (warp == 5) {
    // ================================ MMA issuer (the warp stays converged; one elected lane issues)
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int s = it % MG_STAGES, buf = it & 1;
      const unsigned ph = (unsigned)(it / MG_STAGES) & 1u, dph = (unsigned)(it >> 1) & 1u;
      mbar_wait(&S.ready[s], ph);
      mbar_wait(&S.d_empty[buf], dph ^ 1u);
      tc_fence_after();
      const uint32_t d = tmem + MG_COL_D + (uint32_t)(MG_NT * buf);
      const uint64_t d0 = mg_desc(smem_u32(S.hi[s])), d1 = mg_desc(smem_u32(S.lo[s]));
      const uint32_t hi_lo = (uint32_t)d0, lo_lo = (uint32_t)d1, dhi = (uint32_t)(d0 >> 32);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < MG_KSTEPS; ++ks) {
          const uint32_t off = (uint32_t)(((ks >> 2) * MG_BOX_BYTES + (ks & 3) * 32) >> 4);   // start-address field units
          mg_mma(d, tmem + MG_COL_WHI + 8 * ks, hi_lo + off, dhi, ks > 0);
          mg_mma(d, tmem + MG_COL_WLO + 8 * ks, hi_lo + off, dhi, 1);
          mg_mma(d, tmem + MG_COL_WHI + 8 * ks, lo_lo + off, dhi, 1);
        }
        mg_commit(&S.empty[s]);                           // shared memory of the stage may be refilled
        mg_commit(&S.d_full[buf]);                        // accumulator may be drained
      }
      __syncwarp();
    }
  } else {
    // ================================ split + epilogue (warps 1-4)
    const int t = threadIdx.x - 32;
    const int q = warp & 3;
    const int m = 32 * q + lane;
    const bool vec = (ld_mel % 4 == 0) && (frames_per_item % 4 == 0) && (item_stride % 4 == 0) && frames_per_item >= MG_NT &&
                     ((reinterpret_cast<uintptr_t>(mel) & 15u) == 0);
    auto epilogue = [&](int e) {
      const int buf = e & 1;
      const unsigned dph = (unsigned)(e >> 1) & 1u;
      const long long f0 = (long long)(blockIdx.x + (long long)e * gridDim.x) * MG_NT;
      mbar_wait(&S.d_full[buf], dph);
      tc_fence_after();
      uint32_t r[MG_NT];
      const uint32_t a = tmem + ((uint32_t)(32 * q) << 16) + MG_COL_D + (uint32_t)(MG_NT * buf);
#pragma unroll
      for (int j = 0; j < MG_NT / 16; ++j)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(r[16 * j + 0]), "=r"(r[16 * j + 1]), "=r"(r[16 * j + 2]), "=r"(r[16 * j + 3]),
                       "=r"(r[16 * j + 4]), "=r"(r[16 * j + 5]), "=r"(r[16 * j + 6]), "=r"(r[16 * j + 7]),
                       "=r"(r[16 * j + 8]), "=r"(r[16 * j + 9]), "=r"(r[16 * j + 10]), "=r"(r[16 * j + 11]),
                       "=r"(r[16 * j + 12]), "=r"(r[16 * j + 13]), "=r"(r[16 * j + 14]), "=r"(r[16 * j + 15])
                     : "r"(a + 16 * j) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      tc_fence_before();
      mbar_arrive(&S.d_empty[buf]);                       // the accumulator is in registers: MMAs may reuse it
      if (m < n_mels) {
        // frame f belongs to item f / frames_per_item (one clip's [n_mels][ld_mel] block), column f % frames_per_item
        // (one 64-bit division per tile; a tile of 32 frames crosses at most one item boundary when
        //  frames_per_item >= 32, which the launcher guarantees for the vector path)
        const long long item0 = f0 / frames_per_item, t0 = f0 - item0 * frames_per_item;
        float* rowm = mel + (long long)m * ld_mel + item0 * item_stride;
        const long long wrap = item_stride - frames_per_item;   // added to the column once the tile crosses into the next item
        if (vec && f0 + MG_NT <= n_frames) {
#pragma unroll
          for (int j = 0; j < MG_NT / 4; ++j) {
            long long t = t0 + 4 * j;
            if (t >= frames_per_item) t += wrap;
            stg_stream4(rowm + t, make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                              __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])));
          }
        } else {
#pragma unroll
          for (int j = 0; j < MG_NT; ++j) {
            if (f0 + j < n_frames) {
              long long t = t0 + j;                       // general case: any number of item boundaries
              long long o = 0;
              while (t >= frames_per_item) { t -= frames_per_item; o += item_stride; }
              rowm[o + t] = __uint_as_float(r[j]);
            }
          }
        }
      }
    };
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int s = it % MG_STAGES;
      const unsigned ph = (unsigned)(it / MG_STAGES) & 1u;
      mbar_wait(&S.full[s], ph);
      float4* ph4 = reinterpret_cast<float4*>(S.hi[s]);
      float4* pl4 = reinterpret_cast<float4*>(S.lo[s]);
#pragma unroll 3
      for (int i = t; i < MG_TILE_BYTES / 16; i += MG_WORKERS) {
        const float4 p = ph4[i];
        const float4 h = make_float4(tf32_hi(p.x), tf32_hi(p.y), tf32_hi(p.z), tf32_hi(p.w));
        ph4[i] = h;
        pl4[i] = make_float4(p.x - h.x, p.y - h.y, p.z - h.z, p.w - h.w);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> the tensor core's async reads
      mbar_arrive(&S.ready[s]);
      if (it > 0) epilogue(it - 1);
    }
    if (it > 0) epilogue(it - 1);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(MG_TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------- host
static PFN_cuTensorMapEncodeTiled_v12000 mg_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

cudaError_t launch_mel_gemm(const Tables& tb, const float* power, int64_t n_frames, int64_t ld_power, int n_mels,
                            float* mel, int64_t ld_mel, int64_t frames_per_item, int64_t item_stride, int sm_count,
                            cudaStream_t st, LaunchCtx* lc) {
  if (n_frames <= 0) return cudaSuccess;
  auto encode = mg_encode_fn();
  if (!encode) return cudaErrorNotSupported;
  CUtensorMap tmap;
  const cuuint64_t dims[2] = {(cuuint64_t)N_BINS, (cuuint64_t)n_frames};
  const cuuint64_t strides[1] = {(cuuint64_t)ld_power * sizeof(float)};
  const cuuint32_t box[2] = {32, (cuuint32_t)MG_NT};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(power), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
  const int64_t n_tiles = (n_frames + MG_NT - 1) / MG_NT;
  if (n_tiles > 0x7fffffff) return cudaErrorInvalidValue;
  const size_t smem = sizeof(MgSmem) + 1024;
  cudaError_t e = cudaFuncSetAttribute(k_mel_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int grid = (int)(n_tiles < sm_count ? n_tiles : sm_count);
  const float* F = tb.mel_dense[n_mels == 80 ? 0 : 1];
  lc->begin(KID_MEL_GEMM, st);
  k_mel_gemm<<<grid, MG_THREADS, smem, st>>>(tmap, F, n_mels, (long long)n_frames, mel, (long long)ld_mel,
                                             (long long)frames_per_item, (long long)item_stride, (int)n_tiles);
  lc->end(st);
  return cudaGetLastError();
}

}  // namespace rho
