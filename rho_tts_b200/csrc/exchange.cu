// Multi-GPU record exchange: the ONE exchange of the path is the gather of the 48-byte per-item records.
// One process per GPU.  Every rank owns a gathered-record buffer (two parities x world x n_per_rank records + one flag
// word per source rank); the buffers are exported with CUDA IPC and mapped into every peer process, so that the kernel
// which assembles a record (k_logmel_norm, records_dev.cuh) stores it straight into all ranks' buffers over NVLink /
// NVSwitch peer memory and the last record of a call publishes an epoch flag.  No collective call, no extra launch on
// the producing side; a consumer enqueues k_wait_flags (one warp, bounded spin) where it needs the gathered block.
#include <cstring>
#include "handle.h"

using namespace rho;

namespace rho {

// one lane per source rank: spin (bounded) until its flag reaches `epoch`
__global__ void k_wait_flags(const unsigned* __restrict__ flags, int world, unsigned epoch, unsigned* __restrict__ err,
                             long long timeout_cycles) {
  const int q = threadIdx.x;
  if (q >= world) return;
  const long long t0 = clock64();
  for (;;) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + q) : "memory");
    if ((int)(v - epoch) >= 0) return;
    if (clock64() - t0 > timeout_cycles) { atomicExch(err, 1u + (unsigned)q); return; }   // never hang the GPU
    __nanosleep(200);
  }
}

}  // namespace rho

extern "C" {

int rho_b200_exchange_create(rho_handle* h, int world, int rank, int64_t n_per_rank, void* ipc_handle_out) {
  RHO_ON_DEVICE(h);
  if (world < 1 || world > MAX_RECORD_PEERS) return fail(RHO_ERR_INVALID, "world must be in [1, %d]", MAX_RECORD_PEERS);
  if (rank < 0 || rank >= world || n_per_rank <= 0 || !ipc_handle_out) return fail(RHO_ERR_INVALID, "bad exchange arguments");
  std::lock_guard<std::mutex> lock(h->mu);
  if (h->xch.buf) return fail(RHO_ERR_INVALID, "exchange already created on this handle");
  Exchange& X = h->xch;
  X.world = world; X.rank = rank; X.n_per_rank = n_per_rank; X.epoch = 0;
  X.rec_bytes = sizeof(rho_record) * 2 * (size_t)world * (size_t)n_per_rank;
  X.bytes = align_up(X.rec_bytes, 256) + 256;
  cudaError_t e = cudaMalloc(&X.buf, X.bytes);
  if (e != cudaSuccess) { X.buf = nullptr; return cuda_fail(e, "cudaMalloc(exchange)"); }
  if ((e = cudaMemset(X.buf, 0, X.bytes)) != cudaSuccess || (e = cudaDeviceSynchronize()) != cudaSuccess)
    return cuda_fail(e, "exchange memset");
  cudaIpcMemHandle_t mh;
  if ((e = cudaIpcGetMemHandle(&mh, X.buf)) != cudaSuccess) return cuda_fail(e, "cudaIpcGetMemHandle");
  static_assert(sizeof(mh) == 64, "CUDA IPC handles are 64 bytes");
  memcpy(ipc_handle_out, &mh, 64);
  return RHO_OK;
}

int rho_b200_exchange_connect(rho_handle* h, const void* all_handles) {
  RHO_ON_DEVICE(h);
  if (!all_handles) return fail(RHO_ERR_INVALID, "handles is NULL");
  std::lock_guard<std::mutex> lock(h->mu);
  Exchange& X = h->xch;
  if (!X.buf) return fail(RHO_ERR_INVALID, "exchange not created");
  for (int q = 0; q < X.world; ++q) {
    if (q == X.rank) { X.peer[q] = X.buf; continue; }
    cudaIpcMemHandle_t mh;
    memcpy(&mh, (const char*)all_handles + 64 * q, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      for (int r = 0; r < q; ++r) if (r != X.rank && X.peer[r]) { cudaIpcCloseMemHandle(X.peer[r]); X.peer[r] = nullptr; }
      return cuda_fail(e, "cudaIpcOpenMemHandle (peer record buffer)");
    }
    X.peer[q] = p;
  }
  X.connected = true;
  return RHO_OK;
}

int rho_b200_exchange_wait(rho_handle* h, int64_t epoch, void* stream) {
  RHO_ON_DEVICE(h);
  Exchange& X = h->xch;
  if (!X.connected) return fail(RHO_ERR_INVALID, "exchange not connected");
  if (epoch <= 0) return RHO_OK;
  unsigned* flags = (unsigned*)((char*)X.buf + align_up(X.rec_bytes, 256));
  h->lc.begin(KID_XCH_WAIT, (cudaStream_t)stream);
  k_wait_flags<<<1, 32, 0, (cudaStream_t)stream>>>(flags, X.world, (unsigned)epoch, flags + 32, 4000000000LL);   // ~2 s
  h->lc.end((cudaStream_t)stream);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? RHO_OK : cuda_fail(e, "exchange wait");
}

int64_t rho_b200_exchange_epoch(rho_handle* h) { return h ? (int64_t)h->xch.epoch : (int64_t)-1; }

int rho_b200_exchange_read(rho_handle* h, int64_t epoch, void* dst, int* timed_out, void* stream) {
  RHO_ON_DEVICE(h);
  Exchange& X = h->xch;
  if (!X.buf || !dst) return fail(RHO_ERR_INVALID, "exchange not created or dst is NULL");
  const size_t block = sizeof(rho_record) * (size_t)X.world * (size_t)X.n_per_rank;
  cudaError_t e = cudaMemcpyAsync(dst, (char*)X.buf + block * (size_t)(epoch & 1), block, cudaMemcpyDeviceToDevice,
                                  (cudaStream_t)stream);
  if (e != cudaSuccess)
    return fail(RHO_ERR_CUDA, "exchange read (records, %zu bytes from %p to %p): %s", block,
                (void*)((char*)X.buf + block * (size_t)(epoch & 1)), dst, cudaGetErrorString(e));
  if (timed_out) {
    unsigned v = 0;
    if ((e = cudaStreamSynchronize((cudaStream_t)stream)) != cudaSuccess) return cuda_fail(e, "exchange read (sync)");
    e = cudaMemcpy(&v, (char*)X.buf + align_up(X.rec_bytes, 256) + 128, 4, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return cuda_fail(e, "exchange read (time-out word)");
    *timed_out = (int)v;
  }
  return RHO_OK;
}

int rho_b200_exchange_destroy(rho_handle* h) {
  RHO_ON_DEVICE(h);
  std::lock_guard<std::mutex> lock(h->mu);
  Exchange& X = h->xch;
  cudaDeviceSynchronize();
  for (int q = 0; q < X.world; ++q)
    if (q != X.rank && X.peer[q]) cudaIpcCloseMemHandle(X.peer[q]);
  if (X.buf) cudaFree(X.buf);
  X = Exchange{};
  return RHO_OK;
}

}  // extern "C"
