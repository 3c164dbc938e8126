// The library handle and the helpers shared by the translation units that implement the C ABI (api.cu, host_api.cu).
#pragma once
#include <atomic>
#include <map>
#include <mutex>
#include <vector>
#include <nvtx3/nvToolsExt.h>     // header-only NVTX 3: a no-op unless a tool (ncu --nvtx, nsys) injects itself
#include "kernels.h"

namespace rho {

extern thread_local char g_err[512];
int fail(int code, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* where);

// Every entry point runs on the handle's device, whatever the caller's current device is, and leaves the caller's
// current device as it found it (one process may drive several GPUs; torch keeps its own notion of "current").
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int device) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != device) {
      err = cudaSetDevice(device);
      switched = (err == cudaSuccess);
    }
  }
  ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};
// NVTX range named after the entry point, open for the duration of the call (tracing is a net-new
// deliverable).  Kernel launches get a nested range named after the kernel (LaunchCtx::begin / end).
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};
#define RHO_ON_DEVICE(h)                                            \
  rho::NvtxRange _nvtx(__func__);                                   \
  if (!(h)) return rho::fail(RHO_ERR_INVALID, "handle is NULL");    \
  rho::DeviceGuard _guard((h)->device);                             \
  if (_guard.err != cudaSuccess) return rho::cuda_fail(_guard.err, "cudaSetDevice")

// exchange.cu: this rank's gathered-record buffer and the peer mappings of everybody else's
struct Exchange {
  void* buf = nullptr;                   // [2 parities][world][n_per_rank] records | 32 flag words | error word | done counter
  size_t bytes = 0, rec_bytes = 0;
  int world = 0, rank = 0;
  int64_t n_per_rank = 0;
  uint64_t epoch = 0;                    // calls of rho_b200_validate since the exchange was connected
  bool connected = false;
  void* peer[MAX_RECORD_PEERS] = {};     // peer[q]: rank q's buffer in this process's address space
};

struct HostCtx;   // host_api.cu: streams, device arena and pinned staging of one in-flight host call

}  // namespace rho

struct rho_handle {
  int device;
  int sm_count;            // of `device` (grids of the persistent kernels)
  rho::Tables tb;
  std::vector<void*> allocs;
  rho::LaunchCtx lc;
  std::mutex mu;
  // HOST entry points: one context per concurrent caller, recycled (a handle may be shared by concurrent sessions,
  // ui/state.py:85-87 of the reference: callers do not serialise on each other)
  std::vector<rho::HostCtx*> host_ctx_free;
  // host entry points: fill threads of the next call (0 = default); raised when a call's fill fell behind its copies
  std::atomic<int> host_fill_threads{0};
  // tap tables of rho_b200_resample, one per reduced ratio seen so far: key = (orig << 32) | new
  std::map<uint64_t, float*> resample_taps;
  // rho_b200_pitch_shift: FFT / window / phase-advance tables (built on first use) and windowed tap tables per ratio
  rho::PitchTables pitch_tb{};
  bool pitch_tb_ok = false;
  rho::MfccTables mfcc_tb{};
  bool mfcc_tb_ok = false;
  struct WinTaps { float* taps; int* ilo; };
  std::map<uint64_t, WinTaps> windowed_taps;
  unsigned char* stft_tc_tables = nullptr;   // rho_b200_stft_power_tc: DFT matrices (hi / lo), window, twiddles
  // multi-GPU: records of rho_b200_validate are also stored into every rank's gathered buffer (rho_b200_exchange_*)
  rho::Exchange xch;
};

namespace rho {
void host_ctx_destroy(HostCtx* c);   // host_api.cu
}
