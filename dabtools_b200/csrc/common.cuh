// common.cuh -- shared host/device plumbing for libdabgpu (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/dabgpu.h"
#include "../../include/dabgpu_tables.h"

#define DABGPU_EXPORT extern "C" __attribute__((visibility("default")))

namespace dabgpu {

// ---- error state (per thread), surfaced through dabgpu_last_error[_string] ----------
// status codes are the DABGPU_* macros of include/dabgpu.h

void set_error(int code, const char *fmt, ...);
int last_error_code();

#define CUDA_TRY(expr)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      ::dabgpu::set_error(DABGPU_ERR_CUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, \
                          #expr, cudaGetErrorString(_e));                                   \
      return DABGPU_ERR_CUDA;                                                     \
    }                                                                                       \
  } while (0)

// every kernel launch of the library is counted (bench.py reports it as gpu_launches)
void note_launch();
uint64_t launch_count();
#define LAUNCH_CHECK()               \
  do {                               \
    ::dabgpu::note_launch();         \
    CUDA_TRY(cudaGetLastError());    \
  } while (0)

// ---- the stream every launch of the calling thread goes to ---------------------------
cudaStream_t current_stream();

// ---- grow-only device / pinned-host scratch buffers -----------------------------------
struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes);  // returns status; contents are NOT preserved on growth
  void release();
  template <typename T>
  T *as() const {
    return reinterpret_cast<T *>(p);
  }
};
struct PinBuf {
  void *p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes);
  void release();
  template <typename T>
  T *as() const {
    return reinterpret_cast<T *>(p);
  }
};

// one-time device checks + constant-table upload; returns status
int ensure_device_ready();

// Small control transfers between pinned host memory and the device done by a kernel (zero-copy
// over UVA) instead of the copy engines: a 100 KB descriptor upload issued while a 268 MB sample
// upload is in flight would otherwise queue behind it and stall the kernels that need it.
// Both pointers 4-byte aligned, bytes a multiple of 4.
int launch_ctl_copy(void *dst, const void *src, size_t bytes, cudaStream_t st);

// ---- small device helpers -------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
  return r;
}
__device__ __forceinline__ uint4 ldg_nc_v4(const void *p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
#endif

}  // namespace dabgpu
