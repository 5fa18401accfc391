// common.cuh — error handling, device buffers and small helpers shared by every translation unit
// of libabsb200.so.  Nothing here is visible through the C ABI (include/absb200.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <string>
#include <utility>
#include <vector>

#include "../../include/absb200.h"

namespace absb {

struct Error {
  int code;
  std::string msg;
};

[[noreturn]] inline void fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  throw Error{code, buf};
}

void set_last_error(const std::string& s);  // api.cu

#define ABSB_CUDA(expr)                                                                     \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess)                                                                  \
      ::absb::fail(_e == cudaErrorMemoryAllocation ? ABSB_ERR_OOM : ABSB_ERR_CUDA,          \
                   "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define ABSB_CHECK(cond, code, ...)              \
  do {                                           \
    if (!(cond)) ::absb::fail(code, __VA_ARGS__); \
  } while (0)

// Wraps the body of an extern "C" entry point: no exception crosses the boundary.
#define ABSB_API_BEGIN try {
#define ABSB_API_END                                       \
  return ABSB_OK;                                          \
  }                                                        \
  catch (const ::absb::Error& e) {                         \
    ::absb::set_last_error(e.msg);                         \
    return e.code;                                         \
  }                                                        \
  catch (const std::bad_alloc&) {                          \
    ::absb::set_last_error("host allocation failed");      \
    return ABSB_ERR_OOM;                                   \
  }                                                        \
  catch (const std::exception& e) {                        \
    ::absb::set_last_error(e.what());                      \
    return ABSB_ERR_INVALID;                               \
  }

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    ABSB_CUDA(cudaGetDevice(&prev));
    if (prev != dev) ABSB_CUDA(cudaSetDevice(dev));
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// Owning device buffer; grows geometrically on reserve(), never shrinks.
template <typename T>
struct DBuf {
  T* p = nullptr;
  size_t cap = 0;  // elements
  DBuf() = default;
  DBuf(const DBuf&) = delete;
  DBuf& operator=(const DBuf&) = delete;
  DBuf(DBuf&& o) noexcept : p(o.p), cap(o.cap) { o.p = nullptr; o.cap = 0; }
  DBuf& operator=(DBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; cap = o.cap; o.p = nullptr; o.cap = 0; }
    return *this;
  }
  ~DBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  // contents are NOT preserved
  void reserve(size_t n) {
    if (n <= cap) return;
    release();
    ABSB_CUDA(cudaMalloc(&p, n * sizeof(T)));
    cap = n;
  }
  void alloc_exact(size_t n) {
    release();
    if (n == 0) return;
    ABSB_CUDA(cudaMalloc(&p, n * sizeof(T)));
    cap = n;
  }
};

// Pinned host staging buffer.
template <typename T>
struct HBuf {
  T* p = nullptr;
  size_t cap = 0;
  HBuf() = default;
  HBuf(const HBuf&) = delete;
  HBuf& operator=(const HBuf&) = delete;
  ~HBuf() { if (p) cudaFreeHost(p); }
  void reserve(size_t n) {
    if (n <= cap) return;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    ABSB_CUDA(cudaMallocHost(&p, n * sizeof(T)));
    cap = n;
  }
};

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Ask for the largest shared-memory carveout.  The L1 / shared split of an SM can only change while the SM is
// idle, so kernels of two streams can share an SM only if whichever arrives first has already configured it
// for the sum of both: every kernel that may run while a co-resident scan CTA is resident asks for the maximum.
template <typename Kern>
inline void prefer_max_shared(Kern kern) {
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
}

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------
// A kernel launched with launch_pdl() may START (run its prologue: barrier init, TMEM allocation, tensor-map
// prefetch) while the previous kernel of the stream is still draining; it must call pdl_wait() before its first
// access to global memory (reads of the predecessor's output, and writes the predecessor may still read).
// pdl_trigger() at the top of every kernel lets ITS successor be scheduled as early as SM resources allow.
// On a 28-layer forward of ~230 short kernels this removes the launch gap between consecutive kernels.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

bool pdl_enabled();  // api.cu: ABSB_PDL=0 in the environment switches it off

template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  ABSB_CUDA(cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(std::forward<Args>(args))...));
}

struct DeviceProps {
  int sm_count = 0;
  int cc_major = 0, cc_minor = 0;
  size_t smem_optin = 0;
};
DeviceProps device_props(int device);  // api.cu (cached)

}  // namespace absb
