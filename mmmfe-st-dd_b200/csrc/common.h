// Host-side helpers shared by the engine translation units: error type, CUDA checks, owning device buffers, CSR.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/mmmfe_b200.h"

namespace mmmfe {

struct Error {
  int code;
  std::string msg;
};

[[noreturn]] inline void fail(int code, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  throw Error{code, buf};
}

#define CUDA_CHECK(x)                                                                              \
  do {                                                                                             \
    cudaError_t e_ = (x);                                                                          \
    if (e_ != cudaSuccess) fail(MMMFE_E_CUDA, "%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), \
                                __FILE__, __LINE__);                                               \
  } while (0)

// keeps freed blocks in the device's default memory pool (see DevBuf::release)
inline void dev_pool_retain(int device) {
  cudaMemPool_t pool = nullptr;
  if (cudaDeviceGetDefaultMemPool(&pool, device) != cudaSuccess || !pool) { cudaGetLastError(); return; }
  unsigned long long threshold = ~0ull;
  if (cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold) != cudaSuccess) cudaGetLastError();
}

template <class T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DevBuf &operator=(DevBuf &&o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  // Device memory comes from the device's stream-ordered pool with an unlimited release threshold
  // (dev_pool_retain, called once per context): a release hands the block back to the pool instead of unmapping it.
  // cudaFree of the multi-GB factorisation workspaces cost 0.5-1.3 s per factor group on the B200 boxes (measured,
  // DESIGN.md section 12), and the next group / refinement cycle allocates the same sizes again.  Like cudaFree, a
  // release first waits for the device: the engine's streams are non-blocking, a block must not be handed out again
  // while one of them still uses it.
  void release() {
    if (p) {
      cudaDeviceSynchronize();
      if (cudaFreeAsync(p, cudaStreamLegacy) != cudaSuccess) cudaFree(p);
    }
    p = nullptr;
    n = 0;
  }
  void alloc(size_t count) {
    release();
    n = count;
    if (!count) return;
    if (cudaMallocAsync(reinterpret_cast<void **>(&p), count * sizeof(T), cudaStreamLegacy) != cudaSuccess) {
      cudaGetLastError();
      CUDA_CHECK(cudaMalloc(&p, count * sizeof(T)));
      return;
    }
    CUDA_CHECK(cudaStreamSynchronize(cudaStreamLegacy));  // usable on every stream from here on
  }
  void zero(cudaStream_t st = 0) {
    if (n) CUDA_CHECK(cudaMemsetAsync(p, 0, n * sizeof(T), st));
  }
  void upload(const T *h, size_t count) {
    alloc(count);
    // a pageable-memory copy may return while its staged DMA is still in flight on the legacy stream; the
    // engine's own streams are non-blocking, so wait here before a kernel on one of them can read the buffer
    if (count) {
      CUDA_CHECK(cudaMemcpy(p, h, count * sizeof(T), cudaMemcpyHostToDevice));
      CUDA_CHECK(cudaStreamSynchronize(cudaStreamLegacy));
    }
  }
  void upload(const std::vector<T> &v) { upload(v.data(), v.size()); }
};

struct HostCsr {
  int64_t n_rows = 0, n_cols = 0;
  std::vector<int64_t> rp;
  std::vector<int32_t> col;
  std::vector<double> val;
  void from(const mmmfe_csr &c) {
    n_rows = c.n_rows;
    n_cols = c.n_cols;
    rp.assign(c.row_ptr, c.row_ptr + c.n_rows + 1);
    col.assign(c.col, c.col + rp.back());
    val.assign(c.val, c.val + rp.back());
  }
};

struct DevCsr {
  int64_t n_rows = 0;
  DevBuf<int64_t> rp;
  DevBuf<int32_t> col;
  DevBuf<double> val;
  void upload(const HostCsr &h) {
    n_rows = h.n_rows;
    rp.upload(h.rp);
    col.upload(h.col);
    val.upload(h.val);
  }
};

}  // namespace mmmfe
