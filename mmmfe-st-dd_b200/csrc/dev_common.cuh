// Device helpers shared by the kernel translation units (kernels.cu, wide.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "kernels.cuh"

namespace mmmfe {

__device__ __forceinline__ void dmma_m8n8k4(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return uint32_t(__cvta_generic_to_shared(p)); }

// Asynchronous 8-byte copy global -> shared (LDGSTS); pred == false writes a zero without touching global memory.
// Element-wise so that every operand layout and alignment the callers have works; the copies of one stage are
// committed as a group and waited for STAGES - 2 groups later (multi-stage software pipeline).
__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gsrc, bool pred) {
  const int n = pred ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

__device__ __forceinline__ void bulk_load_start(double *dst, const double *src, uint32_t bytes, uint64_t *bar) {
  // one thread: arm the barrier with the byte count, then issue the bulk copies (<= 64 KB each)
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
  uint32_t done = 0;
  while (done < bytes) {
    const uint32_t n = min(bytes - done, 65536u);
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
            smem_u32(reinterpret_cast<char *>(dst) + done)),
        "l"(reinterpret_cast<const char *>(src) + done), "r"(n), "r"(smem_u32(bar))
        : "memory");
    done += n;
  }
}

__device__ __forceinline__ void mbar_init(uint64_t *bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(bar)) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ int64_t subtree_dof(const SubtreeArgs &a, const InstanceDev &in, int32_t code) {
  const int type = code >> 30, dj = (code >> 15) & 0x7fff, di = code & 0x7fff;
  const int64_t canon = type == 0 ? int64_t(in.J0 + dj) * (a.nx + 1) + in.I0 + di
                                  : a.nxe + int64_t(in.J0 + dj) * a.nx + in.I0 + di;
  return a.dof_of_canon ? int64_t(a.dof_of_canon[canon]) : canon;
}

}  // namespace mmmfe
