// Persistent time-sweep kernel (see sweep.cuh) and the one-time repack of the factor into sweep layout.
#include "sweep.cuh"

#include "kernels.cuh"

namespace mmmfe {

// =================================================================================================================
// repack: old front storage -> contiguous column blocks / row-major subtree blobs
// =================================================================================================================
__global__ void __launch_bounds__(256) k_repack(const RepackDesc *__restrict__ desc, const double *__restrict__ R,
                                                double *__restrict__ Rs) {
  const RepackDesc d = desc[blockIdx.x];
  const int64_t n = int64_t(d.rlen) * d.cw;
  const double *src = R + d.src;
  double *dst = Rs + d.dst;
  if (d.mode == 0) {
    for (int64_t e = threadIdx.x; e < n; e += blockDim.x) {
      const int r = int(e / d.cw), c = int(e % d.cw);
      dst[e] = (d.col0 + c < d.ncols) ? src[int64_t(r) * d.ld + d.col0 + c] : 0.0;
    }
  } else {
    for (int64_t e = threadIdx.x; e < n; e += blockDim.x) {  // read along the rows of R (coalesced), write strided
      const int c = int(e / d.rlen), r = int(e % d.rlen);
      dst[int64_t(r) * d.cw + c] = (d.col0 + c < d.ncols) ? src[int64_t(d.col0 + c) * d.ld + r] : 0.0;
    }
  }
}

void launch_repack(const RepackDesc *d_desc, int32_t n_desc, const double *R_old, double *Rs, cudaStream_t st) {
  for (int32_t off = 0; off < n_desc; off += 1 << 20) {
    const int32_t cnt = n_desc - off < (1 << 20) ? n_desc - off : (1 << 20);
    k_repack<<<cnt, 256, 0, st>>>(d_desc + off, R_old, Rs);
    launch_counter_add(1);
  }
}

__global__ void __launch_bounds__(256) k_repack_sub(const RepackSub *__restrict__ desc, const int32_t *__restrict__ perm,
                                                    const double *__restrict__ R, double *__restrict__ Rs) {
  const RepackSub d = desc[blockIdx.x];
  for (int k = threadIdx.x; k < d.rf_size; k += blockDim.x) Rs[d.new_rf + k] = R[d.old_rf + k];
  const int32_t *pm = perm + d.perm_off;
  for (int k = threadIdx.x; k < d.rb_size; k += blockDim.x) Rs[d.new_rb + k] = pm[k] < 0 ? 0.0 : R[d.old_rb + pm[k]];
}

void launch_repack_sub(const RepackSub *d_desc, int32_t n_desc, const int32_t *perm, const double *R_old, double *Rs,
                       cudaStream_t st) {
  for (int32_t off = 0; off < n_desc; off += 1 << 20) {
    const int32_t cnt = n_desc - off < (1 << 20) ? n_desc - off : (1 << 20);
    k_repack_sub<<<cnt, 256, 0, st>>>(d_desc + off, perm, R_old, Rs);
    launch_counter_add(1);
  }
}

// bitwise comparison of factor ranges (shared subtree blobs must really be identical)
__global__ void __launch_bounds__(256) k_blob_compare(const BlobPair *__restrict__ pairs, const double *__restrict__ R,
                                                      int *__restrict__ mismatch) {
  const BlobPair p = pairs[blockIdx.x];
  const unsigned long long *a = reinterpret_cast<const unsigned long long *>(R + p.a);
  const unsigned long long *b = reinterpret_cast<const unsigned long long *>(R + p.b);
  bool bad = false;
  for (int64_t k = threadIdx.x; k < p.n; k += blockDim.x) bad |= a[k] != b[k];
  if (bad) atomicAdd(mismatch, 1);
}

void launch_blob_compare(const BlobPair *d_pairs, int32_t n_pairs, const double *R, int *d_mismatch, cudaStream_t st) {
  for (int32_t off = 0; off < n_pairs; off += 1 << 20) {
    const int32_t cnt = n_pairs - off < (1 << 20) ? n_pairs - off : (1 << 20);
    k_blob_compare<<<cnt, 256, 0, st>>>(d_pairs + off, R, d_mismatch);
    launch_counter_add(1);
  }
}

// =================================================================================================================
// device helpers
// =================================================================================================================
namespace {

constexpr int NB = SWEEP_NBAR;

__device__ __forceinline__ uint32_t s_u32(const void *p) { return uint32_t(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void bar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(s_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void bar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(s_u32(bar)) : "memory");
}

__device__ __forceinline__ bool bar_try(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(s_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

// Every spin in the kernel is bounded: after ~4 s (or when another thread gave up) the waiter raises the abort flag
// (pinned host memory) and carries on with whatever data it has; the host reports the failure after the launch.
constexpr long long SWEEP_SPIN_LIMIT = 8000000000LL;

// (The flag lives in device memory: a poll is an L2 hit.  It is still only looked at after a long wait.)
__device__ __forceinline__ bool spin_expired(long long t0, volatile int *abort_flag) {
  const long long waited = clock64() - t0;
  if (waited < (1LL << 24)) return false;  // < ~9 ms: nothing can be wrong yet
  if (*abort_flag) return true;
  if (waited > SWEEP_SPIN_LIMIT) {
    *abort_flag = 1;
    return true;
  }
  return false;
}

__device__ __forceinline__ void bar_wait(uint64_t *bar, uint32_t parity, volatile int *abort_flag) {
  if (bar_try(bar, parity)) return;
  const long long t0 = clock64();
  for (int n = 0; !bar_try(bar, parity); ++n)
    if ((n & 1023) == 1023 && spin_expired(t0, abort_flag)) return;
}

__device__ __forceinline__ long long global_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
  return t;
}

__device__ __forceinline__ uint32_t ld_acquire(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void red_release(uint32_t *p) {
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;\n" ::"l"(p) : "memory");
}

__device__ __forceinline__ void red_relaxed(uint32_t *p) {
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;\n" ::"l"(p) : "memory");
}

__device__ __forceinline__ void wait_counter(const uint32_t *p, uint32_t target, volatile int *abort_flag) {
  if (ld_acquire(p) >= target) return;
  const long long t0 = clock64();
  for (int n = 0; ld_acquire(p) < target; ++n) {
    __nanosleep(20);
    if ((n & 255) == 255 && spin_expired(t0, abort_flag)) return;
  }
}

// Completion of entries inside the CTA: monotonic shared-memory counters (one per slot, bumped once per compute
// warp).  The retire and dependency lanes may lag arbitrarily far behind the compute warps, so a phase-parity
// barrier could be missed; a counter cannot.
__device__ __forceinline__ void done_signal(uint32_t *cnt) {
  asm volatile("red.release.cta.shared::cta.add.u32 [%0], 1;\n" ::"r"(s_u32(cnt)) : "memory");
}

__device__ __forceinline__ void done_wait(const uint32_t *cnt, uint32_t target, volatile int *abort_flag) {
  uint32_t v;
  const uint32_t addr = s_u32(cnt);
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];\n" : "=r"(v) : "r"(addr) : "memory");
  if (v >= target) return;
  const long long t0 = clock64();
  for (int n = 0;; ++n) {
    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];\n" : "=r"(v) : "r"(addr) : "memory");
    if (v >= target) return;
    __nanosleep(20);
    if ((n & 255) == 255 && spin_expired(t0, abort_flag)) return;
  }
}

// compute warps only
template <int CT>
__device__ __forceinline__ void csync_t() {
  asm volatile("bar.sync 1, %0;\n" ::"n"(CT) : "memory");
}

// generic-proxy global writes that a later TMA (async proxy) read of this CTA must see
__device__ __forceinline__ void fence_async_global() { asm volatile("fence.proxy.async.global;\n" ::: "memory"); }

__device__ __forceinline__ int t16(const uint16_t *p, int i) {
  const int v = p[i];
  return v == 0xFFFF ? -1 : v;
}

template <int M>
__device__ __forceinline__ void ld_vec(const double *p, double (&v)[M]) {  // L2 load of the M right-hand sides of one entry
  if constexpr (M == 1) {
    v[0] = __ldcg(p);
  } else {
#pragma unroll
    for (int j = 0; j < M; j += 2) {
      const double2 t = __ldcg(reinterpret_cast<const double2 *>(p) + j / 2);
      v[j] = t.x;
      v[j + 1] = t.y;
    }
  }
}

template <int M>
__device__ __forceinline__ void lds_vec(const double *p, double (&v)[M]) {  // shared-memory load of one vector entry
  if constexpr (M == 1) {
    v[0] = *p;
  } else {
#pragma unroll
    for (int j = 0; j < M; j += 2) {
      const double2 t = *(reinterpret_cast<const double2 *>(p) + j / 2);
      v[j] = t.x;
      v[j + 1] = t.y;
    }
  }
}

struct Tmpl {  // views into the shared-memory copy of a packed template (layout: sweep_plan.cpp pack_template)
  int n_levels, n_int, n_rootb, zl_size, w, h;
  const uint16_t *var_code, *cell_edge, *lvl_out, *in_idx;
  const uint4 *fdesc;   // 8 uint16 per forward output
  const uint32_t *sc;   // 2 uint16 per separator variable
  const int4 *lvlb;     // per backward level
  const uint2 *bdesc;   // 4 uint16 per variable
  __device__ __forceinline__ explicit Tmpl(const uint16_t *tb) {
    const int *h32 = reinterpret_cast<const int *>(tb);
    n_levels = h32[0]; n_int = h32[1]; n_rootb = h32[2]; zl_size = h32[3]; w = h32[4]; h = h32[5];
    var_code = tb + h32[9]; cell_edge = tb + h32[10]; lvl_out = tb + h32[11];
    fdesc = reinterpret_cast<const uint4 *>(tb + h32[12]);
    sc = reinterpret_cast<const uint32_t *>(tb + h32[13]);
    lvlb = reinterpret_cast<const int4 *>(tb + h32[14]);
    bdesc = reinterpret_cast<const uint2 *>(tb + h32[15]);
    in_idx = tb + h32[16];
  }
};

__device__ __forceinline__ int64_t sweep_dof(const SweepArgs &a, int I0, int J0, int code) {
  const int type = code >> 14, dj = (code >> 7) & 127, di = code & 127;
  const int64_t canon = type == 0 ? int64_t(J0 + dj) * (a.nx + 1) + I0 + di : a.nxe + int64_t(J0 + dj) * a.nx + I0 + di;
  return a.dof_of_canon ? int64_t(a.dof_of_canon[canon]) : canon;
}

}  // namespace

// =================================================================================================================
// the sweep kernel
// =================================================================================================================
#define SWEEP_PROF(slot)                                     \
  if (a.prof && tid == 0) {                                  \
    const long long t_now = clock64();                       \
    a.prof[cta * 24 + (slot)] += t_now - t_mark;             \
    t_mark = t_now;                                          \
  }

template <int M, int CT>
__global__ void __launch_bounds__(CT + 64, (CT >= 256 ? 1 : (CT >= 128 ? 3 : 6))) k_sweep(const SweepArgs a) {
  extern __shared__ __align__(128) double sm[];
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(sm);  // TMA data of the entry landed
  uint32_t *done_cnt = reinterpret_cast<uint32_t *>(full_bar + NB);  // compute warps that finished the slot's entries
  uint64_t *dep_bar = full_bar + 2 * NB;                  // descriptor staged, dependencies satisfied
  SweepEntry *edesc = reinterpret_cast<SweepEntry *>(sm + 3 * NB);
  uint8_t *gtab = reinterpret_cast<uint8_t *>(sm + 19 * NB);  // compute group of every entry of this CTA
  uint32_t *retired = reinterpret_cast<uint32_t *>(sm + 19 * NB + SWEEP_GTAB_DBL - 1);
  uint16_t *tb_all = reinterpret_cast<uint16_t *>(sm + 19 * NB + SWEEP_GTAB_DBL);
  double *ws_all = sm + 19 * NB + SWEEP_GTAB_DBL + a.ng * a.tmpl_dbl;
  double *ring = ws_all + a.ng * a.ws_dbl;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cta = blockIdx.x;
  const int e_first = a.cta_begin[cta];
  const int P = a.cta_begin[cta + 1] - e_first;
  if (P <= 0) return;
  const SweepEntry *ents = a.entries + e_first;
  const SweepIssue *iss = a.issue + e_first;
  const long long total = (long long)P * a.n_steps;

  if (tid == 0) {
    for (int k = 0; k < NB; ++k) {
      bar_init(full_bar + k, 1);
      done_cnt[k] = 0;
      bar_init(dep_bar + k, 1);
    }
    *retired = 0;
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (a.prof && tid == 0) {
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;\n" : "=r"(smid));
    a.prof[cta * 24 + 15] = smid;
  }
  for (int t = tid; t < P; t += CT + 64) gtab[t] = a.grp[e_first + t];
  __syncthreads();

  if (warp == CT / 32) {
    // =============================================================================================================
    // RETIRE warp: 32 independent lanes, lane l retires the entries G = l (mod 32): completion -> release counter ->
    // TMA copies of the entries this completion frees ring space for.  Which entries a completion releases is
    // static (issue_upto of the previous entry .. issue_upto of this one), so the lanes share no state.
    // =============================================================================================================
    uint64_t policy = 0;
    if (a.evict_first) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(policy));
    auto issue_one = [&](long long g, const SweepIssue &s) {
      uint64_t *fb = full_bar + (g % NB);
      const uint32_t bytes = uint32_t(s.n[0] + s.n[1] + s.n[2] + s.n[3]) * 8u;
      if (bytes == 0) {
        bar_arrive(fb);  // keeps the phase count
        return;
      }
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(s_u32(fb)), "r"(bytes) : "memory");
      double *dst = ring + s.ring_off;
#pragma unroll
      for (int t = 0; t < SWEEP_SEGS; ++t) {
        if (s.n[t] <= 0) continue;
        const double *base = s.kind[t] == SRC_FACTOR ? a.Rs
                             : s.kind[t] == SRC_INDEX ? reinterpret_cast<const double *>(a.index)
                             : s.kind[t] == SRC_XI    ? a.xi
                                                      : a.ptil;
        const char *src = reinterpret_cast<const char *>(base + s.off[t]);
        const bool hint = a.evict_first && s.kind[t] == SRC_FACTOR;
        uint32_t left = uint32_t(s.n[t]) * 8u, done = 0;
        while (left > 0) {
          const uint32_t n = left < 32768u ? left : 32768u;
          if (hint)
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n" ::
                    "r"(s_u32(reinterpret_cast<char *>(dst) + done)),
                "l"(src + done), "r"(n), "r"(s_u32(fb)), "l"(policy)
                : "memory");
          else
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                             s_u32(reinterpret_cast<char *>(dst) + done)),
                         "l"(src + done), "r"(n), "r"(s_u32(fb))
                         : "memory");
          done += n;
          left -= n;
        }
        dst += s.n[t];
      }
    };
    if (lane == 0) {
      long long target = a.cta_prologue[cta];
      if (target > total) target = total;
      for (long long g = 0; g < target; ++g) issue_one(g, iss[g % P]);
    }
    // One converged polling loop: lane l owns slot l and retires the entries G = l (mod NB) in turn.
    long long G = lane < NB ? lane : total;
    int sig = -1, n_sig = 0, sig_off = 0;
    long long first = 0, upto = 0;
    auto load_info = [&](long long g) {
      const int step = int(g / P), i = int(g - (long long)step * P);
      sig = ents[i].sig;
      n_sig = ents[i].n_sig;
      sig_off = ents[i].sig_off;
      first = (i == 0) ? (step == 0 ? (long long)a.cta_prologue[cta] : (long long)(step - 1) * P + ents[P - 1].issue_upto)
                       : (long long)step * P + ents[i - 1].issue_upto;
      upto = (long long)step * P + ents[i].issue_upto;
      if (upto > total) upto = total;
    };
    bool have = G < total;
    if (have) load_info(G);
    int state = 0;  // 0: waiting for the compute warps, 1: completion published, waiting for my turn to free the ring
    long long t0 = clock64();
    for (int spins = 1; __any_sync(0xffffffffu, have); ++spins) {
      bool progress = false;
      if (have && state == 0) {
        uint32_t v;
        asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];\n" : "=r"(v) : "r"(s_u32(done_cnt + (lane & (NB - 1)))) : "memory");
        if (v >= uint32_t(G / NB + 1) * uint32_t(CT / 32 / a.ng)) {  // (the last compute warp published the completion)
          state = 1;
          progress = true;
        }
      }
      // ring space is freed in entry order (the compute groups may finish out of order; the static ring schedule
      // assumes that everything before a finished entry is finished too)
      if (have && state == 1) {
        uint32_t r;
        asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];\n" : "=r"(r) : "r"(s_u32(retired)) : "memory");
        if (a.ng == 1 || r == uint32_t(G)) {  // one group finishes in order by itself
          for (long long g = first; g < upto; ++g) issue_one(g, iss[g % P]);
          asm volatile("st.release.cta.shared::cta.u32 [%0], %1;\n" ::"r"(s_u32(retired)), "r"(uint32_t(G + 1)) : "memory");
          G += NB;
          have = G < total;
          state = 0;
          if (have) load_info(G);
          t0 = clock64();
          progress = true;
        }
      }
      if (!__any_sync(0xffffffffu, progress)) __nanosleep(a.spin_ns);
      if ((spins & 1023) == 0 && have && spin_expired(t0, a.abort_flag)) have = false;
    }
    return;
  }

  if (warp == CT / 32 + 1) {
    // =============================================================================================================
    // DEPENDENCY warp: one converged polling loop, lane l owns slot l: stages the descriptor of its next entry as
    // soon as the slot is free (up to NB entries ahead of the compute warps), then waits for the entry's producers
    // on other CTAs, then opens the slot for the compute warps
    // =============================================================================================================
    long long G = lane < NB ? lane : total;
    bool have = G < total;
    int state = 0;
    int4 d[8];
    int n_dep = 0;
    uint32_t step1 = 0;
    auto load_desc = [&](long long g) {
      const int i = int(g % P);
#pragma unroll
      for (int t = 0; t < 8; ++t) d[t] = reinterpret_cast<const int4 *>(&ents[i])[t];
    };
    if (have) load_desc(G);
    long long t0 = clock64(), t_wait = 0;
    for (int spins = 1; __any_sync(0xffffffffu, have); ++spins) {
      bool progress = false;
      if (have && state == 0) {
        bool free_slot = G < NB;
        if (!free_slot) {
          uint32_t v;
          asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];\n" : "=r"(v) : "r"(s_u32(done_cnt + (lane & (NB - 1)))) : "memory");
          free_slot = v >= uint32_t(G / NB) * uint32_t(CT / 32 / a.ng);
        }
        if (free_slot) {
#pragma unroll
          for (int t = 0; t < 8; ++t) reinterpret_cast<int4 *>(&edesc[lane & (NB - 1)])[t] = d[t];
          step1 = uint32_t(G / P) + 1u;
          n_dep = d[1].w;
          state = 1;
          progress = true;
          if (a.prof) t_wait = clock64();
        }
      }
      // only the few entries the compute warps need next poll the global counters: the others could not be used yet
      // anyway and would only load the L2 slices that hold hot counters
      const long long g_min = __reduce_min_sync(0xffffffffu, have ? (unsigned)(G & 0x7fffffff) : 0x7fffffffu);
      if (have && state == 1 && (n_dep == 0 || (long long)(G & 0x7fffffff) < g_min + a.poll_window)) {
        bool ok = true;
        const int deps[SWEEP_MAX_DEPS] = {d[2].x, d[2].y, d[2].z, d[2].w, d[3].x, d[3].y, d[3].z, d[3].w};
#pragma unroll
        for (int t = 0; t < SWEEP_MAX_DEPS; ++t)
          if (ok && t < n_dep &&
              ld_acquire(a.cnt + (deps[t] & ((1 << SWEEP_DEP_SHIFT) - 1))) < step1 * uint32_t(deps[t] >> SWEEP_DEP_SHIFT))
            ok = false;
        if (ok) {
          if (a.prof && n_dep > 0)
            atomicAdd(reinterpret_cast<unsigned long long *>(a.prof + cta * 24 + 4), (unsigned long long)(clock64() - t_wait));
          bar_arrive(dep_bar + (lane & (NB - 1)));  // release: the descriptor stores above are visible to the waiters
          G += NB;
          have = G < total;
          state = 0;
          if (have) load_desc(G);
          t0 = clock64();
          progress = true;
        }
      }
      if (!__any_sync(0xffffffffu, progress)) __nanosleep(a.spin_ns);
      if ((spins & 255) == 0 && have && spin_expired(t0, a.abort_flag)) have = false;
    }
    return;
  }

  // ===============================================================================================================
  // COMPUTE warps
  // ===============================================================================================================
  // The compute warps form a.ng independent groups of GT threads; group g processes the entries tagged g (static,
  // all chunks of a block carry the same tag) and has its own workspace, reduction buffers and template cache.
  const int GT = CT / a.ng, cgrp = tid / GT, gt = tid - cgrp * GT;
  double *ws = ws_all + cgrp * a.ws_dbl;
  uint16_t *tb = tb_all + cgrp * a.tmpl_dbl * 4;
  auto gsync = [&] { asm volatile("bar.sync %0, %1;\n" ::"r"(cgrp + 1), "r"(GT) : "memory"); };
  int cur_tmpl = -1;

  long long G = 0;
  for (int step = 0; step < a.n_steps; ++step) {
    const int level = a.level0 + step;
    for (int i = 0; i < P; ++i, ++G) {
      if (gtab[i] != cgrp) continue;
      const int slot = int(G % NB);
      const uint32_t parity = uint32_t((G / NB) & 1);
      long long t_entry = 0, t_w = 0, g_entry = 0, g_w = 0;
      const bool tracing = a.trace && tid == 0 && step == a.trace_step;  // stamps only: no read-modify-write, no stall
      if (tracing) g_entry = global_ns();
      if (a.prof && tid == 0) t_entry = clock64();
      bar_wait(dep_bar + slot, parity, a.abort_flag);
      if (tracing) g_w = global_ns();
      if (a.prof && tid == 0) t_w = clock64();
      const SweepEntry &e = edesc[slot];
      const int kind = e.kind, flags = e.flags;
      const double *reg = ring + e.ring_off;
      // the first two completion counters of the entry, loaded now (an L2 round trip that would otherwise sit between
      // the end of the arithmetic and the publication of the completion)
      int32_t sig0 = -1, sig1 = -1;
      if (lane == 0) {
        if (e.n_sig > 0) sig0 = __ldg(a.index + e.sig_off);
        if (e.n_sig > 1) sig1 = __ldg(a.index + e.sig_off + 1);
      }
      bar_wait(full_bar + slot, parity, a.abort_flag);
      if (a.prof && tid == 0) {
        const long long t = clock64();
        a.prof[cta * 24 + 7] += t_w - t_entry;  // waiting for the dependency warp
        a.prof[cta * 24 + 16 + kind] += t_w - t_entry;
        a.prof[cta * 24 + 5] += t - t_w;        // waiting for TMA data
      }
      long long t_mark = (a.prof && tid == 0) ? clock64() : 0;

      if (kind == SW_TOP_FWD || kind == SW_TOP_BWD) {
        // ---------------------------------------------------------------------------------------------------------
        // wave: gather | micro-tiles (one warp each) | reduction of the row parts + output
        // ---------------------------------------------------------------------------------------------------------
        const bool fwd = kind == SW_TOP_FWD;
        const int n_rows = e.a0, n_mt = e.a1, n_cols = e.a2;
        const int32_t *ix = reinterpret_cast<const int32_t *>(reg);
        const double *tiles = reg + e.n[0];
        double *scratch = const_cast<double *>(reg) + e.n[0] + e.n[1];
        double *wsv = (flags & SWF_ROWS_WS) ? ws : scratch;         // [n_rows][M]
        double *part = (flags & SWF_ROWS_WS) ? scratch : scratch + n_rows * M;  // [partial slots][M]
        const int4 *colA = reinterpret_cast<const int4 *>(ix + e.a6);
        // A wave that reuses the gathered rows of the previous wave (same fronts, same CTA) needs no barrier here:
        // the previous wave's first barrier already ordered this CTA's earlier entries before every thread.
        const bool reuse = (flags & SWF_REUSE) != 0;
        if (!reuse) gsync();  // global writes of this CTA's earlier entries are visible to all its threads
        SWEEP_PROF(20)
        // pass-through of the children's contributions to this wave's boundary columns: loaded into registers now,
        // added in the output phase (the L2 round trip overlaps the micro-tiles)
        constexpr int PV = 2;
        double pv[PV][M];
        if (fwd) {
          const int32_t *colB = ix + e.o3;
#pragma unroll
          for (int u = 0; u < PV; ++u) {
            const int o = gt + u * GT;
#pragma unroll
            for (int j = 0; j < M; ++j) pv[u][j] = 0.0;
            if (o < n_cols && (colA[o].y & 0xFFFF) != 0) {  // (columns of direct micro-tiles are loaded by their warp)
              const int c0 = colA[o].w, c1 = colB[o];
              double p0[M], p1[M];
              if (c0 >= 0) ld_vec<M>(a.z + int64_t(c0) * M, p0);
              if (c1 >= 0) ld_vec<M>(a.z + int64_t(c1) * M, p1);
#pragma unroll
              for (int j = 0; j < M; ++j) pv[u][j] = (c0 >= 0 ? p0[j] : 0.0) + (c1 >= 0 ? p1[j] : 0.0);
            }
          }
        }
        if (fwd) {
          // rows of v = rhs of the separator dofs (assemble_rhs_star fused: -K_up p_old) + children's contributions
          const int4 *rows = reinterpret_cast<const int4 *>(ix + e.a4);
          const int32_t *zsd = ix + e.o0;
          constexpr int RB = 2;  // rows per thread in flight
          for (int r0 = gt; r0 < n_rows && !(flags & SWF_REUSE); r0 += RB * GT) {
            int4 gi[RB];
            double pa[RB][M], pb[RB][M], za[RB][M], zb[RB][M];
#pragma unroll
            for (int u = 0; u < RB; ++u) {
              const int r = r0 + u * GT;
              gi[u] = r < n_rows ? rows[r] : make_int4(0, 0, -1, -1);
            }
#pragma unroll
            for (int u = 0; u < RB; ++u) {
              ld_vec<M>(a.ptil + int64_t(gi[u].x & 0x3fffffff) * M, pa[u]);
              ld_vec<M>(a.ptil + int64_t(gi[u].y) * M, pb[u]);
              if (gi[u].z >= 0) ld_vec<M>(a.z + int64_t(gi[u].z) * M, za[u]);
              if (gi[u].w >= 0) ld_vec<M>(a.z + int64_t(gi[u].w) * M, zb[u]);
            }
#pragma unroll
            for (int u = 0; u < RB; ++u) {
              const int r = r0 + u * GT;
              if (r < n_rows) {
                const bool ty = (gi[u].x >> 30) & 1;
                const double k0 = ty ? a.ky0 : a.kx0, k1 = ty ? a.ky1 : a.kx1;
                const int zd = zsd[r];
#pragma unroll
                for (int j = 0; j < M; ++j) {
                  double val = -(k0 * pa[u][j] + k1 * pb[u][j]);
                  if (gi[u].z >= 0) val += za[u][j];
                  if (gi[u].w >= 0) val += zb[u][j];
                  wsv[r * M + j] = val;
                  if (zd >= 0) a.z[int64_t(zd) * M + j] = val;
                }
              }
            }
          }
        } else {
          // rows of w = [accumulated separator rhs ; solution of the boundary dofs]
          const int32_t *rows = ix + e.a4;
          constexpr int RB = 4;
          for (int r0 = gt; r0 < n_rows && !(flags & SWF_REUSE); r0 += RB * GT) {
            double v[RB][M];
#pragma unroll
            for (int u = 0; u < RB; ++u) {
              const int r = r0 + u * GT;
              if (r < n_rows) {
                const int32_t d = rows[r];
                const double *src = d < 0 ? a.z + int64_t(d & 0x7fffffff) * M : a.x + int64_t(d) * M;
                ld_vec<M>(src, v[u]);
              }
            }
#pragma unroll
            for (int u = 0; u < RB; ++u) {
              const int r = r0 + u * GT;
              if (r < n_rows)
#pragma unroll
                for (int j = 0; j < M; ++j) wsv[r * M + j] = v[u][j];
            }
          }
        }
        if (!reuse) gsync();
        SWEEP_PROF(14)
        {
          const SweepMt *mts = reinterpret_cast<const SweepMt *>(ix + e.a5);
          const int wig = gt >> 5, nw = GT >> 5;
          for (int m = wig; m < n_mt; m += nw) {
            const int4 md = *reinterpret_cast<const int4 *>(&mts[m]);  // tile_off, row0, nrows, ncols
            const int4 me = *(reinterpret_cast<const int4 *>(&mts[m]) + 1);  // part_off, col0, direct
            const int part_off = me.x;
            const bool direct = me.z != 0;
            const int nrows = md.z, ncols = md.w;
            const double *vp = wsv + md.y * M;
            // direct micro-tile: the lanes that end up with the column sums (row group 0) write the outputs; the
            // pass-through of the children's contributions (forward) is loaded now, behind the multiply
            const int cp_ = ncols >> 1;
            const bool writer = direct && (ncols == 1 ? lane == 0 : lane < cp_);
            int4 cw0 = make_int4(0, 0, 0, -1), cw1 = make_int4(0, 0, 0, -1);
            double pq0[M], pq1[M];
#pragma unroll
            for (int j = 0; j < M; ++j) pq0[j] = pq1[j] = 0.0;
            if (writer) {
              const int o0 = me.y + (ncols == 1 ? 0 : 2 * lane);
              cw0 = colA[o0];
              if (ncols > 1) cw1 = colA[o0 + 1];
              if (fwd) {
                const int32_t *colB = ix + e.o3;
                const int b0 = colB[o0], b1 = ncols > 1 ? colB[o0 + 1] : -1;
                double t0[M], t1[M], t2[M], t3[M];
                if (cw0.w >= 0) ld_vec<M>(a.z + int64_t(cw0.w) * M, t0);
                if (b0 >= 0) ld_vec<M>(a.z + int64_t(b0) * M, t1);
                if (ncols > 1 && cw1.w >= 0) ld_vec<M>(a.z + int64_t(cw1.w) * M, t2);
                if (b1 >= 0) ld_vec<M>(a.z + int64_t(b1) * M, t3);
#pragma unroll
                for (int j = 0; j < M; ++j) {
                  pq0[j] = (cw0.w >= 0 ? t0[j] : 0.0) + (b0 >= 0 ? t1[j] : 0.0);
                  pq1[j] = ((ncols > 1 && cw1.w >= 0) ? t2[j] : 0.0) + (b1 >= 0 ? t3[j] : 0.0);
                }
              }
            }
            auto put = [&](const int4 &cw, const double (&tot)[M], const double (&pq)[M]) {
              if (fwd) {
#pragma unroll
                for (int j = 0; j < M; ++j) a.z[int64_t(cw.z) * M + j] = tot[j] + pq[j];
              } else {
#pragma unroll
                for (int j = 0; j < M; ++j) a.x[int64_t(cw.z) * M + j] = tot[j];
                if (a.hist)
#pragma unroll
                  for (int j = 0; j < M; ++j)
                    if (a.hist[j]) {
                      double *hp = a.hist[j] + int64_t(level) * a.hist_stride + cw.z;
                      *hp = a.hist_add ? *hp + tot[j] : tot[j];
                    }
              }
            };
            if (ncols == 1) {
              const double *tp = tiles + md.x;
              double acc[M];
#pragma unroll
              for (int j = 0; j < M; ++j) acc[j] = 0.0;
              for (int r = lane; r < nrows; r += 32) {
                const double t = tp[r];
#pragma unroll
                for (int j = 0; j < M; ++j) acc[j] += t * vp[r * M + j];
              }
#pragma unroll
              for (int off = 16; off > 0; off >>= 1)
#pragma unroll
                for (int j = 0; j < M; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], off);
              if (writer) put(cw0, acc, pq0);
              else if (lane == 0)
#pragma unroll
                for (int j = 0; j < M; ++j) part[part_off * M + j] = acc[j];
            } else {
              const int cp = ncols >> 1;                  // lanes per row: each owns two adjacent columns
              const int lcp = 31 - __clz(cp);
              const int rg = 32 >> lcp;                   // rows in flight per warp instruction
              const int my_cp = lane & (cp - 1), my_rg = lane >> lcp;
              const double *tp = tiles + md.x + 2 * my_cp;
              double a0[M], a1[M], b0[M], b1[M];
#pragma unroll
              for (int j = 0; j < M; ++j) a0[j] = a1[j] = b0[j] = b1[j] = 0.0;
              int r = my_rg;
              for (; r + rg < nrows; r += 2 * rg) {
                const double2 t0 = *reinterpret_cast<const double2 *>(tp + r * ncols);
                const double2 t1 = *reinterpret_cast<const double2 *>(tp + (r + rg) * ncols);
                double v0[M], v1[M];
                lds_vec<M>(vp + r * M, v0);
                lds_vec<M>(vp + (r + rg) * M, v1);
#pragma unroll
                for (int j = 0; j < M; ++j) {
                  a0[j] += t0.x * v0[j];
                  a1[j] += t0.y * v0[j];
                  b0[j] += t1.x * v1[j];
                  b1[j] += t1.y * v1[j];
                }
              }
              if (r < nrows) {
                const double2 t0 = *reinterpret_cast<const double2 *>(tp + r * ncols);
                double v0[M];
                lds_vec<M>(vp + r * M, v0);
#pragma unroll
                for (int j = 0; j < M; ++j) {
                  a0[j] += t0.x * v0[j];
                  a1[j] += t0.y * v0[j];
                }
              }
#pragma unroll
              for (int j = 0; j < M; ++j) { a0[j] += b0[j]; a1[j] += b1[j]; }
              for (int off = cp; off < 32; off <<= 1)
#pragma unroll
                for (int j = 0; j < M; ++j) {
                  a0[j] += __shfl_xor_sync(0xffffffffu, a0[j], off);
                  a1[j] += __shfl_xor_sync(0xffffffffu, a1[j], off);
                }
              if (writer) {  // lanes 0 .. cp - 1 are row group 0: my_cp == lane
                put(cw0, a0, pq0);
                put(cw1, a1, pq1);
              } else if (my_rg == 0 && !direct)
#pragma unroll
                for (int j = 0; j < M; ++j) {
                  part[(part_off + 2 * my_cp) * M + j] = a0[j];
                  part[(part_off + 2 * my_cp + 1) * M + j] = a1[j];
                }
            }
          }
        }
        SWEEP_PROF(21)
        if (!(flags & SWF_DIRECT)) gsync();
        SWEEP_PROF(22)
        // outputs: the row parts of every column are added in a fixed order
        for (int o = gt, u = 0; o < n_cols && !(flags & SWF_DIRECT); o += GT, ++u) {
          const int4 ca = colA[o];
          const int np = ca.y & 0xFFFF, stride = ca.y >> 16;
          if (np == 0) continue;  // column of a direct micro-tile
          double tot[M];
#pragma unroll
          for (int j = 0; j < M; ++j) tot[j] = part[ca.x * M + j];
          for (int q = 1; q < np; ++q)
#pragma unroll
            for (int j = 0; j < M; ++j) tot[j] += part[(ca.x + q * stride) * M + j];
          if (fwd) {
            double pass[M];
            if (u == 0) {
#pragma unroll
              for (int j = 0; j < M; ++j) pass[j] = pv[0][j];
            } else if (u == 1) {
#pragma unroll
              for (int j = 0; j < M; ++j) pass[j] = pv[1][j];
            } else {  // more than PV * GT columns in one wave: load here
              const int c0 = ca.w, c1 = (ix + e.o3)[o];
              double p0[M], p1[M];
              if (c0 >= 0) ld_vec<M>(a.z + int64_t(c0) * M, p0);
              if (c1 >= 0) ld_vec<M>(a.z + int64_t(c1) * M, p1);
#pragma unroll
              for (int j = 0; j < M; ++j) pass[j] = (c0 >= 0 ? p0[j] : 0.0) + (c1 >= 0 ? p1[j] : 0.0);
            }
#pragma unroll
            for (int j = 0; j < M; ++j) a.z[int64_t(ca.z) * M + j] = tot[j] + pass[j];
          } else {
#pragma unroll
            for (int j = 0; j < M; ++j) a.x[int64_t(ca.z) * M + j] = tot[j];
            if (a.hist)
#pragma unroll
              for (int j = 0; j < M; ++j)
                if (a.hist[j]) {
                  double *hp = a.hist[j] + int64_t(level) * a.hist_stride + ca.z;
                  *hp = a.hist_add ? *hp + tot[j] : tot[j];
                }
          }
        }
        SWEEP_PROF(23)
      } else {
        // ---------------------------------------------------------------------------------------------------------
        // subtree: factor blob, index lists and this CTA's own vectors are in the ring; walk the levels on chip
        // ---------------------------------------------------------------------------------------------------------
        // (the boundary-table offset of this thread's pack member: an L2 round trip, started before the barrier)
        const int sub_early = gt / (GT / a.pack);
        const int bq_early = sub_early < e.a2 ? __ldg(a.inst_bq + e.a7 + sub_early) : -1;
        gsync();  // every warp is done with the workspace (and the template) of the previous entry
        if (e.a0 != cur_tmpl) {
          const int64_t off = a.tmpl_off[e.a0];
          const int n16 = reinterpret_cast<const int *>(a.tmpl16 + off)[8];
          const int4 *g4 = reinterpret_cast<const int4 *>(a.tmpl16 + off);
          int4 *s4 = reinterpret_cast<int4 *>(tb);
          for (int k = gt; k < (n16 + 7) / 8; k += GT) s4[k] = g4[k];
          cur_tmpl = e.a0;
          gsync();
        }
        const Tmpl T(tb);
        // pack: up to a.pack subtrees of this template, TPS threads each, walked in lockstep (same barriers)
        const int TPS = GT / a.pack, sub = gt / TPS, lt = gt - sub * TPS;
        const bool act = sub < e.a2;
        const int pos = e.a7 + (act ? sub : 0);
        const int bq = bq_early;
        const int ncell = T.w * T.h, ncell_e = (ncell + 1) & ~1, nint_e = (T.n_int + 1) & ~1;
        const int rf_size = reinterpret_cast<const int *>(tb)[6], rb_size = reinterpret_cast<const int *>(tb)[7];
        // pressures of the box (this CTA wrote them last step)
        const double *pc = reg + e.n[0] + e.n[1] + e.n[2] + sub * ncell_e * M;
        double *wsm = ws + sub * a.ws_sub_dbl;
        if (kind == SW_SUB_FWD) {
          const double *blob = reg + ((flags & SWF_SHARED) ? 0 : sub * rf_size);
          double *va = wsm;                  // right-hand sides of the interior variables (+ children, level by level)
          double *zl = va + T.n_int * M;     // boundary contributions of every node
          if (act)
          for (int l = lt; l < T.n_int; l += TPS) {
            const int code = T.var_code[l];
            const int type = code >> 14, dj = (code >> 7) & 127, di = code & 127;
            double val[M];
            if (type == 0) {
#pragma unroll
              for (int j = 0; j < M; ++j)
                val[j] = -((di > 0 ? a.kx0 * pc[(dj * T.w + di - 1) * M + j] : 0.0) +
                           (di < T.w ? a.kx1 * pc[(dj * T.w + di) * M + j] : 0.0));
            } else {
#pragma unroll
              for (int j = 0; j < M; ++j)
                val[j] = -((dj > 0 ? a.ky0 * pc[((dj - 1) * T.w + di) * M + j] : 0.0) +
                           (dj < T.h ? a.ky1 * pc[(dj * T.w + di) * M + j] : 0.0));
            }
            if (bq >= 0) {
#pragma unroll
              for (int j = 0; j < M; ++j) {
                const int q = a.sub_q[bq + l * M + j];
                if (q >= 0) {
                  const double *dp = a.bq_data[q];
                  if (dp) val[j] += a.bq_coef[q] * dp[int64_t(level) * a.bq_stride[q]];
                }
              }
            }
#pragma unroll
            for (int j = 0; j < M; ++j) va[l * M + j] = val[j];
          }
          gsync();
          SWEEP_PROF(11)
          // two phases per level: (i) the separator variables of the level's nodes collect their children's
          // contributions, (ii) every boundary output of every node: pass-through + (-G^T)^T-row times those values
          for (int lv = 0; lv < T.n_levels; ++lv) {
            if (lv > 0) {
              const int4 Lh = T.lvlb[lv];
              if (act)
              for (int v = Lh.x + lt; v < Lh.x + Lh.y; v += TPS) {
                const uint32_t cc = T.sc[v];
                const int k0 = cc & 0xFFFF, k1 = cc >> 16;
#pragma unroll
                for (int j = 0; j < M; ++j) {
                  double vk = va[v * M + j];
                  if (k0 != 0xFFFF) vk += zl[k0 * M + j];
                  if (k1 != 0xFFFF) vk += zl[k1 * M + j];
                  va[v * M + j] = vk;
                }
              }
              gsync();
            }
            if (act)
            for (int o = T.lvl_out[lv] + lt; o < T.lvl_out[lv + 1]; o += TPS) {
              const uint4 D = T.fdesc[o];
              const int rf_off = D.x & 0xFFFF, sb = D.x >> 16, sep0 = D.y & 0xFFFF;
              const int c0 = D.z & 0xFFFF, c1 = D.z >> 16;
              const int s = sb >> 8, b = sb & 255;
              const double *rf = blob + rf_off;
              const double *vv = va + sep0 * M;
              double av[M];
#pragma unroll
              for (int j = 0; j < M; ++j) {
                av[j] = 0.0;
                if (c0 != 0xFFFF) av[j] += zl[c0 * M + j];
                if (c1 != 0xFFFF) av[j] += zl[c1 * M + j];
              }
              int k = 0;
              for (; k + 4 <= s; k += 4) {
                const double r0 = rf[k * b], r1 = rf[(k + 1) * b], r2 = rf[(k + 2) * b], r3 = rf[(k + 3) * b];
#pragma unroll
                for (int j = 0; j < M; ++j)
                  av[j] += r0 * vv[k * M + j] + r1 * vv[(k + 1) * M + j] + r2 * vv[(k + 2) * M + j] + r3 * vv[(k + 3) * M + j];
              }
              for (; k < s; ++k) {
                const double r0 = rf[k * b];
#pragma unroll
                for (int j = 0; j < M; ++j) av[j] += r0 * vv[k * M + j];
              }
#pragma unroll
              for (int j = 0; j < M; ++j) zl[o * M + j] = av[j];
            }
            gsync();
          }
          SWEEP_PROF(12)
          if (act) {
            for (int l = lt; l < T.n_int * M; l += TPS) a.xi[int64_t(e.o1 + sub * nint_e) * M + l] = va[l];
            const int zr = T.zl_size - T.n_rootb;  // the root is the last node
            for (int q = lt; q < T.n_rootb * M; q += TPS) a.z[int64_t(e.o0 + sub * T.n_rootb) * M + q] = zl[zr * M + q];
          }
          fence_async_global();  // xi is streamed back by TMA in the backward entry
          SWEEP_PROF(13)
        } else {
          const double *blob = reg + ((flags & SWF_SHARED) ? 0 : sub * rb_size);
          const int xb_stride = (T.n_rootb + 3) & ~3;  // int32 entries per member
          const int *xb = reinterpret_cast<const int *>(reg + e.n[0]) + sub * xb_stride;
          const double *zs = reg + e.n[0] + e.n[1] + sub * nint_e * M;  // accumulated rhs of the interior variables
          double *xs = wsm;                         // solution: [interior | root boundary]
          if (act)
          for (int q = lt; q < T.n_rootb; q += TPS) {
            double v[M];
            ld_vec<M>(a.x + int64_t(xb[q]) * M, v);
#pragma unroll
            for (int j = 0; j < M; ++j) xs[(T.n_int + q) * M + j] = v[j];
          }
          gsync();
          SWEEP_PROF(8)
          // one phase per level: x_s = [F11^-1 | -G^T] [z_s ; x_b], 2^tl lanes per row, inputs read in place
          for (int lv = T.n_levels - 1; lv >= 0; --lv) {
            const int4 Lh = T.lvlb[lv];
            const int v0 = Lh.x, n_out = Lh.y, tl = Lh.z, estep = Lh.w;
            const int tpo = 1 << tl, per = TPS >> tl;
            const int grp = lt >> tl, rl = lt & (tpo - 1);
            if (act)  // whole warps: TPS is a multiple of 32
            for (int o0 = 0; o0 < n_out; o0 += per) {
              const int o = o0 + grp;
              double av[M];
#pragma unroll
              for (int j = 0; j < M; ++j) av[j] = 0.0;
              if (o < n_out) {
                const uint2 D = T.bdesc[v0 + o];
                const int row_off = D.x & 0xFFFF, len = D.x >> 16, in_off = D.y & 0xFFFF;
                const double *row = blob + row_off;
                const uint16_t *in = T.in_idx + in_off;
#pragma unroll 4
                for (int p = rl; p < len; p += tpo) {
                  const int sidx = in[p * estep];
                  const double r = row[p * estep];
                  const double *from = (sidx & 0x8000) ? xs + (sidx & 0x7fff) * M : zs + sidx * M;
#pragma unroll
                  for (int j = 0; j < M; ++j) av[j] += r * from[j];
                }
              }
              for (int off = tpo >> 1; off > 0; off >>= 1)
#pragma unroll
                for (int j = 0; j < M; ++j) av[j] += __shfl_xor_sync(0xffffffffu, av[j], off);
              if (o < n_out && rl == 0)
#pragma unroll
                for (int j = 0; j < M; ++j) xs[(v0 + o) * M + j] = av[j];
            }
            gsync();
          }
          SWEEP_PROF(9)
          int I0 = 0, J0 = 0;
          if (act && (a.hist || a.has_src)) { I0 = a.inst_ij[2 * pos]; J0 = a.inst_ij[2 * pos + 1]; }
          const int64_t pb = int64_t(e.o2) + sub * ncell_e;
          // pressure recovery p = p_old - M^-1 K_pu u for the cells of the box; next right-hand side source
          if (act)
          for (int c = lt; c < ncell; c += TPS) {
            const int el = T.cell_edge[4 * c], er = T.cell_edge[4 * c + 1], eb = T.cell_edge[4 * c + 2],
                      et = T.cell_edge[4 * c + 3];
            const double mi = a.minv ? a.minv[pb + c] : a.minv_const;
            int64_t pidx = 0;
            if (a.hist || a.has_src) {
              const int dj = c / T.w, di = c - dj * T.w;
              const int64_t cell = int64_t(J0 + dj) * a.nx + I0 + di;
              pidx = a.p_of_cell ? int64_t(a.p_of_cell[cell]) : cell;
            }
#pragma unroll
            for (int j = 0; j < M; ++j) {
              const double div = a.cl * xs[el * M + j] + a.cr * xs[er * M + j] + a.cb * xs[eb * M + j] +
                                 a.ct * xs[et * M + j];
              const double p = pc[c * M + j] - mi * div;
              if (a.hist && a.hist[j]) {
                double *hp = a.hist[j] + int64_t(level) * a.hist_stride + a.n_flux + pidx;
                *hp = a.hist_add ? *hp + p : p;
              }
              double nxt = p;
              if (a.has_src && a.src_next[j] && step + 1 < a.n_steps)
                nxt += mi * a.src_next[j][int64_t(level + 1) * a.n_pressure + pidx];
              a.ptil[(pb + c) * M + j] = nxt;
            }
          }
          fence_async_global();  // the box pressures are streamed back by TMA in the next step
          SWEEP_PROF(10)
          if (bq >= 0) {  // subdom_to_st_distribute: flux traces of the interface dofs
            for (int l = lt; l < T.n_int * M; l += TPS) {
              const int q = a.sub_q[bq + l];
              if (q >= 0) {
                double *tp = a.bq_trace[q];
                if (tp) tp[int64_t(level) * a.bq_stride[q]] = xs[l];
              }
            }
          }
          if (a.hist && act) {
            for (int l = lt; l < T.n_int; l += TPS) {
              const int64_t d = sweep_dof(a, I0, J0, T.var_code[l]);
#pragma unroll
              for (int j = 0; j < M; ++j)
                if (a.hist[j]) {
                  double *hp = a.hist[j] + int64_t(level) * a.hist_stride + d;
                  *hp = a.hist_add ? *hp + xs[l * M + j] : xs[l * M + j];
                }
            }
          }
        }
      }

      // ---- end of entry: every warp reports on its own ----------------------------------------------------------
      if (a.prof && tid == 0) {
        const long long t_end = clock64();
        a.prof[cta * 24 + kind] += t_end - t_entry;
        a.prof[cta * 24 + 6] += 1;
      }
      if (tracing) {
        long long *tr = a.trace + (int64_t(e_first) + i) * 3;
        tr[0] = g_entry; tr[1] = g_w; tr[2] = global_ns();  // nanoseconds of the global timer: comparable across SMs
      }
      // The LAST compute warp of the group to finish the entry publishes its completion counters itself (release at
      // GPU scope; the other warps' writes are ordered before it through the acq_rel update of the shared counter):
      // the consumers on other CTAs see the completion one polling hop earlier than through the retire warp, which
      // only frees the ring space and issues the next TMA copies.  The descriptor is read BEFORE the update: once the
      // count is complete the dependency warp may restage the slot.
      __syncwarp();
      if (lane == 0) {
        const int p_sig = e.sig, p_n = e.n_sig, p_off = e.sig_off;
        uint32_t old;
        asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;\n" : "=r"(old) : "r"(s_u32(done_cnt + slot)) : "memory");
        if (old + 1 == uint32_t(G / NB + 1) * uint32_t(CT / 32 / a.ng)) {
          // (-1: nobody on another CTA waits for that completion, sweep_plan.cpp)
          // one counter: release increment; several (option pub_fence): ONE release fence, then relaxed increments
          const int n_pub = (p_sig >= 0) + (p_n > 0 && sig0 >= 0) + (p_n > 1 && sig1 >= 0) + (p_n > 2 ? p_n - 2 : 0);
          const bool fence = a.pub_fence && n_pub > 1;
          if (fence) asm volatile("fence.acq_rel.gpu;\n" ::: "memory");
          auto pub = [&](int32_t id) {
            if (id < 0) return;
            if (fence) red_relaxed(a.cnt + id);
            else red_release(a.cnt + id);
          };
          pub(p_sig);
          if (p_n > 0) pub(sig0);
          if (p_n > 1) pub(sig1);
          for (int t = 2; t < p_n; ++t) pub(__ldg(a.index + p_off + t));
        }
      }
    }
  }
}

size_t sweep_smem_optin() {
  int dev = 0, optin = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  return size_t(optin);
}

int sweep_fixed_dbl(int, int) { return 19 * SWEEP_NBAR + SWEEP_GTAB_DBL; }

template <int M, int CT>
static int ctas_per_sm_t(size_t smem) {
  cudaFuncSetAttribute(k_sweep<M, CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_sweep<M, CT>, CT + 64, smem) != cudaSuccess) return 0;
  return n;
}

template <int M, int CT>
static cudaError_t launch_t(const SweepArgs &a, int32_t n_cta, size_t smem_bytes, cudaStream_t st) {
  void *args[] = {const_cast<SweepArgs *>(&a)};
  const void *fn = reinterpret_cast<const void *>(k_sweep<M, CT>);
  cudaError_t err = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_bytes));
  if (err != cudaSuccess) return err;
  return cudaLaunchCooperativeKernel(fn, dim3(n_cta), dim3(CT + 64), args, smem_bytes, st);
}

#define SWEEP_DISPATCH(M, CT, CALL)                                            \
  switch ((M) * 1000 + (CT)) {                                                 \
    case 1512: { constexpr int MM = 1, CC = 512; CALL; } break;                \
    case 2512: { constexpr int MM = 2, CC = 512; CALL; } break;                \
    case 4512: { constexpr int MM = 4, CC = 512; CALL; } break;                \
    case 1256: { constexpr int MM = 1, CC = 256; CALL; } break;                \
    case 2256: { constexpr int MM = 2, CC = 256; CALL; } break;                \
    case 4256: { constexpr int MM = 4, CC = 256; CALL; } break;                \
    case 1128: { constexpr int MM = 1, CC = 128; CALL; } break;                \
    case 2128: { constexpr int MM = 2, CC = 128; CALL; } break;                \
    case 4128: { constexpr int MM = 4, CC = 128; CALL; } break;                \
    case 1064: { constexpr int MM = 1, CC = 64; CALL; } break;                 \
    case 2064: { constexpr int MM = 2, CC = 64; CALL; } break;                 \
    case 4064: { constexpr int MM = 4, CC = 64; CALL; } break;                 \
    default: break;                                                            \
  }

int sweep_ctas_per_sm(int M, int ct, size_t smem_bytes) {
  int n = 0;
  SWEEP_DISPATCH(M, ct, (n = ctas_per_sm_t<MM, CC>(smem_bytes)));
  return n;
}

cudaError_t launch_sweep(int M, int ct, const SweepArgs &a, int32_t n_cta, size_t smem_bytes, cudaStream_t st) {
  cudaError_t err = cudaErrorInvalidValue;
  SWEEP_DISPATCH(M, ct, (err = launch_t<MM, CC>(a, n_cta, smem_bytes, st)));
  if (err == cudaSuccess) launch_counter_add(1);
  return err;
}

}  // namespace mmmfe
