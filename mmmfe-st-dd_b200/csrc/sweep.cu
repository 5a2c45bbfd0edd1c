// Persistent time-sweep kernel (see sweep.cuh) and the one-time repack of the factor into sweep layout.
#include "sweep.cuh"

#include "kernels.cuh"

namespace mmmfe {

// =================================================================================================================
// repack: old front storage -> contiguous column blocks
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

// =================================================================================================================
// device helpers
// =================================================================================================================
namespace {

__device__ __forceinline__ uint32_t s_u32(const void *p) { return uint32_t(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void bar_init(uint64_t *bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(s_u32(bar)) : "memory");
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

// Every spin in the kernel is bounded: after ~4 s (or when another CTA gave up) the waiter raises the abort flag
// (pinned host memory) and carries on with whatever data it has; the host reports the failure after the launch.
constexpr long long SWEEP_SPIN_LIMIT = 8000000000LL;

__device__ __forceinline__ bool spin_expired(long long t0, volatile int *abort_flag) {
  if (*abort_flag) return true;
  if (clock64() - t0 > SWEEP_SPIN_LIMIT) {
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

// arm the barrier with the byte count and start the bulk copies (one thread)
__device__ __forceinline__ void tma_issue(double *dst, const double *src, uint32_t bytes, uint64_t *bar,
                                          bool evict_first, uint64_t policy) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
  uint32_t done = 0;
  while (done < bytes) {
    const uint32_t n = min(bytes - done, 32768u);
    if (evict_first) {
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n" ::
              "r"(s_u32(reinterpret_cast<char *>(dst) + done)),
          "l"(reinterpret_cast<const char *>(src) + done), "r"(n), "r"(s_u32(bar)), "l"(policy)
          : "memory");
    } else {
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                       s_u32(reinterpret_cast<char *>(dst) + done)),
                   "l"(reinterpret_cast<const char *>(src) + done), "r"(n), "r"(s_u32(bar))
                   : "memory");
    }
    done += n;
  }
}

__device__ __forceinline__ uint32_t ld_acquire(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void red_release(uint32_t *p) {
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;\n" ::"l"(p) : "memory");
}

__device__ __forceinline__ void wait_counter(const uint32_t *p, uint32_t target, volatile int *abort_flag) {
  if (ld_acquire(p) >= target) return;
  const long long t0 = clock64();
  for (int n = 0; ld_acquire(p) < target; ++n) {
    __nanosleep(20);
    if ((n & 255) == 255 && spin_expired(t0, abort_flag)) return;
  }
}

// packed template accessors (uint16, 0xFFFF = none)
__device__ __forceinline__ int t16(const uint16_t *p, int i) {
  const int v = p[i];
  return v == 0xFFFF ? -1 : v;
}

struct Tmpl {  // views into the shared-memory copy of a packed template
  int n_nodes, n_levels, n_int, n_rootb, zl_size, rf_size, rb_size, w, h;
  const uint16_t *node_s, *node_b, *node_rf, *node_rb, *node_sep, *node_bnd, *node_zl, *var_code, *rootb_code,
      *sep_node, *sep_c0, *sep_c1, *out_node, *out_c0, *out_c1, *bnd_var, *lvl_sep, *lvl_out, *cell_edge;
  __device__ __forceinline__ explicit Tmpl(const uint16_t *tb) {
    const int *h32 = reinterpret_cast<const int *>(tb);
    n_nodes = h32[0]; n_levels = h32[1]; n_int = h32[2]; n_rootb = h32[3]; zl_size = h32[4]; rf_size = h32[5];
    rb_size = h32[6]; w = h32[7]; h = h32[8];
    node_s = tb + h32[10]; node_b = tb + h32[11]; node_rf = tb + h32[12]; node_rb = tb + h32[13];
    node_sep = tb + h32[14]; node_bnd = tb + h32[15]; node_zl = tb + h32[16]; var_code = tb + h32[17];
    rootb_code = tb + h32[18]; sep_node = tb + h32[19]; sep_c0 = tb + h32[20]; sep_c1 = tb + h32[21];
    out_node = tb + h32[22]; out_c0 = tb + h32[23]; out_c1 = tb + h32[24]; bnd_var = tb + h32[25];
    lvl_sep = tb + h32[26]; lvl_out = tb + h32[27]; cell_edge = tb + h32[28];
  }
};

__device__ __forceinline__ int64_t sweep_dof(const SweepArgs &a, int I0, int J0, int code) {
  const int type = code >> 14, dj = (code >> 7) & 127, di = code & 127;
  const int64_t canon = type == 0 ? int64_t(J0 + dj) * (a.nx + 1) + I0 + di : a.nxe + int64_t(J0 + dj) * a.nx + I0 + di;
  return a.dof_of_canon ? int64_t(a.dof_of_canon[canon]) : canon;
}

__device__ __forceinline__ void bar_wait_prof(uint64_t *bar, uint32_t parity, const SweepArgs &a, int cta, int tid) {
  if (a.prof && tid == 0) {
    const long long t0 = clock64();
    bar_wait(bar, parity, a.abort_flag);
    a.prof[cta * 8 + 5] += clock64() - t0;
  } else {
    bar_wait(bar, parity, a.abort_flag);
  }
}

}  // namespace

// =================================================================================================================
// the sweep kernel
// =================================================================================================================
template <int M>
__global__ void __launch_bounds__(SWEEP_THREADS, 2) k_sweep(const SweepArgs a) {
  extern __shared__ __align__(128) double sm[];
  uint64_t *bars = reinterpret_cast<uint64_t *>(sm);
  SweepEntry *ebuf = reinterpret_cast<SweepEntry *>(sm + SWEEP_NBAR);  // two staged entries
  double *red = sm + SWEEP_NBAR + 32;
  uint16_t *tb = reinterpret_cast<uint16_t *>(red + SWEEP_THREADS * M);
  double *ws = red + SWEEP_THREADS * M + a.tmpl_dbl;
  double *ring = ws + a.ws_dbl;
  __shared__ int cur_tmpl;

  const int tid = threadIdx.x;
  const int cta = blockIdx.x;
  const int e_first = a.cta_begin[cta];
  const int P = a.cta_begin[cta + 1] - e_first;
  if (P <= 0) return;
  const SweepEntry *ents = a.entries + e_first;
  const SweepIssue *iss = a.issue + e_first;
  const long long total = (long long)P * a.n_steps;

  uint64_t policy = 0;
  if (a.evict_first) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(policy));

  if (tid == 0) {
    for (int k = 0; k < SWEEP_NBAR; ++k) bar_init(bars + k);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    cur_tmpl = -1;
  }
  if (tid < 32) reinterpret_cast<int *>(&ebuf[0])[tid] = reinterpret_cast<const int *>(&ents[0])[tid];
  __syncthreads();

  // ---- TMA issue state (thread 0 only) ----
  long long issued = 0;
  int issue_i = 0;  // issued % P
  auto issue_one = [&](int64_t src, int32_t n_dbl, int32_t ring_off) {
    uint64_t *ib = bars + (issued % SWEEP_NBAR);
    if (n_dbl > 0) tma_issue(ring + ring_off, a.Rs + src, uint32_t(n_dbl) * 8u, ib, a.evict_first != 0, policy);
    else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(s_u32(ib)) : "memory");  // keeps the phase count
    ++issued;
    if (++issue_i == P) issue_i = 0;
  };
  if (tid == 0) {
    long long target = a.cta_prologue[cta];
    if (target > total) target = total;
    while (issued < target) {
      const SweepIssue s = iss[issue_i];
      issue_one(s.src, s.n_dbl, s.ring_off);
    }
  }

  double acc[M], pass[M];
  int out_dof = -1;
#pragma unroll
  for (int j = 0; j < M; ++j) acc[j] = pass[j] = 0.0;

  long long G = 0;
  for (int step = 0; step < a.n_steps; ++step) {
    const int level = a.level0 + step;
    for (int i = 0; i < P; ++i, ++G) {
      const SweepEntry &e = ebuf[G & 1];
      int pre = 0;
      if (tid >= 32 && tid < 64) {
        const int inext = (i + 1 == P) ? 0 : i + 1;
        pre = reinterpret_cast<const int *>(&ents[inext])[tid - 32];
      }
      const int kind = e.kind, flags = e.flags;
      const long long t_entry = (a.prof && tid == 0) ? clock64() : 0;
      const uint32_t parity = uint32_t((G / SWEEP_NBAR) & 1);
      uint64_t *bar = bars + (G % SWEEP_NBAR);

      if (kind == SW_TOP_FWD || kind == SW_TOP_BWD) {
        // ---------------------------------------------------------------------------------------------------------
        // column block of a top-level front
        // ---------------------------------------------------------------------------------------------------------
        const bool fwd = kind == SW_TOP_FWD;
        const int s = e.a4, b = e.a5, f = s + b;
        const int cw = e.a1, col0 = e.a0;  // cw: power of two below 32, any even width from 32 to 256
        if (flags & SWF_NEW_RUN) {
          if (tid == 0) {
            const long long t0 = a.prof ? clock64() : 0;
            if (e.dep0 >= 0) wait_counter(a.cnt + e.dep0, uint32_t(step + 1) * uint32_t(e.dep0_n), a.abort_flag);
            if (e.dep1 >= 0) wait_counter(a.cnt + e.dep1, uint32_t(step + 1) * uint32_t(e.dep1_n), a.abort_flag);
            if (a.prof) a.prof[cta * 8 + 4] += clock64() - t0;
          }
          __syncthreads();
          if (fwd) {
            // v = rhs of the separator dofs + children contributions (assemble_rhs_star fused: rhs = -K_up p_old)
            const double k0 = e.a6 ? a.ky0 : a.kx0, k1 = e.a6 ? a.ky1 : a.kx1;
            for (int p = tid; p < s; p += SWEEP_THREADS) {
              const int2 cc = a.top_cell[e.o0 + p];
              double val[M];
#pragma unroll
              for (int j = 0; j < M; ++j)
                val[j] = -(k0 * __ldcg(a.ptil + int64_t(cc.x) * M + j) + k1 * __ldcg(a.ptil + int64_t(cc.y) * M + j));
              const int q0 = a.src[e.o1 + p], q1 = a.src[e.o1 + f + p];
              if (q0 >= 0 && e.o4 >= 0)
#pragma unroll
                for (int j = 0; j < M; ++j) val[j] += __ldcg(a.z + int64_t(e.o4 + q0) * M + j);
              if (q1 >= 0 && e.o5 >= 0)
#pragma unroll
                for (int j = 0; j < M; ++j) val[j] += __ldcg(a.z + int64_t(e.o5 + q1) * M + j);
#pragma unroll
              for (int j = 0; j < M; ++j) ws[p * M + j] = val[j];
              if (flags & SWF_WRITE_ZS)
#pragma unroll
                for (int j = 0; j < M; ++j) a.z[int64_t(e.o2 + p) * M + j] = val[j];
            }
          } else {
            // w = [accumulated separator rhs ; solution of the boundary dofs (ancestors' separators)]
            for (int p = tid; p < f; p += SWEEP_THREADS) {
              if (p < s) {
#pragma unroll
                for (int j = 0; j < M; ++j) ws[p * M + j] = __ldcg(a.z + int64_t(e.o2 + p) * M + j);
              } else {
                const int64_t d = a.idx_user[e.o0 + p];
#pragma unroll
                for (int j = 0; j < M; ++j) ws[p * M + j] = __ldcg(a.x + d * M + j);
              }
            }
          }
          __syncthreads();
        }
        if (flags & SWF_FIRST) {
#pragma unroll
          for (int j = 0; j < M; ++j) acc[j] = pass[j] = 0.0;
          out_dof = -1;
          if (tid < cw) {
            if (fwd) {  // pass-through of the children's contributions to this boundary dof
              const int q = col0 + tid;
              if (q < b) {
                const int q0 = a.src[e.o1 + s + q], q1 = a.src[e.o1 + f + s + q];
                if (q0 >= 0 && e.o4 >= 0)
#pragma unroll
                  for (int j = 0; j < M; ++j) pass[j] += __ldcg(a.z + int64_t(e.o4 + q0) * M + j);
                if (q1 >= 0 && e.o5 >= 0)
#pragma unroll
                  for (int j = 0; j < M; ++j) pass[j] += __ldcg(a.z + int64_t(e.o5 + q1) * M + j);
              }
            } else if (col0 + tid < s) {
              out_dof = a.idx_user[e.o0 + col0 + tid];
            }
          }
        }
        if (e.n_dbl > 0) {
          bar_wait_prof(bar, parity, a, cta, tid);
          const int g = tid / cw, col = tid - g * cw, gn = SWEEP_THREADS / cw;
          if (g < gn) {
            const double *tile = ring + e.ring_off + col;
            const int r_beg = e.a2, r_end = e.a3;
            int r = r_beg + g;
            for (; r + 3 * gn < r_end; r += 4 * gn) {
              const double m0 = tile[(r - r_beg) * cw], m1 = tile[(r + gn - r_beg) * cw];
              const double m2 = tile[(r + 2 * gn - r_beg) * cw], m3 = tile[(r + 3 * gn - r_beg) * cw];
#pragma unroll
              for (int j = 0; j < M; ++j)
                acc[j] += m0 * ws[r * M + j] + m1 * ws[(r + gn) * M + j] + m2 * ws[(r + 2 * gn) * M + j] +
                          m3 * ws[(r + 3 * gn) * M + j];
            }
            for (; r < r_end; r += gn) {
              const double m0 = tile[(r - r_beg) * cw];
#pragma unroll
              for (int j = 0; j < M; ++j) acc[j] += m0 * ws[r * M + j];
            }
          }
        }
        if (flags & SWF_LAST) {
          if (cw < 32)
            for (int off = 16; off >= cw; off >>= 1)
#pragma unroll
              for (int j = 0; j < M; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], off);
#pragma unroll
          for (int j = 0; j < M; ++j) red[tid * M + j] = acc[j];
          __syncthreads();
          if (tid < cw) {
            const int stride = cw < 32 ? 32 : cw;
            double tot[M];
#pragma unroll
            for (int j = 0; j < M; ++j) tot[j] = 0.0;
            for (int t = tid; t < SWEEP_THREADS; t += stride)
#pragma unroll
              for (int j = 0; j < M; ++j) tot[j] += red[t * M + j];
            if (fwd) {
              const int q = col0 + tid;
              if (q < b)
#pragma unroll
                for (int j = 0; j < M; ++j) a.z[int64_t(e.o3 + q) * M + j] = tot[j] + pass[j];
            } else if (out_dof >= 0) {
#pragma unroll
              for (int j = 0; j < M; ++j) a.x[int64_t(out_dof) * M + j] = tot[j];
              if (a.hist)
#pragma unroll
                for (int j = 0; j < M; ++j)
                  if (a.hist[j]) {
                    double *hp = a.hist[j] + int64_t(level) * a.hist_stride + out_dof;
                    *hp = a.hist_add ? *hp + tot[j] : tot[j];
                  }
            }
          }
        }
      } else {
        // ---------------------------------------------------------------------------------------------------------
        // subtree: the whole factor blob of the box is in the ring; walk its levels in shared memory
        // ---------------------------------------------------------------------------------------------------------
        if (e.a0 != cur_tmpl) {  // uniform branch: cur_tmpl only changes behind the barrier below
          const int64_t off = a.tmpl_off[e.a0];
          const int n16 = reinterpret_cast<const int *>(a.tmpl16 + off)[9];
          const int4 *g4 = reinterpret_cast<const int4 *>(a.tmpl16 + off);
          int4 *s4 = reinterpret_cast<int4 *>(tb);
          for (int k = tid; k < (n16 + 7) / 8; k += SWEEP_THREADS) s4[k] = g4[k];
          __syncthreads();
          if (tid == 0) cur_tmpl = e.a0;
        }
        const Tmpl T(tb);
        const int I0 = e.a2, J0 = e.a3, bq = e.a1;
        const int ncell = T.w * T.h;
        double *pc = ws;
        if (kind == SW_SUB_FWD) {
          double *vs = pc + ncell * M;
          double *zl = vs + T.n_int * M;
          for (int c = tid; c < ncell; c += SWEEP_THREADS) {
            const int dj = c / T.w, di = c - dj * T.w;
            const int64_t cell = int64_t(J0 + dj) * a.nx + I0 + di;
#pragma unroll
            for (int j = 0; j < M; ++j) pc[c * M + j] = __ldcg(a.ptil + cell * M + j);
          }
          __syncthreads();
          for (int l = tid; l < T.n_int; l += SWEEP_THREADS) {
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
            for (int j = 0; j < M; ++j) vs[l * M + j] = val[j];
          }
          if (e.n_dbl > 0) bar_wait_prof(bar, parity, a, cta, tid);
          __syncthreads();
          const double *blob = ring + e.ring_off;
          for (int lv = 0; lv < T.n_levels; ++lv) {
            if (lv > 0) {  // leaves have no children
              for (int v = T.lvl_sep[lv] + tid; v < T.lvl_sep[lv + 1]; v += SWEEP_THREADS) {
                const int c0 = t16(T.sep_c0, v), c1 = t16(T.sep_c1, v);
#pragma unroll
                for (int j = 0; j < M; ++j) {
                  double val = vs[v * M + j];
                  if (c0 >= 0) val += zl[c0 * M + j];
                  if (c1 >= 0) val += zl[c1 * M + j];
                  vs[v * M + j] = val;
                }
              }
              __syncthreads();
            }
            for (int o = T.lvl_out[lv] + tid; o < T.lvl_out[lv + 1]; o += SWEEP_THREADS) {
              const int n = T.out_node[o];
              const int s = T.node_s[n], b = T.node_b[n];
              const double *rf = blob + T.node_rf[n] + (o - T.node_zl[n]);
              const double *v = vs + int(T.node_sep[n]) * M;
              const int c0 = t16(T.out_c0, o), c1 = t16(T.out_c1, o);
              double av[M];
#pragma unroll
              for (int j = 0; j < M; ++j) {
                av[j] = 0.0;
                if (c0 >= 0) av[j] += zl[c0 * M + j];
                if (c1 >= 0) av[j] += zl[c1 * M + j];
              }
              for (int k = 0; k < s; ++k) {
                const double r = rf[k * b];
#pragma unroll
                for (int j = 0; j < M; ++j) av[j] += r * v[k * M + j];
              }
#pragma unroll
              for (int j = 0; j < M; ++j) zl[o * M + j] = av[j];
            }
            __syncthreads();
          }
          for (int l = tid; l < T.n_int * M; l += SWEEP_THREADS) a.xi[int64_t(e.o1) * M + l] = vs[l];
          const int zr = T.node_zl[T.n_nodes - 1];
          for (int q = tid; q < T.n_rootb * M; q += SWEEP_THREADS) a.z[int64_t(e.o0) * M + q] = zl[zr * M + q];
        } else {
          double *zs = pc + ncell * M;        // accumulated right-hand sides of the interior variables
          double *xs = zs + T.n_int * M;      // solution: [interior | root boundary]
          for (int l = tid; l < T.n_int * M; l += SWEEP_THREADS) zs[l] = __ldcg(a.xi + int64_t(e.o1) * M + l);
          for (int c = tid; c < ncell; c += SWEEP_THREADS) {
            const int dj = c / T.w, di = c - dj * T.w;
            const int64_t cell = int64_t(J0 + dj) * a.nx + I0 + di;
#pragma unroll
            for (int j = 0; j < M; ++j) pc[c * M + j] = __ldcg(a.ptil + cell * M + j);
          }
          if (tid == 0 && e.dep0 >= 0) {
            const long long t0 = a.prof ? clock64() : 0;
            wait_counter(a.cnt + e.dep0, uint32_t(step + 1) * uint32_t(e.dep0_n), a.abort_flag);
            if (a.prof) a.prof[cta * 8 + 4] += clock64() - t0;
          }
          __syncthreads();
          for (int q = tid; q < T.n_rootb; q += SWEEP_THREADS) {
            const int64_t d = sweep_dof(a, I0, J0, T.rootb_code[q]);
#pragma unroll
            for (int j = 0; j < M; ++j) xs[(T.n_int + q) * M + j] = __ldcg(a.x + d * M + j);
          }
          if (e.n_dbl > 0) bar_wait_prof(bar, parity, a, cta, tid);
          __syncthreads();
          const double *blob = ring + e.ring_off;
          for (int lv = T.n_levels - 1; lv >= 0; --lv) {
            for (int v = T.lvl_sep[lv] + tid; v < T.lvl_sep[lv + 1]; v += SWEEP_THREADS) {
              const int n = T.sep_node[v];
              const int s = T.node_s[n], b = T.node_b[n], sep0 = T.node_sep[n];
              const double *rb = blob + T.node_rb[n] + (v - sep0);  // R^T[p][i] at p * s + i
              const uint16_t *bv = T.bnd_var + T.node_bnd[n];
              double av[M];
#pragma unroll
              for (int j = 0; j < M; ++j) av[j] = 0.0;
              for (int p = 0; p < s; ++p) {
                const double r = rb[p * s];
#pragma unroll
                for (int j = 0; j < M; ++j) av[j] += r * zs[(sep0 + p) * M + j];
              }
              rb += s * s;
              for (int q = 0; q < b; ++q) {
                const double r = rb[q * s];
                const int l = bv[q];
#pragma unroll
                for (int j = 0; j < M; ++j) av[j] += r * xs[l * M + j];
              }
#pragma unroll
              for (int j = 0; j < M; ++j) xs[v * M + j] = av[j];
            }
            __syncthreads();
          }
          // pressure recovery p = p_old - M^-1 K_pu u for the cells of the box; next right-hand side source
          for (int c = tid; c < ncell; c += SWEEP_THREADS) {
            const int dj = c / T.w, di = c - dj * T.w;
            const int64_t cell = int64_t(J0 + dj) * a.nx + I0 + di;
            const int el = T.cell_edge[4 * c], er = T.cell_edge[4 * c + 1], eb = T.cell_edge[4 * c + 2],
                      et = T.cell_edge[4 * c + 3];
            const double mi = a.minv[cell];
            const int64_t pidx = a.p_of_cell ? int64_t(a.p_of_cell[cell]) : cell;
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
              a.ptil[cell * M + j] = nxt;
            }
          }
          if (bq >= 0) {  // subdom_to_st_distribute: flux traces of the interface dofs
            for (int l = tid; l < T.n_int * M; l += SWEEP_THREADS) {
              const int q = a.sub_q[bq + l];
              if (q >= 0) {
                double *tp = a.bq_trace[q];
                if (tp) tp[int64_t(level) * a.bq_stride[q]] = xs[l];
              }
            }
          }
          if (a.hist) {
            for (int l = tid; l < T.n_int; l += SWEEP_THREADS) {
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

      // ---- end of entry: stage the next descriptor, publish completion, release ring space --------------------
      const int sig = (flags & SWF_LAST) ? e.sig : -1;
      const long long upto = (long long)step * P + e.issue_upto;
      const int64_t t_src = e.t_src;
      const int32_t t_n = e.t_n_dbl, t_ro = e.t_ring_off;
      if (tid >= 32 && tid < 64) reinterpret_cast<int *>(&ebuf[(G + 1) & 1])[tid - 32] = pre;
      __syncthreads();
      if (tid == 0) {
        if (a.prof) {
          a.prof[cta * 8 + kind] += clock64() - t_entry;
          a.prof[cta * 8 + 6] += 1;
        }
        if (sig >= 0) {
          __threadfence();
          red_release(a.cnt + sig);
        }
        const long long target = upto < total ? upto : total;
        if (issued < target) {
          issue_one(t_src, t_n, t_ro);
          while (issued < target) {
            const SweepIssue s = iss[issue_i];
            issue_one(s.src, s.n_dbl, s.ring_off);
          }
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

template <int M>
static int ctas_per_sm_t(size_t smem) {
  cudaFuncSetAttribute(k_sweep<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_sweep<M>, SWEEP_THREADS, smem) != cudaSuccess) return 0;
  return n;
}

int sweep_ctas_per_sm(int M, size_t smem_bytes) {
  switch (M) {
    case 1: return ctas_per_sm_t<1>(smem_bytes);
    case 2: return ctas_per_sm_t<2>(smem_bytes);
    case 4: return ctas_per_sm_t<4>(smem_bytes);
    default: return 0;
  }
}

cudaError_t launch_sweep(int M, const SweepArgs &a, int32_t n_cta, size_t smem_bytes, cudaStream_t st) {
  void *args[] = {const_cast<SweepArgs *>(&a)};
  const void *fn = nullptr;
  switch (M) {
    case 1: fn = reinterpret_cast<const void *>(k_sweep<1>); break;
    case 2: fn = reinterpret_cast<const void *>(k_sweep<2>); break;
    case 4: fn = reinterpret_cast<const void *>(k_sweep<4>); break;
    default: return cudaErrorInvalidValue;
  }
  cudaError_t err = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_bytes));
  if (err != cudaSuccess) return err;
  err = cudaLaunchCooperativeKernel(fn, dim3(n_cta), dim3(SWEEP_THREADS), args, smem_bytes, st);
  if (err == cudaSuccess) launch_counter_add(1);
  return err;
}

}  // namespace mmmfe
