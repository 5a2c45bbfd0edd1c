// CUDA kernels of the interface-GMRES engine, written for sm_100a (B200).
//
//  * factorisation: grouped FP64 GEMM on the tensor cores (DMMA, mma.sync.m8n8k4.f64 -- tcgen05 has no f64
//    kind), warp-level Gauss-Jordan for the <=32x32 base blocks, front assembly / extend-add;
//  * solve: level-scheduled block-inverse multifrontal sweeps (every front is one dense GEMV, no triangular
//    solves), batched over the M subdomains that share a factor; HBM-bound, coalesced row streaming;
//  * time stepping, projector SpMV, exchange/combine and the classical Gram-Schmidt kernels.
#include "kernels.cuh"
#include "dev_common.cuh"

#include <algorithm>
#include <atomic>

namespace mmmfe {

static std::atomic<long long> g_launches{0};
long long launch_counter() { return g_launches.load(); }
void launch_counter_add(long long n) { g_launches.fetch_add(n); }

// =================================================================================================================
// grouped GEMM  C = alpha * A * B + beta * C   (FP64, DMMA)
// =================================================================================================================
constexpr int GT = 64;         // C tile
constexpr int GK = 16;         // k step
constexpr int AS_LD = GK + 4;  // row strides of the staged tiles: fragment loads of a half-warp hit 16 distinct banks
constexpr int BS_LD = GT + 4;
constexpr int G_STAGES = 4;    // k slabs in flight (cp.async groups)
constexpr int G_STAGE_DBL = GT * AS_LD + GK * BS_LD;
constexpr size_t G_SMEM = size_t(G_STAGES) * G_STAGE_DBL * sizeof(double);

// One 64 x 64 tile of one problem per CTA, 8 warps (4 x 2, 16 x 32 each).  A and B are staged through a four-deep
// cp.async pipeline (three k slabs in flight while one is multiplied, one barrier per k step), so the many short
// problems of the factorisation and of the multi-column solve are not bound by the global-load latency of a k step;
// the copy mapping follows the unit stride of each operand so that every layout is read in full 128-byte lines.
__global__ void __launch_bounds__(256) k_grouped_gemm(const GemmProb *__restrict__ probs,
                                                      const int32_t *__restrict__ tile_start, int32_t n_probs) {
  extern __shared__ __align__(16) double gsm[];
  // locate the problem of this tile
  int lo = 0, hi = n_probs - 1;
  const int tile = blockIdx.x;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tile_start[mid] <= tile) lo = mid; else hi = mid - 1;
  }
  const GemmProb pr = probs[lo];
  const int local = tile - tile_start[lo];
  const int tiles_n = (pr.n + GT - 1) / GT;
  const int tm = local / tiles_n, tn = local % tiles_n;
  const int row0 = tm * GT, col0 = tn * GT;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1;
  double acc[2][4][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const bool a_rowmajor = (pr.acs == 1);
  const bool b_rowmajor = (pr.bcs == 1);
  // this thread's four elements of each staged tile
  int ai[4], ak[4], bk[4], bj[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int e = tid + 256 * r;
    if (a_rowmajor) { ai[r] = e / GK; ak[r] = e % GK; } else { ai[r] = e % GT; ak[r] = e / GT; }
    if (b_rowmajor) { bk[r] = e / GT; bj[r] = e % GT; } else { bk[r] = e % GK; bj[r] = e / GK; }
  }
  auto issue = [&](int stage, int k0) {
    double *as = gsm + stage * G_STAGE_DBL, *bs = as + GT * AS_LD;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int gi = row0 + ai[r], gk = k0 + ak[r];
      const bool pa = gi < pr.m && gk < pr.k;
      cp_async8(as + ai[r] * AS_LD + ak[r], pa ? pr.A + gi * pr.ars + gk * pr.acs : pr.A, pa);
      const int gj = col0 + bj[r], gk2 = k0 + bk[r];
      const bool pb = gj < pr.n && gk2 < pr.k;
      cp_async8(bs + bk[r] * BS_LD + bj[r], pb ? pr.B + gk2 * pr.brs + gj * pr.bcs : pr.B, pb);
    }
  };
  const int nk = (pr.k + GK - 1) / GK;
#pragma unroll
  for (int st = 0; st < G_STAGES - 1; ++st) {
    if (st < nk) issue(st, st * GK);
    cp_async_commit();
  }
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<G_STAGES - 2>();
    __syncthreads();  // slab kt has landed for every thread; everyone is done with the buffer refilled next
    if (kt + G_STAGES - 1 < nk) issue((kt + G_STAGES - 1) % G_STAGES, (kt + G_STAGES - 1) * GK);
    cp_async_commit();
    const double *as = gsm + (kt % G_STAGES) * G_STAGE_DBL, *bs = as + GT * AS_LD;
#pragma unroll
    for (int kk = 0; kk < GK; kk += 4) {
      double a[2], b[4];
#pragma unroll
      for (int i = 0; i < 2; ++i) a[i] = as[(wm * 16 + i * 8 + (lane >> 2)) * AS_LD + kk + (lane & 3)];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = bs[(kk + (lane & 3)) * BS_LD + wn * 32 + j * 8 + (lane >> 2)];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gi = row0 + wm * 16 + i * 8 + (lane >> 2);
      const int gj = col0 + wn * 32 + j * 8 + (lane & 3) * 2;
      if (gi < pr.m) {
        double *crow = pr.C + int64_t(pr.crow ? pr.crow[gi] : gi) * pr.ldc;
#pragma unroll
        for (int t = 0; t < 2; ++t)
          if (gj + t < pr.n) {
            double *c = crow + gj + t;
            const double old = (pr.beta == 0.0) ? 0.0 : pr.beta * (*c);
            *c = pr.alpha * acc[i][j][t] + old;
          }
      }
    }
}

void launch_grouped_gemm(const GemmProb *d_probs, const int32_t *d_tile_start, int32_t n_probs, int32_t n_tiles,
                         cudaStream_t st) {
  if (n_tiles <= 0) return;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(k_grouped_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, int(G_SMEM));
    attr_set = true;
  }
  k_grouped_gemm<<<n_tiles, 256, G_SMEM, st>>>(d_probs, d_tile_start, n_probs);
  launch_counter_add(1);
}

// =================================================================================================================
// base case of the factorisation: A (SPD, n <= 32) -> inverse Cholesky factor L^-1 (lower triangle, zeros above),
// one warp per block.  Forming Schur complements through L^-1 (condition sqrt(cond A)) instead of A^-1 keeps the
// factorisation as accurate as a Cholesky multifrontal method.
// =================================================================================================================
__global__ void __launch_bounds__(64) k_small_chol_inv(const InvProb *__restrict__ probs, int32_t n_probs,
                                                       int32_t *fail) {
  __shared__ double sa[2][32][33];
  __shared__ double sb[2][32][33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pi = blockIdx.x * 2 + warp;
  if (pi >= n_probs) return;
  const InvProb pr = probs[pi];
  const int n = pr.n;
  double(*a)[33] = sa[warp];
  double(*li)[33] = sb[warp];
  for (int i = 0; i < n; ++i)
    if (lane < n) a[i][lane] = pr.A[int64_t(i) * pr.ld + lane];
  __syncwarp();
  // right-looking Cholesky on the lower triangle; lane = row
  for (int k = 0; k < n; ++k) {
    const double d = a[k][k];
    if (!(d > 0.0) && lane == 0) atomicExch(fail, 1);
    const double l = sqrt(d);
    __syncwarp();
    if (lane == k) a[k][k] = l;
    else if (lane > k && lane < n) a[lane][k] /= l;
    __syncwarp();
    for (int j = k + 1; j < n; ++j)
      if (lane >= j && lane < n) a[lane][j] -= a[lane][k] * a[j][k];
    __syncwarp();
  }
  // L^-1 by forward substitution; lane = column
  for (int i = 0; i < n; ++i) {
    double v = (lane == i) ? 1.0 : 0.0;
    for (int k = 0; k < i; ++k)
      if (k >= lane) v -= a[i][k] * li[k][lane];
    if (lane < n) li[i][lane] = (lane <= i) ? v / a[i][i] : 0.0;
    __syncwarp();
  }
  for (int i = 0; i < n; ++i)
    if (lane < n) pr.A[int64_t(i) * pr.ld + lane] = li[i][lane];
}

void launch_small_inverse(const InvProb *d_probs, int32_t n_probs, int32_t *d_fail, cudaStream_t st) {
  if (n_probs <= 0) return;
  k_small_chol_inv<<<(n_probs + 1) / 2, 64, 0, st>>>(d_probs, n_probs, d_fail);
  launch_counter_add(1);
}

__global__ void __launch_bounds__(256) k_transpose(const TransProb *__restrict__ probs) {
  const TransProb pr = probs[blockIdx.x];
  for (int i = blockIdx.y; i < pr.rows; i += gridDim.y)
    for (int j = threadIdx.x; j < pr.cols; j += blockDim.x)
      pr.dst[int64_t(i) * pr.ld_dst + j] = pr.scale * pr.src[int64_t(j) * pr.ld_src + i];
}

void launch_transpose(const TransProb *d_probs, int32_t n_probs, cudaStream_t st) {
  if (n_probs <= 0) return;
  for (int32_t off = 0; off < n_probs; off += 32768) {
    const int32_t cnt = n_probs - off < 32768 ? n_probs - off : 32768;
    k_transpose<<<dim3(cnt, 16), 256, 0, st>>>(d_probs + off);
    launch_counter_add(1);
  }
}

// =================================================================================================================
// front assembly
// =================================================================================================================
__global__ void __launch_bounds__(256)
k_assemble_original(int64_t n_flux, const int64_t *__restrict__ s_rp, const int32_t *__restrict__ s_col,
                    const double *__restrict__ s_val, const int32_t *__restrict__ canon_of_dof,
                    const int32_t *__restrict__ node_of, const int32_t *__restrict__ pos_of,
                    const FrontDesc *__restrict__ fronts, const int64_t *__restrict__ f_off,
                    const int32_t *__restrict__ idx_canon, double *F) {
  const int64_t v = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (v >= n_flux) return;
  const int32_t cv = canon_of_dof ? canon_of_dof[v] : int32_t(v);
  const int32_t node = node_of[cv];
  const int32_t p = pos_of[cv];
  const FrontDesc fd = fronts[node];
  const int64_t f = fd.s + fd.b;
  double *Fp = F + f_off[node];
  const int32_t *bnd = idx_canon + fd.idx_off + fd.s;
  for (int64_t e = s_rp[v]; e < s_rp[v + 1]; ++e) {
    const int32_t c = s_col[e];
    const int32_t cc = canon_of_dof ? canon_of_dof[c] : c;
    const double val = s_val[e];
    if (node_of[cc] == node) {
      Fp[p * f + pos_of[cc]] = val;
    } else {
      int lo = 0, hi = fd.b - 1, q = -1;
      while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        const int32_t t = bnd[mid];
        if (t == cc) { q = mid; break; }
        if (t < cc) lo = mid + 1; else hi = mid - 1;
      }
      if (q >= 0) {
        Fp[p * f + fd.s + q] = val;
        Fp[(fd.s + q) * f + p] = val;
      }
    }
  }
}

void launch_assemble_original(int64_t n_flux, const int64_t *s_rp, const int32_t *s_col, const double *s_val,
                              const int32_t *canon_of_dof, const int32_t *, const int32_t *node_of,
                              const int32_t *pos_of, const FrontDesc *fronts, const int64_t *f_off,
                              const int32_t *idx_canon, double *F, cudaStream_t st) {
  k_assemble_original<<<unsigned((n_flux + 255) / 256), 256, 0, st>>>(n_flux, s_rp, s_col, s_val, canon_of_dof,
                                                                      node_of, pos_of, fronts, f_off, idx_canon, F);
  launch_counter_add(1);
}

__global__ void __launch_bounds__(256)
k_extend_add(const int32_t *__restrict__ nodes, const FrontDesc *__restrict__ fronts,
             const int64_t *__restrict__ f_off, const int32_t *__restrict__ src, double *F) {
  const int32_t node = nodes[blockIdx.x];
  const FrontDesc fd = fronts[node];
  const int32_t f = fd.s + fd.b;
  double *Fp = F + f_off[node];
  for (int c = 0; c < 2; ++c) {
    const int32_t ch = c ? fd.child1 : fd.child0;
    if (ch < 0) continue;
    const FrontDesc cd = fronts[ch];
    const int32_t fc = cd.s + cd.b;
    const double *Fc = F + f_off[ch];
    const int32_t *sm = src + fd.src_off + int64_t(c) * f;
    for (int p1 = blockIdx.y; p1 < f; p1 += gridDim.y) {
      const int32_t q1 = sm[p1];
      if (q1 < 0) continue;
      for (int p2 = threadIdx.x; p2 < f; p2 += blockDim.x) {
        const int32_t q2 = sm[p2];
        if (q2 >= 0) Fp[int64_t(p1) * f + p2] += Fc[int64_t(cd.s + q1) * fc + cd.s + q2];
      }
    }
  }
}

void launch_extend_add(const int32_t *d_nodes, int32_t n_nodes, const FrontDesc *fronts, const int64_t *f_off,
                       const int32_t *src, double *F, cudaStream_t st) {
  for (int32_t off = 0; off < n_nodes; off += 32768) {
    const int32_t cnt = n_nodes - off < 32768 ? n_nodes - off : 32768;
    k_extend_add<<<dim3(cnt, 32), 256, 0, st>>>(d_nodes + off, fronts, f_off, src, F);
    launch_counter_add(1);
  }
}

// =================================================================================================================
// multifrontal solve, top of the tree: one launch per level and direction
// =================================================================================================================
template <int M>
__device__ __forceinline__ void gather_children(const SolveArgs &a, const FrontDesc &fd, int32_t f, int32_t p,
                                                double (&val)[M]) {
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int32_t ch = c ? fd.child1 : fd.child0;
    if (ch < 0) continue;
    const int32_t q = a.src[fd.src_off + int64_t(c) * f + p];
    if (q < 0) continue;
    const FrontDesc cd = a.fronts[ch];
    const int64_t base = cd.z_off + cd.s + q;
    for (int sp = 0; sp < cd.nsplit; ++sp)
#pragma unroll
      for (int j = 0; j < M; ++j) val[j] += a.z[(base + int64_t(sp) * cd.b) * M + j];
  }
}

constexpr int FWD_ROWS = 32;  // rows of -G^T per forward item (== NdNode::nsplit granularity)

// item = rows [r0,r1) x boundary columns [c0,c1) of one front: partial contribution `split`.
// All (<= 32) matrix loads of a thread are issued before anything else, so the gather chain of the input vector
// (index -> child descriptor -> child contribution) overlaps the HBM latency of the factor.
template <int M>
__global__ void __launch_bounds__(256) k_forward_large(SolveArgs a, const WorkItem *__restrict__ items) {
  __shared__ double vs[FWD_ROWS * M];
  __shared__ double red[256 * M];
  const WorkItem it = items[blockIdx.x];
  const FrontDesc fd = a.fronts[it.node];
  const int32_t s = fd.s, f = fd.s + fd.b;
  const int tid = threadIdx.x;
  const int nr = it.r1 - it.r0;
  const int nc = it.c1 - it.c0;
  int cw = 32;
  while (cw < nc && cw < 256) cw <<= 1;
  const int rg = 256 / cw;
  const int cid = tid % cw, rgid = tid / cw;
  double rv[FWD_ROWS];
  {
    const double *Rc = a.R + fd.rf_off + int64_t(it.r0) * fd.ldf + it.c0 + cid;
#pragma unroll
    for (int k = 0; k < FWD_ROWS; ++k) {
      const int i = rgid + k * rg;
      rv[k] = (cid < nc && i < nr) ? __ldg(Rc + int64_t(i) * fd.ldf) : 0.0;
    }
  }
  if (tid < nr) {
    const int32_t p = it.r0 + tid;
    double val[M];
    const int64_t d = a.idx[fd.idx_off + p];
#pragma unroll
    for (int j = 0; j < M; ++j) val[j] = a.x[d * M + j];
    gather_children<M>(a, fd, f, p, val);
#pragma unroll
    for (int j = 0; j < M; ++j) vs[tid * M + j] = val[j];
    if (it.c0 == 0)
#pragma unroll
      for (int j = 0; j < M; ++j) a.z[(fd.z_off + p) * M + j] = val[j];
  }
  double acc[M];
#pragma unroll
  for (int j = 0; j < M; ++j) acc[j] = 0.0;
  const bool writer = rgid == 0 && cid < nc;
  if (writer && it.split == 0) gather_children<M>(a, fd, f, s + it.c0 + cid, acc);
  __syncthreads();
  if (nc <= 0) return;
#pragma unroll
  for (int k = 0; k < FWD_ROWS; ++k) {
    const int i = rgid + k * rg;
    if (i < nr) {
#pragma unroll
      for (int j = 0; j < M; ++j) acc[j] += rv[k] * vs[i * M + j];
    }
  }
  if (rg > 1) {
    if (rgid > 0)
#pragma unroll
      for (int j = 0; j < M; ++j) red[tid * M + j] = acc[j];
    __syncthreads();
    if (writer)
      for (int g = 1; g < rg; ++g)
#pragma unroll
        for (int j = 0; j < M; ++j) acc[j] += red[(g * cw + cid) * M + j];
  }
  if (writer) {
    const int32_t q = it.c0 + cid;
#pragma unroll
    for (int j = 0; j < M; ++j) a.z[(fd.z_off + s + int64_t(it.split) * fd.b + q) * M + j] = acc[j];
  }
}

// item = rows [r0,r1) of R = [F11^-1 | -G^T] (at most 32 rows, 4 per warp); the input vector of the whole front
// is staged once in dynamic shared memory, 8 x rows loads are in flight per lane
template <int M>
__global__ void __launch_bounds__(256) k_backward_large(SolveArgs a, const WorkItem *__restrict__ items) {
  extern __shared__ __align__(16) double ws[];
  const WorkItem it = items[blockIdx.x];
  const FrontDesc fd = a.fronts[it.node];
  const int32_t s = fd.s, f = fd.s + fd.b;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nrow = it.r1 - it.r0;
  const int per_warp = (nrow + 7) / 8;  // 1..4
  const int rbase = it.r0 + warp * per_warp;
  const double *Rr = a.R + fd.rb_off + int64_t(rbase) * fd.ldr;
  double acc[4][M];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int j = 0; j < M; ++j) acc[r][j] = 0.0;
  // first batch of matrix loads goes out before the input vector is gathered
  double rv[4][8];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int t = lane + 32 * u;
      rv[r][u] = (r < per_warp && rbase + r < it.r1 && t < f) ? __ldg(Rr + int64_t(r) * fd.ldr + t) : 0.0;
    }
  for (int p = tid; p < f; p += 256) {
    if (p < s) {
#pragma unroll
      for (int j = 0; j < M; ++j) ws[p * M + j] = a.z[(fd.z_off + p) * M + j];
    } else {
      const int64_t d = a.idx[fd.idx_off + p];
#pragma unroll
      for (int j = 0; j < M; ++j) ws[p * M + j] = a.x[d * M + j];
    }
  }
  __syncthreads();
  for (int t0 = 0; t0 < f; t0 += 256) {
    double nx_[4][8];
    const int t1 = t0 + 256;
    if (t1 < f) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int t = t1 + lane + 32 * u;
          nx_[r][u] = (r < per_warp && rbase + r < it.r1 && t < f) ? __ldg(Rr + int64_t(r) * fd.ldr + t) : 0.0;
        }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int t = t0 + lane + 32 * u;
      if (t < f) {
        double wv[M];
#pragma unroll
        for (int j = 0; j < M; ++j) wv[j] = ws[t * M + j];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int j = 0; j < M; ++j) acc[r][j] += rv[r][u] * wv[j];
      }
    }
    if (t1 < f) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int u = 0; u < 8; ++u) rv[r][u] = nx_[r][u];
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int j = 0; j < M; ++j) {
      double v = acc[r][j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      acc[r][j] = v;
    }
  if (lane == 0)
#pragma unroll
    for (int r = 0; r < 4; ++r)
      if (r < per_warp && rbase + r < it.r1) {
        const int64_t d = a.idx[fd.idx_off + rbase + r];
#pragma unroll
        for (int j = 0; j < M; ++j) a.x[d * M + j] = acc[r][j];
      }
}

// =================================================================================================================
// multifrontal solve, bottom of the tree: one CTA per subtree.  The subtree's factor blob is contiguous in HBM and
// is staged into shared memory by one TMA bulk copy (cp.async.bulk + mbarrier); all index structure comes from
// the shared template, so HBM traffic is the factor, the right-hand side and the solution.
// =================================================================================================================
template <int M>
__global__ void __launch_bounds__(256) k_subtree_forward(SubtreeArgs a) {
  extern __shared__ __align__(16) double sm[];
  const InstanceDev in = a.inst[blockIdx.x];
  const TemplateDev T = a.tmpl[in.tmpl];
  uint64_t *bar = reinterpret_cast<uint64_t *>(sm);
  double *blob = sm + 2;
  double *zl = blob + T.rf_size;
  double *vs = zl + int64_t(T.zl_size) * M;
  const int tid = threadIdx.x;
  if (tid == 0) mbar_init(bar);
  __syncthreads();
  if (tid == 0 && T.rf_size > 0) bulk_load_start(blob, a.R + in.rf_base, uint32_t(T.rf_size) * 8u, bar);
  for (int l = tid; l < T.n_int; l += 256) {
    const int64_t d = subtree_dof(a, in, T.var_code[l]);
#pragma unroll
    for (int j = 0; j < M; ++j) vs[l * M + j] = a.x[d * M + j];
  }
  __syncthreads();
  if (T.rf_size > 0) mbar_wait(bar, 0);
  for (int lv = 0; lv < T.n_levels; ++lv) {
    if (lv > 0) {  // leaves have no children
      for (int e = T.lvl_sep[lv] + tid; e < T.lvl_sep[lv + 1]; e += 256) {
        const int32_t c0 = T.sep_c0[e], c1 = T.sep_c1[e];
#pragma unroll
        for (int j = 0; j < M; ++j) {
          double v = vs[e * M + j];
          if (c0 >= 0) v += zl[c0 * M + j];
          if (c1 >= 0) v += zl[c1 * M + j];
          vs[e * M + j] = v;
        }
      }
      __syncthreads();
    }
    for (int e = T.lvl_out[lv] + tid; e < T.lvl_out[lv + 1]; e += 256) {
      const int32_t n = T.out_node[e];
      const int32_t s = T.node_s[n], b = T.node_b[n];
      const double *rf = blob + T.node_rf[n] + (e - T.node_zl[n]);
      const double *v = vs + int64_t(T.node_sep[n]) * M;
      double acc[M];
      const int32_t c0 = T.out_c0[e], c1 = T.out_c1[e];
#pragma unroll
      for (int j = 0; j < M; ++j) {
        acc[j] = 0.0;
        if (c0 >= 0) acc[j] += zl[c0 * M + j];
        if (c1 >= 0) acc[j] += zl[c1 * M + j];
      }
      for (int i = 0; i < s; ++i) {
        const double r = rf[i * b];
#pragma unroll
        for (int j = 0; j < M; ++j) acc[j] += r * v[i * M + j];
      }
#pragma unroll
      for (int j = 0; j < M; ++j) zl[e * M + j] = acc[j];
    }
    __syncthreads();
  }
  for (int l = tid; l < T.n_int; l += 256) {
    const int64_t d = subtree_dof(a, in, T.var_code[l]);
#pragma unroll
    for (int j = 0; j < M; ++j) a.x[d * M + j] = vs[l * M + j];
  }
  const int32_t zr = T.node_zl[T.n_nodes - 1];
  for (int q = tid; q < T.n_rootb; q += 256)
#pragma unroll
    for (int j = 0; j < M; ++j) a.z[(in.z_root + q) * M + j] = zl[(zr + q) * M + j];
}

template <int M>
__global__ void __launch_bounds__(256) k_subtree_backward(SubtreeArgs a) {
  extern __shared__ __align__(16) double sm[];
  const InstanceDev in = a.inst[blockIdx.x];
  const TemplateDev T = a.tmpl[in.tmpl];
  uint64_t *bar = reinterpret_cast<uint64_t *>(sm);
  double *blob = sm + 2;
  double *zs = blob + T.rb_size;             // accumulated right-hand sides of the interior variables
  double *xs = zs + int64_t(T.n_int) * M;    // solution: [interior | root boundary]
  const int tid = threadIdx.x;
  if (tid == 0) mbar_init(bar);
  __syncthreads();
  if (tid == 0) bulk_load_start(blob, a.R + in.rb_base, uint32_t(T.rb_size) * 8u, bar);
  for (int l = tid; l < T.n_int; l += 256) {
    const int64_t d = subtree_dof(a, in, T.var_code[l]);
#pragma unroll
    for (int j = 0; j < M; ++j) zs[l * M + j] = a.x[d * M + j];
  }
  for (int q = tid; q < T.n_rootb; q += 256) {
    const int64_t d = subtree_dof(a, in, T.rootb_code[q]);
#pragma unroll
    for (int j = 0; j < M; ++j) xs[(T.n_int + q) * M + j] = a.x[d * M + j];
  }
  __syncthreads();
  mbar_wait(bar, 0);
  for (int lv = T.n_levels - 1; lv >= 0; --lv) {
    for (int e = T.lvl_sep[lv] + tid; e < T.lvl_sep[lv + 1]; e += 256) {
      const int32_t n = T.sep_node[e];
      const int32_t s = T.node_s[n], b = T.node_b[n], sep0 = T.node_sep[n];
      const double *rb = blob + T.node_rb[n] + (e - sep0);  // R^T[p][i] at p * s + i
      const int32_t *bv = T.bnd_var + T.node_bnd[n];
      double acc[M];
#pragma unroll
      for (int j = 0; j < M; ++j) acc[j] = 0.0;
      for (int p = 0; p < s; ++p) {
        const double r = rb[p * s];
#pragma unroll
        for (int j = 0; j < M; ++j) acc[j] += r * zs[(sep0 + p) * M + j];
      }
      rb += s * s;
      for (int q = 0; q < b; ++q) {
        const double r = rb[q * s];
        const int32_t l = bv[q];
#pragma unroll
        for (int j = 0; j < M; ++j) acc[j] += r * xs[l * M + j];
      }
#pragma unroll
      for (int j = 0; j < M; ++j) xs[e * M + j] = acc[j];
    }
    __syncthreads();
  }
  for (int l = tid; l < T.n_int; l += 256) {
    const int64_t d = subtree_dof(a, in, T.var_code[l]);
#pragma unroll
    for (int j = 0; j < M; ++j) a.x[d * M + j] = xs[l * M + j];
  }
}

#define MMMFE_DISPATCH_M(M, CALL) \
  switch (M) {                    \
    case 1: { constexpr int MM = 1; CALL; } break; \
    case 2: { constexpr int MM = 2; CALL; } break; \
    case 4: { constexpr int MM = 4; CALL; } break; \
    default: break;               \
  }

void launch_forward_large(int M, const SolveArgs &a, const WorkItem *items, int32_t n_items, cudaStream_t st) {
  if (n_items <= 0) return;
  MMMFE_DISPATCH_M(M, (k_forward_large<MM><<<n_items, 256, 0, st>>>(a, items)));
  launch_counter_add(1);
}
void launch_backward_large(int M, const SolveArgs &a, const WorkItem *items, int32_t n_items, int32_t f_max,
                           cudaStream_t st) {
  if (n_items <= 0) return;
  const size_t smem = size_t(f_max) * M * sizeof(double);
  MMMFE_DISPATCH_M(M, (k_backward_large<MM><<<n_items, 256, smem, st>>>(a, items)));
  launch_counter_add(1);
}

size_t subtree_smem_limit() {
  int dev = 0, optin = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  const size_t lim = size_t(optin);
  cudaFuncSetAttribute(k_backward_large<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(lim));
  cudaFuncSetAttribute(k_backward_large<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(lim));
  cudaFuncSetAttribute(k_backward_large<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(lim));
  cudaFuncSetAttribute(k_subtree_forward<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(lim));
  cudaFuncSetAttribute(k_subtree_forward<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(lim));
  cudaFuncSetAttribute(k_subtree_forward<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(lim));
  cudaFuncSetAttribute(k_subtree_backward<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(lim));
  cudaFuncSetAttribute(k_subtree_backward<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(lim));
  cudaFuncSetAttribute(k_subtree_backward<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(lim));
  return lim;
}

void launch_subtree_forward(int M, const SubtreeArgs &a, int32_t n_inst, size_t smem_bytes, cudaStream_t st) {
  if (n_inst <= 0) return;
  MMMFE_DISPATCH_M(M, (k_subtree_forward<MM><<<n_inst, 256, smem_bytes, st>>>(a)));
  launch_counter_add(1);
}
void launch_subtree_backward(int M, const SubtreeArgs &a, int32_t n_inst, size_t smem_bytes, cudaStream_t st) {
  if (n_inst <= 0) return;
  MMMFE_DISPATCH_M(M, (k_subtree_backward<MM><<<n_inst, 256, smem_bytes, st>>>(a)));
  launch_counter_add(1);
}

// =================================================================================================================
// time stepping
// =================================================================================================================
__global__ void __launch_bounds__(256) k_build_rhs(StepArgs a, const double *__restrict__ lam_bar, int32_t level) {
  const int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (e >= a.n_flux) return;
  if (a.level_dev) level = *a.level_dev;
  const int64_t k0 = a.kup_rp[e], k1 = a.kup_rp[e + 1];
  for (int j = 0; j < a.M; ++j) {
    double acc = 0.0;
    for (int64_t k = k0; k < k1; ++k) acc -= a.kup_val[k] * a.ptil[int64_t(a.kup_col[k]) * a.M + j];
    if (lam_bar) {
      const int32_t q = a.iface_q[e * a.M + j];
      if (q >= 0) acc -= a.tr_sign[q] * lam_bar[a.tr_base[q] + int64_t(level) * a.tr_stride[q]];
    }
    a.x[e * a.M + j] = acc;
  }
}

void launch_build_rhs(const StepArgs &a, const double *lam_bar, int32_t level, cudaStream_t st) {
  k_build_rhs<<<unsigned((a.n_flux + 255) / 256), 256, 0, st>>>(a, lam_bar, level);
  launch_counter_add(1);
}

__global__ void __launch_bounds__(256) k_add_sparse(double *x, int32_t M, int32_t member,
                                                    const int32_t *__restrict__ dofs,
                                                    const double *__restrict__ vals, int32_t n,
                                                    const int32_t *__restrict__ level_dev) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (level_dev) vals += int64_t(*level_dev) * n;
  x[int64_t(dofs[i]) * M + member] += vals[i];
}

void launch_add_sparse(double *x, int32_t M, int32_t member, const int32_t *dofs, const double *vals, int32_t n,
                       cudaStream_t st, const int32_t *level_dev) {
  if (n <= 0) return;
  k_add_sparse<<<(n + 255) / 256, 256, 0, st>>>(x, M, member, dofs, vals, n, level_dev);
  launch_counter_add(1);
}

__global__ void __launch_bounds__(256)
k_update(StepArgs a, double *__restrict__ trace, int32_t level, const double *const *__restrict__ src_next,
         int64_t src_stride, double *const *__restrict__ hist, int64_t hist_level_stride, int add_hist, int32_t n_levels) {
  int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (a.level_dev) level = *a.level_dev;
  if (n_levels > 0 && level + 1 >= n_levels) src_next = nullptr;
  if (t < a.n_pressure) {
    const int64_t c = t;
    const int64_t k0 = a.kpu_rp[c], k1 = a.kpu_rp[c + 1];
    const double mi = a.minv[c];
    for (int j = 0; j < a.M; ++j) {
      double acc = 0.0;
      for (int64_t k = k0; k < k1; ++k) acc += a.kpu_val[k] * a.x[int64_t(a.kpu_col[k]) * a.M + j];
      const double p = a.ptil[c * a.M + j] - mi * acc;
      if (hist && hist[j]) {
        double *h = hist[j] + int64_t(level) * hist_level_stride + a.n_flux + c;
        *h = add_hist ? *h + p : p;
      }
      double nxt = p;
      if (src_next && src_next[j]) nxt += mi * src_next[j][int64_t(level + 1) * src_stride + c];
      a.ptil[c * a.M + j] = nxt;
    }
    return;
  }
  t -= a.n_pressure;
  if (t < a.n_tr) {
    trace[a.tr_base[t] + int64_t(level) * a.tr_stride[t]] = a.x[int64_t(a.tr_dof[t]) * a.M + a.tr_mem[t]];
    return;
  }
  t -= a.n_tr;
  if (hist && t < a.n_flux) {
    for (int j = 0; j < a.M; ++j)
      if (hist[j]) {
        double *h = hist[j] + int64_t(level) * hist_level_stride + t;
        const double v = a.x[t * a.M + j];
        *h = add_hist ? *h + v : v;
      }
  }
}

void launch_update(const StepArgs &a, double *trace, int32_t level, const double *const *src_next,
                   int64_t src_stride, double *const *hist, int64_t hist_level_stride, int add_hist,
                   cudaStream_t st, int32_t n_levels) {
  const int64_t n = a.n_pressure + a.n_tr + (hist ? a.n_flux : 0);
  k_update<<<unsigned((n + 255) / 256), 256, 0, st>>>(a, trace, level, src_next, src_stride, hist,
                                                      hist_level_stride, add_hist, n_levels);
  launch_counter_add(1);
}

__global__ void __launch_bounds__(256) k_init_ptil(StepArgs a, const double *const *__restrict__ p0,
                                                   const double *const *__restrict__ src0) {
  const int64_t c = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (c >= a.n_pressure) return;
  for (int j = 0; j < a.M; ++j) {
    double v = 0.0;
    if (p0 && p0[j]) v = p0[j][c];
    if (src0 && src0[j]) v += a.minv[c] * src0[j][c];
    a.ptil[c * a.M + j] = v;
  }
}

// the same for the persistent kernel's pressures: [box-major cell][M]
__global__ void __launch_bounds__(256) k_init_ptil_box(int64_t n_cells, int32_t M, const int32_t *__restrict__ cell_box,
                                                       const int32_t *__restrict__ p_of_cell, const double *__restrict__ minv_box,
                                                       double minv_const, const double *const *__restrict__ p0,
                                                       const double *const *__restrict__ src0, double *__restrict__ ptil) {
  const int64_t c = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;  // canonical cell
  if (c >= n_cells) return;
  const int64_t pos = cell_box[c], p = p_of_cell ? int64_t(p_of_cell[c]) : c;
  const double mi = minv_box ? minv_box[pos] : minv_const;
  for (int j = 0; j < M; ++j) {
    double v = 0.0;
    if (p0 && p0[j]) v = p0[j][p];
    if (src0 && src0[j]) v += mi * src0[j][p];
    ptil[pos * M + j] = v;
  }
}

void launch_init_ptil_box(int64_t n_cells, int32_t M, const int32_t *cell_box, const int32_t *p_of_cell, const double *minv_box,
                          double minv_const, const double *const *p0, const double *const *src0, double *ptil, cudaStream_t st) {
  k_init_ptil_box<<<unsigned((n_cells + 255) / 256), 256, 0, st>>>(n_cells, M, cell_box, p_of_cell, minv_box, minv_const, p0, src0, ptil);
  launch_counter_add(1);
}

void launch_init_ptil(const StepArgs &a, const double *const *p0, const double *const *src0, cudaStream_t st) {
  k_init_ptil<<<unsigned((a.n_pressure + 255) / 256), 256, 0, st>>>(a, p0, src0);
  launch_counter_add(1);
}

// =================================================================================================================
// interface vector kernels
// =================================================================================================================
__global__ void __launch_bounds__(256) k_spmv(int64_t n_rows, const int64_t *__restrict__ rp,
                                              const int32_t *__restrict__ col, const double *__restrict__ val,
                                              const double *__restrict__ x, double *__restrict__ y) {
  const int64_t r = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (r >= n_rows) return;
  double acc = 0.0;
  for (int64_t k = rp[r]; k < rp[r + 1]; ++k) acc += val[k] * x[col[k]];
  y[r] = acc;
}

void launch_spmv(int64_t n_rows, const int64_t *rp, const int32_t *col, const double *val, const double *x,
                 double *y, cudaStream_t st) {
  if (n_rows <= 0) return;
  k_spmv<<<unsigned((n_rows + 255) / 256), 256, 0, st>>>(n_rows, rp, col, val, x, y);
  launch_counter_add(1);
}

__global__ void __launch_bounds__(256) k_exchange_combine(int64_t n, const double *__restrict__ send,
                                                          const int64_t *__restrict__ partner,
                                                          const double *__restrict__ recv, double sign,
                                                          double *__restrict__ out) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const int64_t p = partner[i];
  const double other = (p >= 0) ? send[p] : recv[-p - 2];
  out[i] = sign * (send[i] + other);
}

void launch_exchange_combine(int64_t n, const double *send, const int64_t *partner, const double *recv, double sign,
                             double *out, cudaStream_t st) {
  if (n <= 0) return;
  k_exchange_combine<<<unsigned((n + 255) / 256), 256, 0, st>>>(n, send, partner, recv, sign, out);
  launch_counter_add(1);
}

__global__ void __launch_bounds__(256) k_gather(int64_t n, const double *__restrict__ srcv,
                                                const int64_t *__restrict__ index, double *__restrict__ dst) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i < n) dst[i] = srcv[index[i]];
}

void launch_gather(int64_t n, const double *srcv, const int64_t *index, double *dst, cudaStream_t st) {
  if (n <= 0) return;
  k_gather<<<unsigned((n + 255) / 256), 256, 0, st>>>(n, srcv, index, dst);
  launch_counter_add(1);
}

__device__ __forceinline__ double block_sum_512(double v, double *sm) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) sm[warp] = v;
  __syncthreads();
  double r = 0.0;
  if (warp == 0) {
    r = (lane < (blockDim.x >> 5)) ? sm[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  return r;  // valid in thread 0
}

// ---- Arnoldi / classical Gram-Schmidt on the interface vector (src/darcy_vtdd.cc:1413-1463) ------------------------
// The interface vector is cut into segments of ARN_SEG entries; a block keeps its segment of w in registers and
// walks a chunk of the Krylov vectors, so w is read once per block and the grid has (segments x vector chunks)
// blocks however few vectors there are.  Partial sums go to scratch[vector][segment]; the LAST block to finish
// (ticket counter) adds them in segment order: one launch, no atomics on the data, bitwise reproducible.
constexpr int ARN_THREADS = 256;
constexpr int ARN_PER = 4;                       // entries per thread
constexpr int ARN_SEG = ARN_THREADS * ARN_PER;   // 1024

__device__ __forceinline__ double block_sum_256(double v, double *sm) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();  // sm may still be read by the previous call
  if (lane == 0) sm[warp] = v;
  __syncthreads();
  double r = 0.0;
  if (warp == 0) {
    r = (lane < (ARN_THREADS >> 5)) ? sm[lane] : 0.0;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  return r;  // valid in thread 0
}

// true in every thread of the last block of the grid to get here (all other blocks' global writes are visible)
__device__ __forceinline__ bool last_block(unsigned int *ticket) {
  __shared__ unsigned int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int total = gridDim.x * gridDim.y;
    const unsigned int t = atomicAdd(ticket, 1u);
    s_last = (t == total - 1) ? 1u : 0u;
    if (s_last) *ticket = 0u;  // ready for the next launch (stream-ordered)
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last != 0u;
}

__global__ void __launch_bounds__(ARN_THREADS) k_multi_dot(int64_t len, const double *__restrict__ w,
                                                           const double *__restrict__ Q, int64_t ldq, int32_t n_vec,
                                                           int32_t vec_per_block, double *__restrict__ scratch,
                                                           unsigned int *ticket, double *__restrict__ h) {
  __shared__ double sm[8];
  const int n_seg = gridDim.x;
  const int64_t j0 = int64_t(blockIdx.x) * ARN_SEG + threadIdx.x;
  double wv[ARN_PER];
#pragma unroll
  for (int u = 0; u < ARN_PER; ++u) {
    const int64_t j = j0 + u * ARN_THREADS;
    wv[u] = j < len ? w[j] : 0.0;
  }
  const int v0 = blockIdx.y * vec_per_block, v1 = min(n_vec, v0 + vec_per_block);
  for (int i = v0; i < v1; ++i) {
    const double *q = Q + int64_t(i) * ldq;
    double acc = 0.0;
#pragma unroll
    for (int u = 0; u < ARN_PER; ++u) {
      const int64_t j = j0 + u * ARN_THREADS;
      if (j < len) acc += wv[u] * q[j];
    }
    const double r = block_sum_256(acc, sm);
    if (threadIdx.x == 0) scratch[int64_t(i) * n_seg + blockIdx.x] = r;
  }
  if (!last_block(ticket)) return;
  // one warp per vector: lane l adds the segments l, l + 32, ... in order, then the lanes are combined by a fixed
  // shuffle tree -- the same order on every launch
  for (int i = threadIdx.x >> 5; i < n_vec; i += ARN_THREADS / 32) {
    const double *p = scratch + int64_t(i) * n_seg;
    double t = 0.0;
    for (int sg = threadIdx.x & 31; sg < n_seg; sg += 32) t += __ldcg(p + sg);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if ((threadIdx.x & 31) == 0) h[i] = t;
  }
}

static int arnoldi_segments(int64_t len) { return int((len + ARN_SEG - 1) / ARN_SEG); }

size_t arnoldi_scratch_doubles(int64_t len, int32_t max_vectors) {
  return size_t(arnoldi_segments(len)) * size_t(max_vectors) + 16;
}

// h[i] = <w, Q_i>, i < k_plus_1.  scratch: arnoldi_scratch_doubles(len, k_plus_1); ticket: one zeroed counter.
void launch_multi_dot(int64_t len, const double *w, const double *Q, int64_t ldq, int32_t k_plus_1, double *h,
                      double *scratch, unsigned int *ticket, cudaStream_t st) {
  const int n_seg = arnoldi_segments(len);
  // enough blocks for four per SM when the vectors allow it
  int chunks = std::max(1, std::min(k_plus_1, (592 + n_seg - 1) / n_seg));
  const int per = (k_plus_1 + chunks - 1) / chunks;
  chunks = (k_plus_1 + per - 1) / per;
  k_multi_dot<<<dim3(n_seg, chunks), ARN_THREADS, 0, st>>>(len, w, Q, ldq, k_plus_1, per, scratch, ticket, h);
  launch_counter_add(1);
}

// w -= sum_i h[i] Q_i; norm2[0] = ||w||^2 (partials per segment, added in segment order by the last block)
__global__ void __launch_bounds__(ARN_THREADS) k_cgs_update(int64_t len, double *__restrict__ w,
                                                            const double *__restrict__ Q, int64_t ldq, int32_t k_plus_1,
                                                            const double *__restrict__ h, double *__restrict__ scratch,
                                                            unsigned int *ticket, double *__restrict__ norm2) {
  __shared__ double sm[8];
  const int64_t j0 = int64_t(blockIdx.x) * ARN_SEG + threadIdx.x;
  double v[ARN_PER];
#pragma unroll
  for (int u = 0; u < ARN_PER; ++u) {
    const int64_t j = j0 + u * ARN_THREADS;
    v[u] = j < len ? w[j] : 0.0;
  }
#pragma unroll 2
  for (int i = 0; i < k_plus_1; ++i) {
    const double hi = h[i];
    const double *q = Q + int64_t(i) * ldq;
#pragma unroll
    for (int u = 0; u < ARN_PER; ++u) {
      const int64_t j = j0 + u * ARN_THREADS;
      if (j < len) v[u] -= hi * q[j];
    }
  }
  double acc2 = 0.0;
#pragma unroll
  for (int u = 0; u < ARN_PER; ++u) {
    const int64_t j = j0 + u * ARN_THREADS;
    if (j < len) { w[j] = v[u]; acc2 += v[u] * v[u]; }
  }
  const double r = block_sum_256(acc2, sm);
  if (threadIdx.x == 0) scratch[blockIdx.x] = r;
  if (!last_block(ticket)) return;
  double t = 0.0;
  for (int sg = threadIdx.x; sg < int(gridDim.x); sg += ARN_THREADS) t += __ldcg(scratch + sg);
  const double tot = block_sum_256(t, sm);  // fixed tree
  if (threadIdx.x == 0) norm2[0] = tot;
}

void launch_cgs_update(int64_t len, double *w, const double *Q, int64_t ldq, int32_t k_plus_1, const double *h,
                       double *norm2, double *scratch, unsigned int *ticket, cudaStream_t st) {
  k_cgs_update<<<arnoldi_segments(len), ARN_THREADS, 0, st>>>(len, w, Q, ldq, k_plus_1, h, scratch, ticket, norm2);
  launch_counter_add(1);
}

__global__ void __launch_bounds__(256) k_scale_store(int64_t len, const double *__restrict__ w, double inv,
                                                     const double *__restrict__ norm2_dev, double *__restrict__ dst) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (norm2_dev) inv = sqrt(norm2_dev[0]);  // IEEE square root: the value the host computes from the same number
  if (i < len) dst[i] = w[i] / inv;
}

// dst = w / inv  (the reference divides: q[i] /= h[k+1], src/darcy_vtdd.cc:1461); norm2_dev != nullptr: inv =
// sqrt(*norm2_dev), read on the device (no host round trip between the norm and the next Krylov vector)
void launch_scale_store(int64_t len, const double *w, double inv, double *dst, cudaStream_t st, const double *norm2_dev) {
  if (len <= 0) return;
  k_scale_store<<<unsigned((len + 255) / 256), 256, 0, st>>>(len, w, inv, norm2_dev, dst);
  launch_counter_add(1);
}

__global__ void __launch_bounds__(256) k_axpy_rows(int64_t len, const double *__restrict__ Q, int64_t ldq,
                                                   int32_t n_rows, const double *__restrict__ y,
                                                   double *__restrict__ out) {
  const int64_t j = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (j >= len) return;
  double acc = 0.0;
  for (int i = 0; i < n_rows; ++i) acc += Q[int64_t(i) * ldq + j] * y[i];
  out[j] = acc;
}

// out = sum_i y[i] Q_i  (lambda reconstruction, src/darcy_vtdd.cc:1517-1521)
void launch_axpy_rows(int64_t len, const double *Q, int64_t ldq, int32_t n_rows, const double *y, double *out,
                      cudaStream_t st) {
  if (len <= 0) return;
  k_axpy_rows<<<unsigned((len + 255) / 256), 256, 0, st>>>(len, Q, ldq, n_rows, y, out);
  launch_counter_add(1);
}

// =================================================================================================================
// diagnostics of the set-up cross-check: reproducible pseudo-random data and ||a - b||^2, ||b||^2
// =================================================================================================================
__global__ void __launch_bounds__(256) k_fill_hash(int64_t n, uint64_t seed, double *__restrict__ out) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  uint64_t x = uint64_t(i) + seed * 0x9e3779b97f4a7c15ull;  // splitmix64
  x += 0x9e3779b97f4a7c15ull; x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull; x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
  x ^= x >> 31;
  out[i] = double(x >> 11) * (2.0 / 9007199254740992.0) - 1.0;  // uniform in [-1, 1)
}

void launch_fill_hash(int64_t n, uint64_t seed, double *out, cudaStream_t st) {
  if (n <= 0) return;
  k_fill_hash<<<unsigned((n + 255) / 256), 256, 0, st>>>(n, seed, out);
  launch_counter_add(1);
}

__global__ void __launch_bounds__(512) k_diff_norms(int64_t n, const double *__restrict__ a, const double *__restrict__ b,
                                                    double *__restrict__ out) {
  __shared__ double sm[16];
  double d2 = 0.0, b2 = 0.0;
  for (int64_t j = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; j < n; j += int64_t(gridDim.x) * blockDim.x) {
    const double d = a[j] - b[j];
    d2 += d * d;
    b2 += b[j] * b[j];
  }
  const double r0 = block_sum_512(d2, sm);
  __syncthreads();
  const double r1 = block_sum_512(b2, sm);
  if (threadIdx.x == 0) { atomicAdd(out, r0); atomicAdd(out + 1, r1); }
}

// out[0] += ||a - b||^2, out[1] += ||b||^2 (out must be zeroed by the caller)
void launch_diff_norms(int64_t n, const double *a, const double *b, double *out, cudaStream_t st) {
  if (n <= 0) return;
  k_diff_norms<<<unsigned(std::min<int64_t>(592, (n + 511) / 512)), 512, 0, st>>>(n, a, b, out);
  launch_counter_add(1);
}

// =================================================================================================================
// measured FP64 tensor-core (DMMA) peak: the denominator of the factorisation / multiscale-basis rooflines
// =================================================================================================================
__global__ void __launch_bounds__(256) k_dmma_peak(double *out, int iters) {
  double acc[8][2];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = 0.0;
  const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
  for (int i = 0; i < iters; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) dmma_m8n8k4(acc[j][0], acc[j][1], a, b);  // 8 independent accumulators per warp
  double sum = 0.0;
#pragma unroll
  for (int j = 0; j < 8; ++j) sum += acc[j][0] + acc[j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
}

// TFLOP/s of register-resident mma.sync.m8n8k4.f64 on all SMs (8 warps x 4 CTAs per SM), best of 3
double measure_dmma_tflops(cudaStream_t st) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int blocks = sms * 4, iters = 20000;
  double *out = nullptr;
  if (cudaMalloc(&out, size_t(blocks) * 256 * sizeof(double)) != cudaSuccess) return 0.0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0, st);
    k_dmma_peak<<<blocks, 256, 0, st>>>(out, iters);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = double(blocks) * 8.0 * iters * 8.0 * 512.0;  // m8n8k4: 256 FMA = 512 flop
    if (rep > 0 && ms > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  launch_counter_add(4);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(out);
  return best;
}

}  // namespace mmmfe
