// Multi-column solve / time-stepping kernels and the block-Toeplitz operator kernel (sm_100a).  See wide.cuh.
#include "wide.cuh"

#include <algorithm>

#include "dev_common.cuh"

namespace mmmfe {

// =================================================================================================================
// time stepping, MW columns: thread = (row, column), columns fastest (coalesced rows of MW doubles)
// =================================================================================================================
// one warp per row, lanes over the columns: the sparse row is read once per warp, the vector rows in full lines
__global__ void __launch_bounds__(256) k_wide_rhs(WideStep a) {
  const int64_t e = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (e >= a.n_flux) return;
  const int lane = threadIdx.x & 31;
  const int64_t k0 = a.kup_rp[e], k1 = a.kup_rp[e + 1];
  for (int j = lane; j < a.MW; j += 32) {
    double acc = 0.0;
    for (int64_t k = k0; k < k1; ++k) acc -= a.kup_val[k] * a.ptil[int64_t(a.kup_col[k]) * a.MW + j];
    a.x[e * a.MW + j] = acc;
  }
}

void launch_wide_rhs(const WideStep &a, cudaStream_t st) {
  k_wide_rhs<<<unsigned((a.n_flux + 7) / 8), 256, 0, st>>>(a);
  launch_counter_add(1);
}

__global__ void __launch_bounds__(256) k_wide_add(WideStep a, const int64_t *__restrict__ lvl_rp,
                                                  const int32_t *__restrict__ row, const int32_t *__restrict__ col,
                                                  const double *__restrict__ val, int32_t n_lvl) {
  const int32_t level = a.state[0], pass = a.state[1];
  if (level >= n_lvl) return;
  const int64_t *rp = lvl_rp + int64_t(pass) * (n_lvl + 1) + level;
  const int64_t e = rp[0] + blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (e >= rp[1]) return;
  a.x[int64_t(row[e]) * a.MW + col[e]] += val[e];  // (row, column) pairs are unique within a level
}

void launch_wide_add(const WideStep &a, const int64_t *lvl_rp, const int32_t *row, const int32_t *col, const double *val,
                     int32_t n_lvl, int64_t max_entries, cudaStream_t st) {
  if (max_entries <= 0) return;
  k_wide_add<<<unsigned((max_entries + 255) / 256), 256, 0, st>>>(a, lvl_rp, row, col, val, n_lvl);
  launch_counter_add(1);
}

__global__ void __launch_bounds__(256) k_wide_update(WideStep a, const int32_t *__restrict__ tr_dof, int32_t n_tr,
                                                     double *__restrict__ trace) {
  const int64_t r = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r < a.n_pressure) {
    const int64_t k0 = a.kpu_rp[r], k1 = a.kpu_rp[r + 1];
    const double mi = a.minv[r];
    for (int j = lane; j < a.MW; j += 32) {
      double acc = 0.0;
      for (int64_t k = k0; k < k1; ++k) acc += a.kpu_val[k] * a.x[int64_t(a.kpu_col[k]) * a.MW + j];
      a.ptil[r * a.MW + j] -= mi * acc;
    }
    return;
  }
  const int64_t t = r - a.n_pressure;
  if (t < n_tr) {
    const double *src = a.x + int64_t(tr_dof[t]) * a.MW;
    double *dst = trace + (int64_t(a.state[0]) * n_tr + t) * a.MW;
    for (int j = lane; j < a.MW; j += 32) dst[j] = src[j];
  }
}

void launch_wide_update(const WideStep &a, const int32_t *tr_dof, int32_t n_tr, double *trace, cudaStream_t st) {
  const int64_t rows = a.n_pressure + n_tr;
  k_wide_update<<<unsigned((rows + 7) / 8), 256, 0, st>>>(a, tr_dof, n_tr, trace);
  launch_counter_add(1);
}

// max |v| over a vector as the bit pattern of a non-negative double (monotone): out must be zeroed by the caller
__global__ void __launch_bounds__(256) k_absmax(int64_t n, const double *__restrict__ v, unsigned long long *out) {
  double m = 0.0;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) m = fmax(m, fabs(v[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.0) atomicMax(out, (unsigned long long)__double_as_longlong(m));
}

void launch_absmax(int64_t n, const double *v, unsigned long long *out, cudaStream_t st) {
  if (n <= 0) return;
  k_absmax<<<unsigned(std::min<int64_t>(1184, (n + 255) / 256)), 256, 0, st>>>(n, v, out);
  launch_counter_add(1);
}

__global__ void k_wide_next_level(int32_t *state) { state[0] += 1; }

void launch_wide_next_level(int32_t *state, cudaStream_t st) {
  k_wide_next_level<<<1, 1, 0, st>>>(state);
  launch_counter_add(1);
}

// =================================================================================================================
// top of the tree: gathers around the per-level grouped GEMMs
// =================================================================================================================
__global__ void __launch_bounds__(256) k_wide_fwd_gather(SolveArgs a, int32_t MW, const WideRow *__restrict__ rows,
                                                         int64_t n_rows) {
  const int64_t w = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (w >= n_rows) return;
  const int lane = threadIdx.x & 31;
  const WideRow rw = rows[w];
  const FrontDesc fd = a.fronts[rw.node];
  const int32_t f = fd.s + fd.b, p = rw.p;
  const double *xr = p < fd.s ? a.x + int64_t(a.idx[fd.idx_off + p]) * MW : nullptr;
  const double *c0 = nullptr, *c1 = nullptr;
  if (fd.child0 >= 0) {
    const int32_t q = a.src[fd.src_off + p];
    if (q >= 0) { const FrontDesc cd = a.fronts[fd.child0]; c0 = a.z + (cd.z_off + cd.s + q) * MW; }
  }
  if (fd.child1 >= 0) {
    const int32_t q = a.src[fd.src_off + f + p];
    if (q >= 0) { const FrontDesc cd = a.fronts[fd.child1]; c1 = a.z + (cd.z_off + cd.s + q) * MW; }
  }
  double *dst = a.z + (fd.z_off + p) * MW;
  for (int j = lane; j < MW; j += 32) {
    double v = xr ? xr[j] : 0.0;
    if (c0) v += c0[j];
    if (c1) v += c1[j];
    dst[j] = v;
  }
}

void launch_wide_fwd_gather(const SolveArgs &a, int32_t MW, const WideRow *rows, int64_t n_rows, cudaStream_t st) {
  if (n_rows <= 0) return;
  k_wide_fwd_gather<<<unsigned((n_rows + 7) / 8), 256, 0, st>>>(a, MW, rows, n_rows);
  launch_counter_add(1);
}

__global__ void __launch_bounds__(256) k_wide_bwd_gather(SolveArgs a, int32_t MW, const WideRow *__restrict__ rows,
                                                         int64_t n_rows) {
  const int64_t w = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (w >= n_rows) return;
  const int lane = threadIdx.x & 31;
  const WideRow rw = rows[w];
  const FrontDesc fd = a.fronts[rw.node];
  const int32_t p = fd.s + rw.p;
  const double *src = a.x + int64_t(a.idx[fd.idx_off + p]) * MW;
  double *dst = a.z + (fd.z_off + p) * MW;
  for (int j = lane; j < MW; j += 32) dst[j] = src[j];
}

void launch_wide_bwd_gather(const SolveArgs &a, int32_t MW, const WideRow *rows, int64_t n_rows, cudaStream_t st) {
  if (n_rows <= 0) return;
  k_wide_bwd_gather<<<unsigned((n_rows + 7) / 8), 256, 0, st>>>(a, MW, rows, n_rows);
  launch_counter_add(1);
}

// =================================================================================================================
// bottom of the tree: one CTA per (subtree, column group).  The subtree's factor blob is staged once by a TMA bulk copy
// and reused for every column chunk of the CTA; a chunk is MC = 8, 16 or 32 columns wide (as wide as the shared memory
// allows: the walk is bound by the latency of its levels, not by arithmetic, so wider chunks are almost free).
// Thread = (row slot, group of 8 columns), row slots fastest: rows of a level are dealt to the row slots, every thread
// keeps 8 column sums in registers; vectors are stored [row][MC + 2] in shared memory (the padding spreads the 16-byte
// accesses of 8 consecutive rows over all banks), global rows are moved as 64-byte segments.
// =================================================================================================================
template <int MC>
__global__ void __launch_bounds__(256) k_wide_subtree_forward(SubtreeArgs a, int32_t MW) {
  constexpr int CG = MC / 8, RT = 256 / CG, LD = MC + 2;
  extern __shared__ __align__(16) double sm[];
  const InstanceDev in = a.inst[blockIdx.x];
  const TemplateDev T = a.tmpl[in.tmpl];
  uint64_t *bar = reinterpret_cast<uint64_t *>(sm);
  double *blob = sm + 2;
  double *zl = blob + T.rf_size;
  double *vs = zl + int64_t(T.zl_size) * LD;
  const int tid = threadIdx.x, rt = tid % RT, cg = tid / RT, cj = cg * 8;
  if (tid == 0) mbar_init(bar);
  __syncthreads();
  if (tid == 0 && T.rf_size > 0) bulk_load_start(blob, a.R + in.rf_base, uint32_t(T.rf_size) * 8u, bar);
  bool waited = T.rf_size <= 0;
  const int32_t zr = T.node_zl[T.n_nodes - 1];
  for (int32_t c0 = blockIdx.y * MC; c0 < MW; c0 += gridDim.y * MC) {  // this CTA's column chunks
    const bool live = c0 + cj < MW;  // MW is a multiple of 8: a group of 8 columns is inside or outside as a whole
    for (int l = rt; l < T.n_int; l += RT) {
      const int64_t d = subtree_dof(a, in, T.var_code[l]);
      const double2 *src = reinterpret_cast<const double2 *>(a.x + d * MW + c0 + cj);
      double2 *dst = reinterpret_cast<double2 *>(vs + l * LD + cj);
#pragma unroll
      for (int h = 0; h < 4; ++h) dst[h] = live ? src[h] : make_double2(0.0, 0.0);
    }
    __syncthreads();
    if (!waited) { mbar_wait(bar, 0); waited = true; }
    for (int lv = 0; lv < T.n_levels; ++lv) {
      if (lv > 0) {
        for (int e = T.lvl_sep[lv] + rt; e < T.lvl_sep[lv + 1]; e += RT) {
          const int32_t s0 = T.sep_c0[e], s1 = T.sep_c1[e];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            double v = vs[e * LD + cj + j];
            if (s0 >= 0) v += zl[s0 * LD + cj + j];
            if (s1 >= 0) v += zl[s1 * LD + cj + j];
            vs[e * LD + cj + j] = v;
          }
        }
        __syncthreads();
      }
      for (int e = T.lvl_out[lv] + rt; e < T.lvl_out[lv + 1]; e += RT) {
        const int32_t n = T.out_node[e];
        const int32_t s = T.node_s[n], b = T.node_b[n];
        const double *rf = blob + T.node_rf[n] + (e - T.node_zl[n]);
        const double *v = vs + int64_t(T.node_sep[n]) * LD + cj;
        double acc[8];
        const int32_t o0 = T.out_c0[e], o1 = T.out_c1[e];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[j] = 0.0;
          if (o0 >= 0) acc[j] += zl[o0 * LD + cj + j];
          if (o1 >= 0) acc[j] += zl[o1 * LD + cj + j];
        }
        for (int i = 0; i < s; ++i) {
          const double r = rf[i * b];
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += r * v[i * LD + j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) zl[e * LD + cj + j] = acc[j];
      }
      __syncthreads();
    }
    if (live) {
      for (int l = rt; l < T.n_int; l += RT) {
        const int64_t d = subtree_dof(a, in, T.var_code[l]);
        double2 *dst = reinterpret_cast<double2 *>(a.x + d * MW + c0 + cj);
        const double2 *src = reinterpret_cast<const double2 *>(vs + l * LD + cj);
#pragma unroll
        for (int h = 0; h < 4; ++h) dst[h] = src[h];
      }
      for (int q = rt; q < T.n_rootb; q += RT) {
        double2 *dst = reinterpret_cast<double2 *>(a.z + (in.z_root + q) * MW + c0 + cj);
        const double2 *src = reinterpret_cast<const double2 *>(zl + (zr + q) * LD + cj);
#pragma unroll
        for (int h = 0; h < 4; ++h) dst[h] = src[h];
      }
    }
    __syncthreads();  // vs / zl are reused by the next column chunk
  }
}

template <int MC>
__global__ void __launch_bounds__(256) k_wide_subtree_backward(SubtreeArgs a, int32_t MW) {
  constexpr int CG = MC / 8, RT = 256 / CG, LD = MC + 2;
  extern __shared__ __align__(16) double sm[];
  const InstanceDev in = a.inst[blockIdx.x];
  const TemplateDev T = a.tmpl[in.tmpl];
  uint64_t *bar = reinterpret_cast<uint64_t *>(sm);
  double *blob = sm + 2;
  double *zs = blob + T.rb_size;            // accumulated right-hand sides of the interior variables
  double *xs = zs + int64_t(T.n_int) * LD;  // solution: [interior | root boundary]
  const int tid = threadIdx.x, rt = tid % RT, cg = tid / RT, cj = cg * 8;
  if (tid == 0) mbar_init(bar);
  __syncthreads();
  if (tid == 0) bulk_load_start(blob, a.R + in.rb_base, uint32_t(T.rb_size) * 8u, bar);
  bool waited = false;
  for (int32_t c0 = blockIdx.y * MC; c0 < MW; c0 += gridDim.y * MC) {
    const bool live = c0 + cj < MW;
    for (int l = rt; l < T.n_int + T.n_rootb; l += RT) {
      const int64_t d = subtree_dof(a, in, l < T.n_int ? T.var_code[l] : T.rootb_code[l - T.n_int]);
      const double2 *src = reinterpret_cast<const double2 *>(a.x + d * MW + c0 + cj);
      double2 *dst = reinterpret_cast<double2 *>((l < T.n_int ? zs + l * LD : xs + l * LD) + cj);
#pragma unroll
      for (int h = 0; h < 4; ++h) dst[h] = live ? src[h] : make_double2(0.0, 0.0);
    }
    __syncthreads();
    if (!waited) { mbar_wait(bar, 0); waited = true; }
    for (int lv = T.n_levels - 1; lv >= 0; --lv) {
      for (int e = T.lvl_sep[lv] + rt; e < T.lvl_sep[lv + 1]; e += RT) {
        const int32_t n = T.sep_node[e];
        const int32_t s = T.node_s[n], b = T.node_b[n], sep0 = T.node_sep[n];
        const double *rb = blob + T.node_rb[n] + (e - sep0);  // R^T[p][i] at p * s + i
        const int32_t *bv = T.bnd_var + T.node_bnd[n];
        double acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.0;
        for (int p = 0; p < s; ++p) {
          const double r = rb[p * s];
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += r * zs[(sep0 + p) * LD + cj + j];
        }
        rb += s * s;
        for (int q = 0; q < b; ++q) {
          const double r = rb[q * s];
          const int32_t l = bv[q];
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += r * xs[l * LD + cj + j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) xs[e * LD + cj + j] = acc[j];
      }
      __syncthreads();
    }
    if (live)
      for (int l = rt; l < T.n_int; l += RT) {
        const int64_t d = subtree_dof(a, in, T.var_code[l]);
        double2 *dst = reinterpret_cast<double2 *>(a.x + d * MW + c0 + cj);
        const double2 *src = reinterpret_cast<const double2 *>(xs + l * LD + cj);
#pragma unroll
        for (int h = 0; h < 4; ++h) dst[h] = src[h];
      }
    __syncthreads();
  }
}

void wide_set_smem_limit(size_t bytes) {
  cudaFuncSetAttribute(k_wide_subtree_forward<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes));
  cudaFuncSetAttribute(k_wide_subtree_forward<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes));
  cudaFuncSetAttribute(k_wide_subtree_forward<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes));
  cudaFuncSetAttribute(k_wide_subtree_backward<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes));
  cudaFuncSetAttribute(k_wide_subtree_backward<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes));
  cudaFuncSetAttribute(k_wide_subtree_backward<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes));
}

// Column chunks are dealt to gridDim.y CTAs per subtree: enough CTAs for ~4 waves of the machine (the chunk loop of one
// CTA is serial), few enough that the factor blob a CTA stages is reused for several chunks when subtrees are many.
static unsigned subtree_column_groups(int MC, int32_t MW, int32_t n_inst) {
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int chunks = (MW + MC - 1) / MC;
  const int want = (4 * sms + n_inst - 1) / n_inst;
  return unsigned(std::max(1, std::min(chunks, want)));
}

void launch_wide_subtree_forward(int MC, int32_t MW, const SubtreeArgs &a, int32_t n_inst, size_t smem_bytes, cudaStream_t st) {
  if (n_inst <= 0) return;
  const dim3 grid(unsigned(n_inst), subtree_column_groups(MC, MW, n_inst));
  if (MC == 32) k_wide_subtree_forward<32><<<grid, 256, smem_bytes, st>>>(a, MW);
  else if (MC == 16) k_wide_subtree_forward<16><<<grid, 256, smem_bytes, st>>>(a, MW);
  else k_wide_subtree_forward<8><<<grid, 256, smem_bytes, st>>>(a, MW);
  launch_counter_add(1);
}

void launch_wide_subtree_backward(int MC, int32_t MW, const SubtreeArgs &a, int32_t n_inst, size_t smem_bytes, cudaStream_t st) {
  if (n_inst <= 0) return;
  const dim3 grid(unsigned(n_inst), subtree_column_groups(MC, MW, n_inst));
  if (MC == 32) k_wide_subtree_backward<32><<<grid, 256, smem_bytes, st>>>(a, MW);
  else if (MC == 16) k_wide_subtree_backward<16><<<grid, 256, smem_bytes, st>>>(a, MW);
  else k_wide_subtree_backward<8><<<grid, 256, smem_bytes, st>>>(a, MW);
  launch_counter_add(1);
}

// =================================================================================================================
// fine -> mortar projection of the stored traces of all columns: one sparse row per operator row
// =================================================================================================================
__global__ void __launch_bounds__(256) k_wide_f2c(int64_t n_rows, const int64_t *__restrict__ rp,
                                                  const int64_t *__restrict__ col, const double *__restrict__ val,
                                                  const double *__restrict__ trace, int32_t MW, int32_t n_cols,
                                                  double *__restrict__ T, int64_t ldt, const int64_t *__restrict__ t_row,
                                                  int64_t col0) {
  const int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  const int64_t r = idx / n_cols;
  if (r >= n_rows) return;
  const int j = int(idx - r * n_cols);
  double acc = 0.0;
  for (int64_t e = rp[r]; e < rp[r + 1]; ++e) acc += val[e] * trace[col[e] * MW + j];
  T[t_row[r] * ldt + col0 + j] = acc;
}

void launch_wide_f2c(int64_t n_rows, const int64_t *rp, const int64_t *col, const double *val, const double *trace,
                     int32_t MW, int32_t n_cols, double *T, int64_t ldt, const int64_t *t_row, int64_t col0,
                     cudaStream_t st) {
  if (n_rows <= 0 || n_cols <= 0) return;
  k_wide_f2c<<<unsigned((n_rows * n_cols + 255) / 256), 256, 0, st>>>(n_rows, rp, col, val, trace, MW, n_cols, T, ldt,
                                                                     t_row, col0);
  launch_counter_add(1);
}

// =================================================================================================================
// block lower-triangular Toeplitz operator application (FP64 DMMA)
//   tile = 64 operator rows (a') x 32 mortar time slabs (I) of one (subdomain, output side);
//   k runs over (d, input side, a): A = T_d rows, B[k][I] = q[(side, I - d, a)] (zero for I < d)
// Every output entry is written by exactly one thread: deterministic, no atomics.
// =================================================================================================================
constexpr int TA_LD = 16 + 4;
constexpr int TB_LD = TOEP_BN + 4;
constexpr int T_STAGES = 4;

template <int BM>
struct ToepCfg {
  static constexpr int WM = BM / 16;            // warps along the rows, 16 rows each
  static constexpr int WN = 8 / WM;             // warps along the slabs
  static constexpr int NJ = TOEP_BN / WN / 8;   // 8-slab blocks per warp
  static constexpr int A_PER = BM * 16 / 256;   // staged A elements per thread and k slab
  static constexpr int STAGE_DBL = BM * TA_LD + 16 * TB_LD;
  static constexpr size_t SMEM = size_t(T_STAGES) * STAGE_DBL * sizeof(double);
};

// BM = 128, 64 or 32 operator rows per tile: large tiles re-read q less often (the large problems are bound by the
// HBM stream of T), small tiles give the small problems enough CTAs to fill the machine.
template <int BM>
__global__ void __launch_bounds__(256) k_toeplitz_apply(const ToepProb *__restrict__ probs,
                                                        const ToepTile *__restrict__ tiles,
                                                        const double *__restrict__ q, double *__restrict__ out_parts,
                                                        int64_t part_stride) {
  using C = ToepCfg<BM>;
  extern __shared__ __align__(16) double tsm[];
  const ToepTile tl = tiles[blockIdx.x];
  const ToepProb pr = probs[tl.prob];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp / C::WN, wn = warp % C::WN;
  double acc[2][C::NJ][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < C::NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  double *out = out_parts + int64_t(tl.part) * part_stride;
  int per_d = 0;
  for (int s = 0; s < pr.n_in; ++s) per_d += (pr.P_in[s] + 15) >> 4;
  const int nk = (tl.d1 - tl.d0) * per_d;  // the host clips [d0, d1) to the delays the tile's slabs can see
  // position of the NEXT k slab to be issued
  int d = tl.d0, s = 0, a0 = 0;
  auto issue = [&](int stage) {
    double *as = tsm + stage * C::STAGE_DBL, *bs = as + BM * TA_LD;
    const int P = pr.P_in[s];
    const double *Td = pr.T + int64_t(d) * pr.d_stride + int64_t(pr.row_base + tl.row0) * pr.ldt + pr.col_base[s] + a0;
#pragma unroll
    for (int r = 0; r < C::A_PER; ++r) {
      const int e = tid + 256 * r, i = e >> 4, kk = e & 15;
      const bool ok = tl.row0 + i < pr.P_out && a0 + kk < P;
      cp_async8(as + i * TA_LD + kk, ok ? Td + int64_t(i) * pr.ldt + kk : pr.T, ok);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int e = tid + 256 * r, n = e >> 4, kk = e & 15;
      const int I = tl.slab0 + n, J = I - d;
      const bool ok = I < pr.m_t && J >= 0 && a0 + kk < P;
      cp_async8(bs + kk * TB_LD + n, ok ? q + pr.in_off[s] + int64_t(J) * P + a0 + kk : q, ok);
    }
    a0 += 16;
    if (a0 >= P) { a0 = 0; ++s; }
    if (s >= pr.n_in) { s = 0; ++d; }
  };
#pragma unroll
  for (int st = 0; st < T_STAGES - 1; ++st) {
    if (st < nk) issue(st);
    cp_async_commit();
  }
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<T_STAGES - 2>();
    __syncthreads();
    if (kt + T_STAGES - 1 < nk) issue((kt + T_STAGES - 1) % T_STAGES);
    cp_async_commit();
    const double *as = tsm + (kt % T_STAGES) * C::STAGE_DBL, *bs = as + BM * TA_LD;
#pragma unroll
    for (int kk = 0; kk < 16; kk += 4) {
      double av[2], bv[C::NJ];
#pragma unroll
      for (int i = 0; i < 2; ++i) av[i] = as[(wm * 16 + i * 8 + (lane >> 2)) * TA_LD + kk + (lane & 3)];
#pragma unroll
      for (int j = 0; j < C::NJ; ++j) bv[j] = bs[(kk + (lane & 3)) * TB_LD + (wn * C::NJ + j) * 8 + (lane >> 2)];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < C::NJ; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], av[i], bv[j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < C::NJ; ++j) {
      const int row = tl.row0 + wm * 16 + i * 8 + (lane >> 2);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int I = tl.slab0 + (wn * C::NJ + j) * 8 + (lane & 3) * 2 + t;
        if (row < pr.P_out && I < pr.m_t) out[pr.out_off + int64_t(I) * pr.P_out + row] = acc[i][j][t];
      }
    }
}

void launch_toeplitz_apply(int bm, const ToepProb *probs, const ToepTile *tiles, int32_t n_tiles, const double *q,
                           double *out_parts, int64_t part_stride, cudaStream_t st) {
  if (n_tiles <= 0) return;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(k_toeplitz_apply<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(ToepCfg<128>::SMEM));
    cudaFuncSetAttribute(k_toeplitz_apply<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(ToepCfg<64>::SMEM));
    cudaFuncSetAttribute(k_toeplitz_apply<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(ToepCfg<32>::SMEM));
    attr_set = true;
  }
  if (bm == 128) k_toeplitz_apply<128><<<n_tiles, 256, ToepCfg<128>::SMEM, st>>>(probs, tiles, q, out_parts, part_stride);
  else if (bm == 64) k_toeplitz_apply<64><<<n_tiles, 256, ToepCfg<64>::SMEM, st>>>(probs, tiles, q, out_parts, part_stride);
  else k_toeplitz_apply<32><<<n_tiles, 256, ToepCfg<32>::SMEM, st>>>(probs, tiles, q, out_parts, part_stride);
  launch_counter_add(1);
}

__global__ void __launch_bounds__(256) k_sum_parts(int64_t n, const double *__restrict__ parts, int64_t part_stride,
                                                   int32_t n_parts, double *__restrict__ out) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  double acc = parts[i];
  for (int p = 1; p < n_parts; ++p) acc += parts[p * part_stride + i];
  out[i] = acc;
}

void launch_sum_parts(int64_t n, const double *parts, int64_t part_stride, int32_t n_parts, double *out, cudaStream_t st) {
  if (n <= 0) return;
  k_sum_parts<<<unsigned((n + 255) / 256), 256, 0, st>>>(n, parts, part_stride, n_parts, out);
  launch_counter_add(1);
}

}  // namespace mmmfe
