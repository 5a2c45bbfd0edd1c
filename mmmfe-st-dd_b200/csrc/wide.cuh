// Multi-column ("wide") variants of the solve and time-stepping kernels, and the time-Toeplitz operator kernel.
//
// The multiscale flux basis (the reference's dormant compute_multiscale_basis, src/darcy_vtdd.cc:959-1003; members
// inc/darcy_vtdd.h:124, 264) is the response of a subdomain to every mortar basis function: many right-hand sides of
// ONE factor.  With MW columns at once every front of the multifrontal solve is a GEMM (FP64 DMMA, k_grouped_gemm)
// instead of a GEMV; vectors are stored [row][MW].  Backward Euler with a time-invariant matrix and a zero start is
// invariant under time translation, so only the basis functions of the FIRST mortar time slab are solved for; the
// interface operator is then block lower-triangular Toeplitz in the mortar time slabs (k_toeplitz_apply).
#pragma once
#include "kernels.cuh"

namespace mmmfe {

struct WideStep {
  int32_t MW;  // columns = leading dimension of x, ptil and of the trace rows
  int64_t n_flux, n_pressure;
  const int64_t *kup_rp;
  const int32_t *kup_col;
  const double *kup_val;
  const int64_t *kpu_rp;
  const int32_t *kpu_col;
  const double *kpu_val;
  const double *minv;
  double *x;       // [n_flux][MW]
  double *ptil;    // [n_pressure][MW]
  int32_t *state;  // device scalars {time level, pass}: the step is one static CUDA graph, replayed nt times
};
// x = -K_up ptil
void launch_wide_rhs(const WideStep &a, cudaStream_t st);
// x[row[e]][col[e]] += val[e] for the entries e of (pass, level): lvl_rp is indexed [pass * (n_lvl + 1) + level];
// levels >= n_lvl carry no data.  The grid covers max_entries.
void launch_wide_add(const WideStep &a, const int64_t *lvl_rp, const int32_t *row, const int32_t *col, const double *val,
                     int32_t n_lvl, int64_t max_entries, cudaStream_t st);
// ptil -= M^-1 K_pu x ; trace[level][t][:] = x[tr_dof[t]][:]
void launch_wide_update(const WideStep &a, const int32_t *tr_dof, int32_t n_tr, double *trace, cudaStream_t st);
void launch_wide_next_level(int32_t *state, cudaStream_t st);
// *out = max(*out, bits of max |v|): zero *out first; the bit pattern of a non-negative double orders like the value
void launch_absmax(int64_t n, const double *v, unsigned long long *out, cudaStream_t st);

struct WideRow {
  int32_t node, p;
};
// forward: z[front][p] = (p < s ? x[idx[p]] : 0) + the children's contributions (separator accumulation and
// pass-through of the boundary entries); the GEMM then adds -G v_s to the boundary part
void launch_wide_fwd_gather(const SolveArgs &a, int32_t MW, const WideRow *rows, int64_t n_rows, cudaStream_t st);
// backward: z[front][s + q] = x[idx[s + q]]
void launch_wide_bwd_gather(const SolveArgs &a, int32_t MW, const WideRow *rows, int64_t n_rows, cudaStream_t st);
// one CTA per subtree: the factor blob is staged once and reused for all column chunks of MC (4 or 8) columns
void launch_wide_subtree_forward(int MC, int32_t MW, const SubtreeArgs &a, int32_t n_inst, size_t smem_bytes, cudaStream_t st);
void launch_wide_subtree_backward(int MC, int32_t MW, const SubtreeArgs &a, int32_t n_inst, size_t smem_bytes, cudaStream_t st);
void wide_set_smem_limit(size_t bytes);

// T[t_row[r] * ldt + col0 + j] = sum_e val[e] * trace[col[e] * MW + j]   for r < n_rows, j < n_cols
void launch_wide_f2c(int64_t n_rows, const int64_t *rp, const int64_t *col, const double *val, const double *trace,
                     int32_t MW, int32_t n_cols, double *T, int64_t ldt, const int64_t *t_row, int64_t col0,
                     cudaStream_t st);

// ---- block lower-triangular Toeplitz operator -----------------------------------------------------------------------
// One problem = one (subdomain, output side): out[(I, a')] = sum_{d <= I} sum_{input sides s} sum_a
//   T[d][row_base + a'][col_base[s] + a] * q[(s, I - d, a)]
struct ToepProb {
  const double *T;  // [m_t][rows][ldt] of the subdomain's factor group
  int64_t ldt, d_stride;
  int32_t m_t, P_out, row_base, n_in;
  int64_t out_off;
  int32_t P_in[4], col_base[4];
  int64_t in_off[4];
};
struct ToepTile {
  int32_t prob, row0, slab0;
  int32_t d0, d1;  // delays [d0, d1) of this tile (split-k: the delay range of a slab tile is cut into parts)
  int32_t part;    // partial-sum buffer this tile writes (0 .. n_parts-1)
};
constexpr int TOEP_BN = 32;  // mortar time slabs per tile; rows per tile: 128, 64 or 32 (chosen by the caller)
// Tiles write out_parts[part * part_stride + ...]; entries no tile of a part covers must be zero (static structure:
// zero the buffer once).  launch_sum_parts adds the parts in fixed order: deterministic.
void launch_toeplitz_apply(int rows_per_tile, const ToepProb *probs, const ToepTile *tiles, int32_t n_tiles,
                           const double *q, double *out_parts, int64_t part_stride, cudaStream_t st);
void launch_sum_parts(int64_t n, const double *parts, int64_t part_stride, int32_t n_parts, double *out, cudaStream_t st);

}  // namespace mmmfe
