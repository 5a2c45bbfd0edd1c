// Device-side data structures and kernel launch wrappers of the interface-GMRES engine (sm_100a).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace mmmfe {

// kernels launched through the wrappers below (graph capture counts once; replays are added by the engine)
long long launch_counter();
void launch_counter_add(long long n);

// One front of the nested-dissection tree as the solve kernels see it.
struct FrontDesc {
  int32_t s, b, ldr, ldf, nsplit;
  int32_t child0, child1;
  int64_t rb_off;   // R = [F11^-1 | -G^T], s rows of ldr doubles (backward)
  int64_t rf_off;   // -G^T, s rows of ldf doubles (forward copy, contiguous)
  int64_t idx_off;  // s + b user flux dof indices (separator first)
  int64_t src_off;  // 2 * (s + b) child gather maps
  int64_t z_off;    // s + nsplit * b per-front vector entries
};

// Subtree templates / instances on the device (see nd_tree.h).
struct TemplateDev {
  int32_t n_nodes, n_levels, n_int, n_rootb, zl_size, rf_size, rb_size;
  const int32_t *node_s, *node_b, *node_rf, *node_rb, *node_sep, *node_bnd, *node_zl;
  const int32_t *var_code, *rootb_code, *sep_node, *sep_c0, *sep_c1, *out_node, *out_c0, *out_c1, *bnd_var;
  const int32_t *lvl_sep, *lvl_out;
};

struct InstanceDev {
  int32_t tmpl, I0, J0, pad;
  int64_t rf_base, rb_base, z_root;  // z_root: first boundary-contribution entry of the root in z
};

struct SubtreeArgs {
  const TemplateDev *tmpl;
  const InstanceDev *inst;
  const double *R;
  double *x, *z;
  const int32_t *dof_of_canon;  // nullptr when the user numbering is canonical
  int32_t nx;
  int64_t nxe;
};
void launch_subtree_forward(int M, const SubtreeArgs &a, int32_t n_inst, size_t smem_bytes, cudaStream_t st);
void launch_subtree_backward(int M, const SubtreeArgs &a, int32_t n_inst, size_t smem_bytes, cudaStream_t st);
// sets the dynamic shared memory limit of the subtree kernels; returns the device's opt-in maximum
size_t subtree_smem_limit();

// Work item of the large-front kernels.
//  forward : rows [r0,r1) of R times boundary columns [c0,c1); writes partial contribution `split`
//  backward: rows [r0,r1) of R
struct WorkItem {
  int32_t node, r0, r1, c0, c1, split;
};

struct GemmProb {
  const double *A, *B;
  double *C;
  int32_t m, n, k, ldc;
  int64_t ars, acs, brs, bcs;  // element (i,k) of A at A[i*ars + k*acs]; (k,j) of B at B[k*brs + j*bcs]
  double alpha, beta;
  const int32_t *crow;  // optional row map of C: row i is written at C + crow[i] * ldc (nullptr: i)
};

struct InvProb {
  double *A;
  int32_t n, ld;
};

struct TransProb {  // dst(i,j) = scale * src(j,i), dst is rows x cols
  const double *src;
  double *dst;
  int32_t rows, cols, ld_src, ld_dst;
  double scale;
};

// ---- factorisation ---------------------------------------------------------------------------------------------
void launch_grouped_gemm(const GemmProb *d_probs, const int32_t *d_tile_start, int32_t n_probs, int32_t n_tiles,
                         cudaStream_t st);
void launch_small_inverse(const InvProb *d_probs, int32_t n_probs, int32_t *d_fail, cudaStream_t st);
void launch_transpose(const TransProb *d_probs, int32_t n_probs, cudaStream_t st);
// scatter S (user numbering CSR, n_flux rows) into the fronts
void launch_assemble_original(int64_t n_flux, const int64_t *s_rp, const int32_t *s_col, const double *s_val,
                              const int32_t *canon_of_dof, const int32_t *dof_of_canon, const int32_t *node_of,
                              const int32_t *pos_of, const FrontDesc *fronts, const int64_t *f_off,
                              const int32_t *idx_canon, double *F, cudaStream_t st);
void launch_extend_add(const int32_t *d_nodes, int32_t n_nodes, const FrontDesc *fronts, const int64_t *f_off,
                       const int32_t *src, double *F, cudaStream_t st);

// ---- solve -----------------------------------------------------------------------------------------------------
struct SolveArgs {
  const FrontDesc *fronts;
  const double *R;
  const int32_t *idx;
  const int32_t *src;
  double *z;  // per-front vectors, [entry][M]
  double *x;  // rhs in / solution out, [flux dof][M]
};
void launch_forward_large(int M, const SolveArgs &a, const WorkItem *items, int32_t n_items, cudaStream_t st);
void launch_backward_large(int M, const SolveArgs &a, const WorkItem *items, int32_t n_items, int32_t f_max,
                           cudaStream_t st);

// ---- time stepping ---------------------------------------------------------------------------------------------
struct StepArgs {
  int32_t M;
  int64_t n_flux, n_pressure;
  // K_up (flux rows x pressure cols) and K_pu (pressure rows x flux cols) of the saddle matrix, 1 / diag(M)
  const int64_t *kup_rp;
  const int32_t *kup_col;
  const double *kup_val;
  const int64_t *kpu_rp;
  const int32_t *kpu_col;
  const double *kpu_val;
  const double *minv;
  double *x;     // [n_flux][M]
  double *ptil;  // [n_pressure][M]  p_old (+ M^-1 src for the bar problem)
  // interface entries of this group: q-th entry = flux dof tr_dof[q] of member tr_mem[q], trace/lam_bar slot
  // tr_base[q] + level * tr_stride[q], outward sign tr_sign[q]
  int32_t n_tr;
  const int32_t *tr_dof, *tr_mem, *tr_stride;
  const int64_t *tr_base;
  const double *tr_sign;
  const int32_t *iface_q;  // [n_flux][M] -> q or -1
  // when set, the kernels take the time level from this device scalar instead of their `level` argument: one time
  // step can then be captured as a static CUDA graph and replayed for every level (bar and final sweeps)
  const int32_t *level_dev;
};
// x = - K_up ptil - sign * lam_bar(level)      (assemble_rhs_star/bar after pressure elimination)
void launch_build_rhs(const StepArgs &a, const double *lam_bar, int32_t level, cudaStream_t st);
// x[dofs[i]][member] += vals[i]
void launch_add_sparse(double *x, int32_t M, int32_t member, const int32_t *dofs, const double *vals, int32_t n,
                       cudaStream_t st, const int32_t *level_dev = nullptr);  // level_dev: vals += *level_dev * n
// p = ptil - M^-1 K_pu x ; ptil_next = p + M^-1 src_next ; trace(level) = x[interface dofs] ; optional history
// n_levels > 0: src_next is only read while level + 1 < n_levels (graph replays pass it for every level)
void launch_update(const StepArgs &a, double *trace, int32_t level, const double *const *src_next,
                   int64_t src_stride, double *const *hist, int64_t hist_level_stride, int add_hist,
                   cudaStream_t st, int32_t n_levels = 0);
void launch_init_ptil(const StepArgs &a, const double *const *p0, const double *const *src0, cudaStream_t st);
// ptil[cell_box[c]][j] = p0[j][p] + minv * src0[j][p], p = p_of_cell[c] (nullptr: c): the pressures of the persistent
// sweep kernel (box-major cell order) at the start of a bar sweep
void launch_init_ptil_box(int64_t n_cells, int32_t M, const int32_t *cell_box, const int32_t *p_of_cell, const double *minv_box,
                          double minv_const, const double *const *p0, const double *const *src0, double *ptil, cudaStream_t st);

// ---- interface vector ------------------------------------------------------------------------------------------
void launch_spmv(int64_t n_rows, const int64_t *rp, const int32_t *col, const double *val, const double *x,
                 double *y, cudaStream_t st);
// ap = -(send + (partner >= 0 ? send[partner] : recv[-partner-2]))
void launch_exchange_combine(int64_t n, const double *send, const int64_t *partner, const double *recv, double sign,
                             double *out, cudaStream_t st);
void launch_gather(int64_t n, const double *srcv, const int64_t *index, double *dst, cudaStream_t st);
// h[i] = <w, Q_i>, i = 0..k  (classical Gram-Schmidt: all against the unmodified w).  scratch holds
// arnoldi_scratch_doubles(len, k + 1) doubles, ticket is one zeroed counter (left zeroed); both are shared with
// launch_cgs_update (stream-ordered).  Grid-wide, one launch, fixed summation order.
size_t arnoldi_scratch_doubles(int64_t len, int32_t max_vectors);
void launch_multi_dot(int64_t len, const double *w, const double *Q, int64_t ldq, int32_t k_plus_1, double *h,
                      double *scratch, unsigned int *ticket, cudaStream_t st);
// w -= sum_i h[i] Q_i ; norm2[0] = sum w^2
void launch_cgs_update(int64_t len, double *w, const double *Q, int64_t ldq, int32_t k_plus_1, const double *h,
                       double *norm2, double *scratch, unsigned int *ticket, cudaStream_t st);
// dst = w / inv, or w / sqrt(*norm2_dev) when norm2_dev is given
void launch_scale_store(int64_t len, const double *w, double inv, double *dst, cudaStream_t st,
                        const double *norm2_dev = nullptr);
void launch_axpy_rows(int64_t len, const double *Q, int64_t ldq, int32_t n_rows, const double *y, double *out,
                      cudaStream_t st);

// ---- diagnostics -----------------------------------------------------------------------------------------------
void launch_fill_hash(int64_t n, uint64_t seed, double *out, cudaStream_t st);
void launch_diff_norms(int64_t n, const double *a, const double *b, double *out, cudaStream_t st);
double measure_dmma_tflops(cudaStream_t st);

}  // namespace mmmfe
