/* mmmfe_b200.h -- C ABI of the B200 (sm_100a) interface-GMRES engine.
 *
 * This is the drop-in boundary for the hot path of MMMFE-ST-DD.  The reference has no
 * FFI; the seams this library replaces are member functions of DarcyVTProblem<2>.  Every entry point names
 * the reference site it stands for (paths relative to the reference repository).
 *
 * Conventions
 *   - plain C, no exceptions cross the boundary; every call returns 0 on success or a negative MMMFE_E_* code,
 *     mmmfe_last_error() returns the message of the last failure on that context;
 *   - one context per process and per refinement cycle, bound to one CUDA device, used from one host thread;
 *   - all pointer arguments are HOST buffers owned by the caller; inputs are copied during the call, outputs are
 *     written before the call returns; device memory is owned by the library;
 *   - "side" is the reference side index 0 bottom, 1 right, 2 top, 3 left (boundary_id - 1,
 *     inc/utilities.h:329-375); outward normal signs {-1,+1,+1,-1} (inc/utilities.h:79-91);
 *   - a flux dof is the flux integral through its edge in the positive coordinate direction (deal.II RT0);
 *   - FP64 throughout.
 */
#ifndef MMMFE_B200_H
#define MMMFE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMMFE_OK 0
#define MMMFE_E_INVALID (-1)     /* bad argument / call order                                   */
#define MMMFE_E_UNSUPPORTED (-2) /* structure outside the path (e.g. c0 == 0, non-diagonal M)   */
#define MMMFE_E_CUDA (-3)        /* CUDA runtime failure (message has the CUDA error string)    */
#define MMMFE_E_NUMERIC (-4)     /* non-positive pivot in a front                               */
#define MMMFE_E_COMM (-5)        /* NCCL failure                                                */

typedef struct mmmfe_ctx mmmfe_ctx;

/* Sparse matrix in CSR, 0-based, host memory. */
typedef struct {
  int64_t n_rows, n_cols;
  const int64_t *row_ptr; /* n_rows + 1 */
  const int32_t *col;     /* nnz        */
  const double *val;      /* nnz        */
} mmmfe_csr;

/* One rectangular subdomain = one MPI rank of the reference (src/darcy_vtdd.cc:2082-2103). */
typedef struct {
  int32_t global_id;       /* reference rank r; ix = r % n0, iy = r / n0 (inc/utilities.h:152-157)         */
  int32_t nx, ny, nt;      /* cells in x, y and backward-Euler steps (mesh_pattern_sub_d<r>)               */
  double dt;               /* final_time / nt (src/darcy_vtdd.cc:2083)                                     */
  /* Saddle matrix [[A, -B^T], [dt B, c0 M]] exactly as assemble_system builds it
   * (src/darcy_vtdd.cc:464-466); unknowns [fluxes ; pressures], n = n_flux + n_pressure.                  */
  mmmfe_csr matrix;
  int32_t n_flux, n_pressure;
  /* canon_of_dof[d] = canonical index of user dof d, or NULL when the user numbering is canonical:
   * x-edges (i=0..nx, j=0..ny-1) at j*(nx+1)+i, y-edges (i=0..nx-1, j=0..ny) at (nx+1)*ny + j*nx + i,
   * cells at n_flux + j*nx + i.  Used only to build the nested-dissection tree.                           */
  const int32_t *canon_of_dof;
  /* Interface sides.  neighbor[s] = global id of the subdomain across side s, or -1
   * (find_neighbors, inc/utilities.h:176-306).  side_dofs[s] = the n_side[s] 2-D interface flux dofs in the
   * order of interface_dofs_subd[s] (src/darcy_vtdd.cc:559-584).                                          */
  int32_t neighbor[4];
  int32_t n_side[4];
  const int32_t *side_dofs[4];
  /* Projector tables of project_mortar (inc/utilities.h:471-505) for each interface side, as explicit
   * sampling operators:
   *   f2c[s] : (mortar_len[s]) x (n_side[s]*nt)  fine 2-D flux trace U[level*n_side+i] -> mortar coefficients
   *   c2f[s] : (n_side[s]*nt) x (mortar_len[s])  mortar coefficients -> lam_bar = (2-D interface dof)/|e|
   * mortar_len[s] = interface_dofs[s].size() (src/darcy_vtdd.cc:524-541).
   * Sides WITHOUT a neighbour may also carry n_side, side_dofs, mortar_len and both tables (else leave them zero):
   * they are never applied to this subdomain, they let ranks share the multiscale-basis work of a factor group.  */
  int32_t mortar_len[4];
  mmmfe_csr f2c[4];
  mmmfe_csr c2f[4];
} mmmfe_subdomain;

/* ---- lifetime ------------------------------------------------------------------------------------------------- */
/* device < 0: use the current device. */
int mmmfe_create(mmmfe_ctx **ctx, int device);
int mmmfe_destroy(mmmfe_ctx *ctx);
const char *mmmfe_last_error(const mmmfe_ctx *ctx);
const char *mmmfe_version(void);

/* Options (set before mmmfe_setup): "keep_history" (default 1: keep every bar level on the device so that the
 * final sweep can return full solutions), "use_graphs" (default 1: replay each sweep as a CUDA graph),
 * "leaf_cells" (default 4: nested-dissection leaf size), "subtree_cells" (default 256: largest box solved by one CTA
 * from shared memory; 0 disables the subtree kernels), "use_sweep" (default 1: run the star sweeps in the
 * persistent sweep kernel when the matrix has the RT0 structure), "sweep_chunk" (doubles per streamed chunk),
 * "sweep_min_fill" (entries the shared-memory ring must hold), "sweep_evict_first" (L2 hint of the factor stream),
 * "sweep_compute_threads" (64, 128 or 256 compute threads per CTA), "sweep_ctas_per_sm" (resident CTAs per SM),
 * "sweep_groups" (independent compute groups per CTA), "sweep_autotune" (default 1: time one star sweep with the
 * persistent kernel and one with the per-level kernels at set-up, compare their traces (set-up fails above 1e-11) and
 * keep the faster; the decision is summed over the ranks), "sweep_crosscheck" (with sweep_autotune = 0: still run and
 * compare both once; "tune_steps", default 32: time levels of the longest sweep that comparison runs on, 0 = all),
 * "sweep_bar" / "sweep_final" (default 1: the bar and final sweeps run through the persistent kernel where the star
 * sweeps do), "sweep_pub_fence" (-1 = by the number of right-hand sides: entries with several completion counters
 * publish with one release fence), "gmres_lookahead" (iterations enqueued ahead of the host's stop test; -1 = 3 on
 * the multiscale-basis operator, 0 on the star sweeps; iteration counts, residuals and lambda do not depend on it).
 * Interface operator: "operator" 0 = nt star solves per subdomain in every GMRES iteration (the reference's loop,
 * src/darcy_vtdd.cc:1316-1411), 1 = multiscale flux basis (the reference's dormant compute_multiscale_basis,
 * src/darcy_vtdd.cc:959-1003): the response to every mortar basis function of the FIRST mortar time slab is computed
 * once at set-up (multi-column multifrontal solves, fronts as FP64 DMMA GEMMs) and S q becomes one block
 * lower-triangular Toeplitz product per subdomain -- the same operator to round-off; 2 (default) = 1 when the
 * projector tables are slab structured and a factor group has at most "basis_auto_columns" (default 2048) basis
 * columns, else 0.  "mortar_time_slabs" = time steps of the mortar mesh (mesh_pattern_mortar nt; required for 1/2),
 * "basis_columns" = columns solved per pass (default 256), "basis_share" (default 1: with several ranks, a factor
 * group that every rank holds is solved for once, each rank taking a block of its columns; needs the projector tables
 * of all four sides, see mmmfe_subdomain), "basis_drop_tol" (default 1e-18; 0 = off): a pass of the basis solve stops
 * at the end of a mortar slab once the pressures -- the only state carried between steps -- are below this fraction of
 * their peak (no data after the first slab, backward Euler contracts), the later delays of the operator are zero.  */
int mmmfe_set_option(mmmfe_ctx *ctx, const char *key, double value);

/* ---- multi-GPU: one process per GPU, subdomains sharded over ranks ------------------------------------------ */
/* Size in bytes of the NCCL unique id blob; mmmfe_comm_unique_id fills one (call on rank 0, broadcast it by any
 * out-of-band means, e.g. torch.distributed), mmmfe_comm_init joins.  Single-rank runs skip both.
 * Communicators are cached per (id, rank, world, device) for the life of the process: the contexts of successive
 * refinement cycles join with the same id and share one communicator.
 * Replaces MPI_COMM_WORLD of the reference (src/darcy_vtdd.cc:89).                                            */
int mmmfe_comm_unique_id_size(void);
int mmmfe_comm_unique_id(void *id_out);
int mmmfe_comm_init(mmmfe_ctx *ctx, int rank, int world, const void *id);
/* owner_rank[g] = rank that holds global subdomain g (length n_global); required before mmmfe_setup when
 * world > 1, optional (all zero) otherwise.                                                                   */
int mmmfe_set_owners(mmmfe_ctx *ctx, int32_t n_global, const int32_t *owner_rank);
/* Opens the NCCL paths this rank will use -- the point-to-point channels to `peers` (the ranks that own a
 * neighbour of one of its subdomains; the relation is symmetric, every peer must list this rank too) and the
 * all-reduce path -- by exchanging one double.  NCCL opens them on first use (measured: 1.5 s on 8 ranks); a
 * front-end calls this right after mmmfe_comm_init from a helper thread while it assembles its matrices, so that
 * the cost is hidden.  Collective over the communicator; no other call on the context may run concurrently.      */
int mmmfe_comm_warmup(mmmfe_ctx *ctx, int32_t n_peers, const int32_t *peers);
/* In-place sum over ranks of n host doubles (error numerators/denominators: MPI_Reduce, src/darcy_vtdd.cc:1812-1813).
 * A no-op on one rank.                                                                                        */
int mmmfe_comm_allreduce_sum(mmmfe_ctx *ctx, double *inout, int32_t n);

/* ---- setup: once per refinement cycle ----------------------------------------------------------------------- */
/* Registers a subdomain held by this rank; returns its local index. */
int mmmfe_add_subdomain(mmmfe_ctx *ctx, const mmmfe_subdomain *sd, int32_t *local_index);
/* Groups subdomains with identical matrices, builds the nested-dissection tree of each group and factorises it
 * on the device (multifrontal, block-inverse fronts, FP64 DMMA); builds the interface layout and exchange plan.
 * Replaces A_direct.initialize(system_matrix) (src/darcy_vtdd.cc:483) and the buffer set-up at the top of
 * local_gmres (src/darcy_vtdd.cc:1149-1223).                                                                  */
int mmmfe_setup(mmmfe_ctx *ctx);

/* Sizes. The local interface vector is the concatenation, over local subdomains in registration order and sides
 * 0..3 with a neighbour, of that side's mortar coefficients (every interface therefore appears on both of its
 * subdomains, as in the reference).                                                                           */
int64_t mmmfe_interface_length(const mmmfe_ctx *ctx);
int64_t mmmfe_interface_offset(const mmmfe_ctx *ctx, int32_t local_index, int32_t side);
int32_t mmmfe_num_factor_groups(const mmmfe_ctx *ctx);
/* Stored factor values of all groups (doubles) and the bytes one star step of every local subdomain streams. */
int64_t mmmfe_factor_values(const mmmfe_ctx *ctx);
int64_t mmmfe_step_bytes(const mmmfe_ctx *ctx);
double mmmfe_factor_flops(const mmmfe_ctx *ctx);
double mmmfe_factor_seconds(const mmmfe_ctx *ctx);
/* Kernels this library has launched on the device since the context was created (graph replays count their
 * kernel nodes).                                                                                              */
int64_t mmmfe_kernel_launches(const mmmfe_ctx *ctx);

/* ---- bar sweep: solve_timestep(0, .) for every level (src/darcy_vtdd.cc:884-890, 902-920) -------------------- */
/* Per local subdomain, the caller provides the data-dependent right-hand sides assemble_rhs_bar would build
 * (src/darcy_vtdd.cc:632-766) WITHOUT the c0*(p_old, w) term, which the library adds:
 *   p0        [n_pressure]            projected initial pressure (src/darcy_vtdd.cc:2172-2176); NULL = 0
 *   src       [nt][n_pressure]        dt * (f(t_n), w) per cell (src/darcy_vtdd.cc:718)
 *   fr_dofs   [n_fr]                  flux dofs that carry outer-boundary data (src/darcy_vtdd.cc:732-755)
 *   fr_vals   [nt][n_fr]              their right-hand-side values per level
 * The bar solution of every level is kept on the device for the final sweep; interface traces are projected to
 * the mortar space and the initial residual r = sum over the two sides of s*f2c(bar trace) is formed
 * (src/darcy_vtdd.cc:1203-1256).                                                                              */
int mmmfe_bar_set_data(mmmfe_ctx *ctx, int32_t local_index, const double *p0, const double *src, int32_t n_fr,
                       const int32_t *fr_dofs, const double *fr_vals);
/* The same, for callers that evaluate the data ON THE DEVICE (the host front-end compiles the inc/data.h function
 * classes for the GPU, host/data_device.cu): the library allocates the three arrays (zero-filled), copies fr_dofs
 * (HOST) and returns DEVICE pointers into its own buffers, laid out as above, for the caller's kernels to fill on
 * the legacy default stream (or any stream it synchronises) before mmmfe_bar_sweep.  The library keeps ownership. */
int mmmfe_bar_data_buffers(mmmfe_ctx *ctx, int32_t local_index, int32_t n_fr, const int32_t *fr_dofs,
                           double **d_p0, double **d_src, double **d_fr_vals);
int mmmfe_bar_sweep(mmmfe_ctx *ctx);

/* ---- interface operator and GMRES ------------------------------------------------------------------------------ */
/* One operator application Ap = -(send + recv) with send = s * f2c(star trace(c2f(q)))
 * (src/darcy_vtdd.cc:1316-1411): HOST in, HOST out, length mmmfe_interface_length.  Collective over ranks.     */
int mmmfe_apply_interface_operator(mmmfe_ctx *ctx, const double *q, double *ap);
/* Initial residual r = sum over both sides of s * f2c(bar trace) (src/darcy_vtdd.cc:1212-1256), HOST out. */
int mmmfe_get_initial_residual(mmmfe_ctx *ctx, double *r);
/* Page-locked host buffers for callers that drive the operator from host memory. */
void *mmmfe_alloc_pinned(int64_t bytes);
void mmmfe_free_pinned(void *p);
/* Same, operating on the library's device-resident buffers; `repeat` applications back to back with q fixed.
 * Used by benchmarks (inputs already in HBM); returns the device time per application in milliseconds.        */
int mmmfe_apply_interface_operator_device(mmmfe_ctx *ctx, int32_t repeat, double *ms_per_apply);

typedef struct {
  int32_t iterations;     /* operator applications = "# GMRES" (cg_iteration, src/darcy_vtdd.cc:1363, 1821) */
  int32_t converged;      /* stop test |Beta_{k+1}| / r_norm < tol hit (src/darcy_vtdd.cc:1478-1492)        */
  double r_norm;          /* printed "r_norm" (every interface counted twice, src/darcy_vtdd.cc:1265-1276)  */
  double final_residual;  /* last relative residual                                                           */
  double seconds;         /* wall time of the loop                                                            */
  double device_ms;       /* CUDA-event time of the iteration loop on the interface stream                    */
} mmmfe_gmres_info;

/* local_gmres (src/darcy_vtdd.cc:1139-1521): classical Gram-Schmidt Arnoldi, Givens least squares, the reference's
 * stop test and its back_solve over k_counter columns (y[k] = 0 on a converged exit).  Requires a bar sweep.
 * residual_history may be NULL, else receives min(history_cap, iterations+1) relative residuals.
 * lambda_out (HOST, interface length) may be NULL.  Collective over ranks.                                    */
int mmmfe_gmres_solve(mmmfe_ctx *ctx, double tol, int32_t max_iter, mmmfe_gmres_info *info,
                      double *residual_history, int32_t history_cap, double *lambda_out);

/* ---- final sweep: solve_timestep(2, .) (src/darcy_vtdd.cc:937-953, 1551-1558) ---------------------------------- */
/* Star solve with the converged lambda plus the stored bar solution.  solution[nt][n_flux + n_pressure] of the
 * given local subdomain (user dof numbering) is written to HOST memory.                                        */
int mmmfe_final_sweep(mmmfe_ctx *ctx);
int mmmfe_get_solution(mmmfe_ctx *ctx, int32_t local_index, double *solution);
/* DEVICE pointer to the same array (valid until the context is destroyed or the next bar sweep): error norms and
 * the space-time output are computed from it without copying the solution to the host.                          */
int mmmfe_solution_device(mmmfe_ctx *ctx, int32_t local_index, double **d_solution);
/* Interface flux traces of the last star (or final) sweep of one side: [nt][n_side]. */
int mmmfe_get_star_trace(mmmfe_ctx *ctx, int32_t local_index, int32_t side, double *trace);

/* ---- single-subdomain kernels exposed for parity tests --------------------------------------------------------- */
/* x = K^{-1} rhs for the saddle matrix of one local subdomain (A_direct.vmult, src/darcy_vtdd.cc:864, 877). */
int mmmfe_solve_saddle(mmmfe_ctx *ctx, int32_t local_index, const double *rhs, double *x);

/* ---- profiling ---------------------------------------------------------------------------------------------------- */
/* Runs ONE star time step of every batch without graphs, with a CUDA event between kernel classes, and reports
 * per class the device milliseconds, the algorithmic bytes streamed and the number of launches.
 * Classes: 0 rhs build, 1 subtree forward (bottom of the tree, one CTA per subtree), 2 forward top levels,
 * 3 backward top levels, 4 subtree backward, 5 pressure update + trace.                                                                                  */
#define MMMFE_N_KERNEL_CLASSES 6
int mmmfe_profile_step(mmmfe_ctx *ctx, double ms[MMMFE_N_KERNEL_CLASSES], double bytes[MMMFE_N_KERNEL_CLASSES],
                       int64_t launches[MMMFE_N_KERNEL_CLASSES]);

/* Times ONE full star sweep (all local time steps, one cooperative launch of the persistent sweep kernel) of every
 * batch that runs on the sweep path: device milliseconds, algorithmic bytes per time step, steps.  n_out = batches
 * reported (0 when the generic kernels are in use).                                                          */
int mmmfe_profile_sweeps(mmmfe_ctx *ctx, int32_t cap, double *ms, double *bytes_per_step, int32_t *n_steps,
                         int32_t *n_out);
/* One instrumented star sweep of sweep batch `batch`: mean SM cycles per CTA spent in {subtree forward, top forward,
 * top backward, subtree backward} entries, [4] completion-counter waits of the dependency warp, [5] TMA (mbarrier)
 * waits, [6] entries, [7] waits for the dependency warp, [8..10] subtree backward: input gather / levels / epilogue,
 * [11..13] subtree forward: right-hand side / levels / output, [14] top-level vector gathers, [15] top-level tiles,
 * [16..19] dependency waits by entry kind, [20..23] entries by kind that waited more than 2000 cycles. */
int mmmfe_profile_sweep_cycles(mmmfe_ctx *ctx, int32_t batch, double out[24]);
/* Clock stamps (start, opened by the dependency warp, end) of every entry of compute group 0 during the second time
 * step of one instrumented sweep of sweep batch `batch`, with {kind, tree height} per entry and the per-CTA entry
 * ranges (sweep_ctas + 1 ints).  Call with stamps = NULL to obtain n_entries.                                   */
int mmmfe_debug_sweep_trace(mmmfe_ctx *ctx, int32_t batch, int64_t cap, int64_t *n_entries, int32_t *cta_begin,
                            int32_t *kind_height, int64_t *stamps);
/* Integer facts about the set-up context: "batches", "sweep_batches", "sweep_ctas", "sweep_smem_bytes",
 * "sweep_entries", "sweep_factor_bytes", "rt0_factors", "sweep_planned" (batches with a sweep plan, whether or
 * not the auto-tuner kept it), "tune_us_sweep" / "tune_us_levels" (microseconds of the timed star sweep of each
 * implementation), "tune_rel_diff_e18" (their relative l2 difference * 1e18), "basis_operator" (1: GMRES runs on the
 * multiscale-basis operator), "basis_us", "basis_columns", "basis_mflop", "basis_passes", "basis_mw", "basis_values", "basis_steps" (time steps
 * solved, summed over passes), "basis_delays" (mortar slabs after which the response is zero), "basis_shared_by",
 * "toeplitz_tiles", "toeplitz_mflop", "toeplitz_mbytes" (per operator application); -1 for an unknown key.        */
int64_t mmmfe_stat(const mmmfe_ctx *ctx, const char *key);

/* Measured peak of the FP64 tensor cores (register-resident mma.sync.m8n8k4.f64 on all SMs), TFLOP/s: the roofline
 * denominator of the factorisation and of the multiscale-basis kernels (MEASURED_PEAKS.json has no FP64 figure).   */
int mmmfe_measure_fp64_tensor_tflops(int device, double *tflops);

/* ---- diagnostics (host only, no device needed) ------------------------------------------------------------------- */
/* The nested-dissection tree the engine builds for an nx x ny RT0 grid, in canonical flux numbering.
 * sizes = {n_nodes, idx_len, src_len, r_total, z_total, n_heights}.
 * node_table: n_nodes x 8 int32 = {s, b, child0, child1, height, nsplit, ldr, parent};
 * offsets:    n_nodes x 4 int64 = {idx_off, src_off, r_off, z_off}.                                            */
/* Device milliseconds of ONE block-Toeplitz product of all local subdomains (the kernel of an operator application on
 * the multiscale basis; exchange and Arnoldi excluded), averaged over `repeat` back-to-back launches.               */
int mmmfe_profile_basis_apply(mmmfe_ctx *ctx, int32_t repeat, double *ms_per_launch);
/* The multiscale-basis operator of a local subdomain's factor group: info = {mortar time slabs m_t, rows = columns
 * P_tot, leading dimension ldt, slots, slot of side 0..3 (-1: none), first row/column of side 0..3 (-1)}; t_out (may be
 * NULL) receives T[d][row][ldt], d = 0..m_t-1: signed mortar response on `row` d slabs after the basis function `col`. */
int mmmfe_debug_get_basis(mmmfe_ctx *ctx, int32_t local_index, int64_t info[12], double *t_out, int64_t capacity);
/* Copies the stored factor (R blocks, layout given by the nd-tree offsets/ldr) of a local subdomain's group. */
int mmmfe_debug_get_factor(mmmfe_ctx *ctx, int32_t local_index, double *r_out, int64_t capacity);
int mmmfe_nd_tree_sizes(int32_t nx, int32_t ny, int32_t leaf_cells, int64_t sizes[6]);
/* Builds the static schedule of the persistent sweep kernel for an nx x ny grid (n_cta CTAs, M right-hand sides,
 * smem_bytes of dynamic shared memory per CTA) and replays n_steps of it on the CPU against the kernel's ring,
 * barrier and completion-counter rules.  0 = every entry completes and no ring range is overwritten early;
 * negative = failure with a message.  stats = {sweep nodes, top fronts, entries, factor doubles, ring doubles,
 * workspace doubles, largest entry doubles, shared memory bytes}.                                             */
int mmmfe_sweep_plan_check(int32_t nx, int32_t ny, int32_t leaf_cells, int32_t subtree_cells, int32_t M,
                           int32_t n_cta, int64_t smem_bytes, int32_t n_steps, int32_t compute_threads,
                           int32_t chunk_doubles, int64_t stats[8], char *msg, int32_t msg_cap);
int mmmfe_nd_tree_fill(int32_t nx, int32_t ny, int32_t leaf_cells, int32_t *node_table, int64_t *offsets,
                       int32_t *idx, int32_t *src);
/* Host-only (no device): the plan of the mortar-face exchange that mmmfe_setup builds for a rank -- replaces the
 * per-side MPI_Sendrecv of src/darcy_vtdd.cc:1390-1401.  Inputs: the interface sides of the rank's subdomains
 * (global id, side, neighbour's global id, offset and length in the rank's interface vector) and the owner rank of
 * every global subdomain.  Outputs: partner[i] >= 0 local partner entry, <= -2 entry -(v + 2) of the receive buffer;
 * send_index = interface entries in send-buffer order; per peer (ascending rank) the slices of the packed buffers.
 * Sender and receiver order a peer's faces by (global id, side) of the sending subdomain.                       */
int mmmfe_exchange_plan(int32_t n_sides, const int32_t *gid, const int32_t *side, const int32_t *neighbor,
                        const int64_t *off, const int64_t *len, int64_t n_iface, int32_t n_global,
                        const int32_t *owner, int32_t rank, int32_t world, int64_t *partner, int64_t *send_index,
                        int64_t send_cap, int64_t *n_send, int32_t *peer_rank, int64_t *peer_send_off,
                        int64_t *peer_recv_off, int64_t *peer_count, int32_t peer_cap, int32_t *n_peers);
/* Host-only: the front-end's ownership rule -- contiguous blocks of subdomains in the reference's rank order
 * (README.md:64), balanced by cost[r] = nx * ny * nt.                                                           */
int mmmfe_assign_owners(int32_t n_global, const double *cost, int32_t world, int32_t *owner);

#ifdef __cplusplus
}
#endif
#endif /* MMMFE_B200_H */
