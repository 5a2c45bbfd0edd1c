/* mmmfe_host.h -- C ABI of the C++ host front-end (mmmfe-st-dd_b200/host), for harnesses that are not C++.
 *
 * The front-end is the B200 build of the reference program: it reads the reference's positional parameter.txt
 * (inc/filecheck_utility.h:34-95), builds DarcyVTProblem<2> (inc/darcy_vtdd.h:94-101) and calls run()
 * (src/darcy_vtdd.cc:2029-2204) exactly as src/darcy_main.cc:23-120 does; the number of subdomains is the number
 * of mesh_pattern_sub_d* lines instead of the MPI process count.  All pointers are host memory.
 */
#ifndef MMMFE_HOST_H
#define MMMFE_HOST_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int32_t rank, world;       /* one process per GPU; subdomains are sharded over ranks                       */
  const void *nccl_id;       /* mmmfe_comm_unique_id blob of rank 0 (world > 1)                               */
  int32_t device;            /* CUDA device, -1 = current                                                     */
  int32_t compute_errors;    /* convergence-table columns (is_manufact_soln only)                            */
  int32_t write_output;      /* .vtu/.pvtu/.tex files under output_dir                                       */
  int32_t final_sweep;       /* solve_timestep(2, .) pass                                                     */
  int32_t verbose;           /* the reference's rank-0 log lines on stdout                                    */
  int32_t only_cycle;        /* -1 = all refinement cycles                                                    */
  int32_t fixed_iterations;  /* >= 0: tolerance 0 and exactly this many GMRES iterations (benchmarks)        */
  int32_t apply_repeat;      /* > 0: time this many device-resident operator applications                    */
  int32_t num_refinement;    /* > 0 overrides the file's num_refinement                                       */
  int32_t mortar_degree;     /* >= 0 overrides the file's mortar_degree                                       */
  const char *output_dir;    /* NULL = "."                                                                    */
  int32_t warmup_iterations; /* with fixed_iterations: untimed GMRES iterations run first                     */
  int32_t e2e_iterations;    /* > 0: also time this many host-driven iterations (pinned host q -> device ->   */
                             /* host Ap, classical Gram-Schmidt on the host)                                   */
  int32_t profile;           /* != 0: fill prof_* with mmmfe_profile_step                                      */
  const char *engine_options; /* NULL or "key=value,key=value" for mmmfe_set_option (before set-up)           */
  int32_t also_solve;        /* with fixed_iterations: then also the converged solve (file tolerance), timed      */
  int32_t skip_fetch;        /* != 0: final sweep on the device only, solutions stay there (timing of big runs)   */
} mmmfe_host_options;

typedef struct {
  int32_t cycle, gmres_iterations, converged, factor_groups;
  double r_norm;
  double velocity_l2l2, pressure_dg, pressure_l2l2, lambda_l2;
  double t_assemble, t_factor, t_bar_data, t_bar, t_gmres, t_final, t_errors, t_output;
  double gmres_device_ms, apply_ms, factor_flops;
  int64_t interface_length, dof_steps, step_bytes, factor_values, kernel_launches;
  double e2e_ms_per_iteration;
  int64_t e2e_h2d_bytes, e2e_d2h_bytes; /* per iteration                                                       */
  int64_t gmres_kernel_launches;        /* kernels launched inside the timed GMRES loop                        */
  double prof_ms[6], prof_bytes[6];
  int64_t prof_launches[6];
  /* persistent sweep kernel: batches it runs (0: the per-level kernels were selected), one timed star sweep each */
  int32_t sweep_batches, sweep_steps[4];
  double sweep_ms[4], sweep_bytes_per_step[4];
  int64_t tune_us_sweep, tune_us_levels, sweep_planned;
  int64_t tune_rel_diff_e18; /* set-up cross-check: relative l2 difference of the two sweep implementations * 1e18, -1 = not run */
  /* multiscale-basis (time-Toeplitz) interface operator: 1 when GMRES ran on it, its precompute cost                */
  int32_t basis_operator, basis_shared_by; /* ranks that split the basis columns of a group among themselves */
  double t_basis, basis_gflop, basis_bytes;
  int64_t basis_columns;
  /* converged solve after the fixed-iteration timing (also_solve)                                                     */
  int32_t solve_iterations, solve_converged;
  double t_solve_gmres, solve_final_residual;
  /* operator on the multiscale basis (profile != 0): one block-Toeplitz product of all local subdomains              */
  double toeplitz_ms, toeplitz_gflop, toeplitz_mbytes;
  int64_t basis_steps, basis_delays, basis_mw, toeplitz_tiles;
} mmmfe_host_report;

void mmmfe_host_default_options(mmmfe_host_options *opt);
/* Runs the whole program on `parameter_file`.  Fills up to max_reports per-cycle reports, returns the number of
 * cycles run (>= 0) or -1 on failure (message via mmmfe_host_last_error).                                      */
int mmmfe_host_run(const char *parameter_file, const mmmfe_host_options *opt, mmmfe_host_report *reports,
                   int32_t max_reports);
/* Residual history / lambda of the last cycle of the last run (copied into the caller's buffers). */
int mmmfe_host_last_residuals(double *out, int32_t cap);
int64_t mmmfe_host_last_lambda(double *out, int64_t cap);
const char *mmmfe_host_last_error(void);
/* Host-only diagnostic (no device): writes the saddle matrix, projector tables and bar data the front-end builds
 * for subdomain r of refinement cycle `cycle` as raw little-endian arrays into `dir`.                          */
int mmmfe_host_dump_subdomain(const char *parameter_file, int32_t cycle, int32_t r, int32_t mortar_degree,
                              const char *dir);

#ifdef __cplusplus
}
#endif
#endif
