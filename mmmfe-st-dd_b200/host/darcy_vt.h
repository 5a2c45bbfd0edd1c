// Host front-end of the B200 build of DarcyVT: mirrors the public surface of the reference's
// DarcyVTProblem<2> (inc/darcy_vtdd.h:90-101), BiotParameters (inc/darcy_vtdd.h:30-51) and
// parameter_pull_in (inc/filecheck_utility.h:34-95).  It does the one-time, host-side work of a refinement
// cycle -- structured RT0 x DGQ0 meshes, saddle-matrix assembly, right-hand-side data from the data.h classes,
// projector tables, error norms, output -- and drives the CUDA engine (include/mmmfe_b200.h) for everything on
// the interface-GMRES path.  One process drives all subdomains it owns; ranks replace "one MPI rank per
// subdomain".
#ifndef MMMFE_HOST_DARCY_VT_H
#define MMMFE_HOST_DARCY_VT_H

#include <functional>
#include <string>
#include <vector>

namespace vt_darcy {

struct BiotParameters {
  BiotParameters(const double dt, const unsigned int nt, const double f_time, const double c = 1.0,
                 const double a = 1.0)
      : time(0.0), time_step(dt), final_time(f_time), num_time_steps(nt), c_0(c), alpha(a) {}
  mutable double time;
  double time_step;
  const double final_time;
  unsigned int num_time_steps;
  const double c_0;
  const double alpha;
};

// What one refinement cycle produced (the reference prints these; harnesses read them).
struct CycleReport {
  int cycle = 0;
  int gmres_iterations = 0;  // "# GMRES"
  int converged = 0;
  double r_norm = 0.0;
  std::vector<double> residuals;
  double velocity_l2l2 = 0.0, pressure_dg = 0.0, pressure_l2l2 = 0.0, lambda_l2 = 0.0;  // convergence-table columns
  // wall-clock seconds of the phases
  double t_assemble = 0.0, t_factor = 0.0, t_bar_data = 0.0, t_bar = 0.0, t_gmres = 0.0, t_final = 0.0,
         t_errors = 0.0, t_output = 0.0;
  double gmres_device_ms = 0.0;
  long long interface_length = 0, dof_steps = 0, step_bytes = 0, factor_values = 0, kernel_launches = 0;
  double factor_flops = 0.0;
  int factor_groups = 0;
  std::vector<double> lambda;  // local interface vector
  double e2e_ms_per_iteration = 0.0;
  long long e2e_h2d_bytes = 0, e2e_d2h_bytes = 0, gmres_kernel_launches = 0;
  double prof_ms[6] = {0, 0, 0, 0, 0, 0}, prof_bytes[6] = {0, 0, 0, 0, 0, 0};
  long long prof_launches[6] = {0, 0, 0, 0, 0, 0};
  // persistent sweep kernel (when the engine selected it): per batch one timed star sweep
  int sweep_batches = 0, sweep_steps[4] = {0, 0, 0, 0};
  double sweep_ms[4] = {0, 0, 0, 0}, sweep_bytes_per_step[4] = {0, 0, 0, 0};
  long long tune_us_sweep = 0, tune_us_levels = 0, sweep_planned = 0;
  // converged solve after a fixed-iteration timing run (RunControl::also_solve)
  int solve_iterations = 0, solve_converged = 0;
  double t_solve_gmres = 0.0, solve_final_residual = 0.0;
  double toeplitz_ms = 0.0, toeplitz_gflop = 0.0, toeplitz_mbytes = 0.0;
  long long basis_steps = 0, basis_delays = 0, basis_mw = 0, toeplitz_tiles = 0;
  long long tune_rel_diff_e18 = -1;  // set-up cross-check of the two star-sweep implementations (relative l2 * 1e18)
  // multiscale-basis (time-Toeplitz) operator, when the engine built it
  int basis_operator = 0, basis_shared_by = 1;
  double t_basis = 0.0, basis_gflop = 0.0, basis_bytes = 0.0;
  long long basis_columns = 0;
};

struct RunControl {
  int rank = 0, world = 1;
  const void *nccl_id = nullptr;  // required when world > 1 (else read from $MMMFE_NCCL_ID_FILE)
  int device = -1;
  bool compute_errors = true;   // is_manufact_soln only
  bool write_output = true;     // .vtu/.pvtu/.tex files
  bool final_sweep = true;
  bool verbose = true;          // rank-0 log lines of the reference
  bool keep_lambda = false;
  int only_cycle = -1;          // run just this refinement index (meshes refined as usual)
  int fixed_iterations = -1;    // >= 0: tolerance 0 and exactly this many GMRES iterations (benchmarks)
  int warmup_iterations = 0;    // with fixed_iterations: an untimed GMRES run of this many iterations first
  int e2e_iterations = 0;       // > 0: host-driven iterations through mmmfe_apply_interface_operator, timed
  bool profile = false;         // per-kernel-class timing of one star step
  int apply_repeat = 0;         // > 0: additionally time this many device-resident operator applications
  double apply_ms = 0.0;        // out
  std::string output_dir = ".";
  // engine options "key=value,key=value" passed to mmmfe_set_option before set-up (after those of the environment
  // variable MMMFE_ENGINE_OPTIONS, same format), e.g. "use_sweep=1,sweep_autotune=0"
  std::string engine_options;
  // with fixed_iterations >= 0: afterwards also run the converged solve (tolerance and max_iteration of the parameter
  // file) in the same context and report its iteration count and time (time-to-solution of the benchmarks)
  bool also_solve = false;
  // false: the final sweep runs on the device but the [nt][n] solutions are not copied to the host (no error norms, no
  // output files): timing runs of problems whose full space-time solution is tens of GB
  bool fetch_solution = true;
};

void parameter_pull_in(double &c_0, double &alpha, int &space_degree, int &mortar_degree, int &num_refinement,
                       double &final_time, double &tolerence, int &max_iteration, bool &need_each_time_step_plot,
                       std::vector<char> &bc_con, std::vector<double> &nm_bc_con_funcs, bool &is_manufact_solution,
                       std::vector<std::vector<int>> &mesh_m3d, unsigned int n_processes,
                       std::string file_name = "parameter.txt");
// Number of mesh_pattern_sub_d* lines: stands in for the MPI process count.
unsigned int count_subdomain_lines(const std::string &file_name);

void dump_subdomain(const std::vector<std::vector<int>> &reps, int r, double final_time, double c0, double alpha,
                    int mortar_degree, bool manufact, const std::vector<char> &bc_type,
                    const std::vector<double> &bc_val, const std::string &dir);

template <int dim = 2>
class DarcyVTProblem {
  static_assert(dim == 2, "only DarcyVTProblem<2> is instantiated by the reference (src/darcy_vtdd.cc:2207)");

 public:
  DarcyVTProblem(const unsigned int degree, const BiotParameters &bprm, const unsigned int mortar_flag = 0,
                 const unsigned int mortar_degree = 0, std::vector<char> bc_condition_vect = {'D', 'D', 'D', 'D'},
                 std::vector<double> bc_const_functs = {0., 0., 0., 0.}, const bool is_manufact_soln = true,
                 const bool need_each_time_step_plot = false);
  void run(const unsigned int refine, const std::vector<std::vector<int>> &reps_st, double tol, unsigned int maxiter,
           unsigned int quad_degree = 3);

  RunControl control;
  const std::vector<CycleReport> &reports() const { return reports_; }

 private:
  BiotParameters prm;
  std::vector<char> bc_condition_vect;
  std::vector<double> bc_const_functs;
  bool is_manufact_solution, need_each_time_step_plot;
  unsigned int degree, mortar_degree, mortar_flag;
  std::vector<CycleReport> reports_;
};

extern template class DarcyVTProblem<2>;

}  // namespace vt_darcy
#endif
