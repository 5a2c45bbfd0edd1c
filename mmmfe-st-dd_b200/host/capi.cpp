// C ABI of the host front-end (include/mmmfe_host.h): the body is the reference's main()
// (src/darcy_main.cc:23-120) without MPI.
#include <algorithm>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "darcy_vt.h"
#include "mmmfe_host.h"

namespace {
std::string g_error;
std::vector<double> g_residuals, g_lambda;
}  // namespace

extern "C" {

void mmmfe_host_default_options(mmmfe_host_options *o) {
  if (!o) return;
  std::memset(o, 0, sizeof *o);
  o->world = 1; o->device = -1; o->compute_errors = 1; o->write_output = 1; o->final_sweep = 1; o->verbose = 1;
  o->only_cycle = -1; o->fixed_iterations = -1; o->num_refinement = 0; o->mortar_degree = -1;
}

int mmmfe_host_run(const char *parameter_file, const mmmfe_host_options *opt, mmmfe_host_report *reports,
                   int32_t max_reports) {
  try {
    using namespace vt_darcy;
    mmmfe_host_options o;
    if (opt) o = *opt; else mmmfe_host_default_options(&o);
    const std::string file = parameter_file ? parameter_file : "parameter.txt";
    double c_0, alpha, final_time, tolerence;
    int space_degree, mortar_degree, num_refinement, max_iteration;
    std::vector<std::vector<int>> mesh_m3d;
    std::vector<char> bc_con(4, 'D');
    std::vector<double> nm_bc_con_funcs(4, 0.0);
    bool is_manufact_solution, need_each_time_step_plot;
    const unsigned n_processes = count_subdomain_lines(file);
    parameter_pull_in(c_0, alpha, space_degree, mortar_degree, num_refinement, final_time, tolerence, max_iteration,
                      need_each_time_step_plot, bc_con, nm_bc_con_funcs, is_manufact_solution, mesh_m3d, n_processes, file);
    if (o.num_refinement > 0) num_refinement = o.num_refinement;
    if (o.mortar_degree >= 0) mortar_degree = o.mortar_degree;
    BiotParameters bparam(1.0, 1, final_time, c_0, alpha);
    DarcyVTProblem<2> problem_2d(space_degree, bparam, 1, mortar_degree, bc_con, nm_bc_con_funcs, is_manufact_solution,
                                 need_each_time_step_plot);
    RunControl &rc = problem_2d.control;
    rc.rank = o.rank; rc.world = o.world; rc.nccl_id = o.nccl_id; rc.device = o.device;
    rc.compute_errors = o.compute_errors != 0; rc.write_output = o.write_output != 0; rc.final_sweep = o.final_sweep != 0;
    rc.verbose = o.verbose != 0; rc.only_cycle = o.only_cycle; rc.fixed_iterations = o.fixed_iterations;
    rc.apply_repeat = o.apply_repeat; rc.keep_lambda = true;
    rc.warmup_iterations = o.warmup_iterations; rc.e2e_iterations = o.e2e_iterations; rc.profile = o.profile != 0;
    if (o.output_dir) rc.output_dir = o.output_dir;
    if (o.engine_options) rc.engine_options = o.engine_options;
    rc.also_solve = o.also_solve != 0; rc.fetch_solution = o.skip_fetch == 0;
    problem_2d.run(num_refinement, mesh_m3d, tolerence, max_iteration, mortar_degree + 1);
    const auto &reps = problem_2d.reports();
    int n = 0;
    for (const auto &r : reps) {
      if (reports && n < max_reports) {
        mmmfe_host_report &d = reports[n];
        d.cycle = r.cycle; d.gmres_iterations = r.gmres_iterations; d.converged = r.converged; d.factor_groups = r.factor_groups;
        d.r_norm = r.r_norm; d.velocity_l2l2 = r.velocity_l2l2; d.pressure_dg = r.pressure_dg;
        d.pressure_l2l2 = r.pressure_l2l2; d.lambda_l2 = r.lambda_l2;
        d.t_assemble = r.t_assemble; d.t_factor = r.t_factor; d.t_bar_data = r.t_bar_data; d.t_bar = r.t_bar;
        d.t_gmres = r.t_gmres; d.t_final = r.t_final; d.t_errors = r.t_errors; d.t_output = r.t_output;
        d.gmres_device_ms = r.gmres_device_ms; d.apply_ms = rc.apply_ms; d.factor_flops = r.factor_flops;
        d.interface_length = r.interface_length; d.dof_steps = r.dof_steps; d.step_bytes = r.step_bytes;
        d.factor_values = r.factor_values; d.kernel_launches = r.kernel_launches;
        d.e2e_ms_per_iteration = r.e2e_ms_per_iteration; d.e2e_h2d_bytes = r.e2e_h2d_bytes; d.e2e_d2h_bytes = r.e2e_d2h_bytes;
        d.gmres_kernel_launches = r.gmres_kernel_launches;
        for (int k = 0; k < 6; ++k) { d.prof_ms[k] = r.prof_ms[k]; d.prof_bytes[k] = r.prof_bytes[k]; d.prof_launches[k] = r.prof_launches[k]; }
        d.sweep_batches = r.sweep_batches;
        for (int k = 0; k < 4; ++k) { d.sweep_steps[k] = r.sweep_steps[k]; d.sweep_ms[k] = r.sweep_ms[k]; d.sweep_bytes_per_step[k] = r.sweep_bytes_per_step[k]; }
        d.tune_us_sweep = r.tune_us_sweep; d.tune_us_levels = r.tune_us_levels; d.sweep_planned = r.sweep_planned;
        d.tune_rel_diff_e18 = r.tune_rel_diff_e18;
        d.basis_operator = r.basis_operator; d.basis_shared_by = r.basis_shared_by; d.t_basis = r.t_basis; d.basis_gflop = r.basis_gflop;
        d.basis_bytes = r.basis_bytes; d.basis_columns = r.basis_columns;
        d.solve_iterations = r.solve_iterations; d.solve_converged = r.solve_converged;
        d.t_solve_gmres = r.t_solve_gmres; d.solve_final_residual = r.solve_final_residual;
        d.toeplitz_ms = r.toeplitz_ms; d.toeplitz_gflop = r.toeplitz_gflop; d.toeplitz_mbytes = r.toeplitz_mbytes;
        d.basis_steps = r.basis_steps; d.basis_delays = r.basis_delays; d.basis_mw = r.basis_mw; d.toeplitz_tiles = r.toeplitz_tiles;
      }
      ++n;
    }
    if (!reps.empty()) { g_residuals = reps.back().residuals; g_lambda = reps.back().lambda; }
    return n;
  } catch (const std::exception &e) {
    g_error = e.what();
    return -1;
  } catch (...) {
    g_error = "unknown exception";
    return -1;
  }
}

int mmmfe_host_last_residuals(double *out, int32_t cap) {
  const int n = int(std::min<size_t>(g_residuals.size(), size_t(cap < 0 ? 0 : cap)));
  if (out) std::memcpy(out, g_residuals.data(), size_t(n) * sizeof(double));
  return int(g_residuals.size());
}

int64_t mmmfe_host_last_lambda(double *out, int64_t cap) {
  const size_t n = std::min<size_t>(g_lambda.size(), size_t(cap < 0 ? 0 : cap));
  if (out) std::memcpy(out, g_lambda.data(), n * sizeof(double));
  return int64_t(g_lambda.size());
}

const char *mmmfe_host_last_error(void) { return g_error.c_str(); }

int mmmfe_host_dump_subdomain(const char *parameter_file, int32_t cycle, int32_t r, int32_t mortar_degree_override,
                              const char *dir) {
  try {
    using namespace vt_darcy;
    const std::string file = parameter_file ? parameter_file : "parameter.txt";
    double c_0, alpha, final_time, tolerence;
    int space_degree, mortar_degree, num_refinement, max_iteration;
    std::vector<std::vector<int>> reps;
    std::vector<char> bc_con(4, 'D');
    std::vector<double> bcv(4, 0.0);
    bool manufact, plots;
    const unsigned n = count_subdomain_lines(file);
    parameter_pull_in(c_0, alpha, space_degree, mortar_degree, num_refinement, final_time, tolerence, max_iteration, plots,
                      bc_con, bcv, manufact, reps, n, file);
    if (mortar_degree_override >= 0) mortar_degree = mortar_degree_override;
    for (int c = 1; c <= cycle; ++c) {  // src/darcy_vtdd.cc:2124-2142
      for (size_t i = 0; i + 1 < reps.size(); ++i) for (int d = 0; d < 3; ++d) reps[i][d] *= 2;
      if (mortar_degree == 1 || c % 2 == 0) for (int d = 0; d < 3; ++d) reps.back()[d] *= 2;
    }
    if (r < 0 || r >= int(n)) throw std::runtime_error("subdomain index out of range");
    dump_subdomain(reps, r, final_time, c_0, alpha, mortar_degree, manufact, bc_con, bcv, dir ? dir : ".");
    return 0;
  } catch (const std::exception &e) {
    g_error = e.what();
    return -1;
  }
}

}  // extern "C"
