// DarcyVT: drop-in for the reference executable (src/darcy_main.cc:23-120).  Run from a directory that holds
// parameter.txt, space-time-plots/ and time-step-plots/.  No arguments are needed; the number of subdomains is the
// number of mesh_pattern_sub_d* lines (the reference takes it from `mpirun -n j`).  Under torchrun/mpirun-style
// launchers RANK, WORLD_SIZE and LOCAL_RANK select one GPU per process; the NCCL id travels through the file named
// by $MMMFE_NCCL_ID_FILE (default ./.mmmfe_nccl_id).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <thread>
#include <vector>

#include "darcy_vt.h"
#include "mmmfe_b200.h"

static int env_int(const char *k, int d) {
  const char *v = std::getenv(k);
  return v ? std::atoi(v) : d;
}

int main(int argc, char *argv[]) {
  try {
    using namespace vt_darcy;
    const std::string file = argc > 1 ? argv[1] : "parameter.txt";
    double c_0, alpha, final_time, tolerence;
    int space_degree, mortar_degree, num_refinement, max_iteration;
    std::vector<std::vector<int>> mesh_m3d;
    std::vector<char> bc_con(4, 'D');
    std::vector<double> nm_bc_con_funcs(4, 0.0);
    bool is_manufact_solution, need_each_time_step_plot;
    const unsigned n_processes = count_subdomain_lines(file);
    parameter_pull_in(c_0, alpha, space_degree, mortar_degree, num_refinement, final_time, tolerence, max_iteration,
                      need_each_time_step_plot, bc_con, nm_bc_con_funcs, is_manufact_solution, mesh_m3d, n_processes, file);

    BiotParameters bparam(1.0, 1, final_time, c_0, alpha);
    DarcyVTProblem<2> problem_2d(space_degree, bparam, 1, mortar_degree, bc_con, nm_bc_con_funcs, is_manufact_solution,
                                 need_each_time_step_plot);
    const int rank = env_int("RANK", 0), world = env_int("WORLD_SIZE", 1);
    std::vector<char> id(size_t(mmmfe_comm_unique_id_size()) + 1);
    problem_2d.control.rank = rank;
    problem_2d.control.world = world;
    problem_2d.control.device = env_int("LOCAL_RANK", -1);
    if (world > 1) {
      const char *p = std::getenv("MMMFE_NCCL_ID_FILE");
      const std::string path = p ? p : ".mmmfe_nccl_id";
      if (rank == 0) {
        if (mmmfe_comm_unique_id(id.data()) != 0) throw std::runtime_error("NCCL is not available");
        std::ofstream tmp(path + ".tmp", std::ios::binary);
        tmp.write(id.data(), mmmfe_comm_unique_id_size());
        tmp.close();
        std::rename((path + ".tmp").c_str(), path.c_str());
      } else {
        for (int tries = 0;; ++tries) {
          std::ifstream in(path, std::ios::binary);
          if (in && in.read(id.data(), mmmfe_comm_unique_id_size())) break;
          if (tries > 600) throw std::runtime_error("timed out waiting for " + path);
          std::this_thread::sleep_for(std::chrono::milliseconds(100));
        }
      }
      problem_2d.control.nccl_id = id.data();
    }
    problem_2d.run(num_refinement, mesh_m3d, tolerence, max_iteration, mortar_degree + 1);
    if (world > 1 && rank == 0) {
      const char *p = std::getenv("MMMFE_NCCL_ID_FILE");
      std::remove(p ? p : ".mmmfe_nccl_id");
    }
  } catch (std::exception &exc) {
    std::cerr << std::endl << std::endl << "----------------------------------------------------" << std::endl;
    std::cerr << "Exception on processing: " << std::endl << exc.what() << std::endl << "Aborting!" << std::endl
              << "----------------------------------------------------" << std::endl;
    return 1;
  } catch (...) {
    std::cerr << std::endl << std::endl << "----------------------------------------------------" << std::endl;
    std::cerr << "Unknown exception!" << std::endl << "Aborting!" << std::endl
              << "----------------------------------------------------" << std::endl;
    return 1;
  }
  return 0;
}
