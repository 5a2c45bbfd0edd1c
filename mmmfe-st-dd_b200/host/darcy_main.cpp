// DarcyVT: drop-in for the reference executable (src/darcy_main.cc:23-120).  Run from a directory that holds
// parameter.txt, space-time-plots/ and time-step-plots/.  No arguments are needed; the number of subdomains is the
// number of mesh_pattern_sub_d* lines (the reference takes it from `mpirun -n j`).  Under torchrun/mpirun-style
// launchers RANK, WORLD_SIZE and LOCAL_RANK select one GPU per process; the NCCL id travels through the file named
// by $MMMFE_NCCL_ID_FILE (default ./.mmmfe_nccl_id.<MASTER_PORT>).  A file left behind by a crashed launch is never
// consumed: rank 0 removes it before writing, the other ranks only accept a file written after they started.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <thread>
#include <vector>

#include <sys/stat.h>

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
    std::string id_path;
    if (world > 1) {
      const char *p = std::getenv("MMMFE_NCCL_ID_FILE");
      const char *port = std::getenv("MASTER_PORT");
      id_path = p ? p : std::string(".mmmfe_nccl_id.") + (port ? port : "0");
      const auto started = std::chrono::system_clock::to_time_t(std::chrono::system_clock::now());
      if (rank == 0) {
        std::remove(id_path.c_str());
        if (mmmfe_comm_unique_id(id.data()) != 0) throw std::runtime_error("NCCL is not available");
        std::ofstream tmp(id_path + ".tmp", std::ios::binary);
        tmp.write(id.data(), mmmfe_comm_unique_id_size());
        tmp.close();
        std::rename((id_path + ".tmp").c_str(), id_path.c_str());
      } else {
        for (int tries = 0;; ++tries) {
          struct stat sb;
          // launchers start the ranks within seconds of each other: anything older than that is a stale file
          if (::stat(id_path.c_str(), &sb) == 0 && sb.st_mtime + 20 >= started) {
            std::ifstream in(id_path, std::ios::binary);
            if (in && in.read(id.data(), mmmfe_comm_unique_id_size())) break;
          }
          if (tries > 600) throw std::runtime_error("timed out waiting for " + id_path);
          std::this_thread::sleep_for(std::chrono::milliseconds(100));
        }
      }
      problem_2d.control.nccl_id = id.data();
    }
    struct IdFileGuard {  // rank 0 removes the id file however run() ends
      std::string path; bool mine;
      ~IdFileGuard() { if (mine && !path.empty()) std::remove(path.c_str()); }
    } id_guard{id_path, rank == 0};
    problem_2d.run(num_refinement, mesh_m3d, tolerence, max_iteration, mortar_degree + 1);
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
