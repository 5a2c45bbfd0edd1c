// Device evaluation of the problem data (inc/data.h function classes) for the host front-end.
//
// The reference evaluates RightHandSidePressure, PressureBoundaryValues, InitialCondition and ExactSolution on the
// host inside its cell loops (assemble_rhs_bar, src/darcy_vtdd.cc:632-766; compute_errors, :1657-1775; output_results,
// :1837-1998).  On the large configurations those loops are O(cells x time levels) host passes of several seconds.
// data_device.cu compiles the SAME, unmodified data header for sm_100a (the virtual member functions of the classes
// become __host__ __device__; objects are constructed on the host and their bytes are handed to the kernels, which
// call the overriders with qualified names, so no device-side vtable is needed) and fills / reduces the engine's
// device buffers directly.  When the header cannot be compiled for the device (build.py tries), the front-end is
// built with MMMFE_NO_DEVICE_DATA and keeps the host loops.
#pragma once
#include <cstdint>

namespace vt_darcy {
namespace devdata {

struct Grid {            // one subdomain (host/darcy_vt.cpp Sub)
  int nx, ny, nt;
  double x0, y0, hx, hy, dt;
  const double *levels_t;  // HOST [nt]: accumulated times (t += dt, src/darcy_vtdd.cc:887)
};

// false when this build has no device data path (MMMFE_NO_DEVICE_DATA) or MMMFE_HOST_DATA=1 is set
bool enabled();

// VectorTools::project(InitialCondition, QGauss(degree + 5)) onto DGQ0 = cell mean with 5x5 Gauss points
// (src/darcy_vtdd.cc:2172-2176): d_p0[n_pressure]
void initial_pressure(const Grid &g, double *d_p0);
// dt (f(t_n), w) per cell with 2x2 Gauss points (src/darcy_vtdd.cc:718): d_src[nt][n_pressure] +=
void bar_source(const Grid &g, double c0, double alpha, double *d_src);
// outer Dirichlet rows -n * int_e g(t_n) with QGauss<1>(qdegree) (src/darcy_vtdd.cc:732-755):
// d_fr_vals[nt][nfr] += for the first n_dir entries; fr_side[e] = side << 24 | position (HOST);
// manufact = 0: the constant bc_val_of_side[side] replaces the function
void boundary_values(const Grid &g, double Lx, double Ly, int n_dir, const int32_t *fr_side, int nfr, int qdegree,
                     int manufact, const double bc_val_of_side[4], double *d_fr_vals);
// d_fr_vals[lvl][slot[k]] += value[k] and d_src[lvl][cell[k]] += value[k] for every level (constant rows of
// system_rhs_bar_bc, src/darcy_vtdd.cc:863)
void add_constant_rows(int nt, int64_t stride, int n, const int32_t *index, const double *value, double *d_rows);
// constraint_bc.distribute (src/darcy_vtdd.cc:866): d_solution[lvl][dof[k]] = value[k]
void set_constant_rows(int nt, int64_t stride, int n, const int32_t *index, const double *value, double *d_rows);

// compute_errors of one subdomain (src/darcy_vtdd.cc:1657-1775) from the device-resident solution [nt][n]:
// per level {velocity num, velocity den, pressure num, pressure den, DG jump} -> out[nt][5] (HOST)
void error_sums(const Grid &g, const double *d_solution, double *out);

// output_results (src/darcy_vtdd.cc:1906-1948): the space-time patches of levels [l0, l1) expanded to their 8
// vertices, into HOST buffers: points[cells*8][3], u[cells*8][3], p[cells*8]  (cells = nx*ny*(l1-l0))
void expand_space_time(const Grid &g, const double *d_solution, int l0, int l1, double *points, double *u, double *p);

}  // namespace devdata
}  // namespace vt_darcy
