// Device evaluation of the problem data: see data_device.h.  Compiled by nvcc for sm_100a with -fmad=false (the
// quadrature-point coordinates and sums are then rounded like the host loops' in darcy_vt.cpp; the transcendental
// functions differ from glibc's in the last bit or two).
//
// How an unmodified inc/data.h reaches the device: every standard header is included FIRST; then `virtual` is
// defined as `__host__ __device__ virtual` while the deal.II shim and the data header are read, which makes exactly
// the overriders (value, vector_value, ...) of the function classes host-device -- nvcc applies the execution space
// of the in-class declaration to the out-of-class definitions the reference's headers use.  Constructors stay
// host-only: the objects are built on the host, their bytes travel as kernel parameters, and the kernels call the
// overriders with qualified names (static dispatch: the host vtable pointer inside the bytes is never used).
#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <complex>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <numeric>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include <cuda_runtime.h>

#include "data_device.h"

#ifndef MMMFE_NO_DEVICE_DATA

#define virtual __host__ __device__ virtual
#include <deal.II/base/function.h>
#include <deal.II/base/tensor_function.h>
#ifndef MMMFE_DATA_HEADER
#define MMMFE_DATA_HEADER "data_default.h"
#endif
#include MMMFE_DATA_HEADER
#undef virtual

namespace vt_darcy {
namespace devdata {

namespace {

void cu(cudaError_t e, const char *what) {
  if (e != cudaSuccess) throw std::runtime_error(std::string("data_device: ") + what + ": " + cudaGetErrorString(e));
}

// the bytes of a host-constructed function object
template <class T>
struct alignas(16) Pod {
  unsigned char b[(sizeof(T) + 15) / 16 * 16];
  __device__ T *get() { return reinterpret_cast<T *>(b); }
};
template <class T>
Pod<T> pod_of(const T &t) {
  Pod<T> r;
  std::memset(r.b, 0, sizeof r.b);
  std::memcpy(r.b, static_cast<const void *>(&t), sizeof(T));
  return r;
}

struct DevGrid {
  int nx, ny, nt;
  double x0, y0, hx, hy, dt;
  const double *t;  // device [nt]
};

struct Levels {  // device copy of levels_t
  double *p = nullptr;
  explicit Levels(const Grid &g) {
    cu(cudaMalloc(&p, sizeof(double) * size_t(g.nt)), "cudaMalloc");
    cu(cudaMemcpy(p, g.levels_t, sizeof(double) * size_t(g.nt), cudaMemcpyHostToDevice), "cudaMemcpy");
  }
  ~Levels() { cudaFree(p); }
  DevGrid grid(const Grid &g) const { return DevGrid{g.nx, g.ny, g.nt, g.x0, g.y0, g.hx, g.hy, g.dt, p}; }
};

struct Quad { double x[8], w[8]; };  // QGauss<1>(n) on [0, 1], n <= 8

// same iteration as gauss01 of darcy_vt.cpp (bitwise the same nodes and weights)
Quad gauss01(int n) {
  if (n > 8) throw std::runtime_error("data_device: quadrature degree beyond 8");
  Quad q{};
  const double pi = std::acos(-1.0);
  for (int i = 0; i < n; ++i) {
    double z = std::cos(pi * (i + 0.75) / (n + 0.5)), pp = 1.0;
    for (int it = 0; it < 100; ++it) {
      double p1 = 1.0, p2 = 0.0;
      for (int j = 0; j < n; ++j) {
        const double p3 = p2;
        p2 = p1;
        p1 = ((2.0 * j + 1.0) * z * p2 - j * p3) / (j + 1.0);
      }
      pp = n * (z * p1 - p2) / (z * z - 1.0);
      const double dz = p1 / pp;
      z -= dz;
      if (std::fabs(dz) < 1e-16) break;
    }
    q.x[n - 1 - i] = 0.5 * (z + 1.0);
    q.w[n - 1 - i] = 1.0 / ((1.0 - z * z) * pp * pp);
  }
  return q;
}

__global__ void __launch_bounds__(256) k_initial_pressure(DevGrid g, Quad q5, Pod<InitialCondition<2>> icb, double *__restrict__ p0) {
  const int64_t c = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (c >= int64_t(g.nx) * g.ny) return;
  const int j = int(c / g.nx), i = int(c - int64_t(j) * g.nx);
  InitialCondition<2> *ic = icb.get();
  dealii::Vector<double> v(3);
  double acc = 0;
  for (int a = 0; a < 5; ++a)
    for (int b = 0; b < 5; ++b) {
      ic->InitialCondition<2>::vector_value(dealii::Point<2>(g.x0 + (i + q5.x[a]) * g.hx, g.y0 + (j + q5.x[b]) * g.hy), v);
      acc += q5.w[a] * q5.w[b] * v(2);
    }
  p0[c] = acc;
}

__global__ void __launch_bounds__(256) k_bar_source(DevGrid g, Quad q2, Pod<RightHandSidePressure<2>> fb, double *__restrict__ src) {
  const int64_t c = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  const int lvl = blockIdx.y;
  if (c >= int64_t(g.nx) * g.ny) return;
  const int j = int(c / g.nx), i = int(c - int64_t(j) * g.nx);
  Pod<RightHandSidePressure<2>> loc = fb;
  RightHandSidePressure<2> *f = loc.get();
  f->dealii::FunctionTime<double>::set_time(g.t[lvl]);
  const double area = g.hx * g.hy;
  double acc = 0;
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b)
      acc += q2.w[a] * q2.w[b] *
             f->RightHandSidePressure<2>::value(dealii::Point<2>(g.x0 + (i + q2.x[a]) * g.hx, g.y0 + (j + q2.x[b]) * g.hy), 0);
  src[int64_t(lvl) * g.nx * g.ny + c] += g.dt * acc * area;
}

struct BcArgs {
  double Lx, Ly;
  int n_dir, nfr, qdegree, manufact;
  double bc_val[4];
  const int32_t *fr_side;  // device
};

__global__ void __launch_bounds__(128) k_boundary_values(DevGrid g, BcArgs a, Quad q, Pod<PressureBoundaryValues<2>> gb,
                                                          double *__restrict__ fr_vals) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int lvl = blockIdx.y;
  if (e >= a.n_dir) return;
  Pod<PressureBoundaryValues<2>> loc = gb;
  PressureBoundaryValues<2> *gf = loc.get();
  gf->dealii::FunctionTime<double>::set_time(g.t[lvl]);
  const int side = a.fr_side[e] >> 24, m = a.fr_side[e] & ((1 << 24) - 1);
  const double normal = (side == 0 || side == 3) ? -1.0 : 1.0;  // bottom, right, top, left (inc/utilities.h:79-91)
  double acc = 0;
  for (int k = 0; k < a.qdegree; ++k) {
    double x, y;
    if (side == 0 || side == 2) { x = g.x0 + (m + q.x[k]) * g.hx; y = (side == 0) ? g.y0 : g.y0 + a.Ly; }
    else { y = g.y0 + (m + q.x[k]) * g.hy; x = (side == 3) ? g.x0 : g.x0 + a.Lx; }
    const double gv = a.manufact ? gf->PressureBoundaryValues<2>::value(dealii::Point<2>(x, y), 0) : a.bc_val[side];
    acc += q.w[k] * gv;
  }
  fr_vals[int64_t(lvl) * a.nfr + e] += -normal * acc;
}

__global__ void __launch_bounds__(256) k_rows(int nt, int64_t stride, int n, const int32_t *__restrict__ index,
                                               const double *__restrict__ value, double *__restrict__ rows, int add) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int lvl = blockIdx.y;
  if (k >= n || lvl >= nt) return;
  double *p = rows + int64_t(lvl) * stride + index[k];
  *p = add ? *p + value[k] : value[k];
}

void rows_op(int nt, int64_t stride, int n, const int32_t *index, const double *value, double *d_rows, int add) {
  if (n <= 0 || nt <= 0) return;
  int32_t *d_i = nullptr;
  double *d_v = nullptr;
  cu(cudaMalloc(&d_i, sizeof(int32_t) * size_t(n)), "cudaMalloc");
  cu(cudaMalloc(&d_v, sizeof(double) * size_t(n)), "cudaMalloc");
  cu(cudaMemcpy(d_i, index, sizeof(int32_t) * size_t(n), cudaMemcpyHostToDevice), "cudaMemcpy");
  cu(cudaMemcpy(d_v, value, sizeof(double) * size_t(n), cudaMemcpyHostToDevice), "cudaMemcpy");
  k_rows<<<dim3((n + 255) / 256, nt), 256>>>(nt, stride, n, d_i, d_v, d_rows, add);
  cu(cudaDeviceSynchronize(), "k_rows");
  cudaFree(d_i);
  cudaFree(d_v);
}

constexpr int ERR_CELLS = 1024;  // cells per block of the error kernel (256 threads x 4)

// One block: ERR_CELLS cells of one level.  Sums in a fixed order (thread-sequential, then a shared-memory tree):
// deterministic run to run.
__global__ void __launch_bounds__(256) k_error_sums(DevGrid g, Pod<ExactSolution<2>> exb, Pod<PressureBoundaryValues<2>> pb,
                                                     const double *__restrict__ sol, double *__restrict__ partial) {
  const int lvl = blockIdx.y;
  const int64_t np = int64_t(g.nx) * g.ny, nxe = int64_t(g.nx + 1) * g.ny, nf = nxe + int64_t(g.nx) * (g.ny + 1), n = nf + np;
  Pod<ExactSolution<2>> eloc = exb;
  ExactSolution<2> *ex = eloc.get();
  ex->dealii::FunctionTime<double>::set_time(g.t[lvl]);
  Pod<PressureBoundaryValues<2>> nloc = pb, oloc = pb;
  PressureBoundaryValues<2> *proj_now = nloc.get(), *proj_old = oloc.get();
  proj_now->dealii::FunctionTime<double>::set_time(g.t[lvl] + 0.5 * g.dt);  // src/darcy_vtdd.cc:1669-1670
  if (lvl > 0) proj_old->dealii::FunctionTime<double>::set_time(g.t[lvl - 1] + 0.5 * g.dt);
  const double *x = sol + int64_t(lvl) * n;
  const double *xo = lvl > 0 ? sol + int64_t(lvl - 1) * n : nullptr;
  const double tq[3] = {0.0, 0.5, 1.0}, tw[3] = {0.25, 0.5, 0.25};  // QIterated(QTrapezoid, 2)
  const double area = g.hx * g.hy;
  double acc[5] = {0, 0, 0, 0, 0};
  dealii::Vector<double> v(3);
  for (int r = 0; r < ERR_CELLS / 256; ++r) {
    const int64_t c = int64_t(blockIdx.x) * ERR_CELLS + r * 256 + threadIdx.x;
    if (c >= np) break;
    const int j = int(c / g.nx), i = int(c - int64_t(j) * g.nx);
    const double p = x[nf + c];
    const double uxl = x[int64_t(j) * (g.nx + 1) + i], uxr = x[int64_t(j) * (g.nx + 1) + i + 1];
    const double uyb = x[nxe + int64_t(j) * g.nx + i], uyt = x[nxe + int64_t(j + 1) * g.nx + i];
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) {
        const double wq = tw[a] * tw[b] * area;
        ex->ExactSolution<2>::vector_value(dealii::Point<2>(g.x0 + (i + tq[a]) * g.hx, g.y0 + (j + tq[b]) * g.hy), v);
        const double uhx = (uxl * (1 - tq[a]) + uxr * tq[a]) / g.hy, uhy = (uyb * (1 - tq[b]) + uyt * tq[b]) / g.hx;
        acc[0] += wq * ((uhx - v(0)) * (uhx - v(0)) + (uhy - v(1)) * (uhy - v(1)));
        acc[1] += wq * (v(0) * v(0) + v(1) * v(1));
        acc[2] += wq * (p - v(2)) * (p - v(2));
        acc[3] += wq * v(2) * v(2);
      }
    if (lvl > 0) {  // :1728-1743
      const dealii::Point<2> ctr(g.x0 + (i + 0.5) * g.hx, g.y0 + (j + 0.5) * g.hy);
      const double d = xo[nf + c] + proj_now->PressureBoundaryValues<2>::value(ctr, 0) - p -
                       proj_old->PressureBoundaryValues<2>::value(ctr, 0);
      acc[4] += d * d * area;
    }
  }
  __shared__ double red[5][256];
  for (int k = 0; k < 5; ++k) red[k][threadIdx.x] = acc[k];
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (threadIdx.x < off)
      for (int k = 0; k < 5; ++k) red[k][threadIdx.x] += red[k][threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x < 5) partial[(int64_t(lvl) * gridDim.x + blockIdx.x) * 5 + threadIdx.x] = red[threadIdx.x][0];
}

__global__ void __launch_bounds__(32) k_error_final(int nblk, const double *__restrict__ partial, double *__restrict__ out) {
  const int lvl = blockIdx.x, k = threadIdx.x;
  if (k >= 5) return;
  double s = 0;
  for (int b = 0; b < nblk; ++b) s += partial[(int64_t(lvl) * nblk + b) * 5 + k];
  out[lvl * 5 + k] = s;
}

__global__ void __launch_bounds__(256) k_expand(DevGrid g, const double *__restrict__ sol, int l0, int64_t cells,
                                                 double *__restrict__ pts, double *__restrict__ u, double *__restrict__ p) {
  const int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;  // (cell, vertex)
  if (idx >= cells * 8) return;
  const int v = int(idx & 7);
  const int64_t c = idx >> 3;
  const int64_t np = int64_t(g.nx) * g.ny, nxe = int64_t(g.nx + 1) * g.ny, nf = nxe + int64_t(g.nx) * (g.ny + 1), n = nf + np;
  const int l = l0 + int(c / np);
  const int64_t c2 = c % np;
  const int j = int(c2 / g.nx), i = int(c2 - int64_t(j) * g.nx);
  const double *x = sol + int64_t(l) * n;
  pts[idx * 3 + 0] = g.x0 + (i + (v & 1)) * g.hx;
  pts[idx * 3 + 1] = g.y0 + (j + ((v >> 1) & 1)) * g.hy;
  pts[idx * 3 + 2] = (l + ((v >> 2) & 1)) * g.dt;
  u[idx * 3 + 0] = x[int64_t(j) * (g.nx + 1) + i + (v & 1)] / g.hy;
  u[idx * 3 + 1] = x[nxe + int64_t(j + ((v >> 1) & 1)) * g.nx + i] / g.hx;
  u[idx * 3 + 2] = 0.0;
  p[idx] = x[nf + c2];
}

}  // namespace

bool enabled() {
  const char *e = std::getenv("MMMFE_HOST_DATA");
  return !(e && e[0] == '1');
}

void initial_pressure(const Grid &g, double *d_p0) {
  Levels lv(g);
  const InitialCondition<2> ic;
  const int64_t np = int64_t(g.nx) * g.ny;
  k_initial_pressure<<<unsigned((np + 255) / 256), 256>>>(lv.grid(g), gauss01(5), pod_of(ic), d_p0);
  cu(cudaDeviceSynchronize(), "k_initial_pressure");
}

void bar_source(const Grid &g, double c0, double alpha, double *d_src) {
  Levels lv(g);
  const RightHandSidePressure<2> f(c0, alpha);
  const int64_t np = int64_t(g.nx) * g.ny;
  k_bar_source<<<dim3(unsigned((np + 255) / 256), g.nt), 256>>>(lv.grid(g), gauss01(2), pod_of(f), d_src);
  cu(cudaDeviceSynchronize(), "k_bar_source");
}

void boundary_values(const Grid &g, double Lx, double Ly, int n_dir, const int32_t *fr_side, int nfr, int qdegree,
                     int manufact, const double bc_val_of_side[4], double *d_fr_vals) {
  if (n_dir <= 0) return;
  Levels lv(g);
  int32_t *d_side = nullptr;
  cu(cudaMalloc(&d_side, sizeof(int32_t) * size_t(n_dir)), "cudaMalloc");
  cu(cudaMemcpy(d_side, fr_side, sizeof(int32_t) * size_t(n_dir), cudaMemcpyHostToDevice), "cudaMemcpy");
  BcArgs a{Lx, Ly, n_dir, nfr, qdegree, manufact, {bc_val_of_side[0], bc_val_of_side[1], bc_val_of_side[2], bc_val_of_side[3]}, d_side};
  const PressureBoundaryValues<2> gf;
  k_boundary_values<<<dim3((n_dir + 127) / 128, g.nt), 128>>>(lv.grid(g), a, gauss01(qdegree), pod_of(gf), d_fr_vals);
  cu(cudaDeviceSynchronize(), "k_boundary_values");
  cudaFree(d_side);
}

void add_constant_rows(int nt, int64_t stride, int n, const int32_t *index, const double *value, double *d_rows) {
  rows_op(nt, stride, n, index, value, d_rows, 1);
}

void set_constant_rows(int nt, int64_t stride, int n, const int32_t *index, const double *value, double *d_rows) {
  rows_op(nt, stride, n, index, value, d_rows, 0);
}

void error_sums(const Grid &g, const double *d_solution, double *out) {
  Levels lv(g);
  const int64_t np = int64_t(g.nx) * g.ny;
  const int nblk = int((np + ERR_CELLS - 1) / ERR_CELLS);
  double *d_partial = nullptr, *d_out = nullptr;
  cu(cudaMalloc(&d_partial, sizeof(double) * size_t(g.nt) * nblk * 5), "cudaMalloc");
  cu(cudaMalloc(&d_out, sizeof(double) * size_t(g.nt) * 5), "cudaMalloc");
  const ExactSolution<2> ex;
  const PressureBoundaryValues<2> pb;
  k_error_sums<<<dim3(nblk, g.nt), 256>>>(lv.grid(g), pod_of(ex), pod_of(pb), d_solution, d_partial);
  k_error_final<<<g.nt, 32>>>(nblk, d_partial, d_out);
  cu(cudaMemcpy(out, d_out, sizeof(double) * size_t(g.nt) * 5, cudaMemcpyDeviceToHost), "error sums");
  cudaFree(d_partial);
  cudaFree(d_out);
}

void expand_space_time(const Grid &g, const double *d_solution, int l0, int l1, double *points, double *u, double *p) {
  const int64_t cells = int64_t(g.nx) * g.ny * (l1 - l0);
  if (cells <= 0) return;
  double *d = nullptr;
  cu(cudaMalloc(&d, sizeof(double) * size_t(cells) * 8 * 7), "cudaMalloc");
  double *d_pts = d, *d_u = d + cells * 24, *d_p = d + cells * 48;
  DevGrid dg{g.nx, g.ny, g.nt, g.x0, g.y0, g.hx, g.hy, g.dt, nullptr};
  k_expand<<<unsigned((cells * 8 + 255) / 256), 256>>>(dg, d_solution, l0, cells, d_pts, d_u, d_p);
  cu(cudaMemcpy(points, d_pts, sizeof(double) * size_t(cells) * 24, cudaMemcpyDeviceToHost), "expand");
  cu(cudaMemcpy(u, d_u, sizeof(double) * size_t(cells) * 24, cudaMemcpyDeviceToHost), "expand");
  cu(cudaMemcpy(p, d_p, sizeof(double) * size_t(cells) * 8, cudaMemcpyDeviceToHost), "expand");
  cudaFree(d);
}

}  // namespace devdata
}  // namespace vt_darcy

#else  // MMMFE_NO_DEVICE_DATA: the data header does not compile for the device; the front-end keeps its host loops

namespace vt_darcy {
namespace devdata {
bool enabled() { return false; }
static void no() { throw std::runtime_error("data_device: built without the device data path"); }
void initial_pressure(const Grid &, double *) { no(); }
void bar_source(const Grid &, double, double, double *) { no(); }
void boundary_values(const Grid &, double, double, int, const int32_t *, int, int, int, const double[4], double *) { no(); }
void add_constant_rows(int, int64_t, int, const int32_t *, const double *, double *) { no(); }
void set_constant_rows(int, int64_t, int, const int32_t *, const double *, double *) { no(); }
void error_sums(const Grid &, const double *, double *) { no(); }
void expand_space_time(const Grid &, const double *, int, int, double *, double *, double *) { no(); }
}  // namespace devdata
}  // namespace vt_darcy

#endif
