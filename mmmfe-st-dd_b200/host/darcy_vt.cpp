// Host front-end: see darcy_vt.h.  Everything here is one-time work per refinement cycle (or one pass over the
// solution); the interface-GMRES hot path runs in the CUDA engine behind include/mmmfe_b200.h.
#include "darcy_vt.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <thread>

#include <omp.h>
#include <zlib.h>

#include "mmmfe_b200.h"
#include "data_device.h"

#ifndef MMMFE_DATA_HEADER
#define MMMFE_DATA_HEADER "data_default.h"
#endif
#include MMMFE_DATA_HEADER

namespace vt_darcy {

namespace {

using clk = std::chrono::steady_clock;
double since(clk::time_point t0) { return std::chrono::duration<double>(clk::now() - t0).count(); }

const double kNormal[4] = {-1.0, 1.0, 1.0, -1.0};  // get_normal_direction, inc/utilities.h:79-91
enum { BOTTOM = 0, RIGHT = 1, TOP = 2, LEFT = 3 };

// QGauss<1>(n) on [0,1]
void gauss01(int n, std::vector<double> &x, std::vector<double> &w) {
  x.assign(n, 0.0);
  w.assign(n, 0.0);
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
    x[n - 1 - i] = 0.5 * (z + 1.0);
    w[n - 1 - i] = 1.0 / ((1.0 - z * z) * pp * pp);
  }
}

// orthonormal Legendre polynomial of degree a on [0,1] (deal.II Polynomials::Legendre)
double legendre01(int a, double xi) {
  const double s = 2.0 * xi - 1.0;
  double p0 = 1.0, p1 = s;
  if (a == 0) return 1.0;
  for (int j = 1; j < a; ++j) {
    const double p2 = ((2.0 * j + 1.0) * s * p1 - j * p0) / (j + 1.0);
    p0 = p1;
    p1 = p2;
  }
  return std::sqrt(2.0 * a + 1.0) * p1;
}

// find_2_divisors, inc/utilities.h:94-110
void find_divisors(unsigned n, unsigned &n0, unsigned &n1) {
  unsigned tmp = unsigned(std::floor(std::sqrt(double(n))));
  while (n % tmp != 0) --tmp;
  n0 = n / tmp;
  n1 = tmp;
}

// find_neighbors (dim == 2), inc/utilities.h:183-205, 297-305; order bottom, right, top, left
void find_neighbors(int r, int n0, int n1, int nb[4]) {
  nb[0] = r - n0; nb[1] = r + 1; nb[2] = r + n0; nb[3] = r - 1;
  if (r % n0 == 0) nb[3] = -1;
  if ((r + 1) % n0 == 0) nb[1] = -1;
  for (int s = 0; s < 4; ++s)
    if (nb[s] < 0 || nb[s] >= n0 * n1) nb[s] = -1;
}

struct Csr {
  long long n_rows = 0, n_cols = 0;
  std::vector<long long> rp;
  std::vector<int> col;
  std::vector<double> val;
  mmmfe_csr view() const { return mmmfe_csr{n_rows, n_cols, (const int64_t *)rp.data(), col.data(), val.data()}; }
};

// One subdomain: canonical dof order x-edges (i=0..nx, j) at j*(nx+1)+i, y-edges at nxe + j*nx + i, cells after.
struct Sub {
  int r, ix, iy, nx, ny, nt, n0 = 1, n1 = 1;
  double x0, y0, Lx, Ly, hx, hy, dt, final_time = 0.0;
  int nxe, nye, nf, np, n;
  int nb[4], n_side[4];
  double edge_len[4];
  std::vector<int> side_edges[4];
  Csr matrix, f2c[4], c2f[4];
  int mortar_len[4] = {0, 0, 0, 0};
  int local = -1;
  // essential (Neumann) conditions of the outer boundary, constraint_bc of src/darcy_vtdd.cc:205-275
  std::vector<int> con_dofs;
  std::vector<double> con_vals, con_diag;
  std::vector<std::pair<int, double>> rhs_bc;  // system_rhs_bar_bc (:479): row, value
  std::vector<double> solution;  // [nt][n] after the final sweep
  std::vector<double> levels_t;  // accumulated times t += dt
};

void init_sub(Sub &s, int r, int n0, int n1, const std::vector<int> &pat, double final_time) {
  s.r = r; s.ix = r % n0; s.iy = r / n0; s.n0 = n0; s.n1 = n1; s.final_time = final_time;
  s.Lx = 1.0 / n0; s.Ly = 1.0 / n1;
  s.x0 = s.ix * s.Lx;                                   // get_subdomain_coordinates, inc/utilities.h:152-157
  s.y0 = s.Ly * std::floor(double(r) / double(n0));
  s.nx = pat[0]; s.ny = pat[1]; s.nt = pat[2];
  s.hx = s.Lx / s.nx; s.hy = s.Ly / s.ny;
  s.dt = final_time / double(s.nt);                     // src/darcy_vtdd.cc:2083
  s.nxe = (s.nx + 1) * s.ny; s.nye = s.nx * (s.ny + 1);
  s.nf = s.nxe + s.nye; s.np = s.nx * s.ny; s.n = s.nf + s.np;
  find_neighbors(r, n0, n1, s.nb);
  s.n_side[0] = s.n_side[2] = s.nx; s.n_side[1] = s.n_side[3] = s.ny;
  s.edge_len[0] = s.edge_len[2] = s.hx; s.edge_len[1] = s.edge_len[3] = s.hy;
  for (int k = 0; k < 4; ++k) s.side_edges[k].clear();
  for (int i = 0; i < s.nx; ++i) { s.side_edges[BOTTOM].push_back(s.nxe + i); s.side_edges[TOP].push_back(s.nxe + s.ny * s.nx + i); }
  for (int j = 0; j < s.ny; ++j) { s.side_edges[RIGHT].push_back(j * (s.nx + 1) + s.nx); s.side_edges[LEFT].push_back(j * (s.nx + 1)); }
  s.levels_t.resize(s.nt);
  double t = 0.0;
  for (int l = 0; l < s.nt; ++l) { t += s.dt; s.levels_t[l] = t; }  // prm.time += prm.time_step, :887
}

// assemble_system, src/darcy_vtdd.cc:401-484: [[A, -B^T], [dt B, c0 M]], QGauss<2>(degree+2) = 2x2 points.
// Row-parallel: every row gathers the contributions of its (at most two) cells in cell order, which is the order
// the reference's cell loop adds them in; explicit zeros of the local matrices are kept (same sparsity pattern).
void assemble_matrix(Sub &s, double c0) {
  const int n = s.n;
  constexpr int W = 12;
  std::vector<double> g, w;
  gauss01(2, g, w);
  const KInverse<2> k_inverse;
  const double div[4] = {-1.0, 1.0, -1.0, 1.0};
  const int nx = s.nx, ny = s.ny;
  // local mass matrices of all cells, 16 doubles each: aloc[e][f], e, f = x-left, x-right, y-bottom, y-top
  std::unique_ptr<double[]> A(new double[size_t(s.np) * 16]);
#pragma omp parallel
  {
    std::vector<dealii::Point<2>> pts(4);
    std::vector<dealii::Tensor<2, 2>> kin(4);
#pragma omp for schedule(static)
    for (int c = 0; c < s.np; ++c) {
      const int j = c / nx, i = c % nx;
      double *aloc = &A[size_t(c) * 16];
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) pts[a * 2 + b] = dealii::Point<2>(s.x0 + (i + g[a]) * s.hx, s.y0 + (j + g[b]) * s.hy);
      k_inverse.value_list(pts, kin);
      for (int t = 0; t < 16; ++t) aloc[t] = 0;
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
          const int q = a * 2 + b;
          const double jxw = w[a] * w[b] * s.hx * s.hy;
          // basis values (flux-integral normalised): x-left, x-right, y-bottom, y-top
          const double phi[4][2] = {{(1 - g[a]) / s.hy, 0}, {g[a] / s.hy, 0}, {0, (1 - g[b]) / s.hx}, {0, g[b] / s.hx}};
          for (int e = 0; e < 4; ++e)
            for (int f = 0; f < 4; ++f) {
              double v = 0;
              for (int d = 0; d < 2; ++d)
                for (int d2 = 0; d2 < 2; ++d2) v += phi[e][d] * kin[q][d][d2] * phi[f][d2];
              aloc[e * 4 + f] += jxw * v;
            }
        }
    }
  }
  // Rows straight into CSR: a flux row holds the 4 edges + the cell of each adjacent cell with the shared edge (the
  // row's own) merged, a pressure row its 4 edges + itself, so the row lengths are known up front.
  Csr &m = s.matrix;
  m.n_rows = m.n_cols = n;
  m.rp.assign(size_t(n) + 1, 0);
#pragma omp parallel for schedule(static)
  for (int r = 0; r < n; ++r) {
    int cells;
    if (r < s.nxe) { const int i = r % (nx + 1); cells = (i > 0) + (i < nx); }
    else if (r < s.nf) { const int j = (r - s.nxe) / nx; cells = (j > 0) + (j < ny); }
    else cells = 0;
    m.rp[size_t(r) + 1] = r < s.nf ? 5 * cells - (cells == 2 ? 1 : 0) : 5;
  }
  for (int r = 0; r < n; ++r) m.rp[size_t(r) + 1] += m.rp[size_t(r)];
  m.col.resize(size_t(m.rp[n]));
  m.val.resize(size_t(m.rp[n]));
  bool overflow = false;
#pragma omp parallel
  {
    std::pair<int, double> row[W];
    int cnt = 0;
    auto add = [&](int c, double v) {
      for (int t = 0; t < cnt; ++t) if (row[t].first == c) { row[t].second += v; return; }
      if (cnt >= W) { overflow = true; return; }
      row[cnt++] = {c, v};
    };
    // the contributions of local edge e of cell (i, j)
    auto edge_row = [&](int i, int j, int e) {
      const int loc[4] = {j * (nx + 1) + i, j * (nx + 1) + i + 1, s.nxe + j * nx + i, s.nxe + (j + 1) * nx + i};
      const double *aloc = &A[size_t(j * nx + i) * 16];
      for (int f = 0; f < 4; ++f) add(loc[f], aloc[e * 4 + f]);
      add(s.nf + j * nx + i, -div[e]);  // -phi_p div phi_u
    };
    auto flush = [&](int r) {
      if (cnt != int(m.rp[size_t(r) + 1] - m.rp[size_t(r)])) { overflow = true; cnt = 0; return; }
      std::sort(row, row + cnt);
      for (int t = 0; t < cnt; ++t) { m.col[size_t(m.rp[r]) + t] = row[t].first; m.val[size_t(m.rp[r]) + t] = row[t].second; }
      cnt = 0;
    };
#pragma omp for schedule(static)
    for (int r = 0; r < s.nxe; ++r) {  // x-edges: left cell, then right cell
      const int j = r / (nx + 1), i = r % (nx + 1);
      if (i > 0) edge_row(i - 1, j, 1);
      if (i < nx) edge_row(i, j, 0);
      flush(r);
    }
#pragma omp for schedule(static)
    for (int k = 0; k < s.nye; ++k) {  // y-edges: lower cell, then upper cell
      const int j = k / nx, i = k % nx;
      if (j > 0) edge_row(i, j - 1, 3);
      if (j < ny) edge_row(i, j, 2);
      flush(s.nxe + k);
    }
#pragma omp for schedule(static)
    for (int c = 0; c < s.np; ++c) {
      const int j = c / nx, i = c % nx, r = s.nf + c;
      const int loc[4] = {j * (nx + 1) + i, j * (nx + 1) + i + 1, s.nxe + j * nx + i, s.nxe + (j + 1) * nx + i};
      for (int e = 0; e < 4; ++e) add(loc[e], s.dt * div[e]);  // dt div phi_u phi_p
      add(r, c0 * s.hx * s.hy);
      flush(r);
    }
  }
  if (overflow) throw std::runtime_error("matrix row overflow");
}

int side_to_bc(int side);

// constraint_bc (src/darcy_vtdd.cc:205-275) applied the way distribute_local_to_global does at :479.  Without a
// manufactured solution the normal flux on every outer side marked 'N' is essential: u.n = bc_const_functs[left,
// bottom, right, top] (:228-231 -- index 3 for the top, unlike the Dirichlet data of :673).  A flux dof is the flux
// through its edge in the positive coordinate direction, so the constrained value is normal * g * |e|.  Constrained
// rows and columns leave the matrix, the diagonal keeps |a_ii| (an outer edge belongs to one cell: local and global
// diagonal agree), every other row j receives -a_ji g_i in system_rhs_bar_bc.
void constrain_matrix(Sub &s, bool manufact, const std::vector<char> &bc_type, const std::vector<double> &bc_val) {
  s.con_dofs.clear(); s.con_vals.clear(); s.con_diag.clear(); s.rhs_bc.clear();
  if (manufact) return;  // :210, 271-273: all outer sides are Dirichlet
  std::vector<double> g(s.n, 0.0);
  std::vector<char> is_con(s.n, 0);
  for (int side = 0; side < 4; ++side) {
    if (s.nb[side] >= 0 || bc_type[side_to_bc(side)] != 'N') continue;
    const double v = kNormal[side] * bc_val[side_to_bc(side)] * s.edge_len[side];
    for (int m = 0; m < s.n_side[side]; ++m) {
      const int d = s.side_edges[side][m];
      s.con_dofs.push_back(d); s.con_vals.push_back(v);
      g[d] = v; is_con[d] = 1;
    }
  }
  if (s.con_dofs.empty()) return;
  Csr &m = s.matrix, out;
  out.n_rows = out.n_cols = m.n_rows;
  out.rp.assign(size_t(s.n) + 1, 0);
  out.col.reserve(m.col.size()); out.val.reserve(m.val.size());
  std::vector<double> diag(s.n, 0.0);
  for (int r = 0; r < s.n; ++r) {
    double bc = 0.0;
    bool touched = false;
    for (long long k = m.rp[r]; k < m.rp[r + 1]; ++k) {
      const int c = m.col[k];
      if (c == r) diag[r] = m.val[k];
      if (is_con[r]) continue;
      if (is_con[c]) { bc -= m.val[k] * g[c]; touched = true; continue; }
      out.col.push_back(c); out.val.push_back(m.val[k]);
    }
    if (is_con[r]) { out.col.push_back(r); out.val.push_back(std::fabs(diag[r])); }
    else if (touched && bc != 0.0) s.rhs_bc.push_back({r, bc});
    out.rp[r + 1] = (long long)out.col.size();
  }
  for (int d : s.con_dofs) s.con_diag.push_back(std::fabs(diag[d]));
  m = std::move(out);
}

// cell index of a coordinate given in units of cells; ties go to the lower cell / earlier level (quirk Q5)
int locate(double u, int n) {
  const double r = std::nearbyint(u);
  int idx = (std::fabs(u - r) < 1e-9) ? int(r) - 1 : int(std::floor(u));
  return std::min(std::max(idx, 0), n - 1);
}

void push_merged(std::vector<int> &rc, std::vector<double> &rv, int c, double v) {
  for (size_t t = 0; t < rc.size(); ++t) if (rc[t] == c) { rv[t] += v; return; }
  rc.push_back(c); rv.push_back(v);
}

void finish_row(Csr &m, std::vector<int> &rc, std::vector<double> &rv) {
  std::vector<std::pair<int, double>> row(rc.size());
  for (size_t t = 0; t < rc.size(); ++t) row[t] = {rc[t], rv[t]};
  std::sort(row.begin(), row.end());
  for (auto &e : row) { m.col.push_back(e.first); m.val.push_back(e.second); }
  m.rp.push_back((long long)m.col.size());
  rc.clear(); rv.clear();
}

// ---- MMMFE_TIE_RULE=dealii91 ---------------------------------------------------------------------------------------
// Which fine cell a mortar Gauss point gets that lies on a fine cell boundary, decided the way deal.II's point
// location (FEFieldFunction, inc/utilities.h:484; 9.0/9.1 series) decides it in IEEE doubles -- the rule report.pdf
// Table 4 was produced with (all three rows, every printed digit: tests/test_oracle_ties.py pins the Python
// restatement oracle/dealii_point_location.py, tests/test_host_frontend.py pins this one against it).  Restated:
// subdivided_hyper_rectangle vertices fl(p1 + fl(i * fl((p2 - p1) / reps))); the mapped quadrature point as the sum
// of the eight hard-coded trilinear shape values times the mortar cell's vertices; find_active_cell_around_point:
// closest fine vertex, its cells in ascending index order, stably sorted by descending scalar product of the
// normalised vertex->centre direction (centre = sum of the 8 vertices / 8) with the normalised vertex->point
// direction; the first cell of that order that holds the point.  No fused multiply-adds (the file is built with
// -ffp-contract=off).  Default (variable unset): the exact-arithmetic limit of this rule, lower cell / earlier level.
bool tie_rule_dealii91() {
  const char *v = std::getenv("MMMFE_TIE_RULE");
  if (!v || !*v || std::string(v) == "lower" || std::string(v) == "exact") return false;
  if (std::string(v) == "dealii91") return true;
  throw std::runtime_error(std::string("MMMFE_TIE_RULE must be 'lower' or 'dealii91', not '") + v + "'");
}

struct Vec3 { double v[3]; };

static Vec3 normalised(Vec3 t) {
  double s = t.v[0] * t.v[0];
  s = s + t.v[1] * t.v[1];
  s = s + t.v[2] * t.v[2];
  const double n = std::sqrt(s);
  if (n > 0) for (double &x : t.v) x = x / n;
  return t;
}

// picks[((ft * ms + fs) * 3 + r) * 3 + q]: bit 0 = space tie goes to the upper cell, bit 1 = time tie to the later level
std::vector<unsigned char> dealii91_tie_picks(const Sub &s, int side, const std::vector<int> &mortar) {
  static const double g[3] = {0.11270166537925831148, 0.5, 0.88729833462074168852};  // QGauss<1>(3) on [0, 1]
  const int n_fine[3] = {s.nx, s.ny, s.nt}, n_mortar[3] = {mortar[0], mortar[1], mortar[2]};
  // get_subdomain_coordinates, inc/utilities.h:152-157; p1_st, p2_st of src/darcy_vtdd.cc:2097
  const double dims[2] = {1.0 / double(s.n0), 1.0 / double(s.n1)};
  const double p1[3] = {double(s.r % s.n0) * dims[0], dims[1] * std::floor(double(s.r) / double(s.n0)), 0.0};
  const double p2[3] = {p1[0] + dims[0], p1[1] + dims[1], s.final_time};
  std::vector<double> fine[3], mort[3];
  for (int a = 0; a < 3; ++a) {
    const double df = (p2[a] - p1[a]) / double(n_fine[a]), dm = (p2[a] - p1[a]) / double(n_mortar[a]);
    for (int i = 0; i <= n_fine[a]; ++i) fine[a].push_back(p1[a] + double(i) * df);
    for (int i = 0; i <= n_mortar[a]; ++i) mort[a].push_back(p1[a] + double(i) * dm);
  }
  const int face = side == LEFT ? 0 : side == RIGHT ? 1 : side == BOTTOM ? 2 : 3;  // deal.II face of the 3-D cell
  const int normal = face < 2 ? 0 : 1, tangent = 1 - normal;
  const int ms = n_mortar[tangent], mnt = n_mortar[2];
  std::vector<unsigned char> picks(size_t(ms) * mnt * 9, 0);
  auto centre_dir = [&](const int c[3], const Vec3 &vc) {
    Vec3 p{{0.0, 0.0, 0.0}};
    for (int v = 0; v < 8; ++v) {
      const int o[3] = {v & 1, (v >> 1) & 1, (v >> 2) & 1};
      for (int a = 0; a < 3; ++a) p.v[a] = p.v[a] + fine[a][c[a] + o[a]];
    }
    Vec3 t;
    for (int a = 0; a < 3; ++a) t.v[a] = p.v[a] / 8.0 - vc.v[a];
    return normalised(t);
  };
  for (int cs = 0; cs < ms; ++cs)
    for (int ct = 0; ct < mnt; ++ct)
      for (int r = 0; r < 3; ++r)
        for (int q = 0; q < 3; ++q) {
          const bool tie_s = q == 1 && ((2LL * cs + 1) * n_fine[tangent]) % (2LL * n_mortar[tangent]) == 0;
          const bool tie_t = r == 1 && ((2LL * ct + 1) * n_fine[2]) % (2LL * n_mortar[2]) == 0;
          if (!tie_s && !tie_t) continue;
          // QProjector::project_to_face: (face, q0, q1) on x-faces, (q1, face, q0) on y-faces; q = space, r = time
          double unit[3];
          int cell[3];
          if (face < 2) { unit[0] = double(face); unit[1] = g[q]; unit[2] = g[r]; cell[0] = face == 0 ? 0 : n_mortar[0] - 1; cell[1] = cs; }
          else { unit[0] = g[q]; unit[1] = double(face - 2); unit[2] = g[r]; cell[0] = cs; cell[1] = face == 2 ? 0 : n_mortar[1] - 1; }
          cell[2] = ct;
          const double x = unit[0], y = unit[1], z = unit[2];
          const double shape[8] = {(1. - x) * (1. - y) * (1. - z), x * (1. - y) * (1. - z), (1. - x) * y * (1. - z),
                                   x * y * (1. - z), (1. - x) * (1. - y) * z, x * (1. - y) * z, (1. - x) * y * z, x * y * z};
          Vec3 p;
          for (int a = 0; a < 3; ++a) p.v[a] = shape[0] * mort[a][cell[a]];
          for (int v = 1; v < 8; ++v) {
            const int o[3] = {v & 1, (v >> 1) & 1, (v >> 2) & 1};
            for (int a = 0; a < 3; ++a) p.v[a] = p.v[a] + shape[v] * mort[a][cell[a] + o[a]];
          }
          int idx[3];
          Vec3 vc;
          for (int a = 0; a < 3; ++a) {
            int best = 0;
            for (int i = 1; i <= n_fine[a]; ++i)
              if (std::fabs(p.v[a] - fine[a][i]) < std::fabs(p.v[a] - fine[a][best])) best = i;
            idx[a] = best; vc.v[a] = fine[a][best];
          }
          Vec3 to_point;
          for (int a = 0; a < 3; ++a) to_point.v[a] = p.v[a] - vc.v[a];
          to_point = normalised(to_point);
          struct Cand { int c[3]; double prod; };
          std::vector<Cand> cand;
          for (int dk = -1; dk <= 0; ++dk)      // ascending cell index: time slowest, x fastest
            for (int dj = -1; dj <= 0; ++dj)
              for (int di = -1; di <= 0; ++di) {
                Cand cd{{idx[0] + di, idx[1] + dj, idx[2] + dk}, 0.0};
                bool in = true;
                for (int a = 0; a < 3; ++a) in = in && cd.c[a] >= 0 && cd.c[a] < n_fine[a];
                if (!in) continue;
                const Vec3 d = centre_dir(cd.c, vc);
                double sp = d.v[0] * to_point.v[0];
                sp = sp + d.v[1] * to_point.v[1];
                sp = sp + d.v[2] * to_point.v[2];
                cd.prod = sp;
                cand.push_back(cd);
              }
          std::stable_sort(cand.begin(), cand.end(), [](const Cand &a, const Cand &b) { return a.prod > b.prod; });
          const Cand *chosen = nullptr;
          for (const Cand &cd : cand) {
            bool holds = true;
            for (int a = 0; a < 3; ++a)
              if (a != normal) holds = holds && fine[a][cd.c[a]] <= p.v[a] && p.v[a] <= fine[a][cd.c[a] + 1];
            if (holds) { chosen = &cd; break; }
          }
          if (!chosen) throw std::runtime_error("dealii91 tie rule: no cell holds a mortar Gauss point");
          const long long s_vertex = (2LL * cs + 1) * n_fine[tangent] / (2LL * n_mortar[tangent]);
          const long long t_vertex = (2LL * ct + 1) * n_fine[2] / (2LL * n_mortar[2]);
          unsigned char bits = 0;
          if (tie_s && chosen->c[tangent] == s_vertex) bits |= 1;
          if (tie_t && chosen->c[2] == t_vertex) bits |= 2;
          picks[((size_t(ct) * ms + cs) * 3 + r) * 3 + q] = bits;
        }
  return picks;
}

// project_mortar as two explicit sampling operators (inc/utilities.h:471-505, inc/projector.h:150-207, 445-524;
// SURVEY.md appendix A.4).  Mortar layout ((ft*ms + fs)*(k+1) + b)*(k+1) + a; fine layout level*n_side + m.
void projector_tables(Sub &s, int side, const std::vector<int> &mortar, int k) {
  const int ms = (side == BOTTOM || side == TOP) ? mortar[0] : mortar[1], mnt = mortar[2];
  const int ns = s.n_side[side], nt = s.nt, nq = k + 1, nd = nq * nq;
  std::vector<double> xi, w;
  gauss01(nq, xi, w);
  std::vector<double> leg(size_t(nq) * nq);
  for (int a = 0; a < nq; ++a) for (int q = 0; q < nq; ++q) leg[a * nq + q] = legendre01(a, xi[q]);
  const double H = ((side == BOTTOM || side == TOP) ? s.Lx : s.Ly) / ms, T = (s.nt * s.dt) / mnt, e_len = s.edge_len[side];
  s.mortar_len[side] = ms * mnt * nd;
  Csr &f2c = s.f2c[side], &c2f = s.c2f[side];
  f2c = Csr(); c2f = Csr();
  f2c.n_rows = s.mortar_len[side]; f2c.n_cols = (long long)ns * nt; f2c.rp.assign(1, 0);
  c2f.n_rows = (long long)ns * nt; c2f.n_cols = s.mortar_len[side]; c2f.rp.assign(1, 0);
  std::vector<int> rc;
  std::vector<double> rv;
  // f2c: Gauss points of each mortar face sample the piecewise-constant fine normal flux U/|e|
  std::vector<unsigned char> picks;  // MMMFE_TIE_RULE=dealii91: per-point tie choices (3-point rule only)
  if (k % 2 == 0 && tie_rule_dealii91()) {  // odd-point Gauss rules have a centre abscissa that can hit a cell boundary
    if (k != 2) throw std::runtime_error("MMMFE_TIE_RULE=dealii91 is restated for mortar_degree 2 only");
    picks = dealii91_tie_picks(s, side, mortar);
  }
  const double coef = H * T / e_len;
  for (int ft = 0; ft < mnt; ++ft)
    for (int fs = 0; fs < ms; ++fs)
      for (int b = 0; b < nq; ++b)
        for (int a = 0; a < nq; ++a) {
          for (int r = 0; r < nq; ++r) {
            const int lvl0 = locate((ft + xi[r]) / mnt * nt, nt);
            for (int q = 0; q < nq; ++q) {
              int m = locate((fs + xi[q]) / ms * ns, ns), lvl = lvl0;
              if (!picks.empty()) {
                const unsigned char bits = picks[((size_t(ft) * ms + fs) * 3 + r) * 3 + q];
                if (bits & 1) m = std::min(m + 1, ns - 1);
                if (bits & 2) lvl = std::min(lvl + 1, nt - 1);
              }
              push_merged(rc, rv, lvl * ns + m, coef * w[q] * w[r] * leg[a * nq + q] * leg[b * nq + r]);
            }
          }
          finish_row(f2c, rc, rv);
        }
  // c2f: Gauss points of each fine space-time face sample the mortar polynomial; result is (2-D dof)/|e|
  for (int lvl = 0; lvl < nt; ++lvl)
    for (int m = 0; m < ns; ++m) {
      for (int r = 0; r < nq; ++r) {
        const double ut = (lvl + xi[r]) / nt * mnt;
        const int ft = locate(ut, mnt);
        for (int q = 0; q < nq; ++q) {
          const double us = (m + xi[q]) / ns * ms;
          const int fs = locate(us, ms);
          for (int b = 0; b < nq; ++b)
            for (int a = 0; a < nq; ++a)
              push_merged(rc, rv, ((ft * ms + fs) * nq + b) * nq + a,
                          w[q] * w[r] * legendre01(a, us - fs) * legendre01(b, ut - ft) / (H * T));
        }
      }
      finish_row(c2f, rc, rv);
    }
}

void side_point(const Sub &s, int side, int m, double xi, double &x, double &y) {
  if (side == BOTTOM || side == TOP) { x = s.x0 + (m + xi) * s.hx; y = (side == BOTTOM) ? s.y0 : s.y0 + s.Ly; }
  else { y = s.y0 + (m + xi) * s.hy; x = (side == LEFT) ? s.x0 : s.x0 + s.Lx; }
}

int side_to_bc(int side) { return side == LEFT ? 0 : side == BOTTOM ? 1 : side == RIGHT ? 2 : 3; }  // :215-217

struct BarData {
  std::vector<double> p0, src, fr_vals;
  std::vector<int> fr_dofs;
};

// Which flux rows carry outer-boundary data and which constant rows system_rhs_bar_bc adds (both bar-data paths).
struct BarLayout {
  std::vector<int> fr_dofs, fr_side;  // fr_side[e] = side << 24 | position for the first n_dir entries, -1 after
  int n_dir = 0;
  std::vector<int> cell_rows, fr_slots;        // pressure rows (cell index) / flux slots that receive a constant
  std::vector<double> cell_vals, fr_slot_vals;
};

void bar_layout(const Sub &s, bool manufact, const std::vector<char> &bc_type, BarLayout &L) {
  for (int side = 0; side < 4; ++side) {
    if (s.nb[side] >= 0) continue;
    if (!manufact && bc_type[side_to_bc(side)] != 'D') continue;  // Neumann sides are essential: constrain_matrix
    for (int m = 0; m < s.n_side[side]; ++m) { L.fr_dofs.push_back(s.side_edges[side][m]); L.fr_side.push_back(side * (1 << 24) + m); }
  }
  L.n_dir = int(L.fr_dofs.size());
  // system_rhs_bar.sadd(1.0, system_rhs_bar_bc) (:863), every level.  Pressure rows go into the source, flux rows
  // into the sparse flux data; the constrained rows themselves get |a_ii| g_i so that the solve returns g_i (the
  // reference solves with 0 there and calls constraint_bc.distribute afterwards, :866 -- same solution).
  std::vector<std::pair<int, double>> extra;  // flux row -> constant right-hand-side value
  for (const auto &rv : s.rhs_bc) {
    if (rv.first >= s.nf) { L.cell_rows.push_back(rv.first - s.nf); L.cell_vals.push_back(rv.second); }
    else extra.push_back(rv);
  }
  for (size_t k = 0; k < s.con_dofs.size(); ++k) extra.push_back({s.con_dofs[k], s.con_diag[k] * s.con_vals[k]});
  std::map<int, int> slot_of;
  for (int e = 0; e < int(L.fr_dofs.size()); ++e) slot_of.insert({L.fr_dofs[e], e});
  for (size_t k = 0; k < extra.size(); ++k) {  // the device adds fr entries without atomics: one entry per dof
    auto it = slot_of.find(extra[k].first);
    int slot;
    if (it != slot_of.end()) slot = it->second;
    else { slot = int(L.fr_dofs.size()); L.fr_dofs.push_back(extra[k].first); L.fr_side.push_back(-1); slot_of.insert({extra[k].first, slot}); }
    // several constants of one slot are summed here, in order, so that both paths add ONE value per slot
    bool merged = false;
    for (size_t q = 0; q < L.fr_slots.size(); ++q) if (L.fr_slots[q] == slot) { L.fr_slot_vals[q] += extra[k].second; merged = true; break; }
    if (!merged) { L.fr_slots.push_back(slot); L.fr_slot_vals.push_back(extra[k].second); }
  }
}

double dirichlet_constant(int side, const std::vector<double> &bc_val) {
  return (side == TOP) ? bc_val[1] : bc_val[side_to_bc(side)];  // the reference reads [1] for the top (:673)
}

// The data-dependent part of assemble_rhs_bar (src/darcy_vtdd.cc:632-766): dt (f, w) per cell and the outer
// Dirichlet rows -s * int g; the c0 (p_old, w) term is added on the device.  HOST evaluation of the data classes
// (builds without the device data path, or MMMFE_HOST_DATA=1).
void bar_data(const Sub &s, int qdegree, bool manufact, const std::vector<char> &bc_type,
              const std::vector<double> &bc_val, double c0, double alpha, BarData &out) {
  std::vector<double> g2, w2, g5, w5, gq, wq;
  gauss01(2, g2, w2);
  gauss01(5, g5, w5);
  gauss01(qdegree, gq, wq);
  // VectorTools::project(IC, QGauss(degree+5)) onto DGQ0 = cell mean (src/darcy_vtdd.cc:2172-2176)
  out.p0.assign(s.np, 0.0);
  {
    const InitialCondition<2> ic;
#pragma omp parallel for schedule(static)
    for (int j = 0; j < s.ny; ++j) {
      dealii::Vector<double> v(3);
      for (int i = 0; i < s.nx; ++i) {
        double acc = 0;
        for (int a = 0; a < 5; ++a)
          for (int b = 0; b < 5; ++b) {
            ic.vector_value(dealii::Point<2>(s.x0 + (i + g5[a]) * s.hx, s.y0 + (j + g5[b]) * s.hy), v);
            acc += w5[a] * w5[b] * v(2);
          }
        out.p0[j * s.nx + i] = acc;
      }
    }
  }
  out.src.assign(size_t(s.nt) * s.np, 0.0);
  const double area = s.hx * s.hy;
#pragma omp parallel for schedule(static)
  for (int lvl = 0; lvl < s.nt; ++lvl) {
    RightHandSidePressure<2> f(c0, alpha);
    f.set_time(s.levels_t[lvl]);
    double *dst = &out.src[size_t(lvl) * s.np];
    for (int j = 0; j < s.ny; ++j)
      for (int i = 0; i < s.nx; ++i) {
        double acc = 0;
        for (int a = 0; a < 2; ++a)
          for (int b = 0; b < 2; ++b)
            acc += w2[a] * w2[b] * f.value(dealii::Point<2>(s.x0 + (i + g2[a]) * s.hx, s.y0 + (j + g2[b]) * s.hy));
        dst[j * s.nx + i] = s.dt * acc * area;  // :718
      }
  }
  BarLayout L;
  bar_layout(s, manufact, bc_type, L);
  out.fr_dofs = L.fr_dofs;
  for (int lvl = 0; lvl < s.nt; ++lvl)
    for (size_t k = 0; k < L.cell_rows.size(); ++k) out.src[size_t(lvl) * s.np + L.cell_rows[k]] += L.cell_vals[k];
  const int nfr = int(out.fr_dofs.size()), n_dir = L.n_dir;
  out.fr_vals.assign(size_t(s.nt) * nfr, 0.0);
  for (int lvl = 0; lvl < s.nt; ++lvl)
    for (size_t k = 0; k < L.fr_slots.size(); ++k) out.fr_vals[size_t(lvl) * nfr + L.fr_slots[k]] += L.fr_slot_vals[k];
#pragma omp parallel for schedule(static)
  for (int lvl = 0; lvl < s.nt; ++lvl) {
    PressureBoundaryValues<2> gfun;
    gfun.set_time(s.levels_t[lvl]);
    for (int e = 0; e < n_dir; ++e) {
      const int side = L.fr_side[e] >> 24, m = L.fr_side[e] & ((1 << 24) - 1);
      double acc = 0;
      for (int q = 0; q < qdegree; ++q) {
        double x, y;
        side_point(s, side, m, gq[q], x, y);
        double gv;
        if (manufact) gv = gfun.value(dealii::Point<2>(x, y));
        else gv = dirichlet_constant(side, bc_val);
        acc += wq[q] * gv;
      }
      out.fr_vals[size_t(lvl) * nfr + e] += -kNormal[side] * acc;  // :748-751
    }
  }
}

devdata::Grid dev_grid(const Sub &s) { return devdata::Grid{s.nx, s.ny, s.nt, s.x0, s.y0, s.hx, s.hy, s.dt, s.levels_t.data()}; }

// The same data evaluated by CUDA kernels straight into the engine's buffers (host/data_device.cu).
void bar_data_device(mmmfe_ctx *ctx, const Sub &s, int qdegree, bool manufact, const std::vector<char> &bc_type,
                     const std::vector<double> &bc_val, double c0, double alpha) {
  BarLayout L;
  bar_layout(s, manufact, bc_type, L);
  const int nfr = int(L.fr_dofs.size());
  double *d_p0 = nullptr, *d_src = nullptr, *d_fr = nullptr;
  if (mmmfe_bar_data_buffers(ctx, s.local, nfr, L.fr_dofs.data(), &d_p0, &d_src, &d_fr) != 0)
    throw std::runtime_error(std::string("mmmfe_bar_data_buffers: ") + mmmfe_last_error(ctx));
  const devdata::Grid g = dev_grid(s);
  devdata::initial_pressure(g, d_p0);
  devdata::bar_source(g, c0, alpha, d_src);
  devdata::add_constant_rows(s.nt, s.np, int(L.cell_rows.size()), L.cell_rows.data(), L.cell_vals.data(), d_src);
  devdata::add_constant_rows(s.nt, nfr, int(L.fr_slots.size()), L.fr_slots.data(), L.fr_slot_vals.data(), d_fr);
  double bc_side[4];
  for (int side = 0; side < 4; ++side) bc_side[side] = dirichlet_constant(side, bc_val);
  devdata::boundary_values(g, s.Lx, s.Ly, L.n_dir, L.fr_side.data(), nfr, qdegree, manufact ? 1 : 0, bc_side, d_fr);
}

// compute_errors per subdomain, src/darcy_vtdd.cc:1657-1775; numerators/denominators {velocity, pressure L2, DG}
void subdomain_errors(const Sub &s, double num[3], double den[3]) {
  const double tq[3] = {0.0, 0.5, 1.0}, tw[3] = {0.25, 0.5, 0.25};  // QIterated(QTrapezoid, 2)
  const int nt = s.nt;
  std::vector<double> n0(nt, 0), d0(nt, 0), n1(nt, 0), d1(nt, 0), jump(nt, 0);
  const double area = s.hx * s.hy;
#pragma omp parallel for schedule(dynamic)
  for (int lvl = 0; lvl < nt; ++lvl) {
    ExactSolution<2> ex;
    ex.set_time(s.levels_t[lvl]);
    PressureBoundaryValues<2> proj_now, proj_old;
    proj_now.set_time(s.levels_t[lvl] + 0.5 * s.dt);  // :1669-1670
    if (lvl > 0) proj_old.set_time(s.levels_t[lvl - 1] + 0.5 * s.dt);
    const double *x = &s.solution[size_t(lvl) * s.n];
    const double *xo = lvl > 0 ? &s.solution[size_t(lvl - 1) * s.n] : nullptr;
    dealii::Vector<double> v(3);
    double a0 = 0, b0 = 0, a1 = 0, b1 = 0, jj = 0;
    for (int j = 0; j < s.ny; ++j)
      for (int i = 0; i < s.nx; ++i) {
        const double p = x[s.nf + j * s.nx + i];
        const double uxl = x[j * (s.nx + 1) + i], uxr = x[j * (s.nx + 1) + i + 1];
        const double uyb = x[s.nxe + j * s.nx + i], uyt = x[s.nxe + (j + 1) * s.nx + i];
        for (int a = 0; a < 3; ++a)
          for (int b = 0; b < 3; ++b) {
            const double wq = tw[a] * tw[b] * area;
            ex.vector_value(dealii::Point<2>(s.x0 + (i + tq[a]) * s.hx, s.y0 + (j + tq[b]) * s.hy), v);
            const double uhx = (uxl * (1 - tq[a]) + uxr * tq[a]) / s.hy, uhy = (uyb * (1 - tq[b]) + uyt * tq[b]) / s.hx;
            a0 += wq * ((uhx - v(0)) * (uhx - v(0)) + (uhy - v(1)) * (uhy - v(1)));
            b0 += wq * (v(0) * v(0) + v(1) * v(1));
            a1 += wq * (p - v(2)) * (p - v(2));
            b1 += wq * v(2) * v(2);
          }
        if (lvl > 0) {  // :1728-1743
          const dealii::Point<2> c(s.x0 + (i + 0.5) * s.hx, s.y0 + (j + 0.5) * s.hy);
          const double d = xo[s.nf + j * s.nx + i] + proj_now.value(c) - p - proj_old.value(c);
          jj += d * d * area;
        }
      }
    n0[lvl] = a0; d0[lvl] = b0; n1[lvl] = a1; d1[lvl] = b1; jump[lvl] = jj;
  }
  for (int k = 0; k < 3; ++k) num[k] = den[k] = 0;
  const double final_time = s.dt * s.nt;
  for (int lvl = 0; lvl < nt; ++lvl) {
    num[0] += n0[lvl]; den[0] += d0[lvl];                // :1774-1775, no dt weight
    num[1] += s.dt * n1[lvl]; den[1] += s.dt * d1[lvl];  // :1714-1715
    num[2] += jump[lvl];
    if (std::fabs(s.levels_t[lvl] - final_time) < 1e-12) { num[2] += n1[lvl]; den[2] += d1[lvl]; }  // :1744-1749
  }
}

// compute_interface_error_l2, src/darcy_vtdd.cc:1566-1616
void lambda_error(const Sub &s, int side, const double *coef, const std::vector<int> &mortar, int k, int qdegree,
                  double &num, double &den) {
  const int ms = (side == BOTTOM || side == TOP) ? mortar[0] : mortar[1], mnt = mortar[2], nq = k + 1;
  std::vector<double> xi, w;
  gauss01(qdegree, xi, w);
  const double H = ((side == BOTTOM || side == TOP) ? s.Lx : s.Ly) / ms, T = (s.nt * s.dt) / mnt;
  PressureBoundaryValues<3> pfun;
  for (int ft = 0; ft < mnt; ++ft)
    for (int fs = 0; fs < ms; ++fs) {
      const double *c = coef + size_t(ft * ms + fs) * nq * nq;
      for (int r = 0; r < qdegree; ++r)
        for (int q = 0; q < qdegree; ++q) {
          double val = 0;
          for (int b = 0; b < nq; ++b)
            for (int a = 0; a < nq; ++a) val += c[b * nq + a] * legendre01(a, xi[q]) * legendre01(b, xi[r]);
          val /= (H * T);
          const double sc = (fs + xi[q]) * H, tc = (ft + xi[r]) * T;
          double x, y;
          if (side == BOTTOM || side == TOP) { x = s.x0 + sc; y = (side == BOTTOM) ? s.y0 : s.y0 + s.Ly; }
          else { y = s.y0 + sc; x = (side == LEFT) ? s.x0 : s.x0 + s.Lx; }
          const double pe = pfun.value(dealii::Point<3>(x, y, tc));
          const double jxw = w[q] * w[r] * H * T;
          num += (val - pe) * (val - pe) * jxw;
          den += pe * pe * jxw;
        }
    }
}

std::string pad4(int v) {
  char b[16];
  std::snprintf(b, sizeof b, "%04d", v);
  return b;
}

// compute_errors from the device-resident solution: the per-level sums are reduced by CUDA kernels
// (host/data_device.cu error_sums), the combination over the levels is the host code of subdomain_errors.
void subdomain_errors_device(const Sub &s, const double *d_solution, double num[3], double den[3]) {
  std::vector<double> sums(size_t(s.nt) * 5);
  devdata::error_sums(dev_grid(s), d_solution, sums.data());
  for (int k = 0; k < 3; ++k) num[k] = den[k] = 0;
  const double final_time = s.dt * s.nt;
  for (int lvl = 0; lvl < s.nt; ++lvl) {
    const double n0 = sums[lvl * 5 + 0], d0 = sums[lvl * 5 + 1], n1 = sums[lvl * 5 + 2], d1 = sums[lvl * 5 + 3], jump = sums[lvl * 5 + 4];
    num[0] += n0; den[0] += d0;                // :1774-1775, no dt weight
    num[1] += s.dt * n1; den[1] += s.dt * d1;  // :1714-1715
    num[2] += jump;
    if (std::fabs(s.levels_t[lvl] - final_time) < 1e-12) { num[2] += n1; den[2] += d1; }  // :1744-1749
  }
}

// ---- VTK XML "binary" data arrays: base64(header) base64(zlib blocks), header = {blocks, block size, size of the
// partial last block, compressed sizes}, UInt64 -- what deal.II's DataOut writes when built with zlib (compressor
// vtkZLibDataCompressor), cut into 1 MiB blocks that are compressed in parallel ----
std::string base64(const unsigned char *d, size_t n) {
  static const char T[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
  std::string o;
  o.resize((n + 2) / 3 * 4);
  size_t k = 0, i = 0;
  for (; i + 3 <= n; i += 3) {
    const unsigned v = (unsigned(d[i]) << 16) | (unsigned(d[i + 1]) << 8) | d[i + 2];
    o[k++] = T[v >> 18]; o[k++] = T[(v >> 12) & 63]; o[k++] = T[(v >> 6) & 63]; o[k++] = T[v & 63];
  }
  if (i < n) {
    const unsigned v = (unsigned(d[i]) << 16) | (i + 1 < n ? unsigned(d[i + 1]) << 8 : 0u);
    o[k++] = T[v >> 18]; o[k++] = T[(v >> 12) & 63]; o[k++] = (i + 1 < n) ? T[(v >> 6) & 63] : '='; o[k++] = '=';
  }
  return o;
}

constexpr size_t kVtuBlock = size_t(1) << 20;

// One data array of a .vtu file, fed in pieces (the space-time file is produced a few time levels at a time).
class VtuArray {
 public:
  void append(const void *data, size_t bytes) {
    const unsigned char *p = static_cast<const unsigned char *>(data);
    total_ += bytes;
    if (!pending_.empty()) {
      const size_t take = std::min(bytes, kVtuBlock - pending_.size());
      pending_.insert(pending_.end(), p, p + take);
      p += take; bytes -= take;
      if (pending_.size() == kVtuBlock) { compress_blocks(pending_.data(), 1); pending_.clear(); }
    }
    const size_t full = bytes / kVtuBlock;
    if (full) compress_blocks(p, full);
    pending_.insert(pending_.end(), p + full * kVtuBlock, p + bytes);
  }
  // <DataArray ...> body
  void write(std::ostream &o) {
    size_t partial = 0;
    if (!pending_.empty()) { partial = pending_.size(); compress_one(pending_.data(), pending_.size(), blocks_.emplace_back()); pending_.clear(); }
    std::vector<uint64_t> head{uint64_t(blocks_.size()), uint64_t(kVtuBlock), uint64_t(partial)};
    size_t csum = 0;
    for (const auto &b : blocks_) { head.push_back(b.size()); csum += b.size(); }
    o << base64(reinterpret_cast<const unsigned char *>(head.data()), head.size() * 8);
    std::vector<unsigned char> all;
    all.reserve(csum);
    for (const auto &b : blocks_) all.insert(all.end(), b.begin(), b.end());
    o << base64(all.data(), all.size()) << "\n";
    blocks_.clear();
  }

 private:
  static void compress_one(const unsigned char *src, size_t n, std::vector<unsigned char> &dst) {
    uLongf cap = compressBound(uLong(n));
    dst.resize(cap);
    if (compress2(dst.data(), &cap, src, uLong(n), Z_BEST_SPEED) != Z_OK) throw std::runtime_error("zlib compress2 failed");
    dst.resize(cap);
  }
  void compress_blocks(const unsigned char *src, size_t n_blocks) {
    const size_t first = blocks_.size();
    blocks_.resize(first + n_blocks);
#pragma omp parallel for schedule(dynamic)
    for (long long b = 0; b < (long long)n_blocks; ++b) compress_one(src + size_t(b) * kVtuBlock, kVtuBlock, blocks_[first + size_t(b)]);
  }
  std::vector<std::vector<unsigned char>> blocks_;
  std::vector<unsigned char> pending_;
  size_t total_ = 0;
};

const char *kVtuHead = "<?xml version=\"1.0\" ?>\n<VTKFile type=\"UnstructuredGrid\" version=\"1.0\" byte_order=\"LittleEndian\" "
                       "header_type=\"UInt64\" compressor=\"vtkZLibDataCompressor\">\n<UnstructuredGrid>\n";

bool vtu_ascii() {
  const char *e = std::getenv("MMMFE_VTU_ASCII");
  return e && e[0] == '1';
}

// patches of levels [l0, l1) expanded on the host (the builds / runs without the device data path)
void expand_space_time_host(const Sub &s, int l0, int l1, double *pts, double *u, double *p) {
#pragma omp parallel for schedule(static)
  for (int l = l0; l < l1; ++l) {
    const double *x = &s.solution[size_t(l) * s.n];
    size_t idx = size_t(l - l0) * s.np * 8;
    for (int j = 0; j < s.ny; ++j)
      for (int i = 0; i < s.nx; ++i)
        for (int v = 0; v < 8; ++v, ++idx) {
          pts[idx * 3] = s.x0 + (i + (v & 1)) * s.hx; pts[idx * 3 + 1] = s.y0 + (j + ((v >> 1) & 1)) * s.hy;
          pts[idx * 3 + 2] = (l + ((v >> 2) & 1)) * s.dt;
          u[idx * 3] = x[j * (s.nx + 1) + i + (v & 1)] / s.hy; u[idx * 3 + 1] = x[s.nxe + (j + ((v >> 1) & 1)) * s.nx + i] / s.hx;
          u[idx * 3 + 2] = 0.0;
          p[idx] = x[s.nf + j * s.nx + i];
        }
  }
}

// output_results, src/darcy_vtdd.cc:1906-1948: one patch per space-time cell, values at its 8 vertices; u3 = 0.
// d_solution != nullptr: the patches are expanded on the device from the solution that lives there, a few time
// levels at a time (host memory stays bounded); else from s.solution.  Binary + zlib like deal.II's DataOut.
void write_st_vtu(const Sub &s, const double *d_solution, const std::string &dir) {
  std::ofstream o(dir + "/space-time-plots/st_solution_d3_p" + pad4(s.r) + ".vtu");
  if (!o) return;
  const long long nc = (long long)s.nx * s.ny * s.nt;
  const bool ascii = vtu_ascii();
  const char *fmt = ascii ? "ascii" : "binary";
  VtuArray a_pts, a_u, a_p, a_conn, a_off, a_types;
  std::ostringstream t_pts, t_u, t_p;
  t_pts << std::setprecision(12); t_u << std::setprecision(12); t_p << std::setprecision(12);
  const int chunk = int(std::max<long long>(1, (long long)(size_t(32) << 20) / (8LL * 8 * s.np)));  // ~32 MiB of p per piece
  std::vector<double> pts, u, p;
  for (int l0 = 0; l0 < s.nt; l0 += chunk) {
    const int l1 = std::min(s.nt, l0 + chunk);
    const size_t cells = size_t(l1 - l0) * s.np;
    pts.resize(cells * 24); u.resize(cells * 24); p.resize(cells * 8);
    if (d_solution) devdata::expand_space_time(dev_grid(s), d_solution, l0, l1, pts.data(), u.data(), p.data());
    else expand_space_time_host(s, l0, l1, pts.data(), u.data(), p.data());
    if (ascii) {
      for (size_t k = 0; k < cells * 8; ++k) {
        t_pts << pts[3 * k] << ' ' << pts[3 * k + 1] << ' ' << pts[3 * k + 2] << '\n';
        t_u << u[3 * k] << ' ' << u[3 * k + 1] << " 0\n";
        t_p << p[k] << '\n';
      }
    } else {
      a_pts.append(pts.data(), pts.size() * 8); a_u.append(u.data(), u.size() * 8); a_p.append(p.data(), p.size() * 8);
    }
  }
  const int perm[8] = {0, 1, 3, 2, 4, 5, 7, 6};  // lexicographic -> VTK_HEXAHEDRON
  if (ascii) o << "<?xml version=\"1.0\" ?>\n<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n<UnstructuredGrid>\n";
  else o << kVtuHead;
  o << "<Piece NumberOfPoints=\"" << nc * 8 << "\" NumberOfCells=\"" << nc << "\">\n<Points>\n<DataArray type=\"Float64\" NumberOfComponents=\"3\" format=\"" << fmt << "\">\n";
  if (ascii) o << t_pts.str(); else a_pts.write(o);
  o << "</DataArray>\n</Points>\n<Cells>\n<DataArray type=\"Int64\" Name=\"connectivity\" format=\"" << fmt << "\">\n";
  if (ascii) {
    for (long long c = 0; c < nc; ++c) { for (int v = 0; v < 8; ++v) o << c * 8 + perm[v] << ' '; o << '\n'; }
  } else {
    std::vector<int64_t> buf;
    const long long step = 1 << 17;
    for (long long c0 = 0; c0 < nc; c0 += step) {
      const long long c1 = std::min(nc, c0 + step);
      buf.resize(size_t(c1 - c0) * 8);
      for (long long c = c0; c < c1; ++c) for (int v = 0; v < 8; ++v) buf[size_t(c - c0) * 8 + v] = c * 8 + perm[v];
      a_conn.append(buf.data(), buf.size() * 8);
    }
    a_conn.write(o);
  }
  o << "</DataArray>\n<DataArray type=\"Int64\" Name=\"offsets\" format=\"" << fmt << "\">\n";
  if (ascii) {
    for (long long c = 1; c <= nc; ++c) o << c * 8 << '\n';
  } else {
    std::vector<int64_t> buf;
    const long long step = 1 << 20;
    for (long long c0 = 0; c0 < nc; c0 += step) {
      const long long c1 = std::min(nc, c0 + step);
      buf.resize(size_t(c1 - c0));
      for (long long c = c0; c < c1; ++c) buf[size_t(c - c0)] = (c + 1) * 8;
      a_off.append(buf.data(), buf.size() * 8);
    }
    a_off.write(o);
  }
  o << "</DataArray>\n<DataArray type=\"UInt8\" Name=\"types\" format=\"" << fmt << "\">\n";
  if (ascii) {
    for (long long c = 0; c < nc; ++c) o << "12\n";
  } else {
    std::vector<unsigned char> buf(size_t(std::min<long long>(nc, 1 << 22)), 12);
    for (long long c0 = 0; c0 < nc; c0 += (long long)buf.size()) a_types.append(buf.data(), size_t(std::min<long long>((long long)buf.size(), nc - c0)));
    a_types.write(o);
  }
  o << "</DataArray>\n</Cells>\n<PointData Scalars=\"scalars\">\n<DataArray type=\"Float64\" Name=\"u\" NumberOfComponents=\"3\" format=\"" << fmt << "\">\n";
  if (ascii) o << t_u.str(); else a_u.write(o);
  o << "</DataArray>\n<DataArray type=\"Float64\" Name=\"p\" format=\"" << fmt << "\">\n";
  if (ascii) o << t_p.str(); else a_p.write(o);
  o << "</DataArray>\n</PointData>\n</Piece>\n</UnstructuredGrid>\n</VTKFile>\n";
}

void write_st_pvtu(int n_sub, const std::string &dir) {
  std::ofstream o(dir + "/space-time-plots/st_solution_d3.pvtu");
  if (!o) return;
  o << "<?xml version=\"1.0\"?>\n<VTKFile type=\"PUnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n<PUnstructuredGrid GhostLevel=\"0\">\n";
  o << "<PPointData Scalars=\"scalars\">\n<PDataArray type=\"Float64\" Name=\"u\" NumberOfComponents=\"3\" format=\"ascii\"/>\n<PDataArray type=\"Float64\" Name=\"p\" format=\"ascii\"/>\n</PPointData>\n";
  o << "<PPoints>\n<PDataArray type=\"Float64\" NumberOfComponents=\"3\"/>\n</PPoints>\n";
  for (int r = 0; r < n_sub; ++r) o << "<Piece Source=\"st_solution_d3_p" << pad4(r) << ".vtu\"/>\n";
  o << "</PUnstructuredGrid>\n</VTKFile>\n";
}

// per-level 2-D output, src/darcy_vtdd.cc:1872-1903
void write_level_vtu(const Sub &s, int level, const std::string &dir, int n_sub, bool master) {
  {
    std::ofstream o(dir + "/time-step-plots/solution_d2_p" + pad4(s.r) + "-" + std::to_string(level + 1) + ".vtu");
    if (!o) return;
    const long long nc = (long long)s.nx * s.ny;
    const double *x = &s.solution[size_t(level) * s.n];
    o << std::setprecision(12);
    o << "<?xml version=\"1.0\" ?>\n<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n<UnstructuredGrid>\n";
    o << "<Piece NumberOfPoints=\"" << nc * 4 << "\" NumberOfCells=\"" << nc << "\">\n<Points>\n<DataArray type=\"Float64\" NumberOfComponents=\"3\" format=\"ascii\">\n";
    for (int j = 0; j < s.ny; ++j) for (int i = 0; i < s.nx; ++i) for (int v = 0; v < 4; ++v)
      o << s.x0 + (i + (v & 1)) * s.hx << ' ' << s.y0 + (j + (v >> 1)) * s.hy << " 0\n";
    o << "</DataArray>\n</Points>\n<Cells>\n<DataArray type=\"Int64\" Name=\"connectivity\" format=\"ascii\">\n";
    const int perm[4] = {0, 1, 3, 2};
    for (long long c = 0; c < nc; ++c) { for (int v = 0; v < 4; ++v) o << c * 4 + perm[v] << ' '; o << '\n'; }
    o << "</DataArray>\n<DataArray type=\"Int64\" Name=\"offsets\" format=\"ascii\">\n";
    for (long long c = 1; c <= nc; ++c) o << c * 4 << '\n';
    o << "</DataArray>\n<DataArray type=\"UInt8\" Name=\"types\" format=\"ascii\">\n";
    for (long long c = 0; c < nc; ++c) o << "9\n";
    o << "</DataArray>\n</Cells>\n<PointData Scalars=\"scalars\">\n<DataArray type=\"Float64\" Name=\"u\" NumberOfComponents=\"3\" format=\"ascii\">\n";
    for (int j = 0; j < s.ny; ++j) for (int i = 0; i < s.nx; ++i) for (int v = 0; v < 4; ++v)
      o << x[j * (s.nx + 1) + i + (v & 1)] / s.hy << ' ' << x[s.nxe + (j + (v >> 1)) * s.nx + i] / s.hx << " 0\n";
    o << "</DataArray>\n<DataArray type=\"Float64\" Name=\"p\" format=\"ascii\">\n";
    for (int j = 0; j < s.ny; ++j) for (int i = 0; i < s.nx; ++i) for (int v = 0; v < 4; ++v) o << x[s.nf + j * s.nx + i] << '\n';
    o << "</DataArray>\n</PointData>\n</Piece>\n</UnstructuredGrid>\n</VTKFile>\n";
  }
  if (master) {
    std::ofstream o(dir + "/time-step-plots/solution_d2-" + std::to_string(level + 1) + ".pvtu");
    if (!o) return;
    o << "<?xml version=\"1.0\"?>\n<VTKFile type=\"PUnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n<PUnstructuredGrid GhostLevel=\"0\">\n";
    o << "<PPointData Scalars=\"scalars\">\n<PDataArray type=\"Float64\" Name=\"u\" NumberOfComponents=\"3\" format=\"ascii\"/>\n<PDataArray type=\"Float64\" Name=\"p\" format=\"ascii\"/>\n</PPointData>\n";
    o << "<PPoints>\n<PDataArray type=\"Float64\" NumberOfComponents=\"3\"/>\n</PPoints>\n";
    for (int r = 0; r < n_sub; ++r) o << "<Piece Source=\"solution_d2_p" << pad4(r) << "-" << level + 1 << ".vtu\"/>\n";
    o << "</PUnstructuredGrid>\n</VTKFile>\n";
  }
}

void check(mmmfe_ctx *ctx, int rc, const char *what) {
  if (rc != 0) throw std::runtime_error(std::string(what) + ": " + mmmfe_last_error(ctx));
}

}  // namespace

unsigned int count_subdomain_lines(const std::string &file_name) {
  std::ifstream f(file_name);
  if (!f.is_open()) throw std::runtime_error("cannot open " + file_name);
  std::string line;
  unsigned n = 0;
  while (std::getline(f, line)) {
    const size_t p = line.find_first_not_of(" \t");
    if (p != std::string::npos && line.compare(p, 18, "mesh_pattern_sub_d") == 0) ++n;
  }
  return n;
}

// inc/filecheck_utility.h:34-95: positional, key strings ignored
void parameter_pull_in(double &c_0, double &alpha, int &space_degree, int &mortar_degree, int &num_refinement,
                       double &final_time, double &tolerence, int &max_iteration, bool &need_each_time_step_plot,
                       std::vector<char> &bc_con, std::vector<double> &nm_bc_con_funcs, bool &is_manufact_solution,
                       std::vector<std::vector<int>> &mesh_m3d, unsigned int n_processes, std::string file_name) {
  std::string dummy;
  std::ifstream in(file_name);
  if (!in.is_open()) throw std::runtime_error("parameter file " + file_name + " could not be opened");
  for (int skip = 0; skip < 4; ++skip) std::getline(in, dummy);
  in >> dummy >> c_0; std::getline(in, dummy);
  in >> dummy >> alpha; std::getline(in, dummy);
  in >> dummy >> space_degree; std::getline(in, dummy);
  in >> dummy >> mortar_degree; std::getline(in, dummy);
  in >> dummy >> num_refinement; std::getline(in, dummy);
  in >> dummy >> final_time; std::getline(in, dummy);
  in >> dummy >> tolerence; std::getline(in, dummy);
  in >> dummy >> max_iteration; std::getline(in, dummy);
  in >> dummy >> need_each_time_step_plot; std::getline(in, dummy);
  bc_con.resize(4); nm_bc_con_funcs.resize(4);
  for (int k = 0; k < 4; ++k) in >> dummy >> bc_con[k] >> nm_bc_con_funcs[k];
  std::getline(in, dummy);
  in >> dummy >> is_manufact_solution; std::getline(in, dummy);
  mesh_m3d.assign(n_processes + 1, std::vector<int>(3, 0));
  for (unsigned sub = 0; sub < n_processes + 1; ++sub) in >> dummy >> mesh_m3d[sub][0] >> mesh_m3d[sub][1] >> mesh_m3d[sub][2];
  if (!in) throw std::runtime_error("parameter file " + file_name + " is incomplete");
  for (char b : bc_con)
    if (b != 'D' && b != 'N')
      throw std::runtime_error("incompatible boundary condition read from parameter file: provide either D or N");
}

// Host-only diagnostic: writes what the front-end would hand to the engine for subdomain r of a cycle as raw
// little-endian arrays under `dir` (used by the CPU tests to compare with the oracle; no device needed).
void dump_subdomain(const std::vector<std::vector<int>> &reps, int r, double final_time, double c0, double alpha,
                    int mortar_degree, bool manufact, const std::vector<char> &bc_type,
                    const std::vector<double> &bc_val, const std::string &dir) {
  const unsigned n_sub = unsigned(reps.size()) - 1;
  unsigned n0, n1;
  find_divisors(n_sub, n0, n1);
  Sub s;
  init_sub(s, r, int(n0), int(n1), reps[r], final_time);
  assemble_matrix(s, c0);
  constrain_matrix(s, manufact, bc_type, bc_val);
  auto wr = [&](const std::string &name, const void *p, size_t bytes) {
    std::ofstream f(dir + "/" + name, std::ios::binary);
    f.write(static_cast<const char *>(p), std::streamsize(bytes));
  };
  auto wr_csr = [&](const std::string &name, const Csr &m) {
    const long long shape[2] = {m.n_rows, m.n_cols};
    wr(name + ".shape", shape, sizeof shape);
    wr(name + ".rp", m.rp.data(), m.rp.size() * sizeof(long long));
    wr(name + ".col", m.col.data(), m.col.size() * sizeof(int));
    wr(name + ".val", m.val.data(), m.val.size() * sizeof(double));
  };
  wr_csr("matrix", s.matrix);
  for (int side = 0; side < 4; ++side)
    if (s.nb[side] >= 0) {
      projector_tables(s, side, reps.back(), mortar_degree);
      wr_csr("f2c" + std::to_string(side), s.f2c[side]);
      wr_csr("c2f" + std::to_string(side), s.c2f[side]);
    }
  BarData bd;
  bar_data(s, mortar_degree + 1, manufact, bc_type, bc_val, c0, alpha, bd);
  wr("p0", bd.p0.data(), bd.p0.size() * sizeof(double));
  wr("src", bd.src.data(), bd.src.size() * sizeof(double));
  wr("fr_dofs", bd.fr_dofs.data(), bd.fr_dofs.size() * sizeof(int));
  wr("fr_vals", bd.fr_vals.data(), bd.fr_vals.size() * sizeof(double));
  const int nb[4] = {s.nb[0], s.nb[1], s.nb[2], s.nb[3]};
  wr("neighbors", nb, sizeof nb);
}

template <int dim>
DarcyVTProblem<dim>::DarcyVTProblem(const unsigned int degree_, const BiotParameters &bprm, const unsigned int mortar_flag_,
                                    const unsigned int mortar_degree_, std::vector<char> bc, std::vector<double> bcv,
                                    const bool manufact, const bool plots)
    : prm(bprm), bc_condition_vect(bc), bc_const_functs(bcv), is_manufact_solution(manufact),
      need_each_time_step_plot(plots), degree(degree_), mortar_degree(mortar_degree_), mortar_flag(mortar_flag_) {}

template <int dim>
void DarcyVTProblem<dim>::run(const unsigned int refine, const std::vector<std::vector<int>> &reps_st, double tol,
                              unsigned int maxiter, unsigned int quad_degree) {
  const int rank = control.rank, world = control.world;
  const bool root = rank == 0;
  const bool say = root && control.verbose;
  if (reps_st.empty() || reps_st[0].size() != 3) throw std::runtime_error("mesh patterns must have 3 entries");
  const unsigned n_sub = unsigned(reps_st.size()) - 1;
  if (say) std::cout << "\n\n Total number of processes is " << n_sub << "\n\n";
  if (mortar_flag != 1) throw std::runtime_error("only the mortar path (mortar_flag = 1) exists, as in the reference driver");
  if (n_sub <= 1) throw std::runtime_error("Mortar MFEM is impossible with 1 subdomain");  // :2066
  if (degree != 0) throw std::runtime_error("space_degree >= 1 is not valid on the space-time mortar path (see DESIGN.md, quirk Q7)");
  // every rank derives the same ownership table, so this fails on all ranks at once (no rank is left in a collective)
  if (world > int(n_sub)) throw std::runtime_error("more ranks than subdomains: launch at most one process per subdomain");
  const int qdegree = int(quad_degree);
  // The ranks of one node share its cores: every rank takes its share of them for the host-side loops (matrix and
  // projector tables, right-hand-side data, error norms).  torchrun exports OMP_NUM_THREADS=1 to its workers, which
  // made the bar data of two ranks 6x slower than that of one (30 s against 5 s): with several ranks the share is set
  // explicitly; $MMMFE_HOST_THREADS overrides.
  if (const char *ht = std::getenv("MMMFE_HOST_THREADS")) {
    omp_set_num_threads(std::max(1, std::atoi(ht)));
  } else if (world > 1) {
    int local_world = world;
    if (const char *lw = std::getenv("LOCAL_WORLD_SIZE")) local_world = std::max(1, std::atoi(lw));
    omp_set_num_threads(std::max(1, omp_get_num_procs() / local_world));
  }
  unsigned n0, n1;
  find_divisors(n_sub, n0, n1);
  std::vector<std::vector<int>> reps = reps_st;
  reports_.clear();

  for (unsigned cycle = 0; cycle < refine; ++cycle) {
    if (cycle > 0) {  // src/darcy_vtdd.cc:2124-2142
      for (unsigned i = 0; i + 1 < reps.size(); ++i) for (int d = 0; d < 3; ++d) reps[i][d] *= 2;
      if (mortar_degree == 1 || cycle % 2 == 0) for (int d = 0; d < 3; ++d) reps.back()[d] *= 2;
    }
    if (control.only_cycle >= 0 && int(cycle) != control.only_cycle) continue;
    CycleReport rep;
    rep.cycle = int(cycle);
    const std::vector<int> &mortar = reps.back();
    if (say) {
      std::cout << "Final time= " << prm.final_time << "\n";
      std::cout << "Mortar mesh has " << (long long)mortar[0] * mortar[1] * mortar[2] << " cells" << std::endl;
      std::cout << "Making grid and DOFs...\n";
    }
    // ---- ownership: contiguous blocks of subdomains balanced by nt * cells ----
    std::vector<int> owner(n_sub, 0);
    {
      std::vector<double> cost(n_sub);
      for (unsigned r = 0; r < n_sub; ++r) cost[r] = double(reps[r][0]) * reps[r][1] * reps[r][2];
      if (mmmfe_assign_owners(int(n_sub), cost.data(), world, owner.data()) != 0) throw std::runtime_error("mmmfe_assign_owners failed");
    }
    // ---- engine context first: with several ranks NCCL is initialised and its paths are opened by a helper thread
    // WHILE this thread assembles the matrices (communicator set-up + first-use channel opening cost 2-3 s on 8 ranks
    // and would otherwise sit in front of / inside mmmfe_setup)
    mmmfe_ctx *ctx = nullptr;
    if (mmmfe_create(&ctx, control.device) != 0)
      throw std::runtime_error("no usable CUDA device: the interface-GMRES path has no CPU fallback");
    struct Guard { mmmfe_ctx *c; ~Guard() { mmmfe_destroy(c); } } guard{ctx};
    std::thread comm_thread;
    int comm_rc = 0;
    std::string comm_what;
    struct Joiner { std::thread &t; ~Joiner() { if (t.joinable()) t.join(); } } joiner{comm_thread};  // (before guard runs)
    if (world > 1) {
      std::vector<int32_t> peers;
      for (unsigned r = 0; r < n_sub; ++r) {
        if (owner[r] != rank) continue;
        int nb[4];
        find_neighbors(int(r), int(n0), int(n1), nb);
        for (int k = 0; k < 4; ++k)
          if (nb[k] >= 0 && owner[nb[k]] != rank && std::find(peers.begin(), peers.end(), owner[nb[k]]) == peers.end())
            peers.push_back(owner[nb[k]]);
      }
      comm_thread = std::thread([this, ctx, rank, world, peers, &comm_rc, &comm_what] {
        comm_rc = mmmfe_comm_init(ctx, rank, world, control.nccl_id);
        comm_what = "mmmfe_comm_init";
        if (comm_rc == 0) { comm_rc = mmmfe_comm_warmup(ctx, int32_t(peers.size()), peers.data()); comm_what = "mmmfe_comm_warmup"; }
      });
    }
    auto t0 = clk::now();
    std::vector<std::unique_ptr<Sub>> subs;
    for (unsigned r = 0; r < n_sub; ++r) {
      if (owner[r] != rank) continue;
      auto s = std::make_unique<Sub>();
      init_sub(*s, int(r), int(n0), int(n1), reps[r], prm.final_time);
      subs.push_back(std::move(s));
    }
    const bool host_prof = std::getenv("MMMFE_SETUP_PROFILE") != nullptr;
    for (auto &sp : subs) {  // row-parallel inside
      assemble_matrix(*sp, prm.c_0);
      constrain_matrix(*sp, is_manufact_solution, bc_condition_vect, bc_const_functs);
    }
    if (host_prof) std::fprintf(stderr, "[setup profile] host: saddle matrices %8.1f ms\n", since(t0) * 1e3);
    {
      // outer sides too when several ranks run: lets the ranks share the multiscale-basis work of a factor group
      std::vector<std::pair<int, int>> jobs;
      for (int li = 0; li < int(subs.size()); ++li)
        for (int side = 0; side < 4; ++side)
          if (subs[li]->nb[side] >= 0 || world > 1) jobs.push_back({li, side});
      const auto t1 = clk::now();
      // an invalid MMMFE_TIE_RULE is reported here, not from inside the parallel loop
      if (tie_rule_dealii91() && mortar_degree % 2 == 0 && mortar_degree != 2)
        throw std::runtime_error("MMMFE_TIE_RULE=dealii91 is restated for mortar_degree 2 only");
#pragma omp parallel for schedule(dynamic)
      for (int q = 0; q < int(jobs.size()); ++q) projector_tables(*subs[jobs[q].first], jobs[q].second, mortar, int(mortar_degree));
      if (host_prof) std::fprintf(stderr, "[setup profile] host: projector tables %8.1f ms\n", since(t1) * 1e3);
    }
    rep.t_assemble = since(t0);
    if (say && !subs.empty()) {
      std::cout << "N flux subdom dofs: " << subs[0]->nf << std::endl;
      std::cout << "N pressure subdom dofs: " << subs[0]->np << std::endl;
      std::cout << "Assembling system...\n";
    }
    // ---- engine ----
    mmmfe_set_option(ctx, "keep_history", control.final_sweep ? 1 : 0);
    mmmfe_set_option(ctx, "mortar_time_slabs", mortar[2]);  // lets the engine build the multiscale-basis operator
    {
      std::string all;
      if (const char *env = std::getenv("MMMFE_ENGINE_OPTIONS")) all = env;
      if (!control.engine_options.empty()) all += (all.empty() ? "" : ",") + control.engine_options;
      std::stringstream ss(all);
      std::string item;
      while (std::getline(ss, item, ',')) {
        const size_t eq = item.find('=');
        if (item.empty()) continue;
        if (eq == std::string::npos) throw std::runtime_error("engine option '" + item + "' is not key=value");
        check(ctx, mmmfe_set_option(ctx, item.substr(0, eq).c_str(), std::atof(item.c_str() + eq + 1)), "mmmfe_set_option");
      }
    }
    if (world > 1) {
      comm_thread.join();
      check(ctx, comm_rc, comm_what.c_str());
      check(ctx, mmmfe_set_owners(ctx, int(n_sub), owner.data()), "mmmfe_set_owners");
    }
    for (auto &sp : subs) {
      Sub &s = *sp;
      mmmfe_subdomain d{};
      d.global_id = s.r; d.nx = s.nx; d.ny = s.ny; d.nt = s.nt; d.dt = s.dt;
      d.matrix = s.matrix.view(); d.n_flux = s.nf; d.n_pressure = s.np; d.canon_of_dof = nullptr;
      for (int k = 0; k < 4; ++k) {
        d.neighbor[k] = s.nb[k]; d.n_side[k] = s.n_side[k]; d.side_dofs[k] = s.side_edges[k].data();
        d.mortar_len[k] = s.mortar_len[k];
        if (s.nb[k] >= 0 || (world > 1 && !s.f2c[k].rp.empty())) { d.f2c[k] = s.f2c[k].view(); d.c2f[k] = s.c2f[k].view(); }
      }
      check(ctx, mmmfe_add_subdomain(ctx, &d, &s.local), "mmmfe_add_subdomain");
    }
    t0 = clk::now();
    check(ctx, mmmfe_setup(ctx), "mmmfe_setup");
    rep.t_factor = since(t0);
    if (say) std::cout << "  ...factorized..." << "\n";
    rep.interface_length = mmmfe_interface_length(ctx);
    rep.step_bytes = mmmfe_step_bytes(ctx);
    rep.factor_values = mmmfe_factor_values(ctx);
    rep.factor_flops = mmmfe_factor_flops(ctx);
    rep.factor_groups = mmmfe_num_factor_groups(ctx);
    rep.sweep_batches = int(mmmfe_stat(ctx, "sweep_batches"));
    rep.tune_rel_diff_e18 = mmmfe_stat(ctx, "tune_rel_diff_e18");
    rep.basis_operator = int(mmmfe_stat(ctx, "basis_operator"));
    rep.t_basis = double(mmmfe_stat(ctx, "basis_us")) * 1e-6;
    rep.basis_gflop = double(mmmfe_stat(ctx, "basis_mflop")) * 1e-3;
    rep.basis_bytes = double(mmmfe_stat(ctx, "basis_mbytes")) * 1e6;
    rep.basis_columns = mmmfe_stat(ctx, "basis_columns");
    rep.basis_shared_by = int(std::max<long long>(1, mmmfe_stat(ctx, "basis_shared_by")));
    rep.basis_steps = mmmfe_stat(ctx, "basis_steps"); rep.basis_delays = mmmfe_stat(ctx, "basis_delays");
    rep.basis_mw = mmmfe_stat(ctx, "basis_mw"); rep.toeplitz_tiles = mmmfe_stat(ctx, "toeplitz_tiles");
    rep.toeplitz_gflop = double(mmmfe_stat(ctx, "toeplitz_mflop")) * 1e-3;
    rep.toeplitz_mbytes = double(mmmfe_stat(ctx, "toeplitz_mbytes"));
    for (auto &sp : subs) { rep.dof_steps += (long long)sp->n * sp->nt; sp->matrix = Csr(); }
    // ---- bar problems ----
    t0 = clk::now();
    const bool dev_data = devdata::enabled();  // the data classes are evaluated by CUDA kernels (host/data_device.cu)
    for (auto &sp : subs) {
      if (dev_data) {
        bar_data_device(ctx, *sp, qdegree, is_manufact_solution, bc_condition_vect, bc_const_functs, prm.c_0, prm.alpha);
        continue;
      }
      BarData bd;
      bar_data(*sp, qdegree, is_manufact_solution, bc_condition_vect, bc_const_functs, prm.c_0, prm.alpha, bd);
      check(ctx, mmmfe_bar_set_data(ctx, sp->local, bd.p0.data(), bd.src.data(), int(bd.fr_dofs.size()),
                                    bd.fr_dofs.data(), bd.fr_vals.data()), "mmmfe_bar_set_data");
    }
    rep.t_bar_data = since(t0);
    if (world > 1) {  // meet the other ranks (and let NCCL connect) before the sweep is timed
      double one = 1.0;
      check(ctx, mmmfe_comm_allreduce_sum(ctx, &one, 1), "mmmfe_comm_allreduce_sum");
    }
    t0 = clk::now();
    check(ctx, mmmfe_bar_sweep(ctx), "mmmfe_bar_sweep");
    rep.t_bar = since(t0);
    if (say) std::cout << "\nStarting GMRES iterations.........\n";
    // ---- interface GMRES ----
    const bool fixed = control.fixed_iterations >= 0;
    const int max_it = fixed ? std::max(1, control.fixed_iterations) : int(maxiter);
    mmmfe_gmres_info info{};
    std::vector<double> hist(size_t(max_it) + 2, 0.0);
    rep.lambda.assign(size_t(rep.interface_length), 0.0);
    if (fixed && control.warmup_iterations > 0)
      check(ctx, mmmfe_gmres_solve(ctx, 0.0, control.warmup_iterations, &info, nullptr, 0, nullptr), "gmres warm-up");
    const long long launches_before = mmmfe_kernel_launches(ctx);
    t0 = clk::now();
    check(ctx, mmmfe_gmres_solve(ctx, fixed ? 0.0 : tol, max_it, &info, hist.data(), int(hist.size()), rep.lambda.data()),
          "mmmfe_gmres_solve");
    rep.t_gmres = since(t0);
    rep.gmres_kernel_launches = mmmfe_kernel_launches(ctx) - launches_before;
    rep.gmres_iterations = info.iterations; rep.converged = info.converged; rep.r_norm = info.r_norm;
    rep.gmres_device_ms = info.device_ms;
    rep.residuals.assign(hist.begin(), hist.begin() + std::min<size_t>(hist.size(), size_t(info.iterations) + 1));
    if (say) {
      std::cout << "\n\n r_norm is " << info.r_norm << " target is " << info.r_norm * tol << "\n\n";
      if (info.converged)
        std::cout << "\n  GMRES converges in " << info.iterations << " iterations!\n and residual is "
                  << info.final_residual * info.r_norm << "\n";
      else if (!fixed)
        std::cout << "\n  GMRES doesn't converge after  " << info.iterations << " iterations!\n";
    }
    if (control.apply_repeat > 0)
      check(ctx, mmmfe_apply_interface_operator_device(ctx, control.apply_repeat, &control.apply_ms), "apply");
    if (control.profile && rep.basis_operator)
      check(ctx, mmmfe_profile_basis_apply(ctx, 20, &rep.toeplitz_ms), "profile basis apply");
    if (control.profile) {
      check(ctx, mmmfe_profile_step(ctx, rep.prof_ms, rep.prof_bytes, (int64_t *)rep.prof_launches), "profile");
      int32_t nsw = 0;
      check(ctx, mmmfe_profile_sweeps(ctx, 4, rep.sweep_ms, rep.sweep_bytes_per_step, rep.sweep_steps, &nsw), "profile sweeps");
      rep.sweep_batches = nsw;
      rep.tune_us_sweep = mmmfe_stat(ctx, "tune_us_sweep");
      rep.tune_us_levels = mmmfe_stat(ctx, "tune_us_levels");
      rep.sweep_planned = mmmfe_stat(ctx, "sweep_planned");
    }
    if (control.e2e_iterations > 0) {
      // The same iteration driven from host memory: every step copies the Krylov vector host -> device from a
      // pinned buffer, applies the operator, copies A q back and does the Arnoldi update on the host
      // (what local_gmres does between its MPI calls, src/darcy_vtdd.cc:1413-1463).
      const long long len = rep.interface_length;
      const int ne = control.e2e_iterations;
      double *q = static_cast<double *>(mmmfe_alloc_pinned((len + 1) * sizeof(double)));
      double *ap = static_cast<double *>(mmmfe_alloc_pinned((len + 1) * sizeof(double)));
      if (!q || !ap) throw std::runtime_error("pinned allocation failed");
      std::vector<std::vector<double>> Q;
      std::vector<double> r(size_t(len) + 1);
      check(ctx, mmmfe_get_initial_residual(ctx, r.data()), "initial residual");
      // (the host loops below run on all threads of the rank, in fixed chunks: reproducible run to run)
      auto dot = [&](const double *x, const double *y) {
        const int nchunk = 64;
        double part[nchunk];
#pragma omp parallel for schedule(static)
        for (int cidx = 0; cidx < nchunk; ++cidx) {
          const long long j0 = len * cidx / nchunk, j1 = len * (cidx + 1) / nchunk;
          double a = 0;
#pragma omp simd reduction(+ : a)
          for (long long j = j0; j < j1; ++j) a += x[j] * y[j];
          part[cidx] = a;
        }
        double a = 0;
        for (int cidx = 0; cidx < nchunk; ++cidx) a += part[cidx];
        return a;
      };
      double nn[1] = {dot(r.data(), r.data())};
      check(ctx, mmmfe_comm_allreduce_sum(ctx, nn, 1), "allreduce");
      Q.emplace_back(r.begin(), r.begin() + len);
      for (auto &v : Q[0]) v /= std::sqrt(nn[0]);
      double total = 0;
      for (int k = 0; k < ne + 1; ++k) {  // iteration 0 is an untimed warm-up
        const auto ti = clk::now();
        std::copy(Q[k].begin(), Q[k].end(), q);
        check(ctx, mmmfe_apply_interface_operator(ctx, q, ap), "apply (host buffers)");
        std::vector<double> h(size_t(k) + 2, 0.0);
        for (int i = 0; i <= k; ++i) h[i] = dot(ap, Q[i].data());
        check(ctx, mmmfe_comm_allreduce_sum(ctx, h.data(), k + 1), "allreduce");
#pragma omp parallel for schedule(static)
        for (long long j = 0; j < len; ++j) {
          double v = ap[j];
          for (int i = 0; i <= k; ++i) v -= h[i] * Q[i][j];
          ap[j] = v;
        }
        double n2[1] = {dot(ap, ap)};
        check(ctx, mmmfe_comm_allreduce_sum(ctx, n2, 1), "allreduce");
        const double hk = std::sqrt(n2[0]);
        Q.emplace_back(size_t(len));
        {
          double *qn = Q.back().data();
#pragma omp parallel for schedule(static)
          for (long long j = 0; j < len; ++j) qn[j] = ap[j] / hk;
        }
        if (k > 0) total += since(ti);
      }
      rep.e2e_ms_per_iteration = 1e3 * total / ne;
      rep.e2e_h2d_bytes = len * 8;
      rep.e2e_d2h_bytes = len * 8;
      mmmfe_free_pinned(q);
      mmmfe_free_pinned(ap);
    }
    if (fixed && control.also_solve) {  // the converged solve, timed in the same context (lambda of THIS solve is kept)
      mmmfe_gmres_info info2{};
      t0 = clk::now();
      check(ctx, mmmfe_gmres_solve(ctx, tol, int(maxiter), &info2, nullptr, 0, rep.lambda.data()), "mmmfe_gmres_solve (converged)");
      rep.t_solve_gmres = since(t0);
      rep.solve_iterations = info2.iterations; rep.solve_converged = info2.converged;
      rep.solve_final_residual = info2.final_residual;
    }
    // ---- final sweep, errors, output ----
    if (control.final_sweep) {
      t0 = clk::now();
      check(ctx, mmmfe_final_sweep(ctx), "mmmfe_final_sweep");
      // With the device data path the solution stays on the GPU: the essential values are set there, the error sums
      // are reduced there and the output patches are expanded there.  The host copy is only made for the host
      // paths (no device data, or the per-level 2-D plots).
      std::vector<double *> d_sol(subs.size(), nullptr);
      if (dev_data)
        for (size_t li = 0; li < subs.size(); ++li) {
          Sub &sd = *subs[li];
          check(ctx, mmmfe_solution_device(ctx, sd.local, &d_sol[li]), "mmmfe_solution_device");
          devdata::set_constant_rows(sd.nt, sd.n, int(sd.con_dofs.size()), sd.con_dofs.data(), sd.con_vals.data(), d_sol[li]);
        }
      const bool want_files = control.write_output && cycle + 1 == refine;  // :1847
      const bool host_copy = control.fetch_solution && (!dev_data || (want_files && need_each_time_step_plot));
      if (host_copy)
      for (auto &sp : subs) {
        sp->solution.resize(size_t(sp->nt) * sp->n);
        check(ctx, mmmfe_get_solution(ctx, sp->local, sp->solution.data()), "mmmfe_get_solution");
        for (int lvl = 0; lvl < sp->nt; ++lvl)  // constraint_bc.distribute (:866): the essential values, exactly
          for (size_t k = 0; k < sp->con_dofs.size(); ++k) sp->solution[size_t(lvl) * sp->n + sp->con_dofs[k]] = sp->con_vals[k];
      }
      rep.t_final = since(t0);
      if (say) std::cout << "finished local_gmres" << "\n";
      if (is_manufact_solution && control.compute_errors && (dev_data || control.fetch_solution)) {
        t0 = clk::now();
        double buf[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // num{vel, pL2, DG, lambda}, den{...}
        for (size_t li = 0; li < subs.size(); ++li) {
          Sub *sp = subs[li].get();
          double num[3], den[3];
          if (dev_data) subdomain_errors_device(*sp, d_sol[li], num, den);
          else subdomain_errors(*sp, num, den);
          for (int k = 0; k < 3; ++k) { buf[k] += num[k]; buf[4 + k] += den[k]; }
          for (int side = 0; side < 4; ++side)
            if (sp->nb[side] >= 0) {
              const long long off = mmmfe_interface_offset(ctx, sp->local, side);
              lambda_error(*sp, side, rep.lambda.data() + off, mortar, int(mortar_degree), qdegree, buf[3], buf[7]);
            }
        }
        if (world > 1) check(ctx, mmmfe_comm_allreduce_sum(ctx, buf, 8), "allreduce");  // MPI_Reduce, :1812-1813
        rep.velocity_l2l2 = std::sqrt(buf[0]) / std::sqrt(buf[4]);
        rep.pressure_l2l2 = std::sqrt(buf[1]) / std::sqrt(buf[5]);
        rep.pressure_dg = std::sqrt(buf[2]) / std::sqrt(buf[6]);
        rep.lambda_l2 = std::sqrt(buf[3]) / std::sqrt(buf[7]);
        rep.t_errors = since(t0);
      }
      if (want_files && (dev_data || control.fetch_solution)) {
        t0 = clk::now();
        for (size_t li = 0; li < subs.size(); ++li) {
          Sub *sp = subs[li].get();
          write_st_vtu(*sp, dev_data ? d_sol[li] : nullptr, control.output_dir);
          if (need_each_time_step_plot)
            for (int l = 0; l < sp->nt; ++l) write_level_vtu(*sp, l, control.output_dir, int(n_sub), root && sp->local == 0);
        }
        if (root) write_st_pvtu(int(n_sub), control.output_dir);
        rep.t_output = since(t0);
      }
    }
    rep.kernel_launches = mmmfe_kernel_launches(ctx);
    if (!control.keep_lambda) rep.lambda.clear();
    reports_.push_back(rep);
  }
  // convergence table, src/darcy_vtdd.cc:1957-1996
  if (root && is_manufact_solution && control.compute_errors && control.final_sweep && !reports_.empty()) {
    std::ostringstream txt, tex;
    txt << "cycle # GMRES Velocity,L2-L2 Pressure,DG Pressure,L2-L2 Lambda,Darcy_L2\n";
    tex << "\\begin{table}[H]\n\\begin{center}\n\\begin{tabular}{|c|c|c|c|c|c|c|c|c|c|c|}\\hline\ncycle & \\# gmres & & $ \\|z - z_h\\|_{L^{2}(L^2)} $ & & $ \\|p - p_h\\|_{DG} $ & & $ \\|p - p_h\\|_{L^{2}(L^2)} $ & & $ \\|p - \\lambda_p_H\\|_{L^{2}(L^2)} $ & \\\\ \\hline\n";
    for (size_t i = 0; i < reports_.size(); ++i) {
      const CycleReport &r = reports_[i];
      auto rate = [&](double cur, double prev) {
        std::ostringstream s;
        if (i == 0) s << "-"; else s << std::fixed << std::setprecision(2) << std::log2(prev / cur);
        return s.str();
      };
      const CycleReport &p = reports_[i ? i - 1 : 0];
      std::ostringstream row;
      row << std::scientific << std::setprecision(3);
      txt << r.cycle << " " << r.gmres_iterations << " " << rate(r.gmres_iterations, p.gmres_iterations) << " ";
      txt << std::scientific << std::setprecision(3) << r.velocity_l2l2 << " " << rate(r.velocity_l2l2, p.velocity_l2l2) << " "
          << r.pressure_dg << " " << rate(r.pressure_dg, p.pressure_dg) << " " << r.pressure_l2l2 << " "
          << rate(r.pressure_l2l2, p.pressure_l2l2) << " " << r.lambda_l2 << " " << rate(r.lambda_l2, p.lambda_l2) << "\n";
      tex << r.cycle << " & " << r.gmres_iterations << " & " << rate(r.gmres_iterations, p.gmres_iterations) << " & "
          << std::scientific << std::setprecision(3) << r.velocity_l2l2 << " & " << rate(r.velocity_l2l2, p.velocity_l2l2)
          << " & " << r.pressure_dg << " & " << rate(r.pressure_dg, p.pressure_dg) << " & " << r.pressure_l2l2 << " & "
          << rate(r.pressure_l2l2, p.pressure_l2l2) << " & " << r.lambda_l2 << " & " << rate(r.lambda_l2, p.lambda_l2)
          << "\\\\ \\hline\n";
    }
    tex << "\\end{tabular}\n\\end{center}\n\\end{table}\n";
    if (control.verbose) std::cout << std::endl << txt.str();
    if (control.write_output) {
      std::ofstream f(control.output_dir + "/error" + std::to_string(n_sub) + "domains.tex");
      f << tex.str();
    }
  }
}

template class DarcyVTProblem<2>;

}  // namespace vt_darcy
