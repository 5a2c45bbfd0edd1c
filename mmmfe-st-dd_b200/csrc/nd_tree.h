// Geometric nested dissection of one structured rectangular RT0 subdomain (host side, symbolic only).
//
// Unknowns are the flux dofs (edges) of the nx x ny cell grid after the cell pressures have been eliminated
// (DGQ0 mass matrix is diagonal): S = A + dt/c0 B^T M^-1 B couples two edges iff they bound a common cell, so a
// line of ny x-edges (or nx y-edges) separates the grid exactly like a line of nodes separates a 5-point grid.
// The tree replaces the ordering step inside SparseDirectUMFPACK::initialize (src/darcy_vtdd.cc:483).
#pragma once
#include <cstdint>
#include <vector>

namespace mmmfe {

struct NdNode {
  int32_t s = 0;            // fully summed (separator / leaf-interior) variables
  int32_t b = 0;            // boundary variables = box perimeter edges owned by ancestors
  int32_t child[2] = {-1, -1};
  int32_t parent = -1;
  int32_t height = 0;       // leaves 0; forward solve runs heights ascending, backward descending
  int32_t nsplit = 1;       // row splits of the forward update (partial contribution vectors)
  int64_t idx_off = 0;      // into NdTree::idx (s + b canonical flux indices, sep first)
  int64_t src_off = 0;      // into NdTree::src (2 * (s + b) entries: child-0 / child-1 source index or -1)
  int64_t r_off = 0;        // into the factor storage: s rows of ldr doubles, R = [F11^-1 | -G^T]
  int32_t ldr = 0;          // row stride of R (>= s + b, even)
  int64_t z_off = 0;        // into the per-front vector storage (s + nsplit * b entries per right-hand side)
  int64_t f_off = 0;        // into the factorisation workspace ((s + b)^2 doubles)
};

struct NdTree {
  int32_t nx = 0, ny = 0;
  int64_t n_flux = 0;
  std::vector<NdNode> nodes;        // post-order: every subtree is a contiguous index range
  std::vector<int32_t> idx;         // canonical flux index of every front variable
  std::vector<int32_t> src;         // gather maps from the children's boundary entries
  std::vector<std::vector<int32_t>> by_height;
  std::vector<int32_t> node_of;     // canonical flux index -> node that eliminates it
  std::vector<int32_t> pos_of;      // canonical flux index -> position in that node's separator list
  int64_t r_total = 0, z_total = 0, f_total = 0;
  double factor_flops = 0.0;

  // leaf_cells: boxes with at most this many cells are eliminated as one dense block.
  void build(int32_t nx, int32_t ny, int32_t leaf_cells = 4);
};

}  // namespace mmmfe
