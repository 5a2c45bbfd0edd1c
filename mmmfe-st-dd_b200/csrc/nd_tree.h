// Geometric nested dissection of one structured rectangular RT0 subdomain (host side, symbolic only).
//
// Unknowns are the flux dofs (edges) of the nx x ny cell grid after the cell pressures have been eliminated
// (DGQ0 mass matrix is diagonal): S = A + dt/c0 B^T M^-1 B couples two edges iff they bound a common cell, so a
// line of ny x-edges (or nx y-edges) separates the grid exactly like a line of nodes separates a 5-point grid.
// The tree replaces the ordering step inside SparseDirectUMFPACK::initialize (src/darcy_vtdd.cc:483).
//
// The tree is cut into two parts for the solve:
//   * SUBTREES: maximal boxes whose whole factor fits in shared memory; one CTA streams the subtree's factor
//     blob (contiguous in HBM) and walks its levels on chip.  Subtrees of equal shape share one TEMPLATE (all the
//     index structure), so the only per-subtree HBM traffic is the factor itself and the vectors.
//   * TOP nodes: everything above, processed level by level by grid-wide kernels.
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
  int32_t nsplit = 1;       // row splits of the forward update of a top node (partial contribution vectors)
  int32_t i0 = 0, i1 = 0, j0 = 0, j1 = 0;  // cell box
  int32_t subtree = -1;     // instance index when the node lives inside a subtree
  int32_t tnode = -1;       // node index inside the subtree template
  int64_t idx_off = 0;      // into NdTree::idx (s + b canonical flux indices, sep first)
  int64_t src_off = 0;      // into NdTree::src (2 * (s + b) entries: child-0 / child-1 source index or -1)
  // factor storage (offsets in doubles into one allocation)
  //   top node    : rb_off -> R  = [F11^-1 | -G^T]  s rows x ldr      rf_off -> -G^T  s rows x ldf (forward copy)
  //   subtree node: rb_off -> R^T  (s + b) rows x s                   rf_off -> -G^T  s rows x b
  int64_t rb_off = 0, rf_off = 0;
  int32_t ldr = 0, ldf = 0;
  int64_t z_off = 0;        // into the per-front vector storage (s + nsplit * b entries per right-hand side)
  int64_t f_off = 0;        // into the factorisation workspace ((s + b)^2 doubles)
};

// Index structure of one subtree shape; every array is shared by all subtrees of that shape.
struct SubtreeTemplate {
  int32_t w = 0, h = 0, flags = 0;
  int32_t n_nodes = 0, n_levels = 0;
  int32_t n_int = 0;     // variables eliminated inside the subtree
  int32_t n_rootb = 0;   // boundary variables of the subtree root
  int32_t zl_size = 0;   // sum of b over nodes = boundary-output entries
  int32_t rf_size = 0, rb_size = 0;  // doubles per subtree in the forward / backward blob (even)
  // per node (processing order: heights ascending)
  std::vector<int32_t> node_s, node_b, node_rf, node_rb, node_sep, node_bnd, node_zl;
  // per interior variable: code of its canonical index relative to the root box (type<<30 | dj<<15 | di)
  std::vector<int32_t> var_code;
  std::vector<int32_t> rootb_code;
  // forward pass 1 (separator accumulation), flat over nodes (flat index = local variable id): owning node and
  // absolute zl index of the two children's contribution (-1 = none)
  std::vector<int32_t> sep_node, sep_c0, sep_c1;
  // forward pass 2 (boundary outputs), flat = zl index: owning node and the children's pass-through entries
  std::vector<int32_t> out_node, out_c0, out_c1;
  // backward: local index (into [interior | root boundary]) of every node boundary variable
  std::vector<int32_t> bnd_var;
  // level ranges (by height ascending) into the flat sep / out lists
  std::vector<int32_t> lvl_sep, lvl_out;
  // per cell of the box (row-major, dj * w + di): local index (into [interior | root boundary]) of its four
  // edges x(di,dj), x(di+1,dj), y(di,dj), y(di,dj+1) -- the fused pressure update of the sweep kernel
  std::vector<int32_t> cell_edge;
};

struct SubtreeInstance {
  int32_t tmpl = 0, root = 0;  // template, root node id in NdTree::nodes
  int32_t I0 = 0, J0 = 0;
  int64_t rf_base = 0, rb_base = 0;
};

struct NdTree {
  int32_t nx = 0, ny = 0;
  int64_t n_flux = 0;
  std::vector<NdNode> nodes;        // post-order: every subtree is a contiguous index range
  std::vector<int32_t> idx;         // canonical flux index of every front variable
  std::vector<int32_t> src;         // gather maps from the children's boundary entries
  std::vector<std::vector<int32_t>> top_by_height;  // top nodes only
  std::vector<std::vector<int32_t>> by_height;      // all nodes (factorisation order)
  std::vector<int32_t> node_of;     // canonical flux index -> node that eliminates it
  std::vector<int32_t> pos_of;      // canonical flux index -> position in that node's separator list
  std::vector<SubtreeTemplate> templates;
  std::vector<SubtreeInstance> instances;
  int64_t r_total = 0, z_total = 0, f_total = 0;
  double factor_flops = 0.0;

  // leaf_cells: boxes with at most this many cells are eliminated as one dense block.
  // subtree_cells: largest box handled by one CTA (0 = no subtrees, everything is a top node).
  void build(int32_t nx, int32_t ny, int32_t leaf_cells = 4, int32_t subtree_cells = 256);
};

}  // namespace mmmfe
