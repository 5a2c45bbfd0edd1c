#include "nd_tree.h"

#include <algorithm>
#include <functional>
#include <map>
#include <tuple>

namespace mmmfe {

void NdTree::build(int32_t nx_, int32_t ny_, int32_t leaf_cells, int32_t subtree_cells) {
  nx = nx_;
  ny = ny_;
  const int64_t nxe = int64_t(nx + 1) * ny;
  n_flux = nxe + int64_t(nx) * (ny + 1);
  nodes.clear();
  idx.clear();
  src.clear();
  templates.clear();
  instances.clear();
  node_of.assign(n_flux, -1);
  pos_of.assign(n_flux, -1);
  auto xe = [&](int i, int j) { return int32_t(int64_t(j) * (nx + 1) + i); };
  auto ye = [&](int i, int j) { return int32_t(nxe + int64_t(j) * nx + i); };

  std::vector<int32_t> first_desc;       // first node id of every node's subtree (post-order)

  // Pass 1 (serial, sizes only): the recursion fixes the post-order numbering, boxes, children, heights and the sizes
  // s / b of every front, so that all offsets are known.  Pass 2 fills the index lists of all fronts in parallel:
  // the separator and boundary lists of a front are a function of its box alone (ascending canonical index), and the
  // gather maps from a child's boundary entries are found by binary search in those sorted lists -- no state is
  // shared between fronts.  (The serial one-pass version took 0.45 s for a 1024^2 grid.)
  nodes.reserve(size_t(4) * size_t(nx) * ny / size_t(std::max(1, leaf_cells)) + 64);
  first_desc.reserve(nodes.capacity());
  std::function<int32_t(int, int, int, int)> rec = [&](int i0, int i1, int j0, int j1) -> int32_t {
    const int w = i1 - i0, h = j1 - j0;
    int32_t c0 = -1, c1 = -1, s = 0;
    const int32_t first = int32_t(nodes.size());
    if (int64_t(w) * h <= leaf_cells || (w == 1 && h == 1)) {
      s = h * ((w - 1) + (i0 == 0) + (i1 == nx)) + w * ((h - 1) + (j0 == 0) + (j1 == ny));
      if (s == 0) return -1;
    } else if (w >= h) {
      const int is = i0 + w / 2;
      c0 = rec(i0, is, j0, j1);
      c1 = rec(is, i1, j0, j1);
      s = h;
    } else {
      const int js = j0 + h / 2;
      c0 = rec(i0, i1, j0, js);
      c1 = rec(i0, i1, js, j1);
      s = w;
    }
    NdNode nd;
    nd.s = s;
    nd.b = h * ((i0 > 0) + (i1 < nx)) + w * ((j0 > 0) + (j1 < ny));
    nd.child[0] = c0;
    nd.child[1] = c1;
    nd.i0 = i0; nd.i1 = i1; nd.j0 = j0; nd.j1 = j1;
    const int32_t id = int32_t(nodes.size());
    for (int c = 0; c < 2; ++c)
      if (nd.child[c] >= 0) {
        nodes[nd.child[c]].parent = id;
        nd.height = std::max(nd.height, nodes[nd.child[c]].height + 1);
      }
    nodes.push_back(nd);
    first_desc.push_back(first);
    return id;
  };
  rec(0, nx, 0, ny);
  {
    int64_t io = 0, so = 0;
    for (NdNode &nd : nodes) {
      nd.idx_off = io; io += nd.s + nd.b;
      nd.src_off = so; so += 2 * int64_t(nd.s + nd.b);
    }
    idx.assign(size_t(io), -1);
    src.assign(size_t(so), -1);
  }
  // separator / leaf-interior list of a box, ascending (x-edges before y-edges, each row-major = ascending index)
  auto fill_sep = [&](const NdNode &nd, int32_t *out) {
    const int i0 = nd.i0, i1 = nd.i1, j0 = nd.j0, j1 = nd.j1, w = i1 - i0, h = j1 - j0;
    int32_t n = 0;
    if (nd.child[0] < 0 && nd.child[1] < 0) {
      for (int j = j0; j < j1; ++j)
        for (int i = i0; i <= i1; ++i) {
          const bool inner = i > i0 && i < i1;
          if (inner || (i == i0 && i0 == 0) || (i == i1 && i1 == nx)) out[n++] = xe(i, j);
        }
      for (int j = j0; j <= j1; ++j) {
        const bool inner = j > j0 && j < j1;
        if (inner || (j == j0 && j0 == 0) || (j == j1 && j1 == ny))
          for (int i = i0; i < i1; ++i) out[n++] = ye(i, j);
      }
    } else if (w >= h) {
      const int is = i0 + w / 2;
      for (int j = j0; j < j1; ++j) out[n++] = xe(is, j);
    } else {
      const int js = j0 + h / 2;
      for (int i = i0; i < i1; ++i) out[n++] = ye(i, js);
    }
    return n;
  };
  // boundary list of a box, ascending: x-edges of the two vertical sides interleave row by row, then the y-edges of
  // the bottom row, then those of the top row
  auto fill_bnd = [&](const NdNode &nd, int32_t *out) {
    const int i0 = nd.i0, i1 = nd.i1, j0 = nd.j0, j1 = nd.j1;
    int32_t n = 0;
    for (int j = j0; j < j1; ++j) {
      if (i0 > 0) out[n++] = xe(i0, j);
      if (i1 < nx) out[n++] = xe(i1, j);
    }
    if (j0 > 0) for (int i = i0; i < i1; ++i) out[n++] = ye(i, j0);
    if (j1 < ny) for (int i = i0; i < i1; ++i) out[n++] = ye(i, j1);
    return n;
  };
  bool bad = false;
#pragma omp parallel
  {
    std::vector<int32_t> cb;  // a child's boundary list
#pragma omp for schedule(dynamic, 256)
    for (int32_t id = 0; id < int32_t(nodes.size()); ++id) {
      const NdNode &nd = nodes[id];
      int32_t *sep = idx.data() + nd.idx_off, *bnd = sep + nd.s;
      if (fill_sep(nd, sep) != nd.s || fill_bnd(nd, bnd) != nd.b) { bad = true; continue; }
      for (int32_t p = 0; p < nd.s; ++p) { node_of[sep[p]] = id; pos_of[sep[p]] = p; }
      const int32_t f = nd.s + nd.b;
      for (int c = 0; c < 2; ++c) {
        const int32_t ch = nd.child[c];
        if (ch < 0) continue;
        const NdNode &cn = nodes[ch];
        cb.resize(size_t(cn.b));
        fill_bnd(cn, cb.data());
        for (int32_t q = 0; q < cn.b; ++q) {  // the child's boundary is part of this front
          const int32_t v = cb[q];
          const int32_t *ps = std::lower_bound(sep, sep + nd.s, v);
          int32_t pos;
          if (ps != sep + nd.s && *ps == v) pos = int32_t(ps - sep);
          else {
            const int32_t *pb = std::lower_bound(bnd, bnd + nd.b, v);
            if (pb == bnd + nd.b || *pb != v) { bad = true; continue; }
            pos = nd.s + int32_t(pb - bnd);
          }
          src[nd.src_off + int64_t(c) * f + pos] = q;
        }
      }
    }
  }
  if (bad) { nodes.clear(); return; }  // (cannot happen: the lists are a function of the boxes)
  const int32_t nn = int32_t(nodes.size());

  // ---- cut the tree: maximal boxes of at most subtree_cells cells become subtrees --------------------------------
  auto fits = [&](const NdNode &n) {
    return subtree_cells > 0 && int64_t(n.i1 - n.i0) * (n.j1 - n.j0) <= subtree_cells;
  };
  std::map<std::tuple<int, int, int>, int32_t> tmpl_of;
  for (int32_t id = 0; id < nn; ++id) {
    NdNode &root = nodes[id];
    if (!fits(root) || (root.parent >= 0 && fits(nodes[root.parent]))) continue;
    const int w = root.i1 - root.i0, h = root.j1 - root.j0;
    const int flags = (root.i0 == 0) | ((root.i1 == nx) << 1) | ((root.j0 == 0) << 2) | ((root.j1 == ny) << 3);
    // nodes of the subtree in processing order (height ascending, then id)
    std::vector<int32_t> order;
    for (int32_t k = first_desc[id]; k <= id; ++k) order.push_back(k);
    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return nodes[a].height < nodes[b].height; });
    const int32_t inst_id = int32_t(instances.size());
    for (size_t t = 0; t < order.size(); ++t) {
      nodes[order[t]].subtree = inst_id;
      nodes[order[t]].tnode = int32_t(t);
    }
    const auto key = std::make_tuple(w, h, flags);
    auto it = tmpl_of.find(key);
    if (it == tmpl_of.end()) {
      SubtreeTemplate T;
      T.w = w; T.h = h; T.flags = flags;
      T.n_nodes = int32_t(order.size());
      int32_t sep_run = 0, zl_run = 0, bnd_run = 0, rf_run = 0, rb_run = 0;
      for (int32_t k : order) {
        const NdNode &n = nodes[k];
        T.node_s.push_back(n.s); T.node_b.push_back(n.b);
        T.node_sep.push_back(sep_run); T.node_zl.push_back(zl_run); T.node_bnd.push_back(bnd_run);
        T.node_rf.push_back(rf_run); T.node_rb.push_back(rb_run);
        sep_run += n.s; zl_run += n.b; bnd_run += n.b;
        rf_run += n.s * n.b; rb_run += (n.s + n.b) * n.s;
      }
      T.n_int = sep_run; T.zl_size = zl_run; T.n_rootb = root.b;
      T.rf_size = (rf_run + 1) & ~1; T.rb_size = (rb_run + 1) & ~1;
      auto code_of = [&](int32_t v) {
        int type, i, j;
        if (v < nxe) { type = 0; j = v / (nx + 1); i = v % (nx + 1); }
        else { type = 1; const int64_t k = v - nxe; j = int(k / nx); i = int(k % nx); }
        return int32_t((type << 30) | ((j - root.j0) << 15) | (i - root.i0));
      };
      std::map<int32_t, int32_t> rootb_pos;
      for (int32_t q = 0; q < root.b; ++q) {
        const int32_t v = idx[root.idx_off + root.s + q];
        rootb_pos[v] = q;
        T.rootb_code.push_back(code_of(v));
      }
      int32_t cur_h = -1;
      for (size_t t = 0; t < order.size(); ++t) {
        const NdNode &n = nodes[order[t]];
        const int32_t f = n.s + n.b;
        if (n.height != cur_h) {
          cur_h = n.height;
          T.lvl_sep.push_back(T.node_sep[t]);
          T.lvl_out.push_back(T.node_zl[t]);
        }
        auto child_entry = [&](int c, int32_t p) -> int32_t {
          const int32_t ch = n.child[c];
          if (ch < 0) return -1;
          const int32_t q = src[n.src_off + int64_t(c) * f + p];
          return q < 0 ? -1 : T.node_zl[nodes[ch].tnode] + q;
        };
        for (int32_t p = 0; p < n.s; ++p) {
          T.var_code.push_back(code_of(idx[n.idx_off + p]));
          T.sep_node.push_back(int32_t(t));
          T.sep_c0.push_back(child_entry(0, p));
          T.sep_c1.push_back(child_entry(1, p));
        }
        for (int32_t q = 0; q < n.b; ++q) {
          T.out_node.push_back(int32_t(t));
          T.out_c0.push_back(child_entry(0, n.s + q));
          T.out_c1.push_back(child_entry(1, n.s + q));
          const int32_t v = idx[n.idx_off + n.s + q];
          const int32_t owner = node_of[v];
          if (owner >= first_desc[id] && owner <= id) T.bnd_var.push_back(T.node_sep[nodes[owner].tnode] + pos_of[v]);
          else T.bnd_var.push_back(T.n_int + rootb_pos.at(v));
        }
      }
      {
        std::map<int32_t, int32_t> local_of;  // canonical flux index -> local variable
        int32_t l = 0;
        for (int32_t k : order)
          for (int32_t p = 0; p < nodes[k].s; ++p) local_of[idx[nodes[k].idx_off + p]] = l++;
        for (const auto &kv : rootb_pos) local_of[kv.first] = T.n_int + kv.second;
        for (int j = root.j0; j < root.j1; ++j)
          for (int i = root.i0; i < root.i1; ++i)
            for (int32_t v : {xe(i, j), xe(i + 1, j), ye(i, j), ye(i, j + 1)}) T.cell_edge.push_back(local_of.at(v));
      }
      T.n_levels = int32_t(T.lvl_sep.size());
      T.lvl_sep.push_back(T.n_int);
      T.lvl_out.push_back(T.zl_size);
      it = tmpl_of.emplace(key, int32_t(templates.size())).first;
      templates.push_back(std::move(T));
    }
    SubtreeInstance inst;
    inst.tmpl = it->second;
    inst.root = id;
    inst.I0 = root.i0; inst.J0 = root.j0;
    instances.push_back(inst);
  }

  // ---- storage offsets ------------------------------------------------------------------------------------------
  r_total = z_total = f_total = 0;
  factor_flops = 0.0;
  int32_t hmax = 0;
  for (auto &nd : nodes) {  // top nodes: row-major [F11^-1 | -G^T] plus a contiguous forward copy of -G^T
    const int64_t f = nd.s + nd.b;
    nd.f_off = f_total;
    f_total += f * f;
    hmax = std::max(hmax, nd.height);
    const double s = nd.s, b = nd.b;
    factor_flops += 2.0 * s * s * s + 2.0 * b * s * s + 2.0 * b * b * s;
    if (nd.subtree >= 0) continue;
    nd.ldr = int32_t((f + 1) & ~int64_t(1));
    nd.ldf = int32_t((nd.b + 1) & ~1);
    nd.nsplit = std::max(1, (nd.s + 31) / 32);
    nd.rb_off = r_total;
    r_total += int64_t(nd.s) * nd.ldr;
    nd.rf_off = r_total;
    r_total += int64_t(nd.s) * nd.ldf;
    nd.z_off = z_total;
    z_total += nd.s + int64_t(nd.nsplit) * nd.b;
  }
  for (auto &inst : instances) {
    const SubtreeTemplate &T = templates[inst.tmpl];
    inst.rf_base = r_total;
    r_total += T.rf_size;
    inst.rb_base = r_total;
    r_total += T.rb_size;
    NdNode &root = nodes[inst.root];
    root.z_off = z_total;
    z_total += root.s + root.b;
  }
  for (auto &nd : nodes) {
    if (nd.subtree < 0) continue;
    const SubtreeInstance &inst = instances[nd.subtree];
    const SubtreeTemplate &T = templates[inst.tmpl];
    nd.nsplit = 1;
    nd.ldr = nd.s;  // R^T: (s + b) rows of s
    nd.ldf = nd.b;
    nd.rb_off = inst.rb_base + T.node_rb[nd.tnode];
    nd.rf_off = inst.rf_base + T.node_rf[nd.tnode];
    if (nd.tnode != T.n_nodes - 1) nd.z_off = -1;
  }
  by_height.assign(size_t(hmax) + 1, {});
  top_by_height.assign(size_t(hmax) + 1, {});
  for (int32_t i = 0; i < nn; ++i) {
    by_height[nodes[i].height].push_back(i);
    if (nodes[i].subtree < 0) top_by_height[nodes[i].height].push_back(i);
  }
}

}  // namespace mmmfe

// ---- diagnostics of include/mmmfe_b200.h ---------------------------------------------------------------------------
extern "C" int mmmfe_nd_tree_sizes(int32_t nx, int32_t ny, int32_t leaf_cells, int64_t sizes[6]) {
  if (nx < 1 || ny < 1 || !sizes) return -1;
  mmmfe::NdTree t;
  t.build(nx, ny, leaf_cells, 0);
  sizes[0] = int64_t(t.nodes.size());
  sizes[1] = int64_t(t.idx.size());
  sizes[2] = int64_t(t.src.size());
  sizes[3] = t.r_total;
  sizes[4] = t.z_total;
  sizes[5] = int64_t(t.by_height.size());
  return 0;
}

extern "C" int mmmfe_nd_tree_fill(int32_t nx, int32_t ny, int32_t leaf_cells, int32_t *node_table, int64_t *offsets,
                                  int32_t *idx, int32_t *src) {
  if (nx < 1 || ny < 1 || !node_table || !offsets || !idx || !src) return -1;
  mmmfe::NdTree t;
  t.build(nx, ny, leaf_cells, 0);
  for (size_t i = 0; i < t.nodes.size(); ++i) {
    const mmmfe::NdNode &n = t.nodes[i];
    int32_t *row = node_table + 8 * i;
    row[0] = n.s; row[1] = n.b; row[2] = n.child[0]; row[3] = n.child[1];
    row[4] = n.height; row[5] = n.nsplit; row[6] = n.ldr; row[7] = n.parent;
    int64_t *o = offsets + 4 * i;
    o[0] = n.idx_off; o[1] = n.src_off; o[2] = n.rb_off; o[3] = n.z_off;
  }
  std::copy(t.idx.begin(), t.idx.end(), idx);
  std::copy(t.src.begin(), t.src.end(), src);
  return 0;
}
