#include "nd_tree.h"

#include <algorithm>
#include <functional>

namespace mmmfe {

void NdTree::build(int32_t nx_, int32_t ny_, int32_t leaf_cells) {
  nx = nx_;
  ny = ny_;
  const int64_t nxe = int64_t(nx + 1) * ny;
  n_flux = nxe + int64_t(nx) * (ny + 1);
  nodes.clear();
  idx.clear();
  src.clear();
  node_of.assign(n_flux, -1);
  pos_of.assign(n_flux, -1);
  auto xe = [&](int i, int j) { return int32_t(int64_t(j) * (nx + 1) + i); };
  auto ye = [&](int i, int j) { return int32_t(nxe + int64_t(j) * nx + i); };

  std::vector<int32_t> pos(n_flux, -1);  // scratch: position of a variable in the front being built

  std::function<int32_t(int, int, int, int)> rec = [&](int i0, int i1, int j0, int j1) -> int32_t {
    const int w = i1 - i0, h = j1 - j0;
    std::vector<int32_t> sep;
    int32_t c0 = -1, c1 = -1;
    if (int64_t(w) * h <= leaf_cells || (w == 1 && h == 1)) {
      for (int j = j0; j < j1; ++j)
        for (int i = i0; i <= i1; ++i) {
          const bool inner = i > i0 && i < i1;
          if (inner || (i == i0 && i0 == 0) || (i == i1 && i1 == nx)) sep.push_back(xe(i, j));
        }
      for (int j = j0; j <= j1; ++j) {
        const bool inner = j > j0 && j < j1;
        if (inner || (j == j0 && j0 == 0) || (j == j1 && j1 == ny))
          for (int i = i0; i < i1; ++i) sep.push_back(ye(i, j));
      }
      if (sep.empty()) return -1;
    } else if (w >= h) {
      const int is = i0 + w / 2;
      c0 = rec(i0, is, j0, j1);
      c1 = rec(is, i1, j0, j1);
      for (int j = j0; j < j1; ++j) sep.push_back(xe(is, j));
    } else {
      const int js = j0 + h / 2;
      c0 = rec(i0, i1, j0, js);
      c1 = rec(i0, i1, js, j1);
      for (int i = i0; i < i1; ++i) sep.push_back(ye(i, js));
    }
    std::vector<int32_t> bnd;
    if (i0 > 0) for (int j = j0; j < j1; ++j) bnd.push_back(xe(i0, j));
    if (i1 < nx) for (int j = j0; j < j1; ++j) bnd.push_back(xe(i1, j));
    if (j0 > 0) for (int i = i0; i < i1; ++i) bnd.push_back(ye(i, j0));
    if (j1 < ny) for (int i = i0; i < i1; ++i) bnd.push_back(ye(i, j1));
    std::sort(sep.begin(), sep.end());
    std::sort(bnd.begin(), bnd.end());

    NdNode nd;
    nd.s = int32_t(sep.size());
    nd.b = int32_t(bnd.size());
    nd.child[0] = c0;
    nd.child[1] = c1;
    nd.idx_off = int64_t(idx.size());
    const int32_t id = int32_t(nodes.size());
    for (int32_t p = 0; p < nd.s; ++p) {
      idx.push_back(sep[p]);
      node_of[sep[p]] = id;
      pos_of[sep[p]] = p;
      pos[sep[p]] = p;
    }
    for (int32_t q = 0; q < nd.b; ++q) {
      idx.push_back(bnd[q]);
      pos[bnd[q]] = nd.s + q;
    }
    const int32_t f = nd.s + nd.b;
    nd.src_off = int64_t(src.size());
    src.resize(src.size() + 2 * size_t(f), -1);
    for (int c = 0; c < 2; ++c) {
      const int32_t ch = nd.child[c];
      if (ch < 0) continue;
      nodes[ch].parent = id;
      nd.height = std::max(nd.height, nodes[ch].height + 1);
      const NdNode &cn = nodes[ch];
      for (int32_t q = 0; q < cn.b; ++q) {
        const int32_t v = idx[cn.idx_off + cn.s + q];
        src[nd.src_off + int64_t(c) * f + pos[v]] = q;  // pos[v] is valid: child boundary is in this front
      }
    }
    for (int32_t p = 0; p < f; ++p) pos[idx[nd.idx_off + p]] = -1;
    nodes.push_back(nd);
    return id;
  };
  rec(0, nx, 0, ny);

  r_total = z_total = f_total = 0;
  factor_flops = 0.0;
  int32_t hmax = 0;
  for (auto &nd : nodes) {
    const int64_t f = nd.s + nd.b;
    nd.ldr = int32_t((f + 1) & ~int64_t(1));
    nd.nsplit = std::max(1, (nd.s + 127) / 128);
    nd.r_off = r_total;
    r_total += int64_t(nd.s) * nd.ldr;
    nd.z_off = z_total;
    z_total += nd.s + int64_t(nd.nsplit) * nd.b;
    nd.f_off = f_total;
    f_total += f * f;
    hmax = std::max(hmax, nd.height);
    const double s = nd.s, b = nd.b;
    factor_flops += 2.0 * s * s * s + 2.0 * b * s * s + 2.0 * b * b * s;
  }
  by_height.assign(size_t(hmax) + 1, {});
  for (int32_t i = 0; i < int32_t(nodes.size()); ++i) by_height[nodes[i].height].push_back(i);
}

}  // namespace mmmfe

// ---- diagnostics of include/mmmfe_b200.h ---------------------------------------------------------------------------
extern "C" int mmmfe_nd_tree_sizes(int32_t nx, int32_t ny, int32_t leaf_cells, int64_t sizes[6]) {
  if (nx < 1 || ny < 1 || !sizes) return -1;
  mmmfe::NdTree t;
  t.build(nx, ny, leaf_cells);
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
  t.build(nx, ny, leaf_cells);
  for (size_t i = 0; i < t.nodes.size(); ++i) {
    const mmmfe::NdNode &n = t.nodes[i];
    int32_t *row = node_table + 8 * i;
    row[0] = n.s; row[1] = n.b; row[2] = n.child[0]; row[3] = n.child[1];
    row[4] = n.height; row[5] = n.nsplit; row[6] = n.ldr; row[7] = n.parent;
    int64_t *o = offsets + 4 * i;
    o[0] = n.idx_off; o[1] = n.src_off; o[2] = n.r_off; o[3] = n.z_off;
  }
  std::copy(t.idx.begin(), t.idx.end(), idx);
  std::copy(t.src.begin(), t.src.end(), src);
  return 0;
}
