#include "sweep_plan.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>

namespace mmmfe {

namespace {

int32_t pow2_floor(int64_t v) {
  int32_t p = 1;
  while (int64_t(p) * 2 <= v) p *= 2;
  return p;
}
int32_t pow2_ceil(int64_t v) {
  int32_t p = 1;
  while (p < v) p *= 2;
  return p;
}
int32_t even(int64_t v) { return int32_t((v + 1) & ~int64_t(1)); }

char g_msg[256];

// Packed subtree template (shared-memory resident while its subtrees are processed).
// Header, int32 words: 0 n_levels, 1 n_int, 2 n_rootb, 3 zl_size, 4 w, 5 h, 6 rf_size, 7 rb_size (sweep layout),
// 8 uint16 words in total; offsets in uint16 units from the blob start: 9 var_code, 10 cell_edge, 11 lvl_out,
// 12 fdesc (8 uint16 per forward output: rf_off, s << 8 | b, sep0, writer flag, child-0 zl, child-1 zl, -, -),
// 13 sc (2 uint16 per separator variable: child-0 zl, child-1 zl), 14 lvlb (4 int32 per backward level: first
// variable, outputs, log2 lanes per output, element step), 15 bdesc (4 uint16 per variable: row offset, row
// length, input-list offset, -), 16 in_idx (input of every row element: separator rhs index, or 0x8000 | solution
// index).
// Backward blob: levels with one lane per output are stored interleaved (element p of output o at
// base + p * n_out + o: conflict-free), the others row-major; perm[k] = old position of element k, or -1 (zero).
const char *pack_template(const SubtreeTemplate &T, int32_t ct, std::vector<uint16_t> &out, std::vector<int32_t> &perm,
                          int32_t &rb2_size) {
  if (T.w > 127 || T.h > 127) return "subtree box wider than 127 cells";
  const size_t base = out.size();
  out.resize(base + 2 * SWEEP_TMPL_HDR, 0);
  std::vector<int32_t> hdr(SWEEP_TMPL_HDR, 0);
  bool overflow = false;
  auto put = [&](int slot, const std::vector<int32_t> &v) {
    while ((out.size() - base) % 8) out.push_back(0);  // every array starts on a 16-byte boundary
    hdr[slot] = int32_t(out.size() - base);
    for (int32_t y : v) {
      if (y < -1 || y >= 0xFFFF) overflow = true;
      out.push_back(y < 0 ? uint16_t(0xFFFF) : uint16_t(y));
    }
  };
  auto put32 = [&](int slot, const std::vector<int32_t> &v) {
    while ((out.size() - base) % 8) out.push_back(0);
    hdr[slot] = int32_t(out.size() - base);
    for (int32_t y : v) {
      out.push_back(uint16_t(uint32_t(y) & 0xFFFF));
      out.push_back(uint16_t(uint32_t(y) >> 16));
    }
  };
  std::vector<int32_t> code(T.var_code.size());
  for (size_t l = 0; l < code.size(); ++l) {
    const int32_t x = T.var_code[l];
    code[l] = ((x >> 30) << 14) | (((x >> 15) & 0x7fff) << 7) | (x & 0x7fff);
  }
  // forward
  std::vector<int32_t> fdesc(size_t(T.zl_size) * 8, 0), sc(size_t(T.n_int) * 2, -1);
  for (int32_t o = 0; o < T.zl_size; ++o) {
    const int32_t n = T.out_node[o];
    if (T.node_s[n] > 255 || T.node_b[n] > 255) overflow = true;
    fdesc[8 * o + 0] = T.node_rf[n] + (o - T.node_zl[n]);
    fdesc[8 * o + 1] = (T.node_s[n] << 8) | T.node_b[n];
    fdesc[8 * o + 2] = T.node_sep[n];
    fdesc[8 * o + 3] = (o == T.node_zl[n]) ? 1 : 0;  // the first output of a node stores its accumulated rhs
    fdesc[8 * o + 4] = T.out_c0[o];
    fdesc[8 * o + 5] = T.out_c1[o];
  }
  for (int32_t v = 0; v < T.n_int; ++v) { sc[2 * v] = T.sep_c0[v]; sc[2 * v + 1] = T.sep_c1[v]; }
  // backward
  std::vector<int32_t> lvlb(size_t(T.n_levels) * 4, 0), bdesc(size_t(T.n_int) * 4, 0), in_idx;
  perm.clear();
  int32_t node = 0;
  for (int32_t lv = 0; lv < T.n_levels; ++lv) {
    const int32_t v0 = T.lvl_sep[lv], n_out = T.lvl_sep[lv + 1] - v0;
    const int32_t tpo = std::max(1, std::min(32, pow2_floor(std::max(1, ct / std::max(1, n_out)))));
    int32_t tl = 0;
    while ((1 << tl) < tpo) ++tl;
    const int32_t node0 = node;
    int32_t max_len = 0;
    while (node < T.n_nodes && T.node_sep[node] < T.lvl_sep[lv + 1]) { max_len = std::max(max_len, T.node_s[node] + T.node_b[node]); ++node; }
    const bool interleaved = tpo == 1;
    const int32_t estep = interleaved ? n_out : 1;
    lvlb[4 * lv + 0] = v0; lvlb[4 * lv + 1] = n_out; lvlb[4 * lv + 2] = tl; lvlb[4 * lv + 3] = estep;
    const int32_t blob0 = int32_t(perm.size()), idx0 = int32_t(in_idx.size());
    if (interleaved) {
      perm.resize(size_t(blob0) + size_t(max_len) * n_out, -1);
      in_idx.resize(size_t(idx0) + size_t(max_len) * n_out, 0);
    }
    for (int32_t nd = node0; nd < node; ++nd) {
      const int32_t s = T.node_s[nd], b = T.node_b[nd], f = s + b, sep0 = T.node_sep[nd];
      for (int32_t i = 0; i < s; ++i) {
        const int32_t v = sep0 + i, o = v - v0;
        int32_t row_off, in_off;
        if (interleaved) { row_off = blob0 + o; in_off = idx0 + o; }
        else { row_off = int32_t(perm.size()); in_off = int32_t(in_idx.size()); perm.resize(perm.size() + f, -1); in_idx.resize(in_idx.size() + f, 0); }
        for (int32_t pp = 0; pp < f; ++pp) {
          perm[size_t(row_off) + size_t(pp) * estep] = T.node_rb[nd] + pp * s + i;  // old layout: R^T[p][i]
          in_idx[size_t(in_off) + size_t(pp) * estep] = pp < s ? sep0 + pp : (0x8000 | T.bnd_var[T.node_bnd[nd] + pp - s]);
        }
        bdesc[4 * v + 0] = row_off; bdesc[4 * v + 1] = interleaved ? max_len : f; bdesc[4 * v + 2] = in_off;
      }
    }
  }
  if (perm.size() % 2) perm.push_back(-1);
  rb2_size = int32_t(perm.size());
  if (T.n_int + T.n_rootb >= 0x8000) overflow = true;
  hdr[0] = T.n_levels; hdr[1] = T.n_int; hdr[2] = T.n_rootb; hdr[3] = T.zl_size; hdr[4] = T.w; hdr[5] = T.h;
  hdr[6] = T.rf_size; hdr[7] = rb2_size;
  put(9, code); put(10, T.cell_edge); put(11, T.lvl_out); put(12, fdesc); put(13, sc); put32(14, lvlb);
  put(15, bdesc); put(16, in_idx);
  while ((out.size() - base) % 8) out.push_back(0);  // 16-byte granules
  hdr[8] = int32_t(out.size() - base);
  std::memcpy(out.data() + base, hdr.data(), SWEEP_TMPL_HDR * sizeof(int32_t));
  return overflow ? "subtree template field beyond 16 bits" : nullptr;
}

}  // namespace

const char *build_sweep_layout(const NdTree &tr, const std::vector<int32_t> &idx_user,
                               const std::vector<int32_t> &dof_of_canon, int32_t n_cta, int32_t chunk_dbl, int32_t ct,
                               SweepLayout &L, int32_t ng, int32_t mt_dbl, int32_t pack, const std::vector<int32_t> &inst_class) {
  L = SweepLayout();
  L.n_cta = n_cta;
  L.chunk_dbl = chunk_dbl;
  L.ct = ct;
  L.ng = std::max(1, ng);
  L.mt_dbl = std::max(64, mt_dbl);
  L.pack = (pack == 2 || pack == 4 || pack == 8) && ct / pack >= 32 ? pack : 1;
  if (const char *env = std::getenv("MMMFE_SWEEP_WS_ROWS")) L.ws_rows = std::atoi(env);
  if (const char *env = std::getenv("MMMFE_SWEEP_DIRECT")) L.direct = std::atoi(env) != 0;
  const int32_t nn = int32_t(tr.nodes.size());
  if (tr.instances.empty()) return "no subtrees (subtree_cells = 0)";
  if (tr.idx.size() >= (size_t(1) << 30) || tr.src.size() >= (size_t(1) << 31)) return "index structure beyond 32 bits";
  const int64_t nxe = int64_t(tr.nx + 1) * tr.ny;
  auto user = [&](int64_t canon) { return dof_of_canon.empty() ? int32_t(canon) : dof_of_canon[canon]; };

  // ---- templates, instances, box-major cell order ----
  std::vector<int32_t> perm_off;
  for (const auto &T : tr.templates) {
    L.tmpl_off.push_back(int64_t(L.tmpl16.size()));
    const size_t before = L.tmpl16.size();
    std::vector<int32_t> perm;
    int32_t rb2 = 0;
    if (const char *err = pack_template(T, ct / L.pack, L.tmpl16, perm, rb2)) return err;
    L.tmpl16_max = std::max(L.tmpl16_max, int32_t(L.tmpl16.size() - before));
    perm_off.push_back(int32_t(L.perm.size()));
    L.perm.insert(L.perm.end(), perm.begin(), perm.end());
    L.rb2_size.push_back(rb2);
    L.sub_ws_max = std::max(L.sub_ws_max, std::max(T.n_int + T.zl_size, T.n_int + T.n_rootb));
  }
  L.perm.push_back(0);
  // ---- subtrees -> CTAs, in units of PACKS: a pack costs the same whether it has one member or L.pack (lockstep), so
  // the subtrees are first grouped by (template, class) -- what may share a pack -- cut into packs in instance order
  // (neighbouring boxes stay together), and the packs are dealt out in contiguous runs: every CTA gets
  // floor or ceil(n_packs / n_cta) of them.  (Dealing out subtrees instead left CTAs with up to 17 mostly
  // half-empty packs against a mean of 14.9 on a 1024^2 grid.) ----
  const int32_t n_inst = int32_t(tr.instances.size());
  L.cta_of_inst.assign(n_inst, 0);
  {
    std::map<std::pair<int32_t, int32_t>, std::vector<int32_t>> groups;  // (interior first, template, class) -> instances
    for (int32_t k = 0; k < n_inst; ++k) {
      const int32_t t = tr.instances[k].tmpl;
      const int32_t key0 = (tr.templates[t].flags == 0 ? 0 : 1 << 20) + t;
      groups[{key0, inst_class.empty() ? -1 : inst_class[k]}].push_back(k);
    }
    std::vector<std::pair<int32_t, int32_t>> pack_of;  // (first index into order, members)
    std::vector<int32_t> order;
    for (const auto &g : groups)
      for (size_t i = 0; i < g.second.size(); i += size_t(L.pack)) {
        const size_t n = std::min(size_t(L.pack), g.second.size() - i);
        pack_of.push_back({int32_t(order.size()), int32_t(n)});
        order.insert(order.end(), g.second.begin() + i, g.second.begin() + i + n);
      }
    // boundary packs (they also read interface data and write traces: ~30 % slower) are spread evenly over the CTAs
    // first; the interior packs then fill every CTA up to its quota, in contiguous runs
    const int64_t n_packs = int64_t(pack_of.size());
    std::vector<int64_t> bnd, inr;
    for (int64_t j = 0; j < n_packs; ++j)
      (tr.templates[tr.instances[order[pack_of[j].first]].tmpl].flags == 0 ? inr : bnd).push_back(j);
    std::vector<int32_t> load(n_cta, 0);
    auto place = [&](int64_t j, int32_t cta) {
      for (int32_t m = 0; m < pack_of[j].second; ++m) L.cta_of_inst[order[pack_of[j].first + m]] = cta;
      load[cta] += 1;
    };
    const int64_t nb = int64_t(bnd.size());
    for (int64_t t = 0; t < nb; ++t) place(bnd[t], int32_t(std::min<int64_t>(n_cta - 1, ((2 * t + 1) * n_cta) / (2 * nb))));
    size_t next = 0;
    for (int32_t cta = 0; cta < n_cta && next < inr.size(); ++cta) {
      const int64_t quota = (int64_t(cta) + 1) * n_packs / n_cta - int64_t(cta) * n_packs / n_cta;
      while (load[cta] < quota && next < inr.size()) place(inr[next++], cta);
    }
    for (int32_t cta = 0; next < inr.size(); cta = (cta + 1) % n_cta) place(inr[next++], cta);  // left-overs
  }
  // allocation order (CTA, template, instance): the members of a pack are neighbours in every array
  L.inst_at_pos.resize(n_inst);
  for (int32_t k = 0; k < n_inst; ++k) L.inst_at_pos[k] = k;
  auto class_of = [&](int32_t k) { return inst_class.empty() ? -1 : inst_class[k]; };
  std::stable_sort(L.inst_at_pos.begin(), L.inst_at_pos.end(), [&](int32_t x, int32_t y) {
    if (L.cta_of_inst[x] != L.cta_of_inst[y]) return L.cta_of_inst[x] < L.cta_of_inst[y];
    if (tr.instances[x].tmpl != tr.instances[y].tmpl) return tr.instances[x].tmpl < tr.instances[y].tmpl;
    return class_of(x) < class_of(y);
  });
  L.cell_box.assign(size_t(tr.nx) * tr.ny, -1);
  L.inst.resize(tr.instances.size());
  for (int32_t pos = 0; pos < n_inst; ++pos) {
    const int32_t k = L.inst_at_pos[pos];
    const SubtreeInstance &in = tr.instances[k];
    const SubtreeTemplate &T = tr.templates[in.tmpl];
    SweepInstH &h = L.inst[k];
    h.pos = pos;
    L.inst_ij.push_back(in.I0); L.inst_ij.push_back(in.J0);
    h.xi_off = int32_t(L.xi_total);
    L.xi_total += even(T.n_int);
    h.pb_off = int32_t(L.pb_total);
    L.pb_total += even(int64_t(T.w) * T.h);
    for (int dj = 0; dj < T.h; ++dj)
      for (int di = 0; di < T.w; ++di) L.cell_box[size_t(in.J0 + dj) * tr.nx + in.I0 + di] = h.pb_off + dj * T.w + di;
    h.xb = int64_t(L.index.size());
    for (int32_t q = 0; q < T.n_rootb; ++q) {
      const int32_t c = T.rootb_code[q];
      const int type = c >> 30, dj = (c >> 15) & 0x7fff, di = c & 0x7fff;
      const int64_t canon = type == 0 ? int64_t(in.J0 + dj) * (tr.nx + 1) + in.I0 + di
                                      : nxe + int64_t(in.J0 + dj) * tr.nx + in.I0 + di;
      L.index.push_back(user(canon));
    }
    while (L.index.size() % 4) L.index.push_back(0);
    // a pack never mixes classes: either all members share one blob, or (no class information) none does
    if (L.packs.empty() || L.packs.back().n >= L.pack || L.packs.back().tmpl != in.tmpl ||
        L.packs.back().cta != L.cta_of_inst[k] || L.packs.back().cls != class_of(k)) {
      SweepPackH P;
      P.pos0 = pos; P.tmpl = in.tmpl; P.cta = L.cta_of_inst[k]; P.cls = class_of(k);
      L.packs.push_back(P);
    }
    L.packs.back().n += 1;
    L.packs.back().shared = (L.packs.back().cls >= 0 && L.packs.back().n > 1) ? 1 : 0;
  }
  for (int32_t cb : L.cell_box)
    if (cb < 0) return "a cell lies outside every subtree";
  if (L.xi_total >= (int64_t(1) << 29) || L.pb_total >= (int64_t(1) << 29)) return "vector storage beyond 32 bits";

  // ---- sweep nodes ----
  std::vector<int32_t> sid(nn, -1);
  for (int32_t i = 0; i < nn; ++i) {
    const NdNode &n = tr.nodes[i];
    if (n.subtree >= 0) continue;
    if (n.child[0] < 0 && n.child[1] < 0) return "a leaf front lies outside every subtree";
    sid[i] = int32_t(L.nodes.size());
    SweepNodeH h;
    h.old_id = i; h.s = n.s; h.b = n.b; h.height = n.height;
    h.sep_type = tr.idx[n.idx_off] < nxe ? 0 : 1;
    L.nodes.push_back(h);
  }
  L.n_top = int32_t(L.nodes.size());
  for (size_t k = 0; k < tr.instances.size(); ++k) {
    const SubtreeInstance &in = tr.instances[k];
    sid[in.root] = int32_t(L.nodes.size());
    SweepNodeH h;
    h.old_id = in.root; h.s = 0; h.b = tr.templates[in.tmpl].n_rootb; h.height = tr.nodes[in.root].height;
    h.n_fwd = 1; h.n_bwd = 1;
    L.nodes.push_back(h);
  }
  for (auto &h : L.nodes) {
    const NdNode &n = tr.nodes[h.old_id];
    if (h.s > 0)
      for (int c = 0; c < 2; ++c)
        if (n.child[c] >= 0) {
          if (sid[n.child[c]] < 0) return "child of a top front is neither a top front nor a subtree root";
          h.child[c] = sid[n.child[c]];
        }
    if (n.parent >= 0) h.parent = sid[n.parent];
    if (h.s > 0) {
      h.zs_off = int32_t(L.z_total); L.z_total += h.s;
      h.zc_off = int32_t(L.z_total); L.z_total += h.b;
    }
  }
  for (int32_t pos = 0; pos < n_inst; ++pos) {  // root-boundary contributions of the subtrees: allocation order
    SweepNodeH &h = L.nodes[L.n_top + L.inst_at_pos[pos]];
    h.zc_off = int32_t(L.z_total);
    L.z_total += h.b;
  }
  if (L.z_total >= (int64_t(1) << 29)) return "vector storage beyond 32 bits";

  // producers of every completion counter (2 * node: forward, 2 * node + 1: backward): the CTA when it is a single
  // one.  A consumer that runs later on the same CTA needs no counter (one compute group: program order + barrier).
  std::vector<int32_t> prod_cta(2 * L.nodes.size(), -1);  // -1 none yet, -2 several CTAs
  auto add_producer = [&](int32_t counter, int32_t cta) {
    int32_t &p = prod_cta[counter];
    p = (p == -1 || p == cta) ? cta : -2;
  };
  for (int32_t k = 0; k < n_inst; ++k) add_producer(2 * (L.n_top + k), L.cta_of_inst[k]);

  // ---- waves: per level and direction, the level's tile doubles are cut into n_cta contiguous shares (whole small
  // fronts, column ranges of large ones); a CTA's share is packed into waves of at most chunk_dbl tile doubles ----
  const size_t nh = tr.top_by_height.size();
  L.fwd.assign(nh, {});
  L.bwd.assign(nh, {});
  int64_t rs = 0;
  struct Piece { int32_t node, c0, nc; };
  auto plan_level = [&](size_t h, int pass) -> const char * {
    auto &waves = pass == 0 ? L.fwd[h] : L.bwd[h];
    const auto &ids = tr.top_by_height[h];
    double level_dbl = 0;
    for (int32_t id : ids) {
      const NdNode &n = tr.nodes[id];
      level_dbl += pass == 0 ? double(n.s) * n.b : double(n.s) * (n.s + n.b);
    }
    const double share = std::max(1.0, level_dbl / n_cta);
    std::vector<std::vector<Piece>> pieces(n_cta);
    double cum = 0;
    for (int32_t id : ids) {
      const NdNode &n = tr.nodes[id];
      const int32_t rlen = pass == 0 ? n.s : n.s + n.b, ncols = pass == 0 ? n.b : n.s;
      const double dbl = double(rlen) * ncols;
      auto cta_at = [&](double pos) { return std::max(0, std::min(n_cta - 1, int32_t(pos / share))); };
      if (ncols == 0) {  // root, forward: rows only (accumulates and stores the separator right-hand side)
        pieces[cta_at(cum)].push_back({sid[id], 0, 0});
      } else if (dbl <= 0.75 * share || ncols <= 4) {
        pieces[cta_at(cum + 0.5 * dbl)].push_back({sid[id], 0, ncols});
      } else {
        int32_t c = 0, k = cta_at(cum + 0.5 * rlen);
        while (c < ncols) {
          const double pos = cum + double(c) * rlen;
          int32_t take = int32_t(((k + 1) * share - pos) / rlen + 0.5);
          take = std::max(take, 1);
          if (k == n_cta - 1 || ncols - (c + take) < 2) take = ncols - c;
          take = std::min(take, ncols - c);
          pieces[k].push_back({sid[id], c, take});
          c += take;
          k = std::min(k + 1, n_cta - 1);
        }
      }
      cum += dbl;
    }
    for (int32_t cta = 0; cta < n_cta; ++cta) {
      if (pieces[cta].empty()) continue;
      // strips of every piece: widths 64, then the binary digits of the rest; never more than chunk_dbl doubles
      struct Strip { int32_t node, c0, w, parts, direct; };
      std::vector<Strip> strips;
      for (const Piece &pc : pieces[cta]) {
        const SweepNodeH &sn = L.nodes[pc.node];
        const int32_t rlen = pass == 0 ? sn.s : sn.s + sn.b;
        if (pc.nc == 0) { strips.push_back({pc.node, 0, 0, 0, 0}); continue; }
        const int32_t wmax = std::max(1, std::min(64, pow2_floor(std::max<int64_t>(1, L.chunk_dbl / std::max(1, rlen)))));
        // DIRECT strips: as wide as ONE micro-tile (up to 1.5 mt_dbl doubles) can hold with all rows, so that the warp
        // that multiplies it owns the complete column sums and writes the outputs itself (no partial sums in shared
        // memory, no second barrier, no output phase).  Fronts with more rows than that keep the split tiles.
        const int64_t cap = (3 * int64_t(L.mt_dbl)) / 2;
        const int32_t wdir = (L.direct && rlen <= cap) ? std::min(wmax, pow2_floor(std::max<int64_t>(1, cap / rlen))) : 0;
        int32_t c = pc.c0, left = pc.nc;
        while (left > 0) {
          const int32_t w = std::min(wdir > 0 ? wdir : wmax, pow2_floor(left));
          const int32_t parts = wdir > 0 ? 1 : std::max(1, std::min(rlen, int32_t((int64_t(rlen) * w + L.mt_dbl / 2) / L.mt_dbl)));
          strips.push_back({pc.node, c, w, parts, (L.direct && parts == 1) ? 1 : 0});
          c += w; left -= w;
        }
      }
      // A front that several consecutive waves of this CTA work on is signalled ONCE, by the last of them: the warps
      // run a CTA's entries in order (one compute group), so that wave's completion implies the earlier ones', and
      // the consumers cannot start before it anyway.  Fewer release operations (each one drains the SM's stores).
      std::map<int32_t, size_t> last_strip;
      for (size_t q = 0; q < strips.size(); ++q) last_strip[strips[q].node] = q;
      size_t i = 0;
      std::vector<int32_t> prev_fronts;
      while (i < strips.size()) {
        SweepWaveH W;
        W.kind = pass == 0 ? SW_TOP_FWD : SW_TOP_BWD; W.cta = cta; W.height = int32_t(h);
        std::vector<int32_t> row0_of;      // per front of the wave: first row in the gathered vector
        std::vector<int32_t> rows, zsd, colA, colB;
        std::vector<SweepMt> mts;
        std::vector<RepackDesc> reps;
        int64_t tile = 0;
        bool all_direct = true;
        for (; i < strips.size(); ++i) {
          const Strip &sp = strips[i];
          const SweepNodeH &sn = L.nodes[sp.node];
          const NdNode &n = tr.nodes[sn.old_id];
          const int32_t rlen = pass == 0 ? sn.s : sn.s + sn.b, ncols = pass == 0 ? sn.b : sn.s, f = sn.s + sn.b;
          auto child_z = [&](int c, int32_t p) -> int32_t {
            if (sn.child[c] < 0) return -1;
            const int32_t q = tr.src[n.src_off + int64_t(c) * f + p];
            return q < 0 ? -1 : L.nodes[sn.child[c]].zc_off + q;
          };
          int64_t sdbl = 0;
          for (int32_t q = 0; q < sp.parts; ++q) {
            const int32_t r0 = int32_t(int64_t(rlen) * q / sp.parts), r1 = int32_t(int64_t(rlen) * (q + 1) / sp.parts);
            sdbl += even(int64_t(r1 - r0) * sp.w);
          }
          auto it = std::find(W.fronts.begin(), W.fronts.end(), sp.node);
          const bool new_front = it == W.fronts.end();
          // dependencies the new front would add
          std::vector<int32_t> nd;
          if (new_front) {
            auto want = [&](int32_t counter, int32_t count) {
              if (L.ng == 1 && prod_cta[counter] == cta) return;
              const int32_t packed = counter | (count << SWEEP_DEP_SHIFT);
              if (std::find(W.deps.begin(), W.deps.end(), packed) == W.deps.end() &&
                  std::find(nd.begin(), nd.end(), packed) == nd.end()) nd.push_back(packed);
            };
            if (pass == 0) {
              for (int c = 0; c < 2; ++c)
                if (sn.child[c] >= 0) want(2 * sn.child[c], L.nodes[sn.child[c]].n_fwd);
            } else if (sn.parent >= 0) want(2 * sn.parent + 1, L.nodes[sn.parent].n_bwd);
            else want(2 * sp.node, sn.n_fwd);
          }
          if (!W.fronts.empty() && (tile + sdbl > L.chunk_dbl || W.deps.size() + nd.size() > size_t(SWEEP_MAX_DEPS))) break;
          if (nd.size() > size_t(SWEEP_MAX_DEPS)) return "a front has more dependencies than an entry can hold";
          int32_t fi;
          if (new_front) {
            fi = int32_t(W.fronts.size());
            W.fronts.push_back(sp.node);
            W.deps.insert(W.deps.end(), nd.begin(), nd.end());
            row0_of.push_back(int32_t(pass == 0 ? rows.size() / 4 : rows.size()));
            if (pass == 0) {
              // the wave that holds column 0 of the front (or its only, column-less piece) stores the accumulated
              // separator right-hand side for the backward pass
              const bool writes_zs = sp.c0 == 0;
              for (int32_t p = 0; p < sn.s; ++p) {
                const int64_t v = tr.idx[n.idx_off + p];
                int32_t c0, c1;
                if (v < nxe) {
                  const int32_t j = int32_t(v / (tr.nx + 1)), ii = int32_t(v % (tr.nx + 1));
                  if (ii <= 0 || ii >= tr.nx) return "separator edge on the domain boundary";
                  c0 = j * tr.nx + ii - 1; c1 = j * tr.nx + ii;
                } else {
                  const int64_t k = v - nxe;
                  const int32_t j = int32_t(k / tr.nx), ii = int32_t(k % tr.nx);
                  if (j <= 0 || j >= tr.ny) return "separator edge on the domain boundary";
                  c0 = (j - 1) * tr.nx + ii; c1 = j * tr.nx + ii;
                }
                rows.push_back(L.cell_box[c0] | (sn.sep_type << 30)); rows.push_back(L.cell_box[c1]);
                rows.push_back(child_z(0, p)); rows.push_back(child_z(1, p));
                zsd.push_back(writes_zs ? sn.zs_off + p : -1);
              }
            } else {
              for (int32_t q = 0; q < f; ++q) rows.push_back(q < sn.s ? int32_t(0x80000000u | uint32_t(sn.zs_off + q)) : idx_user[n.idx_off + q]);
            }
          } else {
            fi = int32_t(it - W.fronts.begin());
          }
          if (sp.w == 0) continue;  // rows only
          all_direct = all_direct && sp.direct != 0;
          const int32_t part0 = W.n_parts;
          const int32_t rec0 = int32_t(colA.size() / 4);
          for (int32_t q = 0; q < sp.parts; ++q) {
            const int32_t r0 = int32_t(int64_t(rlen) * q / sp.parts), r1 = int32_t(int64_t(rlen) * (q + 1) / sp.parts);
            SweepMt m;
            std::memset(&m, 0, sizeof m);
            m.tile_off = int32_t(tile); m.row0 = row0_of[fi] + r0; m.nrows = r1 - r0; m.ncols = sp.w;
            m.part_off = W.n_parts;
            m.col0 = rec0; m.direct = sp.direct;
            mts.push_back(m);
            RepackDesc d;
            d.dst = tile;  // relative to the wave; rebased below
            d.src = pass == 0 ? n.rf_off + int64_t(r0) * n.ldf : n.rb_off + r0;
            d.ld = pass == 0 ? n.ldf : n.ldr;
            d.rlen = r1 - r0; d.col0 = sp.c0; d.cw = sp.w; d.ncols = ncols; d.mode = pass;
            reps.push_back(d);
            tile += even(int64_t(r1 - r0) * sp.w);
            W.n_parts += sp.w;
          }
          for (int32_t c = 0; c < sp.w; ++c) {
            const int32_t col = sp.c0 + c;  // < ncols: strips never pad
            colA.push_back(part0 + c); colA.push_back((sp.direct ? 0 : sp.parts) | (sp.w << 16));
            if (pass == 0) {
              colA.push_back(sn.zc_off + col); colA.push_back(child_z(0, sn.s + col));
              colB.push_back(child_z(1, sn.s + col));
            } else {
              colA.push_back(idx_user[n.idx_off + col]); colA.push_back(0);
            }
          }
        }
        // balance the micro-tiles over the warps: largest first, dealt round-robin (warp w takes m = w, w + nw, ...)
        std::vector<int32_t> order(mts.size());
        for (size_t m = 0; m < order.size(); ++m) order[m] = int32_t(m);
        std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) {
          return int64_t(mts[x].nrows) * mts[x].ncols > int64_t(mts[y].nrows) * mts[y].ncols;
        });
        W.n_rows = int32_t(pass == 0 ? rows.size() / 4 : rows.size());
        if (W.n_rows <= L.ws_rows) {
          W.flags |= SWF_ROWS_WS;
          L.wave_ws_rows = std::max(L.wave_ws_rows, W.n_rows);
          if (L.ng == 1 && W.fronts == prev_fronts) {  // the previous wave of this CTA left exactly these rows in place
            W.flags |= SWF_REUSE;
            rows.clear();
            zsd.clear();
          }
          prev_fronts = W.fronts;
        } else {
          prev_fronts.clear();
        }
        if (all_direct && !mts.empty()) W.flags |= SWF_DIRECT;
        W.n_mt = int32_t(mts.size());
        W.n_cols = int32_t(colA.size() / 4);
        W.tile_src = rs; W.tile_dbl = int32_t(tile);
        for (auto &d : reps) { d.dst += rs; L.repack.push_back(d); }
        rs += tile;
        // segment 0
        auto pad4 = [&] { while (L.index.size() % 4) L.index.push_back(0); };
        pad4();
        W.idx_off = int64_t(L.index.size());
        auto here = [&] { pad4(); return int32_t(int64_t(L.index.size()) - W.idx_off); };
        W.off_rows = here(); L.index.insert(L.index.end(), rows.begin(), rows.end());
        W.off_zs = here(); L.index.insert(L.index.end(), zsd.begin(), zsd.end());
        W.off_colA = here(); L.index.insert(L.index.end(), colA.begin(), colA.end());
        W.off_colB = here(); L.index.insert(L.index.end(), colB.begin(), colB.end());
        W.off_mt = here();
        for (int32_t m : order) {
          const int32_t *raw = reinterpret_cast<const int32_t *>(&mts[m]);
          L.index.insert(L.index.end(), raw, raw + sizeof(SweepMt) / 4);
        }
        W.idx_len = here();
        W.sig_off = int64_t(L.index.size());
        W.n_sig = 0;
        for (int32_t fnode : W.fronts) {
          if (L.ng == 1 && last_strip[fnode] >= i) continue;  // a later wave of this CTA finishes the front
          L.index.push_back(2 * fnode + pass);
          (pass == 0 ? L.nodes[fnode].n_fwd : L.nodes[fnode].n_bwd) += 1;
          ++W.n_sig;
        }
        pad4();
        waves.push_back(std::move(W));
      }
    }
    // (the counts n_fwd / n_bwd of this level are final now: consumers are planned after their producers)
    for (const auto &W : waves)
      for (int32_t fnode : W.fronts) add_producer(2 * fnode + pass, W.cta);
    return nullptr;
  };
  for (size_t h = 0; h < nh; ++h)
    if (const char *err = plan_level(h, 0)) return err;
  for (size_t h = nh; h-- > 0;)
    if (const char *err = plan_level(h, 1)) return err;
  for (const auto &sn : L.nodes)
    if (std::max(sn.n_fwd, sn.n_bwd) >= (1 << (31 - SWEEP_DEP_SHIFT))) return "too many waves per front";
  if (2 * L.nodes.size() >= (size_t(1) << SWEEP_DEP_SHIFT)) return "too many completion counters";
  L.prod_cta = prod_cta;
  // ---- subtree blobs: per pack the forward blobs of the members, then their backward blobs; a shared pack stores
  // one forward and one backward blob ----
  double logical_extra = 0;
  std::vector<uint8_t> writes_blob(tr.instances.size(), 1);
  for (SweepPackH &P : L.packs) {
    for (int pass = 0; pass < 2; ++pass)
      for (int32_t m = 0; m < P.n; ++m) {
        const int32_t k = L.inst_at_pos[P.pos0 + m];
        const int32_t tm = tr.instances[k].tmpl;
        const int32_t sz = pass == 0 ? tr.templates[tm].rf_size : L.rb2_size[tm];
        if (P.shared && m > 0) {
          const int32_t k0 = L.inst_at_pos[P.pos0];
          (pass == 0 ? L.inst[k].rf : L.inst[k].rb) = pass == 0 ? L.inst[k0].rf : L.inst[k0].rb;
          writes_blob[k] = 0;
          logical_extra += 8.0 * sz;
          continue;
        }
        (pass == 0 ? L.inst[k].rf : L.inst[k].rb) = rs;
        rs += sz;
      }
    P.sig_off = int64_t(L.index.size());
    for (int32_t m = 0; m < P.n; ++m) L.index.push_back(2 * (L.n_top + L.inst_at_pos[P.pos0 + m]));
    while (L.index.size() % 4) L.index.push_back(0);
  }
  // ---- completions nobody on another CTA waits for are not published: a counter is needed only when some entry
  // lists it as a dependency (same-CTA consumers rely on program order, their dependencies were elided above).  On a
  // uniform grid the shares of consecutive levels line up, so most fronts of the lower top levels are produced and
  // consumed by one CTA: their release operations (a store drain each) disappear.  Entries -1 are skipped. ----
  {
    std::vector<uint8_t> used(2 * L.nodes.size(), 0);
    const int32_t mask = (1 << SWEEP_DEP_SHIFT) - 1;
    for (const auto *lv : {&L.fwd, &L.bwd})
      for (const auto &waves : *lv)
        for (const SweepWaveH &W : waves)
          for (int32_t d : W.deps) used[d & mask] = 1;
    for (int32_t k = 0; k < n_inst; ++k) {  // subtree backward entries (build_sweep_schedule adds the dependency)
      const SweepNodeH &sn = L.nodes[L.n_top + k];
      if (sn.parent >= 0 && !(L.ng == 1 && prod_cta[2 * sn.parent + 1] == L.cta_of_inst[k])) used[2 * sn.parent + 1] = 1;
    }
    auto prune = [&](int64_t off, int32_t n) {
      for (int32_t t = 0; t < n; ++t) {
        int32_t &id = L.index[size_t(off) + t];
        if (id >= 0 && !used[id]) {
          SweepNodeH &sn = L.nodes[id >> 1];
          ((id & 1) ? sn.n_bwd : sn.n_fwd) -= 1;
          id = -1;
        }
      }
    };
    const bool stats = std::getenv("MMMFE_PLAN_STATS") != nullptr;
    int pass_id = 0;
    for (auto *lv : {&L.fwd, &L.bwd}) {
      for (size_t h = 0; h < lv->size(); ++h) {
        int64_t before = 0, after = 0, nd = 0;
        for (SweepWaveH &W : (*lv)[h]) {
          before += W.n_sig;
          prune(W.sig_off, W.n_sig);
          for (int32_t t = 0; t < W.n_sig; ++t) after += L.index[size_t(W.sig_off) + t] >= 0;
          nd += int64_t(W.deps.size());
        }
        if (stats && !(*lv)[h].empty())
          std::fprintf(stderr, "[plan] %s level %2zu: %5zu waves, %5lld cross-CTA dependencies, signals %5lld -> %5lld\n",
                       pass_id == 0 ? "fwd" : "bwd", h, (*lv)[h].size(), (long long)nd, (long long)before, (long long)after);
      }
      ++pass_id;
    }
    int64_t pb = 0, pa = 0;
    for (SweepPackH &P : L.packs) {
      pb += P.n;
      prune(P.sig_off, P.n);
      for (int32_t t = 0; t < P.n; ++t) pa += L.index[size_t(P.sig_off) + t] >= 0;
    }
    if (stats) std::fprintf(stderr, "[plan] subtree packs: signals %lld -> %lld\n", (long long)pb, (long long)pa);
  }
  for (size_t k = 0; k < tr.instances.size(); ++k) {
    if (!writes_blob[k]) continue;
    const SubtreeInstance &in = tr.instances[k];
    const SubtreeTemplate &T = tr.templates[in.tmpl];
    RepackSub d;
    d.old_rf = in.rf_base; d.old_rb = in.rb_base; d.new_rf = L.inst[k].rf; d.new_rb = L.inst[k].rb;
    d.rf_size = T.rf_size; d.rb_size = L.rb2_size[in.tmpl]; d.perm_off = perm_off[in.tmpl]; d.pad = 0;
    L.repack_sub.push_back(d);
  }
  L.rs_total = rs;
  L.bytes_per_step = 8.0 * double(rs);
  L.logical_bytes_per_step = L.bytes_per_step + logical_extra;
  for (int t = 0; t < 8; ++t) L.index.push_back(0);  // the padded tail of the last list may be read
  if (L.index.size() >= (size_t(1) << 31)) return "index lists beyond 32 bits";
  return nullptr;
}

const char *build_sweep_schedule(const NdTree &tr, const SweepLayout &L, int32_t M, int32_t ng, size_t smem_budget,
                                 double min_fill, SweepSchedule &S) {
  S = SweepSchedule();
  S.ng = ng;
  const int32_t n_cta = L.n_cta;
  // compute group of every entry: blocks (with all their chunks) and subtrees are dealt out round-robin
  std::vector<std::vector<uint8_t>> per_grp(n_cta);
  std::vector<int32_t> next_grp(n_cta, 0);
  // ---- static assignment ----
  std::vector<std::vector<SweepEntry>> per(n_cta);
  std::vector<std::vector<SweepIssue>> per_iss(n_cta);
  auto blank = [] {
    SweepEntry e;
    std::memset(&e, 0, sizeof e);
    e.sig = -1;
    return e;
  };
  auto blank_issue = [] {
    SweepIssue s;
    std::memset(&s, 0, sizeof s);
    return s;
  };
  auto pack_dep = [](int32_t counter, int32_t count) { return counter | (count << SWEEP_DEP_SHIFT); };
  auto sub_entry = [&](const SweepPackH &P, bool fwd, SweepIssue &iss) {
    const int32_t k0 = L.inst_at_pos[P.pos0];
    const SubtreeInstance &in = tr.instances[k0];
    const SubtreeTemplate &T = tr.templates[in.tmpl];
    const SweepInstH &ih = L.inst[k0];
    SweepEntry e = blank();
    iss = blank_issue();
    e.kind = fwd ? SW_SUB_FWD : SW_SUB_BWD;
    e.flags = SWF_FIRST | SWF_LAST;
    e.a0 = in.tmpl; e.a1 = -1; e.a2 = P.n; e.a7 = P.pos0;
    e.o0 = L.nodes[L.n_top + k0].zc_off; e.o1 = ih.xi_off; e.o2 = ih.pb_off;
    e.n[0] = (P.shared ? 1 : P.n) * (fwd ? T.rf_size : L.rb2_size[in.tmpl]);
    if (P.shared) e.flags |= SWF_SHARED;
    iss.off[0] = fwd ? ih.rf : ih.rb; iss.kind[0] = SRC_FACTOR;
    if (!fwd) {
      e.n[1] = P.n * even((T.n_rootb + 1) / 2);
      iss.off[1] = ih.xb / 2; iss.kind[1] = SRC_INDEX;
      e.n[2] = P.n * even(T.n_int) * M;
      iss.off[2] = int64_t(ih.xi_off) * M; iss.kind[2] = SRC_XI;
    }
    e.n[3] = P.n * even(int64_t(T.w) * T.h) * M;
    iss.off[3] = int64_t(ih.pb_off) * M; iss.kind[3] = SRC_PTIL;
    for (int t = 0; t < SWEEP_SEGS; ++t) iss.n[t] = e.n[t];
    if (fwd) {
      e.n_sig = P.n;
      e.sig_off = int32_t(P.sig_off);
    } else {
      for (int32_t m = 0; m < P.n; ++m) {
        const SweepNodeH &sn = L.nodes[L.n_top + L.inst_at_pos[P.pos0 + m]];
        if (sn.parent < 0 || (ng == 1 && L.prod_cta[2 * sn.parent + 1] == P.cta)) continue;
        const int32_t d = pack_dep(2 * sn.parent + 1, L.nodes[sn.parent].n_bwd);
        if (std::find(e.dep, e.dep + e.n_dep, d) == e.dep + e.n_dep) e.dep[e.n_dep++] = d;
      }
    }
    return e;
  };
  auto take_grp = [&](int32_t cta) {
    const int32_t g = next_grp[cta];
    next_grp[cta] = (g + 1) % ng;
    return g;
  };
  for (const SweepPackH &P : L.packs) {
    SweepIssue iss;
    per[P.cta].push_back(sub_entry(P, true, iss));
    per_iss[P.cta].push_back(iss);
    per_grp[P.cta].push_back(uint8_t(take_grp(P.cta)));
  }
  auto place_waves = [&](const std::vector<SweepWaveH> &waves) -> const char * {
    for (const SweepWaveH &W : waves) {
      SweepEntry e = blank();
      SweepIssue iss = blank_issue();
      e.kind = W.kind;
      e.flags = SWF_FIRST | SWF_LAST | W.flags;
      e.a0 = W.n_rows; e.a1 = W.n_mt; e.a2 = W.n_cols; e.a3 = W.n_parts;
      e.a4 = W.off_rows; e.a5 = W.off_mt; e.a6 = W.off_colA; e.a7 = W.height;
      e.o0 = W.off_zs; e.o3 = W.off_colB;
      e.n[0] = W.idx_len / 2; iss.off[0] = W.idx_off / 2; iss.kind[0] = SRC_INDEX; iss.n[0] = e.n[0];
      e.n[1] = W.tile_dbl; iss.off[1] = W.tile_src; iss.kind[1] = SRC_FACTOR; iss.n[1] = e.n[1];
      e.n[2] = even((int64_t((W.flags & SWF_ROWS_WS) ? 0 : W.n_rows) + W.n_parts) * M);  // scratch (not streamed)
      e.sig = -1; e.n_sig = W.n_sig;
      if (W.sig_off >= (int64_t(1) << 31)) return "index lists beyond 32 bits";
      e.sig_off = int32_t(W.sig_off);
      if (W.deps.size() > size_t(SWEEP_MAX_DEPS)) return "wave with too many dependencies";
      e.n_dep = int32_t(W.deps.size());
      for (size_t t = 0; t < W.deps.size(); ++t) e.dep[t] = W.deps[t];
      per[W.cta].push_back(e);
      per_iss[W.cta].push_back(iss);
      per_grp[W.cta].push_back(uint8_t(take_grp(W.cta)));
    }
    return nullptr;
  };
  for (size_t h = 0; h < L.fwd.size(); ++h)
    if (const char *err = place_waves(L.fwd[h])) return err;
  for (size_t h = L.bwd.size(); h-- > 0;)
    if (const char *err = place_waves(L.bwd[h])) return err;
  for (const SweepPackH &P : L.packs) {
    if (P.sig_off >= (int64_t(1) << 31)) return "index lists beyond 32 bits";
    SweepIssue iss;
    per[P.cta].push_back(sub_entry(P, false, iss));
    per_iss[P.cta].push_back(iss);
    per_grp[P.cta].push_back(uint8_t(take_grp(P.cta)));
  }

  // ---- shared-memory partition (doubles) ----
  for (const auto &list : per)
    for (const auto &e : list) S.entry_max = std::max(S.entry_max, e.n[0] + e.n[1] + e.n[2] + e.n[3]);
  const int64_t total_dbl = int64_t(smem_budget / 8);
  const int64_t fixed = sweep_fixed_dbl(M, L.ct * ng);
  for (const auto &list : per)
    if (int64_t(list.size()) > SWEEP_MAX_ENTRIES) return "too many entries per CTA for the group table";
  S.tmpl_dbl = even((L.tmpl16_max + 3) / 4);
  S.ws_sub_dbl = even(int64_t(L.sub_ws_max) * M);
  S.ws_dbl = std::max(L.pack * S.ws_sub_dbl, even(int64_t(L.wave_ws_rows) * M));
  const int64_t ring = (total_dbl - fixed - int64_t(ng) * (S.tmpl_dbl + S.ws_dbl)) & ~int64_t(1);
  if (double(ring) < std::max(1.0, min_fill) * double(S.entry_max)) {
    std::snprintf(g_msg, sizeof g_msg, "shared-memory ring (%lld doubles) is too small for entries of %d doubles",
                  (long long)ring, S.entry_max);
    return g_msg;
  }
  S.ring_dbl = int32_t(ring);
  S.smem_bytes = size_t(fixed + int64_t(ng) * (S.tmpl_dbl + S.ws_dbl) + S.ring_dbl) * 8;

  // ---- ring schedule per CTA ----
  S.cta_begin.assign(n_cta + 1, 0);
  S.cta_prologue.assign(n_cta, 0);
  for (int32_t c = 0; c < n_cta; ++c) {
    auto &list = per[c];
    const int32_t P = int32_t(list.size());
    S.cta_begin[c + 1] = S.cta_begin[c] + P;
    if (P == 0) continue;
    std::vector<int32_t> size(P), producer(P, -1);  // producer: distance back to the entry that writes what this streams
    {
      std::map<int32_t, int32_t> fwd_at, bwd_at;
      for (int32_t i = 0; i < P; ++i) {
        if (list[i].kind == SW_SUB_FWD) fwd_at[list[i].a7] = i;
        if (list[i].kind == SW_SUB_BWD) bwd_at[list[i].a7] = i;
      }
      for (int32_t i = 0; i < P; ++i) {
        if (list[i].kind == SW_SUB_BWD) producer[i] = i - fwd_at.at(list[i].a7);        // xi of this step
        if (list[i].kind == SW_SUB_FWD) producer[i] = P - (bwd_at.at(list[i].a7) - i);  // pressures of the last step
      }
    }
    int64_t pos = 0;
    for (int32_t i = 0; i < P; ++i) {
      SweepEntry &e = list[i];
      size[i] = e.n[0] + e.n[1] + e.n[2] + e.n[3];
      if (size[i] > 0) {
        if (pos + size[i] > S.ring_dbl) pos = 0;
        e.ring_off = int32_t(pos);
        pos += size[i];
      }
    }
    // dependency of every entry on the completion of an earlier one (ring space, barrier reuse, producer of a
    // streamed vector), over three periods of the periodic sequence; eff = running maximum because issue order is
    // entry order
    const int32_t U = 3 * P;
    std::vector<int64_t> eff(U);
    for (int32_t u = 0; u < U; ++u) {
      const int32_t i = u % P;
      const SweepEntry &e = list[i];
      int64_t dep = int64_t(u) - SWEEP_NBAR;
      if (producer[i] > 0) dep = std::max<int64_t>(dep, int64_t(u) - producer[i]);
      if (size[i] > 0) {
        int64_t seen = 0;
        for (int32_t v = u - 1, n = 0; v >= 0 && n < 4096 && seen < 2 * int64_t(S.ring_dbl); --v, ++n) {
          const SweepEntry &o = list[v % P];
          const int32_t osz = size[v % P];
          if (osz <= 0) continue;
          seen += osz;
          if (o.ring_off < e.ring_off + size[i] && e.ring_off < o.ring_off + osz) { dep = std::max<int64_t>(dep, v); break; }
        }
      }
      eff[u] = u > 0 ? std::max(dep, eff[u - 1]) : dep;
    }
    // steady state relative to the start of a step: rel(u') = eff[2P + u' mod P] - 2P + (u' div P) P
    auto rel = [&](int64_t up) { return eff[2 * P + up % P] - 2 * int64_t(P) + (up / P) * P; };
    int64_t ptr = 0;
    while (ptr < SWEEP_NBAR && rel(ptr) < 0) ++ptr;
    S.cta_prologue[c] = int32_t(ptr);
    int64_t upto = ptr;
    for (int32_t i = 0; i < P; ++i) {
      while (upto < int64_t(i) + 1 + SWEEP_NBAR && rel(upto) <= i) ++upto;
      list[i].issue_upto = int32_t(upto);
    }
  }
  S.entries.reserve(S.cta_begin[n_cta]);
  for (int32_t c = 0; c < n_cta; ++c)
    for (size_t i = 0; i < per[c].size(); ++i) {
      per_iss[c][i].ring_off = per[c][i].ring_off;
      S.entries.push_back(per[c][i]);
      S.issue.push_back(per_iss[c][i]);
      S.grp.push_back(per_grp[c][i]);
    }
  return nullptr;
}

const char *check_sweep_schedule(const SweepLayout &L, const SweepSchedule &S, int32_t n_steps) {
  const int32_t n_cta = L.n_cta;
  std::vector<uint32_t> cnt(2 * L.nodes.size(), 0);
  struct Cta { int64_t G = 0, issued = 0; std::map<int32_t, int32_t> fwd_at, bwd_at; };
  std::vector<Cta> st(n_cta);
  for (int32_t c = 0; c < n_cta; ++c) {
    const int32_t b0 = S.cta_begin[c], P = S.cta_begin[c + 1] - b0;
    for (int32_t i = 0; i < P; ++i) {
      const SweepEntry &e = S.entries[b0 + i];
      if (e.kind == SW_SUB_FWD) st[c].fwd_at[e.a7] = i;
      if (e.kind == SW_SUB_BWD) st[c].bwd_at[e.a7] = i;
    }
  }
  auto size_of = [](const SweepEntry &e) { return e.n[0] + e.n[1] + e.n[2] + e.n[3]; };
  auto issue_check = [&](int32_t c, int64_t upto) -> const char * {
    const int32_t b0 = S.cta_begin[c], P = S.cta_begin[c + 1] - b0;
    Cta &s = st[c];
    const int64_t total = int64_t(P) * n_steps;
    upto = std::min(upto, total);
    for (; s.issued < upto; ++s.issued) {
      const SweepEntry &e = S.entries[b0 + s.issued % P];
      const int64_t step = s.issued / P;
      if (s.issued - SWEEP_NBAR >= s.G) { std::snprintf(g_msg, sizeof g_msg, "cta %d: barrier of entry %lld re-armed before use", c, (long long)s.issued); return g_msg; }
      // streamed vectors must have been written by a retired entry
      if (e.kind == SW_SUB_BWD && step * P + s.fwd_at.at(e.a7) >= s.G) { std::snprintf(g_msg, sizeof g_msg, "cta %d: entry %lld streams xi before its forward pass retired", c, (long long)s.issued); return g_msg; }
      if (e.kind == SW_SUB_FWD && step > 0 && (step - 1) * P + s.bwd_at.at(e.a7) >= s.G) { std::snprintf(g_msg, sizeof g_msg, "cta %d: entry %lld streams pressures before the last step's backward pass retired", c, (long long)s.issued); return g_msg; }
      const int32_t sz = size_of(e);
      if (sz <= 0) continue;
      bool bad = e.ring_off < 0 || e.ring_off + sz > S.ring_dbl || (e.ring_off & 1);
      for (int t = 0; t < SWEEP_SEGS; ++t) bad |= (e.n[t] & 1) != 0;
      if (bad) { std::snprintf(g_msg, sizeof g_msg, "cta %d: ring range of entry %lld out of bounds or misaligned", c, (long long)s.issued); return g_msg; }
      for (int64_t v = s.G; v < s.issued; ++v) {
        const SweepEntry &o = S.entries[b0 + v % P];
        const int32_t osz = size_of(o);
        if (osz > 0 && o.ring_off < e.ring_off + sz && e.ring_off < o.ring_off + osz) {
          std::snprintf(g_msg, sizeof g_msg, "cta %d: entry %lld overwrites the ring range of unfinished entry %lld", c, (long long)s.issued, (long long)v);
          return g_msg;
        }
      }
    }
    return nullptr;
  };
  for (int32_t c = 0; c < n_cta; ++c)
    if (S.cta_begin[c + 1] > S.cta_begin[c])
      if (const char *err = issue_check(c, S.cta_prologue[c])) return err;
  int64_t remaining = 0;
  for (int32_t c = 0; c < n_cta; ++c) remaining += int64_t(S.cta_begin[c + 1] - S.cta_begin[c]) * n_steps;
  while (remaining > 0) {
    bool progress = false;
    for (int32_t c = 0; c < n_cta; ++c) {
      const int32_t b0 = S.cta_begin[c], P = S.cta_begin[c + 1] - b0;
      Cta &s = st[c];
      while (P > 0 && s.G < int64_t(P) * n_steps) {
        const int64_t step = s.G / P;
        const SweepEntry &e = S.entries[b0 + s.G % P];
        if (s.issued <= s.G) { std::snprintf(g_msg, sizeof g_msg, "cta %d: entry %lld consumed before it was issued", c, (long long)s.G); return g_msg; }
        bool blocked = false;
        for (int32_t t = 0; t < e.n_dep; ++t) {
          const int32_t counter = e.dep[t] & ((1 << SWEEP_DEP_SHIFT) - 1), count = e.dep[t] >> SWEEP_DEP_SHIFT;
          if (cnt[counter] < uint32_t(step + 1) * uint32_t(count)) blocked = true;
        }
        if (blocked) break;
        if (e.sig >= 0) ++cnt[e.sig];
        for (int32_t t = 0; t < e.n_sig; ++t)
          if (L.index[size_t(e.sig_off) + t] >= 0) ++cnt[L.index[size_t(e.sig_off) + t]];
        ++s.G;
        --remaining;
        progress = true;
        if (const char *err = issue_check(c, step * P + e.issue_upto)) return err;
      }
    }
    if (!progress) { std::snprintf(g_msg, sizeof g_msg, "deadlock: %lld entries cannot complete", (long long)remaining); return g_msg; }
  }
  for (size_t n = 0; n < L.nodes.size(); ++n) {
    if (cnt[2 * n] != uint32_t(n_steps) * uint32_t(L.nodes[n].n_fwd)) { std::snprintf(g_msg, sizeof g_msg, "forward counter of node %zu ends at %u", n, cnt[2 * n]); return g_msg; }
    if (int32_t(n) < L.n_top && cnt[2 * n + 1] != uint32_t(n_steps) * uint32_t(L.nodes[n].n_bwd)) { std::snprintf(g_msg, sizeof g_msg, "backward counter of node %zu ends at %u", n, cnt[2 * n + 1]); return g_msg; }
  }
  return nullptr;
}

}  // namespace mmmfe

// ---- diagnostics of include/mmmfe_b200.h (host only) ---------------------------------------------------------------
extern "C" int mmmfe_sweep_plan_check(int32_t nx, int32_t ny, int32_t leaf_cells, int32_t subtree_cells, int32_t M,
                                      int32_t n_cta, int64_t smem_bytes, int32_t n_steps, int32_t ct,
                                      int32_t chunk_dbl, int64_t stats[8], char *msg, int32_t msg_cap) {
  auto say = [&](const char *m) {
    if (msg && msg_cap > 0) std::snprintf(msg, size_t(msg_cap), "%s", m ? m : "");
  };
  say("");
  if (nx < 1 || ny < 1 || n_cta < 1 || (M != 1 && M != 2 && M != 4) || n_steps < 1 || chunk_dbl < 64 ||
      (ct != 64 && ct != 128 && ct != 256)) { say("bad argument"); return -1; }
  mmmfe::NdTree t;
  t.build(nx, ny, leaf_cells, subtree_cells);
  mmmfe::SweepLayout L;
  const int32_t pack = std::getenv("MMMFE_PLAN_PACK") ? std::atoi(std::getenv("MMMFE_PLAN_PACK")) : 1;
  std::vector<int32_t> cls;  // MMMFE_PLAN_SHARE: plan as if all subtrees of one shape had identical factors (K = I grids)
  if (std::getenv("MMMFE_PLAN_SHARE"))
    for (const auto &in : t.instances) cls.push_back(in.tmpl);
  if (const char *e = mmmfe::build_sweep_layout(t, t.idx, std::vector<int32_t>(), n_cta, chunk_dbl, ct, L, 1, 512, pack, cls)) { say(e); return -2; }
  if (std::getenv("MMMFE_PLAN_DEBUG")) {
    for (int pass = 0; pass < 2; ++pass)
      for (size_t h = 0; h < L.fwd.size(); ++h) {
        const auto &ws = pass == 0 ? L.fwd[h] : L.bwd[h];
        if (ws.empty()) continue;
        int64_t tile = 0, rows = 0, mts = 0, deps = 0, fr = 0;
        std::vector<int> per(n_cta, 0);
        for (const auto &w : ws) { tile += w.tile_dbl; rows += w.n_rows; mts += w.n_mt; deps += int64_t(w.deps.size()); fr += int64_t(w.fronts.size()); ++per[w.cta]; }
        std::fprintf(stderr, "%s h %2zu: waves %5zu (max/cta %d) tile/wave %6lld rows/wave %5lld mt/wave %4.1f fronts/wave %4.1f deps/wave %4.2f\n",
                     pass == 0 ? "fwd" : "bwd", h, ws.size(), *std::max_element(per.begin(), per.end()), (long long)(tile / int64_t(ws.size())),
                     (long long)(rows / int64_t(ws.size())), double(mts) / ws.size(), double(fr) / ws.size(), double(deps) / ws.size());
      }
  }
  mmmfe::SweepSchedule S;
  if (const char *e = mmmfe::build_sweep_schedule(t, L, M, 1, size_t(smem_bytes), std::getenv("MMMFE_PLAN_FILL") ? std::atof(std::getenv("MMMFE_PLAN_FILL")) : 2.0, S)) { say(e); return -3; }
  if (const char *path = std::getenv("MMMFE_PLAN_DUMP")) {  // per entry, CTA-major: features for cost modelling
    if (FILE *f = std::fopen(path, "w")) {
      for (int32_t c = 0; c < n_cta; ++c)
        for (int32_t i = S.cta_begin[c]; i < S.cta_begin[c + 1]; ++i) {
          const mmmfe::SweepEntry &e = S.entries[size_t(i)];
          const bool wave = e.kind == mmmfe::SW_TOP_FWD || e.kind == mmmfe::SW_TOP_BWD;
          int32_t max_load = 0, n_w = 8;
          if (wave) {  // largest per-warp micro-tile load (doubles), warps deal m = w, w + 8, ...
            std::vector<int32_t> load(n_w, 0);
            const int32_t *ix = L.index.data() + 2 * S.issue[size_t(i)].off[0];
            const mmmfe::SweepMt *mts = reinterpret_cast<const mmmfe::SweepMt *>(ix + e.a5);
            for (int32_t m = 0; m < e.a1; ++m) load[m % n_w] += mts[m].nrows * mts[m].ncols;
            max_load = *std::max_element(load.begin(), load.end());
          }
          std::fprintf(f, "%d %d %d %d %d %d %d %d %d %d %d\n", c, e.kind, wave ? e.a7 : 0, e.flags, wave ? e.a0 : 0,
                       wave ? e.a1 : 0, wave ? e.a2 : 0, wave ? e.a3 : 0, e.n[1], max_load, e.n_dep);
        }
      std::fclose(f);
    }
  }
  if (stats) {
    stats[0] = int64_t(L.nodes.size()); stats[1] = L.n_top; stats[2] = int64_t(S.entries.size()); stats[3] = L.rs_total;
    stats[4] = S.ring_dbl; stats[5] = S.ws_dbl; stats[6] = S.entry_max; stats[7] = int64_t(S.smem_bytes);
  }
  if (const char *e = mmmfe::check_sweep_schedule(L, S, n_steps)) { say(e); return -4; }
  return 0;
}
