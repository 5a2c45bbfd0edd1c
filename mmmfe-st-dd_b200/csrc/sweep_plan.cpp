#include "sweep_plan.h"

#include <algorithm>
#include <cstdio>
#include <cstring>

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

char g_msg[256];

// uint16 arrays of one template behind a 32-word int32 header (offsets in uint16 units from the blob start)
const char *pack_template(const SubtreeTemplate &T, std::vector<uint16_t> &out) {
  if (T.w > 127 || T.h > 127) return "subtree box wider than 127 cells";
  const size_t base = out.size();
  out.resize(base + 2 * SWEEP_TMPL_HDR, 0);
  std::vector<int32_t> hdr(SWEEP_TMPL_HDR, 0);
  hdr[0] = T.n_nodes; hdr[1] = T.n_levels; hdr[2] = T.n_int; hdr[3] = T.n_rootb; hdr[4] = T.zl_size;
  hdr[5] = T.rf_size; hdr[6] = T.rb_size; hdr[7] = T.w; hdr[8] = T.h;
  bool overflow = false;
  auto put = [&](int slot, const std::vector<int32_t> &v, bool is_code) {
    hdr[slot] = int32_t(out.size() - base);
    for (int32_t x : v) {
      int32_t y = x;
      if (is_code) y = ((x >> 30) << 14) | (((x >> 15) & 0x7fff) << 7) | (x & 0x7fff);
      if (y < -1 || y >= 0xFFFF) overflow = true;
      out.push_back(y < 0 ? uint16_t(0xFFFF) : uint16_t(y));
    }
  };
  put(10, T.node_s, false); put(11, T.node_b, false); put(12, T.node_rf, false); put(13, T.node_rb, false);
  put(14, T.node_sep, false); put(15, T.node_bnd, false); put(16, T.node_zl, false); put(17, T.var_code, true);
  put(18, T.rootb_code, true); put(19, T.sep_node, false); put(20, T.sep_c0, false); put(21, T.sep_c1, false);
  put(22, T.out_node, false); put(23, T.out_c0, false); put(24, T.out_c1, false); put(25, T.bnd_var, false);
  put(26, T.lvl_sep, false); put(27, T.lvl_out, false); put(28, T.cell_edge, false);
  while ((out.size() - base) % 8) out.push_back(0);  // 16-byte granules
  hdr[9] = int32_t(out.size() - base);
  std::memcpy(out.data() + base, hdr.data(), SWEEP_TMPL_HDR * sizeof(int32_t));
  return overflow ? "subtree template field beyond 16 bits" : nullptr;
}

}  // namespace

const char *build_sweep_layout(const NdTree &tr, int32_t n_cta, int32_t chunk_dbl, SweepLayout &L) {
  L = SweepLayout();
  L.n_cta = n_cta;
  L.chunk_dbl = chunk_dbl;
  const int32_t nn = int32_t(tr.nodes.size());
  if (tr.instances.empty()) return "no subtrees (subtree_cells = 0)";
  if (tr.idx.size() >= (size_t(1) << 30) || tr.src.size() >= (size_t(1) << 31)) return "index structure beyond 32 bits";
  const int64_t nxe = int64_t(tr.nx + 1) * tr.ny;
  std::vector<int32_t> sid(nn, -1);
  for (int32_t i = 0; i < nn; ++i) {
    const NdNode &n = tr.nodes[i];
    if (n.subtree >= 0) continue;
    if (n.child[0] < 0 && n.child[1] < 0) return "a leaf front lies outside every subtree";
    sid[i] = int32_t(L.nodes.size());
    SweepNodeH h;
    h.old_id = i; h.s = n.s; h.b = n.b; h.height = n.height;
    h.idx_off = int32_t(n.idx_off); h.src_off = int32_t(n.src_off);
    h.sep_type = tr.idx[n.idx_off] < nxe ? 0 : 1;
    L.nodes.push_back(h);
    L.f_max = std::max(L.f_max, n.s + n.b);
  }
  L.n_top = int32_t(L.nodes.size());
  L.inst_xi.resize(tr.instances.size());
  for (size_t k = 0; k < tr.instances.size(); ++k) {
    const SubtreeInstance &in = tr.instances[k];
    const SubtreeTemplate &T = tr.templates[in.tmpl];
    sid[in.root] = int32_t(L.nodes.size());
    SweepNodeH h;
    h.old_id = in.root; h.s = 0; h.b = T.n_rootb; h.height = tr.nodes[in.root].height; h.n_fwd = 1; h.n_bwd = 1;
    L.nodes.push_back(h);
    L.inst_xi[k] = int32_t(L.xi_total);
    L.xi_total += T.n_int;
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
    if (h.s > 0) { h.zs_off = int32_t(L.z_total); L.z_total += h.s; }
    h.zc_off = int32_t(L.z_total);
    L.z_total += h.b;
  }
  if (L.z_total >= (int64_t(1) << 30) || L.xi_total >= (int64_t(1) << 30)) return "vector storage beyond 32 bits";

  // cells of the separator edges of the top fronts (right-hand side -K_up p_old is formed on the fly)
  L.top_cell.assign(2 * tr.idx.size(), -1);
  for (int32_t t = 0; t < L.n_top; ++t) {
    const NdNode &n = tr.nodes[L.nodes[t].old_id];
    for (int32_t p = 0; p < n.s; ++p) {
      const int64_t v = tr.idx[n.idx_off + p];
      int32_t c0, c1;
      if (v < nxe) {
        const int32_t j = int32_t(v / (tr.nx + 1)), i = int32_t(v % (tr.nx + 1));
        if (i <= 0 || i >= tr.nx) return "separator edge on the domain boundary";
        c0 = j * tr.nx + i - 1; c1 = j * tr.nx + i;
      } else {
        const int64_t k = v - nxe;
        const int32_t j = int32_t(k / tr.nx), i = int32_t(k % tr.nx);
        if (j <= 0 || j >= tr.ny) return "separator edge on the domain boundary";
        c0 = (j - 1) * tr.nx + i; c1 = j * tr.nx + i;
      }
      L.top_cell[2 * (n.idx_off + p)] = c0;
      L.top_cell[2 * (n.idx_off + p) + 1] = c1;
    }
  }

  // ---- top items: column blocks, sized per level so that every CTA gets work ----
  const size_t nh = tr.top_by_height.size();
  L.fwd.assign(nh, {});
  L.bwd.assign(nh, {});
  int64_t rs = 0;
  for (size_t h = 0; h < nh; ++h) {
    int64_t fdbl = 0, bdbl = 0;
    for (int32_t id : tr.top_by_height[h]) {
      const NdNode &n = tr.nodes[id];
      fdbl += int64_t(n.s) * n.b;
      bdbl += int64_t(n.s) * (n.s + n.b);
    }
    for (int pass = 0; pass < 2; ++pass) {
      const int64_t level_dbl = pass == 0 ? fdbl : bdbl;
      const int64_t target = std::min<int64_t>(chunk_dbl, std::max<int64_t>(256, level_dbl / std::max(1, n_cta)));
      for (int32_t id : tr.top_by_height[h]) {
        const NdNode &n = tr.nodes[id];
        SweepNodeH &sn = L.nodes[sid[id]];
        const int32_t rlen = pass == 0 ? n.s : n.s + n.b;
        const int32_t ncols = pass == 0 ? n.b : n.s;
        auto &items = pass == 0 ? L.fwd[h] : L.bwd[h];
        if (ncols == 0) {  // root: forward entry that only accumulates the separator right-hand side
          items.push_back({SW_TOP_FWD, sid[id], 0, 2, 0, 0});
          sn.n_fwd = 1;
          continue;
        }
        // block width: a power of two below 32 (shuffle reduction), else any even width up to 256 with at most
        // one padded column per block
        const int64_t cap = std::max<int64_t>(2, std::min<int64_t>(256, target / rlen));
        int32_t cw;
        if (cap >= ncols) cw = ncols >= 32 ? ((ncols + 1) & ~1) : std::max(2, pow2_ceil(ncols));
        else if (cap >= 32) { const int32_t nb = int32_t((ncols + cap - 1) / cap); cw = (((ncols + nb - 1) / nb) + 1) & ~1; }
        else cw = pow2_floor(cap);
        const int32_t nblk = (ncols + cw - 1) / cw;
        for (int32_t k = 0; k < nblk; ++k) {
          items.push_back({pass == 0 ? SW_TOP_FWD : SW_TOP_BWD, sid[id], k * cw, cw, rlen, rs});
          RepackDesc d;
          d.dst = rs;
          d.src = pass == 0 ? n.rf_off : n.rb_off;
          d.ld = pass == 0 ? n.ldf : n.ldr;
          d.rlen = rlen; d.col0 = k * cw; d.cw = cw; d.ncols = ncols; d.mode = pass;
          L.repack.push_back(d);
          rs += int64_t(rlen) * cw;
        }
        (pass == 0 ? sn.n_fwd : sn.n_bwd) = nblk;
      }
    }
  }
  L.bytes_per_step = 8.0 * double(rs);
  // ---- subtree blobs: one contiguous range of the front storage, copied verbatim ----
  L.sub_old_base = tr.instances.front().rf_base;
  L.sub_dbl = tr.r_total - L.sub_old_base;
  L.sub_rs_base = rs;
  rs += L.sub_dbl;
  L.rs_total = rs;
  L.bytes_per_step += 8.0 * double(L.sub_dbl);
  L.blob_max = chunk_dbl;
  for (const auto &T : tr.templates) {
    L.tmpl_off.push_back(int64_t(L.tmpl16.size()));
    const size_t before = L.tmpl16.size();
    if (const char *err = pack_template(T, L.tmpl16)) return err;
    L.tmpl16_max = std::max(L.tmpl16_max, int32_t(L.tmpl16.size() - before));
    const int32_t ncell = T.w * T.h;
    L.sub_ws_max = std::max(L.sub_ws_max, ncell + std::max(T.n_int + T.zl_size, 2 * T.n_int + T.n_rootb));
    L.blob_max = std::max(L.blob_max, std::max(T.rf_size, T.rb_size));
  }
  return nullptr;
}

const char *build_sweep_schedule(const NdTree &tr, const SweepLayout &L, int32_t M, size_t smem_budget,
                                 double min_fill, SweepSchedule &S) {
  S = SweepSchedule();
  const int32_t n_cta = L.n_cta;
  // ---- shared-memory partition (doubles) ----
  const int64_t total_dbl = int64_t(smem_budget / 8) - 16;  // static shared memory of the kernel
  const int64_t fixed = SWEEP_NBAR + 32 + int64_t(SWEEP_THREADS) * M;
  S.tmpl_dbl = int32_t(((L.tmpl16_max + 3) / 4 + 1) & ~1);
  S.ws_dbl = int32_t((int64_t(std::max(L.f_max, L.sub_ws_max)) * M + 1) & ~int64_t(1));
  const int64_t ring = (total_dbl - fixed - S.tmpl_dbl - S.ws_dbl) & ~int64_t(1);
  if (double(ring) < std::max(2.0, min_fill) * double(L.blob_max)) {
    std::snprintf(g_msg, sizeof g_msg, "shared-memory ring (%lld doubles) is too small for entries of %d doubles",
                  (long long)ring, L.blob_max);
    return g_msg;
  }
  S.ring_dbl = int32_t(ring);
  S.smem_bytes = size_t(fixed + S.tmpl_dbl + S.ws_dbl + S.ring_dbl) * 8;

  // ---- static assignment ----
  std::vector<std::vector<SweepEntry>> per(n_cta);
  auto blank = [] {
    SweepEntry e;
    std::memset(&e, 0, sizeof e);
    e.sig = e.dep0 = e.dep1 = -1;
    e.o4 = e.o5 = -1;
    return e;
  };
  const int32_t n_inst = int32_t(tr.instances.size());
  auto sub_entry = [&](int32_t k, bool fwd) {
    const SubtreeInstance &in = tr.instances[k];
    const SubtreeTemplate &T = tr.templates[in.tmpl];
    const SweepNodeH &sn = L.nodes[L.n_top + k];
    SweepEntry e = blank();
    e.kind = fwd ? SW_SUB_FWD : SW_SUB_BWD;
    e.flags = SWF_FIRST | SWF_LAST;
    e.n_dbl = fwd ? T.rf_size : T.rb_size;
    e.src = L.sub_rs_base + ((fwd ? in.rf_base : in.rb_base) - L.sub_old_base);
    e.a0 = in.tmpl; e.a1 = -1; e.a2 = in.I0; e.a3 = in.J0; e.a7 = k;
    e.o0 = sn.zc_off; e.o1 = L.inst_xi[k];
    if (fwd) e.sig = 2 * (L.n_top + k);
    else if (sn.parent >= 0) { e.dep0 = 2 * sn.parent + 1; e.dep0_n = L.nodes[sn.parent].n_bwd; }
    return e;
  };
  auto cta_of_inst = [&](int32_t k) { return int32_t(int64_t(k) * n_cta / n_inst); };
  for (int32_t k = 0; k < n_inst; ++k) per[cta_of_inst(k)].push_back(sub_entry(k, true));

  int32_t rot = 0;
  auto place_level = [&](const std::vector<SweepItem> &items) {
    if (items.empty()) return;
    int64_t level_dbl = 0;
    for (const auto &it : items) level_dbl += int64_t(it.rlen) * it.cw;
    int64_t cum = 0;
    int32_t used = 0;
    for (const auto &it : items) {
      const int64_t sz = int64_t(it.rlen) * it.cw;
      int32_t slot = level_dbl > 0 ? int32_t((cum + sz / 2) * n_cta / level_dbl) : 0;
      slot = std::min(slot, n_cta - 1);
      used = std::max(used, slot + 1);
      cum += sz;
      auto &list = per[(slot + rot) % n_cta];
      const SweepNodeH &sn = L.nodes[it.node];
      const bool fwd = it.kind == SW_TOP_FWD;
      const int32_t cw = it.cw;
      const int32_t rows_per = std::max(1, L.chunk_dbl / cw);
      int32_t r = 0;
      do {
        const int32_t r1 = std::min(it.rlen, r + rows_per);
        SweepEntry e = blank();
        e.kind = it.kind;
        e.flags = (r == 0 ? SWF_FIRST : 0) | (r1 >= it.rlen ? SWF_LAST : 0);
        if (r == 0) {
          const bool same_run = !list.empty() && list.back().kind == it.kind && list.back().a7 == it.node;
          if (!same_run) e.flags |= SWF_NEW_RUN;
          if (fwd && it.col0 == 0) e.flags |= SWF_WRITE_ZS;
        }
        e.n_dbl = (r1 - r) * cw;
        e.src = it.src + int64_t(r) * cw;
        e.a0 = it.col0; e.a1 = it.cw; e.a2 = r; e.a3 = r1; e.a4 = sn.s; e.a5 = sn.b; e.a6 = sn.sep_type;
        e.a7 = it.node;
        e.o0 = sn.idx_off; e.o1 = sn.src_off; e.o2 = sn.zs_off; e.o3 = sn.zc_off;
        e.o4 = sn.child[0] >= 0 ? L.nodes[sn.child[0]].zc_off : -1;
        e.o5 = sn.child[1] >= 0 ? L.nodes[sn.child[1]].zc_off : -1;
        if (fwd) {
          e.sig = 2 * it.node;
          if (sn.child[0] >= 0) { e.dep0 = 2 * sn.child[0]; e.dep0_n = L.nodes[sn.child[0]].n_fwd; }
          if (sn.child[1] >= 0) { e.dep1 = 2 * sn.child[1]; e.dep1_n = L.nodes[sn.child[1]].n_fwd; }
        } else {
          e.sig = 2 * it.node + 1;
          if (sn.parent >= 0) { e.dep0 = 2 * sn.parent + 1; e.dep0_n = L.nodes[sn.parent].n_bwd; }
          else { e.dep0 = 2 * it.node; e.dep0_n = sn.n_fwd; }
        }
        list.push_back(e);
        r = r1;
      } while (r < it.rlen);
    }
    rot = (rot + used + 1) % n_cta;
  };
  for (size_t h = 0; h < L.fwd.size(); ++h) place_level(L.fwd[h]);
  for (size_t h = L.bwd.size(); h-- > 0;) place_level(L.bwd[h]);
  for (int32_t k = 0; k < n_inst; ++k) per[cta_of_inst(k)].push_back(sub_entry(k, false));

  // ---- ring schedule per CTA ----
  S.cta_begin.assign(n_cta + 1, 0);
  S.cta_prologue.assign(n_cta, 0);
  for (int32_t c = 0; c < n_cta; ++c) {
    auto &list = per[c];
    const int32_t P = int32_t(list.size());
    S.cta_begin[c + 1] = S.cta_begin[c] + P;
    if (P == 0) continue;
    int64_t pos = 0;
    for (auto &e : list) {
      if (e.n_dbl > 0) {
        if (pos + e.n_dbl > S.ring_dbl) pos = 0;
        e.ring_off = int32_t(pos);
        pos += e.n_dbl;
      }
    }
    // dependency of every entry on the completion of an earlier one (ring space and barrier reuse), over three
    // periods of the periodic sequence; eff = running maximum because issue order is entry order
    const int32_t U = 3 * P;
    std::vector<int64_t> eff(U);
    for (int32_t u = 0; u < U; ++u) {
      const SweepEntry &e = list[u % P];
      int64_t dep = int64_t(u) - SWEEP_NBAR;
      if (e.n_dbl > 0) {
        int64_t seen = 0;
        for (int32_t v = u - 1, n = 0; v >= 0 && n < 4096 && seen < 2 * int64_t(S.ring_dbl); --v, ++n) {
          const SweepEntry &o = list[v % P];
          if (o.n_dbl <= 0) continue;
          seen += o.n_dbl;
          if (o.ring_off < e.ring_off + e.n_dbl && e.ring_off < o.ring_off + o.n_dbl) { dep = std::max<int64_t>(dep, v); break; }
        }
      }
      eff[u] = u > 0 ? std::max(dep, eff[u - 1]) : dep;
    }
    // steady state relative to the start of a step: rel(u') = eff[2P + u' mod P] - 2P + (u' div P) P
    auto rel = [&](int64_t up) { return eff[2 * P + up % P] - 2 * int64_t(P) + (up / P) * P; };
    int64_t ptr = 0;
    while (ptr < 2 * int64_t(P) && rel(ptr) < 0) ++ptr;
    S.cta_prologue[c] = int32_t(std::min<int64_t>(ptr, SWEEP_NBAR));
    int64_t upto = S.cta_prologue[c];
    for (int32_t i = 0; i < P; ++i) {
      const int64_t first = upto;
      while (upto < int64_t(i) + 1 + SWEEP_NBAR && rel(upto) <= i) ++upto;
      if (upto < i + 2 && P > 1) upto = i + 2;  // cannot happen: entry i + 1 only depends on entries <= i
      list[i].issue_upto = int32_t(upto);
      const SweepEntry &t = list[first % P];
      list[i].t_src = t.src; list[i].t_n_dbl = t.n_dbl; list[i].t_ring_off = t.ring_off;
    }
  }
  S.entries.reserve(S.cta_begin[n_cta]);
  for (int32_t c = 0; c < n_cta; ++c)
    for (const auto &e : per[c]) {
      S.entries.push_back(e);
      S.issue.push_back({e.src, e.n_dbl, e.ring_off});
    }
  return nullptr;
}

const char *check_sweep_schedule(const SweepLayout &L, const SweepSchedule &S, int32_t n_steps) {
  const int32_t n_cta = L.n_cta;
  std::vector<uint32_t> cnt(2 * L.nodes.size(), 0);
  struct Cta { int64_t G = 0, issued = 0; };
  std::vector<Cta> st(n_cta);
  auto issue_check = [&](int32_t c, int64_t upto) -> const char * {
    const int32_t b0 = S.cta_begin[c], P = S.cta_begin[c + 1] - b0;
    Cta &s = st[c];
    const int64_t total = int64_t(P) * n_steps;
    upto = std::min(upto, total);
    for (; s.issued < upto; ++s.issued) {
      const SweepEntry &e = S.entries[b0 + s.issued % P];
      if (s.issued - SWEEP_NBAR >= s.G) { std::snprintf(g_msg, sizeof g_msg, "cta %d: barrier of entry %lld re-armed before use", c, (long long)s.issued); return g_msg; }
      if (e.n_dbl <= 0) continue;
      if (e.ring_off < 0 || e.ring_off + e.n_dbl > S.ring_dbl || (e.ring_off & 1) || (e.n_dbl & 1)) { std::snprintf(g_msg, sizeof g_msg, "cta %d: ring range of entry %lld out of bounds or misaligned", c, (long long)s.issued); return g_msg; }
      for (int64_t v = s.G; v < s.issued; ++v) {
        const SweepEntry &o = S.entries[b0 + v % P];
        if (o.n_dbl > 0 && o.ring_off < e.ring_off + e.n_dbl && e.ring_off < o.ring_off + o.n_dbl) {
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
        const bool waits = (e.kind == SW_SUB_BWD) || ((e.kind == SW_TOP_FWD || e.kind == SW_TOP_BWD) && (e.flags & SWF_NEW_RUN));
        if (waits) {
          if (e.dep0 >= 0 && cnt[e.dep0] < uint32_t(step + 1) * uint32_t(e.dep0_n)) break;
          if (e.dep1 >= 0 && cnt[e.dep1] < uint32_t(step + 1) * uint32_t(e.dep1_n)) break;
        }
        if ((e.flags & SWF_LAST) && e.sig >= 0) ++cnt[e.sig];
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
  }
  return nullptr;
}

}  // namespace mmmfe

// ---- diagnostics of include/mmmfe_b200.h (host only) ---------------------------------------------------------------
extern "C" int mmmfe_sweep_plan_check(int32_t nx, int32_t ny, int32_t leaf_cells, int32_t subtree_cells, int32_t M,
                                      int32_t n_cta, int64_t smem_bytes, int32_t n_steps, int64_t stats[8], char *msg,
                                      int32_t msg_cap) {
  auto say = [&](const char *m) {
    if (msg && msg_cap > 0) std::snprintf(msg, size_t(msg_cap), "%s", m ? m : "");
  };
  say("");
  if (nx < 1 || ny < 1 || n_cta < 1 || (M != 1 && M != 2 && M != 4) || n_steps < 1) { say("bad argument"); return -1; }
  mmmfe::NdTree t;
  t.build(nx, ny, leaf_cells, subtree_cells);
  mmmfe::SweepLayout L;
  if (const char *e = mmmfe::build_sweep_layout(t, n_cta, 2048, L)) { say(e); return -2; }
  mmmfe::SweepSchedule S;
  if (const char *e = mmmfe::build_sweep_schedule(t, L, M, size_t(smem_bytes), 2.0, S)) { say(e); return -3; }
  if (stats) {
    stats[0] = int64_t(L.nodes.size()); stats[1] = L.n_top; stats[2] = int64_t(S.entries.size()); stats[3] = L.rs_total;
    stats[4] = S.ring_dbl; stats[5] = S.ws_dbl; stats[6] = L.blob_max; stats[7] = int64_t(S.smem_bytes);
  }
  if (const char *e = mmmfe::check_sweep_schedule(L, S, n_steps)) { say(e); return -4; }
  return 0;
}
