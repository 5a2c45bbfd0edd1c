#include "exchange_plan.h"

#include <algorithm>
#include <cstdio>
#include <map>
#include <utility>

#include "../../include/mmmfe_b200.h"

namespace mmmfe {

namespace {
const int kOppositeSide[4] = {2, 3, 0, 1};
char g_plan_msg[160];
}  // namespace

const char *build_exchange_plan(const std::vector<ExchangeSide> &sides, int64_t n_iface,
                                const std::vector<int32_t> &owner, int32_t rank, int32_t world, ExchangePlan &P) {
  P = ExchangePlan();
  P.partner.assign(size_t(n_iface), -1);
  std::map<std::pair<int32_t, int32_t>, int64_t> local_off;
  for (const ExchangeSide &s : sides) {
    if (s.side < 0 || s.side > 3 || s.off < 0 || s.len < 0 || s.off + s.len > n_iface) return "interface side out of range";
    local_off[{s.gid, s.side}] = s.off;
  }
  struct Seg { int32_t peer, key_gid, key_side; int64_t off, len; };
  std::vector<Seg> send_segs, recv_segs;
  for (const ExchangeSide &s : sides) {
    int32_t own = rank;
    if (world > 1 && !owner.empty()) {
      if (s.neighbor < 0 || size_t(s.neighbor) >= owner.size()) return "neighbour id outside the owner table";
      own = owner[s.neighbor];
    }
    if (own == rank) {
      auto it = local_off.find({s.neighbor, kOppositeSide[s.side]});
      if (it == local_off.end()) {
        std::snprintf(g_plan_msg, sizeof g_plan_msg, "neighbour %d of subdomain %d is not registered", s.neighbor, s.gid);
        return g_plan_msg;
      }
      for (int64_t i = 0; i < s.len; ++i) P.partner[s.off + i] = it->second + i;
    } else {
      send_segs.push_back({own, s.gid, s.side, s.off, s.len});
      recv_segs.push_back({own, s.neighbor, kOppositeSide[s.side], s.off, s.len});  // keyed like the peer's send
    }
  }
  auto by_key = [](const Seg &a, const Seg &b) {
    if (a.peer != b.peer) return a.peer < b.peer;
    if (a.key_gid != b.key_gid) return a.key_gid < b.key_gid;
    return a.key_side < b.key_side;
  };
  std::sort(send_segs.begin(), send_segs.end(), by_key);
  std::sort(recv_segs.begin(), recv_segs.end(), by_key);
  int64_t roff = 0;
  for (size_t i = 0; i < send_segs.size(); ++i) {
    const Seg &sg = send_segs[i], &rg = recv_segs[i];
    if (P.peers.empty() || P.peers.back().rank != sg.peer) P.peers.push_back({sg.peer, int64_t(P.send_index.size()), roff, 0});
    for (int64_t k = 0; k < sg.len; ++k) P.send_index.push_back(sg.off + k);
    for (int64_t k = 0; k < rg.len; ++k) P.partner[rg.off + k] = -(roff + k) - 2;
    roff += rg.len;
    P.peers.back().count += sg.len;
  }
  for (int64_t v : P.partner)
    if (v == -1) return "interface entry without a partner";
  return nullptr;
}

// Contiguous blocks of subdomains in rank order (the reference launches rank r on subdomain r, README.md:64) that
// minimise the largest block cost; every rank below min(world, n) owns at least one subdomain, owners are
// non-decreasing.  Exact linear-partition dynamic programme (n <= a few thousand, world <= 8: microseconds).
void assign_owners(const std::vector<double> &cost, int32_t world, std::vector<int32_t> &owner) {
  const int n = int(cost.size());
  owner.assign(cost.size(), 0);
  if (world <= 1 || n == 0) return;
  const int w = std::min(world, n);
  std::vector<double> pre(size_t(n) + 1, 0.0);
  for (int i = 0; i < n; ++i) pre[i + 1] = pre[i] + std::max(cost[i], 0.0);
  // best[k][i] = smallest possible largest-block cost when the first i subdomains go to k non-empty blocks
  const double inf = 1e300;
  std::vector<std::vector<double>> best(size_t(w) + 1, std::vector<double>(size_t(n) + 1, inf));
  std::vector<std::vector<int>> cut(size_t(w) + 1, std::vector<int>(size_t(n) + 1, 0));
  best[0][0] = 0.0;
  for (int k = 1; k <= w; ++k)
    for (int i = k; i <= n - (w - k); ++i)
      for (int j = k - 1; j < i; ++j) {
        if (best[k - 1][j] >= inf) continue;
        const double v = std::max(best[k - 1][j], pre[i] - pre[j]);
        // ties: the later cut (more even block sizes towards the end of the rank order)
        if (v < best[k][i] || (v == best[k][i] && j > cut[k][i])) { best[k][i] = v; cut[k][i] = j; }
      }
  int i = n;
  for (int k = w; k >= 1; --k) {
    const int j = cut[k][i];
    for (int r = j; r < i; ++r) owner[r] = k - 1;
    i = j;
  }
}

}  // namespace mmmfe

// ---- host-only diagnostics of include/mmmfe_b200.h ---------------------------------------------------------------
extern "C" int mmmfe_exchange_plan(int32_t n_sides, const int32_t *gid, const int32_t *side, const int32_t *neighbor,
                                   const int64_t *off, const int64_t *len, int64_t n_iface, int32_t n_global,
                                   const int32_t *owner, int32_t rank, int32_t world, int64_t *partner,
                                   int64_t *send_index, int64_t send_cap, int64_t *n_send, int32_t *peer_rank,
                                   int64_t *peer_send_off, int64_t *peer_recv_off, int64_t *peer_count,
                                   int32_t peer_cap, int32_t *n_peers) {
  if (n_sides < 0 || !gid || !side || !neighbor || !off || !len || !partner || !n_send || !n_peers) return MMMFE_E_INVALID;
  std::vector<mmmfe::ExchangeSide> sides(static_cast<size_t>(n_sides));
  for (int32_t i = 0; i < n_sides; ++i) sides[i] = {gid[i], side[i], neighbor[i], off[i], len[i]};
  std::vector<int32_t> own;
  if (owner && n_global > 0) own.assign(owner, owner + n_global);
  mmmfe::ExchangePlan P;
  if (mmmfe::build_exchange_plan(sides, n_iface, own, rank, world, P) != nullptr) return MMMFE_E_INVALID;
  if (int64_t(P.send_index.size()) > send_cap || int32_t(P.peers.size()) > peer_cap) return MMMFE_E_INVALID;
  for (int64_t i = 0; i < n_iface; ++i) partner[i] = P.partner[size_t(i)];
  for (size_t i = 0; i < P.send_index.size(); ++i) send_index[i] = P.send_index[i];
  *n_send = int64_t(P.send_index.size());
  for (size_t i = 0; i < P.peers.size(); ++i) {
    peer_rank[i] = P.peers[i].rank; peer_send_off[i] = P.peers[i].send_off;
    peer_recv_off[i] = P.peers[i].recv_off; peer_count[i] = P.peers[i].count;
  }
  *n_peers = int32_t(P.peers.size());
  return MMMFE_OK;
}

extern "C" int mmmfe_assign_owners(int32_t n_global, const double *cost, int32_t world, int32_t *owner) {
  if (n_global <= 0 || !cost || !owner || world < 1) return MMMFE_E_INVALID;
  std::vector<int32_t> o;
  mmmfe::assign_owners(std::vector<double>(cost, cost + n_global), world, o);
  for (int32_t i = 0; i < n_global; ++i) owner[i] = o[size_t(i)];
  return MMMFE_OK;
}
