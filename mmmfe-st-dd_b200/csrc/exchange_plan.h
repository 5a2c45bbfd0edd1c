// Host-side plan of the mortar-face exchange (src/darcy_vtdd.cc:1379-1411 without MPI): which interface entries of
// this rank are combined locally, which are packed and sent to which peer, and where the received ones land.
// Pure C++ (no device calls): the plan is checked by world_size-2 CPU tests (tests/test_multirank_cpu.py).
#pragma once
#include <cstdint>
#include <vector>

namespace mmmfe {

struct ExchangeSide {   // one interface side of a local subdomain
  int32_t gid, side, neighbor;  // global subdomain id, side 0 bottom / 1 right / 2 top / 3 left, neighbour's global id
  int64_t off, len;             // position of the side's mortar coefficients in the rank's interface vector
};

struct ExchangePeer {
  int32_t rank;
  int64_t send_off, recv_off, count;  // slices of the packed send / receive buffers (doubles)
};

struct ExchangePlan {
  // per interface entry: >= 0 local partner entry; <= -2: entry -(v + 2) of the receive buffer
  std::vector<int64_t> partner;
  std::vector<int64_t> send_index;    // interface entries in send-buffer order
  std::vector<ExchangePeer> peers;    // ascending rank
};

// owner: rank of every global subdomain (empty or world == 1: everything is local).  Both sides of a remote face
// order their messages by (global id, side) of the SENDING subdomain, so sender and receiver agree without
// negotiation.  Returns nullptr or an error message.
const char *build_exchange_plan(const std::vector<ExchangeSide> &sides, int64_t n_iface,
                                const std::vector<int32_t> &owner, int32_t rank, int32_t world, ExchangePlan &P);

// Ownership rule of the front-end: contiguous blocks of subdomains (rank order of the reference, README.md:64)
// balanced by nx * ny * nt.
void assign_owners(const std::vector<double> &cost, int32_t world, std::vector<int32_t> &owner);

}  // namespace mmmfe
