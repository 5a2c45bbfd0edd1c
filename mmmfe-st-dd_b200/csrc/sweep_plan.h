// Host-side planning of the persistent sweep kernel (sweep.cuh): the factor layout in streaming order, the static
// per-CTA entry lists, and the shared-memory ring schedule.  Pure C++ (no device calls), so the plans can be checked
// on a machine without a GPU (mmmfe_sweep_plan_check).
#pragma once
#include <cstdint>
#include <vector>

#include "nd_tree.h"
#include "sweep.cuh"

namespace mmmfe {

struct SweepNodeH {       // top front, or the root of a subtree (pseudo node: s = 0, b = n_rootb)
  int32_t old_id = -1;    // NdTree node
  int32_t s = 0, b = 0;
  int32_t child[2] = {-1, -1}, parent = -1;  // sweep node ids
  int32_t n_fwd = 0, n_bwd = 0;              // completions per time step
  int32_t sep_type = 0, height = 0;
  int32_t idx_off = 0, src_off = 0, zs_off = -1, zc_off = -1;
};

struct SweepItem {        // one column block of a top front, streamed as one or more row chunks
  int32_t kind, node, col0, cw, rlen;
  int64_t src;            // offset of the tile (rlen x cw doubles) in the sweep-layout factor
};

struct SweepLayout {
  int32_t n_cta = 0, chunk_dbl = 2048;
  int32_t n_top = 0;
  std::vector<SweepNodeH> nodes;                 // top fronts (any order), then one pseudo node per subtree instance
  std::vector<std::vector<SweepItem>> fwd, bwd;  // top items by height (ascending)
  std::vector<RepackDesc> repack;                // builds the top part of the sweep layout from the front storage
  int64_t rs_total = 0;                          // doubles of the sweep-layout factor
  int64_t sub_rs_base = 0, sub_old_base = 0, sub_dbl = 0;  // subtree blobs are copied verbatim (one contiguous range)
  std::vector<int32_t> top_cell;                 // 2 canonical cells per front variable (separator entries only)
  std::vector<uint16_t> tmpl16;                  // packed templates
  std::vector<int64_t> tmpl_off;
  std::vector<int32_t> inst_xi;                  // per instance: offset of its interior variables in xi
  int64_t z_total = 0, xi_total = 0;
  int32_t f_max = 0;                             // longest top-level gather vector
  int32_t tmpl16_max = 0;                        // uint16 words of the largest template
  int32_t sub_ws_max = 0;                        // vector workspace entries (per right-hand side) of a subtree entry
  int32_t blob_max = 0;                          // doubles of the largest single entry
  double bytes_per_step = 0;                     // factor bytes one step streams
};

struct SweepSchedule {
  std::vector<SweepEntry> entries;
  std::vector<SweepIssue> issue;
  std::vector<int32_t> cta_begin, cta_prologue;
  int32_t tmpl_dbl = 0, ws_dbl = 0, ring_dbl = 0;
  size_t smem_bytes = 0;
};

// Fails (returns a message) when the tree cannot be run by the sweep kernel: a leaf outside a subtree, boxes wider
// than 127 cells, template fields beyond 16 bits.
const char *build_sweep_layout(const NdTree &tr, int32_t n_cta, int32_t chunk_dbl, SweepLayout &L);

// Shared-memory partition for M right-hand sides within smem_budget bytes per CTA, entry lists and ring schedule.
// Returns nullptr on success; a message when the ring cannot hold max(2, min_fill) of the largest entries.
const char *build_sweep_schedule(const NdTree &tr, const SweepLayout &L, int32_t M, size_t smem_budget,
                                 double min_fill, SweepSchedule &S);

// CPU model of the kernel's control flow: replays n_steps of every CTA's entry list against the completion counters
// and the ring / barrier rules.  Returns nullptr when every entry completes and no ring range is overwritten
// before its consumer finished; else a description of the first violation.
const char *check_sweep_schedule(const SweepLayout &L, const SweepSchedule &S, int32_t n_steps);

}  // namespace mmmfe
