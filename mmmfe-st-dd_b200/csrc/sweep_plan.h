// Host-side planning of the persistent sweep kernel (sweep.cuh): the factor layout in streaming order, the static
// index lists, the per-CTA entry lists, and the shared-memory ring schedule.  Pure C++ (no device calls), so the
// plans can be checked on a machine without a GPU (mmmfe_sweep_plan_check).
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
  int32_t zs_off = -1, zc_off = -1;
};

struct SweepWaveH {       // one wave: the work of ONE CTA at one tree level and direction (sweep.cuh)
  int32_t kind = 0, cta = 0, height = 0, flags = 0;  // flags: SWF_ROWS_WS, SWF_REUSE
  int32_t n_rows = 0, n_mt = 0, n_cols = 0, n_parts = 0;
  int64_t idx_off = 0;    // segment 0 in the index array (int32 units, multiple of 4)
  int32_t idx_len = 0;    // int32 units, multiple of 4
  int32_t off_rows = 0, off_zs = 0, off_colA = 0, off_colB = 0, off_mt = 0;  // int32 units from idx_off
  int64_t sig_off = 0;    // completion counters of the wave (index array, int32 units)
  int32_t n_sig = 0;
  int64_t tile_src = 0;   // micro-tiles in the sweep-layout factor (doubles)
  int32_t tile_dbl = 0;
  std::vector<int32_t> fronts;  // sweep nodes with a piece in the wave
  std::vector<int32_t> deps;    // packed dependencies (counter | count << SWEEP_DEP_SHIFT)
};

struct SweepPackH {       // up to SweepLayout::pack subtrees of one template that one CTA walks in lockstep
  int32_t pos0 = 0, n = 0, tmpl = 0, cta = 0;
  int32_t cls = -1, shared = 0;  // shared: all members have bitwise identical blobs, stored and streamed once
  int64_t sig_off = 0;    // forward completion counters of the members (index array, int32 units)
};

struct SweepInstH {       // one subtree instance
  int32_t pos = 0;                 // position in allocation order: (CTA, template, instance)
  int32_t xi_off = 0, pb_off = 0;  // entries (per right-hand side) into xi / box-major ptil; both even
  int64_t xb = 0;                  // root-boundary user dofs (offset into the index array, int32 units)
  int64_t rf = 0, rb = 0;          // blobs in the sweep-layout factor
};

struct SweepLayout {
  int32_t n_cta = 0, chunk_dbl = 2048, ct = 256;  // CTAs, micro-tile doubles per wave, compute threads per GROUP
  int32_t n_top = 0;
  std::vector<SweepNodeH> nodes;                 // top fronts (any order), then one pseudo node per subtree instance
  std::vector<SweepInstH> inst;
  std::vector<int32_t> cta_of_inst;              // CTA that runs each subtree
  std::vector<int32_t> inst_at_pos;              // allocation order -> instance
  std::vector<int32_t> inst_ij;                  // (I0, J0) of the instance at every position
  std::vector<SweepPackH> packs;                 // CTA-major
  int32_t pack = 1;                              // subtrees per pack (1, 2 or 4)
  std::vector<int32_t> prod_cta;                 // per completion counter: the CTA when a single one bumps it, else -2
  int32_t ng = 1, mt_dbl = 512;                  // compute groups per CTA; target doubles of a micro-tile
  bool direct = true;                            // strips that fit one micro-tile are written out by their warp
  int32_t ws_rows = 4096;                        // waves with at most this many rows gather them into the workspace
  int32_t wave_ws_rows = 0;                      // most rows any wave keeps in the workspace
  std::vector<std::vector<SweepWaveH>> fwd, bwd;  // waves by height (ascending), CTA-major within a height
  std::vector<RepackDesc> repack;                // micro-tiles of the sweep layout from the front storage
  std::vector<RepackSub> repack_sub;             // subtree blobs
  std::vector<int32_t> perm;                     // backward-blob permutations of the templates (-1: zero padding)
  std::vector<int32_t> rb2_size;                 // doubles of the backward blob of each template in sweep layout
  int64_t rs_total = 0;                          // doubles of the sweep-layout factor
  std::vector<int32_t> index;                    // static index lists
  std::vector<int32_t> cell_box;                 // canonical cell -> box-major position
  std::vector<uint16_t> tmpl16;                  // packed templates
  std::vector<int64_t> tmpl_off;
  int64_t z_total = 0, xi_total = 0, pb_total = 0;
  int32_t tmpl16_max = 0;                        // uint16 words of the largest template
  int32_t sub_ws_max = 0;                        // vector workspace entries (per right-hand side) of a subtree entry
  double bytes_per_step = 0;                     // factor bytes one step streams (with block padding)
  double logical_bytes_per_step = 0;             // the same with every shared subtree blob counted per member
};

struct SweepSchedule {
  std::vector<SweepEntry> entries;
  std::vector<SweepIssue> issue;
  std::vector<int32_t> cta_begin, cta_prologue;
  std::vector<uint8_t> grp;      // compute group of every entry
  int32_t ng = 1;
  int32_t tmpl_dbl = 0, ws_dbl = 0, ws_sub_dbl = 0, ring_dbl = 0, entry_max = 0;  // ws_sub_dbl: per pack member
  size_t smem_bytes = 0;
};

// idx_user: user flux dof of every front variable (NdTree::idx order); dof_of_canon may be empty (canonical).
// Fails (returns a message) when the tree cannot be run by the sweep kernel: a leaf or a cell outside every subtree,
// boxes wider than 127 cells, template fields beyond 16 bits.
const char *build_sweep_layout(const NdTree &tr, const std::vector<int32_t> &idx_user,
                               const std::vector<int32_t> &dof_of_canon, int32_t n_cta, int32_t chunk_dbl, int32_t ct,
                               SweepLayout &L, int32_t ng = 1, int32_t mt_dbl = 512, int32_t pack = 1,
                               const std::vector<int32_t> &inst_class = std::vector<int32_t>());

// Shared-memory partition for M right-hand sides within smem_budget bytes per CTA, entry lists and ring schedule.
// Returns nullptr on success; a message when the ring cannot hold max(2, min_fill) of the largest entries.
const char *build_sweep_schedule(const NdTree &tr, const SweepLayout &L, int32_t M, int32_t ng, size_t smem_budget,
                                 double min_fill, SweepSchedule &S);

// CPU model of the kernel's control flow: replays n_steps of every CTA's entry list against the completion counters
// and the ring / barrier rules.  Returns nullptr when every entry completes, no ring range is overwritten before
// its consumer finished and no vector is streamed before its producer retired; else the first violation.
const char *check_sweep_schedule(const SweepLayout &L, const SweepSchedule &S, int32_t n_steps);

}  // namespace mmmfe
