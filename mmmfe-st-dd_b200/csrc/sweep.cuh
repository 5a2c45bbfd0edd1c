// Persistent time-sweep kernel of the interface-GMRES engine (sm_100a): ALL local time steps of one factor batch
// (forward substitution, backward substitution, right-hand-side build, pressure update and interface traces) run in
// ONE cooperative launch.
//
//  * every CTA owns a static, ordered list of ENTRIES (a subtree, or a column block of one top-level front) that is
//    identical for every time step.  Everything an entry reads that is known ahead of time -- its factor bytes
//    ("sweep layout", contiguous per entry, built once by the repack kernels), its index lists, and the vectors this
//    CTA itself produced earlier (subtree right-hand sides, the pressures of the box) -- is streamed into a
//    shared-memory ring by TMA bulk copies (cp.async.bulk + mbarrier) that run several entries AHEAD of the
//    arithmetic: the HBM stream never stops at a tree level;
//  * warp-specialised: 8 compute warps, one RETIRE warp (publishes completions with release semantics, frees ring
//    space, issues the next TMA copies) and one DEPENDENCY warp (stages entry descriptors, waits on the completion
//    counters of other CTAs ahead of the compute warps).  Hand-offs are shared-memory mbarriers; the compute warps
//    only ever synchronise among themselves (named barrier);
//  * ordering between fronts is data flow: per-front completion counters in global memory (release/acquire), no
//    grid-wide barrier anywhere, also not between time steps;
//  * the RT0 structure of the saddle matrix (K_up = -+1, K_pu = +-dt, diagonal M) is exploited to fuse
//    assemble_rhs_star (src/darcy_vtdd.cc:769-855), the pressure recovery and the trace shuttles
//    (src/darcy_vtdd.cc:1006-1033) into the subtree entries: fluxes of subtree-interior edges never touch HBM.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace mmmfe {

// compute threads per CTA: 256, 128 or 64 (template parameter of the kernel); + retire warp + dependency warp
constexpr int SWEEP_NBAR = 16;                     // slots = most entries in flight per CTA (one service lane each)
constexpr int SWEEP_TMPL_HDR = 48;                 // int32 header words of a packed subtree template
constexpr int SWEEP_SEGS = 4;                      // TMA copies per entry
constexpr int SWEEP_GTAB_DBL = 128;                // shared-memory doubles of the per-entry group table (+ counter)
constexpr int SWEEP_MAX_ENTRIES = 8 * SWEEP_GTAB_DBL - 8;  // entries per CTA and step the group table can hold

enum SweepKind : int32_t { SW_SUB_FWD = 0, SW_TOP_FWD = 1, SW_TOP_BWD = 2, SW_SUB_BWD = 3 };
// SWF_ROWS_WS: the gathered rows of a wave live in the group's workspace (not in the ring scratch); SWF_REUSE: the
// previous wave of this CTA gathered exactly these rows and left them there
// SWF_SHARED: the members of a subtree pack have bitwise identical factor blobs; segment 0 holds one copy
// SWF_DIRECT: every micro-tile of the wave holds ALL rows of its columns: the warps write the outputs themselves, the
// wave has no partial sums, no second barrier and no output phase
enum SweepFlag : int32_t { SWF_FIRST = 1, SWF_LAST = 2, SWF_ROWS_WS = 4, SWF_REUSE = 8, SWF_SHARED = 16, SWF_DIRECT = 32 };
enum SweepSrc : int32_t { SRC_FACTOR = 0, SRC_INDEX = 1, SRC_XI = 2, SRC_PTIL = 3 };

// One unit of work of one CTA (128 bytes).  The ring region of an entry holds its streamed segments back to back
// (n[0..3] doubles each; a segment that the TMA issuer does not copy -- SweepIssue::n = 0 -- is scratch space).
//
//   WAVES (SW_TOP_FWD / SW_TOP_BWD): everything ONE CTA does for one tree level and direction within a shared-memory
//   budget: pieces (column ranges, all rows) of one or more top fronts, cut into MICRO-TILES of nrows x ncols doubles
//   (ncols = 2..64, a power of two; stored row-major and contiguous) that single warps multiply independently.
//     segment 0 (index array): row records, column records, micro-tile descriptors
//     segment 1 (factor):      the micro-tiles
//     segment 2 (scratch):     gathered vector rows [n_rows][M] (unless SWF_ROWS_WS), partial sums
//     a0 n_rows, a1 n_mt, a2 n_cols, a3 partial slots, a4 rows, a5 micro-tiles, a6 column records A, o3 column
//     records B (forward), o0 zs destinations (forward)   (offsets in int32 units from the start of segment 0),
//     a7 tree height
//     phases: gather (rows and pass-through from L2) | micro-tiles (one warp each, partial sums to shared memory) |
//     reduction of the row parts in a fixed order + output.
//       FWD  row record int4 {cell A | sep-type << 30, cell B, child-0 z, child-1 z} (+ zs destination or -1):
//            v = -(K_up p_old) + children (assemble_rhs_star fused); column record A {partial slot, parts | stride
//            << 16, z destination, child-0 z}, B {child-1 z};  micro-tile of -G^T (rows: separator, columns: boundary)
//       BWD  row record int {x dof, or 1 << 31 | z index for the separator rows}; column record A {slot, parts |
//            stride << 16, x dof, -}; micro-tile of [F11^-1 ; -G] (rows: separator + boundary, columns: separator)
//   subtree entries (PACKS of a2 <= SweepArgs::pack subtrees of one template, neighbours in every array, walked in
//     lockstep by GT / pack threads each): a0 template, a2 members, a7 allocation position of the first member,
//     o0 zc_off (root boundary contributions), o1 xi_off, o2 ptil offset of the box (box-major cell order) of the
//     first member; segments (members back to back): 0 factor blobs, 1 root-boundary x dofs (BWD), 2 interior
//     right-hand sides xi (BWD), 3 box pressures
//
// Completion: sig >= 0: one counter (subtrees); n_sig > 0: the counters index[sig_off ...] (waves: one per front).
// Dependencies: dep[t] = counter | count << 21; satisfied when counter >= (step + 1) * count.
constexpr int SWEEP_MAX_DEPS = 8;
constexpr int SWEEP_DEP_SHIFT = 21;
struct alignas(16) SweepEntry {
  int32_t kind, flags, ring_off, issue_upto;
  int32_t sig, n_sig, sig_off, n_dep;
  int32_t dep[SWEEP_MAX_DEPS];
  int32_t a0, a1, a2, a3;
  int32_t a4, a5, a6, a7;
  int32_t o0, o1, o2, o3;
  int32_t n[SWEEP_SEGS];
};
static_assert(sizeof(SweepEntry) == 128, "SweepEntry must be 128 bytes");

struct alignas(16) SweepMt {  // micro-tile descriptor (32 bytes)
  int32_t tile_off;           // doubles from the start of segment 1
  int32_t row0, nrows;        // rows [row0, row0 + nrows) of the wave's gathered vector
  int32_t ncols;              // 2, 4, ..., 64
  int32_t part_off;           // first partial slot
  int32_t col0;               // column record of the tile's first column
  int32_t direct;             // 1: the tile holds all rows of its columns -> the warp writes their outputs itself
  int32_t pad;
};

struct alignas(16) SweepIssue {  // what the TMA issuer needs of an entry
  int64_t off[SWEEP_SEGS];       // in doubles, relative to the base of the source
  int32_t n[SWEEP_SEGS];         // doubles
  int32_t kind[SWEEP_SEGS];      // SweepSrc
  int32_t ring_off, pad[3];
};
static_assert(sizeof(SweepIssue) == 80, "SweepIssue must be 80 bytes");

struct SweepArgs {
  const SweepEntry *entries;
  const SweepIssue *issue;
  const int32_t *cta_begin;     // n_cta + 1
  const int32_t *cta_prologue;  // entries issued before the first one is processed
  const uint8_t *grp;           // compute group of every entry
  int32_t ng;                   // compute groups per CTA (1, 2 or 4)
  const double *Rs;             // factor in sweep layout
  const int32_t *index;         // static index lists (gather / column lists, root-boundary dofs)
  const uint16_t *tmpl16;       // packed subtree templates
  const int64_t *tmpl_off;      // per template, in uint16 units
  const int32_t *dof_of_canon;  // nullptr when the user numbering is canonical
  double *x;                    // [n_flux][M]   solution of top-level separator dofs
  double *z;                    // per-front vectors
  double *xi;                   // subtree-interior right-hand sides after the forward pass
  double *ptil;                 // [cell][M], box-major cell order
  uint32_t *cnt;                // completion counters (2 per sweep node: forward, backward)
  volatile int *abort_flag;     // device memory: set when a spin-wait gave up
  long long *prof;              // optional [n_cta][24] cycle counters
  long long *trace;             // optional [entries][3] global-timer stamps of step trace_step (start, opened, end)
  int32_t trace_step;
  const double *minv;           // 1 / (c0 |cell|), box-major cell order; nullptr: minv_const everywhere
  double minv_const;
  // boundary entries of the batch (interface dofs; in the bar problem also outer Dirichlet dofs)
  const int32_t *sub_q;            // per boundary subtree: [n_int][M] -> entry or -1
  const double *const *bq_data;    // entry -> right-hand-side data of level 0 (nullptr: none)
  double *const *bq_trace;         // entry -> trace slot of level 0 (nullptr: none)
  const int32_t *bq_stride;        // entry -> stride per time level (data and trace)
  const double *bq_coef;           // rhs += coef * data[level * stride]
  // optional per-member arrays (bar / final sweeps), M pointers each, user numbering
  const double *const *src_next;  // ptil_next += minv * src[(level+1) * n_pressure + p]
  double *const *hist;            // solution history [level][n_flux + n_pressure]
  const int32_t *p_of_cell;       // canonical cell -> user pressure index (nullptr: identity)
  int32_t hist_add, has_src;
  int64_t hist_stride;
  double kx0, kx1, ky0, ky1;    // K_up entries of an x-edge (left, right cell) and a y-edge (bottom, top cell)
  double cl, cr, cb, ct;        // K_pu entries of a cell (left, right, bottom, top edge)
  int32_t nx, ny;
  int64_t nxe, n_flux, n_pressure;
  int32_t n_steps, level0;
  int32_t tmpl_dbl, ws_dbl, ring_dbl;  // shared-memory partition in doubles (template and workspace: per group)
  int32_t pack, ws_sub_dbl;            // subtrees per pack (1, 2, 4) and workspace doubles of one pack member
  const int32_t *inst_bq;              // per subtree (allocation order): boundary-entry table offset or -1
  const int32_t *inst_ij;              // per subtree (allocation order): I0, J0
  int32_t evict_first;
  int32_t pub_fence;            // entries with several completion counters: one release fence + relaxed increments
  uint32_t spin_ns;             // back-off of the service warps' polling loops
  int32_t poll_window;          // entries ahead of the oldest unopened one that poll the completion counters
};

// One block of the sweep layout: dst[r * cw + c] = value(r, col0 + c) (0 beyond ncols), r < rlen.
//   mode 0: value(r, c) = src[r * ld + c]   (forward blocks from -G^T, s rows of ldf)
//   mode 1: value(r, c) = src[c * ld + r]   (backward blocks: transpose of R = [F11^-1 | -G^T], s rows of ldr)
struct RepackDesc {
  int64_t dst, src;
  int32_t ld, rlen, col0, cw, ncols, mode;
};
void launch_repack(const RepackDesc *d_desc, int32_t n_desc, const double *R_old, double *Rs, cudaStream_t st);

// Subtree blobs: forward blob copied, backward blob permuted (perm[k] = old position of new element k, per template)
struct RepackSub {
  int64_t old_rf, old_rb, new_rf, new_rb;
  int32_t rf_size, rb_size, perm_off, pad;
};
void launch_repack_sub(const RepackSub *d_desc, int32_t n_desc, const int32_t *perm, const double *R_old, double *Rs,
                       cudaStream_t st);

struct BlobPair { int64_t a, b, n; };  // ranges [a, a + n) and [b, b + n) of the factor (doubles)
// counts the pairs that differ in any bit into *d_mismatch
void launch_blob_compare(const BlobPair *d_pairs, int32_t n_pairs, const double *R, int *d_mismatch, cudaStream_t st);

size_t sweep_smem_optin();
// resident CTAs per SM of the sweep kernel (ct compute threads) for this much dynamic shared memory
int sweep_ctas_per_sm(int M, int ct, size_t smem_bytes);
// fixed shared-memory doubles of the kernel besides template, workspace and ring
int sweep_fixed_dbl(int M, int ct);
// cooperative launch; returns the CUDA error code
cudaError_t launch_sweep(int M, int ct, const SweepArgs &a, int32_t n_cta, size_t smem_bytes, cudaStream_t st);

}  // namespace mmmfe
