// Persistent time-sweep kernel of the interface-GMRES engine (sm_100a): ALL local time steps of one factor batch
// (forward substitution, backward substitution, right-hand-side build, pressure update and interface traces) run in
// ONE cooperative launch.
//
//  * every CTA owns a static, ordered list of ENTRIES (a subtree, or a column block of one top-level front) that is
//    identical for every time step; the factor bytes of an entry are contiguous in HBM ("sweep layout", built once
//    by launch_repack) and are streamed into a shared-memory ring by TMA bulk copies (cp.async.bulk + mbarrier)
//    that run several entries AHEAD of the arithmetic -- the loads do not depend on the right-hand side, so the
//    HBM stream never stops at a tree level;
//  * ordering between fronts is data flow: per-front completion counters in global memory (release/acquire), no
//    grid-wide barrier anywhere, also not between time steps;
//  * the RT0 structure of the saddle matrix (K_up = -+1, K_pu = +-dt, diagonal M) is exploited to fuse
//    assemble_rhs_star (src/darcy_vtdd.cc:769-855), the pressure recovery and the trace shuttles
//    (src/darcy_vtdd.cc:1006-1033) into the subtree entries: fluxes of subtree-interior edges never touch HBM.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace mmmfe {

constexpr int SWEEP_THREADS = 256;
constexpr int SWEEP_NBAR = 8;       // mbarriers = most entries in flight per CTA
constexpr int SWEEP_TMPL_HDR = 32;  // int32 header words of a packed subtree template

enum SweepKind : int32_t { SW_SUB_FWD = 0, SW_TOP_FWD = 1, SW_TOP_BWD = 2, SW_SUB_BWD = 3 };
enum SweepFlag : int32_t { SWF_FIRST = 1, SWF_LAST = 2, SWF_NEW_RUN = 4, SWF_WRITE_ZS = 8 };

// One unit of work of one CTA (128 bytes, staged through shared memory one entry ahead).
//   top entries  (column block [col0, col0 + cw) of a front, tile rows [r_beg, r_end)):
//     a0 col0, a1 cw, a2 r_beg, a3 r_end, a4 s, a5 b, a6 separator type (0 x-edges, 1 y-edges)
//     o0 idx_off, o1 src_off, o2 zs_off, o3 zc_off, o4 / o5 zc_off of child 0 / 1 (-1: none)
//     FWD tile = -G^T (s x b), reduces over the s separator rows;  BWD tile = [F11^-1 ; -G] ((s+b) x s)
//   subtree entries: a0 template, a1 boundary-entry table offset (-1: box touches no domain side), a2 I0, a3 J0,
//     o0 zc_off (root boundary contributions), o1 xi_off (interior right-hand sides kept between the passes)
struct alignas(16) SweepEntry {
  int32_t kind, flags, n_dbl, ring_off;
  int32_t issue_upto, sig, dep0, dep0_n;  // sig: counter bumped on completion; dep: counter >= (step+1) * dep_n
  int32_t dep1, dep1_n, a0, a1;
  int32_t a2, a3, a4, a5;
  int32_t a6, a7, o0, o1;
  int32_t o2, o3, o4, o5;
  int64_t src, t_src;              // factor offsets (doubles) of this entry / of the first entry it releases
  int32_t t_n_dbl, t_ring_off, pad0, pad1;
};
static_assert(sizeof(SweepEntry) == 128, "SweepEntry must be 128 bytes");

struct SweepIssue {  // what the TMA issuer needs of an entry
  int64_t src;
  int32_t n_dbl, ring_off;
};

struct SweepArgs {
  const SweepEntry *entries;
  const SweepIssue *issue;
  const int32_t *cta_begin;     // n_cta + 1
  const int32_t *cta_prologue;  // entries issued before the first one is processed
  const double *Rs;             // factor in sweep layout
  const int32_t *idx_user;      // front variables -> user flux dof (separator first)
  const int32_t *src;           // child gather maps
  const int2 *top_cell;         // front variables -> the two canonical cells of the edge
  const uint16_t *tmpl16;       // packed subtree templates
  const int64_t *tmpl_off;      // per template, in uint16 units
  const int32_t *dof_of_canon;  // nullptr when the user numbering is canonical
  double *x;                    // [n_flux][M]   solution of top-level separator dofs
  double *z;                    // per-front vectors
  double *xi;                   // subtree-interior right-hand sides after the forward pass
  double *ptil;                 // [cell][M] canonical cell order
  uint32_t *cnt;                // completion counters (2 per sweep node: forward, backward)
  volatile int *abort_flag;     // pinned host memory: set when a spin-wait gave up
  long long *prof;              // optional [n_cta][8] cycle counters: per entry kind 0..3, 4 counter waits, 5 TMA waits
  const double *minv;           // 1 / (c0 |cell|), canonical cell order
  // boundary entries of the batch (interface dofs; in the bar problem also outer Dirichlet dofs)
  const int32_t *sub_q;            // per boundary subtree: [n_int][M] -> entry or -1
  const double *const *bq_data;    // entry -> right-hand-side data of level 0 (nullptr: none)
  double *const *bq_trace;         // entry -> trace slot of level 0 (nullptr: none)
  const int32_t *bq_stride;        // entry -> stride per time level (data and trace)
  const double *bq_coef;           // rhs += coef * data[level * stride]
  // optional per-member arrays (bar / final sweeps), M pointers each, user numbering
  const double *const *src_next;  // ptil_next += minv * src[(level+1) * n_pressure + cell]
  double *const *hist;            // solution history [level][n_flux + n_pressure]
  const int32_t *p_of_cell;       // canonical cell -> user pressure index (nullptr: identity)
  int32_t hist_add, has_src;
  int64_t hist_stride;
  double kx0, kx1, ky0, ky1;    // K_up entries of an x-edge (left, right cell) and a y-edge (bottom, top cell)
  double cl, cr, cb, ct;        // K_pu entries of a cell (left, right, bottom, top edge)
  int32_t nx, ny;
  int64_t nxe, n_flux, n_pressure;
  int32_t n_steps, level0;
  int32_t tmpl_dbl, ws_dbl, ring_dbl;  // shared-memory partition in doubles
  int32_t evict_first;
};

// One block of the sweep layout: dst[r * cw + c] = value(r, col0 + c) (0 beyond ncols), r < rlen.
//   mode 0: value(r, c) = src[r * ld + c]   (forward blocks from -G^T, s rows of ldf)
//   mode 1: value(r, c) = src[c * ld + r]   (backward blocks: transpose of R = [F11^-1 | -G^T], s rows of ldr)
struct RepackDesc {
  int64_t dst, src;
  int32_t ld, rlen, col0, cw, ncols, mode;
};
void launch_repack(const RepackDesc *d_desc, int32_t n_desc, const double *R_old, double *Rs, cudaStream_t st);

size_t sweep_smem_optin();
// resident CTAs per SM of the sweep kernel for this much dynamic shared memory
int sweep_ctas_per_sm(int M, size_t smem_bytes);
// cooperative launch; returns the CUDA error code
cudaError_t launch_sweep(int M, const SweepArgs &a, int32_t n_cta, size_t smem_bytes, cudaStream_t st);

}  // namespace mmmfe
