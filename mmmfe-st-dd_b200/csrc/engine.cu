// Host driver of the B200 interface-GMRES engine: C ABI (include/mmmfe_b200.h), factor grouping, the device
// multifrontal factorisation, time-step sweeps as CUDA graphs on per-group streams, interface operator, GMRES.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/mmmfe_b200.h"
#include "common.h"
#include "exchange_plan.h"
#include "kernels.cuh"
#include "nd_tree.h"
#include "sweep.cuh"
#include "sweep_plan.h"
#include "wide.cuh"

#ifdef MMMFE_WITH_NCCL
#include <dlfcn.h>
#include <nccl.h>
#endif

namespace mmmfe {

#ifdef MMMFE_WITH_NCCL
// NCCL is bound at run time (dlopen), never at link time: inside a Python process the copy torch has already
// loaded is reused (same SONAME), a stand-alone DarcyVT picks up the system library.
struct NcclApi {
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  bool ok = false;
};

static NcclApi &nccl() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW);
  if (!h) return api;
#define MMMFE_NCCL_SYM(name) api.name = reinterpret_cast<decltype(api.name)>(dlsym(h, "nccl" #name))
  MMMFE_NCCL_SYM(GetUniqueId); MMMFE_NCCL_SYM(CommInitRank); MMMFE_NCCL_SYM(CommDestroy); MMMFE_NCCL_SYM(Send);
  MMMFE_NCCL_SYM(Recv); MMMFE_NCCL_SYM(GroupStart); MMMFE_NCCL_SYM(GroupEnd); MMMFE_NCCL_SYM(AllReduce);
#undef MMMFE_NCCL_SYM
  api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.Send && api.Recv && api.GroupStart &&
           api.GroupEnd && api.AllReduce;
  return api;
}
#endif

#ifdef MMMFE_WITH_NCCL
static std::mutex &comm_cache_mutex() { static std::mutex m; return m; }
static std::map<std::string, ncclComm_t> &comm_cache() { static std::map<std::string, ncclComm_t> m; return m; }
#endif

static const double kNormalSign[4] = {-1.0, 1.0, 1.0, -1.0};  // inc/utilities.h:79-91
static const int kOpposite[4] = {2, 3, 0, 1};

// ---- one registered subdomain ------------------------------------------------------------------------------------
struct Sub {
  int32_t gid, nx, ny, nt, nf, np;
  double dt;
  HostCsr matrix;
  std::vector<int32_t> canon;  // empty = canonical
  int32_t neighbor[4], n_side[4], mortar_len[4];
  std::vector<int32_t> side_dofs[4];
  HostCsr f2c[4], c2f[4];
  int64_t iface_off[4] = {-1, -1, -1, -1};  // into the local interface vector
  int64_t fine_off[4] = {-1, -1, -1, -1};   // into the fine trace space (nt * n_side each)
  int32_t batch = -1, member = -1;
  // bar data on the device
  DevBuf<double> p0, src, fr_vals, hist;
  DevBuf<int32_t> fr_dofs;
  std::vector<int32_t> fr_dofs_h;
  int32_t n_fr = 0;
  bool has_bar_data = false;
};

// ---- multiscale flux basis of one factor group (src/darcy_vtdd.cc:959-1003, inc/darcy_vtdd.h:124, 264) ---------------
// "slot" = one of the sides that is an interface on at least one member of the group.  Rows and columns of the operator
// share one numbering: (slot, a) -> base[slot] + a, a = mortar dof within ONE time slab of that side.
// T[d][row][col], d = delay in mortar time slabs: response on row's side, d slabs later, to the basis function col.
struct Basis {
  int32_t m_t = 0, lvl_per_slab = 0, nt = 0;
  int32_t n_slots = 0, slot_of_side[4] = {-1, -1, -1, -1}, side_of_slot[4] = {0, 0, 0, 0};
  int32_t P[4] = {0, 0, 0, 0}, n_side[4] = {0, 0, 0, 0}, base[4] = {0, 0, 0, 0}, tr_base[4] = {0, 0, 0, 0};
  int32_t rep[4] = {-1, -1, -1, -1};  // local subdomain whose tables represent the slot
  int32_t P_tot = 0, n_tr = 0, MW = 0, n_pass = 0, MC = 8;
  int32_t col_lo = 0, col_hi = 0;  // the columns THIS rank solves for (all of them unless the ranks share the group)
  int32_t shared_by = 1;           // ranks that share the work of this group
  int32_t d_used = 0;              // delays d >= d_used are identically zero (the response has died out)
  int64_t steps = 0;               // time steps actually solved, summed over the passes
  int64_t ldt = 0;
  DevBuf<double> T;
  double seconds = 0.0, flops = 0.0, bytes = 0.0;
  long long graph_kernels = 0;
};

// ---- one factor: nested-dissection tree + block-inverse fronts on the device -------------------------------------
struct LevelWork {  // one height level of the top of the tree
  DevBuf<WorkItem> fwd, bwd;
  int32_t n_fwd = 0, n_bwd = 0, f_max = 0;
  double bytes_fwd = 0, bytes_bwd = 0;
};

struct Factor {
  int32_t nx, ny, nf, np;
  NdTree tree;
  DevBuf<FrontDesc> fronts;
  DevBuf<double> R;
  DevBuf<int32_t> idx_user, src;
  std::vector<LevelWork> levels;
  // subtrees (bottom of the tree): templates shared by shape, one instance per subtree
  DevBuf<int32_t> tmpl_pool, dof_of_canon;
  DevBuf<TemplateDev> tmpl_dev;
  DevBuf<InstanceDev> inst_dev;
  int32_t n_inst = 0;
  size_t smem_fwd[5] = {0, 0, 0, 0, 0}, smem_bwd[5] = {0, 0, 0, 0, 0};  // by M
  double bytes_sub_fwd = 0, bytes_sub_bwd = 0;
  DevCsr kup, kpu;
  DevBuf<double> minv;
  int64_t solve_bytes = 0;  // factor + index bytes one forward+backward pass streams
  double seconds = 0.0;
  // ---- persistent sweep path (sweep.cuh): RT0 structure of the saddle matrix + factor in streaming layout ----
  bool rt0 = false, sweep_ok = false;
  double kup_c[4] = {0, 0, 0, 0}, kpu_c[4] = {0, 0, 0, 0};
  std::vector<int32_t> p_of_cell_h;  // canonical cell -> user pressure index
  bool p_identity = true;
  SweepLayout lay;
  int32_t sweep_n_cta = 0, sweep_ct = 128, sweep_ng = 1;
  double sm_share = 1.0;              // fraction of the SMs the sweep kernels of this factor run on
  cudaStream_t sweep_stream = nullptr;  // concurrent mode: all sweep launches of this factor, in order
  int64_t shared_blob_pairs = 0;
  size_t sweep_budget = 0;
  DevBuf<double> Rs, minv_c;
  double minv_const = 0.0;
  bool minv_uniform = false;
  DevBuf<int32_t> sweep_index, p_of_cell, inst_ij, cell_box;
  DevBuf<uint16_t> tmpl16;
  DevBuf<int64_t> tmpl_off;
  std::unique_ptr<Basis> basis;  // multiscale flux basis of this factor group (time-Toeplitz operator)
};


// ---- one batch: up to 4 subdomains solved as simultaneous right-hand sides of one factor ---------------------------
struct Batch {
  std::shared_ptr<Factor> factor;
  std::vector<int32_t> members;  // local subdomain indices
  int32_t M = 1, nt = 0;
  DevBuf<double> x, z, ptil;
  DevBuf<int32_t> tr_dof, tr_mem, tr_stride, iface_q;
  DevBuf<int64_t> tr_base;
  DevBuf<double> tr_sign;
  int32_t n_tr = 0;
  DevBuf<const double *> p0_ptr, src_ptr;
  DevBuf<double *> hist_ptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;
  cudaGraphExec_t graph_star = nullptr;
  long long graph_kernels = 0;
  DevBuf<int32_t> level_dev;  // time level of the step graphs of the bar and final sweeps
  cudaGraphExec_t graph_step = nullptr;
  StepArgs step{};
  // persistent sweep path
  bool sweep = false;
  size_t sweep_smem = 0;
  DevBuf<SweepEntry> s_entries;
  DevBuf<SweepIssue> s_issue;
  DevBuf<int32_t> s_cta_begin, s_cta_pro, s_sub_q, s_bq_stride, s_inst_bq;
  DevBuf<uint8_t> s_grp;
  DevBuf<double> s_z, s_xi, s_ptil, s_bq_coef;
  DevBuf<uint32_t> s_cnt;
  DevBuf<const double *> s_bq_data;
  DevBuf<double *> s_bq_trace;
  SweepArgs sargs{};
  // bar sweep through the persistent kernel: boundary-entry tables that also carry the outer Dirichlet rows
  std::vector<int32_t> h_iface_q, h_bq_stride;
  std::vector<double *> h_bq_trace;
  bool bar_planned = false, bar_sweep = false;
  DevBuf<int32_t> sb_sub_q, sb_bq_stride, sb_inst_bq;
  DevBuf<double> sb_bq_coef;
  DevBuf<const double *> sb_bq_data;
  DevBuf<double *> sb_bq_trace;
};

struct Arena {
  DevBuf<double> buf;
  size_t top = 0;
  double *take(size_t n) {
    if (top + n > buf.n) fail(MMMFE_E_INVALID, "factorisation arena exhausted (%zu + %zu > %zu)", top, n, buf.n);
    double *p = buf.p + top;
    top += n;
    return p;
  }
};

}  // namespace mmmfe

using namespace mmmfe;

struct mmmfe_ctx {
  int device = 0;
  std::string err;
  bool keep_history = true, use_graphs = true, use_sweep = true, sweep_evict_first = true;
  int leaf_cells = 8, subtree_cells = 256, sweep_chunk = 7168, sweep_ct = 256, sweep_ctas_per_sm = 1, sweep_groups = 1;
  bool sweep_autotune = true;
  int sweep_pub_fence = -1;        // SweepArgs::pub_fence: 0 / 1, -1 = by the number of right-hand sides
  bool sweep_debug_stall = false;  // test hook, see plan_batch_sweep
  bool sweep_bar = true;           // bar sweep through the persistent kernel where the star sweeps use it
  bool sweep_final = true;         // final sweep through the persistent kernel where the star sweeps use it
  int tune_steps = 32;             // time levels of the set-up comparison of the two sweep implementations (0: all)
  bool sweep_crosscheck = false;  // with sweep_autotune = 0: still run both implementations once and compare their traces
  double tune_ms_sweep = 0.0, tune_ms_levels = 0.0, tune_rel_diff = -1.0;
  double sweep_min_fill = 1.75;
  int sweep_mt = 896, sweep_pack = 4;  // 7168 / 896 = one micro-tile per compute warp and wave
  bool sweep_share_blobs = true;
  // 1: the persistent kernels of different factors run CONCURRENTLY, each on its own share of the SMs (proportional
  // to nt * factor bytes of its batches); 0: one after the other on all SMs
  bool sweep_concurrent = true;
  double sweep_share_exponent = 1.0;  // SM share of a factor group ~ (streamed bytes)^exponent
  int sweep_spin_ns = 100;
  bool is_setup = false, bar_done = false, final_done = false;
  int rank = 0, world = 1;
  std::vector<int32_t> owner;
#ifdef MMMFE_WITH_NCCL
  ncclComm_t comm = nullptr;
#endif
  std::vector<std::unique_ptr<Sub>> subs;
  std::vector<std::shared_ptr<Factor>> factors;
  std::vector<std::unique_ptr<Batch>> batches;
  cudaStream_t stream = nullptr;  // interface-vector stream
  cudaEvent_t ev_start = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
  int64_t n_iface = 0, n_fine = 0;
  DevCsr c2f_all, f2c_all;  // f2c_all carries the outward sign
  DevBuf<int64_t> partner;
  DevBuf<double> lam_bar, trace, send, ap, qin, lam, r0;
  // remote exchange plan
  struct Peer {
    int rank;
    int64_t send_off, recv_off, count;
  };
  std::vector<Peer> peers;
  DevBuf<int64_t> send_index;
  DevBuf<double> send_buf, recv_buf;
  // GMRES workspace
  DevBuf<double> Q, h_dev, norm_dev, arn_scratch, ar_buf;
  DevBuf<unsigned int> arn_ticket;
  int gmres_lookahead = -1;        // iterations enqueued ahead of the host's stop test; -1: 3 on the basis operator, else 0
  int gmres_surplus = 0;           // iterations of the last solve that ran beyond the converged one
  double *gmres_ring = nullptr;    // pinned
  size_t gmres_ring_dbl = 0;
  std::vector<cudaEvent_t> gmres_events;
  int64_t q_rows = 0;
  double factor_seconds = 0.0;
  long long launch_base = 0;
  int *abort_flag = nullptr;  // device memory: raised by the sweep kernel's watchdog
  // interface operator: 0 = star sweeps every iteration (the reference's algorithm), 1 = multiscale flux basis /
  // time-Toeplitz operator precomputed at set-up, 2 = basis when the structure allows it and it has at most
  // basis_auto_columns columns per factor group, sweeps otherwise
  int operator_mode = 2, mortar_slabs = 0, basis_cols_max = 256, basis_auto_columns = 2048;
  bool basis_share = true;  // several ranks: split the columns of a group every rank holds among the ranks
  double basis_drop_tol = 1e-18;  // a pass of the basis solve stops once the pressures fell below this fraction of their peak
  bool use_basis = false;
  std::string basis_note;  // why the basis operator is not in use
  DevBuf<ToepProb> toep_probs;
  DevBuf<ToepTile> toep_tiles;
  int32_t n_toep_tiles = 0, toep_bm = 64, toep_parts = 1;
  DevBuf<double> toep_part_buf;
  double basis_seconds = 0.0, basis_flops = 0.0, basis_bytes = 0.0, toep_flops = 0.0, toep_bytes = 0.0;
  int64_t basis_columns = 0;
  // temporaries of the factorisation
  DevBuf<char> desc_arena;  // problem descriptors of the factorisation batches (desc_upload)
  size_t desc_top = 0;
  DevBuf<int32_t> d_fail;
};

namespace mmmfe {

// =================================================================================================================
// numeric factorisation
// =================================================================================================================
// Problem descriptors of the batched factorisation kernels go into a device bump arena: every batch gets its own
// range, so the host never waits for a kernel to finish before it uploads the next batch (the stream orders the
// work; cudaMemcpyAsync from pageable memory returns once the source has been staged).  The arena is rewound -- after
// a stream synchronisation -- only when it is full.
template <class T>
static T *desc_upload(mmmfe_ctx *c, const T *h, size_t count) {
  const size_t bytes = (count * sizeof(T) + 255) & ~size_t(255);
  if (c->desc_top + bytes > c->desc_arena.n) {
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->desc_top = 0;
    if (bytes > c->desc_arena.n) c->desc_arena.alloc(std::max(bytes * 2, size_t(64) << 20));
  }
  T *p = reinterpret_cast<T *>(c->desc_arena.p + c->desc_top);
  c->desc_top += bytes;
  CUDA_CHECK(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, c->stream));
  return p;
}

static void run_gemm_batch(mmmfe_ctx *c, const std::vector<GemmProb> &pr) {
  if (pr.empty()) return;
  std::vector<int32_t> ts(pr.size() + 1);
  int64_t tiles = 0;
  for (size_t i = 0; i < pr.size(); ++i) {
    ts[i] = int32_t(tiles);
    tiles += int64_t((pr[i].m + 63) / 64) * ((pr[i].n + 63) / 64);
  }
  ts[pr.size()] = int32_t(tiles);
  if (tiles > 2000000000LL) fail(MMMFE_E_INVALID, "too many GEMM tiles");
  const GemmProb *d_pr = desc_upload(c, pr.data(), pr.size());
  const int32_t *d_ts = desc_upload(c, ts.data(), ts.size());
  launch_grouped_gemm(d_pr, d_ts, int32_t(pr.size()), int32_t(tiles), c->stream);
}

static void run_trans_batch(mmmfe_ctx *c, const std::vector<TransProb> &pr) {
  if (pr.empty()) return;
  launch_transpose(desc_upload(c, pr.data(), pr.size()), int32_t(pr.size()), c->stream);
}

// In place: SPD block A (full storage) -> L^-1, the inverse of its Cholesky factor (lower triangle, exact zeros
// above).  Recursive: everything above the 32x32 base blocks is GEMM on the FP64 tensor cores.
//   A = [[A11, .], [A21, A22]] :  Li11 = rec(A11);  L21 = A21 Li11^T;  A22 -= L21 L21^T;  Li22 = rec(A22);
//   L^-1 = [[Li11, 0], [-Li22 L21 Li11, Li22]]
static void chol_inv_batch(mmmfe_ctx *c, const std::vector<InvProb> &probs, Arena &arena) {
  std::vector<InvProb> small, big;
  for (const auto &p : probs) (p.n <= 32 ? small : big).push_back(p);
  if (!small.empty())
    launch_small_inverse(desc_upload(c, small.data(), small.size()), int32_t(small.size()), c->d_fail.p, c->stream);
  if (big.empty()) return;
  const size_t mark = arena.top;
  std::vector<InvProb> a11, a22;
  std::vector<double *> T(big.size()), T2(big.size());
  for (size_t i = 0; i < big.size(); ++i) {
    const int n = big[i].n, h = n / 2, ld = big[i].ld;
    a11.push_back({big[i].A, h, ld});
    a22.push_back({big[i].A + int64_t(h) * ld + h, n - h, ld});
  }
  chol_inv_batch(c, a11, arena);
  std::vector<GemmProb> g1, g2, g3, g4;
  std::vector<TransProb> zr;
  for (size_t i = 0; i < big.size(); ++i) {
    const int n = big[i].n, h = n / 2, r = n - h, ld = big[i].ld;
    double *A = big[i].A, *A21 = A + int64_t(h) * ld, *A22 = A21 + h;
    T[i] = arena.take(size_t(r) * h);
    g1.push_back({A21, A, T[i], r, h, h, h, ld, 1, 1, ld, 1.0, 0.0});       // L21 = A21 * Li11^T
    g2.push_back({T[i], T[i], A22, r, r, h, ld, h, 1, 1, h, -1.0, 1.0});    // A22 -= L21 * L21^T
  }
  run_gemm_batch(c, g1);
  run_gemm_batch(c, g2);
  chol_inv_batch(c, a22, arena);
  for (size_t i = 0; i < big.size(); ++i) {
    const int n = big[i].n, h = n / 2, r = n - h, ld = big[i].ld;
    double *A = big[i].A, *A21 = A + int64_t(h) * ld, *A22 = A21 + h;
    T2[i] = arena.take(size_t(r) * h);
    g3.push_back({T[i], A, T2[i], r, h, h, h, h, 1, ld, 1, 1.0, 0.0});      // T2 = L21 * Li11
    g4.push_back({A22, T2[i], A21, r, h, r, ld, ld, 1, h, 1, -1.0, 0.0});   // A21 = -Li22 * T2
    zr.push_back({A21, A + h, h, r, ld, ld, 0.0});                          // upper-right block = 0
  }
  run_gemm_batch(c, g3);
  run_gemm_batch(c, g4);
  run_trans_batch(c, zr);
  arena.top = mark;
}

// S = A - K_up diag(1/M) K_pu in user flux numbering; also extracts K_up, K_pu, 1/M.
static void build_schur(const Sub &sd, HostCsr &S, HostCsr &kup, HostCsr &kpu, std::vector<double> &minv) {
  const int32_t nf = sd.nf, np = sd.np;
  const HostCsr &K = sd.matrix;
  if (K.n_rows != nf + np || K.n_cols != nf + np) fail(MMMFE_E_INVALID, "matrix must be (n_flux+n_pressure)^2");
  minv.assign(np, 0.0);
  // Two passes per block (count, prefix sum, fill), rows in parallel; the entries of a row keep the order of the
  // serial construction (first appearance), so the matrices are the same bit for bit.
  int bad = 0;
  kpu.n_rows = np; kpu.n_cols = nf; kpu.rp.assign(size_t(np) + 1, 0);
#pragma omp parallel for schedule(static) reduction(max : bad)
  for (int32_t c = 0; c < np; ++c) {
    double d = 0.0;
    int64_t cnt = 0;
    for (int64_t e = K.rp[nf + c]; e < K.rp[nf + c + 1]; ++e) {
      const int32_t col = K.col[e];
      if (col < nf) cnt += K.val[e] != 0.0;
      else if (col == nf + c) d += K.val[e];
      else if (K.val[e] != 0.0) bad = std::max(bad, 1);
    }
    if (!(d > 0.0)) bad = std::max(bad, 2);
    minv[c] = 1.0 / d;
    kpu.rp[c + 1] = cnt;
  }
  if (bad == 1) fail(MMMFE_E_UNSUPPORTED, "pressure block is not diagonal (only DGQ0 is on the path)");
  if (bad == 2) fail(MMMFE_E_UNSUPPORTED, "c0 * |cell| must be positive (c0 = 0 is not on the accelerated path)");
  for (int32_t c = 0; c < np; ++c) kpu.rp[c + 1] += kpu.rp[c];
  kpu.col.resize(size_t(kpu.rp[np])); kpu.val.resize(size_t(kpu.rp[np]));
#pragma omp parallel for schedule(static)
  for (int32_t c = 0; c < np; ++c) {
    int64_t o = kpu.rp[c];
    for (int64_t e = K.rp[nf + c]; e < K.rp[nf + c + 1]; ++e)
      if (K.col[e] < nf && K.val[e] != 0.0) { kpu.col[o] = K.col[e]; kpu.val[o] = K.val[e]; ++o; }
  }
  kup.n_rows = nf; kup.n_cols = np; kup.rp.assign(size_t(nf) + 1, 0);
  S.n_rows = S.n_cols = nf; S.rp.assign(size_t(nf) + 1, 0);
  // one row of S = A + (dt / c0) B^T M^-1 B and of K_up, into small fixed buffers
  constexpr int RW = 64;
  auto row_of = [&](int32_t e, int32_t *rc, double *rv, int &n_s, int32_t *uc, double *uv, int &n_u) {
    n_s = n_u = 0;
    auto add = [&](int32_t col, double v) {
      for (int t = 0; t < n_s; ++t) if (rc[t] == col) { rv[t] += v; return; }
      if (n_s < RW) { rc[n_s] = col; rv[n_s] = v; }
      ++n_s;
    };
    for (int64_t k = K.rp[e]; k < K.rp[e + 1]; ++k) {
      const int32_t col = K.col[k];
      if (col < nf) add(col, K.val[k]);
      else if (K.val[k] != 0.0) {
        const int32_t cell = col - nf;
        if (n_u < RW) { uc[n_u] = cell; uv[n_u] = K.val[k]; }
        ++n_u;
        const double f = -K.val[k] * minv[cell];
        for (int64_t t = kpu.rp[cell]; t < kpu.rp[cell + 1]; ++t) add(kpu.col[t], f * kpu.val[t]);
      }
    }
  };
  int wide = 0;
#pragma omp parallel for schedule(static) reduction(max : wide)
  for (int32_t e = 0; e < nf; ++e) {
    int32_t rc[RW], uc[RW];
    double rv[RW], uv[RW];
    int n_s, n_u;
    row_of(e, rc, rv, n_s, uc, uv, n_u);
    if (n_s > RW || n_u > RW) wide = 1;
    S.rp[e + 1] = n_s; kup.rp[e + 1] = n_u;
  }
  if (wide) fail(MMMFE_E_UNSUPPORTED, "a flux row of the saddle matrix couples more than 64 dofs");
  for (int32_t e = 0; e < nf; ++e) { S.rp[e + 1] += S.rp[e]; kup.rp[e + 1] += kup.rp[e]; }
  S.col.resize(size_t(S.rp[nf])); S.val.resize(size_t(S.rp[nf]));
  kup.col.resize(size_t(kup.rp[nf])); kup.val.resize(size_t(kup.rp[nf]));
#pragma omp parallel for schedule(static)
  for (int32_t e = 0; e < nf; ++e) {
    int32_t rc[RW], uc[RW];
    double rv[RW], uv[RW];
    int n_s, n_u;
    row_of(e, rc, rv, n_s, uc, uv, n_u);
    for (int t = 0; t < n_s; ++t) { S.col[S.rp[e] + t] = rc[t]; S.val[S.rp[e] + t] = rv[t]; }
    for (int t = 0; t < n_u; ++t) { kup.col[kup.rp[e] + t] = uc[t]; kup.val[kup.rp[e] + t] = uv[t]; }
  }
}

// RT0 structure of the off-diagonal blocks (src/darcy_vtdd.cc:464-466): every flux row of K_up holds the same two
// constants for its lower/upper cell, every pressure row of K_pu the same four for its left/right/bottom/top edge.
// Recovers the cell of every user pressure index from the matrix itself (x-edges, walking each row of cells from the
// left boundary).  False = some other discretisation: the generic kernels are used.
static bool detect_rt0(const Sub &sd, const HostCsr &kup, const HostCsr &kpu, const std::vector<int32_t> &dof_of_canon,
                       Factor &F) {
  const int nx = sd.nx, ny = sd.ny;
  const int64_t nxe = int64_t(nx + 1) * ny;
  auto user = [&](int64_t canon) { return dof_of_canon.empty() ? int32_t(canon) : dof_of_canon[canon]; };
  std::vector<int32_t> p_of_cell(size_t(nx) * ny, -1), cell_of_p(size_t(nx) * ny, -1);
  double k[4] = {0, 0, 0, 0};
  bool have[4] = {false, false, false, false};
  auto set_const = [&](int slot, double v) {
    if (!have[slot]) { have[slot] = true; k[slot] = v; return true; }
    return k[slot] == v;
  };
  for (int j = 0; j < ny; ++j)
    for (int i = 0; i <= nx; ++i) {
      const int32_t d = user(int64_t(j) * (nx + 1) + i);
      const int64_t e0 = kup.rp[d], e1 = kup.rp[d + 1];
      const int expect = (i > 0) + (i < nx);
      if (e1 - e0 != expect) return false;
      int32_t right = -1;
      for (int64_t e = e0; e < e1; ++e) {
        const int32_t col = kup.col[e];
        if (i > 0 && col == p_of_cell[size_t(j) * nx + i - 1]) { if (!set_const(0, kup.val[e])) return false; }
        else if (right < 0 && i < nx) { right = col; if (!set_const(1, kup.val[e])) return false; }
        else return false;
      }
      if (i < nx) {
        if (right < 0 || right >= nx * ny || cell_of_p[right] >= 0) return false;
        p_of_cell[size_t(j) * nx + i] = right;
        cell_of_p[right] = j * nx + i;
      }
    }
  for (int j = 0; j <= ny; ++j)
    for (int i = 0; i < nx; ++i) {
      const int32_t d = user(nxe + int64_t(j) * nx + i);
      const int64_t e0 = kup.rp[d], e1 = kup.rp[d + 1];
      if (e1 - e0 != (j > 0) + (j < ny)) return false;
      for (int64_t e = e0; e < e1; ++e) {
        const int32_t col = kup.col[e];
        if (j > 0 && col == p_of_cell[size_t(j - 1) * nx + i]) { if (!set_const(2, kup.val[e])) return false; }
        else if (j < ny && col == p_of_cell[size_t(j) * nx + i]) { if (!set_const(3, kup.val[e])) return false; }
        else return false;
      }
    }
  double q[4] = {0, 0, 0, 0};
  bool hq[4] = {false, false, false, false};
  for (int j = 0; j < ny; ++j)
    for (int i = 0; i < nx; ++i) {
      const int32_t r = p_of_cell[size_t(j) * nx + i];
      const int32_t edge[4] = {user(int64_t(j) * (nx + 1) + i), user(int64_t(j) * (nx + 1) + i + 1),
                               user(nxe + int64_t(j) * nx + i), user(nxe + int64_t(j + 1) * nx + i)};
      if (kpu.rp[r + 1] - kpu.rp[r] != 4) return false;
      for (int64_t e = kpu.rp[r]; e < kpu.rp[r + 1]; ++e) {
        int slot = -1;
        for (int t = 0; t < 4; ++t) if (kpu.col[e] == edge[t]) slot = t;
        if (slot < 0) return false;
        if (!hq[slot]) { hq[slot] = true; q[slot] = kpu.val[e]; }
        else if (q[slot] != kpu.val[e]) return false;
      }
    }
  for (int t = 0; t < 4; ++t) { F.kup_c[t] = k[t]; F.kpu_c[t] = q[t]; }
  F.p_identity = true;
  for (size_t cidx = 0; cidx < p_of_cell.size(); ++cidx) if (p_of_cell[cidx] != int32_t(cidx)) F.p_identity = false;
  F.p_of_cell_h = std::move(p_of_cell);
  return true;
}

// resident sweep CTAs per SM and the dynamic shared memory each may use
static void sweep_geometry(int device, int M, int ct, int want_per_sm, double sm_share, int32_t &n_cta, size_t &budget) {
  int sms = 0, per_sm = 0, reserved = 0, optin = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  cudaDeviceGetAttribute(&per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device);
  cudaDeviceGetAttribute(&reserved, cudaDevAttrReservedSharedMemoryPerBlock, device);
  cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  int per = std::max(1, want_per_sm);
  for (; per >= 1; --per) {  // as many CTAs per SM as asked for, fewer when registers or threads do not allow it
    budget = (std::min<size_t>(size_t(optin), size_t(per_sm) / per - size_t(reserved)) - 128) & ~size_t(127);
    if (sweep_ctas_per_sm(M, ct, budget) >= per) break;
  }
  per = std::max(per, 1);
  n_cta = per * std::max(1, std::min(sms, int(std::lround(sm_share * sms))));
}

// Factor in streaming order + the device copies of the layout tables (after the numeric factorisation).
static void build_sweep_factor(mmmfe_ctx *c, Factor &F, const std::vector<double> &minv) {
  const SweepLayout &L = F.lay;
  CUDA_CHECK(cudaDeviceSynchronize());
  F.Rs.alloc(size_t(L.rs_total) + 2);
  DevBuf<RepackDesc> d_desc;
  DevBuf<RepackSub> d_sub;
  DevBuf<int32_t> d_perm;
  d_desc.upload(L.repack);
  d_sub.upload(L.repack_sub);
  d_perm.upload(L.perm);
  {  // shared subtree blobs: the members of such a pack must hold bitwise identical factors
    std::vector<BlobPair> pairs;
    const NdTree &tr = F.tree;
    for (const SweepPackH &P : L.packs) {
      if (!P.shared) continue;
      const SubtreeInstance &i0 = tr.instances[L.inst_at_pos[P.pos0]];
      const SubtreeTemplate &T = tr.templates[i0.tmpl];
      for (int32_t m = 1; m < P.n; ++m) {
        const SubtreeInstance &im = tr.instances[L.inst_at_pos[P.pos0 + m]];
        pairs.push_back({i0.rf_base, im.rf_base, T.rf_size});
        pairs.push_back({i0.rb_base, im.rb_base, T.rb_size});
      }
    }
    F.shared_blob_pairs = int64_t(pairs.size());
    if (!pairs.empty()) {
      DevBuf<BlobPair> d_pairs;
      DevBuf<int> d_bad;
      d_pairs.upload(pairs);
      d_bad.alloc(1); d_bad.zero(c->stream);
      launch_blob_compare(d_pairs.p, int32_t(pairs.size()), F.R.p, d_bad.p, c->stream);
      int bad = 0;
      CUDA_CHECK(cudaMemcpyAsync(&bad, d_bad.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
      CUDA_CHECK(cudaStreamSynchronize(c->stream));
      if (bad) fail(MMMFE_E_NUMERIC, "%d subtree factor blobs that were planned as shared differ (set option sweep_share_blobs = 0)", bad);
    }
  }
  launch_repack(d_desc.p, int32_t(L.repack.size()), F.R.p, F.Rs.p, c->stream);
  launch_repack_sub(d_sub.p, int32_t(L.repack_sub.size()), d_perm.p, F.R.p, F.Rs.p, c->stream);
  CUDA_CHECK(cudaGetLastError());
  F.sweep_index.upload(L.index);
  {
    std::vector<int32_t> ij = L.inst_ij;
    ij.push_back(0); ij.push_back(0);
    F.inst_ij.upload(ij);
  }
  F.tmpl16.upload(L.tmpl16);
  F.tmpl_off.upload(L.tmpl_off);
  // 1 / (c0 |cell|) in box-major cell order (a scalar on uniform grids)
  std::vector<double> mc(size_t(L.pb_total) + 2, 0.0);
  F.minv_uniform = true;
  for (size_t cell = 0; cell < L.cell_box.size(); ++cell) {
    mc[L.cell_box[cell]] = minv[F.p_of_cell_h[cell]];
    if (mc[L.cell_box[cell]] != minv[F.p_of_cell_h[0]]) F.minv_uniform = false;
  }
  F.minv_const = minv[F.p_of_cell_h[0]];
  F.minv_c.upload(mc);
  F.cell_box.upload(L.cell_box);
  if (!F.p_identity) F.p_of_cell.upload(F.p_of_cell_h);
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  F.sweep_ok = true;
}

// Subtrees whose part of the flux matrix is bitwise identical (same shape, same entries in template-local numbering:
// every interior box of a uniform grid with constant K) get the same class: their factor blobs are identical too
// (same deterministic kernels on the same numbers), so a pack of them streams ONE blob.  Only entries with an
// interior variable matter: the others are assembled above the subtree.  Verified on the device after the numeric
// factorisation (verify_shared_blobs).
static std::vector<int32_t> subtree_classes(const NdTree &tr, const HostCsr &S, const std::vector<int32_t> &dof_of_canon) {
  const int64_t nxe = int64_t(tr.nx + 1) * tr.ny;
  std::vector<int32_t> cls(tr.instances.size(), -1);
  std::vector<uint64_t> sigs(tr.instances.size(), 0);
  auto mix = [](uint64_t x) {
    x += 0x9e3779b97f4a7c15ull; x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull; x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
  };
#pragma omp parallel
  {
    std::vector<int32_t> local(size_t(tr.n_flux), -1);  // per thread: neighbouring subtrees share boundary variables
#pragma omp for schedule(dynamic, 16)
    for (int64_t k = 0; k < int64_t(tr.instances.size()); ++k) {
      const SubtreeInstance &in = tr.instances[size_t(k)];
      const SubtreeTemplate &T = tr.templates[in.tmpl];
      auto user = [&](int32_t code) {
        const int type = code >> 30, dj = (code >> 15) & 0x7fff, di = code & 0x7fff;
        const int64_t canon = type == 0 ? int64_t(in.J0 + dj) * (tr.nx + 1) + in.I0 + di : nxe + int64_t(in.J0 + dj) * tr.nx + in.I0 + di;
        return dof_of_canon.empty() ? int32_t(canon) : dof_of_canon[canon];
      };
      for (int32_t l = 0; l < T.n_int; ++l) local[user(T.var_code[l])] = l;
      for (int32_t q = 0; q < T.n_rootb; ++q) local[user(T.rootb_code[q])] = T.n_int + q;
      uint64_t sig = 0;
      for (int32_t l = 0; l < T.n_int; ++l) {
        const int32_t row = user(T.var_code[l]);
        for (int64_t e = S.rp[row]; e < S.rp[row + 1]; ++e) {
          uint64_t bits;
          std::memcpy(&bits, &S.val[e], sizeof bits);
          sig += mix(mix(uint64_t(l) << 32 | uint32_t(local[S.col[e]])) ^ bits);
        }
      }
      for (int32_t l = 0; l < T.n_int; ++l) local[user(T.var_code[l])] = -1;
      for (int32_t q = 0; q < T.n_rootb; ++q) local[user(T.rootb_code[q])] = -1;
      sigs[size_t(k)] = sig;
    }
  }
  std::map<std::pair<int32_t, uint64_t>, int32_t> seen;
  for (size_t k = 0; k < tr.instances.size(); ++k) {  // class numbers in instance order, as before
    const int32_t tm = tr.instances[k].tmpl;
    auto it = seen.find({tm, sigs[k]});
    if (it == seen.end()) it = seen.insert({{tm, sigs[k]}, int32_t(seen.size())}).first;
    cls[k] = it->second;
  }
  return cls;
}

static double nd_factor_values(int nx, int ny, int leaf_cells);

static std::shared_ptr<Factor> factorize(mmmfe_ctx *c, const Sub &sd, int Mmax, double sm_share) {
  const auto t0 = std::chrono::steady_clock::now();
  const bool prof = std::getenv("MMMFE_SETUP_PROFILE") != nullptr;
  auto t_last = t0;
  auto lap = [&](const char *what) {
    if (!prof) return;
    cudaDeviceSynchronize();
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[setup profile] %dx%d %-28s %8.1f ms\n", sd.nx, sd.ny, what, std::chrono::duration<double>(now - t_last).count() * 1e3);
    t_last = now;
  };
  auto F = std::make_shared<Factor>();
  F->sm_share = sm_share;
  F->nx = sd.nx; F->ny = sd.ny; F->nf = sd.nf; F->np = sd.np;
  HostCsr S, kup, kpu;
  std::vector<double> minv;
  build_schur(sd, S, kup, kpu, minv);
  lap("schur complement (host)");
  F->kup.upload(kup); F->kpu.upload(kpu); F->minv.upload(minv);
  NdTree &tr = F->tree;
  const size_t smem_limit = subtree_smem_limit();
  std::vector<int32_t> dof_of_canon;
  if (!sd.canon.empty()) {
    dof_of_canon.assign(sd.nf, -1);
    for (int32_t d = 0; d < sd.nf; ++d) {
      const int32_t cidx = sd.canon[d];
      if (cidx < 0 || cidx >= sd.nf || dof_of_canon[cidx] >= 0) fail(MMMFE_E_INVALID, "canon_of_dof is not a permutation");
      dof_of_canon[cidx] = d;
    }
  }
  if (int64_t(sd.nx + 1) * sd.ny + int64_t(sd.nx) * (sd.ny + 1) != sd.nf)
    fail(MMMFE_E_INVALID, "n_flux %d does not match an RT0 grid %dx%d", sd.nf, sd.nx, sd.ny);
  F->rt0 = detect_rt0(sd, kup, kpu, dof_of_canon, *F);
  bool want_sweep = c->use_sweep && F->rt0 && c->subtree_cells > 0;
  F->sweep_ct = c->sweep_ct;
  F->sweep_ng = (c->sweep_ct / c->sweep_groups >= 32) ? c->sweep_groups : 1;
  if (want_sweep) sweep_geometry(c->device, Mmax, F->sweep_ct, c->sweep_ctas_per_sm, sm_share, F->sweep_n_cta, F->sweep_budget);
  bool sweep_planned = false;
  for (int cells = c->subtree_cells;; cells /= 2) {  // largest subtrees whose factor blob + vectors fit on chip twice per SM
    tr.build(sd.nx, sd.ny, c->leaf_cells, cells);
    for (int M = 1; M <= 4; M *= 2) F->smem_fwd[M] = F->smem_bwd[M] = 0;
    for (const auto &T : tr.templates)
      for (int M = 1; M <= 4; M *= 2) {
        F->smem_fwd[M] = std::max(F->smem_fwd[M], 16 + 8 * (size_t(T.rf_size) + size_t(T.zl_size) * M + size_t(T.n_int) * M));
        F->smem_bwd[M] = std::max(F->smem_bwd[M], 16 + 8 * (size_t(T.rb_size) + size_t(2 * T.n_int + T.n_rootb) * M));
      }
    if (cells == 0) break;
    if (F->smem_fwd[Mmax] > smem_limit / 2 || F->smem_bwd[Mmax] > smem_limit / 2) continue;
    if (!want_sweep) break;
    // the persistent sweep kernel additionally wants a ring of several entries: the waves are as large as asked for
    // (sweep_chunk) unless the ring then holds fewer than 2.6 of them -- the vectors of M right-hand sides take
    // their share of the shared memory, and a ring of barely two waves stalls the TMA stream (measured at M = 2, ring
    // of 19 478 doubles: 7168-double waves 221 ms, 8192-double waves 335 ms) -- in which case they shrink in steps
    // of 1/8
    {
      std::vector<int32_t> iu(tr.idx.size());
      for (size_t i = 0; i < tr.idx.size(); ++i) iu[i] = dof_of_canon.empty() ? tr.idx[i] : dof_of_canon[tr.idx[i]];
      std::vector<int32_t> cls;
      if (c->sweep_share_blobs) cls = subtree_classes(tr, S, dof_of_canon);
      bool layout_ok = true;
      for (int chunk = c->sweep_chunk; chunk >= 1024; chunk = (chunk * 7 / 8) & ~1) {
        if (build_sweep_layout(tr, iu, dof_of_canon, F->sweep_n_cta, chunk, F->sweep_ct / F->sweep_ng, F->lay, F->sweep_ng, std::min(c->sweep_mt, chunk), c->sweep_pack, cls) != nullptr) { layout_ok = false; break; }
        SweepSchedule probe;
        if (build_sweep_schedule(tr, F->lay, Mmax, F->sweep_ng, F->sweep_budget, c->sweep_min_fill, probe) != nullptr) continue;
        if (10 * int64_t(probe.ring_dbl) < 26 * int64_t(chunk) && chunk * 7 / 8 >= 1024) continue;
        sweep_planned = true;
        break;
      }
      if (!layout_ok) { want_sweep = false; break; }
      if (sweep_planned) break;
    }
    if (cells / 2 < std::max(1, c->leaf_cells)) {  // cannot shrink further: generic kernels with the widest subtrees
      want_sweep = false;
      cells = c->subtree_cells * 2;
    }
  }
  lap("tree + sweep plan (host)");
  if (prof) {  // the size-only recursion that sizes the SM shares must agree with the tree
    double vals = 0;
    for (const NdNode &n : tr.nodes) vals += double(n.s) * (double(n.s) + 2.0 * n.b);
    const double est = nd_factor_values(sd.nx, sd.ny, c->leaf_cells);
    if (vals != est) fail(MMMFE_E_INVALID, "nd_factor_values %.0f != tree %.0f", est, vals);
  }
  const int32_t nn = int32_t(tr.nodes.size());
  // user numbering of the front variables
  std::vector<int32_t> idx_user(tr.idx.size());
  for (size_t i = 0; i < tr.idx.size(); ++i) idx_user[i] = dof_of_canon.empty() ? tr.idx[i] : dof_of_canon[tr.idx[i]];
  std::vector<FrontDesc> fd(nn);
  std::vector<int64_t> f_off(nn);
  for (int32_t i = 0; i < nn; ++i) {
    const NdNode &n = tr.nodes[i];
    fd[i] = {n.s, n.b, n.ldr, n.ldf, n.nsplit, n.child[0], n.child[1], n.rb_off, n.rf_off, n.idx_off, n.src_off, n.z_off};
    f_off[i] = n.f_off;
  }
  F->fronts.upload(fd);
  F->idx_user.upload(idx_user);
  F->src.upload(tr.src);
  F->R.alloc(size_t(tr.r_total));
  F->R.zero(c->stream);
  if (!dof_of_canon.empty()) F->dof_of_canon.upload(dof_of_canon);
  {  // subtree templates: one int32 pool, pointers patched on the host
    std::vector<int32_t> pool;
    std::vector<std::vector<size_t>> offs;
    for (const auto &T : tr.templates) {
      std::vector<size_t> o;
      for (const std::vector<int32_t> *v : {&T.node_s, &T.node_b, &T.node_rf, &T.node_rb, &T.node_sep, &T.node_bnd, &T.node_zl,
                                            &T.var_code, &T.rootb_code, &T.sep_node, &T.sep_c0, &T.sep_c1, &T.out_node,
                                            &T.out_c0, &T.out_c1, &T.bnd_var, &T.lvl_sep, &T.lvl_out}) {
        o.push_back(pool.size());
        pool.insert(pool.end(), v->begin(), v->end());
      }
      offs.push_back(o);
    }
    pool.push_back(0);
    F->tmpl_pool.upload(pool);
    std::vector<TemplateDev> td;
    for (size_t t = 0; t < tr.templates.size(); ++t) {
      const auto &T = tr.templates[t];
      const int32_t *b = F->tmpl_pool.p;
      const auto &o = offs[t];
      td.push_back({T.n_nodes, T.n_levels, T.n_int, T.n_rootb, T.zl_size, T.rf_size, T.rb_size, b + o[0], b + o[1], b + o[2],
                    b + o[3], b + o[4], b + o[5], b + o[6], b + o[7], b + o[8], b + o[9], b + o[10], b + o[11], b + o[12],
                    b + o[13], b + o[14], b + o[15], b + o[16], b + o[17]});
    }
    td.push_back(TemplateDev{});
    F->tmpl_dev.upload(td);
    std::vector<InstanceDev> id;
    for (const auto &in : tr.instances) {
      const NdNode &root = tr.nodes[in.root];
      const auto &T = tr.templates[in.tmpl];
      id.push_back({in.tmpl, in.I0, in.J0, 0, in.rf_base, in.rb_base, root.z_off + root.s});
      F->bytes_sub_fwd += 8.0 * T.rf_size;
      F->bytes_sub_bwd += 8.0 * T.rb_size;
    }
    F->n_inst = int32_t(id.size());
    id.push_back(InstanceDev{});
    F->inst_dev.upload(id);
  }

  lap("descriptor uploads");
  // ---- numeric phase ----
  DevCsr dS;
  dS.upload(S);
  DevBuf<int32_t> d_canon, d_node_of, d_pos_of, d_idx_canon;
  if (!sd.canon.empty()) d_canon.upload(sd.canon);
  d_node_of.upload(tr.node_of);
  d_pos_of.upload(tr.pos_of);
  d_idx_canon.upload(tr.idx);
  DevBuf<int64_t> d_foff;
  d_foff.upload(f_off);
  DevBuf<double> W;
  W.alloc(size_t(tr.f_total));
  W.zero(c->stream);
  c->d_fail.alloc(1);
  c->d_fail.zero(c->stream);
  launch_assemble_original(sd.nf, dS.rp.p, dS.col.p, dS.val.p, d_canon.p, nullptr, d_node_of.p, d_pos_of.p,
                           F->fronts.p, d_foff.p, d_idx_canon.p, W.p, c->stream);
  Arena arena;
  size_t arena_need = 0;
  for (const auto &lvl : tr.by_height) {
    size_t need = 0;
    for (int32_t id : lvl) need += size_t(tr.nodes[id].s) * (tr.nodes[id].s + tr.nodes[id].b);
    arena_need = std::max(arena_need, need);
  }
  arena.buf.alloc(arena_need + 16);
  {  // room for the descriptors of the whole factorisation (~0.6 KB per front), at most 512 MB
    const size_t want = std::min(size_t(512) << 20, tr.nodes.size() * 700 + (size_t(16) << 20));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    if (c->desc_arena.n < want) c->desc_arena.alloc(want);
    c->desc_top = 0;
  }
  for (size_t h = 0; h < tr.by_height.size(); ++h) {
    const auto &lvl = tr.by_height[h];
    const int64_t nl = int64_t(lvl.size());
    if (h > 0)
      launch_extend_add(desc_upload(c, lvl.data(), lvl.size()), int32_t(nl), F->fronts.p, d_foff.p, F->src.p, W.p, c->stream);
    arena.top = 0;
    // the descriptor arrays of a level are filled in parallel (the leaf levels of a 1024^2 grid have 131 072 fronts)
    const size_t nls = size_t(nl);
    std::vector<InvProb> inv(nls);
#pragma omp parallel for schedule(static) if (nl > 4096)
    for (int64_t k = 0; k < nl; ++k) {
      const NdNode &n = tr.nodes[lvl[size_t(k)]];
      inv[size_t(k)] = {W.p + n.f_off, n.s, n.s + n.b};
    }
    chol_inv_batch(c, inv, arena);  // F11 block now holds Li = L^-1
    // positions of the fronts with a boundary part (all but the root) and of their b x s work blocks in the arena
    std::vector<int64_t> slot(size_t(nl) + 1, 0), wt_off(size_t(nl) + 1, 0);
    for (int64_t k = 0; k < nl; ++k) {
      const NdNode &n = tr.nodes[lvl[size_t(k)]];
      slot[size_t(k) + 1] = slot[size_t(k)] + (n.b > 0 ? 1 : 0);
      wt_off[size_t(k) + 1] = wt_off[size_t(k)] + int64_t(n.b) * n.s;
    }
    double *wt_base = arena.take(size_t(wt_off[size_t(nl)]));
    const size_t nb = size_t(slot[size_t(nl)]);
    std::vector<GemmProb> gW(nb), gU(nb), gX(nls), gG(nb), gF(nb);
#pragma omp parallel for schedule(static) if (nl > 4096)
    for (int64_t k = 0; k < nl; ++k) {
      const NdNode &n = tr.nodes[lvl[size_t(k)]];
      const int s = n.s, b = n.b, f = s + b;
      double *Fp = W.p + n.f_off;
      double *Rb = F->R.p + n.rb_off, *Rf = F->R.p + n.rf_off;
      const bool sub = n.subtree >= 0;
      // F11^-1 = Li^T Li: rows of R (top node, stride ldr) or the first s rows of R^T (subtree node, stride s)
      gX[size_t(k)] = {Fp, Fp, Rb, s, s, s, sub ? s : n.ldr, 1, f, f, 1, 1.0, 0.0};
      if (b == 0) continue;
      const size_t q = size_t(slot[size_t(k)]);
      double *Wt = wt_base + wt_off[size_t(k)];
      double *F21 = Fp + int64_t(s) * f, *F22 = F21 + s;
      gW[q] = {F21, Fp, Wt, b, s, s, s, f, 1, 1, f, 1.0, 0.0};          // Wt = F21 Li^T        (b x s)
      gU[q] = {Wt, Wt, F22, b, b, s, f, s, 1, 1, s, -1.0, 1.0};         // U  = F22 - Wt Wt^T
      // -G^T = -Li^T Wt^T (s x b): forward copy, and the right part of R for top nodes
      gF[q] = {Fp, Wt, Rf, s, b, s, n.ldf, 1, f, 1, s, -1.0, 0.0};
      if (sub) gG[q] = {Wt, Fp, Rb + int64_t(s) * s, b, s, s, s, s, 1, f, 1, -1.0, 0.0};  // rows s.. of R^T = -G = -Wt Li
      else gG[q] = {Fp, Wt, Rb + s, s, b, s, n.ldr, 1, f, 1, s, -1.0, 0.0};
    }
    run_gemm_batch(c, gW);
    run_gemm_batch(c, gU);
    run_gemm_batch(c, gX);
    run_gemm_batch(c, gG);
    run_gemm_batch(c, gF);
  }
  lap("numeric factorisation");
  int32_t failed = 0;
  CUDA_CHECK(cudaMemcpy(&failed, c->d_fail.p, sizeof(int32_t), cudaMemcpyDeviceToHost));
  if (failed) fail(MMMFE_E_NUMERIC, "non-positive pivot in a front: the flux Schur complement is not SPD");

  // ---- work lists of the solve (top of the tree only; subtrees are one CTA each) ----
  int64_t bytes = 0;
  for (const auto &n : tr.nodes) bytes += 8 * (int64_t(n.s) * n.b + int64_t(n.s) * (n.s + n.b));
  for (size_t h = 0; h < tr.top_by_height.size(); ++h) {
    if (tr.top_by_height[h].empty()) continue;
    F->levels.emplace_back();
    LevelWork &lw = F->levels.back();
    std::vector<WorkItem> fwd, bwd;
    for (int32_t id : tr.top_by_height[h]) {
      const NdNode &n = tr.nodes[id];
      const int64_t f = n.s + n.b;
      bytes += 4 * (2 * f + 2 * f);
      lw.f_max = std::max(lw.f_max, int32_t(f));
      lw.bytes_fwd += 8.0 * n.s * n.b + 4.0 * 3 * f;
      lw.bytes_bwd += 8.0 * n.s * f + 4.0 * f;
      const int rows_per = (n.s + n.nsplit - 1) / n.nsplit;  // <= 32
      for (int sp = 0; sp < n.nsplit; ++sp) {
        const int r0 = sp * rows_per, r1 = std::min(n.s, r0 + rows_per);
        if (n.b == 0) fwd.push_back({id, r0, r1, 0, 0, sp});
        for (int c0 = 0; c0 < n.b; c0 += 256) fwd.push_back({id, r0, r1, c0, std::min(n.b, c0 + 256), sp});
      }
      // ~64 KB of R per backward item: 8..32 rows (1..4 per warp)
      const int rows_b = int(std::min<int64_t>(32, std::max<int64_t>(8, (65536 / (8 * f)) / 8 * 8)));
      for (int r0 = 0; r0 < n.s; r0 += rows_b) bwd.push_back({id, r0, std::min(n.s, r0 + rows_b), 0, 0, 0});
    }
    lw.n_fwd = int32_t(fwd.size()); lw.n_bwd = int32_t(bwd.size());
    lw.fwd.upload(fwd); lw.bwd.upload(bwd);
  }
  F->solve_bytes = bytes;
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  lap("solve work lists");
  if (sweep_planned) build_sweep_factor(c, *F, minv);
  lap("sweep factor repack");
  F->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return F;
}

// after a stream synchronisation: did a spin-wait of the sweep kernel give up?
static void check_watchdog(mmmfe_ctx *c) {
  if (!c->abort_flag) return;
  int flag = 0;
  CUDA_CHECK(cudaMemcpy(&flag, c->abort_flag, sizeof(int), cudaMemcpyDeviceToHost));
  if (flag) {
    CUDA_CHECK(cudaMemset(c->abort_flag, 0, sizeof(int)));
    fail(MMMFE_E_CUDA, "sweep kernel watchdog: a data-flow wait did not complete (results discarded)");
  }
}

static SubtreeArgs subtree_args(const Factor &F, double *x, double *z) {
  return SubtreeArgs{F.tmpl_dev.p, F.inst_dev.p, F.R.p, x, z, F.dof_of_canon.p, F.nx, int64_t(F.nx + 1) * F.ny};
}

static void run_solve(const Factor &F, int M, double *x, double *z, cudaStream_t st) {
  SolveArgs a{F.fronts.p, F.R.p, F.idx_user.p, F.src.p, z, x};
  const SubtreeArgs sa = subtree_args(F, x, z);
  launch_subtree_forward(M, sa, F.n_inst, F.smem_fwd[M], st);
  for (const LevelWork &lw : F.levels) launch_forward_large(M, a, lw.fwd.p, lw.n_fwd, st);
  for (auto it = F.levels.rbegin(); it != F.levels.rend(); ++it) launch_backward_large(M, a, it->bwd.p, it->n_bwd, it->f_max, st);
  launch_subtree_backward(M, sa, F.n_inst, F.smem_bwd[M], st);
}

// =================================================================================================================
// sweeps
// =================================================================================================================
enum SweepMode { SWEEP_BAR, SWEEP_STAR, SWEEP_FINAL };

static void plan_batch_sweep(mmmfe_ctx *c, Batch &b, const std::vector<int32_t> &iface_q,
                             const std::vector<int64_t> &tr_base, const std::vector<int32_t> &tr_stride,
                             const std::vector<double> &tr_sign);
static void run_sweep_kernel(mmmfe_ctx *c, Batch &b, cudaStream_t st, int mode = 0);

// one time step of a batch on the per-level kernels; lvl < 0: the kernels read the level from b.level_dev (step graph)
static void enqueue_step(mmmfe_ctx *c, Batch &b, SweepMode mode, int lvl) {
  const Factor &F = *b.factor;
  cudaStream_t st = b.stream;
  const bool bar = mode == SWEEP_BAR;
  const bool with_hist = (bar && c->keep_history) || (mode == SWEEP_FINAL && c->keep_history);
  StepArgs sa = b.step;
  const bool from_dev = lvl < 0;
  if (from_dev) { sa.level_dev = b.level_dev.p; lvl = 0; }
  launch_build_rhs(sa, bar ? nullptr : c->lam_bar.p, lvl, st);
  if (bar)
    for (int j = 0; j < int(b.members.size()); ++j) {
      const Sub &sd = *c->subs[b.members[j]];
      launch_add_sparse(b.x.p, b.M, j, sd.fr_dofs.p, from_dev ? sd.fr_vals.p : sd.fr_vals.p + int64_t(lvl) * sd.n_fr, sd.n_fr, st,
                        from_dev ? b.level_dev.p : nullptr);
    }
  run_solve(F, b.M, b.x.p, b.z.p, st);
  const bool has_next = bar && (from_dev || lvl + 1 < b.nt);
  launch_update(sa, c->trace.p, lvl, has_next ? b.src_ptr.p : nullptr, F.np, with_hist ? b.hist_ptr.p : nullptr,
                int64_t(F.nf) + F.np, mode == SWEEP_FINAL ? 1 : 0, st, from_dev ? b.nt : 0);
}

static void enqueue_sweep(mmmfe_ctx *c, Batch &b, SweepMode mode) {
  const bool bar = mode == SWEEP_BAR;
  launch_init_ptil(b.step, bar ? b.p0_ptr.p : nullptr, bar ? b.src_ptr.p : nullptr, b.stream);
  for (int lvl = 0; lvl < b.nt; ++lvl) enqueue_step(c, b, mode, lvl);
}

// bar and final sweeps: ONE step captured as a CUDA graph (the level comes from a device scalar) and replayed nt times
// -- a sweep of 512 steps is ~15 000 kernel launches otherwise, bound by the launch rate of the host thread
static void enqueue_sweep_by_step_graph(mmmfe_ctx *c, Batch &b, SweepMode mode) {
  const bool bar = mode == SWEEP_BAR;
  cudaStream_t st = b.stream;
  if (!b.level_dev.p) b.level_dev.alloc(2);
  if (b.graph_step) {  // of an earlier sweep (its pointers may be stale): that sweep has long finished
    CUDA_CHECK(cudaStreamSynchronize(st));
    cudaGraphExecDestroy(b.graph_step);
    b.graph_step = nullptr;
  }
  cudaGraph_t g = nullptr;
  const long long before = launch_counter();
  CUDA_CHECK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  enqueue_step(c, b, mode, -1);
  launch_wide_next_level(b.level_dev.p, st);
  CUDA_CHECK(cudaStreamEndCapture(st, &g));
  const long long per_step = launch_counter() - before;
  launch_counter_add(-per_step);
  CUDA_CHECK(cudaGraphInstantiate(&b.graph_step, g, 0));
  CUDA_CHECK(cudaGraphDestroy(g));
  CUDA_CHECK(cudaMemsetAsync(b.level_dev.p, 0, 2 * sizeof(int32_t), st));
  launch_init_ptil(b.step, bar ? b.p0_ptr.p : nullptr, bar ? b.src_ptr.p : nullptr, st);
  for (int lvl = 0; lvl < b.nt; ++lvl) CUDA_CHECK(cudaGraphLaunch(b.graph_step, st));
  launch_counter_add(per_step * b.nt);
}

static void run_star_sweeps(mmmfe_ctx *c, SweepMode mode) {
  CUDA_CHECK(cudaEventRecord(c->ev_start, c->stream));
  for (auto &bp : c->batches) {
    Batch &b = *bp;
    // the final sweep is a star sweep that adds its solution to the stored bar levels: the persistent kernel does that
    // through its history pointers (hist_add)
    const bool final_by_sweep = mode == SWEEP_FINAL && c->keep_history && c->sweep_final && b.hist_ptr.p;
    const bool bar_by_sweep = mode == SWEEP_BAR && b.bar_sweep;
    const int sk_mode = final_by_sweep ? 1 : (bar_by_sweep ? 2 : 0);  // SweepKernelMode
    if ((mode == SWEEP_STAR || final_by_sweep || bar_by_sweep) && b.sweep) {
      Factor &F = *b.factor;
      if (F.sm_share >= 1.0) {  // the persistent kernel owns the whole GPU: batches one after the other
        run_sweep_kernel(c, b, c->stream, sk_mode);
      } else {                  // its factor's share of the SMs: the factors run side by side
        if (!F.sweep_stream) CUDA_CHECK(cudaStreamCreateWithFlags(&F.sweep_stream, cudaStreamNonBlocking));
        CUDA_CHECK(cudaStreamWaitEvent(F.sweep_stream, c->ev_start, 0));
        run_sweep_kernel(c, b, F.sweep_stream, sk_mode);
        CUDA_CHECK(cudaEventRecord(b.done, F.sweep_stream));
        CUDA_CHECK(cudaStreamWaitEvent(c->stream, b.done, 0));
      }
      continue;
    }
    CUDA_CHECK(cudaStreamWaitEvent(b.stream, c->ev_start, 0));
    if (mode == SWEEP_STAR && c->use_graphs) {
      if (!b.graph_star) {
        cudaGraph_t g = nullptr;
        const long long before = launch_counter();
        CUDA_CHECK(cudaStreamBeginCapture(b.stream, cudaStreamCaptureModeThreadLocal));
        enqueue_sweep(c, b, SWEEP_STAR);
        CUDA_CHECK(cudaStreamEndCapture(b.stream, &g));
        b.graph_kernels = launch_counter() - before;
        launch_counter_add(-b.graph_kernels);  // capture launches nothing; every replay adds them back
        CUDA_CHECK(cudaGraphInstantiate(&b.graph_star, g, 0));
        CUDA_CHECK(cudaGraphDestroy(g));
      }
      CUDA_CHECK(cudaGraphLaunch(b.graph_star, b.stream));
      launch_counter_add(b.graph_kernels);
    } else if (c->use_graphs && mode != SWEEP_STAR && b.nt >= 8) {
      enqueue_sweep_by_step_graph(c, b, mode);
    } else {
      enqueue_sweep(c, b, mode);
    }
    CUDA_CHECK(cudaEventRecord(b.done, b.stream));
    CUDA_CHECK(cudaStreamWaitEvent(c->stream, b.done, 0));
  }
}

// send = s * f2c(trace); exchange with the neighbours; out = sign * (send + recv)
static void exchange_and_combine(mmmfe_ctx *c, double sign, double *out, bool send_ready = false) {
  if (!send_ready)
    launch_spmv(c->n_iface, c->f2c_all.rp.p, c->f2c_all.col.p, c->f2c_all.val.p, c->trace.p, c->send.p, c->stream);
#ifdef MMMFE_WITH_NCCL
  if (!c->peers.empty()) {
    launch_gather(int64_t(c->send_index.n), c->send.p, c->send_index.p, c->send_buf.p, c->stream);
    nccl().GroupStart();
    for (const auto &p : c->peers) {
      nccl().Send(c->send_buf.p + p.send_off, size_t(p.count), ncclDouble, p.rank, c->comm, c->stream);
      nccl().Recv(c->recv_buf.p + p.recv_off, size_t(p.count), ncclDouble, p.rank, c->comm, c->stream);
    }
    if (nccl().GroupEnd() != ncclSuccess) fail(MMMFE_E_COMM, "NCCL face exchange failed");
  }
#endif
  launch_exchange_combine(c->n_iface, c->send.p, c->partner.p, c->recv_buf.p, sign, out, c->stream);
}

// Ap = S q for q resident in d_q (src/darcy_vtdd.cc:1316-1411)
static void apply_operator(mmmfe_ctx *c, const double *d_q, double *d_ap) {
  if (c->use_basis) {  // send = (signed f2c . star sweeps . c2f) q as one block-Toeplitz product per subdomain
    if (c->toep_parts <= 1) {
      launch_toeplitz_apply(c->toep_bm, c->toep_probs.p, c->toep_tiles.p, c->n_toep_tiles, d_q, c->send.p, 0, c->stream);
    } else {  // split over the delays: partial sums, added in fixed order
      launch_toeplitz_apply(c->toep_bm, c->toep_probs.p, c->toep_tiles.p, c->n_toep_tiles, d_q, c->toep_part_buf.p, c->n_iface, c->stream);
      launch_sum_parts(c->n_iface, c->toep_part_buf.p, c->n_iface, c->toep_parts, c->send.p, c->stream);
    }
    exchange_and_combine(c, -1.0, d_ap, true);
    return;
  }
  launch_spmv(c->n_fine, c->c2f_all.rp.p, c->c2f_all.col.p, c->c2f_all.val.p, d_q, c->lam_bar.p, c->stream);
  run_star_sweeps(c, SWEEP_STAR);
  exchange_and_combine(c, -1.0, d_ap);
}

static void allreduce_sum(mmmfe_ctx *c, double *d, size_t n) {
#ifdef MMMFE_WITH_NCCL
  if (c->world > 1) {
    if (nccl().AllReduce(d, d, n, ncclDouble, ncclSum, c->comm, c->stream) != ncclSuccess)
      fail(MMMFE_E_COMM, "NCCL allreduce failed");
  }
#else
  (void)c; (void)d; (void)n;
#endif
}


// =================================================================================================================
// multiscale flux basis / time-Toeplitz interface operator
// =================================================================================================================
// The tables of one interface side are slab-structured when (i) the mortar dofs are ordered slab-major (P per slab),
// (ii) c2f and f2c couple a mortar slab only with the lvl_per_slab fine levels it covers and (iii) every slab's block
// equals the first slab's (translation invariance in time).  project_mortar builds exactly such tables whenever the
// mortar time mesh divides the subdomain's (inc/utilities.h:471-505; front-end: projector_tables).
static bool slab_structured(const HostCsr &c2f, const HostCsr &f2c, int64_t ns, int64_t nt, int64_t m_t, int64_t P,
                            std::string &why) {
  const int64_t r = nt / m_t;
  // entries that are zero in exact arithmetic (odd Legendre modes at a midpoint) come out as round-off of either
  // sign: compare against the size of the table, not of the entry
  double c2f_max = 0, f2c_max = 0;
  for (double v : c2f.val) c2f_max = std::max(c2f_max, std::fabs(v));
  for (double v : f2c.val) f2c_max = std::max(f2c_max, std::fabs(v));
  double scale = c2f_max;
  auto close = [&](double a, double b) { return std::fabs(a - b) <= 1e-12 * scale; };
  for (int64_t lvl = 0; lvl < nt; ++lvl) {
    const int64_t J = lvl / r;
    for (int64_t i = 0; i < ns; ++i) {
      const int64_t row = lvl * ns + i, row0 = (lvl - J * r) * ns + i;
      if (c2f.rp[row + 1] - c2f.rp[row] != c2f.rp[row0 + 1] - c2f.rp[row0]) { why = "c2f rows differ between time slabs"; return false; }
      for (int64_t e = c2f.rp[row], e0 = c2f.rp[row0]; e < c2f.rp[row + 1]; ++e, ++e0) {
        if (c2f.col[e] < J * P || c2f.col[e] >= (J + 1) * P) { why = "c2f couples a fine level with another mortar time slab"; return false; }
        if (c2f.col[e] - J * P != c2f.col[e0] || !close(c2f.val[e], c2f.val[e0])) { why = "c2f is not translation invariant in time"; return false; }
      }
    }
  }
  scale = f2c_max;
  for (int64_t J = 0; J < m_t; ++J)
    for (int64_t a = 0; a < P; ++a) {
      const int64_t row = J * P + a;
      if (f2c.rp[row + 1] - f2c.rp[row] != f2c.rp[a + 1] - f2c.rp[a]) { why = "f2c rows differ between time slabs"; return false; }
      for (int64_t e = f2c.rp[row], e0 = f2c.rp[a]; e < f2c.rp[row + 1]; ++e, ++e0) {
        const int64_t lvl = f2c.col[e] / ns;
        if (lvl < J * r || lvl >= (J + 1) * r) { why = "f2c couples a mortar time slab with fine levels outside it"; return false; }
        if (f2c.col[e] - J * r * ns != f2c.col[e0] || !close(f2c.val[e], f2c.val[e0])) { why = "f2c is not translation invariant in time"; return false; }
      }
    }
  return true;
}

static bool same_csr(const HostCsr &a, const HostCsr &b) {
  return a.n_rows == b.n_rows && a.n_cols == b.n_cols && a.rp == b.rp && a.col == b.col && a.val == b.val;
}

// Slots, sizes and eligibility of the basis of one factor group; no device work.
static bool has_tables(const Sub &sd, int side) { return !sd.f2c[side].rp.empty() && !sd.c2f[side].rp.empty(); }

// Slots, sizes and eligibility of the basis of one factor group for the sides in `mask`; no device work.  The tables of
// a side are taken from any member that carries them (the front-end passes the tables of all four sides, so that ranks
// can share the work of a group whose members have different interface sides on different ranks).
static std::unique_ptr<Basis> plan_basis(mmmfe_ctx *c, const std::vector<int32_t> &members, int mask, std::string &why) {
  auto B = std::make_unique<Basis>();
  const int m_t = c->mortar_slabs;
  if (m_t < 1) { why = "option mortar_time_slabs is not set"; return nullptr; }
  const Sub &first = *c->subs[members[0]];
  if (first.nt % m_t != 0) { why = "the mortar time mesh does not divide the subdomain's time steps"; return nullptr; }
  B->m_t = m_t; B->nt = first.nt; B->lvl_per_slab = first.nt / m_t;
  for (int side = 0; side < 4; ++side) {
    if (!(mask & (1 << side))) continue;
    int slot = -1;
    for (int32_t li : members) {
      const Sub &sd = *c->subs[li];
      if (!has_tables(sd, side)) continue;
      if (sd.mortar_len[side] % m_t != 0) { why = "mortar_len is not a multiple of the number of mortar time slabs"; return nullptr; }
      if (slot < 0) {
        slot = B->n_slots++;
        B->slot_of_side[side] = slot;
        B->side_of_slot[slot] = side;
        B->P[slot] = sd.mortar_len[side] / m_t;
        B->n_side[slot] = sd.n_side[side];
        B->rep[slot] = li;
        if (!slab_structured(sd.c2f[side], sd.f2c[side], sd.n_side[side], sd.nt, m_t, B->P[slot], why)) return nullptr;
      } else {
        const Sub &rp = *c->subs[B->rep[slot]];
        if (!same_csr(rp.c2f[side], sd.c2f[side]) || !same_csr(rp.f2c[side], sd.f2c[side]) || rp.side_dofs[side] != sd.side_dofs[side]) {
          why = "subdomains that share a factor have different projector tables on the same side";
          return nullptr;
        }
      }
    }
    if (slot < 0) { why = "no projector tables for a side the multiscale basis needs"; return nullptr; }
  }
  if (B->n_slots == 0) { why = "no interface side"; return nullptr; }
  for (int k = 0; k < B->n_slots; ++k) {
    B->base[k] = B->P_tot; B->P_tot += B->P[k];
    B->tr_base[k] = B->n_tr; B->n_tr += B->n_side[k];
  }
  B->ldt = (int64_t(B->P_tot) + 7) & ~int64_t(7);
  B->col_lo = 0; B->col_hi = B->P_tot;
  return B;
}

// Solves for the responses of all first-slab basis functions of the group (MW columns per pass, every time step one
// replay of a static CUDA graph: rhs, multi-column multifrontal solve with the fronts as DMMA GEMMs, pressure update,
// traces) and projects the traces to the mortar space: T[d][row][col].
static void compute_basis(mmmfe_ctx *c, Factor &F, Basis &B) {
  const auto t_begin = std::chrono::steady_clock::now();
  cudaStream_t st = c->stream;
  const NdTree &tr = F.tree;
  const int nt = B.nt, r = B.lvl_per_slab;
  // ---- columns per pass: as many as asked for, fewer when the work vectors would not fit ----
  size_t free_b = 0, total_b = 0;
  CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
  const double t_bytes = 8.0 * B.m_t * double(B.P_tot) * double(B.ldt);
  const double per_col = 8.0 * (double(F.nf) + double(tr.z_total) + double(F.np) + double(nt) * B.n_tr);
  int cap = std::max(8, c->basis_cols_max & ~7);
  while (cap > 8 && per_col * cap > 0.6 * (double(free_b) - t_bytes)) cap -= 8;
  if (per_col * cap + t_bytes > 0.9 * double(free_b)) fail(MMMFE_E_CUDA, "not enough device memory for the multiscale basis (%.1f GB needed)", (per_col * cap + t_bytes) / 1e9);
  const int n_local = B.col_hi - B.col_lo;  // this rank's share of the columns
  B.T.alloc(size_t(B.m_t) * B.P_tot * B.ldt);
  B.T.zero(st);
  if (n_local <= 0) { B.n_pass = 0; B.MW = 8; CUDA_CHECK(cudaStreamSynchronize(st)); return; }
  B.n_pass = (n_local + cap - 1) / cap;
  B.MW = (((n_local + B.n_pass - 1) / B.n_pass) + 7) & ~7;
  const int MW = B.MW;
  // ---- subtree kernels: column chunks as wide as the shared memory allows next to the factor blob, preferring two
  // resident CTAs per SM (the level walk is latency-bound: a wider chunk costs almost the same as a narrow one) ----
  const size_t smem_limit = subtree_smem_limit();
  wide_set_smem_limit(smem_limit);
  size_t smem_f = 0, smem_b = 0;
  auto smem_for = [&](int MC) {
    smem_f = smem_b = 0;
    for (const auto &T : tr.templates) {
      smem_f = std::max(smem_f, 16 + 8 * (size_t(T.rf_size) + (size_t(T.zl_size) + size_t(T.n_int)) * (MC + 2)));
      smem_b = std::max(smem_b, 16 + 8 * (size_t(T.rb_size) + size_t(2 * T.n_int + T.n_rootb) * (MC + 2)));
    }
    return std::max(smem_f, smem_b);
  };
  B.MC = 8;
  for (int MC : {32, 16, 8}) {
    const size_t need = smem_for(MC);
    if (MC > 8 && need > (smem_limit - 2048) / 2) continue;  // two CTAs per SM (1 KB reserved per block)
    B.MC = MC;
    break;
  }
  if (smem_for(B.MC) > smem_limit) fail(MMMFE_E_UNSUPPORTED, "subtree factor blobs do not fit the multi-column kernels");
  // ---- host tables ----
  // interface data of the first slab: x[dof][column] += -sign * c2f(level, i; a), CSR by (pass, level)
  std::vector<int64_t> add_rp(size_t(B.n_pass) * (r + 1) + 1, 0);
  std::vector<int32_t> add_row, add_col;
  std::vector<double> add_val;
  int64_t add_max = 0;
  for (int pass = 0; pass < B.n_pass; ++pass) {
    for (int lvl = 0; lvl < r; ++lvl) {
      add_rp[size_t(pass) * (r + 1) + lvl] = int64_t(add_row.size());
      for (int k = 0; k < B.n_slots; ++k) {
        const int side = B.side_of_slot[k];
        const Sub &sd = *c->subs[B.rep[k]];
        const HostCsr &t = sd.c2f[side];
        const int64_t ns = B.n_side[k];
        for (int64_t i = 0; i < ns; ++i) {
          const int64_t row = int64_t(lvl) * ns + i;
          for (int64_t e = t.rp[row]; e < t.rp[row + 1]; ++e) {
            const int col = B.base[k] + t.col[e] - B.col_lo;  // first slab: t.col[e] < P (slab_structured)
            if (col < 0 || col >= n_local || col / MW != pass) continue;
            add_row.push_back(sd.side_dofs[side][size_t(i)]);
            add_col.push_back(col % MW);
            add_val.push_back(-kNormalSign[side] * t.val[e]);  // rhs_i = -s * lam_bar, src/darcy_vtdd.cc:840-844
          }
        }
      }
      add_max = std::max(add_max, int64_t(add_row.size()) - add_rp[size_t(pass) * (r + 1) + lvl]);
    }
    add_rp[size_t(pass) * (r + 1) + r] = int64_t(add_row.size());
  }
  std::vector<int32_t> tr_dof(size_t(B.n_tr));
  std::vector<int64_t> f_rp(1, 0), f_col, f_trow;
  std::vector<double> f_val;
  for (int k = 0; k < B.n_slots; ++k) {
    const int side = B.side_of_slot[k];
    const Sub &sd = *c->subs[B.rep[k]];
    for (int i = 0; i < B.n_side[k]; ++i) tr_dof[size_t(B.tr_base[k] + i)] = sd.side_dofs[side][size_t(i)];
    const HostCsr &t = sd.f2c[side];
    const int64_t ns = B.n_side[k];
    for (int64_t m = 0; m < t.n_rows; ++m) {
      for (int64_t e = t.rp[m]; e < t.rp[m + 1]; ++e) {
        const int64_t lvl = t.col[e] / ns, i = t.col[e] % ns;
        f_col.push_back(lvl * B.n_tr + B.tr_base[k] + i);
        f_val.push_back(kNormalSign[side] * t.val[e]);
      }
      f_rp.push_back(int64_t(f_col.size()));
      f_trow.push_back((m / B.P[k]) * B.P_tot + B.base[k] + m % B.P[k]);
    }
  }
  // per tree level: gather rows and GEMM problems
  struct Lvl {
    DevBuf<WideRow> fwd_rows, bwd_rows;
    DevBuf<GemmProb> fwd_probs, bwd_probs;
    DevBuf<int32_t> fwd_tiles, bwd_tiles;
    int32_t n_fwd = 0, n_bwd = 0, t_fwd = 0, t_bwd = 0;
  };
  DevBuf<double> x, z, ptil, trace;
  DevBuf<int32_t> state;
  x.alloc(size_t(F.nf) * MW); z.alloc(size_t(tr.z_total) * MW + 8); ptil.alloc(size_t(F.np) * MW);
  trace.alloc(size_t(nt) * B.n_tr * MW);
  state.alloc(2);
  z.zero(st);
  std::vector<std::unique_ptr<Lvl>> lvls;
  const int tiles_n = (MW + 63) / 64;
  for (const auto &level : tr.top_by_height) {
    if (level.empty()) continue;
    auto L = std::make_unique<Lvl>();
    std::vector<WideRow> fr, br;
    std::vector<GemmProb> fp, bp;
    std::vector<int32_t> ft(1, 0), bt(1, 0);
    for (int32_t id : level) {
      const NdNode &n = tr.nodes[id];
      const int32_t f = n.s + n.b;
      for (int32_t p = 0; p < f; ++p) fr.push_back({id, p});
      for (int32_t q = 0; q < n.b; ++q) br.push_back({id, q});
      double *zs = z.p + n.z_off * MW;
      if (n.b > 0) {  // contributions += (-G) v_s:  C(q, j) += sum_i rf[i][q] v[i][j]
        fp.push_back({F.R.p + n.rf_off, zs, zs + int64_t(n.s) * MW, n.b, MW, n.s, MW, 1, n.ldf, MW, 1, 1.0, 1.0, nullptr});
        ft.push_back(ft.back() + ((n.b + 63) / 64) * tiles_n);
      }
      // x_s = [F11^-1 | -G^T] [z_s ; x_b], written straight to the rows of x (row map = the front's dof indices)
      bp.push_back({F.R.p + n.rb_off, zs, x.p, n.s, MW, f, MW, n.ldr, 1, MW, 1, 1.0, 0.0, F.idx_user.p + n.idx_off});
      bt.push_back(bt.back() + ((n.s + 63) / 64) * tiles_n);
    }
    L->n_fwd = int32_t(fp.size()); L->n_bwd = int32_t(bp.size());
    L->t_fwd = ft.back(); L->t_bwd = bt.back();
    L->fwd_rows.upload(fr); L->bwd_rows.upload(br);
    L->fwd_probs.upload(fp); L->bwd_probs.upload(bp);
    L->fwd_tiles.upload(ft); L->bwd_tiles.upload(bt);
    lvls.push_back(std::move(L));
  }
  DevBuf<int64_t> d_add_rp, d_frp, d_fcol, d_ftrow;
  DevBuf<int32_t> d_add_row, d_add_col, d_tr_dof;
  DevBuf<double> d_add_val, d_fval;
  add_row.push_back(0); add_col.push_back(0); add_val.push_back(0.0);
  d_add_rp.upload(add_rp); d_add_row.upload(add_row); d_add_col.upload(add_col); d_add_val.upload(add_val);
  d_tr_dof.upload(tr_dof);
  d_frp.upload(f_rp); d_fcol.upload(f_col); d_ftrow.upload(f_trow); d_fval.upload(f_val);

  const WideStep ws{MW, F.nf, F.np, F.kup.rp.p, F.kup.col.p, F.kup.val.p, F.kpu.rp.p, F.kpu.col.p, F.kpu.val.p,
                    F.minv.p, x.p, ptil.p, state.p};
  const SolveArgs sa{F.fronts.p, F.R.p, F.idx_user.p, F.src.p, z.p, x.p};
  const SubtreeArgs sub = subtree_args(F, x.p, z.p);
  auto enqueue_step = [&]() {
    launch_wide_rhs(ws, st);
    launch_wide_add(ws, d_add_rp.p, d_add_row.p, d_add_col.p, d_add_val.p, r, add_max, st);
    launch_wide_subtree_forward(B.MC, MW, sub, F.n_inst, smem_f, st);
    for (auto &L : lvls) {
      launch_wide_fwd_gather(sa, MW, L->fwd_rows.p, int64_t(L->fwd_rows.n), st);
      launch_grouped_gemm(L->fwd_probs.p, L->fwd_tiles.p, L->n_fwd, L->t_fwd, st);
    }
    for (auto it = lvls.rbegin(); it != lvls.rend(); ++it) {
      launch_wide_bwd_gather(sa, MW, (*it)->bwd_rows.p, int64_t((*it)->bwd_rows.n), st);
      launch_grouped_gemm((*it)->bwd_probs.p, (*it)->bwd_tiles.p, (*it)->n_bwd, (*it)->t_bwd, st);
    }
    launch_wide_subtree_backward(B.MC, MW, sub, F.n_inst, smem_b, st);
    launch_wide_update(ws, d_tr_dof.p, B.n_tr, trace.p, st);
    launch_wide_next_level(state.p, st);
  };
  CUDA_CHECK(cudaStreamSynchronize(st));
  if (std::getenv("MMMFE_BASIS_PROFILE")) {  // one instrumented step (level 0 of pass 0), microseconds per kernel class
    const int32_t init[2] = {0, 0};
    CUDA_CHECK(cudaMemcpy(state.p, init, sizeof init, cudaMemcpyHostToDevice));
    ptil.zero(st);
    std::vector<cudaEvent_t> ev;
    std::vector<const char *> name;
    auto mark = [&](const char *nm) {
      cudaEvent_t e;
      CUDA_CHECK(cudaEventCreate(&e));
      CUDA_CHECK(cudaEventRecord(e, st));
      ev.push_back(e); name.push_back(nm);
    };
    for (int rep = 0; rep < 2; ++rep) {  // second repetition is reported (warm caches, attributes set)
      ev.clear(); name.clear();
      mark("start");
      launch_wide_rhs(ws, st);
      launch_wide_add(ws, d_add_rp.p, d_add_row.p, d_add_col.p, d_add_val.p, r, add_max, st); mark("rhs");
      launch_wide_subtree_forward(B.MC, MW, sub, F.n_inst, smem_f, st); mark("subtree_fwd");
      for (auto &L : lvls) launch_wide_fwd_gather(sa, MW, L->fwd_rows.p, int64_t(L->fwd_rows.n), st);
      mark("fwd_gathers(all levels, without their gemms)");
      for (auto &L : lvls) launch_grouped_gemm(L->fwd_probs.p, L->fwd_tiles.p, L->n_fwd, L->t_fwd, st);
      mark("fwd_gemms");
      for (auto it = lvls.rbegin(); it != lvls.rend(); ++it) launch_wide_bwd_gather(sa, MW, (*it)->bwd_rows.p, int64_t((*it)->bwd_rows.n), st);
      mark("bwd_gathers");
      for (auto it = lvls.rbegin(); it != lvls.rend(); ++it) launch_grouped_gemm((*it)->bwd_probs.p, (*it)->bwd_tiles.p, (*it)->n_bwd, (*it)->t_bwd, st);
      mark("bwd_gemms");
      launch_wide_subtree_backward(B.MC, MW, sub, F.n_inst, smem_b, st); mark("subtree_bwd");
      launch_wide_update(ws, d_tr_dof.p, B.n_tr, trace.p, st); mark("update");
      CUDA_CHECK(cudaStreamSynchronize(st));
    }
    std::fprintf(stderr, "[basis profile] grid %dx%d nt %d columns %d MW %d passes %d MC %d levels %zu subtrees %d smem %zu/%zu:",
                 F.nx, F.ny, nt, B.P_tot, MW, B.n_pass, B.MC, lvls.size(), F.n_inst, smem_f, smem_b);
    for (size_t i = 1; i < ev.size(); ++i) {
      float ms = 0;
      cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
      std::fprintf(stderr, " %s %.1f us;", name[i], ms * 1e3);
    }
    std::fprintf(stderr, "\n");
    for (auto e : ev) cudaEventDestroy(e);
    if (std::atoi(std::getenv("MMMFE_BASIS_PROFILE")) == 2) fail(MMMFE_E_INVALID, "MMMFE_BASIS_PROFILE=2: profile only");
  }
  cudaGraph_t g = nullptr;
  cudaGraphExec_t ge = nullptr;
  if (c->use_graphs) {
    const long long before = launch_counter();
    CUDA_CHECK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    enqueue_step();
    CUDA_CHECK(cudaStreamEndCapture(st, &g));
    B.graph_kernels = launch_counter() - before;
    launch_counter_add(-B.graph_kernels);
    CUDA_CHECK(cudaGraphInstantiate(&ge, g, 0));
    CUDA_CHECK(cudaGraphDestroy(g));
  }
  // The star problem has zero data after the first mortar slab and backward Euler contracts: once the pressures
  // (the only state carried from step to step) have decayed below basis_drop_tol (default 1e-18) of their peak, every
  // later trace is below that fraction of the peak response too -- far under the rounding of the sums it would enter
  // -- and the remaining steps of the pass are skipped (their traces stay zero).  Checked at the end of every slab.
  DevBuf<unsigned long long> d_peak;
  d_peak.alloc(1);
  int64_t steps_done = 0;
  B.d_used = 0;
  for (int pass = 0; pass < B.n_pass; ++pass) {
    const int32_t init[2] = {0, pass};
    CUDA_CHECK(cudaMemcpyAsync(state.p, init, sizeof init, cudaMemcpyHostToDevice, st));
    CUDA_CHECK(cudaStreamSynchronize(st));  // `init` is a stack variable
    ptil.zero(st);
    trace.zero(st);
    double peak = 0.0;
    int done = nt;
    for (int lvl = 0; lvl < nt; ++lvl) {
      if (ge) { CUDA_CHECK(cudaGraphLaunch(ge, st)); launch_counter_add(B.graph_kernels); }
      else enqueue_step();
      if (c->basis_drop_tol > 0 && (lvl + 1) % r == 0 && lvl + 1 < nt) {
        d_peak.zero(st);
        launch_absmax(int64_t(ptil.n), ptil.p, d_peak.p, st);
        unsigned long long bits = 0;
        CUDA_CHECK(cudaMemcpyAsync(&bits, d_peak.p, sizeof bits, cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
        double v;
        std::memcpy(&v, &bits, sizeof v);
        peak = std::max(peak, v);
        if (v <= c->basis_drop_tol * peak) { done = lvl + 1; break; }
      }
    }
    steps_done += done;
    B.d_used = std::max(B.d_used, (done + r - 1) / r);
    const int n_cols = std::min(MW, n_local - pass * MW);
    launch_wide_f2c(int64_t(f_trow.size()), d_frp.p, d_fcol.p, d_fval.p, trace.p, MW, n_cols, B.T.p, B.ldt, d_ftrow.p,
                    int64_t(B.col_lo) + int64_t(pass) * MW, st);
  }
  CUDA_CHECK(cudaStreamSynchronize(st));
  CUDA_CHECK(cudaGetLastError());
  if (ge) cudaGraphExecDestroy(ge);
  double vals = 0;
  for (const NdNode &n : tr.nodes) vals += double(n.s) * (double(n.s) + 2.0 * n.b);
  B.flops = 2.0 * vals * double(steps_done) * MW;
  B.bytes = double(steps_done) * (8.0 * vals + 8.0 * MW * (4.0 * F.nf + 3.0 * F.np));
  B.steps = steps_done;
  B.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
}

static uint64_t mix64(uint64_t h, uint64_t x) {
  h ^= x + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
  h = (h ^ (h >> 30)) * 0xbf58476d1ce4e5b9ull;
  return h ^ (h >> 27);
}
template <class T>
static uint64_t hash_vec(uint64_t h, const std::vector<T> &v) {
  h = mix64(h, v.size());
  for (const T &e : v) {
    uint64_t bits = 0;
    std::memcpy(&bits, &e, sizeof(T) < 8 ? sizeof(T) : 8);
    h = mix64(h, bits);
  }
  return h;
}
// identifies a factor group across ranks: same matrix, numbering, time steps and projector tables of every side
static uint64_t group_signature(const mmmfe_ctx *c, const std::vector<int32_t> &members) {
  const Sub &sd = *c->subs[members[0]];
  uint64_t h = 0x1234567ull;
  for (int v : {sd.nx, sd.ny, sd.nt, sd.nf, sd.np, c->mortar_slabs}) h = mix64(h, uint64_t(v));
  uint64_t bits;
  std::memcpy(&bits, &sd.dt, 8);
  h = mix64(h, bits);
  h = hash_vec(h, sd.matrix.rp); h = hash_vec(h, sd.matrix.col); h = hash_vec(h, sd.matrix.val); h = hash_vec(h, sd.canon);
  for (int side = 0; side < 4; ++side)
    for (int32_t li : members) {
      const Sub &m = *c->subs[li];
      if (!has_tables(m, side)) continue;
      h = mix64(h, uint64_t(side));
      h = hash_vec(h, m.c2f[side].col); h = hash_vec(h, m.c2f[side].val);
      h = hash_vec(h, m.f2c[side].col); h = hash_vec(h, m.f2c[side].val); h = hash_vec(h, m.side_dofs[side]);
      break;
    }
  return h;
}

// Decides whether the interface operator runs on the multiscale basis, computes the bases and the tile list of the
// block-Toeplitz kernel.  With several ranks the groups are matched by signature: when EVERY rank holds a group of a
// signature and has the tables of all the sides any of them needs, the ranks split that group's columns among
// themselves and sum the (disjoint) results with one all-reduce -- a checkerboard over 8 GPUs solves for an eighth of
// the columns per GPU.  All ranks take the same decisions from the same exchanged data.
static void setup_basis_operator(mmmfe_ctx *c, const std::vector<int> &group_of) {
  c->use_basis = false;
  c->basis_note.clear();
  if (c->operator_mode == 0) { c->basis_note = "operator = 0"; return; }
  const int ns = int(c->subs.size());
  const int ng = int(c->factors.size());
  std::vector<std::vector<int32_t>> members(ng);
  for (int i = 0; i < ns; ++i) members[group_of[i]].push_back(i);
  std::vector<int> mask(ng, 0);
  for (int g = 0; g < ng; ++g)
    for (int32_t li : members[g])
      for (int side = 0; side < 4; ++side)
        if (c->subs[li]->neighbor[side] >= 0) mask[g] |= 1 << side;
  std::vector<int> share_ranks(ng, 1), share_index(ng, 0);
  // ---- match groups across ranks ----
  constexpr int MAXG = 8, REC = 6;
  const bool try_share = c->world > 1 && c->basis_share && ng <= MAXG;
  std::vector<double> info;
  if (c->world > 1) {
    // every rank takes part in this exchange whatever its options are (the record says whether it can share)
    info.assign(size_t(c->world) * MAXG * REC, 0.0);
    for (int g = 0; g < ng && g < MAXG; ++g) {
      std::string why;
      const bool own_ok = plan_basis(c, members[g], mask[g], why) != nullptr;
      bool all_tables = true;
      for (int side = 0; side < 4; ++side) {
        bool any = false;
        for (int32_t li : members[g]) any |= has_tables(*c->subs[li], side);
        all_tables &= any;
      }
      const uint64_t sig = group_signature(c, members[g]);
      double *rec = &info[(size_t(c->rank) * MAXG + g) * REC];
      rec[0] = 1.0; rec[1] = double(sig & 0xffffffffull); rec[2] = double(sig >> 32); rec[3] = double(mask[g]);
      rec[4] = (try_share && own_ok && all_tables) ? 1.0 : 0.0;
    }
    DevBuf<double> d;
    d.upload(info);
    allreduce_sum(c, d.p, info.size());
    CUDA_CHECK(cudaMemcpyAsync(info.data(), d.p, info.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    for (int g = 0; g < ng && g < MAXG; ++g) {
      const double *mine = &info[(size_t(c->rank) * MAXG + g) * REC];
      int holders = 0, can = 0, uni = 0, before = 0;
      for (int rk = 0; rk < c->world; ++rk)
        for (int k = 0; k < MAXG; ++k) {
          const double *rec = &info[(size_t(rk) * MAXG + k) * REC];
          if (rec[0] != 1.0 || rec[1] != mine[1] || rec[2] != mine[2]) continue;
          ++holders; can += rec[4] == 1.0; uni |= int(rec[3]);
          if (rk < c->rank) ++before;
          break;  // one group per signature and rank
        }
      if (holders == c->world && can == c->world) {  // every rank holds it and can share
        share_ranks[g] = c->world; share_index[g] = before; mask[g] = uni;
      }
    }
  }
  // ---- plans (identical on the ranks that share a group) ----
  std::vector<std::unique_ptr<Basis>> plans(ng);
  for (int g = 0; g < ng; ++g) {
    std::string why;
    plans[g] = plan_basis(c, members[g], mask[g], why);
    if (!plans[g]) { c->basis_note = why; break; }
    if (c->operator_mode == 2 && plans[g]->P_tot > c->basis_auto_columns) {
      c->basis_note = "more basis columns per factor group than basis_auto_columns";
      plans[g].reset();
      break;
    }
    Basis &B = *plans[g];
    B.shared_by = share_ranks[g];
    if (B.shared_by > 1) {  // contiguous column blocks, multiples of 8 columns
      const int64_t blocks = (B.P_tot + 7) / 8;
      B.col_lo = int32_t(std::min<int64_t>(B.P_tot, 8 * (blocks * share_index[g] / B.shared_by)));
      B.col_hi = int32_t(std::min<int64_t>(B.P_tot, 8 * (blocks * (share_index[g] + 1) / B.shared_by)));
    }
  }
  bool all = true;
  for (int g = 0; g < ng; ++g) all &= plans[g] != nullptr;
  // One operator for the whole run: the basis is used only when EVERY rank can use it (also keeps the collectives of
  // the shared groups matched).
  if (c->world > 1) {
    DevBuf<double> flag;
    const double mine = all ? 1.0 : 0.0;
    flag.upload(&mine, 1);
    allreduce_sum(c, flag.p, 1);
    double sum = 0;
    CUDA_CHECK(cudaMemcpyAsync(&sum, flag.p, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    if (all && sum != double(c->world)) { all = false; c->basis_note = "another rank cannot use the multiscale basis"; }
  }
  if (!all) {
    if (c->operator_mode == 1) fail(MMMFE_E_UNSUPPORTED, "operator = 1 (multiscale basis) is not possible here: %s", c->basis_note.c_str());
    return;
  }
  for (int g = 0; g < ng; ++g) {
    Factor &F = *c->factors[g];
    F.basis = std::move(plans[g]);
    compute_basis(c, F, *F.basis);
    c->basis_seconds += F.basis->seconds;
    c->basis_flops += F.basis->flops;
    c->basis_bytes += F.basis->bytes;
    c->basis_columns += F.basis->col_hi - F.basis->col_lo;
  }
  // shared groups: the ranks hold disjoint column blocks (zeros elsewhere); groups in the order of their signature
  {
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::pair<std::pair<double, double>, int>> order;
    for (int g = 0; g < ng && g < MAXG; ++g)
      if (share_ranks[g] > 1) {
        const double *mine = &info[(size_t(c->rank) * MAXG + g) * REC];
        order.push_back({{mine[2], mine[1]}, g});
      }
    std::sort(order.begin(), order.end());
    std::vector<double> used(order.size() * size_t(c->world), 0.0);
    for (size_t k = 0; k < order.size(); ++k) {
      Basis &B = *c->factors[order[k].second]->basis;
      allreduce_sum(c, B.T.p, B.T.n);
      used[k * c->world + c->rank] = B.d_used;
    }
    if (!order.empty()) {  // the delays any rank's columns reach
      DevBuf<double> d;
      d.upload(used);
      allreduce_sum(c, d.p, used.size());
      CUDA_CHECK(cudaMemcpyAsync(used.data(), d.p, used.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
      CUDA_CHECK(cudaStreamSynchronize(c->stream));
      for (size_t k = 0; k < order.size(); ++k) {
        Basis &B = *c->factors[order[k].second]->basis;
        for (int rk = 0; rk < c->world; ++rk) B.d_used = std::max(B.d_used, int32_t(used[k * c->world + rk]));
      }
      c->basis_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
  }
  std::vector<ToepProb> probs;
  std::vector<ToepTile> tiles;
  std::vector<int> prob_dmax;
  c->toep_flops = c->toep_bytes = 0;
  for (int i = 0; i < ns; ++i) {
    const Sub &sd = *c->subs[i];
    const Basis &B = *c->factors[group_of[i]]->basis;
    for (int so = 0; so < 4; ++so) {
      if (sd.neighbor[so] < 0) continue;
      ToepProb pr{};
      const int ko = B.slot_of_side[so];
      pr.T = B.T.p; pr.ldt = B.ldt; pr.d_stride = int64_t(B.P_tot) * B.ldt; pr.m_t = B.m_t;
      pr.P_out = B.P[ko]; pr.row_base = B.base[ko]; pr.out_off = sd.iface_off[so];
      double k_len = 0;
      for (int si = 0; si < 4; ++si) {
        if (sd.neighbor[si] < 0) continue;
        const int ki = B.slot_of_side[si];
        pr.P_in[pr.n_in] = B.P[ki]; pr.col_base[pr.n_in] = B.base[ki]; pr.in_off[pr.n_in] = sd.iface_off[si];
        ++pr.n_in;
        k_len += B.P[ki];
      }
      const double du = std::max(1, B.d_used);
      c->toep_flops += 2.0 * pr.P_out * k_len * (du * B.m_t - 0.5 * du * (du - 1.0));  // slab I sees delays 0 .. min(I, du-1)
      c->toep_bytes += 8.0 * pr.P_out * k_len * du;
      probs.push_back(pr);
      prob_dmax.push_back(std::max(1, B.d_used));
    }
  }
  // rows per tile: large tiles re-read q less often, small problems need small tiles to yield enough CTAs; the delay
  // range of every slab tile is then cut into parts (split-k) until the machine is filled about four times
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
  auto base_tiles = [&](int bm) {
    int64_t n = 0;
    for (const ToepProb &pr : probs) n += int64_t((pr.P_out + bm - 1) / bm) * ((pr.m_t + TOEP_BN - 1) / TOEP_BN);
    return n;
  };
  c->toep_bm = base_tiles(128) >= 96 ? 128 : (base_tiles(64) >= 48 ? 64 : 32);
  const int64_t n_base = std::max<int64_t>(1, base_tiles(c->toep_bm));
  c->toep_parts = int(std::min<int64_t>(16, std::max<int64_t>(1, (4 * sms + n_base - 1) / n_base)));
  int max_part = 0;
  for (size_t p = 0; p < probs.size(); ++p)
    for (int row0 = 0; row0 < probs[p].P_out; row0 += c->toep_bm)
      for (int slab0 = 0; slab0 < probs[p].m_t; slab0 += TOEP_BN) {
        const int nd = std::min(std::min(probs[p].m_t, slab0 + TOEP_BN), prob_dmax[p]);  // delays 0 .. nd-1 reach this tile's slabs and carry data
        const int parts = std::min(c->toep_parts, nd);
        for (int k = 0; k < parts; ++k) {
          tiles.push_back({int32_t(p), row0, slab0, int32_t(int64_t(k) * nd / parts), int32_t(int64_t(k + 1) * nd / parts), k});
          max_part = std::max(max_part, k);
        }
      }
  c->toep_parts = max_part + 1;
  if (c->toep_parts > 1) { c->toep_part_buf.alloc(size_t(c->toep_parts) * size_t(c->n_iface)); c->toep_part_buf.zero(c->stream); }
  // long tiles first
  std::stable_sort(tiles.begin(), tiles.end(), [](const ToepTile &a, const ToepTile &b) { return a.d1 - a.d0 > b.d1 - b.d0; });
  c->toep_probs.upload(probs);
  c->toep_tiles.upload(tiles);
  c->n_toep_tiles = int32_t(tiles.size());
  c->use_basis = true;
}

// =================================================================================================================
// setup
// =================================================================================================================
// sum over the fronts of s (s + 2 b) for the nested dissection NdTree::build would produce (same bisection rule,
// sizes only): the factor values one solve multiplies with, without building the tree
static double nd_factor_values(int nx, int ny, int leaf_cells) {
  std::function<double(int, int, int, int)> rec = [&](int i0, int i1, int j0, int j1) -> double {
    const int w = i1 - i0, h = j1 - j0;
    double below = 0.0, s;
    if (int64_t(w) * h <= leaf_cells || (w == 1 && h == 1)) {
      s = double(h) * ((w - 1) + (i0 == 0) + (i1 == nx)) + double(w) * ((h - 1) + (j0 == 0) + (j1 == ny));
      if (s == 0.0) return 0.0;
    } else if (w >= h) {
      const int is = i0 + w / 2;
      below = rec(i0, is, j0, j1) + rec(is, i1, j0, j1);
      s = h;
    } else {
      const int js = j0 + h / 2;
      below = rec(i0, i1, j0, js) + rec(i0, i1, js, j1);
      s = w;
    }
    const double b = double(h) * ((i0 > 0) + (i1 < nx)) + double(w) * ((j0 > 0) + (j1 < ny));
    return below + s * (s + 2.0 * b);
  };
  return rec(0, nx, 0, ny);
}

static bool same_matrix(const Sub &a, const Sub &b) {
  return a.nx == b.nx && a.ny == b.ny && a.nt == b.nt && a.nf == b.nf && a.np == b.np && a.dt == b.dt &&
         a.canon == b.canon && a.matrix.rp == b.matrix.rp && a.matrix.col == b.matrix.col &&
         a.matrix.val == b.matrix.val;
}

static void setup(mmmfe_ctx *c) {
  if (c->is_setup) fail(MMMFE_E_INVALID, "mmmfe_setup called twice");
  if (c->subs.empty()) fail(MMMFE_E_INVALID, "no subdomains registered");
  CUDA_CHECK(cudaSetDevice(c->device));
  if (!c->stream) CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_start, cudaEventDisableTiming));
  CUDA_CHECK(cudaEventCreate(&c->ev_t0));
  CUDA_CHECK(cudaEventCreate(&c->ev_t1));
  const int ns = int(c->subs.size());
  const bool prof = std::getenv("MMMFE_SETUP_PROFILE") != nullptr;
  auto t_last = std::chrono::steady_clock::now();
  auto lap = [&](const char *what) {
    if (!prof) return;
    cudaDeviceSynchronize();
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[setup profile] set-up: %-36s %8.1f ms\n", what, std::chrono::duration<double>(now - t_last).count() * 1e3);
    t_last = now;
  };
  // ---- group identical matrices, factorise one representative each ----
  std::vector<int> group_of(ns, -1);
  std::vector<int> reps;
  for (int i = 0; i < ns; ++i) {
    for (size_t g = 0; g < reps.size(); ++g)
      if (same_matrix(*c->subs[reps[g]], *c->subs[i])) { group_of[i] = int(g); break; }
    if (group_of[i] < 0) { group_of[i] = int(reps.size()); reps.push_back(i); }
  }
  // Concurrent mode: every factor group gets a share of the SMs proportional to the bytes its batches stream in one
  // operator application (nt * factor values per batch; the factor size is known from the elimination tree alone).
  std::vector<double> share(reps.size(), 1.0);
  if (c->sweep_concurrent && c->use_sweep && reps.size() > 1) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    std::vector<double> w(reps.size(), 0.0);
    double tot = 0;
    for (size_t g = 0; g < reps.size(); ++g) {
      const Sub &sd = *c->subs[reps[g]];
      const double vals = nd_factor_values(sd.nx, sd.ny, c->leaf_cells);
      int cnt = 0;
      for (int i = 0; i < ns; ++i) cnt += group_of[i] == int(g);
      w[g] = vals * sd.nt * ((cnt + 3) / 4);
      w[g] = std::pow(w[g], c->sweep_share_exponent);  // < 1 favours the small groups (they are latency-bound)
      tot += w[g];
    }
    // at least 1/16 of the machine each; the shares add up to all SMs
    int left = sms;
    std::vector<int> n_sm(reps.size(), 0);
    for (size_t g = 0; g < reps.size(); ++g) { n_sm[g] = std::max(sms / 16, int(std::floor(w[g] / tot * sms))); left -= n_sm[g]; }
    size_t big = 0;
    for (size_t g = 1; g < reps.size(); ++g) if (w[g] > w[big]) big = g;
    n_sm[big] += left;
    if (n_sm[big] >= sms / 16)
      for (size_t g = 0; g < reps.size(); ++g) share[g] = double(n_sm[g]) / sms;
  }
  lap("matrix groups + SM shares");
  for (size_t g = 0; g < reps.size(); ++g) {
    int cnt = 0;
    for (int i = 0; i < ns; ++i) cnt += group_of[i] == int(g);
    const auto tf0 = std::chrono::steady_clock::now();
    c->factors.push_back(factorize(c, *c->subs[reps[g]], cnt <= 1 ? 1 : (cnt == 2 ? 2 : 4), share[g]));
    c->factor_seconds += c->factors.back()->seconds;
    if (prof)
      std::fprintf(stderr, "[setup profile] set-up: factorize() body %.1f ms, with the release of its work buffers %.1f ms\n",
                   c->factors.back()->seconds * 1e3, std::chrono::duration<double>(std::chrono::steady_clock::now() - tf0).count() * 1e3);
  }
  lap("factorize (all groups)");
  // ---- interface and fine-trace layouts ----
  int64_t io = 0, fo = 0;
  std::map<std::pair<int, int>, int64_t> local_off;  // (gid, side) -> offset
  for (auto &sp : c->subs)
    for (int s = 0; s < 4; ++s)
      if (sp->neighbor[s] >= 0) {
        sp->iface_off[s] = io; io += sp->mortar_len[s];
        sp->fine_off[s] = fo; fo += int64_t(sp->nt) * sp->n_side[s];
        local_off[{sp->gid, s}] = sp->iface_off[s];
      }
  c->n_iface = io; c->n_fine = fo;
  HostCsr c2f, f2c;
  c2f.n_rows = fo; c2f.n_cols = io; c2f.rp.assign(1, 0);
  f2c.n_rows = io; f2c.n_cols = fo; f2c.rp.assign(1, 0);
  for (auto &sp : c->subs)
    for (int s = 0; s < 4; ++s)
      if (sp->neighbor[s] >= 0) {
        const HostCsr &a = sp->c2f[s], &b = sp->f2c[s];
        if (a.n_rows != int64_t(sp->nt) * sp->n_side[s] || a.n_cols != sp->mortar_len[s] ||
            b.n_cols != a.n_rows || b.n_rows != a.n_cols)
          fail(MMMFE_E_INVALID, "projector table shapes of subdomain %d side %d are inconsistent", sp->gid, s);
        for (int64_t r = 0; r < a.n_rows; ++r) {
          for (int64_t k = a.rp[r]; k < a.rp[r + 1]; ++k) {
            c2f.col.push_back(int32_t(sp->iface_off[s] + a.col[k])); c2f.val.push_back(a.val[k]);
          }
          c2f.rp.push_back(int64_t(c2f.col.size()));
        }
        for (int64_t r = 0; r < b.n_rows; ++r) {
          for (int64_t k = b.rp[r]; k < b.rp[r + 1]; ++k) {
            f2c.col.push_back(int32_t(sp->fine_off[s] + b.col[k])); f2c.val.push_back(kNormalSign[s] * b.val[k]);
          }
          f2c.rp.push_back(int64_t(f2c.col.size()));
        }
      }
  c->c2f_all.upload(c2f); c->f2c_all.upload(f2c);
  // ---- exchange plan (exchange_plan.cpp: host only, checked by the 2-rank CPU tests) ----
  std::vector<ExchangeSide> ex_sides;
  for (auto &sp : c->subs)
    for (int s = 0; s < 4; ++s)
      if (sp->neighbor[s] >= 0) ex_sides.push_back({sp->gid, s, sp->neighbor[s], sp->iface_off[s], sp->mortar_len[s]});
  ExchangePlan xp;
  if (const char *err = build_exchange_plan(ex_sides, io, c->owner, c->rank, c->world, xp)) fail(MMMFE_E_INVALID, "%s", err);
  std::vector<int64_t> &partner = xp.partner, &send_index = xp.send_index;
  int64_t roff = 0;
  for (const ExchangePeer &p : xp.peers) { c->peers.push_back({p.rank, p.send_off, p.recv_off, p.count}); roff += p.count; }
  for (int64_t v : partner) if (v == -1) fail(MMMFE_E_INVALID, "interface entry without a partner");
  c->partner.upload(partner);
  c->send_index.upload(send_index);
  c->send_buf.alloc(send_index.size() + 1);
  c->recv_buf.alloc(size_t(roff) + 1);
  c->lam_bar.alloc(size_t(fo)); c->trace.alloc(size_t(fo));
  c->send.alloc(size_t(io)); c->ap.alloc(size_t(io)); c->qin.alloc(size_t(io)); c->lam.alloc(size_t(io));
  c->r0.alloc(size_t(io));
  c->lam_bar.zero(); c->trace.zero(); c->qin.zero(); c->lam.zero();
  lap("projector + exchange tables");
  // ---- batches ----
  for (size_t g = 0; g < reps.size(); ++g) {
    std::vector<int32_t> mem;
    for (int i = 0; i < ns; ++i) if (group_of[i] == int(g)) mem.push_back(i);
    for (size_t m0 = 0; m0 < mem.size(); m0 += 4) {
      auto b = std::make_unique<Batch>();
      b->factor = c->factors[g];
      b->members.assign(mem.begin() + m0, mem.begin() + std::min(mem.size(), m0 + 4));
      const int cnt = int(b->members.size());
      b->M = cnt <= 1 ? 1 : (cnt == 2 ? 2 : 4);
      b->nt = c->subs[b->members[0]]->nt;
      const Factor &F = *b->factor;
      b->x.alloc(size_t(F.nf) * b->M); b->x.zero();
      b->z.alloc(size_t(F.tree.z_total) * b->M); b->z.zero();
      b->ptil.alloc(size_t(F.np) * b->M); b->ptil.zero();
      std::vector<int32_t> tr_dof, tr_mem, tr_stride, iface_q(size_t(F.nf) * b->M, -1);
      std::vector<int64_t> tr_base;
      std::vector<double> tr_sign;
      for (int j = 0; j < cnt; ++j) {
        Sub &sd = *c->subs[b->members[j]];
        sd.batch = int32_t(c->batches.size()); sd.member = j;
        for (int s = 0; s < 4; ++s)
          if (sd.neighbor[s] >= 0)
            for (int i = 0; i < sd.n_side[s]; ++i) {
              const int32_t d = sd.side_dofs[s][i];
              if (d < 0 || d >= sd.nf) fail(MMMFE_E_INVALID, "side_dofs out of range");
              iface_q[size_t(d) * b->M + j] = int32_t(tr_dof.size());
              tr_dof.push_back(d); tr_mem.push_back(j); tr_stride.push_back(sd.n_side[s]);
              tr_base.push_back(sd.fine_off[s] + i); tr_sign.push_back(kNormalSign[s]);
            }
      }
      b->n_tr = int32_t(tr_dof.size());
      b->tr_dof.upload(tr_dof); b->tr_mem.upload(tr_mem); b->tr_stride.upload(tr_stride);
      b->tr_base.upload(tr_base); b->tr_sign.upload(tr_sign); b->iface_q.upload(iface_q);
      CUDA_CHECK(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
      CUDA_CHECK(cudaEventCreateWithFlags(&b->done, cudaEventDisableTiming));
      b->step = StepArgs{b->M, F.nf, F.np, F.kup.rp.p, F.kup.col.p, F.kup.val.p, F.kpu.rp.p, F.kpu.col.p,
                         F.kpu.val.p, F.minv.p, b->x.p, b->ptil.p, b->n_tr, b->tr_dof.p, b->tr_mem.p,
                         b->tr_stride.p, b->tr_base.p, b->tr_sign.p, b->iface_q.p, nullptr};
      if (c->use_sweep) plan_batch_sweep(c, *b, iface_q, tr_base, tr_stride, tr_sign);
      c->batches.push_back(std::move(b));
    }
  }
  CUDA_CHECK(cudaDeviceSynchronize());
  c->is_setup = true;
  lap("batches + sweep schedules");
  // Both star-sweep implementations are exact; keep the one that is faster for THIS configuration.  One star sweep
  // each on the SAME non-zero pseudo-random interface data: the traces of the two implementations must agree to
  // 1e-11 (set-up fails otherwise -- the persistent kernel's plan depends on the grid, the SM share and M, so this is
  // the check that the instantiation actually used on this problem is right), and the times decide.  The times are
  // summed over the ranks, so every rank takes the same decision.  The persistent kernel runs the batches one after
  // the other, the per-level kernels run them concurrently on their own streams, so the winner depends on the mix.
  bool any_sweep = false;
  for (auto &b : c->batches) any_sweep |= b->sweep;
  // A rank without a sweep plan (only constrained or very small subdomains) still takes part in the all-reduce of the
  // times below: the condition of a collective must not depend on what a rank happens to own.
  if (!any_sweep && c->world > 1 && (c->sweep_autotune || c->sweep_crosscheck)) {
    DevBuf<double> tune;
    tune.alloc(4);
    tune.zero(c->stream);
    allreduce_sum(c, tune.p, 2);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
  }
  if (any_sweep && (c->sweep_autotune || c->sweep_crosscheck)) {
    // The comparison runs on the first tune_steps of the longest sweep's time levels (every level walks the same entry lists and the
    // same kernels, so both the check and the time ratio carry over): a full sweep of each implementation, twice,
    // was 1.1 s of the set-up on configs[1].
    auto timed = [&](bool with_sweep) {
      std::vector<bool> keep;
      std::vector<int> keep_nt;
      int nt_max = 1;
      for (auto &b : c->batches) nt_max = std::max(nt_max, b->nt);
      for (auto &b : c->batches) {
        keep.push_back(b->sweep); b->sweep = b->sweep && with_sweep;
        keep_nt.push_back(b->nt);
        // every batch keeps its share of the sweep (the SM shares of concurrent kernels are sized for whole sweeps)
        if (c->tune_steps > 0 && nt_max > c->tune_steps)
          b->nt = std::max(1, int((int64_t(b->nt) * c->tune_steps + nt_max - 1) / nt_max));
        b->sargs.n_steps = b->nt;
      }
      float best = 0;
      for (int rep = 0; rep < 2; ++rep) {  // first pass: graph capture (per-level path), first-launch costs (persistent kernel)
        CUDA_CHECK(cudaEventRecord(c->ev_t0, c->stream));
        run_star_sweeps(c, SWEEP_STAR);
        CUDA_CHECK(cudaEventRecord(c->ev_t1, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        CUDA_CHECK(cudaEventElapsedTime(&best, c->ev_t0, c->ev_t1));
      }
      for (size_t i = 0; i < keep.size(); ++i) {
        Batch &b = *c->batches[i];
        b.sweep = keep[i];
        if (b.nt != keep_nt[i] && b.graph_star) {  // captured for the shortened sweep
          cudaGraphExecDestroy(b.graph_star);
          b.graph_star = nullptr;
        }
        b.nt = keep_nt[i];
        b.sargs.n_steps = b.nt;
      }
      return double(best);
    };
    DevBuf<double> ref_trace, tune;
    ref_trace.alloc(size_t(c->n_fine));
    tune.alloc(4);
    tune.zero(c->stream);
    launch_fill_hash(c->n_fine, 0x5eedull, c->lam_bar.p, c->stream);
    c->tune_ms_sweep = timed(true);
    check_watchdog(c);
    CUDA_CHECK(cudaMemcpyAsync(ref_trace.p, c->trace.p, size_t(c->n_fine) * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    c->tune_ms_levels = timed(false);
    launch_diff_norms(c->n_fine, ref_trace.p, c->trace.p, tune.p, c->stream);
    double h[4] = {0, 0, 0, 0};
    CUDA_CHECK(cudaMemcpyAsync(h, tune.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->tune_rel_diff = h[1] > 0 ? std::sqrt(h[0] / h[1]) : (h[0] > 0 ? 1.0 : 0.0);
    if (!(c->tune_rel_diff <= 1e-11))
      fail(MMMFE_E_NUMERIC, "set-up cross-check: persistent sweep kernel and per-level kernels disagree (relative l2 %.3e)", c->tune_rel_diff);
    h[0] = c->tune_ms_sweep; h[1] = c->tune_ms_levels;
    if (c->world > 1) {
      CUDA_CHECK(cudaMemcpyAsync(tune.p, h, 2 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
      allreduce_sum(c, tune.p, 2);
      CUDA_CHECK(cudaMemcpyAsync(h, tune.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
      CUDA_CHECK(cudaStreamSynchronize(c->stream));
    }
    if (c->sweep_autotune && h[1] < h[0])
      for (auto &b : c->batches) b->sweep = false;
    c->lam_bar.zero(c->stream); c->trace.zero(c->stream);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
  }
  if (prof) std::fprintf(stderr, "[setup profile] set-up: cross-check on %d levels: persistent %.3f ms, per-level %.3f ms, rel diff %.2e\n",
                         c->tune_steps, c->tune_ms_sweep, c->tune_ms_levels, c->tune_rel_diff);
  lap("sweep cross-check / selection");
  setup_basis_operator(c, group_of);
  lap("multiscale basis");
  // NCCL opens its point-to-point channels on first use (measured: 1.5 s on 8 ranks, charged to the bar sweep): do one
  // exchange of zeros now, where set-up time belongs
  if (!c->peers.empty()) {
    exchange_and_combine(c, 1.0, c->r0.p);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
  }
}

// Boundary-entry tables of the persistent kernel: per boundary subtree a table [interior variable][member] -> entry
// (q_of[dof * M + member], -1 = none), and per subtree position the offset of its table (-1 = no entries).
// referenced (optional): set to 1 for every entry some subtree variable refers to.
static void build_sub_q(const Factor &F, int M, const std::vector<int32_t> &q_of, const std::vector<int32_t> &dof_of_canon,
                        std::vector<int32_t> &sub_q, std::vector<int32_t> &bq_pos, std::vector<uint8_t> *referenced) {
  const NdTree &tr = F.tree;
  const int64_t nxe = int64_t(F.nx + 1) * F.ny;
  std::vector<int32_t> bq_off(tr.instances.size(), -1);
  sub_q.clear();
  for (size_t k = 0; k < tr.instances.size(); ++k) {
    const SubtreeInstance &in = tr.instances[k];
    const SubtreeTemplate &T = tr.templates[in.tmpl];
    if (T.flags == 0) continue;
    const size_t base = sub_q.size();
    bool any = false;
    for (int32_t l = 0; l < T.n_int; ++l) {
      const int32_t code = T.var_code[l];
      const int type = code >> 30, dj = (code >> 15) & 0x7fff, di = code & 0x7fff;
      const int64_t canon = type == 0 ? int64_t(in.J0 + dj) * (F.nx + 1) + in.I0 + di : nxe + int64_t(in.J0 + dj) * F.nx + in.I0 + di;
      const int64_t d = dof_of_canon.empty() ? canon : dof_of_canon[canon];
      for (int j = 0; j < M; ++j) {
        const int32_t q = q_of[size_t(d) * M + j];
        sub_q.push_back(q);
        any |= q >= 0;
        if (q >= 0 && referenced) (*referenced)[q] = 1;
      }
    }
    if (any) bq_off[k] = int32_t(base);
    else sub_q.resize(base);
  }
  bq_pos.assign(bq_off.size() + 1, -1);
  for (size_t pos = 0; pos < bq_off.size(); ++pos) bq_pos[pos] = bq_off[F.lay.inst_at_pos[pos]];
  sub_q.push_back(-1);
}

// Entry lists, ring schedule and boundary-entry tables of one batch for the persistent sweep kernel.
static void plan_batch_sweep(mmmfe_ctx *c, Batch &b, const std::vector<int32_t> &iface_q,
                             const std::vector<int64_t> &tr_base, const std::vector<int32_t> &tr_stride,
                             const std::vector<double> &tr_sign) {
  Factor &F = *b.factor;
  if (!F.sweep_ok) return;
  const NdTree &tr = F.tree;
  SweepSchedule S;
  if (build_sweep_schedule(tr, F.lay, b.M, F.sweep_ng, F.sweep_budget, 1.0, S) != nullptr) return;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
  if (sweep_ctas_per_sm(b.M, F.sweep_ct, S.smem_bytes) * sms < F.sweep_n_cta) return;  // grid would not be co-resident
  const int M = b.M;
  // boundary entries: the interface dofs of every member (same enumeration as tr_*); per boundary subtree a table
  // [interior variable][member] -> entry
  const int64_t nxe = int64_t(F.nx + 1) * F.ny;
  std::vector<int32_t> dof_of_canon;
  if (F.dof_of_canon.n) {
    dof_of_canon.resize(F.dof_of_canon.n);
    CUDA_CHECK(cudaMemcpy(dof_of_canon.data(), F.dof_of_canon.p, F.dof_of_canon.n * sizeof(int32_t), cudaMemcpyDeviceToHost));
  }
  std::vector<int32_t> sub_q, bq_pos;
  build_sub_q(F, M, iface_q, dof_of_canon, sub_q, bq_pos, nullptr);
  b.s_inst_bq.upload(bq_pos);
  const size_t nq = tr_base.size();
  std::vector<const double *> data(nq + 1, nullptr);
  std::vector<double *> trace(nq + 1, nullptr);
  std::vector<double> coef(nq + 1, 0.0);
  std::vector<int32_t> stride(nq + 1, 0);
  for (size_t q = 0; q < nq; ++q) {
    data[q] = c->lam_bar.p + tr_base[q];
    trace[q] = c->trace.p + tr_base[q];
    coef[q] = -tr_sign[q];  // rhs_i = -s * lam_bar, src/darcy_vtdd.cc:840-844
    stride[q] = tr_stride[q];
  }
  if (c->sweep_debug_stall) {
    // test hook: the first cross-CTA dependency of the batch can never be satisfied -> the kernel's bounded spin-waits
    // must give up (watchdog flag) and the host must report it instead of hanging or returning garbage
    for (SweepEntry &e : S.entries)
      if (e.n_dep > 0) { e.dep[0] |= ((1 << (31 - SWEEP_DEP_SHIFT)) - 1) << SWEEP_DEP_SHIFT; break; }
  }
  b.s_entries.upload(S.entries); b.s_issue.upload(S.issue);
  b.s_cta_begin.upload(S.cta_begin); b.s_cta_pro.upload(S.cta_prologue);
  S.grp.push_back(0);
  b.s_grp.upload(S.grp);
  b.s_sub_q.upload(sub_q); b.s_bq_stride.upload(stride); b.s_bq_coef.upload(coef);
  b.s_bq_data.upload(data); b.s_bq_trace.upload(trace);
  b.h_iface_q = iface_q; b.h_bq_stride = stride; b.h_bq_trace = trace;
  b.s_z.alloc(size_t(F.lay.z_total + 1) * M); b.s_z.zero();
  b.s_xi.alloc(size_t(F.lay.xi_total + 2) * M); b.s_xi.zero();
  b.s_ptil.alloc(size_t(F.lay.pb_total + 2) * M); b.s_ptil.zero();
  b.s_cnt.alloc(2 * F.lay.nodes.size() + 2); b.s_cnt.zero();
  SweepArgs &a = b.sargs;
  std::memset(&a, 0, sizeof a);
  a.entries = b.s_entries.p; a.issue = b.s_issue.p; a.cta_begin = b.s_cta_begin.p; a.cta_prologue = b.s_cta_pro.p;
  a.grp = b.s_grp.p; a.ng = F.sweep_ng;
  a.Rs = F.Rs.p; a.index = F.sweep_index.p;
  a.tmpl16 = F.tmpl16.p; a.tmpl_off = F.tmpl_off.p; a.dof_of_canon = F.dof_of_canon.p;
  a.x = b.x.p; a.z = b.s_z.p; a.xi = b.s_xi.p; a.ptil = b.s_ptil.p; a.cnt = b.s_cnt.p;
  a.minv = F.minv_uniform ? nullptr : F.minv_c.p; a.minv_const = F.minv_const;
  a.sub_q = b.s_sub_q.p; a.bq_data = b.s_bq_data.p; a.bq_trace = b.s_bq_trace.p; a.bq_stride = b.s_bq_stride.p;
  a.bq_coef = b.s_bq_coef.p;
  a.p_of_cell = F.p_of_cell.p;
  a.kx0 = F.kup_c[0]; a.kx1 = F.kup_c[1]; a.ky0 = F.kup_c[2]; a.ky1 = F.kup_c[3];
  a.cl = F.kpu_c[0]; a.cr = F.kpu_c[1]; a.cb = F.kpu_c[2]; a.ct = F.kpu_c[3];
  a.nx = F.nx; a.ny = F.ny; a.nxe = nxe; a.n_flux = F.nf; a.n_pressure = F.np;
  a.n_steps = b.nt; a.level0 = 0;
  a.tmpl_dbl = S.tmpl_dbl; a.ws_dbl = S.ws_dbl; a.ring_dbl = S.ring_dbl;
  a.pack = F.lay.pack; a.ws_sub_dbl = S.ws_sub_dbl;
  a.inst_bq = b.s_inst_bq.p; a.inst_ij = F.inst_ij.p;
  a.evict_first = c->sweep_evict_first ? 1 : 0;
  // measured (same box, A/B): M = 4 batches of 256^2 / 128^2 grids 94.1 -> 90.2 ms per application with the fence,
  // M = 2 batches of 1024^2 / 512^2 grids 181.0 -> 182.6 ms: on by default for M >= 4 only
  a.pub_fence = c->sweep_pub_fence < 0 ? (M >= 4 ? 1 : 0) : (c->sweep_pub_fence ? 1 : 0);
  a.spin_ns = uint32_t(c->sweep_spin_ns);
  a.poll_window = std::max(4, 2 * F.sweep_ng);
  if (!c->abort_flag) {
    CUDA_CHECK(cudaMalloc(reinterpret_cast<void **>(&c->abort_flag), sizeof(int)));
    CUDA_CHECK(cudaMemset(c->abort_flag, 0, sizeof(int)));
  }
  a.abort_flag = c->abort_flag;
  b.sweep_smem = S.smem_bytes;
  b.sweep = true;
}

// all nt star steps of one batch in one cooperative launch on the interface stream
// The bar problem on the persistent kernel: the boundary-entry tables additionally carry the outer Dirichlet rows of
// every member (data = its fr_vals column, no trace); the interface entries keep their trace slots and have no data
// (lambda = 0).  False when a data row is not an interior variable of a boundary subtree (rows that constrain_matrix
// adds on separators): the per-level kernels run the bar sweep then.
static bool plan_batch_bar(mmmfe_ctx *c, Batch &b) {
  const Factor &F = *b.factor;
  const int M = b.M;
  std::vector<int32_t> q_of = b.h_iface_q;
  const size_t nq = b.h_bq_trace.size() - 1;  // interface entries
  std::vector<const double *> data(nq, nullptr);
  std::vector<double *> trace(b.h_bq_trace.begin(), b.h_bq_trace.begin() + nq);
  std::vector<int32_t> stride(b.h_bq_stride.begin(), b.h_bq_stride.begin() + nq);
  std::vector<double> coef(nq, 0.0);
  for (size_t j = 0; j < b.members.size(); ++j) {
    const Sub &sd = *c->subs[b.members[j]];
    for (int32_t e = 0; e < sd.n_fr; ++e) {
      const int32_t d = sd.fr_dofs_h[e];
      if (d < 0 || d >= F.nf || q_of[size_t(d) * M + j] >= 0) return false;
      q_of[size_t(d) * M + j] = int32_t(data.size());
      data.push_back(sd.fr_vals.p + e); trace.push_back(nullptr); stride.push_back(sd.n_fr); coef.push_back(1.0);
    }
  }
  std::vector<int32_t> dof_of_canon;
  if (F.dof_of_canon.n) {
    dof_of_canon.resize(F.dof_of_canon.n);
    CUDA_CHECK(cudaMemcpy(dof_of_canon.data(), F.dof_of_canon.p, F.dof_of_canon.n * sizeof(int32_t), cudaMemcpyDeviceToHost));
  }
  std::vector<uint8_t> referenced(data.size() + 1, 0);
  std::vector<int32_t> sub_q, bq_pos;
  build_sub_q(F, M, q_of, dof_of_canon, sub_q, bq_pos, &referenced);
  for (size_t q = nq; q < data.size(); ++q) if (!referenced[q]) return false;
  data.push_back(nullptr); trace.push_back(nullptr); stride.push_back(0); coef.push_back(0.0);
  b.sb_sub_q.upload(sub_q); b.sb_inst_bq.upload(bq_pos);
  b.sb_bq_data.upload(data); b.sb_bq_trace.upload(trace); b.sb_bq_stride.upload(stride); b.sb_bq_coef.upload(coef);
  return true;
}

enum SweepKernelMode { SK_STAR, SK_FINAL, SK_BAR };

static void run_sweep_kernel(mmmfe_ctx *c, Batch &b, cudaStream_t st, int mode) {
  const Factor &F = *b.factor;
  SweepArgs a = b.sargs;
  if (mode == SK_BAR) {  // solve_timestep(0, .), src/darcy_vtdd.cc:902-920
    launch_init_ptil_box(F.np, b.M, F.cell_box.p, F.p_of_cell.p, F.minv_uniform ? nullptr : F.minv_c.p, F.minv_const,
                         b.p0_ptr.p, b.src_ptr.p, b.s_ptil.p, st);
    a.sub_q = b.sb_sub_q.p; a.inst_bq = b.sb_inst_bq.p;
    a.bq_data = b.sb_bq_data.p; a.bq_trace = b.sb_bq_trace.p; a.bq_stride = b.sb_bq_stride.p; a.bq_coef = b.sb_bq_coef.p;
    a.has_src = 1; a.src_next = b.src_ptr.p;
    if (c->keep_history) { a.hist = b.hist_ptr.p; a.hist_add = 0; a.hist_stride = int64_t(F.nf) + F.np; }
  } else {
    CUDA_CHECK(cudaMemsetAsync(b.s_ptil.p, 0, b.s_ptil.n * sizeof(double), st));
  }
  CUDA_CHECK(cudaMemsetAsync(b.s_cnt.p, 0, b.s_cnt.n * sizeof(uint32_t), st));
  if (mode == SK_FINAL) {  // solution = star + bar_collection[level], src/darcy_vtdd.cc:944-950
    a.hist = b.hist_ptr.p;
    a.hist_add = 1;
    a.hist_stride = int64_t(F.nf) + F.np;
  }
  const cudaError_t err = launch_sweep(b.M, b.factor->sweep_ct, a, b.factor->sweep_n_cta, b.sweep_smem, st);
  if (err != cudaSuccess) fail(MMMFE_E_CUDA, "sweep kernel launch failed: %s", cudaGetErrorString(err));
}

static void bar_sweep(mmmfe_ctx *c) {
  if (!c->is_setup) fail(MMMFE_E_INVALID, "mmmfe_setup must precede the bar sweep");
  for (auto &bp : c->batches) {
    Batch &b = *bp;
    std::vector<const double *> p0(b.M, nullptr), src(b.M, nullptr);
    std::vector<double *> hist(b.M, nullptr);
    for (size_t j = 0; j < b.members.size(); ++j) {
      Sub &sd = *c->subs[b.members[j]];
      if (!sd.has_bar_data) fail(MMMFE_E_INVALID, "subdomain %d has no bar data", sd.gid);
      p0[j] = sd.p0.p; src[j] = sd.src.p;
      if (c->keep_history) {
        sd.hist.alloc(size_t(sd.nt) * (size_t(sd.nf) + sd.np));
        hist[j] = sd.hist.p;
      }
    }
    b.p0_ptr.upload(p0); b.src_ptr.upload(src); b.hist_ptr.upload(hist);
    if (b.sweep && c->sweep_bar && !b.bar_planned) {
      b.bar_sweep = plan_batch_bar(c, b);
      b.bar_planned = true;
    }
    if (!b.sweep || !c->sweep_bar) b.bar_sweep = false;
  }
  run_star_sweeps(c, SWEEP_BAR);
  // initial residual r = send + recv of the signed bar traces in mortar space (src/darcy_vtdd.cc:1203-1256)
  exchange_and_combine(c, 1.0, c->r0.p);
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  c->bar_done = true;
  c->final_done = false;
}

// local_gmres, src/darcy_vtdd.cc:1139-1521.
//
// The reference decides after every iteration (Givens rotation + stop test on the host).  Here the GPU never waits
// for that decision: the norm h[k+1] stays on the device (the next Krylov vector is scaled by a kernel that reads
// it), the Hessenberg column travels to a pinned ring slot behind an event, and the host enqueues up to
// `gmres_lookahead` iterations AHEAD of the one whose column it is rotating.  When the stop test fires at iteration
// j, the iterations enqueued beyond j are surplus Krylov vectors that are simply not used (lambda is built from the
// first k_counter columns, as in the reference), so iteration counts, residual history and lambda are those of the
// in-order algorithm.  Lookahead 0 is the in-order loop; it is the default whenever one operator application is long
// against a launch (star sweeps), where running ahead would only waste sweeps after convergence.
static void gmres(mmmfe_ctx *c, double tol, int max_iter, mmmfe_gmres_info *info, double *hist, int hist_cap,
                  double *lambda_out) {
  if (!c->bar_done) fail(MMMFE_E_INVALID, "mmmfe_bar_sweep must precede mmmfe_gmres_solve");
  const auto t_begin = std::chrono::steady_clock::now();
  const int64_t len = c->n_iface;
  const int64_t rows = int64_t(max_iter) + 2;
  if (c->q_rows < rows) { c->Q.alloc(size_t(rows) * len); c->q_rows = rows; }
  c->h_dev.alloc(size_t(max_iter) + 4);
  c->norm_dev.alloc(8);
  c->arn_scratch.alloc(arnoldi_scratch_doubles(len, max_iter + 2));
  if (!c->arn_ticket.p) { c->arn_ticket.alloc(4); c->arn_ticket.zero(c->stream); }
  cudaStream_t st = c->stream;
  double *Q = c->Q.p;
  const int lag = c->gmres_lookahead >= 0 ? c->gmres_lookahead : (c->use_basis ? 3 : 0);
  const int L = lag + 1;
  // pinned ring: per slot the Hessenberg column (max_iter + 1), ||w||^2 and the sweep kernel's watchdog flag
  const size_t slot_dbl = size_t(max_iter) + 4;
  if (c->gmres_ring_dbl < slot_dbl * L) {
    if (c->gmres_ring) cudaFreeHost(c->gmres_ring);
    CUDA_CHECK(cudaMallocHost(&c->gmres_ring, slot_dbl * L * sizeof(double)));
    c->gmres_ring_dbl = slot_dbl * L;
  }
  while (int(c->gmres_events.size()) < L) {
    cudaEvent_t e;
    CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c->gmres_events.push_back(e);
  }
  // initial residual r was formed at the end of the bar sweep
  launch_multi_dot(len, c->r0.p, c->r0.p, len, 1, c->h_dev.p, c->arn_scratch.p, c->arn_ticket.p, st);
  allreduce_sum(c, c->h_dev.p, 1);
  double rr = 0.0;
  CUDA_CHECK(cudaMemcpyAsync(&rr, c->h_dev.p, sizeof(double), cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  const double r_norm = std::sqrt(rr);  // :1265-1276
  launch_scale_store(len, c->r0.p, r_norm, Q, st);  // Q[0] = r / r_norm, :1280-1285
  std::vector<double> cs, sn, beta{r_norm};
  std::vector<std::vector<double>> H;
  std::vector<double> res{1.0};
  int k = 0, its = 0, converged = 0;
  CUDA_CHECK(cudaEventRecord(c->ev_t0, st));
  // one Arnoldi step on the device, everything stream-ordered, nothing waits for the host
  auto enqueue = [&](int kk) {
    double *slot = c->gmres_ring + size_t(kk % L) * slot_dbl;
    apply_operator(c, Q + int64_t(kk) * len, c->ap.p);
    launch_multi_dot(len, c->ap.p, Q, len, kk + 1, c->h_dev.p, c->arn_scratch.p, c->arn_ticket.p, st);  // :1416-1422
    allreduce_sum(c, c->h_dev.p, size_t(kk) + 1);                                                        // :1427-1432
    launch_cgs_update(len, c->ap.p, Q, len, kk + 1, c->h_dev.p, c->norm_dev.p, c->arn_scratch.p, c->arn_ticket.p, st);  // :1436-1441
    allreduce_sum(c, c->norm_dev.p, 1);                                                                  // :1449-1454
    CUDA_CHECK(cudaMemcpyAsync(slot, c->h_dev.p, (size_t(kk) + 1) * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaMemcpyAsync(slot + kk + 1, c->norm_dev.p, sizeof(double), cudaMemcpyDeviceToHost, st));
    if (c->abort_flag)
      CUDA_CHECK(cudaMemcpyAsync(slot + kk + 2, const_cast<int *>(c->abort_flag), sizeof(int), cudaMemcpyDeviceToHost, st));
    else
      std::memset(slot + kk + 2, 0, sizeof(double));
    CUDA_CHECK(cudaEventRecord(c->gmres_events[kk % L], st));
    launch_scale_store(len, c->ap.p, 0.0, Q + int64_t(kk + 1) * len, st, c->norm_dev.p);  // :1458-1463
  };
  int enq = 0;  // iterations enqueued so far
  while (k < max_iter) {
    while (enq < max_iter && enq - k <= lag) enqueue(enq++);
    CUDA_CHECK(cudaEventSynchronize(c->gmres_events[k % L]));
    const double *slot = c->gmres_ring + size_t(k % L) * slot_dbl;
    int flag = 0;
    std::memcpy(&flag, slot + k + 2, sizeof(int));
    if (flag) {
      CUDA_CHECK(cudaStreamSynchronize(st));
      check_watchdog(c);
    }
    ++its;  // cg_iteration++, :1363
    std::vector<double> h(slot, slot + k + 1);
    h.push_back(std::sqrt(slot[k + 1]));  // :1455
    // apply_givens_rotation, :1092-1118
    for (int i = 0; i < k; ++i) {
      const double temp = cs[i] * h[i] + sn[i] * h[i + 1];
      h[i + 1] = -sn[i] * h[i] + cs[i] * h[i + 1];
      h[i] = temp;
    }
    double cs_k, sn_k;
    const double v1 = h[k], v2 = h[k + 1];
    if (std::fabs(v1) < 1e-15) { cs_k = 0; sn_k = 1; }  // givens_rotation, :1076-1087
    else { const double t = std::sqrt(v1 * v1 + v2 * v2); cs_k = std::fabs(v1) / t; sn_k = cs_k * v2 / v1; }
    h[k] = cs_k * h[k] + sn_k * h[k + 1];
    h[k + 1] = 0.0;
    cs.push_back(cs_k); sn.push_back(sn_k);
    H.push_back(h);
    beta.push_back(-sn[k] * beta[k]);  // :1474
    beta[k] *= cs[k];                  // :1475
    const double err = std::fabs(beta[k + 1]) / r_norm;  // :1478
    res.push_back(err);
    if (err < tol) { converged = 1; break; }  // :1488-1492, k not incremented (quirk Q1)
    ++k;
  }
  CUDA_CHECK(cudaEventRecord(c->ev_t1, st));
  c->gmres_surplus = enq - its;
  // back_solve over k columns, :1121-1135; y[k] stays 0
  std::vector<double> y(size_t(k) + 1, 0.0);
  for (int i = k - 1; i >= 0; --i) {
    y[i] = beta[i] / H[i][i];
    for (int j = i + 1; j <= k - 1; ++j) y[i] -= H[j][i] * y[j] / H[i][i];
  }
  CUDA_CHECK(cudaMemcpyAsync(c->h_dev.p, y.data(), y.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  launch_axpy_rows(len, Q, len, k + 1, c->h_dev.p, c->lam.p, st);  // :1517-1521
  if (lambda_out) CUDA_CHECK(cudaMemcpyAsync(lambda_out, c->lam.p, size_t(len) * sizeof(double), cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  check_watchdog(c);
  if (info) {
    info->iterations = its; info->converged = converged; info->r_norm = r_norm; info->final_residual = res.back();
    info->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
    float ms = 0;
    CUDA_CHECK(cudaEventElapsedTime(&ms, c->ev_t0, c->ev_t1));
    info->device_ms = ms;
  }
  if (hist) for (int i = 0; i < hist_cap && i < int(res.size()); ++i) hist[i] = res[i];
}

}  // namespace mmmfe

// =================================================================================================================
// C ABI
// =================================================================================================================
#define MMMFE_TRY(ctx, ...)                        \
  try {                                            \
    __VA_ARGS__;                                   \
    return MMMFE_OK;                               \
  } catch (const mmmfe::Error &e) {                \
    if (ctx) (ctx)->err = e.msg;                   \
    return e.code;                                 \
  } catch (const std::exception &e) {              \
    if (ctx) (ctx)->err = e.what();                \
    return MMMFE_E_INVALID;                        \
  }

extern "C" {

const char *mmmfe_version(void) { return "mmmfe-b200 0.1 (sm_100a)"; }

int mmmfe_create(mmmfe_ctx **out, int device) {
  if (!out) return MMMFE_E_INVALID;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return MMMFE_E_CUDA;
  auto *c = new mmmfe_ctx();
  if (device < 0) cudaGetDevice(&device);
  c->device = device;
  c->launch_base = launch_counter();
  if (cudaSetDevice(device) != cudaSuccess) { delete c; return MMMFE_E_CUDA; }
  dev_pool_retain(device);
  *out = c;
  return MMMFE_OK;
}

int mmmfe_destroy(mmmfe_ctx *c) {
  if (!c) return MMMFE_OK;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (auto &b : c->batches) {
    if (b->graph_star) cudaGraphExecDestroy(b->graph_star);
    if (b->graph_step) cudaGraphExecDestroy(b->graph_step);
    if (b->done) cudaEventDestroy(b->done);
    if (b->stream) cudaStreamDestroy(b->stream);
  }
  for (auto &f : c->factors)
    if (f->sweep_stream) { cudaStreamDestroy(f->sweep_stream); f->sweep_stream = nullptr; }
  // the NCCL communicator belongs to the process-wide cache (mmmfe_comm_init), not to the context
  if (c->ev_start) cudaEventDestroy(c->ev_start);
  if (c->ev_t0) cudaEventDestroy(c->ev_t0);
  if (c->ev_t1) cudaEventDestroy(c->ev_t1);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->abort_flag) cudaFree(c->abort_flag);
  if (c->gmres_ring) cudaFreeHost(c->gmres_ring);
  for (cudaEvent_t e : c->gmres_events) cudaEventDestroy(e);
  delete c;
  return MMMFE_OK;
}

const char *mmmfe_last_error(const mmmfe_ctx *c) { return c ? c->err.c_str() : "null context"; }

int mmmfe_set_option(mmmfe_ctx *c, const char *key, double value) {
  if (!c || !key) return MMMFE_E_INVALID;
  const std::string k(key);
  if (k == "keep_history") c->keep_history = value != 0;
  else if (k == "use_graphs") c->use_graphs = value != 0;
  else if (k == "leaf_cells") c->leaf_cells = std::max(1, int(value));
  else if (k == "subtree_cells") c->subtree_cells = std::max(0, int(value));
  else if (k == "use_sweep") c->use_sweep = value != 0;
  else if (k == "sweep_evict_first") c->sweep_evict_first = value != 0;
  else if (k == "sweep_chunk") c->sweep_chunk = std::max(256, int(value)) & ~1;
  else if (k == "sweep_compute_threads") {
    if (value != 64 && value != 128 && value != 256 && value != 512) { c->err = "sweep_compute_threads must be 64, 128, 256 or 512"; return MMMFE_E_INVALID; }
    c->sweep_ct = int(value);
  }
  else if (k == "sweep_ctas_per_sm") c->sweep_ctas_per_sm = std::max(1, int(value));
  else if (k == "sweep_spin_ns") c->sweep_spin_ns = std::max(0, int(value));
  else if (k == "sweep_autotune") c->sweep_autotune = value != 0;
  else if (k == "tune_steps") c->tune_steps = int(value);
  else if (k == "sweep_final") c->sweep_final = value != 0;
  else if (k == "sweep_bar") c->sweep_bar = value != 0;
  else if (k == "sweep_debug_stall") c->sweep_debug_stall = value != 0;
  else if (k == "sweep_pub_fence") c->sweep_pub_fence = int(value);
  else if (k == "gmres_lookahead") c->gmres_lookahead = int(value);
  else if (k == "sweep_crosscheck") c->sweep_crosscheck = value != 0;
  else if (k == "operator") {
    if (value != 0 && value != 1 && value != 2) { c->err = "operator must be 0 (sweeps), 1 (multiscale basis) or 2 (auto)"; return MMMFE_E_INVALID; }
    c->operator_mode = int(value);
  }
  else if (k == "mortar_time_slabs") c->mortar_slabs = std::max(0, int(value));
  else if (k == "basis_columns") c->basis_cols_max = std::max(8, int(value));
  else if (k == "basis_auto_columns") c->basis_auto_columns = std::max(0, int(value));
  else if (k == "basis_share") c->basis_share = value != 0;
  else if (k == "basis_drop_tol") c->basis_drop_tol = std::max(0.0, value);
  else if (k == "sweep_groups") {
    if (value != 1 && value != 2 && value != 4 && value != 8) { c->err = "sweep_groups must be 1, 2, 4 or 8"; return MMMFE_E_INVALID; }
    c->sweep_groups = int(value);
  }
  else if (k == "sweep_min_fill") c->sweep_min_fill = value;
  else if (k == "sweep_micro_tile") c->sweep_mt = std::max(64, int(value));
  else if (k == "sweep_share_blobs") c->sweep_share_blobs = value != 0;
  else if (k == "sweep_concurrent") c->sweep_concurrent = value != 0;
  else if (k == "sweep_share_exponent") c->sweep_share_exponent = std::min(2.0, std::max(0.1, value));
  else if (k == "sweep_pack") {
    if (value != 1 && value != 2 && value != 4 && value != 8) { c->err = "sweep_pack must be 1, 2, 4 or 8"; return MMMFE_E_INVALID; }
    c->sweep_pack = int(value);
  }
  else { c->err = "unknown option " + k; return MMMFE_E_INVALID; }
  return MMMFE_OK;
}

int mmmfe_comm_unique_id_size(void) {
#ifdef MMMFE_WITH_NCCL
  return int(sizeof(ncclUniqueId));
#else
  return 0;
#endif
}

int mmmfe_comm_unique_id(void *id_out) {
#ifdef MMMFE_WITH_NCCL
  ncclUniqueId id;
  if (!nccl().ok || nccl().GetUniqueId(&id) != ncclSuccess) return MMMFE_E_COMM;
  std::memcpy(id_out, &id, sizeof id);
  return MMMFE_OK;
#else
  (void)id_out;
  return MMMFE_E_UNSUPPORTED;
#endif
}

int mmmfe_comm_init(mmmfe_ctx *c, int rank, int world, const void *id) {
  if (!c || world < 1 || rank < 0 || rank >= world) return MMMFE_E_INVALID;
  c->rank = rank; c->world = world;
  if (world == 1) return MMMFE_OK;
#ifdef MMMFE_WITH_NCCL
  if (!id) { c->err = "mmmfe_comm_init: the unique id is required when world > 1"; return MMMFE_E_INVALID; }
  ncclUniqueId uid;
  std::memcpy(&uid, id, sizeof uid);
  cudaSetDevice(c->device);
  if (!nccl().ok) { c->err = "libnccl.so.2 could not be loaded"; return MMMFE_E_COMM; }
  // A unique id can initialise ONE communicator, but the front-end creates a context per refinement cycle
  // (src/darcy_vtdd.cc:2079: everything is rebuilt per cycle) with the id it was launched with: communicators are
  // cached per (id, rank, world, device) for the life of the process and shared by the contexts.
  {
    std::lock_guard<std::mutex> lock(comm_cache_mutex());
    std::string key(reinterpret_cast<const char *>(&uid), sizeof uid);
    key += "/" + std::to_string(rank) + "/" + std::to_string(world) + "/" + std::to_string(c->device);
    auto &cache = comm_cache();
    auto it = cache.find(key);
    if (it == cache.end()) {
      ncclComm_t comm = nullptr;
      if (nccl().CommInitRank(&comm, world, uid, rank) != ncclSuccess) { c->err = "ncclCommInitRank failed"; return MMMFE_E_COMM; }
      it = cache.emplace(key, comm).first;
    }
    c->comm = it->second;
  }
  return MMMFE_OK;
#else
  (void)id;
  c->err = "built without NCCL";
  return MMMFE_E_UNSUPPORTED;
#endif
}

int mmmfe_set_owners(mmmfe_ctx *c, int32_t n_global, const int32_t *owner_rank) {
  if (!c || n_global <= 0 || !owner_rank) return MMMFE_E_INVALID;
  c->owner.assign(owner_rank, owner_rank + n_global);
  return MMMFE_OK;
}

int mmmfe_comm_warmup(mmmfe_ctx *c, int32_t n_peers, const int32_t *peers) {
  if (!c || n_peers < 0 || (n_peers > 0 && !peers)) return MMMFE_E_INVALID;
  if (c->world <= 1) return MMMFE_OK;
#ifdef MMMFE_WITH_NCCL
  MMMFE_TRY(c, {
    if (!c->comm) fail(MMMFE_E_INVALID, "mmmfe_comm_init first");
    CUDA_CHECK(cudaSetDevice(c->device));
    if (!c->stream) CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    DevBuf<double> buf;
    buf.alloc(size_t(2 * n_peers) + 2);
    buf.zero(c->stream);
    nccl().GroupStart();
    for (int32_t k = 0; k < n_peers; ++k) {
      if (peers[k] < 0 || peers[k] >= c->world || peers[k] == c->rank) continue;
      nccl().Send(buf.p + 2 * k, 1, ncclDouble, peers[k], c->comm, c->stream);
      nccl().Recv(buf.p + 2 * k + 1, 1, ncclDouble, peers[k], c->comm, c->stream);
    }
    if (nccl().GroupEnd() != ncclSuccess) fail(MMMFE_E_COMM, "NCCL warm-up exchange failed");
    allreduce_sum(c, buf.p + 2 * n_peers, 1);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
  });
#else
  (void)peers;
  return MMMFE_E_UNSUPPORTED;
#endif
}

int mmmfe_comm_allreduce_sum(mmmfe_ctx *c, double *inout, int32_t n) {
  if (!c || !inout || n < 0) return MMMFE_E_INVALID;
  if (c->world <= 1 || n == 0) return MMMFE_OK;
  MMMFE_TRY(c, {
    if (!c->stream) CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    if (c->ar_buf.n < size_t(n)) c->ar_buf.alloc(std::max(size_t(n) * 2, size_t(1024)));  // kept: no allocation per call
    CUDA_CHECK(cudaMemcpyAsync(c->ar_buf.p, inout, size_t(n) * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    allreduce_sum(c, c->ar_buf.p, size_t(n));
    CUDA_CHECK(cudaMemcpyAsync(inout, c->ar_buf.p, size_t(n) * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
  });
}

int mmmfe_measure_fp64_tensor_tflops(int device, double *tflops) {
  if (!tflops) return MMMFE_E_INVALID;
  if (device >= 0 && cudaSetDevice(device) != cudaSuccess) return MMMFE_E_CUDA;
  *tflops = measure_dmma_tflops(nullptr);
  return *tflops > 0 ? MMMFE_OK : MMMFE_E_CUDA;
}

int64_t mmmfe_kernel_launches(const mmmfe_ctx *c) { return c ? launch_counter() - c->launch_base : 0; }

int mmmfe_add_subdomain(mmmfe_ctx *c, const mmmfe_subdomain *d, int32_t *local_index) {
  if (!c || !d) return MMMFE_E_INVALID;
  MMMFE_TRY(c, {
    if (c->is_setup) fail(MMMFE_E_INVALID, "subdomains must be added before mmmfe_setup");
    if (d->nx < 1 || d->ny < 1 || d->nt < 1 || !(d->dt > 0)) fail(MMMFE_E_INVALID, "bad mesh pattern");
    auto s = std::make_unique<Sub>();
    s->gid = d->global_id; s->nx = d->nx; s->ny = d->ny; s->nt = d->nt; s->dt = d->dt;
    s->nf = d->n_flux; s->np = d->n_pressure;
    if (int64_t(s->np) != int64_t(d->nx) * d->ny) fail(MMMFE_E_UNSUPPORTED, "only DGQ0 pressures (one per cell) are on the path");
    s->matrix.from(d->matrix);
    if (d->canon_of_dof) s->canon.assign(d->canon_of_dof, d->canon_of_dof + d->n_flux);
    for (int k = 0; k < 4; ++k) {
      s->neighbor[k] = d->neighbor[k]; s->n_side[k] = d->n_side[k]; s->mortar_len[k] = d->mortar_len[k];
      if (d->neighbor[k] >= 0) {
        if (!d->side_dofs[k]) fail(MMMFE_E_INVALID, "side_dofs missing for side %d", k);
        s->side_dofs[k].assign(d->side_dofs[k], d->side_dofs[k] + d->n_side[k]);
        s->f2c[k].from(d->f2c[k]);
        s->c2f[k].from(d->c2f[k]);
      } else if (d->side_dofs[k] && d->f2c[k].row_ptr && d->c2f[k].row_ptr && d->n_side[k] > 0 && d->mortar_len[k] > 0) {
        // tables of an outer side: never applied to this subdomain, only used to solve basis columns for other ranks
        s->side_dofs[k].assign(d->side_dofs[k], d->side_dofs[k] + d->n_side[k]);
        s->f2c[k].from(d->f2c[k]);
        s->c2f[k].from(d->c2f[k]);
      }
    }
    if (local_index) *local_index = int32_t(c->subs.size());
    c->subs.push_back(std::move(s));
  });
}

int mmmfe_setup(mmmfe_ctx *c) {
  if (!c) return MMMFE_E_INVALID;
  MMMFE_TRY(c, setup(c));
}

int64_t mmmfe_interface_length(const mmmfe_ctx *c) { return c ? c->n_iface : -1; }

int64_t mmmfe_interface_offset(const mmmfe_ctx *c, int32_t li, int32_t side) {
  if (!c || li < 0 || li >= int32_t(c->subs.size()) || side < 0 || side > 3) return -1;
  return c->subs[li]->iface_off[side];
}

int32_t mmmfe_num_factor_groups(const mmmfe_ctx *c) { return c ? int32_t(c->factors.size()) : -1; }

int64_t mmmfe_factor_values(const mmmfe_ctx *c) {
  int64_t v = 0;
  if (c) for (auto &f : c->factors) v += f->tree.r_total;
  return v;
}

int64_t mmmfe_step_bytes(const mmmfe_ctx *c) {
  // one star step of every local subdomain: each batch streams its factor once (forward G, backward [F11^-1|G])
  // plus 8 * (2 n + 2 sum n_side) bytes of vector traffic per member
  int64_t v = 0;
  if (!c) return 0;
  for (auto &b : c->batches) {
    v += b->factor->solve_bytes;
    v += int64_t(b->members.size()) * 16 * (int64_t(b->factor->nf) + b->factor->np);
    v += 16 * int64_t(b->n_tr);
  }
  return v;
}

double mmmfe_factor_flops(const mmmfe_ctx *c) {
  double v = 0;
  if (c) for (auto &f : c->factors) v += f->tree.factor_flops;
  return v;
}

double mmmfe_factor_seconds(const mmmfe_ctx *c) { return c ? c->factor_seconds : 0.0; }

int mmmfe_bar_set_data(mmmfe_ctx *c, int32_t li, const double *p0, const double *src, int32_t n_fr,
                       const int32_t *fr_dofs, const double *fr_vals) {
  if (!c || li < 0 || li >= int32_t(c->subs.size()) || !src) return MMMFE_E_INVALID;
  MMMFE_TRY(c, {
    Sub &sd = *c->subs[li];
    CUDA_CHECK(cudaSetDevice(c->device));
    if (p0) sd.p0.upload(p0, size_t(sd.np)); else { sd.p0.alloc(size_t(sd.np)); sd.p0.zero(); }
    sd.src.upload(src, size_t(sd.nt) * sd.np);
    sd.n_fr = n_fr;
    sd.fr_dofs.upload(fr_dofs, size_t(n_fr));
    sd.fr_dofs_h.assign(fr_dofs, fr_dofs + n_fr);
    sd.fr_vals.upload(fr_vals, size_t(n_fr) * sd.nt);
    sd.has_bar_data = true;
    if (sd.batch >= 0) c->batches[sd.batch]->bar_planned = false;
  });
}

int mmmfe_bar_data_buffers(mmmfe_ctx *c, int32_t li, int32_t n_fr, const int32_t *fr_dofs, double **d_p0, double **d_src,
                           double **d_fr_vals) {
  if (!c || li < 0 || li >= int32_t(c->subs.size()) || n_fr < 0 || (n_fr > 0 && !fr_dofs) || !d_p0 || !d_src || !d_fr_vals)
    return MMMFE_E_INVALID;
  MMMFE_TRY(c, {
    Sub &sd = *c->subs[li];
    CUDA_CHECK(cudaSetDevice(c->device));
    sd.p0.alloc(size_t(sd.np)); sd.p0.zero();
    sd.src.alloc(size_t(sd.nt) * sd.np); sd.src.zero();
    sd.n_fr = n_fr;
    sd.fr_dofs.upload(fr_dofs, size_t(n_fr));
    sd.fr_dofs_h.assign(fr_dofs, fr_dofs + n_fr);
    if (sd.batch >= 0) c->batches[sd.batch]->bar_planned = false;
    sd.fr_vals.alloc(size_t(n_fr) * sd.nt); sd.fr_vals.zero();
    CUDA_CHECK(cudaDeviceSynchronize());
    sd.has_bar_data = true;
    *d_p0 = sd.p0.p; *d_src = sd.src.p; *d_fr_vals = sd.fr_vals.p;
  });
}

int mmmfe_bar_sweep(mmmfe_ctx *c) {
  if (!c) return MMMFE_E_INVALID;
  MMMFE_TRY(c, bar_sweep(c));
}

int mmmfe_apply_interface_operator(mmmfe_ctx *c, const double *q, double *ap) {
  if (!c || !q || !ap) return MMMFE_E_INVALID;
  MMMFE_TRY(c, {
    if (!c->is_setup) fail(MMMFE_E_INVALID, "mmmfe_setup first");
    const size_t bytes = size_t(c->n_iface) * sizeof(double);
    CUDA_CHECK(cudaMemcpyAsync(c->qin.p, q, bytes, cudaMemcpyHostToDevice, c->stream));
    apply_operator(c, c->qin.p, c->ap.p);
    CUDA_CHECK(cudaMemcpyAsync(ap, c->ap.p, bytes, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    check_watchdog(c);
  });
}

int mmmfe_get_initial_residual(mmmfe_ctx *c, double *r) {
  if (!c || !r) return MMMFE_E_INVALID;
  MMMFE_TRY(c, {
    if (!c->bar_done) fail(MMMFE_E_INVALID, "bar sweep first");
    CUDA_CHECK(cudaMemcpy(r, c->r0.p, size_t(c->n_iface) * sizeof(double), cudaMemcpyDeviceToHost));
  });
}

void *mmmfe_alloc_pinned(int64_t bytes) {
  void *p = nullptr;
  if (bytes <= 0 || cudaMallocHost(&p, size_t(bytes)) != cudaSuccess) return nullptr;
  return p;
}

void mmmfe_free_pinned(void *p) {
  if (p) cudaFreeHost(p);
}

int mmmfe_apply_interface_operator_device(mmmfe_ctx *c, int32_t repeat, double *ms) {
  if (!c || repeat < 1) return MMMFE_E_INVALID;
  MMMFE_TRY(c, {
    if (!c->is_setup) fail(MMMFE_E_INVALID, "mmmfe_setup first");
    CUDA_CHECK(cudaEventRecord(c->ev_t0, c->stream));
    for (int i = 0; i < repeat; ++i) apply_operator(c, c->qin.p, c->ap.p);
    CUDA_CHECK(cudaEventRecord(c->ev_t1, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    check_watchdog(c);
    float t = 0;
    CUDA_CHECK(cudaEventElapsedTime(&t, c->ev_t0, c->ev_t1));
    if (ms) *ms = double(t) / repeat;
  });
}

int mmmfe_gmres_solve(mmmfe_ctx *c, double tol, int32_t max_iter, mmmfe_gmres_info *info, double *hist,
                      int32_t hist_cap, double *lambda_out) {
  if (!c || max_iter < 1) return MMMFE_E_INVALID;
  MMMFE_TRY(c, gmres(c, tol, max_iter, info, hist, hist_cap, lambda_out));
}

int mmmfe_final_sweep(mmmfe_ctx *c) {
  if (!c) return MMMFE_E_INVALID;
  MMMFE_TRY(c, {
    if (!c->bar_done) fail(MMMFE_E_INVALID, "bar sweep first");
    if (c->final_done) fail(MMMFE_E_INVALID, "final sweep already done for this bar sweep");
    launch_spmv(c->n_fine, c->c2f_all.rp.p, c->c2f_all.col.p, c->c2f_all.val.p, c->lam.p, c->lam_bar.p, c->stream);
    run_star_sweeps(c, SWEEP_FINAL);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    check_watchdog(c);
    c->final_done = true;
  });
}

int mmmfe_get_solution(mmmfe_ctx *c, int32_t li, double *solution) {
  if (!c || li < 0 || li >= int32_t(c->subs.size()) || !solution) return MMMFE_E_INVALID;
  MMMFE_TRY(c, {
    Sub &sd = *c->subs[li];
    if (!sd.hist.p) fail(MMMFE_E_INVALID, "no stored solution (keep_history = 0 or no sweep yet)");
    CUDA_CHECK(cudaMemcpy(solution, sd.hist.p, sd.hist.n * sizeof(double), cudaMemcpyDeviceToHost));
  });
}

int mmmfe_solution_device(mmmfe_ctx *c, int32_t li, double **d_solution) {
  if (!c || li < 0 || li >= int32_t(c->subs.size()) || !d_solution) return MMMFE_E_INVALID;
  MMMFE_TRY(c, {
    Sub &sd = *c->subs[li];
    if (!sd.hist.p) fail(MMMFE_E_INVALID, "no stored solution (keep_history = 0 or no sweep yet)");
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    *d_solution = sd.hist.p;
  });
}

int mmmfe_get_star_trace(mmmfe_ctx *c, int32_t li, int32_t side, double *trace) {
  if (!c || li < 0 || li >= int32_t(c->subs.size()) || side < 0 || side > 3 || !trace) return MMMFE_E_INVALID;
  MMMFE_TRY(c, {
    Sub &sd = *c->subs[li];
    if (sd.fine_off[side] < 0) fail(MMMFE_E_INVALID, "side %d has no neighbour", side);
    CUDA_CHECK(cudaMemcpy(trace, c->trace.p + sd.fine_off[side], size_t(sd.nt) * sd.n_side[side] * sizeof(double),
                          cudaMemcpyDeviceToHost));
  });
}

int mmmfe_solve_saddle(mmmfe_ctx *c, int32_t li, const double *rhs, double *x) {
  if (!c || li < 0 || li >= int32_t(c->subs.size()) || !rhs || !x) return MMMFE_E_INVALID;
  MMMFE_TRY(c, {
    if (!c->is_setup) fail(MMMFE_E_INVALID, "mmmfe_setup first");
    Sub &sd = *c->subs[li];
    const Factor &F = *c->batches[sd.batch]->factor;
    cudaStream_t st = c->stream;
    DevBuf<double> d_rhs, xv, z, pt, out;
    d_rhs.upload(rhs, size_t(F.nf) + F.np);
    xv.alloc(size_t(F.nf)); z.alloc(size_t(F.tree.z_total)); pt.alloc(size_t(F.np)); out.alloc(size_t(F.nf) + F.np);
    DevBuf<int32_t> dofs;
    std::vector<int32_t> all(F.nf);
    for (int32_t i = 0; i < F.nf; ++i) all[i] = i;
    dofs.upload(all);
    std::vector<const double *> srcp{d_rhs.p + F.nf};
    std::vector<double *> histp{out.p};
    DevBuf<const double *> d_srcp;
    DevBuf<double *> d_histp;
    d_srcp.upload(srcp); d_histp.upload(histp);
    StepArgs a{1, F.nf, F.np, F.kup.rp.p, F.kup.col.p, F.kup.val.p, F.kpu.rp.p, F.kpu.col.p, F.kpu.val.p,
               F.minv.p, xv.p, pt.p, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    launch_init_ptil(a, nullptr, d_srcp.p, st);
    launch_build_rhs(a, nullptr, 0, st);
    launch_add_sparse(xv.p, 1, 0, dofs.p, d_rhs.p, F.nf, st);
    run_solve(F, 1, xv.p, z.p, st);
    launch_update(a, nullptr, 0, nullptr, F.np, d_histp.p, int64_t(F.nf) + F.np, 0, st);
    CUDA_CHECK(cudaMemcpyAsync(x, out.p, out.n * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
  });
}

int mmmfe_profile_step(mmmfe_ctx *c, double ms[MMMFE_N_KERNEL_CLASSES], double bytes[MMMFE_N_KERNEL_CLASSES],
                       int64_t launches[MMMFE_N_KERNEL_CLASSES]) {
  if (!c || !ms || !bytes || !launches) return MMMFE_E_INVALID;
  MMMFE_TRY(c, {
    if (!c->is_setup) fail(MMMFE_E_INVALID, "mmmfe_setup first");
    for (int k = 0; k < MMMFE_N_KERNEL_CLASSES; ++k) { ms[k] = 0; bytes[k] = 0; launches[k] = 0; }
    cudaStream_t st = c->stream;
    CUDA_CHECK(cudaStreamSynchronize(st));
    std::vector<cudaEvent_t> ev;
    std::vector<int> cls;
    auto mark = [&](int k) {
      cudaEvent_t e;
      CUDA_CHECK(cudaEventCreate(&e));
      CUDA_CHECK(cudaEventRecord(e, st));
      ev.push_back(e);
      cls.push_back(k);
    };
    for (auto &bp : c->batches) {
      Batch &b = *bp;
      const Factor &F = *b.factor;
      const int M = b.M;
      SolveArgs a{F.fronts.p, F.R.p, F.idx_user.p, F.src.p, b.z.p, b.x.p};
      const double vec = 8.0 * M;
      mark(-1);
      launch_build_rhs(b.step, c->lam_bar.p, 0, st);
      mark(0); bytes[0] += vec * (F.nf + 2.0 * F.nf) + 12.0 * 2 * F.nf; launches[0] += 1;
      const SubtreeArgs sa = subtree_args(F, b.x.p, b.z.p);
      if (F.n_inst) { launch_subtree_forward(M, sa, F.n_inst, F.smem_fwd[M], st); mark(1); bytes[1] += F.bytes_sub_fwd + vec * F.nf; launches[1] += 1; }
      for (const LevelWork &lw : F.levels)
        if (lw.n_fwd) { launch_forward_large(M, a, lw.fwd.p, lw.n_fwd, st); mark(2); bytes[2] += lw.bytes_fwd; launches[2] += 1; }
      for (auto it = F.levels.rbegin(); it != F.levels.rend(); ++it)
        if (it->n_bwd) { launch_backward_large(M, a, it->bwd.p, it->n_bwd, it->f_max, st); mark(3); bytes[3] += it->bytes_bwd; launches[3] += 1; }
      if (F.n_inst) { launch_subtree_backward(M, sa, F.n_inst, F.smem_bwd[M], st); mark(4); bytes[4] += F.bytes_sub_bwd + 2 * vec * F.nf; launches[4] += 1; }
      launch_update(b.step, c->trace.p, 0, nullptr, F.np, nullptr, int64_t(F.nf) + F.np, 0, st);
      mark(5); bytes[5] += vec * (4.0 * F.np + 2.0 * F.np) + 12.0 * 4 * F.np; launches[5] += 1;
    }
    CUDA_CHECK(cudaStreamSynchronize(st));
    for (size_t i = 1; i < ev.size(); ++i)
      if (cls[i] >= 0) {
        float t = 0;
        CUDA_CHECK(cudaEventElapsedTime(&t, ev[i - 1], ev[i]));
        ms[cls[i]] += t;
      }
    for (auto e : ev) cudaEventDestroy(e);
  });
}

int64_t mmmfe_stat(const mmmfe_ctx *c, const char *key) {
  if (!c || !key) return -1;
  const std::string k(key);
  int64_t v = 0;
  if (k == "batches") return int64_t(c->batches.size());
  if (k == "sweep_batches") { for (auto &b : c->batches) v += b->sweep; return v; }
  if (k == "sweep_ctas") { for (auto &b : c->batches) if (b->sweep) v = std::max<int64_t>(v, b->factor->sweep_n_cta); return v; }
  if (k == "sweep_smem_bytes") { for (auto &b : c->batches) if (b->sweep) v = std::max<int64_t>(v, int64_t(b->sweep_smem)); return v; }
  if (k == "sweep_entries") { for (auto &b : c->batches) if (b->sweep) v += int64_t(b->s_entries.n); return v; }
  if (k == "sweep_shared_blob_pairs") { for (auto &f : c->factors) if (f->sweep_ok) v += f->shared_blob_pairs; return v; }
  if (k == "sweep_factor_bytes") { for (auto &f : c->factors) if (f->sweep_ok) v += int64_t(f->lay.bytes_per_step); return v; }
  if (k == "sweep_chunk_used") { for (auto &f : c->factors) if (f->sweep_ok) v = std::max<int64_t>(v, f->lay.chunk_dbl); return v; }
  if (k == "sweep_ring_doubles") { for (auto &b : c->batches) if (b->sweep) v = std::max<int64_t>(v, b->sargs.ring_dbl); return v; }
  if (k == "rt0_factors") { for (auto &f : c->factors) v += f->rt0; return v; }
  if (k == "sweep_planned") { for (auto &b : c->batches) v += b->s_entries.n > 0; return v; }
  if (k == "basis_operator") return c->use_basis ? 1 : 0;
  if (k == "basis_us") return int64_t(c->basis_seconds * 1e6);
  if (k == "basis_columns") return c->basis_columns;
  if (k == "basis_mflop") return int64_t(c->basis_flops / 1e6);
  if (k == "basis_mbytes") return int64_t(c->basis_bytes / 1e6);
  if (k == "basis_shared_by") { for (auto &f : c->factors) if (f->basis) v = std::max<int64_t>(v, f->basis->shared_by); return v; }
  if (k == "basis_steps") { for (auto &f : c->factors) if (f->basis) v += f->basis->steps; return v; }
  if (k == "basis_delays") { for (auto &f : c->factors) if (f->basis) v = std::max<int64_t>(v, f->basis->d_used); return v; }
  if (k == "basis_passes") { for (auto &f : c->factors) if (f->basis) v += f->basis->n_pass; return v; }
  if (k == "basis_mw") { for (auto &f : c->factors) if (f->basis) v = std::max<int64_t>(v, f->basis->MW); return v; }
  if (k == "basis_values") { for (auto &f : c->factors) if (f->basis) v += int64_t(f->basis->T.n); return v; }
  if (k == "toeplitz_tiles") return c->n_toep_tiles;
  if (k == "toeplitz_mflop") return int64_t(c->toep_flops / 1e6);
  if (k == "toeplitz_mbytes") return int64_t(c->toep_bytes / 1e6);
  if (k == "gmres_surplus") return c->gmres_surplus;
  if (k == "bar_sweep_batches") { int64_t n = 0; for (auto &b : c->batches) n += b->sweep && b->bar_sweep; return n; }
  if (k == "tune_us_sweep") return int64_t(c->tune_ms_sweep * 1e3);
  if (k == "tune_us_levels") return int64_t(c->tune_ms_levels * 1e3);
  // relative l2 difference of the two star-sweep implementations at set-up, in units of 1e-18 (-1: not run)
  if (k == "tune_rel_diff_e18") return c->tune_rel_diff < 0 ? -1 : int64_t(std::min(c->tune_rel_diff * 1e18, 9e18));
  return -1;
}

int mmmfe_debug_sweep_trace(mmmfe_ctx *c, int32_t batch, int64_t cap, int64_t *n_entries, int32_t *cta_begin_out,
                            int32_t *kind_height, int64_t *stamps) {
  if (!c || !n_entries) return MMMFE_E_INVALID;
  MMMFE_TRY(c, {
    int32_t k = 0;
    for (auto &bp : c->batches) {
      Batch &b = *bp;
      if (!b.sweep) continue;
      if (k++ != batch) continue;
      const int32_t n_cta = b.factor->sweep_n_cta;
      const int64_t ne = int64_t(b.s_entries.n);
      *n_entries = ne;
      if (!stamps || cap < ne) return MMMFE_OK;
      DevBuf<long long> prof, trace;
      prof.alloc(size_t(n_cta) * 24); prof.zero(c->stream);
      trace.alloc(size_t(ne) * 3); trace.zero(c->stream);
      SweepArgs a = b.sargs;
      a.prof = std::getenv("MMMFE_TRACE_WITH_PROF") ? prof.p : nullptr;  // default: stamps only, undisturbed timing
      a.trace = trace.p;
      a.trace_step = std::min(std::max(1, b.nt / 2), b.nt - 1);
      CUDA_CHECK(cudaMemsetAsync(b.s_ptil.p, 0, b.s_ptil.n * sizeof(double), c->stream));
      CUDA_CHECK(cudaMemsetAsync(b.s_cnt.p, 0, b.s_cnt.n * sizeof(uint32_t), c->stream));
      const cudaError_t err = launch_sweep(b.M, b.factor->sweep_ct, a, n_cta, b.sweep_smem, c->stream);
      if (err != cudaSuccess) fail(MMMFE_E_CUDA, "sweep kernel launch failed: %s", cudaGetErrorString(err));
      CUDA_CHECK(cudaStreamSynchronize(c->stream));
      check_watchdog(c);
      CUDA_CHECK(cudaMemcpy(stamps, trace.p, size_t(ne) * 3 * sizeof(long long), cudaMemcpyDeviceToHost));
      std::vector<SweepEntry> he(static_cast<size_t>(ne));
      CUDA_CHECK(cudaMemcpy(he.data(), b.s_entries.p, size_t(ne) * sizeof(SweepEntry), cudaMemcpyDeviceToHost));
      for (int64_t i = 0; i < ne; ++i) {
        kind_height[2 * i] = he[i].kind;
        const bool top = he[i].kind == SW_TOP_FWD || he[i].kind == SW_TOP_BWD;
        kind_height[2 * i + 1] = top ? he[i].a7 : 0;
      }
      CUDA_CHECK(cudaMemcpy(cta_begin_out, b.s_cta_begin.p, size_t(n_cta + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost));
      return MMMFE_OK;
    }
    fail(MMMFE_E_INVALID, "no such sweep batch");
  });
}

int mmmfe_profile_sweep_cycles(mmmfe_ctx *c, int32_t batch, double out[24]) {
  if (!c || !out) return MMMFE_E_INVALID;
  MMMFE_TRY(c, {
    if (!c->is_setup) fail(MMMFE_E_INVALID, "mmmfe_setup first");
    int32_t k = 0;
    for (auto &bp : c->batches) {
      Batch &b = *bp;
      if (!b.sweep) continue;
      if (k++ != batch) continue;
      const int32_t n_cta = b.factor->sweep_n_cta;
      DevBuf<long long> prof;
      prof.alloc(size_t(n_cta) * 24);
      prof.zero(c->stream);
      SweepArgs a = b.sargs;
      a.prof = prof.p;
      CUDA_CHECK(cudaMemsetAsync(b.s_ptil.p, 0, b.s_ptil.n * sizeof(double), c->stream));
      CUDA_CHECK(cudaMemsetAsync(b.s_cnt.p, 0, b.s_cnt.n * sizeof(uint32_t), c->stream));
      const cudaError_t err = launch_sweep(b.M, b.factor->sweep_ct, a, n_cta, b.sweep_smem, c->stream);
      if (err != cudaSuccess) fail(MMMFE_E_CUDA, "sweep kernel launch failed: %s", cudaGetErrorString(err));
      CUDA_CHECK(cudaStreamSynchronize(c->stream));
      check_watchdog(c);
      std::vector<long long> h(size_t(n_cta) * 24);
      CUDA_CHECK(cudaMemcpy(h.data(), prof.p, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
      for (int t = 0; t < 24; ++t) out[t] = 0;
      for (int32_t cta = 0; cta < n_cta; ++cta)
        for (int t = 0; t < 24; ++t) out[t] += double(h[size_t(cta) * 24 + t]) / n_cta;
      if (const char *dump = std::getenv("MMMFE_PROF_DUMP")) {  // per-CTA counters (slot 15: SM id)
        if (FILE *f = std::fopen((std::string(dump) + "." + std::to_string(batch)).c_str(), "w")) {
          for (int32_t cta = 0; cta < n_cta; ++cta) {
            for (int t = 0; t < 24; ++t) std::fprintf(f, "%lld ", h[size_t(cta) * 24 + t]);
            std::fprintf(f, "\n");
          }
          std::fclose(f);
        }
      }
      return MMMFE_OK;
    }
    fail(MMMFE_E_INVALID, "no such sweep batch");
  });
}

int mmmfe_profile_sweeps(mmmfe_ctx *c, int32_t cap, double *ms, double *bytes_per_step, int32_t *n_steps,
                         int32_t *n_out) {
  if (!c || !ms || !bytes_per_step || !n_steps || !n_out) return MMMFE_E_INVALID;
  MMMFE_TRY(c, {
    if (!c->is_setup) fail(MMMFE_E_INVALID, "mmmfe_setup first");
    int32_t k = 0;
    for (auto &bp : c->batches) {
      Batch &b = *bp;
      if (!b.sweep || k >= cap) continue;
      CUDA_CHECK(cudaStreamSynchronize(c->stream));
      CUDA_CHECK(cudaMemsetAsync(b.s_ptil.p, 0, b.s_ptil.n * sizeof(double), c->stream));
      CUDA_CHECK(cudaMemsetAsync(b.s_cnt.p, 0, b.s_cnt.n * sizeof(uint32_t), c->stream));
      CUDA_CHECK(cudaEventRecord(c->ev_t0, c->stream));
      const cudaError_t err = launch_sweep(b.M, b.factor->sweep_ct, b.sargs, b.factor->sweep_n_cta, b.sweep_smem, c->stream);
      if (err != cudaSuccess) fail(MMMFE_E_CUDA, "sweep kernel launch failed: %s", cudaGetErrorString(err));
      CUDA_CHECK(cudaEventRecord(c->ev_t1, c->stream));
      CUDA_CHECK(cudaStreamSynchronize(c->stream));
      check_watchdog(c);
      float t = 0;
      CUDA_CHECK(cudaEventElapsedTime(&t, c->ev_t0, c->ev_t1));
      ms[k] = t;
      // factor stream + 8 (2 n + 2 sum n_side) bytes of vector traffic per member
      bytes_per_step[k] = double(b.factor->solve_bytes) +
                          double(b.members.size()) * 16.0 * (double(b.factor->nf) + b.factor->np) + 16.0 * b.n_tr;
      n_steps[k] = b.nt;
      ++k;
    }
    *n_out = k;
  });
}

int mmmfe_profile_basis_apply(mmmfe_ctx *c, int32_t repeat, double *ms_per_launch) {
  if (!c || repeat < 1 || !ms_per_launch) return MMMFE_E_INVALID;
  MMMFE_TRY(c, {
    if (!c->use_basis) fail(MMMFE_E_INVALID, "the multiscale-basis operator is not in use (%s)", c->basis_note.c_str());
    cudaStream_t st = c->stream;
    auto once = [&]() {
      if (c->toep_parts <= 1) {
        launch_toeplitz_apply(c->toep_bm, c->toep_probs.p, c->toep_tiles.p, c->n_toep_tiles, c->qin.p, c->send.p, 0, st);
      } else {
        launch_toeplitz_apply(c->toep_bm, c->toep_probs.p, c->toep_tiles.p, c->n_toep_tiles, c->qin.p, c->toep_part_buf.p, c->n_iface, st);
        launch_sum_parts(c->n_iface, c->toep_part_buf.p, c->n_iface, c->toep_parts, c->send.p, st);
      }
    };
    once();  // warm-up
    CUDA_CHECK(cudaEventRecord(c->ev_t0, st));
    for (int i = 0; i < repeat; ++i) once();
    CUDA_CHECK(cudaEventRecord(c->ev_t1, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    float t = 0;
    CUDA_CHECK(cudaEventElapsedTime(&t, c->ev_t0, c->ev_t1));
    *ms_per_launch = double(t) / repeat;
  });
}

int mmmfe_debug_get_basis(mmmfe_ctx *c, int32_t li, int64_t info[12], double *t_out, int64_t capacity) {
  if (!c || li < 0 || li >= int32_t(c->subs.size()) || !info) return MMMFE_E_INVALID;
  MMMFE_TRY(c, {
    if (!c->is_setup) fail(MMMFE_E_INVALID, "mmmfe_setup first");
    const Factor &F = *c->batches[c->subs[li]->batch]->factor;
    if (!c->use_basis || !F.basis) fail(MMMFE_E_INVALID, "the multiscale-basis operator is not in use (%s)", c->basis_note.c_str());
    const Basis &B = *F.basis;
    info[0] = B.m_t; info[1] = B.P_tot; info[2] = B.ldt; info[3] = B.n_slots;
    for (int sd = 0; sd < 4; ++sd) {
      info[4 + sd] = B.slot_of_side[sd];
      info[8 + sd] = B.slot_of_side[sd] >= 0 ? B.base[B.slot_of_side[sd]] : -1;
    }
    if (t_out) {
      if (capacity < int64_t(B.T.n)) fail(MMMFE_E_INVALID, "buffer too small (%lld needed)", (long long)B.T.n);
      CUDA_CHECK(cudaMemcpy(t_out, B.T.p, B.T.n * sizeof(double), cudaMemcpyDeviceToHost));
    }
  });
}

int mmmfe_debug_get_factor(mmmfe_ctx *c, int32_t li, double *r_out, int64_t capacity) {
  if (!c || li < 0 || li >= int32_t(c->subs.size()) || !r_out) return MMMFE_E_INVALID;
  MMMFE_TRY(c, {
    if (!c->is_setup) fail(MMMFE_E_INVALID, "mmmfe_setup first");
    const Factor &F = *c->batches[c->subs[li]->batch]->factor;
    if (capacity < int64_t(F.R.n)) fail(MMMFE_E_INVALID, "buffer too small (%lld needed)", (long long)F.R.n);
    CUDA_CHECK(cudaMemcpy(r_out, F.R.p, F.R.n * sizeof(double), cudaMemcpyDeviceToHost));
  });
}

}  // extern "C"
