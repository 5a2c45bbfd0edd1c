"""ctypes binding of the B200 interface-GMRES engine (include/mmmfe_b200.h) and of the C++ host front-end
(include/mmmfe_host.h).  Python is plumbing only: tests, bench.py and launch scripts call the same C ABI the
``DarcyVT`` executable uses.  There is no CPU fallback: a missing library or device raises.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ENGINE_PATH = os.path.join(HERE, "libmmmfe_b200.so")
HOST_PATH = os.path.join(HERE, "libmmmfe_host.so")

c_i32p = ctypes.POINTER(ctypes.c_int32)
c_i64p = ctypes.POINTER(ctypes.c_int64)
c_f64p = ctypes.POINTER(ctypes.c_double)


class MmmfeError(RuntimeError):
    pass


class Csr(ctypes.Structure):
    _fields_ = [("n_rows", ctypes.c_int64), ("n_cols", ctypes.c_int64), ("row_ptr", c_i64p), ("col", c_i32p),
                ("val", c_f64p)]


class SubdomainDesc(ctypes.Structure):
    _fields_ = [("global_id", ctypes.c_int32), ("nx", ctypes.c_int32), ("ny", ctypes.c_int32),
                ("nt", ctypes.c_int32), ("dt", ctypes.c_double), ("matrix", Csr), ("n_flux", ctypes.c_int32),
                ("n_pressure", ctypes.c_int32), ("canon_of_dof", c_i32p), ("neighbor", ctypes.c_int32 * 4),
                ("n_side", ctypes.c_int32 * 4), ("side_dofs", c_i32p * 4), ("mortar_len", ctypes.c_int32 * 4),
                ("f2c", Csr * 4), ("c2f", Csr * 4)]


class GmresInfo(ctypes.Structure):
    _fields_ = [("iterations", ctypes.c_int32), ("converged", ctypes.c_int32), ("r_norm", ctypes.c_double),
                ("final_residual", ctypes.c_double), ("seconds", ctypes.c_double), ("device_ms", ctypes.c_double)]


_engine = None


def load_engine():
    """dlopen libmmmfe_b200.so (built in-tree by build.py); raises when it is absent."""
    global _engine
    if _engine is not None:
        return _engine
    if not os.path.exists(ENGINE_PATH):
        raise MmmfeError(f"{ENGINE_PATH} is missing: run `python mmmfe-st-dd_b200/build.py` (no CPU fallback exists)")
    lib = ctypes.CDLL(ENGINE_PATH)
    vp = ctypes.c_void_p
    sig = {
        "mmmfe_create": (ctypes.c_int, [ctypes.POINTER(vp), ctypes.c_int]),
        "mmmfe_destroy": (ctypes.c_int, [vp]),
        "mmmfe_last_error": (ctypes.c_char_p, [vp]),
        "mmmfe_version": (ctypes.c_char_p, []),
        "mmmfe_set_option": (ctypes.c_int, [vp, ctypes.c_char_p, ctypes.c_double]),
        "mmmfe_comm_unique_id_size": (ctypes.c_int, []),
        "mmmfe_comm_unique_id": (ctypes.c_int, [vp]),
        "mmmfe_comm_init": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, vp]),
        "mmmfe_set_owners": (ctypes.c_int, [vp, ctypes.c_int32, c_i32p]),
        "mmmfe_comm_allreduce_sum": (ctypes.c_int, [vp, c_f64p, ctypes.c_int32]),
        "mmmfe_kernel_launches": (ctypes.c_int64, [vp]),
        "mmmfe_measure_fp64_tensor_tflops": (ctypes.c_int, [ctypes.c_int, c_f64p]),
        "mmmfe_add_subdomain": (ctypes.c_int, [vp, ctypes.POINTER(SubdomainDesc), c_i32p]),
        "mmmfe_setup": (ctypes.c_int, [vp]),
        "mmmfe_interface_length": (ctypes.c_int64, [vp]),
        "mmmfe_interface_offset": (ctypes.c_int64, [vp, ctypes.c_int32, ctypes.c_int32]),
        "mmmfe_num_factor_groups": (ctypes.c_int32, [vp]),
        "mmmfe_factor_values": (ctypes.c_int64, [vp]),
        "mmmfe_step_bytes": (ctypes.c_int64, [vp]),
        "mmmfe_factor_flops": (ctypes.c_double, [vp]),
        "mmmfe_factor_seconds": (ctypes.c_double, [vp]),
        "mmmfe_bar_set_data": (ctypes.c_int, [vp, ctypes.c_int32, c_f64p, c_f64p, ctypes.c_int32, c_i32p, c_f64p]),
        "mmmfe_bar_sweep": (ctypes.c_int, [vp]),
        "mmmfe_apply_interface_operator": (ctypes.c_int, [vp, c_f64p, c_f64p]),
        "mmmfe_get_initial_residual": (ctypes.c_int, [vp, c_f64p]),
        "mmmfe_alloc_pinned": (ctypes.c_void_p, [ctypes.c_int64]),
        "mmmfe_free_pinned": (None, [ctypes.c_void_p]),
        "mmmfe_apply_interface_operator_device": (ctypes.c_int, [vp, ctypes.c_int32, c_f64p]),
        "mmmfe_gmres_solve": (ctypes.c_int, [vp, ctypes.c_double, ctypes.c_int32, ctypes.POINTER(GmresInfo), c_f64p,
                                             ctypes.c_int32, c_f64p]),
        "mmmfe_final_sweep": (ctypes.c_int, [vp]),
        "mmmfe_get_solution": (ctypes.c_int, [vp, ctypes.c_int32, c_f64p]),
        "mmmfe_comm_warmup": (ctypes.c_int, [vp, ctypes.c_int32, ctypes.POINTER(ctypes.c_int32)]),
        "mmmfe_solution_device": (ctypes.c_int, [vp, ctypes.c_int32, ctypes.POINTER(ctypes.c_void_p)]),
        "mmmfe_bar_data_buffers": (ctypes.c_int, [vp, ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(ctypes.c_int32),
                                                  ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p),
                                                  ctypes.POINTER(ctypes.c_void_p)]),
        "mmmfe_get_star_trace": (ctypes.c_int, [vp, ctypes.c_int32, ctypes.c_int32, c_f64p]),
        "mmmfe_solve_saddle": (ctypes.c_int, [vp, ctypes.c_int32, c_f64p, c_f64p]),
        "mmmfe_profile_step": (ctypes.c_int, [vp, c_f64p, c_f64p, c_i64p]),
        "mmmfe_debug_get_factor": (ctypes.c_int, [vp, ctypes.c_int32, c_f64p, ctypes.c_int64]),
        "mmmfe_debug_get_basis": (ctypes.c_int, [vp, ctypes.c_int32, c_i64p, c_f64p, ctypes.c_int64]),
        "mmmfe_profile_basis_apply": (ctypes.c_int, [vp, ctypes.c_int32, c_f64p]),
        "mmmfe_profile_sweeps": (ctypes.c_int, [vp, ctypes.c_int32, c_f64p, c_f64p, c_i32p, c_i32p]),
        "mmmfe_stat": (ctypes.c_int64, [vp, ctypes.c_char_p]),
        "mmmfe_profile_sweep_cycles": (ctypes.c_int, [vp, ctypes.c_int32, c_f64p]),
        "mmmfe_debug_sweep_trace": (ctypes.c_int, [vp, ctypes.c_int32, ctypes.c_int64, c_i64p, c_i32p, c_i32p, c_i64p]),
        "mmmfe_sweep_plan_check": (ctypes.c_int, [ctypes.c_int32] * 6 + [ctypes.c_int64, ctypes.c_int32, ctypes.c_int32,
                                                  ctypes.c_int32, c_i64p, ctypes.c_char_p, ctypes.c_int32]),
        "mmmfe_exchange_plan": (ctypes.c_int, [ctypes.c_int32, c_i32p, c_i32p, c_i32p, c_i64p, c_i64p, ctypes.c_int64,
                                               ctypes.c_int32, c_i32p, ctypes.c_int32, ctypes.c_int32, c_i64p, c_i64p,
                                               ctypes.c_int64, c_i64p, c_i32p, c_i64p, c_i64p, c_i64p, ctypes.c_int32,
                                               c_i32p]),
        "mmmfe_assign_owners": (ctypes.c_int, [ctypes.c_int32, c_f64p, ctypes.c_int32, c_i32p]),
        "mmmfe_nd_tree_sizes": (ctypes.c_int, [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, c_i64p]),
        "mmmfe_nd_tree_fill": (ctypes.c_int, [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, c_i32p, c_i64p, c_i32p,
                                              c_i32p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    lib._signatures = sig
    _engine = lib
    return lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a, typ):
    return a.ctypes.data_as(typ) if a is not None else None


def _csr(mat, keep):
    """scipy CSR -> Csr struct (arrays kept alive in `keep`)."""
    mat = mat.tocsr()
    rp = np.ascontiguousarray(mat.indptr, dtype=np.int64)
    col = np.ascontiguousarray(mat.indices, dtype=np.int32)
    val = _f64(mat.data)
    keep.extend([rp, col, val])
    return Csr(mat.shape[0], mat.shape[1], _ptr(rp, c_i64p), _ptr(col, c_i32p), _ptr(val, c_f64p))


class Engine:
    """Thin object wrapper over the C ABI; one per process and refinement cycle."""

    def __init__(self, device=-1):
        self.lib = load_engine()
        self.h = ctypes.c_void_p()
        rc = self.lib.mmmfe_create(ctypes.byref(self.h), device)
        if rc != 0:
            raise MmmfeError(f"mmmfe_create failed ({rc}): no usable CUDA device; this package has no CPU path")
        self.subs = []

    def close(self):
        if self.h:
            self.lib.mmmfe_destroy(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise MmmfeError(f"mmmfe error {rc}: {self.lib.mmmfe_last_error(self.h).decode()}")

    def set_option(self, key, value):
        self._check(self.lib.mmmfe_set_option(self.h, key.encode(), float(value)))

    def comm_init(self, rank, world, uid_bytes):
        buf = ctypes.create_string_buffer(uid_bytes, len(uid_bytes))
        self._check(self.lib.mmmfe_comm_init(self.h, rank, world, ctypes.cast(buf, ctypes.c_void_p)))

    def unique_id(self):
        n = self.lib.mmmfe_comm_unique_id_size()
        buf = ctypes.create_string_buffer(n)
        rc = self.lib.mmmfe_comm_unique_id(ctypes.cast(buf, ctypes.c_void_p))
        if rc != 0:
            raise MmmfeError("ncclGetUniqueId failed")
        return buf.raw

    def set_owners(self, owners):
        o = np.ascontiguousarray(owners, dtype=np.int32)
        self._check(self.lib.mmmfe_set_owners(self.h, len(o), _ptr(o, c_i32p)))

    def add_subdomain(self, *, global_id, nx, ny, nt, dt, matrix, n_flux, n_pressure, neighbor, side_dofs, f2c, c2f,
                      canon_of_dof=None):
        keep = []
        d = SubdomainDesc()
        d.global_id, d.nx, d.ny, d.nt, d.dt = global_id, nx, ny, nt, dt
        d.matrix = _csr(matrix, keep)
        d.n_flux, d.n_pressure = n_flux, n_pressure
        if canon_of_dof is not None:
            cn = np.ascontiguousarray(canon_of_dof, dtype=np.int32)
            keep.append(cn)
            d.canon_of_dof = _ptr(cn, c_i32p)
        for s in range(4):
            d.neighbor[s] = neighbor[s]
            if neighbor[s] >= 0:
                sdofs = np.ascontiguousarray(side_dofs[s], dtype=np.int32)
                keep.append(sdofs)
                d.n_side[s] = len(sdofs)
                d.side_dofs[s] = _ptr(sdofs, c_i32p)
                d.mortar_len[s] = f2c[s].shape[0]
                d.f2c[s] = _csr(f2c[s], keep)
                d.c2f[s] = _csr(c2f[s], keep)
        li = ctypes.c_int32(-1)
        self._check(self.lib.mmmfe_add_subdomain(self.h, ctypes.byref(d), ctypes.byref(li)))
        self.subs.append(dict(nt=nt, n=n_flux + n_pressure, n_side=[d.n_side[s] for s in range(4)]))
        return li.value

    def setup(self):
        self._check(self.lib.mmmfe_setup(self.h))

    @property
    def interface_length(self):
        return int(self.lib.mmmfe_interface_length(self.h))

    def interface_offset(self, li, side):
        return int(self.lib.mmmfe_interface_offset(self.h, li, side))

    def stats(self):
        return dict(factor_groups=int(self.lib.mmmfe_num_factor_groups(self.h)),
                    factor_values=int(self.lib.mmmfe_factor_values(self.h)),
                    step_bytes=int(self.lib.mmmfe_step_bytes(self.h)),
                    factor_flops=float(self.lib.mmmfe_factor_flops(self.h)),
                    factor_seconds=float(self.lib.mmmfe_factor_seconds(self.h)))

    @property
    def kernel_launches(self):
        return int(self.lib.mmmfe_kernel_launches(self.h))

    def bar_set_data(self, li, p0, src, fr_dofs, fr_vals):
        p0 = _f64(p0) if p0 is not None else None
        src = _f64(src)
        fr_dofs = np.ascontiguousarray(fr_dofs, dtype=np.int32)
        fr_vals = _f64(fr_vals)
        self._check(self.lib.mmmfe_bar_set_data(self.h, li, _ptr(p0, c_f64p), _ptr(src, c_f64p), len(fr_dofs),
                                                _ptr(fr_dofs, c_i32p), _ptr(fr_vals, c_f64p)))

    def bar_sweep(self):
        self._check(self.lib.mmmfe_bar_sweep(self.h))

    def apply(self, q):
        q = _f64(q)
        out = np.empty_like(q)
        self._check(self.lib.mmmfe_apply_interface_operator(self.h, _ptr(q, c_f64p), _ptr(out, c_f64p)))
        return out

    def apply_device(self, repeat=1):
        ms = ctypes.c_double(0.0)
        self._check(self.lib.mmmfe_apply_interface_operator_device(self.h, repeat, ctypes.byref(ms)))
        return ms.value

    def gmres(self, tol, max_iter):
        info = GmresInfo()
        hist = np.zeros(max_iter + 2)
        lam = np.zeros(self.interface_length)
        self._check(self.lib.mmmfe_gmres_solve(self.h, tol, max_iter, ctypes.byref(info), _ptr(hist, c_f64p),
                                               len(hist), _ptr(lam, c_f64p)))
        return dict(iterations=info.iterations, converged=bool(info.converged), r_norm=info.r_norm,
                    final_residual=info.final_residual, seconds=info.seconds, device_ms=info.device_ms,
                    residuals=hist[:info.iterations + 1].copy(), lam=lam)

    def final_sweep(self):
        self._check(self.lib.mmmfe_final_sweep(self.h))

    def get_solution(self, li):
        s = self.subs[li]
        out = np.empty((s["nt"], s["n"]))
        self._check(self.lib.mmmfe_get_solution(self.h, li, _ptr(out, c_f64p)))
        return out

    def get_star_trace(self, li, side):
        s = self.subs[li]
        out = np.empty((s["nt"], s["n_side"][side]))
        self._check(self.lib.mmmfe_get_star_trace(self.h, li, side, _ptr(out, c_f64p)))
        return out

    def profile_step(self):
        ms, by, ln = np.zeros(6), np.zeros(6), np.zeros(6, dtype=np.int64)
        self._check(self.lib.mmmfe_profile_step(self.h, _ptr(ms, c_f64p), _ptr(by, c_f64p), _ptr(ln, c_i64p)))
        names = ["rhs_build", "subtree_forward", "forward_top", "backward_top", "subtree_backward", "update_trace"]
        return {n: dict(ms=float(ms[i]), bytes=float(by[i]), launches=int(ln[i])) for i, n in enumerate(names)}

    def stat(self, key):
        return int(self.lib.mmmfe_stat(self.h, key.encode()))

    def profile_sweeps(self):
        """One timed star sweep per batch on the persistent-kernel path: [(ms, bytes_per_step, n_steps)]."""
        cap = 64
        ms, by = np.zeros(cap), np.zeros(cap)
        ns, n = np.zeros(cap, dtype=np.int32), np.zeros(1, dtype=np.int32)
        self._check(self.lib.mmmfe_profile_sweeps(self.h, cap, _ptr(ms, c_f64p), _ptr(by, c_f64p), _ptr(ns, c_i32p),
                                                  _ptr(n, c_i32p)))
        return [(float(ms[i]), float(by[i]), int(ns[i])) for i in range(int(n[0]))]

    def sweep_trace(self, batch):
        n = np.zeros(1, dtype=np.int64)
        self._check(self.lib.mmmfe_debug_sweep_trace(self.h, batch, 0, _ptr(n, c_i64p), None, None, None))
        ne, ncta = int(n[0]), self.stat("sweep_ctas")
        cb = np.zeros(ncta + 1, dtype=np.int32)
        kh = np.zeros((ne, 2), dtype=np.int32)
        st = np.zeros((ne, 3), dtype=np.int64)
        self._check(self.lib.mmmfe_debug_sweep_trace(self.h, batch, ne, _ptr(n, c_i64p), _ptr(cb, c_i32p), _ptr(kh, c_i32p),
                                                     _ptr(st, c_i64p)))
        return cb, kh, st

    def profile_sweep_cycles(self, batch):
        out = np.zeros(24)
        self._check(self.lib.mmmfe_profile_sweep_cycles(self.h, batch, _ptr(out, c_f64p)))
        names = ["sub_fwd", "top_fwd", "top_bwd", "sub_bwd", "counter_wait", "tma_wait", "entries", "open_wait",
                 "sb_gather", "sb_levels", "sb_epilogue", "sf_rhs", "sf_levels", "sf_output", "top_gather", "top_tile",
                 "wait_sub_fwd", "wait_top_fwd", "wait_top_bwd", "wait_sub_bwd", "wave_sync0", "wave_tile",
                 "wave_sync1", "wave_out"]
        return dict(zip(names, (float(v) for v in out)))

    def debug_get_factor(self, li, r_total):
        out = np.empty(r_total)
        self._check(self.lib.mmmfe_debug_get_factor(self.h, li, _ptr(out, c_f64p), r_total))
        return out

    def get_basis(self, li):
        """Multiscale-basis operator of the factor group of local subdomain li: (info dict, T[m_t][P_tot][P_tot])."""
        info = np.zeros(12, dtype=np.int64)
        self._check(self.lib.mmmfe_debug_get_basis(self.h, li, _ptr(info, c_i64p), None, 0))
        m_t, p_tot, ldt = int(info[0]), int(info[1]), int(info[2])
        t = np.empty(m_t * p_tot * ldt)
        self._check(self.lib.mmmfe_debug_get_basis(self.h, li, _ptr(info, c_i64p), _ptr(t, c_f64p), t.size))
        meta = dict(m_t=m_t, P_tot=p_tot, slot_of_side=[int(v) for v in info[4:8]], base_of_side=[int(v) for v in info[8:12]])
        return meta, t.reshape(m_t, p_tot, ldt)[:, :, :p_tot].copy()

    def solve_saddle(self, li, rhs):
        rhs = _f64(rhs)
        out = np.empty_like(rhs)
        self._check(self.lib.mmmfe_solve_saddle(self.h, li, _ptr(rhs, c_f64p), _ptr(out, c_f64p)))
        return out


def fp64_tensor_peak_tflops(device=-1):
    """Measured DMMA peak of the device (TFLOP/s)."""
    lib = load_engine()
    v = ctypes.c_double(0.0)
    if lib.mmmfe_measure_fp64_tensor_tflops(device, ctypes.byref(v)) != 0:
        raise MmmfeError("DMMA peak measurement failed")
    return v.value


def sweep_plan_check(nx, ny, leaf_cells=4, subtree_cells=128, M=1, n_cta=296, smem_bytes=115000, n_steps=3,
                     compute_threads=256, chunk_doubles=2048):
    """Builds the persistent-kernel schedule on the host and replays it on the CPU; returns (rc, stats, message)."""
    lib = load_engine()
    stats = np.zeros(8, dtype=np.int64)
    msg = ctypes.create_string_buffer(256)
    rc = lib.mmmfe_sweep_plan_check(nx, ny, leaf_cells, subtree_cells, M, n_cta, smem_bytes, n_steps,
                                    compute_threads, chunk_doubles, _ptr(stats, c_i64p), msg, 256)
    keys = ["nodes", "top_fronts", "entries", "factor_doubles", "ring_doubles", "ws_doubles", "max_entry_doubles",
            "smem_bytes"]
    return rc, dict(zip(keys, (int(v) for v in stats))), msg.value.decode()


def exchange_plan(sides, n_iface, owner, rank, world):
    """Host only: the face-exchange plan mmmfe_setup builds for `rank` (csrc/exchange_plan.cpp).
    sides: list of (global id, side, neighbour global id, offset, length).  Returns (partner, send_index, peers) with
    peers = [(rank, send_off, recv_off, count)]."""
    lib = load_engine()
    n = len(sides)
    gid = np.array([s[0] for s in sides], dtype=np.int32)
    side = np.array([s[1] for s in sides], dtype=np.int32)
    nb = np.array([s[2] for s in sides], dtype=np.int32)
    off = np.array([s[3] for s in sides], dtype=np.int64)
    ln = np.array([s[4] for s in sides], dtype=np.int64)
    own = np.asarray(owner, dtype=np.int32)
    partner = np.zeros(max(1, n_iface), dtype=np.int64)
    send_index = np.zeros(max(1, n_iface), dtype=np.int64)
    n_send = np.zeros(1, dtype=np.int64)
    cap = max(1, world)
    pr = np.zeros(cap, dtype=np.int32)
    ps, prc, pc = (np.zeros(cap, dtype=np.int64) for _ in range(3))
    n_peers = np.zeros(1, dtype=np.int32)
    rc = lib.mmmfe_exchange_plan(n, _ptr(gid, c_i32p), _ptr(side, c_i32p), _ptr(nb, c_i32p), _ptr(off, c_i64p),
                                 _ptr(ln, c_i64p), ctypes.c_int64(n_iface), len(own), _ptr(own, c_i32p), rank, world,
                                 _ptr(partner, c_i64p), _ptr(send_index, c_i64p), ctypes.c_int64(len(send_index)),
                                 _ptr(n_send, c_i64p), _ptr(pr, c_i32p), _ptr(ps, c_i64p), _ptr(prc, c_i64p),
                                 _ptr(pc, c_i64p), cap, _ptr(n_peers, c_i32p))
    if rc != 0:
        raise MmmfeError(f"mmmfe_exchange_plan failed ({rc})")
    k = int(n_peers[0])
    return partner[:n_iface], send_index[:int(n_send[0])], [(int(pr[i]), int(ps[i]), int(prc[i]), int(pc[i])) for i in range(k)]


def assign_owners(cost, world):
    """Host only: the front-end's ownership rule (contiguous blocks balanced by nx*ny*nt)."""
    lib = load_engine()
    cost = np.asarray(cost, dtype=np.float64)
    owner = np.zeros(len(cost), dtype=np.int32)
    if lib.mmmfe_assign_owners(len(cost), _ptr(cost, c_f64p), world, _ptr(owner, c_i32p)) != 0:
        raise MmmfeError("mmmfe_assign_owners failed")
    return owner


def nd_tree(nx, ny, leaf_cells=4):
    """The engine's nested-dissection tree as numpy arrays (host only)."""
    lib = load_engine()
    sizes = np.zeros(6, dtype=np.int64)
    if lib.mmmfe_nd_tree_sizes(nx, ny, leaf_cells, _ptr(sizes, c_i64p)) != 0:
        raise MmmfeError("bad grid")
    nodes = np.zeros((sizes[0], 8), dtype=np.int32)
    offs = np.zeros((sizes[0], 4), dtype=np.int64)
    idx = np.zeros(sizes[1], dtype=np.int32)
    src = np.zeros(sizes[2], dtype=np.int32)
    lib.mmmfe_nd_tree_fill(nx, ny, leaf_cells, _ptr(nodes, c_i32p), _ptr(offs, c_i64p), _ptr(idx, c_i32p),
                           _ptr(src, c_i32p))
    return dict(nodes=nodes, offsets=offs, idx=idx, src=src, r_total=int(sizes[3]), z_total=int(sizes[4]),
                n_heights=int(sizes[5]))


# ---- C++ host front-end (include/mmmfe_host.h) --------------------------------------------------------------------
class HostOptions(ctypes.Structure):
    _fields_ = [("rank", ctypes.c_int32), ("world", ctypes.c_int32), ("nccl_id", ctypes.c_void_p),
                ("device", ctypes.c_int32), ("compute_errors", ctypes.c_int32), ("write_output", ctypes.c_int32),
                ("final_sweep", ctypes.c_int32), ("verbose", ctypes.c_int32), ("only_cycle", ctypes.c_int32),
                ("fixed_iterations", ctypes.c_int32), ("apply_repeat", ctypes.c_int32),
                ("num_refinement", ctypes.c_int32), ("mortar_degree", ctypes.c_int32), ("output_dir", ctypes.c_char_p),
                ("warmup_iterations", ctypes.c_int32), ("e2e_iterations", ctypes.c_int32), ("profile", ctypes.c_int32),
                ("engine_options", ctypes.c_char_p), ("also_solve", ctypes.c_int32), ("skip_fetch", ctypes.c_int32)]


class HostReport(ctypes.Structure):
    _fields_ = [("cycle", ctypes.c_int32), ("gmres_iterations", ctypes.c_int32), ("converged", ctypes.c_int32),
                ("factor_groups", ctypes.c_int32), ("r_norm", ctypes.c_double), ("velocity_l2l2", ctypes.c_double),
                ("pressure_dg", ctypes.c_double), ("pressure_l2l2", ctypes.c_double), ("lambda_l2", ctypes.c_double),
                ("t_assemble", ctypes.c_double), ("t_factor", ctypes.c_double), ("t_bar_data", ctypes.c_double),
                ("t_bar", ctypes.c_double), ("t_gmres", ctypes.c_double), ("t_final", ctypes.c_double),
                ("t_errors", ctypes.c_double), ("t_output", ctypes.c_double), ("gmres_device_ms", ctypes.c_double),
                ("apply_ms", ctypes.c_double), ("factor_flops", ctypes.c_double),
                ("interface_length", ctypes.c_int64), ("dof_steps", ctypes.c_int64), ("step_bytes", ctypes.c_int64),
                ("factor_values", ctypes.c_int64), ("kernel_launches", ctypes.c_int64),
                ("e2e_ms_per_iteration", ctypes.c_double), ("e2e_h2d_bytes", ctypes.c_int64),
                ("e2e_d2h_bytes", ctypes.c_int64), ("gmres_kernel_launches", ctypes.c_int64),
                ("prof_ms", ctypes.c_double * 6), ("prof_bytes", ctypes.c_double * 6),
                ("prof_launches", ctypes.c_int64 * 6), ("sweep_batches", ctypes.c_int32),
                ("sweep_steps", ctypes.c_int32 * 4), ("sweep_ms", ctypes.c_double * 4),
                ("sweep_bytes_per_step", ctypes.c_double * 4), ("tune_us_sweep", ctypes.c_int64),
                ("tune_us_levels", ctypes.c_int64), ("sweep_planned", ctypes.c_int64),
                ("tune_rel_diff_e18", ctypes.c_int64), ("basis_operator", ctypes.c_int32), ("basis_shared_by", ctypes.c_int32),
                ("t_basis", ctypes.c_double), ("basis_gflop", ctypes.c_double), ("basis_bytes", ctypes.c_double),
                ("basis_columns", ctypes.c_int64), ("solve_iterations", ctypes.c_int32),
                ("solve_converged", ctypes.c_int32), ("t_solve_gmres", ctypes.c_double),
                ("solve_final_residual", ctypes.c_double), ("toeplitz_ms", ctypes.c_double),
                ("toeplitz_gflop", ctypes.c_double), ("toeplitz_mbytes", ctypes.c_double),
                ("basis_steps", ctypes.c_int64), ("basis_delays", ctypes.c_int64), ("basis_mw", ctypes.c_int64),
                ("toeplitz_tiles", ctypes.c_int64)]


_host = None


def load_host():
    """dlopen libmmmfe_host.so (the C++ front-end); raises when it is absent."""
    global _host
    if _host is not None:
        return _host
    load_engine()
    if not os.path.exists(HOST_PATH):
        raise MmmfeError(f"{HOST_PATH} is missing: run `python mmmfe-st-dd_b200/build.py`")
    lib = ctypes.CDLL(HOST_PATH)
    sig = {
        "mmmfe_host_default_options": (None, [ctypes.POINTER(HostOptions)]),
        "mmmfe_host_run": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(HostOptions), ctypes.POINTER(HostReport),
                                          ctypes.c_int32]),
        "mmmfe_host_last_residuals": (ctypes.c_int, [c_f64p, ctypes.c_int32]),
        "mmmfe_host_last_lambda": (ctypes.c_int64, [c_f64p, ctypes.c_int64]),
        "mmmfe_host_last_error": (ctypes.c_char_p, []),
        "mmmfe_host_dump_subdomain": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                                     ctypes.c_char_p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    lib._signatures = sig
    _host = lib
    return lib


def host_run(parameter_file, nccl_id=None, **options):
    """Runs the DarcyVT front-end on a parameter file; returns (reports, residuals, lambda)."""
    lib = load_host()
    opt = HostOptions()
    lib.mmmfe_host_default_options(ctypes.byref(opt))
    keep = None
    if nccl_id is not None:
        keep = ctypes.create_string_buffer(nccl_id, len(nccl_id))
        opt.nccl_id = ctypes.cast(keep, ctypes.c_void_p)
    for k, v in options.items():
        if k in ("output_dir", "engine_options"):
            v = v.encode()
        setattr(opt, k, v)
    reps = (HostReport * 16)()
    n = lib.mmmfe_host_run(os.fsencode(parameter_file), ctypes.byref(opt), reps, 16)
    if n < 0:
        raise MmmfeError(lib.mmmfe_host_last_error().decode())
    out = []
    for i in range(min(n, 16)):
        d = {}
        for f, _ in HostReport._fields_:
            v = getattr(reps[i], f)
            d[f] = list(v) if hasattr(v, "__len__") else v
        out.append(d)
    nres = lib.mmmfe_host_last_residuals(None, 0)
    res = np.zeros(nres)
    lib.mmmfe_host_last_residuals(_ptr(res, c_f64p), nres)
    nl = lib.mmmfe_host_last_lambda(None, 0)
    lam = np.zeros(nl)
    lib.mmmfe_host_last_lambda(_ptr(lam, c_f64p), nl)
    return out, res, lam


def write_parameter_file(path, mesh, mortar, mortar_degree=1, num_refinement=1, c0=1.0, final_time=0.5, tol=1e-6,
                         max_iteration=500, is_manufact=1, bc=("D 0.0", "D 0.0", "D 0.0", "D 0.0"), plots=0):
    """Writes a parameter.txt in the reference's positional format (inc/filecheck_utility.h:44-84)."""
    lines = ["synthetic DarcyVT parameter file", "-" * 40, "positional format: keep one key/value per line", "-" * 40,
             f"c0: {c0}", "", "alpha: 1.0", "", "space_degree: 0", "", f"mortar_degree: {mortar_degree}", "",
             f"num_refinement: {num_refinement}", "", f"final_time: {final_time}", "", f"tolerence: {tol}", "",
             f"max_iteration: {max_iteration}", "", f"need_plot_at_each_time_step(bool): {plots}", "",
             f"left_bc: {bc[0]}", f"bottom_bc: {bc[1]}", f"right_bc: {bc[2]}", f"top_bc: {bc[3]}", "",
             f"is_manufact_soln(bool): {is_manufact}", ""]
    for i, m in enumerate(mesh):
        lines.append(f"mesh_pattern_sub_d{i}: {m[0]} {m[1]} {m[2]}")
    lines.append(f"mesh_pattern_mortar: {mortar[0]} {mortar[1]} {mortar[2]}")
    with open(path, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    return path
