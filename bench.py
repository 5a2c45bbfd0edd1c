#!/usr/bin/env python
"""Benchmark of the interface-GMRES hot path (BASELINE.json metric) -- prints ONE JSON line on rank 0.

  python bench.py --gpus N --steps K --warmup W              # the B200 engine through the DarcyVT front-end
  python bench.py --impl reference --gpus N --steps K ...    # the CPU restatement of the reference on host cores

A "step" is one interface-GMRES iteration: mortar->fine projection, nt backward-Euler star solves of every
subdomain, fine->mortar projection, face exchange, classical Gram-Schmidt update.  The workload at N = 1 is
BASELINE.json configs[1] (2x2 subdomains, non-matching space/time grids with ratio 2: 1024^2 x 256 steps on the
fine colour, 512^2 x 128 on the coarse one, mortar 256 x 256 x 64, degree 1).  For N > 1 the subdomain grid grows
to 4 N subdomains of the same sizes (checkerboard), 4 per GPU: weak scaling.  `--config cfg3` / `cfg4` run
BASELINE.json configs[2] / configs[3] (8 subdomains per GPU: 4x4 at --gpus 2, 8x8 at --gpus 8) for DESIGN.md section 7;
the driver's bench line is the default cfg2.
`value` = subdomain DOF*timesteps per second over all GPUs (DOF*steps of one operator application x iterations/s).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (fine pattern, coarse pattern, mortar pattern, mortar degree[, subdomains per GPU (default 4)])
    "cfg2": ([1024, 1024, 256], [512, 512, 128], [256, 256, 64], 1),
    "cfg2_half": ([512, 512, 128], [256, 256, 64], [128, 128, 32], 1),
    "cfg2_small": ([128, 128, 32], [64, 64, 16], [32, 32, 8], 1),
    # BASELINE.json configs[2] (4x4 checkerboard at --gpus 2) and configs[3] (8x8 = 64 subdomains, 2048^2 global x 512
    # fine steps, quadratic mortar, at --gpus 8): not the bench line, measured for DESIGN.md section 7
    "cfg3": ([512, 512, 128], [256, 256, 64], [128, 128, 32], 1, 8),
    "cfg4": ([256, 256, 512], [128, 128, 256], [16, 16, 32], 2, 8),
}


def find_divisors(n):
    import math
    t = math.floor(math.sqrt(n))
    while n % t:
        t -= 1
    return n // t, t


def checkerboard(n_sub, fine, coarse):
    n0, _ = find_divisors(n_sub)
    return [list(fine) if ((r % n0) + (r // n0)) % 2 == 0 else list(coarse) for r in range(n_sub)]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.device)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def measured_traffic(kernel_key, workload):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (bench.py cannot run
    under a profiler): profiles/traffic.json maps kernel -> {workload, dram_bytes_per_launch, source}."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None, None
    try:
        t = json.load(open(p)).get(kernel_key)
    except (ValueError, OSError):
        return None, None
    if not t or t.get("workload") != workload:
        return None, None
    return t.get("dram_bytes_per_launch"), t.get("source")


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle restatement of the reference on host cores
# ----------------------------------------------------------------------------------------------------------------
def _cpu_sample(args):
    """One worker = one subdomain type: ND-ordered sparse LU (SuperLU), `n_solves` star steps, and one real pass of
    the mortar projections of that subdomain (c2f of a mortar vector, f2c of a trace).  Returns seconds."""
    pattern, mortar, k, n_solves = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import mmmfe_oracle as mo
    prm = mo.Params()
    prm.mortar_degree = k
    sd = mo.Subdomain(0, 2, 2, pattern[0], pattern[1], pattern[2], prm, mo.DataDefault())
    t0 = time.time()
    sd.factorize("nd")
    t_factor = time.time() - t0
    rng = np.random.default_rng(0)
    lam = [rng.standard_normal(sd.n_side[s]) for s in range(4)]
    sd.neighbors = [1, 1, 1, 1]
    p_old = np.zeros(sd.np)
    t0 = time.time()
    for i in range(n_solves):
        x = sd.lu.solve(sd.rhs_star(lam, p_old, first=(i == 0)))
        p_old = x[sd.nf:]
    t_step = (time.time() - t0) / n_solves
    # project_mortar both ways on two sides (a 2x2 subdomain has two interfaces): sparse products with the sampling tables
    f2c, c2f = mo.projector_tables(sd, 1, mortar, k)
    q = rng.standard_normal(c2f.shape[1])
    t0 = time.time()
    for _ in range(2):
        tr = c2f @ q
        f2c @ tr
    t_proj = time.time() - t0
    return t_step, t_factor, sd.n, t_proj, c2f.shape[1]


def cpu_iteration_seconds(fine, coarse, mortar, k, n_sub, n_solves, arnoldi_k=10):
    """Seconds per GMRES iteration of the CPU restatement with one core per subdomain (what `mpirun -n j DarcyVT`
    can use: the reference runs exactly one rank per subdomain).  EXTRAPOLATED: `n_solves` star steps of one fine and
    one coarse subdomain are timed and multiplied by nt; the projections of one subdomain and a classical Gram-Schmidt
    update against `arnoldi_k` Krylov vectors are timed for real and added."""
    import multiprocessing as mp
    import numpy as np
    with mp.get_context("fork").Pool(2) as pool:
        (tf, ff, nf_, pf, lf), (tc, fc, nc_, pc, lc) = pool.map(
            _cpu_sample, [(fine, mortar, k, n_solves), (coarse, mortar, k, n_solves)])
    cores = min(n_sub, os.cpu_count() or 1)
    # ranks run concurrently; an iteration ends when the slowest rank has done its nt steps.  With fewer cores
    # than subdomains the ranks share cores.
    work = [fine[2] * tf + pf if i % 2 == 0 else coarse[2] * tc + pc for i in range(n_sub)]
    t_sweeps = max(max(work), sum(work) / cores)
    # Arnoldi on one rank's share of the interface vector (src/darcy_vtdd.cc:1413-1463)
    rng = np.random.default_rng(1)
    length = 2 * lf
    Q = rng.standard_normal((arnoldi_k + 1, length))
    w = rng.standard_normal(length)
    t0 = time.time()
    h = Q[:arnoldi_k] @ w
    w = w - h @ Q[:arnoldi_k]
    Q[arnoldi_k] = w / np.sqrt(w @ w)
    t_arnoldi = time.time() - t0
    t_iter = t_sweeps + t_arnoldi
    return t_iter, cores, dict(fine_s_per_step=tf, coarse_s_per_step=tc, fine_factor_s=ff, coarse_factor_s=fc,
                               fine_dofs=nf_, coarse_dofs=nc_, fine_projections_s=pf, coarse_projections_s=pc,
                               arnoldi_s=t_arnoldi, arnoldi_vectors=arnoldi_k)


def cpu_baseline_record(fine, coarse, mortar, k, n_sub, n_solves, dof_steps):
    t_iter, cores, detail = cpu_iteration_seconds(fine, coarse, mortar, k, n_sub, n_solves)
    return t_iter, {"value": dof_steps / t_iter, "unit": "DOF*timesteps/s", "cores": cores, "kind": "port",
                    "extrapolated": True, "timed_steps": n_solves,
                    "sample": f"oracle port (not the DarcyVT binary: deal.II/MPI/UMFPACK are not in the image): ND-ordered "
                              f"SuperLU factorisation + {n_solves} star steps of one fine and one coarse subdomain, multiplied by "
                              f"nt; one real pass of the mortar projections and one real Arnoldi update added; one core per "
                              f"subdomain", "detail": detail}


def init_distributed(pkg, rank, world, local_rank):
    """torch.distributed (NCCL) for the barriers/reductions of the harness, and the NCCL unique id of the engine's own
    communicator (rank 0 creates it, broadcast through torch).  Returns the id bytes, None on one rank."""
    if world <= 1:
        return None
    import ctypes
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    eng_lib = pkg.load_engine()
    n = eng_lib.mmmfe_comm_unique_id_size()
    buf = torch.zeros(n, dtype=torch.uint8, device="cuda")
    if rank == 0:
        raw = ctypes.create_string_buffer(n)
        assert eng_lib.mmmfe_comm_unique_id(ctypes.cast(raw, ctypes.c_void_p)) == 0
        buf.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
    dist.broadcast(buf, 0)
    return bytes(buf.cpu().numpy().tobytes())


def dof_of(p):
    return 3 * p[0] * p[1] + p[0] + p[1]


def workload(name, gpus):
    fine, coarse, mortar, k = CONFIGS[name][:4]
    per_gpu = CONFIGS[name][4] if len(CONFIGS[name]) > 4 else 4
    n_sub = per_gpu * max(1, gpus)
    mesh = checkerboard(n_sub, fine, coarse)
    dof_steps = sum(dof_of(m) * m[2] for m in mesh)
    n0, n1 = find_divisors(n_sub)
    text = (f"{name}: {n_sub} subdomains ({n0}x{n1}) checkerboard {fine[0]}^2x{fine[2]} / {coarse[0]}^2x{coarse[2]}, mortar "
            f"{mortar[0]}x{mortar[1]}x{mortar[2]} degree {k}, RT0xDGQ0, c0=1, T=0.5, K=I "
            f"(BASELINE.json configs[{dict(cfg3=2, cfg4=3).get(name, 1)}])")
    return fine, coarse, mortar, k, per_gpu, n_sub, mesh, dof_steps, text


def solve_record(r, dof_steps):
    """Time-to-solution of one refinement cycle from a front-end report that ran the converged solve."""
    setup = r["t_assemble"] + (r["t_factor"] - r["t_basis"]) + r["t_bar_data"]
    solve = r["t_bar"] + r["t_basis"] + r["t_solve_gmres"] + r["t_final"]
    its = max(1, r["solve_iterations"])
    return {"gmres_iterations_to_tol": r["solve_iterations"], "converged": bool(r["solve_converged"]),
            "final_relative_residual": r["solve_final_residual"],
            "operator": "multiscale basis (time-Toeplitz)" if r["basis_operator"] else "star sweeps",
            "gmres_s": r["t_solve_gmres"], "ms_per_iteration": 1e3 * r["t_solve_gmres"] / its,
            "basis_precompute_s": r["t_basis"], "bar_sweep_s": r["t_bar"], "final_sweep_s": r["t_final"],
            "time_to_solution_s": solve,
            "note": "solve_darcy_vt (src/darcy_vtdd.cc:881-895): bar sweep + [basis precompute] + GMRES to tol 1e-6 + final sweep",
            "setup_s": setup, "setup": {"assemble_s": r["t_assemble"], "factor_and_plan_s": r["t_factor"] - r["t_basis"],
                                        "bar_data_s": r["t_bar_data"]},
            "whole_cycle_s": setup + solve,
            "dof_steps_per_s_of_the_solve": dof_steps * its / max(r["t_solve_gmres"], 1e-12)}


def basis_record(r, pkg, local_rank):
    """The operator on the multiscale basis: precompute cost and the roofline of the block-Toeplitz kernel."""
    if not r["basis_operator"]:
        return None
    peak = pkg.fp64_tensor_peak_tflops(local_rank)
    ms = r["toeplitz_ms"]
    tf = r["toeplitz_gflop"] / ms if ms > 0 else 0.0  # GFLOP / ms = TFLOP/s
    hbm, _ = measured_peak_gbs()
    return {"columns_solved_on_this_rank": r["basis_columns"], "ranks_sharing_a_group": r["basis_shared_by"],
            "columns_per_pass": r["basis_mw"], "time_steps_solved": r["basis_steps"],
            "delays_with_data": r["basis_delays"], "precompute_s": r["t_basis"], "precompute_gflop": r["basis_gflop"],
            "precompute_tflops": r["basis_gflop"] / 1e3 / r["t_basis"] if r["t_basis"] > 0 else None,
            "toeplitz_kernel": {"bound": "tensor", "kernel": "k_toeplitz_apply (FP64 DMMA, mma.sync.m8n8k4.f64)",
                                "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak if peak else None,
                                "peak_source": "measured in this process: register-resident DMMA on all SMs "
                                               "(mmmfe_measure_fp64_tensor_tflops)",
                                "ms_per_launch": ms, "gflop_per_launch": r["toeplitz_gflop"],
                                "algorithmic_MB_per_launch": r["toeplitz_mbytes"],
                                "GBps_algorithmic": r["toeplitz_mbytes"] / ms if ms > 0 else None,
                                "hbm_peak_GBps": hbm, "tiles": r["toeplitz_tiles"]}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default=os.environ.get("MMMFE_BENCH_CONFIG", "cfg2"), choices=sorted(CONFIGS))
    ap.add_argument("--operator", default="auto", choices=["auto", "sweeps", "basis"],
                    help="interface operator of the timed GMRES loop (engine option `operator`)")
    ap.add_argument("--cpu-solves", type=int, default=4, help="star steps per subdomain type in the CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-solve", action="store_true", help="skip the converged solve (time-to-solution) of the N = 1 line")
    ap.add_argument("--no-north-star", action="store_true", help="skip the nested configs[3] record")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    fine, coarse, mortar, k, per_gpu, n_sub, mesh, dof_steps, text = workload(a.config, a.gpus)
    config = {"workload": text, "subdomains": n_sub, "dof_steps_per_iteration": dof_steps,
              "parallelism": f"{max(1, a.gpus)} x {per_gpu} subdomains per GPU",
              "l2": "factor stream per step (>= 0.7 GB) exceeds the 126 MB L2; no flush needed"}

    if a.impl == "reference":
        if rank != 0:
            return
        t0 = time.time()
        samples = [cpu_baseline_record(fine, coarse, mortar, k, n_sub, a.cpu_solves, dof_steps)
                   for _ in range(max(1, min(a.steps, 2)))]
        t_iter, base = min(samples, key=lambda s: s[0])
        val = dof_steps / t_iter
        print(json.dumps({"impl": "reference", "metric": "subdomain DOF*timesteps/s of the interface-GMRES loop",
                          "value": val, "unit": "DOF*timesteps/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                          "ms_per_step": 1e3 * t_iter, "gmres_iterations_per_s": 1.0 / t_iter, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                          "extrapolated": True, "cpu_baseline": base,
                          "e2e": {"value": val, "unit": "DOF*timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0, "wall_s": time.time() - t0}))
        return

    import torch
    import mmmfe_import
    pkg = mmmfe_import.load()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    nccl_id = init_distributed(pkg, rank, world, local_rank)
    dist = torch.distributed if world > 1 else None

    tmp = tempfile.mkdtemp(prefix="mmmfe_bench_")
    pfile = pkg.write_parameter_file(os.path.join(tmp, f"parameter_{rank}.txt"), mesh, mortar, mortar_degree=k, num_refinement=1)
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_wall = time.time()
    with_solve = a.gpus == 1 and not a.no_solve
    eng_opts = {"auto": "", "sweeps": "operator=0", "basis": "operator=1"}[a.operator]
    reps, _, _ = pkg.host_run(pfile, nccl_id=nccl_id, rank=rank, world=world, device=local_rank, compute_errors=0,
                              write_output=0, final_sweep=1 if with_solve else 0, skip_fetch=1, verbose=0, only_cycle=0,
                              fixed_iterations=a.steps, warmup_iterations=a.warmup, e2e_iterations=a.steps, profile=1,
                              also_solve=1 if with_solve else 0, engine_options=eng_opts)
    torch.cuda.synchronize()
    clocks = sampler.stop()
    r = reps[0]
    ms = r["gmres_device_ms"]
    if dist:
        t = torch.tensor([ms, r["e2e_ms_per_iteration"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_it = float(t[0]), float(t[1])
        launches = torch.tensor([r["gmres_kernel_launches"]], dtype=torch.int64, device="cuda")
        dist.all_reduce(launches)
        gpu_launches = int(launches[0])
        dist.barrier()
    else:
        e2e_it = r["e2e_ms_per_iteration"]
        gpu_launches = int(r["gmres_kernel_launches"])
    ms_per_step = ms / a.steps
    value = dof_steps / (ms_per_step * 1e-3)
    peak, peak_src = measured_peak_gbs()
    on_basis = bool(r["basis_operator"])
    out = None
    if rank == 0:
        # roofline of the dominant kernel (live CUDA-event timing inside the front-end run)
        names = ["rhs_build", "subtree_forward", "forward_top", "backward_top", "subtree_backward", "update_trace"]
        prof = {n: {"ms": r["prof_ms"][i], "MB": r["prof_bytes"][i] / 1e6, "launches": r["prof_launches"][i],
                    "GBps": (r["prof_bytes"][i] / 1e9) / (r["prof_ms"][i] * 1e-3) if r["prof_ms"][i] > 0 else 0.0}
                for i, n in enumerate(names)}
        tune = {"persistent_kernel_ms": r["tune_us_sweep"] / 1e3, "per_level_kernels_ms": r["tune_us_levels"] / 1e3,
                "compared_on": "the first 32 time levels of the longest sweep, every batch its share (engine option tune_steps)",
                "batches_with_sweep_plan": r["sweep_planned"],
                "relative_l2_difference_of_their_traces": r["tune_rel_diff_e18"] * 1e-18 if r["tune_rel_diff_e18"] >= 0 else None,
                "selected": "persistent sweep kernel" if r["sweep_batches"] > 0 else "per-level kernels"}
        basis = basis_record(r, pkg, local_rank)
        if on_basis:
            tk = basis["toeplitz_kernel"]
            roof = dict(tk)
            roof["note"] = ("the timed GMRES loop runs on the multiscale-basis operator: its dominant kernel is the "
                            "block-Toeplitz product; the star-sweep kernel (k_sweep) is not on this loop")
        else:
            if r["sweep_batches"] > 0:
                # the star sweeps run in the persistent kernel: one launch = all time steps of one batch
                sw = [{"ms": r["sweep_ms"][i], "steps": r["sweep_steps"][i], "MB_per_step": r["sweep_bytes_per_step"][i] / 1e6,
                       "GBps": r["sweep_bytes_per_step"][i] * r["sweep_steps"][i] / 1e9 / (r["sweep_ms"][i] * 1e-3)}
                      for i in range(r["sweep_batches"])]
                top = max(sw, key=lambda b: b["ms"])
                dom, achieved = "k_sweep (persistent star sweep, one launch per batch)", top["GBps"]
                per_launch_ms, bytes_per_launch = top["ms"], top["MB_per_step"] * 1e6 * top["steps"]
                prof = {"sweep_batches": sw, "per_level_classes_one_step": prof}
            else:
                dom = max(names, key=lambda n: prof[n]["ms"])
                per_launch_ms = prof[dom]["ms"] / max(1, prof[dom]["launches"])
                achieved = prof[dom]["GBps"]
                bytes_per_launch = prof[dom]["MB"] * 1e6 / max(1, prof[dom]["launches"])
            traffic, traffic_src = measured_traffic("k_sweep" if r["sweep_batches"] > 0 else dom, a.config)
            roof = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                    "frac_dram": (traffic / 1e9 / (per_launch_ms * 1e-3) / peak) if traffic else None,
                    "peak_source": peak_src, "bytes_per_launch": bytes_per_launch, "ms_per_launch": per_launch_ms,
                    "per_class": prof, "star_sweep_autotune": tune}
        out = {"metric": "subdomain DOF*timesteps/s of the interface-GMRES loop", "value": value, "unit": "DOF*timesteps/s",
               "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_per_step,
               "gmres_iterations_per_s": 1e3 / ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f64", "data": "synthetic", "config": config,
               "operator": {"requested": a.operator,
                            "used": "multiscale basis (time-Toeplitz, precomputed at set-up)" if on_basis else
                                    "star sweeps (nt backward-Euler solves per subdomain and iteration)",
                            "rule": "auto = basis when the mortar time mesh divides the subdomain's and a factor group has at "
                                    "most 2048 basis columns (engine option basis_auto_columns)"},
               "roofline": roof,
               "e2e": {"value": dof_steps / (e2e_it * 1e-3) if e2e_it > 0 else None, "unit": "DOF*timesteps/s",
                       "ms_per_step": e2e_it, "iterations_timed": a.steps, "h2d_bytes_per_step": r["e2e_h2d_bytes"],
                       "d2h_bytes_per_step": r["e2e_d2h_bytes"]},
               "gpu_launches": gpu_launches, "clocks": clocks,
               "setup": {"assemble_s": r["t_assemble"], "factor_s": r["t_factor"] - r["t_basis"],
                         "factor_gflops": r["factor_flops"] / 1e9, "basis_precompute_s": r["t_basis"],
                         "bar_data_s": r["t_bar_data"], "bar_sweep_s": r["t_bar"], "factor_groups": r["factor_groups"],
                         "factor_values": r["factor_values"], "step_bytes": r["step_bytes"],
                         "interface_length": r["interface_length"]}}
        if basis:
            out["basis_operator"] = basis
        if with_solve:
            out["time_to_solution"] = solve_record(r, dof_steps)
    # ---- nested record on BASELINE.json configs[3] (the north-star problem: 64 subdomains on 8 GPUs; one GPU's share of
    # it -- 8 subdomains -- otherwise): converged solve on the operator the engine selects there (the multiscale basis)
    if not a.no_north_star and a.config == "cfg2" and a.gpus in (1, 8):
        f4, c4, m4, k4, pg4, ns4, mesh4, ds4, text4 = workload("cfg4", a.gpus)
        p4 = pkg.write_parameter_file(os.path.join(tmp, f"north_star_{rank}.txt"), mesh4, m4, mortar_degree=k4, num_refinement=1)
        nid = init_distributed(pkg, rank, world, local_rank)  # a fresh communicator id for the second front-end run
        r4 = pkg.host_run(p4, nccl_id=nid, rank=rank, world=world, device=local_rank, compute_errors=0, write_output=0,
                          final_sweep=1, skip_fetch=1, verbose=0, only_cycle=0, fixed_iterations=20, warmup_iterations=3,
                          profile=1, also_solve=1)[0][0]
        vals = [r4[key] for key in ("t_assemble", "t_factor", "t_basis", "t_bar_data", "t_bar", "t_solve_gmres", "t_final",
                                    "gmres_device_ms", "toeplitz_ms")]
        if dist:
            t = torch.tensor(vals, dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            vals = [float(v) for v in t]
            dist.barrier()
        if rank == 0:
            for key, v in zip(("t_assemble", "t_factor", "t_basis", "t_bar_data", "t_bar", "t_solve_gmres", "t_final",
                               "gmres_device_ms", "toeplitz_ms"), vals):
                r4[key] = v
            rec = solve_record(r4, ds4)
            rec["workload"] = text4 + (" -- ONE GPU's share of the 64-subdomain problem" if a.gpus == 1 else "")
            rec["tie_rule"] = (os.environ.get("MMMFE_TIE_RULE") or "lower") + (
                " (quadratic mortar: centre Gauss points on fine cell boundaries sample the lower cell / earlier level, "
                "the exact-arithmetic limit of deal.II's point location; MMMFE_TIE_RULE=dealii91 = its floating-point "
                "choice, under which the operator is not time-Toeplitz and the engine falls back to the star sweeps; "
                "DESIGN.md section 6)")
            rec["ms_per_iteration_device_timed_20_fixed"] = r4["gmres_device_ms"] / 20
            rec["gmres_solves_per_s"] = 1.0 / max(r4["t_solve_gmres"], 1e-12)
            rec["solve_darcy_vt_per_s"] = 1.0 / max(rec["time_to_solution_s"], 1e-12)
            rec["basis_operator"] = basis_record(r4, pkg, local_rank)
            out["north_star"] = rec
    if rank != 0:
        return
    out["wall_s"] = time.time() - t_wall
    if a.gpus == 1 and not a.no_cpu_baseline:
        _, out["cpu_baseline"] = cpu_baseline_record(fine, coarse, mortar, k, n_sub, a.cpu_solves, dof_steps)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
