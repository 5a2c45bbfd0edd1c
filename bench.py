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
    """One worker = one subdomain type: ND-ordered sparse LU (SuperLU) + `n_solves` star steps; returns s/step."""
    pattern, n_solves = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import mmmfe_oracle as mo
    prm = mo.Params()
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
    return (time.time() - t0) / n_solves, t_factor, sd.n


def cpu_iteration_seconds(fine, coarse, n_sub, n_solves):
    """Seconds per GMRES iteration of the CPU restatement with one core per subdomain (what `mpirun -n j DarcyVT`
    can use: the reference runs exactly one rank per subdomain), extrapolated from a bounded sample of star steps."""
    import multiprocessing as mp
    with mp.get_context("fork").Pool(2) as pool:
        (tf, ff, nf_), (tc, fc, nc_) = pool.map(_cpu_sample, [(fine, n_solves), (coarse, n_solves)])
    cores = min(n_sub, os.cpu_count() or 1)
    # ranks run concurrently; an iteration ends when the slowest rank has done its nt steps.  With fewer cores
    # than subdomains the ranks share cores.
    work = [fine[2] * tf if i % 2 == 0 else coarse[2] * tc for i in range(n_sub)]
    t_iter = max(max(work), sum(work) / cores)
    return t_iter, cores, dict(fine_s_per_step=tf, coarse_s_per_step=tc, fine_factor_s=ff, coarse_factor_s=fc,
                               fine_dofs=nf_, coarse_dofs=nc_)


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default=os.environ.get("MMMFE_BENCH_CONFIG", "cfg2"), choices=sorted(CONFIGS))
    ap.add_argument("--cpu-solves", type=int, default=4, help="star steps per subdomain type in the CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    fine, coarse, mortar, k = CONFIGS[a.config][:4]
    per_gpu = CONFIGS[a.config][4] if len(CONFIGS[a.config]) > 4 else 4
    n_sub = per_gpu * max(1, a.gpus)
    mesh = checkerboard(n_sub, fine, coarse)
    dof = lambda p: 3 * p[0] * p[1] + p[0] + p[1]
    dof_steps = sum(dof(m) * m[2] for m in mesh)
    config = {"workload": f"{a.config}: {n_sub} subdomains ({find_divisors(n_sub)[0]}x{find_divisors(n_sub)[1]}) checkerboard "
                          f"{fine[0]}^2x{fine[2]} / {coarse[0]}^2x{coarse[2]}, mortar {mortar[0]}x{mortar[1]}x{mortar[2]} degree {k}, "
                          f"RT0xDGQ0, c0=1, T=0.5, K=I (BASELINE.json configs[{dict(cfg3=2, cfg4=3).get(a.config, 1)}])",
              "subdomains": n_sub, "dof_steps_per_iteration": dof_steps, "parallelism": f"{max(1, a.gpus)} x {per_gpu} subdomains per GPU",
              "l2": "factor stream per step (>= 0.7 GB) exceeds the 126 MB L2; no flush needed"}

    if a.impl == "reference":
        if rank != 0:
            return
        t0 = time.time()
        samples = []
        for _ in range(max(1, min(a.steps, 2))):
            samples.append(cpu_iteration_seconds(fine, coarse, n_sub, a.cpu_solves))
        t_iter, cores, detail = min(samples, key=lambda s: s[0])
        val = dof_steps / t_iter
        base = {"value": val, "unit": "DOF*timesteps/s", "cores": cores, "kind": "port",
                "sample": f"ND-ordered SuperLU factorisation + {a.cpu_solves} star steps of one fine and one coarse subdomain, "
                          f"extrapolated to nt steps with one core per subdomain (oracle restatement, not the DarcyVT binary)",
                "detail": detail}
        print(json.dumps({"impl": "reference", "metric": "subdomain DOF*timesteps/s of the interface-GMRES loop",
                          "value": val, "unit": "DOF*timesteps/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                          "ms_per_step": 1e3 * t_iter, "gmres_iterations_per_s": 1.0 / t_iter, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                          "cpu_baseline": base, "e2e": {"value": val, "unit": "DOF*timesteps/s", "h2d_bytes_per_step": 0,
                                                        "d2h_bytes_per_step": 0},
                          "gpu_launches": 0, "wall_s": time.time() - t0}))
        return

    import numpy as np
    import torch
    import mmmfe_import
    pkg = mmmfe_import.load()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    nccl_id = init_distributed(pkg, rank, world, local_rank)

    tmp = tempfile.mkdtemp(prefix="mmmfe_bench_")
    pfile = pkg.write_parameter_file(os.path.join(tmp, "parameter.txt"), mesh, mortar, mortar_degree=k, num_refinement=1)
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_wall = time.time()
    reps, _, _ = pkg.host_run(pfile, nccl_id=nccl_id, rank=rank, world=world, device=local_rank, compute_errors=0,
                              write_output=0, final_sweep=0, verbose=0, only_cycle=0, fixed_iterations=a.steps,
                              warmup_iterations=a.warmup, e2e_iterations=min(a.steps, 3), profile=1)
    torch.cuda.synchronize()
    clocks = sampler.stop()
    r = reps[0]
    ms = r["gmres_device_ms"]
    e2e_ms = r["e2e_ms_per_iteration"] * min(a.steps, 3)
    if world > 1:
        t = torch.tensor([ms, r["e2e_ms_per_iteration"]], dtype=torch.float64, device="cuda")
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms, e2e_it = float(t[0]), float(t[1])
        launches = torch.tensor([r["gmres_kernel_launches"]], dtype=torch.int64, device="cuda")
        torch.distributed.all_reduce(launches)
        gpu_launches = int(launches[0])
        torch.distributed.barrier()
    else:
        e2e_it = r["e2e_ms_per_iteration"]
        gpu_launches = int(r["gmres_kernel_launches"])
    if rank != 0:
        return
    ms_per_step = ms / a.steps
    value = dof_steps / (ms_per_step * 1e-3)
    peak, peak_src = measured_peak_gbs()
    # roofline of the dominant kernel class of one star step (live CUDA-event timing, mmmfe_profile_step)
    names = ["rhs_build", "subtree_forward", "forward_top", "backward_top", "subtree_backward", "update_trace"]
    prof = {n: {"ms": r["prof_ms"][i], "MB": r["prof_bytes"][i] / 1e6, "launches": r["prof_launches"][i],
                "GBps": (r["prof_bytes"][i] / 1e9) / (r["prof_ms"][i] * 1e-3) if r["prof_ms"][i] > 0 else 0.0}
            for i, n in enumerate(names)}
    tune = {"persistent_kernel_ms": r["tune_us_sweep"] / 1e3, "per_level_kernels_ms": r["tune_us_levels"] / 1e3,
            "batches_with_sweep_plan": r["sweep_planned"],
            "selected": "persistent sweep kernel" if r["sweep_batches"] > 0 else "per-level kernels"}
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
    out = {"metric": "subdomain DOF*timesteps/s of the interface-GMRES loop", "value": value, "unit": "DOF*timesteps/s",
           "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_per_step,
           "gmres_iterations_per_s": 1e3 / ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic", "config": config,
           "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                        "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                        "bytes_per_launch": bytes_per_launch, "ms_per_launch": per_launch_ms, "per_class": prof,
                        "star_sweep_autotune": tune},
           "e2e": {"value": dof_steps / (e2e_it * 1e-3) if e2e_it > 0 else None, "unit": "DOF*timesteps/s",
                   "ms_per_step": e2e_it, "h2d_bytes_per_step": r["e2e_h2d_bytes"], "d2h_bytes_per_step": r["e2e_d2h_bytes"]},
           "gpu_launches": gpu_launches, "clocks": clocks,
           "setup": {"assemble_s": r["t_assemble"], "factor_s": r["t_factor"], "factor_gflops": r["factor_flops"] / 1e9,
                     "bar_data_s": r["t_bar_data"], "bar_sweep_s": r["t_bar"], "factor_groups": r["factor_groups"],
                     "factor_values": r["factor_values"], "step_bytes": r["step_bytes"],
                     "interface_length": r["interface_length"]},
           "wall_s": time.time() - t_wall}
    if a.gpus == 1 and not a.no_cpu_baseline:
        t_iter, cores, detail = cpu_iteration_seconds(fine, coarse, n_sub, a.cpu_solves)
        out["cpu_baseline"] = {"value": dof_steps / t_iter, "unit": "DOF*timesteps/s", "cores": cores, "kind": "port",
                               "sample": f"ND-ordered SuperLU factorisation + {a.cpu_solves} star steps of one fine and one coarse "
                                         f"subdomain, extrapolated to nt steps, one core per subdomain", "detail": detail}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
