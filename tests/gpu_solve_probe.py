"""Time-to-solution probe: one DarcyVT cycle of a bench configuration with the converged GMRES solve (tolerance of the
parameter file), bar sweep and final sweep timed; prints one JSON line on rank 0.
Usage: [torchrun ...] gpu_solve_probe.py cfg2 [engine options]     (N ranks: per_gpu * N subdomains, like bench.py)"""
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import mmmfe_import  # noqa: E402

if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    opts = sys.argv[2] if len(sys.argv) > 2 else ""
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    fine, coarse, mortar, k = bench.CONFIGS[name][:4]
    per_gpu = bench.CONFIGS[name][4] if len(bench.CONFIGS[name]) > 4 else 4
    mesh = bench.checkerboard(per_gpu * world, fine, coarse)
    pkg = mmmfe_import.load()
    import torch
    torch.cuda.set_device(local)
    nccl_id = bench.init_distributed(pkg, rank, world, local)
    tmp = tempfile.mkdtemp(prefix="mmmfe_solve_")
    p = pkg.write_parameter_file(os.path.join(tmp, "parameter.txt"), mesh, mortar, mortar_degree=k, num_refinement=1)
    reps, res, _ = pkg.host_run(p, nccl_id=nccl_id, rank=rank, world=world, device=local, verbose=0, write_output=0,
                                compute_errors=0, final_sweep=1, skip_fetch=1, fixed_iterations=3, warmup_iterations=1,
                                also_solve=0 if os.environ.get("MMMFE_PROBE_NOSOLVE") else 1, engine_options=opts, apply_repeat=10)
    r = reps[0]
    keys = ["t_assemble", "t_factor", "t_bar_data", "t_bar", "t_gmres", "t_solve_gmres", "t_final", "solve_iterations",
            "solve_converged", "solve_final_residual", "gmres_device_ms", "sweep_batches", "interface_length",
            "basis_operator", "basis_shared_by", "t_basis", "basis_columns", "basis_gflop", "tune_rel_diff_e18", "apply_ms"]
    if world > 1:
        torch.distributed.barrier()
    if rank == 0:
        print(json.dumps({"config": name, "world": world, "subdomains": len(mesh), "options": opts, **{k_: r[k_] for k_ in keys}}))
    if world > 1:
        torch.distributed.destroy_process_group()
