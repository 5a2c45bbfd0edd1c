"""Quick device timing of one interface-operator application for a 2x2 checkerboard (no oracle LU)."""
import sys
import time

import numpy as np

from helpers import default_params, mo, pkg


def build(nf, ntf, nc, ntc, mortar, k=1):
    prm = default_params(k)
    mesh = [[nf, nf, ntf], [nc, nc, ntc], [nc, nc, ntc], [nf, nf, ntf]]
    prob = mo.Problem(prm, 4, mesh, mortar)
    t0 = time.time()
    for sd in prob.subs:
        sd.matrix = sd.assemble_matrix()
        sd.f2c, sd.c2f = {}, {}
        for side in range(4):
            if sd.neighbors[side] >= 0:
                sd.f2c[side], sd.c2f[side] = mo.projector_tables(sd, side, prob.mortar, prob.k)
    print(f"host assembly {time.time() - t0:.1f}s", flush=True)
    return prob


if __name__ == "__main__":
    nf, ntf, nc, ntc, ms, mt = [int(a) for a in sys.argv[1:7]]
    reps = int(sys.argv[7]) if len(sys.argv) > 7 else 3
    prob = build(nf, ntf, nc, ntc, [ms, ms, mt])
    eng = pkg().Engine()
    import os
    for key in ("subtree_cells", "leaf_cells", "use_graphs", "use_sweep", "sweep_chunk", "sweep_min_fill", "sweep_evict_first", "sweep_compute_threads", "sweep_ctas_per_sm", "sweep_spin_ns", "sweep_groups", "sweep_autotune", "sweep_micro_tile", "sweep_pack", "sweep_share_blobs", "sweep_concurrent", "sweep_share_exponent"):
        if os.environ.get("MMMFE_" + key.upper()):
            eng.set_option(key, float(os.environ["MMMFE_" + key.upper()]))
    for sd in prob.subs:
        eng.add_subdomain(global_id=sd.r, nx=sd.nx, ny=sd.ny, nt=sd.nt, dt=sd.dt, matrix=sd.matrix.tocsr(),
                          n_flux=sd.nf, n_pressure=sd.np, neighbor=sd.neighbors,
                          side_dofs=[np.asarray(e) for e in sd.side_edges], f2c=sd.f2c, c2f=sd.c2f)
    t0 = time.time()
    eng.setup()
    st = eng.stats()
    print(f"setup {time.time() - t0:.2f}s  stats {st}", flush=True)
    q = np.random.default_rng(0).standard_normal(prob.n_iface)
    t0 = time.time()
    ap = eng.apply(q)
    print(f"first apply (graph capture) {time.time() - t0:.3f}s  |ap| {np.linalg.norm(ap):.6e}", flush=True)
    ms_ = eng.apply_device(reps)
    steps = sum(sd.nt for sd in prob.subs)
    # bytes per apply: every batch streams its factor nt times
    print(f"apply {ms_:.3f} ms  ({1e3 / ms_:.2f} applies/s)  step_bytes(all batches, 1 step) {st['step_bytes'] / 1e6:.1f} MB")
    dofsteps = sum(sd.n * sd.nt for sd in prob.subs)
    print(f"DOF*steps/s {dofsteps / (ms_ * 1e-3):.3e}")
    print("sweep batches", eng.stat("sweep_batches"), "of", eng.stat("batches"), "ctas", eng.stat("sweep_ctas"),
          "smem", eng.stat("sweep_smem_bytes"), "entries", eng.stat("sweep_entries"), "chunk", eng.stat("sweep_chunk_used"),
          "ring", eng.stat("sweep_ring_doubles"))
    for ms_sw, by, ns in eng.profile_sweeps():
        print(f"  sweep kernel: {ms_sw:9.3f} ms for {ns} steps = {ms_sw / ns * 1e3:8.1f} us/step, "
              f"{by / 1e6:8.1f} MB/step -> {by * ns / ms_sw / 1e6:7.0f} GB/s")
    for bi in range(eng.stat("sweep_batches")):
        cyc = eng.profile_sweep_cycles(bi)
        print("  cycles/CTA:", {k: round(v) for k, v in cyc.items()})
    if os.environ.get("MMMFE_TRACE_ALL"):
        cb, kh, st = eng.sweep_trace(0)
        np.savez(os.environ["MMMFE_TRACE_ALL"], cb=cb, kh=kh, st=st)
    if os.environ.get("MMMFE_TRACE"):
        cb, kh, st = eng.sweep_trace(0)
        for cta in [int(x) for x in os.environ["MMMFE_TRACE"].split(",")]:
            b0, b1 = cb[cta], cb[cta + 1]
            t0 = st[b0:b1, 0][st[b0:b1, 0] > 0].min()
            print(f"--- cta {cta}: entries {b1 - b0}")
            for i in range(b0, b1):
                if st[i, 0] == 0:
                    continue
                print(f"  {i - b0:4d} kind {kh[i, 0]} h {kh[i, 1]:2d} start {st[i, 0] - t0:8d} wait {st[i, 1] - st[i, 0]:7d} "
                      f"run {st[i, 2] - st[i, 1]:7d}")
    for k, v in eng.profile_step().items():
        print(f"  {k:18s} {v['ms']*1e3:8.1f} us  {v['bytes']/1e6:8.1f} MB  {v['launches']:3d} launches  {v['bytes']/1e9/max(v['ms'],1e-9)*1e3:7.0f} GB/s")
    eng.close()
