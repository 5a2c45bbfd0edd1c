"""Step-by-step GPU bring-up script (prints diagnostics instead of stopping at the first failure)."""
import sys
import time
import traceback

import numpy as np

from helpers import default_params, engine_from_oracle, make_problem, rel_l2, set_bar_data


def step(name, fn):
    t0 = time.time()
    try:
        out = fn()
        print(f"[ok] {name} ({time.time() - t0:.2f}s): {out}", flush=True)
    except Exception:
        print(f"[FAIL] {name}", flush=True)
        traceback.print_exc()
        sys.stdout.flush()


def saddle(nx, ny):
    prm = default_params()
    prob = make_problem(prm, n_sub=2, mesh=[[nx, ny, 2], [nx, ny, 2]], mortar=[1, 1, 1])
    eng = engine_from_oracle(prob)
    eng.setup()
    sd = prob.subs[0]
    rhs = np.random.default_rng(1).standard_normal(sd.n)
    x = eng.solve_saddle(0, rhs)
    ref = sd.lu.solve(rhs)
    st = eng.stats()
    eng.close()
    return dict(err=rel_l2(x, ref), factor_s=st["factor_seconds"], vals=st["factor_values"])


def operator(k, cycle):
    prm = default_params(k, cycle + 1)
    prob = make_problem(prm, cycle)
    eng = engine_from_oracle(prob)
    eng.setup()
    q = np.random.default_rng(3).standard_normal(prob.n_iface)
    a = eng.apply(q)
    b = prob.apply_S(q)
    eng.close()
    return dict(err=rel_l2(a, b), n=prob.n_iface)


def full(k, cycle):
    prm = default_params(k, cycle + 1)
    prob = make_problem(prm, cycle)
    ref = prob.run_cycle(cycle)
    eng = engine_from_oracle(prob)
    eng.setup()
    set_bar_data(eng, prob)
    eng.bar_sweep()
    out = eng.gmres(prm.tolerance, prm.max_iteration)
    eng.final_sweep()
    ep = max(rel_l2(eng.get_solution(i)[:, sd.nf:], ref.solution_p[i]) for i, sd in enumerate(prob.subs))
    eu = max(rel_l2(eng.get_solution(i)[:, :sd.nf], ref.solution_u[i]) for i, sd in enumerate(prob.subs))
    eng.close()
    return dict(its=(out["iterations"], ref.gmres_iterations), r_norm=(out["r_norm"], ref.r_norm),
                lam=rel_l2(out["lam"], ref.lam), p=ep, u=eu, gmres_s=out["seconds"])


if __name__ == "__main__":
    for shp in [(2, 2), (3, 3), (5, 7), (16, 16), (64, 64), (100, 90), (256, 256)]:
        step(f"saddle {shp}", lambda: saddle(*shp))
    for k, c in [(1, 0), (2, 0), (1, 2)]:
        step(f"operator k={k} cycle={c}", lambda: operator(k, c))
    for k, c in [(1, 0), (1, 2)]:
        step(f"full k={k} cycle={c}", lambda: full(k, c))
