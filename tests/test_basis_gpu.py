"""The multiscale-flux-basis / time-Toeplitz interface operator (engine option operator = 1; SURVEY.md 8(f)1, the
reference's dormant compute_multiscale_basis, src/darcy_vtdd.cc:959-1003) against the oracle and against the sweep
operator: it must be the SAME operator to round-off, whatever the tree shape, the number of passes or the mortar."""
import numpy as np
import pytest

from helpers import default_params, engine_from_oracle, make_problem, mo, rel_l2, set_bar_data

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _two_by_two(nxs, nts, mortar, k=1):
    prm = default_params(k)
    mesh = [[n, n, t] for n, t in zip(nxs, nts)]
    return make_problem(prm, n_sub=4, mesh=mesh, mortar=mortar)


def _basis_engine(prob, extra=None, **kw):
    opts = {"operator": 1, "mortar_time_slabs": prob.mortar[2]}
    opts.update(extra or {})
    eng = engine_from_oracle(prob, options=opts, **kw)
    eng.setup()
    assert eng.stat("basis_operator") == 1
    return eng


def _dense_oracle_operator(prob, sd, meta):
    """T[d][row][col] of one subdomain from the oracle: star sweeps of unit mortar vectors of the first time slab."""
    m_t, k = prob.mortar[2], prob.k
    nd = (k + 1) ** 2
    sides = [s for s in range(4) if sd.neighbors[s] >= 0]
    P = {s: prob.offsets[(sd.r, s)][1] // m_t for s in sides}
    T = np.zeros((m_t, meta["P_tot"], meta["P_tot"]))
    for s in sides:
        for a in range(P[s]):
            lam = {t: np.zeros(prob.offsets[(sd.r, t)][1]) for t in sides}
            lam[s][a] = 1.0
            tr, _ = prob.star_sweep(sd, lam)
            for so in sides:
                out = mo.NORMAL_SIGN[so] * (sd.f2c[so] @ tr[so])
                T[:, meta["base_of_side"][so]:meta["base_of_side"][so] + P[so], meta["base_of_side"][s] + a] = out.reshape(m_t, P[so])
    return T, sides, P


@pytest.mark.parametrize("nxs,nts,mortar,k", [
    ((8, 4, 4, 8), (8, 4, 4, 8), [2, 2, 2], 1),
    ((12, 6, 6, 12), (6, 3, 3, 6), [3, 3, 3], 2),
])
def test_basis_blocks_match_the_dense_oracle_operator(nxs, nts, mortar, k):
    prob = _two_by_two(nxs, nts, mortar, k)
    eng = _basis_engine(prob)
    for li in (0, 1):
        sd = prob.subs[li]
        meta, T = eng.get_basis(li)
        ref, sides, P = _dense_oracle_operator(prob, sd, meta)
        # the group operator also holds the sides of the OTHER member of the factor group: compare this member's blocks
        for so in sides:
            for si in sides:
                r0, c0 = meta["base_of_side"][so], meta["base_of_side"][si]
                blk, rblk = T[:, r0:r0 + P[so], c0:c0 + P[si]], ref[:, r0:r0 + P[so], c0:c0 + P[si]]
                assert np.max(np.abs(blk - rblk)) < TOL * np.max(np.abs(ref)), (li, so, si)
    eng.close()


@pytest.mark.parametrize("nxs,nts,mortar,k,extra", [
    ((3, 2, 4, 3), (3, 2, 4, 3), [1, 1, 1], 1, {}),                              # one subtree, one slab
    ((16, 8, 8, 16), (8, 4, 4, 8), [4, 4, 2], 1, {"basis_columns": 8}),           # several passes
    ((37, 23, 23, 37), (6, 3, 3, 6), [5, 5, 3], 2, {"subtree_cells": 64}),        # ragged boxes, quadratic mortar
    ((96, 48, 48, 96), (8, 4, 4, 8), [12, 12, 2], 1, {"subtree_cells": 64}),      # several top levels
    ((48, 24, 24, 48), (8, 4, 4, 8), [6, 6, 4], 1, {"subtree_cells": 0}),         # no subtrees: every front is a GEMM
    ((40, 20, 20, 40), (12, 6, 6, 12), [5, 5, 6], 2, {"basis_columns": 24}),      # passes that do not divide the columns
])
def test_basis_operator_is_the_sweep_operator(nxs, nts, mortar, k, extra):
    prob = _two_by_two(nxs, nts, mortar, k)
    q = np.random.default_rng(41).standard_normal(prob.n_iface)
    ref = prob.apply_S(q)
    eng = _basis_engine(prob, extra)
    out = eng.apply(q)
    assert np.array_equal(out, eng.apply(q))  # deterministic
    assert rel_l2(out, ref) < TOL
    eng.close()
    sw = engine_from_oracle(prob, options={"operator": 0, **{k_: v for k_, v in extra.items() if k_ == "subtree_cells"}})
    sw.setup()
    assert sw.stat("basis_operator") == 0
    assert rel_l2(out, sw.apply(q)) < 1e-11
    sw.close()


def test_basis_operator_checkerboards():
    """4x4 linear and 8x8 quadratic checkerboards (configs[2], configs[3] scaled down): two factor groups whose members
    have different interface sides (corner, edge and interior subdomains share one group operator)."""
    prm = default_params(1)
    mesh = [[16, 16, 8] if ((r % 4) + (r // 4)) % 2 == 0 else [8, 8, 4] for r in range(16)]
    prob = make_problem(prm, n_sub=16, mesh=mesh, mortar=[4, 4, 2])
    eng = _basis_engine(prob)
    assert eng.stats()["factor_groups"] == 2
    q = np.random.default_rng(6).standard_normal(prob.n_iface)
    assert rel_l2(eng.apply(q), prob.apply_S(q)) < TOL
    eng.close()
    prm = default_params(2)
    mesh = [[8, 8, 16] if ((r % 8) + (r // 8)) % 2 == 0 else [4, 4, 8] for r in range(64)]
    prob = make_problem(prm, n_sub=64, mesh=mesh, mortar=[2, 2, 4])
    eng = _basis_engine(prob)
    q = np.random.default_rng(8).standard_normal(prob.n_iface)
    assert rel_l2(eng.apply(q), prob.apply_S(q)) < TOL
    eng.close()


def test_basis_passes_stop_when_the_response_has_died_out():
    """8x8 subdomains of size 1/8: the slowest mode of a star problem decays by 1/(1 + 2 pi^2 dt / L^2) per step, so the
    response to first-slab data falls below 1e-18 of its peak long before the final time; the remaining steps are
    skipped and the Toeplitz sums stop at the last delay that carries data -- same operator to round-off."""
    prm = default_params(1)
    mesh = [[8, 8, 32] if ((r % 8) + (r // 8)) % 2 == 0 else [4, 4, 16] for r in range(64)]
    prob = make_problem(prm, n_sub=64, mesh=mesh, mortar=[2, 2, 8])
    q = np.random.default_rng(12).standard_normal(prob.n_iface)
    ref = prob.apply_S(q)
    eng = _basis_engine(prob)
    assert 1 <= eng.stat("basis_delays") < 8
    assert eng.stat("basis_steps") < eng.stat("basis_passes") // 2 * (32 + 16)
    assert rel_l2(eng.apply(q), ref) < TOL
    eng.close()
    eng = _basis_engine(prob, {"basis_drop_tol": 0})  # switched off: all delays, all steps
    assert eng.stat("basis_delays") == 8
    assert rel_l2(eng.apply(q), ref) < TOL
    eng.close()


def test_basis_operator_with_permuted_dofs_and_neumann_sides():
    prm = default_params(1, 2)
    prob = make_problem(prm, 1)
    q = np.random.default_rng(9).standard_normal(prob.n_iface)
    eng = _basis_engine(prob, perm_seed=21)
    assert rel_l2(eng.apply(q), prob.apply_S(q)) < TOL
    eng.close()
    prm = default_params(1, 2)
    prm.is_manufact, prm.bc_type, prm.bc_value = False, ["N", "D", "D", "N"], [0.5, 1.0, 2.0, -0.25]
    prob = make_problem(prm, 1)
    eng = _basis_engine(prob)  # constrained matrices: no RT0 stencil, generic kernels + basis
    assert rel_l2(eng.apply(q), prob.apply_S(q)) < TOL
    eng.close()


@pytest.mark.parametrize("k,cycle", [(1, 1), (2, 0), (1, 2)])
def test_full_solve_on_the_basis_operator_matches_oracle(k, cycle):
    prm = default_params(k, cycle + 1)
    prob = make_problem(prm, cycle)
    ref = prob.run_cycle(cycle)
    eng = _basis_engine(prob)
    set_bar_data(eng, prob)
    eng.bar_sweep()
    out = eng.gmres(prm.tolerance, prm.max_iteration)
    assert out["iterations"] == ref.gmres_iterations
    assert out["r_norm"] == pytest.approx(ref.r_norm, rel=1e-12)
    np.testing.assert_allclose(out["residuals"], ref.residuals, rtol=1e-6, atol=1e-13)
    assert rel_l2(out["lam"], ref.lam) < TOL
    eng.final_sweep()
    for li, sd in enumerate(prob.subs):
        sol = eng.get_solution(li)
        assert rel_l2(sol[:, sd.nf:], ref.solution_p[li]) < TOL
        assert rel_l2(sol[:, :sd.nf], ref.solution_u[li]) < TOL
    eng.close()


def test_basis_needs_a_mortar_time_mesh_that_divides_the_time_steps():
    prob = _two_by_two((8, 4, 4, 8), (6, 3, 3, 6), [2, 2, 4])  # 3 steps against 4 mortar slabs
    eng = engine_from_oracle(prob, options={"operator": 1, "mortar_time_slabs": 4})
    with pytest.raises(Exception, match="multiscale basis"):
        eng.setup()
    eng.close()
    eng = engine_from_oracle(prob, options={"operator": 2, "mortar_time_slabs": 4})
    eng.setup()  # auto: falls back to the sweeps
    assert eng.stat("basis_operator") == 0
    q = np.random.default_rng(2).standard_normal(prob.n_iface)
    assert rel_l2(eng.apply(q), prob.apply_S(q)) < TOL
    eng.close()
