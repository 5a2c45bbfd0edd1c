"""Pins the CPU oracle against the only golden numbers the reference publishes (report.pdf Tables 2, 4)."""
import json
import os

import numpy as np
import pytest

import mmmfe_oracle as mo

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = json.load(open(os.path.join(HERE, "golden", "report_tables.json")))
COLS = ["velocity_l2l2", "pressure_dg", "pressure_l2l2", "lambda_l2"]


def _params(mortar_degree, cycles):
    prm = mo.parse_parameter_file(os.path.join(ROOT, "parameter.txt"))
    prm.mortar_degree = mortar_degree
    prm.num_refinement = cycles
    return prm


def _check_row(res, row, its_slack):
    for c in COLS:
        got = float(f"{res.errors[c]:.3e}")  # the table prints 4 significant digits
        assert got == pytest.approx(row[c], rel=1.01e-3 / 1.0), (c, res.errors[c], row[c])
    assert abs(res.gmres_iterations - row["gmres"]) <= its_slack


def test_table2_linear_mortar_cycles_0_to_3():
    """Every printed digit of all four error columns; GMRES count within the +-1 the contract allows.

    The current source stops at residual_k < tol, the March-2020 table at residual_k / r_norm < tol
    (test_published_gmres_counts_follow_from_the_residual_history): one fewer in cycles 0-2 where r_norm > 1.
    """
    res = mo.run(_params(1, 4))
    for r, row in zip(res, GOLD["table2_linear_mortar"]["rows"]):
        _check_row(r, row, its_slack=1)


def test_table2_linear_mortar_cycle_4():
    res = mo.run(_params(1, 5), cycles=[4])
    # errors match in every digit; 82 iterations against 86 published: the report-era stop test was scaled by
    # r_norm once more (test_published_count_of_table2_row5_follows_from_the_residual_history reproduces the 86).
    _check_row(res[0], GOLD["table2_linear_mortar"]["rows"][4], its_slack=4)


def _first_below(res, threshold):
    return next(k for k in range(1, len(res.residuals)) if res.residuals[k] < threshold)


def _both_counts(mortar_degree, cycle, tol=1e-6):
    """(count of the current source, count under the report-era stop rule) from ONE residual history: the run uses a
    ten times smaller tolerance so that the history reaches past both stops."""
    prm = _params(mortar_degree, cycle + 1)
    prm.tolerance = 0.1 * tol
    res = mo.run(prm, cycles=[cycle])[0]
    # current source: |Beta[k+1]| / r_norm < tol (src/darcy_vtdd.cc:1478, 1488)
    # report era:     |Beta[k+1]| / r_norm^2 < tol -- the relative residual divided once more by r_norm
    return _first_below(res, tol), _first_below(res, tol * res.r_norm), res


@pytest.mark.parametrize("cycle", [0, 1, 2, 3])
def test_published_gmres_counts_follow_from_the_residual_history(cycle):
    """report.pdf Table 2 prints 11/23/39/59/86 iterations, the current source gives 12/24/40/59/82.  The difference
    is ONE scaling in the stop test and nothing else: the first k with residual_k / r_norm < tol (i.e. the scaled
    residual of :1478 divided once more by r_norm; r_norm = 4.13, 2.30, 1.50, 0.92, 0.60 over the cycles, which is
    why the published count is smaller in cycles 0-2, equal in cycle 3 and larger in cycle 4) is the published
    count in every published row -- 5 of Table 2 here, 3 of Table 4 (rows 3 and 5 with the deal.II tie rule) in
    tests/test_oracle_ties.py.  The residual HISTORY of the oracle is therefore pinned to the report by eight
    integers, two of them with a margin of a few
    per cent (cycle 3: 8.72e-7 < 9.22e-7 <= 1.011e-6, cycle 4: 5.81e-7 < 5.95e-7 <= 6.56e-7); what is left open
    is only which scaling the March-2020 binary used (:1289 still carries `// r_norm; scaled residual`)."""
    now, then, res = _both_counts(1, cycle)
    assert now == [12, 24, 40, 59][cycle]
    assert then == GOLD["table2_linear_mortar"]["rows"][cycle]["gmres"]


def test_published_count_of_table4_row1_follows_from_the_residual_history():
    now, then, _ = _both_counts(2, 0)
    assert (now, then) == (19, GOLD["table4_quadratic_mortar"]["rows"][0]["gmres"])


def test_published_count_of_table2_row5_follows_from_the_residual_history():
    now, then, res = _both_counts(1, 4)
    assert (now, then) == (82, 86) and GOLD["table2_linear_mortar"]["rows"][4]["gmres"] == 86
    assert res.residuals[86] < 1e-6 * res.r_norm <= res.residuals[85]


def test_table4_quadratic_mortar_cycle_0():
    """Quadratic mortar: reproduced only with the lower-cell / earlier-level tie rule (quirk Q5); rows 3 and 5 are
    reproduced with the restated deal.II tie rule -- tests/test_oracle_ties.py."""
    res = mo.run(_params(2, 1))
    _check_row(res[0], GOLD["table4_quadratic_mortar"]["rows"][0], its_slack=1)


def test_survey_probe_values_cycle0():
    """Full-precision values of an independent restatement, 1e-7 level."""
    res = mo.run(_params(1, 1))[0]
    assert res.r_norm == pytest.approx(4.1313268417737685, rel=1e-12)
    assert res.gmres_iterations == 12
    assert res.errors["velocity_l2l2"] == pytest.approx(0.6504239772938007, rel=1e-7)
    assert res.errors["pressure_l2l2"] == pytest.approx(0.7905736841187648, rel=1e-7)
    assert res.residuals[-3:] == pytest.approx([1.667135e-05, 1.844421e-06, 2.380891e-07], rel=1e-5)


def test_c2f_reproduces_constants_and_operator_is_linear():
    prm = _params(1, 1)
    mesh = mo.refined_meshes(prm, 4)[0]
    prob = mo.Problem(prm, 4, mesh[:-1], mesh[-1])
    prob.setup()
    sd = prob.subs[0]
    for side, c2f in sd.c2f.items():
        ms = prob.mortar[0] if side in (0, 2) else prob.mortar[1]
        big_h = (sd.Lx if side in (0, 2) else sd.Ly) / ms
        big_t = sd.dt * sd.nt / prob.mortar[2]
        coef = np.zeros(c2f.shape[1])
        coef[0::4] = 3.5 * big_h * big_t  # constant 3.5: only the (0,0) Legendre mode, value c/|F_H|
        np.testing.assert_allclose(c2f @ coef, 3.5, rtol=1e-13)
    rng = np.random.default_rng(0)
    a, b = rng.standard_normal(prob.n_iface), rng.standard_normal(prob.n_iface)
    lhs = prob.apply_S(2.0 * a - 3.0 * b)
    rhs = 2.0 * prob.apply_S(a) - 3.0 * prob.apply_S(b)
    np.testing.assert_allclose(lhs, rhs, rtol=1e-10, atol=1e-12)


def test_parameter_parser_and_topology():
    prm = mo.parse_parameter_file(os.path.join(ROOT, "parameter.txt"))
    assert (prm.c0, prm.mortar_degree, prm.num_refinement, prm.max_iteration) == (1.0, 1, 2, 500)
    assert prm.bc_type == ["D", "N", "D", "N"] and prm.bc_value == [0.0, 0.0, 0.0, 1.0]
    assert prm.mesh == [[3, 3, 3], [2, 2, 2], [4, 4, 4], [3, 3, 3], [1, 1, 1]]
    assert mo.find_divisors(4) == (2, 2) and mo.find_divisors(8) == (4, 2) and mo.find_divisors(64) == (8, 8)
    assert mo.find_neighbors(0, 2, 2) == [-1, 1, 2, -1]
    assert mo.find_neighbors(3, 2, 2) == [1, -1, -1, 2]
    assert mo.find_neighbors(5, 4, 4) == [1, 6, 9, 4]
