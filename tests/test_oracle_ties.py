"""The point-location hazard of the quadratic mortar (quirk Q5): which fine cell a mortar-face Gauss point samples
when it lies exactly on a fine cell boundary (FEFieldFunction, inc/utilities.h:484).

deal.II tries the candidate cells in the order of GridTools' compare_point_association (scalar product of
cell-centre direction and point direction, equal products keep ascending cell index).  With dyadic coordinates the
two products are bitwise equal and the lower cell / earlier level wins; with non-dyadic coordinates the last bit of
the mapped quadrature point decides, point by point.  Consequences pinned here:
  * report.pdf Table 4 row 1 (ties only on the dyadic 2- and 4-cell subdomains) is reproduced digit for digit by the
    lower/earlier rule and by no other uniform rule;
  * row 3 (ties with ratio 6 on the non-dyadic 12-cell subdomains 0 and 3) lies strictly between the uniform choices
    on those two subdomains -- it cannot be reproduced without deal.II's round-off, and is bracketed instead;
  * on BASELINE.json configs[3] (8x8 subdomains, 256^2 x 512 / 128^2 x 256, mortar 16 x 16 x 32: every coordinate
    dyadic) the rule is deterministic, and the scaled-down case shows how much is at stake if it were not.
"""
import itertools
import json
import os

import numpy as np
import pytest

import mmmfe_oracle as mo
from helpers import ROOT, default_params

GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "report_tables.json")))
COLS = ["velocity_l2l2", "pressure_dg", "pressure_l2l2", "lambda_l2"]
RULES = [dict(space=s, time=t) for s, t in itertools.product(("lower", "upper"), repeat=2)]


def _cycle(k, idx, tie_rules):
    prm = default_params(k, idx + 1)
    mesh = mo.refined_meshes(prm, 4)[idx]
    return mo.Problem(prm, 4, mesh[:-1], mesh[-1], tie_rules=tie_rules).run_cycle(idx)


def test_row1_only_the_lower_rule_on_the_dyadic_subdomains():
    row = GOLD["table4_quadratic_mortar"]["rows"][0]
    hits = []
    for rule in RULES:
        res = _cycle(2, 0, {1: rule, 2: rule})  # subdomains 1 (2 cells) and 2 (4 cells) are the only ones with ties
        ok = all(float(f"{res.errors[c]:.3e}") == pytest.approx(row[c], rel=1.01e-3) for c in COLS)
        hits.append(ok)
    assert hits == [True, False, False, False]
    # the 3-cell subdomains have no ties at all: their rule changes nothing
    a = _cycle(2, 0, {0: RULES[3], 3: RULES[3]})
    b = _cycle(2, 0, None)
    assert a.errors == b.errors and a.gmres_iterations == b.gmres_iterations


def test_row3_is_bracketed_by_the_choices_on_the_non_dyadic_subdomains():
    row = GOLD["table4_quadratic_mortar"]["rows"][1]
    vals = {c: [] for c in COLS}
    for rule in RULES:
        res = _cycle(2, 2, {0: rule, 3: rule})  # 12-cell subdomains: h = 1/24, ties at ratio 6
        for c in COLS:
            vals[c].append(res.errors[c])
    for c in COLS:
        assert min(vals[c]) < row[c] < max(vals[c]), (c, vals[c], row[c])
        assert vals[c][0] == min(vals[c])  # lower/earlier is one end of the bracket
    # within 3 % (velocity, pressure L2) of the published row with the default rule
    assert vals["velocity_l2l2"][0] == pytest.approx(row["velocity_l2l2"], rel=0.03)
    assert vals["pressure_l2l2"][0] == pytest.approx(row["pressure_l2l2"], rel=0.05)


def test_configs3_coordinates_are_exact_so_ties_are_deterministic():
    """Every fine vertex coordinate and every mortar Gauss-point tie coordinate of configs[3] is a dyadic rational
    with few bits: x0 + i*h and the mapped face centre are computed without rounding, the candidates' scalar
    products are bitwise equal, deal.II keeps ascending cell order = lower cell / earlier level."""
    from fractions import Fraction
    n0 = 8
    for nx, nt, ms, mt in ((256, 512, 16, 32), (128, 256, 16, 32)):
        h, dt = Fraction(1, n0 * nx), Fraction(1, 2 * nt)
        big_h, big_t = Fraction(1, n0 * ms), Fraction(1, 2 * mt)
        for frac in (h, dt, big_h, big_t):
            assert frac.denominator & (frac.denominator - 1) == 0  # power of two
        # the centre Gauss point of mortar face f sits at (f + 1/2) H: a fine vertex, exactly representable
        for f in (0, 5, ms - 1):
            tie = (f + Fraction(1, 2)) * big_h
            assert (tie / h).denominator == 1
            assert float(tie) == float(f + 0.5) * float(big_h) == float(int(tie / h)) * float(h)


def test_what_the_tie_rule_is_worth_on_a_scaled_configs3():
    """8x8 checkerboard 8^2 x 16 / 4^2 x 8, mortar 2 x 2 x 4, degree 2 (every centre Gauss point is a tie): the
    operator moves by tens of per cent between the uniform rules -- the rule is not a detail, which is why it is
    fixed to what deal.II does on dyadic grids and stated in DESIGN.md."""
    prm = default_params(2)
    mesh = [[8, 8, 16] if ((r % 8) + (r // 8)) % 2 == 0 else [4, 4, 8] for r in range(64)]
    q = np.random.default_rng(0).standard_normal(2 * 112 * 2 * 4 * 9)
    outs = []
    for rule in (RULES[0], RULES[3]):
        prob = mo.Problem(prm, 64, mesh, [2, 2, 4], tie_rules={r: rule for r in range(64)})
        prob.setup()
        outs.append(prob.apply_S(q))
    rel = np.linalg.norm(outs[0] - outs[1]) / np.linalg.norm(outs[0])
    assert 0.05 < rel < 2.0
