"""The point-location hazard of the quadratic mortar (quirk Q5): which fine cell a mortar-face Gauss point samples
when it lies exactly on a fine cell boundary (FEFieldFunction, inc/utilities.h:484).

deal.II (9.1 series, restated in oracle/dealii_point_location.py) tries the cells around the closest vertex in the
order of the scalar product of cell-centre direction and point direction; equal products keep ascending cell index.
In exact arithmetic the two candidates of a tie have equal products and the lower cell / earlier level wins -- the rule
the oracle and the product implement.  In floating point the mapped Gauss point is a sum of eight rounded products
(its OTHER two abscissae are irrational) and the centre directions are normalised sums, so the last bits decide,
point by point, unless every coordinate has a single significant bit.  Consequences pinned here:
  * report.pdf Table 4 row 1 (ties only on the 2- and 4-cell subdomains, coordinates 0, 1/2, 1: everything exact) is
    reproduced digit for digit by the lower/earlier rule and by no other uniform rule -- and the restated candidate
    order predicts exactly that rule there;
  * rows 3 and 5 (ties with ratio 6 and 12 on the 12- and 48-cell subdomains 0 and 3, and a few on the dyadic
    subdomains 1 and 2) lie strictly between the uniform choices; with the restated candidate order on ALL FOUR
    subdomains -- no fitted parameter -- the oracle reproduces both rows in every printed digit of the four error
    columns and their report-era GMRES counts 34 and 57;
  * on BASELINE.json configs[3] (8x8 subdomains, 256^2 x 512 / 128^2 x 256, mortar 16 x 16 x 32) the tie COORDINATES
    are exact, but the mapped Gauss points are not: under the restated order about a tenth of the tie points leave
    the lower cell on the scaled-down case and the operator moves by ~17 %.  What the reference returns there depends
    on the deal.II version and its round-off (README.md:42 names 9.5.2, whose point location is a different
    algorithm again); the product implements the exact-arithmetic limit and says so.
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


def test_configs3_tie_coordinates_are_exact():
    """Every fine vertex coordinate and every mortar-face centre of configs[3] is a dyadic rational with few bits:
    the TIE ITSELF is exact (x0 + i*h and the face centre are computed without rounding), so the exact-arithmetic
    rule -- lower cell / earlier level -- is well defined there.  (What deal.II's floating-point point location
    makes of it is a different matter: test_restated_candidate_order_on_a_scaled_configs3.)"""
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


# ---- the restated deal.II 9.1 candidate order (oracle/dealii_point_location.py) -----------------------------------
def _restated(idx, tol=1e-7):
    prm = default_params(2, idx + 1)
    prm.tolerance = tol
    mesh = mo.refined_meshes(prm, 4)[idx]
    prob = mo.Problem(prm, 4, mesh[:-1], mesh[-1], tie_rules="dealii91")
    res = prob.run_cycle(idx)
    report_era = next(k for k in range(1, len(res.residuals)) if res.residuals[k] < 1e-6 * res.r_norm)
    current = next(k for k in range(1, len(res.residuals)) if res.residuals[k] < 1e-6)
    return res, report_era, current, prob


def _digits(res, row):
    return all(float(f"{res.errors[c]:.3e}") == pytest.approx(row[c], rel=1.01e-3) for c in COLS)


def test_restated_candidate_order_predicts_the_lower_rule_where_row1_pins_it():
    """Cycle 0: the only ties are on the 2- and 4-cell subdomains 1 and 2 (coordinates with one significant bit).
    The restated order keeps every one of them in the lower cell / earlier level -- the rule that alone reproduces
    Table 4 row 1 (test_row1_only_the_lower_rule_on_the_dyadic_subdomains) -- and the row follows digit for digit."""
    res, report_era, current, prob = _restated(0)
    picks = [v for r in (1, 2) for side in prob.subs[r].tie_points.values() for v in side.values()]
    assert len(picks) == 20 and not any(a or b for a, b in picks)
    assert not any(prob.subs[r].tie_points[side] for r in (0, 3) for side in prob.subs[r].tie_points)  # 3 cells: no ties
    row = GOLD["table4_quadratic_mortar"]["rows"][0]
    assert _digits(res, row) and (report_era, current) == (row["gmres"], 19)


def test_row3_is_reproduced_by_the_restated_candidate_order():
    """Table 4 row 3 (cycle 2), every printed digit and the published GMRES count.  The picks: 32 of 40 tie points of
    subdomain 0 and 20 of 40 of subdomain 3 leave the lower cell, and 2 of 40 on each of the DYADIC subdomains 1, 2."""
    row = GOLD["table4_quadratic_mortar"]["rows"][1]
    res, report_era, current, prob = _restated(2)
    assert _digits(res, row), {c: res.errors[c] for c in COLS}
    assert report_era == row["gmres"] == 34
    moved = [sum(1 for side in prob.subs[r].tie_points.values() for v in side.values() if any(v)) for r in range(4)]
    assert moved == [32, 2, 2, 20]
    default = _cycle(2, 2, None)
    assert not _digits(default, row) and all(abs(default.errors[c] / row[c] - 1) > 0.025 for c in COLS)


def test_row5_is_reproduced_by_the_restated_candidate_order():
    """Table 4 row 5 (cycle 4, 48- and 64-cell subdomains): every printed digit and the published count 57 -- the
    out-of-sample row (the restatement was written against row 3)."""
    row = GOLD["table4_quadratic_mortar"]["rows"][2]
    res, report_era, current, _ = _restated(4)
    assert _digits(res, row), {c: res.errors[c] for c in COLS}
    assert report_era == row["gmres"] == 57


def test_restated_candidate_order_on_a_scaled_configs3():
    """8x8 checkerboard 8^2 x 16 / 4^2 x 8, mortar 2 x 2 x 4, degree 2: dyadic coordinates, yet under the restated
    floating-point order a tenth of the tie points leave the lower cell (the mapped Gauss point is off the vertex by
    an ulp) and the operator moves by ~17 %.  The picks are not invariant under time translation either, so the
    time-Toeplitz structure of the operator exists only under the exact-arithmetic rule."""
    prm = default_params(2)
    mesh = [[8, 8, 16] if ((r % 8) + (r // 8)) % 2 == 0 else [4, 4, 8] for r in range(64)]
    q = np.random.default_rng(0).standard_normal(2 * 112 * 2 * 4 * 9)
    outs = []
    for tr in (None, "dealii91"):
        prob = mo.Problem(prm, 64, mesh, [2, 2, 4], tie_rules=tr)
        prob.setup()
        outs.append(prob.apply_S(q))
    picks = [v for sd in prob.subs for side in sd.tie_points.values() for v in side.values()]
    moved = sum(1 for a, b in picks if a or b)
    assert len(picks) == 8960 and 0.05 < moved / len(picks) < 0.2
    rel = np.linalg.norm(outs[0] - outs[1]) / np.linalg.norm(outs[0])
    assert 0.05 < rel < 0.5
    # time translation: the time picks of a side differ between mortar time slabs
    sd = prob.subs[0]
    side = next(iter(sd.tie_points))
    by_slab = {}
    for (fs, ft, qq, rr), pick in sd.tie_points[side].items():
        by_slab.setdefault(ft, {})[(fs, qq, rr)] = pick
    assert any(by_slab[ft] != by_slab[0] for ft in by_slab)
