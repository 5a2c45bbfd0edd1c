"""The committed full-precision oracle outputs (tests/golden/oracle_vectors.json, written by
tests/golden/make_oracle_vectors.py) pin the oracle against drift: any edit of oracle/ that moves a GMRES count, a
residual history, an error column or lambda of the shipped parameter.txt shows up here."""
import json
import os

import numpy as np
import pytest

import mmmfe_oracle as mo
from helpers import ROOT, default_params

VEC = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_vectors.json")))


@pytest.mark.parametrize("name,degree,rule", [("linear", 1, None), ("quadratic_lower", 2, None),
                                              ("quadratic_dealii91", 2, "dealii91")])
def test_oracle_reproduces_its_committed_vectors(name, degree, rule):
    res = mo.run(default_params(degree, 3), tie_rules=rule)
    assert len(res) == len(VEC[name]) == 3
    for r, g in zip(res, VEC[name]):
        assert r.gmres_iterations == g["gmres_iterations"]
        assert r.r_norm == pytest.approx(g["r_norm"], rel=1e-13)
        for c, v in g["errors"].items():
            assert r.errors[c] == pytest.approx(v, rel=1e-10), c
        np.testing.assert_allclose(r.residuals, g["residuals"], rtol=1e-7, atol=1e-14)
        lam = np.asarray(g["lambda"])
        assert np.linalg.norm(r.lam - lam) <= 1e-10 * np.linalg.norm(lam)


def test_the_two_tie_rules_differ_from_cycle_2_on():
    a, b = VEC["quadratic_lower"], VEC["quadratic_dealii91"]
    assert a[0] == b[0] and a[1]["gmres_iterations"] == b[1]["gmres_iterations"]
    assert (a[2]["gmres_iterations"], b[2]["gmres_iterations"]) == (39, 38)
    assert float(f"{b[2]['errors']['lambda_l2']:.3e}") == 0.282  # report.pdf Table 4 row 3
