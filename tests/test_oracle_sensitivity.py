"""How far round-off in the interface operator moves the result of the reference's GMRES (classical Gram-Schmidt,
no restart, no re-orthogonalisation: src/darcy_vtdd.cc:1413-1463).

The CUDA path and the oracle solve the subdomain problems with different direct solvers (block-inverse multifrontal
on the flux Schur complement vs SuperLU on the saddle matrix; the reference uses UMFPACK), so their operators agree
to ~1e-12, not to the last bit.  These tests measure, on the oracle alone, the factor by which a relative
perturbation eps of every operator application shows up in lambda, and that the iteration count does not move.
The factor is problem dependent (75-150 on the cycles of the shipped parameter.txt, ~4e7 on the 2x2 non-matching
96^2/48^2 case), so the long GPU runs measure it on their own problem (helpers.lambda_tolerance).
"""
import numpy as np
import pytest

import mmmfe_oracle as mo
from helpers import LAMBDA_AMPLIFICATION, default_params, make_problem, perturbed_gmres


def _spread(cycle, eps_list, seeds=(1, 2)):
    prm = default_params(1, cycle + 1)
    prob = make_problem(prm, cycle)
    prob.bar_sweep()
    base, its0, _ = perturbed_gmres(prob, 0.0, 0)
    out = []
    for eps in eps_list:
        for seed in seeds:
            lam, its, res = perturbed_gmres(prob, eps, seed)
            out.append((eps, its, float(np.linalg.norm(lam - base) / np.linalg.norm(base))))
    return its0, out


def test_lambda_amplification_at_59_iterations():
    """Default parameter.txt, cycle 3 (59 iterations): measured amplification 75-115 for eps = 1e-16 ... 1e-9."""
    its0, rows = _spread(3, [1.1e-16, 1e-13, 1e-10])
    assert its0 == 59
    for eps, its, rel in rows:
        assert its == its0
        assert 10 * eps < rel < LAMBDA_AMPLIFICATION * eps, (eps, rel)


@pytest.mark.slow
def test_iteration_count_of_table2_row5_is_not_a_roundoff_effect():
    """Cycle 4: 82 iterations here against 86 in report.pdf Table 2.  Perturbing every operator application by up
    to 1e-8 relative leaves the count at 82 (1e-7: 83) and moves lambda by ~150 eps, so the published 86 is not
    explained by solver round-off (UMFPACK vs SuperLU).  It is explained by the scaling of the stop test the
    March-2020 report was produced with (residual_k / r_norm < tol: tests/test_oracle_golden.py and tests/test_oracle_ties.py reproduce all eight
    published counts from the oracle's residual history).  The error columns agree in every printed digit."""
    its0, rows = _spread(4, [1e-12, 1e-8], seeds=(1,))
    assert its0 == 82
    for eps, its, rel in rows:
        assert its == 82
        assert rel < LAMBDA_AMPLIFICATION * eps


@pytest.mark.slow
def test_nonmatching_medium_case_is_far_more_sensitive():
    """2x2, 96^2 x 24 / 48^2 x 12, mortar 24 x 24 x 6 (configs[1] scaled down), 95 iterations: a 1-ulp perturbation of
    the operator moves lambda by ~5e-9 and the late residual history by ~0.3 %.  This is the reference algorithm
    (classical Gram-Schmidt without re-orthogonalisation), not a property of either implementation; it is why the GPU
    tests compare lambda against a per-problem measured tolerance and check the true residual instead."""
    prm = default_params(1, 1)
    prob = make_problem(prm, n_sub=4, mesh=[[96, 96, 24], [48, 48, 12], [48, 48, 12], [96, 96, 24]], mortar=[24, 24, 6])
    prob.bar_sweep()
    base, its0, res0 = perturbed_gmres(prob, 0.0, 0)
    lam, its, res = perturbed_gmres(prob, 1.1e-16, 1)
    assert its == its0 == 95
    rel = np.linalg.norm(lam - base) / np.linalg.norm(base)
    assert 1e-10 < rel < 1e-6
    assert np.max(np.abs(np.array(res) / np.array(res0) - 1)) > 1e-5
