"""A-priori rates of the refinement sweep (BASELINE.json configs[4] asks for them; its RT1/DGQ1 spaces are refused --
the reference's own space-time transfer is broken for space_degree >= 1, quirk Q7 -- so this is the RT0 x DGQ0 sweep)
on the smooth data of inc/data_sinx_siny.h: p = e^t sin x sin y, K = I.  Backward Euler + lowest-order mixed elements
with a linear mortar: every error column converges with order 1."""
import math

import numpy as np
import pytest

import mmmfe_oracle as mo
from helpers import default_params


class DataSinxSiny:
    """inc/data_sinx_siny.h: KInverse :34-62, RightHandSidePressure :92-105, PressureBoundaryValues :121-137,
    ExactSolution :154-180, InitialCondition :275-294."""

    def k_inverse(self, x, y):
        one, zero = np.ones_like(x), np.zeros_like(x)
        return one, zero, zero, one

    def rhs(self, x, y, t, c0=1.0):
        return c0 * math.exp(t) * np.sin(x) * np.sin(y) + math.exp(t) * 2.0 * np.sin(x) * np.sin(y)

    def pressure_bc(self, x, y, t):
        return np.exp(t) * np.sin(x) * np.sin(y)

    def exact(self, x, y, t):
        return -np.exp(t) * np.cos(x) * np.sin(y), -np.exp(t) * np.sin(x) * np.cos(y), np.exp(t) * np.sin(x) * np.sin(y)

    def initial_pressure(self, x, y):
        return np.sin(x) * np.sin(y)


def test_first_order_rates_on_the_smooth_data():
    res = mo.run(default_params(1, 4), data=DataSinxSiny())
    assert [r.gmres_iterations > 0 for r in res] == [True] * 4
    for col in ("velocity_l2l2", "pressure_l2l2", "pressure_dg", "lambda_l2"):
        errs = [r.errors[col] for r in res]
        rates = [math.log2(a / b) for a, b in zip(errs, errs[1:])]
        assert all(e > 0 for e in errs) and errs[-1] < errs[0] / 5, (col, errs)
        assert all(0.95 < r < 1.05 for r in rates), (col, rates)  # measured: 0.99 ... 1.01 in every column and cycle
