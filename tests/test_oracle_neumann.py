"""CPU checks of the oracle's essential (Neumann) outer sides (src/darcy_vtdd.cc:205-275, 479, 863-866): the
constrained system of distribute_local_to_global + constraint_bc.distribute is the elimination of the prescribed
fluxes, and the prescribed normal fluxes come back in the solution.  (The reference has no published numbers for
is_manufact_soln = 0: parity unpinned, consistency only.)"""
import numpy as np
import scipy.sparse.linalg as spla

from helpers import mo

MESH = [[3, 3, 3], [2, 2, 2], [4, 4, 4], [3, 3, 3]]


def _problem(bc_type, bc_value):
    prm = mo.Params()
    prm.is_manufact, prm.bc_type, prm.bc_value = False, bc_type, bc_value
    prob = mo.Problem(prm, 4, MESH, [1, 1, 1], keep_solution=True)
    return prm, prob


def test_constrained_system_is_the_elimination_of_the_prescribed_fluxes():
    prm, prob = _problem(["N", "D", "D", "N"], [0.5, 1.0, 2.0, -0.25])
    prob.setup()
    prob.bar_sweep()
    n_con = 0
    for sd in prob.subs:
        if not len(sd.con_dofs):
            continue
        n_con += len(sd.con_dofs)
        full = sd.assemble_matrix().tocsr()
        free = np.setdiff1d(np.arange(sd.n), sd.con_dofs)
        g = np.zeros(sd.n)
        g[sd.con_dofs] = sd.con_vals
        plain = mo.Subdomain(sd.r, 2, 2, sd.nx, sd.ny, sd.nt, prm, prob.data)  # no constraint attributes: raw rhs
        rhs = plain.rhs_bar(sd.dt, sd.project_initial_condition(), prob.qdegree, prm)
        x = g.copy()
        x[free] = spla.spsolve(full[free][:, free].tocsc(), (rhs - full @ g)[free])
        np.testing.assert_allclose(sd.bar_solution[0], x, rtol=0, atol=1e-13 * np.abs(x).max())
        # flux dof = outward-normal sign * g * |edge| in the positive coordinate direction (:228-231)
        np.testing.assert_array_equal(sd.bar_solution[-1][sd.con_dofs], sd.con_vals)
        # constrained rows and columns left the matrix, the diagonal kept |a_ii| (:479)
        m = sd.matrix.tocsr()
        for d in sd.con_dofs:
            row = m.getrow(d)
            assert row.nnz == 1 and row.indices[0] == d and row.data[0] == abs(full[d, d])
            assert m.getcol(d).nnz == 1
    assert n_con == 3 + 4 + 4 + 3  # left side of subdomains 0, 2; top side of subdomains 2, 3


def test_side_values_follow_the_reference_indexing():
    """Neumann data read bc_const_functs[left, bottom, right, top] (:228-231); Dirichlet data on the TOP side read
    index 1 (:673, a quirk of the reference that the oracle keeps)."""
    prm, prob = _problem(["D", "D", "N", "N"], [1.0, 2.0, 3.0, 4.0])
    prob.setup()
    top_right = prob.subs[3]
    right = np.asarray(top_right.side_edges[mo.RIGHT])
    top = np.asarray(top_right.side_edges[mo.TOP])
    got = dict(zip(top_right.con_dofs.tolist(), top_right.con_vals.tolist()))
    assert all(got[d] == 3.0 * top_right.hy for d in right)
    assert all(got[d] == 4.0 * top_right.hx for d in top)
    prm2, prob2 = _problem(["D", "D", "D", "D"], [1.0, 2.0, 3.0, 4.0])
    assert mo._dirichlet_constant(prm2, mo.TOP) == 2.0 and mo._dirichlet_constant(prm2, mo.RIGHT) == 3.0


def test_manufactured_runs_ignore_the_bc_lines():
    prm = mo.Params()
    prm.bc_type = ["N", "N", "N", "N"]  # is_manufact stays True: all outer sides Dirichlet (:210, 271-273)
    prob = mo.Problem(prm, 4, MESH, [1, 1, 1])
    prob.setup()
    assert all(len(sd.con_dofs) == 0 for sd in prob.subs)
