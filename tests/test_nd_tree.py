"""CPU check of the symbolic nested dissection: a NumPy emulation of the block-inverse multifrontal
factorisation and solve, driven by the engine's own tree, must reproduce a sparse direct solve."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from helpers import default_params, make_problem, pkg, rel_l2


def _schur(sd):
    """S = A + dt/c0 B^T M^-1 B on the flux dofs (canonical numbering) from the oracle's saddle matrix."""
    K = sd.matrix.tocsr()
    A = K[:sd.nf, :sd.nf]
    Kup = K[:sd.nf, sd.nf:]
    Kpu = K[sd.nf:, :sd.nf]
    M = K[sd.nf:, sd.nf:].diagonal()
    return (A - Kup @ sp.diags(1.0 / M) @ Kpu).tocsr(), Kup, Kpu, M


def _emulate(tree, S, rhs):
    nodes, offs, idx, src = tree["nodes"], tree["offsets"], tree["idx"], tree["src"]
    nn = len(nodes)
    S = S.tocsr()
    node_of = np.full(S.shape[0], -1)
    for i in range(nn):
        s = nodes[i, 0]
        node_of[idx[offs[i, 0]:offs[i, 0] + s]] = i
    assert (node_of >= 0).all(), "every flux dof is eliminated by exactly one node"
    order = np.argsort(nodes[:, 4], kind="stable")
    U, R, Z = {}, {}, {}
    x = rhs.copy()
    for i in order:  # heights ascending
        s, b = nodes[i, 0], nodes[i, 1]
        f = s + b
        v = idx[offs[i, 0]:offs[i, 0] + f]
        pos = {int(g): p for p, g in enumerate(v)}
        F = np.zeros((f, f))
        for p in range(s):
            row = S.getrow(v[p])
            for c, val in zip(row.indices, row.data):
                if node_of[c] == i:
                    F[p, pos[int(c)]] = val
                elif int(c) in pos:
                    F[p, pos[int(c)]] = val
                    F[pos[int(c)], p] = val
        vec = np.zeros(f)
        vec[:s] = x[v[:s]]
        for c in range(2):
            ch = nodes[i, 2 + c]
            if ch < 0:
                continue
            m = src[offs[i, 1] + c * f: offs[i, 1] + (c + 1) * f]
            sel = np.nonzero(m >= 0)[0]
            assert len(sel) == nodes[ch, 1], "every child boundary variable lands in the parent front"
            F[np.ix_(sel, sel)] += U[ch][np.ix_(m[sel], m[sel])]
            vec[sel] += Z[ch][m[sel]]
        F11i = np.linalg.inv(F[:s, :s])
        G = F[s:, :s] @ F11i
        U[i] = F[s:, s:] - G @ F[:s, s:]
        R[i] = (F11i, G)
        Z[i] = vec[s:] - G @ vec[:s]
        R[i] = (F11i, G, vec[:s].copy())
    for i in order[::-1]:
        s, b = nodes[i, 0], nodes[i, 1]
        v = idx[offs[i, 0]:offs[i, 0] + s + b]
        F11i, G, zs = R[i]
        x[v[:s]] = F11i @ zs - G.T @ x[v[s:]]
    return x


@pytest.mark.parametrize("shape,leaf", [((1, 1), 4), ((2, 2), 4), ((3, 3), 4), ((4, 4), 1), ((5, 7), 4), ((8, 8), 4),
                                        ((13, 6), 2), ((16, 16), 4), ((9, 20), 8)])
def test_tree_drives_a_correct_multifrontal_solve(shape, leaf):
    nx, ny = shape
    prm = default_params()
    prob = make_problem(prm, n_sub=2, mesh=[[nx, ny, 2], [nx, ny, 2]], mortar=[1, 1, 1])
    sd = prob.subs[0]
    S, Kup, Kpu, M = _schur(sd)
    tree = pkg().nd_tree(nx, ny, leaf)
    nodes = tree["nodes"]
    assert nodes[:, 0].sum() == sd.nf
    # post-order: children precede parents, heights consistent
    for i, nd in enumerate(nodes):
        for ch in nd[2:4]:
            if ch >= 0:
                assert ch < i and nodes[ch, 7] == i and nodes[ch, 4] < nd[4]
    rhs = np.random.default_rng(0).standard_normal(sd.nf)
    x = _emulate(tree, S, rhs)
    ref = spla.spsolve(S.tocsc(), rhs)
    assert rel_l2(x, ref) < 1e-11


def test_pressure_elimination_is_equivalent_to_the_saddle_system():
    prm = default_params()
    prob = make_problem(prm, n_sub=2, mesh=[[6, 5, 2], [6, 5, 2]], mortar=[1, 1, 1])
    sd = prob.subs[0]
    S, Kup, Kpu, M = _schur(sd)
    rhs = np.random.default_rng(1).standard_normal(sd.n)
    ptil = rhs[sd.nf:] / M
    u = spla.spsolve(S.tocsc(), rhs[:sd.nf] - Kup @ ptil)
    p = ptil - (Kpu @ u) / M
    ref = sd.lu.solve(rhs)
    assert rel_l2(np.concatenate([u, p]), ref) < 1e-11


def test_factor_size_scales_like_n_log_n():
    vals = {}
    for n in (32, 64, 128):
        t = pkg().nd_tree(n, n, 4)
        vals[n] = t["r_total"] / (2 * n * (n + 1))
    assert vals[32] < vals[64] < vals[128] < 80  # stored factor values per flux dof
