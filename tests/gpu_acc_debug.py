import sys
import numpy as np, scipy.sparse.linalg as spla, scipy.linalg as sl
from helpers import default_params, engine_from_oracle, make_problem, pkg, rel_l2
from test_nd_tree import _schur

n = int(sys.argv[1])
prm = default_params()
prob = make_problem(prm, n_sub=2, mesh=[[n, n, 2], [n, n, 2]], mortar=[1, 1, 1])
sd = prob.subs[0]
S, Kup, Kpu, M = _schur(sd)
eng = engine_from_oracle(prob)
eng.setup()
rng = np.random.default_rng(1)
# (a) flux-only rhs
rhs = np.zeros(sd.n); rhs[:sd.nf] = rng.standard_normal(sd.nf)
x = eng.solve_saddle(0, rhs)
luS = spla.splu(S.tocsc())
print("flux-only rhs: u err vs splu(S)", rel_l2(x[:sd.nf], luS.solve(rhs[:sd.nf])), " vs splu(K)", rel_l2(x, sd.lu.solve(rhs)))
# (b) saddle rhs
rhs = rng.standard_normal(sd.n)
x = eng.solve_saddle(0, rhs); ref = sd.lu.solve(rhs)
print("saddle rhs: u err", rel_l2(x[:sd.nf], ref[:sd.nf]), " p err", rel_l2(x[sd.nf:], ref[sd.nf:]))
# (c) per-front factor accuracy
tree = pkg().nd_tree(n, n, 4)
R = eng.debug_get_factor(0, tree["r_total"])
nodes, offs, idx, src = tree["nodes"], tree["offsets"], tree["idx"], tree["src"]
Sc = S.tocsr(); nn = len(nodes)
node_of = np.full(Sc.shape[0], -1)
for i in range(nn): node_of[idx[offs[i,0]:offs[i,0]+nodes[i,0]]] = i
order = np.argsort(nodes[:,4], kind="stable")
U = {}; worst = {}
for i in order:
    s,b = nodes[i,0], nodes[i,1]; f=s+b
    v = idx[offs[i,0]:offs[i,0]+f]; pos = {int(g):p for p,g in enumerate(v)}
    F = np.zeros((f,f), dtype=np.longdouble)
    for p in range(s):
        row = Sc.getrow(v[p])
        for c,val in zip(row.indices,row.data):
            if int(c) in pos:
                F[p,pos[int(c)]] = val
                if node_of[c] != i: F[pos[int(c)],p] = val
    for c in range(2):
        ch = nodes[i,2+c]
        if ch<0: continue
        m = src[offs[i,1]+c*f: offs[i,1]+(c+1)*f]; sel = np.nonzero(m>=0)[0]
        F[np.ix_(sel,sel)] += U[ch][np.ix_(m[sel],m[sel])]
    F64 = np.asarray(F, dtype=np.float64)
    # high accuracy inverse via float64 inverse + two Newton steps in long double
    X = np.linalg.inv(F64[:s,:s]).astype(np.longdouble)
    F11 = F[:s,:s]
    for _ in range(3): X = X + X @ (np.eye(s, dtype=np.longdouble) - F11 @ X)
    G = F[s:,:s] @ X
    U[i] = F[s:,s:] - G @ F[:s,s:]
    ldr = nodes[i,6]; Rg = R[offs[i,2]:offs[i,2]+s*ldr].reshape(s, ldr)
    ex = np.abs(Rg[:, :s] - np.asarray(X, dtype=np.float64)).max() / np.abs(X).max()
    eg = (np.abs(Rg[:, s:f] + np.asarray(G, dtype=np.float64).T).max() / max(np.abs(G).max(), 1e-300)) if b else 0.0
    h = int(nodes[i,4])
    w = worst.setdefault(h, [0.0, 0.0, 0, 0.0])
    w[0] = max(w[0], float(ex)); w[1] = max(w[1], float(eg)); w[2] = max(w[2], int(s)); w[3] = max(w[3], float(np.linalg.cond(F64[:s,:s])))
for h in sorted(worst): print("height", h, "max rel err X %.2e  G %.2e  max s %d  cond(F11) %.2e" % tuple(worst[h]))
