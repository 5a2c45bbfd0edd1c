"""Which fine cell deal.II 9.1 hands a mortar Gauss point that lies exactly on a fine cell boundary.

TEST INFRASTRUCTURE ONLY (see mmmfe_oracle.py).  The reference samples the fine trace at the mortar Gauss points
through ``Functions::FEFieldFunction`` (inc/utilities.h:484), i.e. through deal.II's point location.  deal.II is a
third-party dependency that is absent from /root/reference and unpinned (CMakeLists.txt:23 asks for >= 8.4,
README.md:42 names 9.5.2, report.pdf of March 2020 was produced with 9.1 or older); what follows restates the published
algorithm of the 9.0/9.1 series, in IEEE double arithmetic without fused multiply-adds:

  * ``GridGenerator::subdivided_hyper_rectangle``: vertex i of a direction = fl(p1 + fl(i * fl((p2 - p1) / reps)));
  * ``QGauss``/``QProjector::project_to_face`` + ``MappingQGeneric`` (degree 1, hard-coded trilinear shape functions
    ``(1-x)(1-y)(1-z) ...`` multiplied left to right): quadrature point = sum_k shape_k * vertex_k, accumulated in
    vertex order;
  * ``GridTools::find_active_cell_around_point``: closest vertex V, the cells around V in ascending index order,
    sorted (stable for so few cells) by descending scalar product of the normalised direction V -> cell centre
    (``cell->center()`` = sum of the 8 vertices / 8) with the normalised direction V -> point; the first cell of that
    order that contains the point is taken.

In exact arithmetic the two candidates of a tie have equal products and the lower cell (earlier level) wins: quirk
Q5, the default of the oracle and of the product.  In floating point the mapped Gauss point (its other two abscissae
are irrational) and the normalised centre directions carry round-off, and the last bits decide point by point --
also on the dyadic subdomains once a coordinate has more than one significant bit.

Pinned by every golden number the reference holds for the quadratic mortar (tests/test_oracle_ties.py): with this
rule on all four subdomains the oracle reproduces report.pdf Table 4 rows 1, 3 AND 5 in every printed digit of the
four error columns, and their GMRES counts 18 / 34 / 57 under the report-era stop test (tests/test_oracle_golden.py).
Not restated: the strict ``is_inside_unit_cell`` test on deal.II's inverse map (whether it is emulated with per
direction quotients or left out changes none of the picks on these meshes); fused multiply-adds (none assumed).

``tie_points(sub, n_fine, n_mortar, ...)`` returns {side: {(fs, ft, q, r): (space upper?, time upper?)}} in the form
``mmmfe_oracle.Problem(..., tie_rules={sub: {"points": ...}})`` takes; ``Problem(..., tie_rules="dealii91")`` calls it
for every subdomain.  3-point Gauss rule (mortar degree 2) only.
"""
from decimal import Decimal, getcontext

import numpy as np

f64 = np.float64
# deal.II face -> side index of the reference (0 bottom, 1 right, 2 top, 3 left): x=0 left, x=1 right, y=0 bottom, y=1 top
_FACE_TO_SIDE = {0: 3, 1: 1, 2: 0, 3: 2}


def _gauss3():
    getcontext().prec = 50
    d = Decimal(15).sqrt() / Decimal(10)
    return [float(Decimal("0.5") - d), 0.5, float(Decimal("0.5") + d)]


def _grid(p1, p2, reps):
    delta = (f64(p2) - f64(p1)) / f64(reps)
    return [f64(p1) + f64(i) * delta for i in range(reps + 1)]


def _shapes(x, y, z):
    x, y, z, one = f64(x), f64(y), f64(z), f64(1.0)
    return [(one - x) * (one - y) * (one - z), x * (one - y) * (one - z), (one - x) * y * (one - z), x * y * (one - z),
            (one - x) * (one - y) * z, x * (one - y) * z, (one - x) * y * z, x * y * z]


def _mapped(shape, verts):
    res = [shape[0] * verts[0][i] for i in range(3)]
    for k in range(1, 8):
        for i in range(3):
            res[i] = res[i] + shape[k] * verts[k][i]
    return res


def _normalised(t):
    s = t[0] * t[0]
    s = s + t[1] * t[1]
    s = s + t[2] * t[2]
    n = np.sqrt(s)
    return [x / n for x in t] if n > 0 else list(t)


def _dot(a, b):
    s = a[0] * b[0]
    s = s + a[1] * b[1]
    return s + a[2] * b[2]


def tie_points(sub, n_fine, n_mortar, final_time=0.5, n0=2, n1=2):
    """Subdomain `sub` of the n0 x n1 decomposition of the unit square (get_subdomain_coordinates,
    inc/utilities.h:152-157): n_fine = (nx, ny, nt) fine space-time cells (or one number for all three), n_mortar
    likewise, 3-point Gauss rule (quadratic mortar)."""
    g = _gauss3()
    n_fine = (n_fine,) * 3 if np.isscalar(n_fine) else tuple(n_fine)
    n_mortar = (n_mortar,) * 3 if np.isscalar(n_mortar) else tuple(n_mortar)
    dims = [f64(1.0) / f64(n0), f64(1.0) / f64(n1)]
    p1 = [f64(sub % n0) * dims[0], dims[1] * f64(sub // n0), 0.0]
    p2 = [p1[0] + dims[0], p1[1] + dims[1], final_time]
    fine = [_grid(p1[i], p2[i], n_fine[i]) for i in range(3)]
    mort = [_grid(p1[i], p2[i], n_mortar[i]) for i in range(3)]

    def vert(i, j, k):
        return (fine[0][i], fine[1][j], fine[2][k])

    def centre(c):
        p = [f64(0.0)] * 3
        for v in range(8):
            vv = vert(c[0] + (v & 1), c[1] + ((v >> 1) & 1), c[2] + ((v >> 2) & 1))
            p = [p[a] + vv[a] for a in range(3)]
        return [x / f64(8.0) for x in p]

    out = {}
    ix, iy = sub % n0, sub // n0
    faces = [fc for fc, inner in ((0, ix > 0), (1, ix < n0 - 1), (2, iy > 0), (3, iy < n1 - 1)) if inner]
    for face in faces:  # the interface sides of the subdomain
        side = _FACE_TO_SIDE[face]
        normal = 0 if face in (0, 1) else 1
        tangent = 1 - normal
        table = out.setdefault(side, {})
        for cs in range(n_mortar[tangent]):
            for ct in range(n_mortar[2]):
                for i0 in range(3):
                    for i1 in range(3):
                        if face in (0, 1):  # QProjector: (face, q0, q1) on x-faces, (q1, face, q0) on y-faces
                            unit, q, r = (float(face), g[i0], g[i1]), i0, i1
                            cell = ((0 if face == 0 else n_mortar[0] - 1), cs, ct)
                        else:
                            unit, q, r = (g[i1], float(face - 2), g[i0]), i1, i0
                            cell = (cs, (0 if face == 2 else n_mortar[1] - 1), ct)
                        tie_s = q == 1 and ((2 * cs + 1) * n_fine[tangent]) % (2 * n_mortar[tangent]) == 0
                        tie_t = r == 1 and ((2 * ct + 1) * n_fine[2]) % (2 * n_mortar[2]) == 0
                        if not (tie_s or tie_t):
                            continue  # only centre abscissae can hit a fine cell boundary
                        mv = [[mort[0][cell[0] + (k & 1)], mort[1][cell[1] + ((k >> 1) & 1)],
                               mort[2][cell[2] + ((k >> 2) & 1)]] for k in range(8)]
                        p = _mapped(_shapes(*unit), mv)
                        idx = [int(np.argmin([abs(p[a] - x) for x in fine[a]])) for a in range(3)]
                        v_close = vert(*idx)
                        to_point = _normalised([p[a] - v_close[a] for a in range(3)])
                        cells = [(idx[0] + di, idx[1] + dj, idx[2] + dk) for dk in (-1, 0) for dj in (-1, 0)
                                 for di in (-1, 0)]
                        cells = [c for c in cells if all(0 <= c[a] < n_fine[a] for a in range(3))]
                        cells.sort(key=lambda c: (c[2], c[1], c[0]))  # std::set of active cell iterators
                        prod = [_dot(_normalised([x - y for x, y in zip(centre(c), v_close)]), to_point) for c in cells]
                        order = sorted(range(len(cells)), key=lambda a: -prod[a])  # stable, like insertion sort
                        chosen = next(cells[a] for a in order
                                      if all(a_ == normal or fine[a_][cells[a][a_]] <= p[a_] <= fine[a_][cells[a][a_] + 1]
                                             for a_ in range(3)))
                        s_vertex = (2 * cs + 1) * n_fine[tangent] // (2 * n_mortar[tangent])
                        t_vertex = (2 * ct + 1) * n_fine[2] // (2 * n_mortar[2])
                        table[(cs, ct, q, r)] = (tie_s and chosen[tangent] == s_vertex, tie_t and chosen[2] == t_vertex)
    return out
