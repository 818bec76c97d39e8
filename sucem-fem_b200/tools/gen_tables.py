#!/usr/bin/env python
"""Generate csrc/generated/n1curl_tables.inc: pre-integrated reference tensors of the first-kind Nedelec
(N1curl) elements of order 1 and 2 on the UFC reference tetrahedron and triangle.

The reference gets these numbers from FFC-generated ``tabulate_tensor`` code for the forms in
/root/reference/sucemfem/Forms.py:76,84 and /root/reference/sucemfem/BoundaryConditions/ABC.py:57 on the element
named at /root/reference/sucemfem/ProblemConfigurations/EMProblem.py:36.  Here they are derived symbolically
(sympy, exact rationals) and rounded once to fp64.  This is a build-time tool of the product; it is written
independently of oracle/refelem.py (different algebra back-end), and tests compare the two.

Per (dim, order) the tables are, with ab running over the symmetric index pairs
(3-D: 00,11,22,01,02,12; 2-D: 00,11,01) and off-diagonal pairs holding T[a][b]+T[b][a]:

  mass  [i][j][ab] = int_K phi_i^a phi_j^b            (contracted with |detJ| eps_r (J^-1 J^-T)_ab)
  curl  [i][j][ab] = int_K (curl phi_i)^a (curl phi_j)^b (3-D; with (J^T J)_ab /(|detJ| mu_r)); 2-D: scalar [i][j]
  facet [lf][i][j][ab] = int over the unit facet simplex of phi_i^a phi_j^b on local facet lf
"""
import os
import sys
from math import factorial

import sympy as sp

TET_EDGES = ((2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1))
TET_FACES = ((1, 2, 3), (0, 2, 3), (0, 1, 3), (0, 1, 2))
TRI_EDGES = ((1, 2), (0, 2), (0, 1))


def simplex_integral(expr, syms):
    """Exact integral of a polynomial over the unit simplex in the given symbols."""
    expr = sp.expand(expr)
    if expr == 0:
        return sp.Integer(0)
    poly = sp.Poly(expr, *syms)
    tot = sp.Integer(0)
    d = len(syms)
    for mon, c in poly.terms():
        num = 1
        for k in mon:
            num *= factorial(k)
        tot += c * sp.Rational(num, factorial(sum(mon) + d))
    return tot


def build_basis(dim, order):
    X = sp.symbols('x y z')[:dim]
    verts = [sp.Matrix([0] * dim)] + [sp.Matrix([1 if i == d else 0 for i in range(dim)]) for d in range(dim)]
    Xv = sp.Matrix(X)
    # generators of R_order = P_{order-1}^dim (+) {p in hom. P_order^dim : p.X = 0}
    monos = [sp.Integer(1)] + (list(X) if order == 2 else [])
    gens = [sp.Matrix([m if dd == d else 0 for dd in range(dim)]) for m in monos for d in range(dim)]
    homs = [sp.Integer(1)] if order == 1 else list(X)
    if dim == 3:
        extra = [Xv.cross(sp.Matrix([h if dd == d else 0 for dd in range(3)])) for h in homs for d in range(3)]
        if order == 2:
            extra = extra[:-1]       # the nine fields satisfy one relation (X x X = 0)
    else:
        extra = [sp.Matrix([-X[1] * h, X[0] * h]) for h in homs]
    gens += extra
    edges = TET_EDGES if dim == 3 else TRI_EDGES
    fun = []                         # functionals as callables on a vector field
    def point_eval(pt, t):
        sub = dict(zip(X, pt))
        return lambda u: (u.subs(sub).T * t)[0, 0]
    for (a, b) in edges:
        t = verts[b] - verts[a]
        if order == 1:
            fun.append(point_eval(verts[a] + t / 2, t))
        else:
            for k in range(2):
                fun.append(point_eval(verts[a] + sp.Rational(k + 1, 3) * t, t))
    if order == 2 and dim == 3:
        for (a, b, c) in TET_FACES:
            cen = (verts[a] + verts[b] + verts[c]) / 3
            fun.append(point_eval(cen, verts[b] - verts[a]))
            fun.append(point_eval(cen, verts[c] - verts[a]))
    if order == 2 and dim == 2:
        # interior moments against e_x, e_y (scale not pinned by any reference fixture)
        for d in range(2):
            fun.append(lambda u, d=d: simplex_integral(u[d], X))
    n = len(fun)
    assert len(gens) == n
    V = sp.Matrix(n, n, lambda i, m: fun[i](gens[m]))
    Vi = V.inv()
    basis = []
    for j in range(n):
        phi = sp.zeros(dim, 1)
        for m in range(n):
            phi += Vi[m, j] * gens[m]
        basis.append(phi.applyfunc(sp.expand))
    return X, verts, basis


def curl(phi, X):
    if len(X) == 3:
        return sp.Matrix([sp.diff(phi[2], X[1]) - sp.diff(phi[1], X[2]),
                          sp.diff(phi[0], X[2]) - sp.diff(phi[2], X[0]),
                          sp.diff(phi[1], X[0]) - sp.diff(phi[0], X[1])])
    return sp.Matrix([sp.diff(phi[1], X[0]) - sp.diff(phi[0], X[1])])


def sym_pairs(d):
    return [(a, a) for a in range(d)] + [(a, b) for a in range(d) for b in range(a + 1, d)]


def fmt(x):
    return repr(float(x)) if x != 0 else '0.0'


def emit_array(name, dims, flat):
    assert len(flat) == eval('*'.join(map(str, dims)))
    s = 'SFEM_TABLE double %s%s = {\n' % (name, ''.join('[%d]' % d for d in dims))
    for i in range(0, len(flat), 6):
        s += '  ' + ', '.join(fmt(v) for v in flat[i:i + 6]) + ',\n'
    return s + '};\n'


def tables_for(dim, order):
    X, verts, basis = build_basis(dim, order)
    n = len(basis)
    curls = [curl(p, X) for p in basis]
    pairs = sym_pairs(dim)
    mass, crl = [], []
    for i in range(n):
        for j in range(n):
            for (a, b) in pairs:
                v = simplex_integral(basis[i][a] * basis[j][b], X)
                if a != b:
                    v += simplex_integral(basis[i][b] * basis[j][a], X)
                mass.append(v)
            if dim == 3:
                for (a, b) in pairs:
                    v = simplex_integral(curls[i][a] * curls[j][b], X)
                    if a != b:
                        v += simplex_integral(curls[i][b] * curls[j][a], X)
                    crl.append(v)
            else:
                crl.append(simplex_integral(curls[i][0] * curls[j][0], X))
    facets = TET_FACES if dim == 3 else TRI_EDGES
    S = sp.symbols('s t')[:dim - 1]
    fac = []
    for fv in facets:
        p = verts[fv[0]]
        for m, v in enumerate(fv[1:]):
            p = p + S[m] * (verts[v] - verts[fv[0]])
        sub = dict(zip(X, p))
        R = [b.subs(sub) for b in basis]
        for i in range(n):
            for j in range(n):
                for (a, b) in pairs:
                    v = simplex_integral(R[i][a] * R[j][b], S)
                    if a != b:
                        v += simplex_integral(R[i][b] * R[j][a], S)
                    fac.append(v)
    tag = '%s%d' % ('Tet' if dim == 3 else 'Tri', order)
    np_ = len(pairs)
    out = emit_array('kMass' + tag, [n, n, np_], mass)
    out += emit_array('kCurl' + tag, [n, n, np_] if dim == 3 else [n, n], crl)
    out += emit_array('kFacet' + tag, [len(facets), n, n, np_], fac)
    return out


def basis_module():
    """sucemfem/_generated_basis.py: monomial coefficients of the nodal basis functions on the reference tetrahedron
    (orders 1, 2), for point evaluation on the host (sources)."""
    out = '"""GENERATED by tools/gen_tables.py -- do not edit.  BASIS[order] = (exponents [n_mono][3], coefficients\n'
    out += '[ndof][3][n_mono]): reference basis function j, component d = sum_m coef[j][d][m] * x^e0 y^e1 z^e2."""\n'
    out += 'BASIS = {}\n'
    for order in (1, 2):
        X, verts, basis = build_basis(3, order)
        monos = sorted({m for phi in basis for d in range(3) for m in sp.Poly(phi[d], *X).monoms()})
        out += 'BASIS[%d] = (%r,\n [' % (order, [list(m) for m in monos])
        for phi in basis:
            rows = []
            for d in range(3):
                poly = sp.Poly(phi[d], *X)
                cd = dict(zip(poly.monoms(), poly.coeffs()))
                rows.append([float(cd.get(m, 0)) for m in monos])
            out += repr(rows) + ',\n  '
        out += '])\n'
    return out


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, '..', 'sucemfem', '_generated_basis.py'), 'w') as f:
        f.write(basis_module())
    dst = os.path.join(here, '..', 'csrc', 'generated', 'n1curl_tables.inc')
    out = '// GENERATED by tools/gen_tables.py -- do not edit.  See that file for definitions.\n'
    for dim in (3, 2):
        for order in (1, 2):
            sys.stderr.write('dim %d order %d\n' % (dim, order))
            out += tables_for(dim, order)
    with open(dst, 'w') as f:
        f.write(out)


if __name__ == '__main__':
    main()
