"""TEST INFRASTRUCTURE -- CPU oracle, never imported by the product path.

Exact (rational) N1curl reference elements of order 1 and 2 on the UFC reference
tetrahedron / triangle, and the pre-integrated reference tensors the oracle
contracts with per-cell geometry.

The reference (cemagg/sucem-fem) never defines the element itself: it asks DOLFIN
for ``FunctionSpace(mesh, "Nedelec 1st kind H(curl)", order)``
(/root/reference/sucemfem/ProblemConfigurations/EMProblem.py:36,100-105) and
FFC/FIAT supply the basis.  FIAT is an un-vendored, un-pinned dependency, so this
file restates FIAT's published construction (nodal basis dual to tangential point
evaluations on edges and tangential point evaluations at face centroids, tangents
un-normalised, entities oriented low->high vertex), and is pinned against the
reference's own fixtures in tests/test_oracle_golden.py.

Conventions restated here:
  * reference tet  v0=(0,0,0) v1=(1,0,0) v2=(0,1,0) v3=(0,0,1)
  * UFC edges      e0=(2,3) e1=(1,3) e2=(1,2) e3=(0,3) e4=(0,2) e5=(0,1)
  * UFC faces      f_i opposite v_i:  f0=(1,2,3) f1=(0,2,3) f2=(0,1,3) f3=(0,1,2)
  * triangle edges e_i opposite v_i:  e0=(1,2) e1=(0,2) e2=(0,1)
  * order 1: l_e(u)     = u(mid_e) . (x_b - x_a)
  * order 2: l_{e,k}(u) = u(x_a + (k+1)/3 (x_b-x_a)) . (x_b - x_a),   k=0,1
             l_{f,k}(u) = u(centroid_f) . t_k,  t_0=x_b-x_a, t_1=x_c-x_a
  * local dof order: entity-major, k-minor: [e0k0,e0k1,...,e5k1,f0k0,...,f3k1]
"""
from fractions import Fraction as Fr
from functools import lru_cache
from math import factorial

import numpy as np

TET_EDGES = ((2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1))
TET_FACES = ((1, 2, 3), (0, 2, 3), (0, 1, 3), (0, 1, 2))
TRI_EDGES = ((1, 2), (0, 2), (0, 1))


# --------------------------------------------------------------------------------------
# tiny exact polynomial algebra: poly = {exponent tuple: Fraction}
# --------------------------------------------------------------------------------------
def p_add(a, b, sb=1):
    r = dict(a)
    for k, v in b.items():
        r[k] = r.get(k, 0) + sb * v
        if r[k] == 0:
            del r[k]
    return r


def p_mul(a, b):
    r = {}
    for ka, va in a.items():
        for kb, vb in b.items():
            k = tuple(x + y for x, y in zip(ka, kb))
            r[k] = r.get(k, 0) + va * vb
    return {k: v for k, v in r.items() if v != 0}


def p_scale(a, s):
    return {k: v * s for k, v in a.items()} if s != 0 else {}


def p_diff(a, d):
    r = {}
    for k, v in a.items():
        if k[d] > 0:
            kk = list(k)
            kk[d] -= 1
            r[tuple(kk)] = r.get(tuple(kk), 0) + v * k[d]
    return r


def p_eval(a, x):
    s = Fr(0)
    for k, v in a.items():
        t = v
        for xi, ki in zip(x, k):
            t = t * xi ** ki
        s += t
    return s


def p_int_simplex(a):
    """Integral over the unit simplex: int x^a y^b z^c = a! b! c! / (a+b+c+dim)!"""
    s = Fr(0)
    for k, v in a.items():
        num = 1
        for ki in k:
            num *= factorial(ki)
        s += v * Fr(num, factorial(sum(k) + len(k)))
    return s


def p_const(dim, c):
    return {(0,) * dim: Fr(c)} if c != 0 else {}


def p_var(dim, d):
    k = [0] * dim
    k[d] = 1
    return {tuple(k): Fr(1)}


def p_compose_affine(a, origin, dirs, ndim_out):
    """Substitute x_d = origin[d] + sum_m dirs[m][d] * s_m; result is a poly in s (ndim_out vars)."""
    dim = len(origin)
    lin = []
    for d in range(dim):
        q = p_const(ndim_out, origin[d])
        for m in range(ndim_out):
            q = p_add(q, p_scale(p_var(ndim_out, m), dirs[m][d]))
        lin.append(q)
    out = {}
    for k, v in a.items():
        term = p_const(ndim_out, v)
        for d in range(dim):
            for _ in range(k[d]):
                term = p_mul(term, lin[d])
        out = p_add(out, term)
    return out


# --------------------------------------------------------------------------------------
# the element
# --------------------------------------------------------------------------------------
def _ref_vertices(dim):
    z = [Fr(0)] * dim
    vs = [tuple(z)]
    for d in range(dim):
        v = list(z)
        v[d] = Fr(1)
        vs.append(tuple(v))
    return vs


def _space_generators(dim, order):
    """Generators (vector polynomials) of the first-kind Nedelec space R_order = P_{order-1}^dim + S_order."""
    X = [p_var(dim, d) for d in range(dim)]
    one = p_const(dim, 1)
    gens = []
    monos = [one] if order == 1 else [one] + X  # basis of P_{order-1}
    for m in monos:
        for d in range(dim):
            gens.append(tuple(m if dd == d else {} for dd in range(dim)))
    homs = [one] if order == 1 else X  # homogeneous degree order-1
    if dim == 3:
        # X x (h e_d); at order 2 the combination sum_d X x (x_d e_d) = X x X = 0, so drop the last one
        cand = []
        for h in homs:
            for d in range(3):
                q = [{} for _ in range(3)]
                q[d] = h
                cr = (p_add(p_mul(X[1], q[2]), p_mul(X[2], q[1]), -1),
                      p_add(p_mul(X[2], q[0]), p_mul(X[0], q[2]), -1),
                      p_add(p_mul(X[0], q[1]), p_mul(X[1], q[0]), -1))
                cand.append(cr)
        if order == 2:
            cand = cand[:-1]
        gens += cand
    else:
        for h in homs:  # rot(X) h = (-y, x) h
            gens.append((p_scale(p_mul(X[1], h), -1), p_mul(X[0], h)))
    return gens


def _functionals(dim, order, interior_scale=None):
    """List of (point, direction) tangential point evaluations, in local dof order, plus
    (2-D order 2 only) interior moment functionals given as ('moment', component, scale)."""
    V = _ref_vertices(dim)
    sub = lambda a, b: tuple(x - y for x, y in zip(a, b))
    lerp = lambda a, b, t: tuple(x + t * (y - x) for x, y in zip(a, b))
    edges = TET_EDGES if dim == 3 else TRI_EDGES
    fs = []
    for (a, b) in edges:
        t = sub(V[b], V[a])
        if order == 1:
            fs.append((lerp(V[a], V[b], Fr(1, 2)), t))
        else:
            for k in range(2):
                fs.append((lerp(V[a], V[b], Fr(k + 1, 3)), t))
    if order == 2:
        if dim == 3:
            for (a, b, c) in TET_FACES:
                cen = tuple((x + y + z) / 3 for x, y, z in zip(V[a], V[b], V[c]))
                fs.append((cen, sub(V[b], V[a])))
                fs.append((cen, sub(V[c], V[a])))
        else:
            # interior moments against the constant vector fields e_x, e_y.  FIAT scales them by an
            # internal orthonormalisation constant that no reference fixture pins (SURVEY 8.1-2D,
            # [UNPINNED]); the scale only rescales the two interior basis functions.
            fs.append(('moment', 0))
            fs.append(('moment', 1))
    return fs


@lru_cache(maxsize=None)
def nodal_basis(dim, order):
    """Exact nodal basis: list of vector polynomials phi_j (tuples of dim polys)."""
    gens = _space_generators(dim, order)
    fs = _functionals(dim, order)
    n = len(fs)
    assert len(gens) == n, (len(gens), n)
    Vm = [[Fr(0)] * n for _ in range(n)]  # Vm[i][m] = l_i(gen_m)
    for i, f in enumerate(fs):
        for m, g in enumerate(gens):
            if f[0] == 'moment':
                Vm[i][m] = p_int_simplex(g[f[1]])
            else:
                pt, t = f
                Vm[i][m] = sum(p_eval(g[d], pt) * t[d] for d in range(dim))
    # invert exactly (Gauss-Jordan on Fractions)
    A = [row[:] + [Fr(int(i == j)) for j in range(n)] for i, row in enumerate(Vm)]
    for c in range(n):
        piv = next(r for r in range(c, n) if A[r][c] != 0)
        A[c], A[piv] = A[piv], A[c]
        pv = A[c][c]
        A[c] = [x / pv for x in A[c]]
        for r in range(n):
            if r != c and A[r][c] != 0:
                f_ = A[r][c]
                A[r] = [x - f_ * y for x, y in zip(A[r], A[c])]
    inv = [row[n:] for row in A]  # inv[m][j]: phi_j = sum_m inv[m][j] gen_m
    basis = []
    for j in range(n):
        comp = [{} for _ in range(dim)]
        for m in range(n):
            if inv[m][j] != 0:
                for d in range(dim):
                    comp[d] = p_add(comp[d], p_scale(gens[m][d], inv[m][j]))
        basis.append(tuple(comp))
    return basis


def curl_of(phi, dim):
    if dim == 3:
        return (p_add(p_diff(phi[2], 1), p_diff(phi[1], 2), -1),
                p_add(p_diff(phi[0], 2), p_diff(phi[2], 0), -1),
                p_add(p_diff(phi[1], 0), p_diff(phi[0], 1), -1))
    return (p_add(p_diff(phi[1], 0), p_diff(phi[0], 1), -1),)


def _to_np(t):
    return np.array(t, dtype=object).astype(np.float64)


@lru_cache(maxsize=None)
def mass_tensor(dim, order):
    """Mhat[a,b,i,j] = int_Khat phi_i^a phi_j^b  (fp64, correctly rounded from exact rationals)."""
    B = nodal_basis(dim, order)
    n = len(B)
    out = np.zeros((dim, dim, n, n))
    for a in range(dim):
        for b in range(dim):
            for i in range(n):
                for j in range(n):
                    out[a, b, i, j] = float(p_int_simplex(p_mul(B[i][a], B[j][b])))
    return out


@lru_cache(maxsize=None)
def curl_tensor(dim, order):
    """Chat[a,b,i,j] = int_Khat (curl phi_i)^a (curl phi_j)^b ; in 2-D the curl is a scalar (a=b=0)."""
    B = nodal_basis(dim, order)
    C = [curl_of(p, dim) for p in B]
    n = len(B)
    cd = 3 if dim == 3 else 1
    out = np.zeros((cd, cd, n, n))
    for a in range(cd):
        for b in range(cd):
            for i in range(n):
                for j in range(n):
                    out[a, b, i, j] = float(p_int_simplex(p_mul(C[i][a], C[j][b])))
    return out


@lru_cache(maxsize=None)
def facet_tensor(dim, order):
    """Bhat[lf,a,b,i,j] = int over the unit (dim-1)-simplex of phi_i^a phi_j^b restricted to local facet lf,
    parametrised as p0 + sum_m s_m (p_m - p0).  Physical integral = (facet measure * (dim-1)!) * Bhat."""
    B = nodal_basis(dim, order)
    V = _ref_vertices(dim)
    facets = TET_FACES if dim == 3 else TRI_EDGES
    n = len(B)
    out = np.zeros((len(facets), dim, dim, n, n))
    for lf, fv in enumerate(facets):
        p0 = V[fv[0]]
        dirs = [tuple(x - y for x, y in zip(V[v], p0)) for v in fv[1:]]
        R = [[p_compose_affine(B[i][a], p0, dirs, dim - 1) for a in range(dim)] for i in range(n)]
        for a in range(dim):
            for b in range(dim):
                for i in range(n):
                    for j in range(n):
                        out[lf, a, b, i, j] = float(p_int_simplex(p_mul(R[i][a], R[j][b])))
    return out


def eval_basis(dim, order, Xref):
    """Reference basis values at reference points Xref [npts, dim] -> [npts, ndof, dim] (fp64)."""
    B = nodal_basis(dim, order)
    Xref = np.atleast_2d(np.asarray(Xref, dtype=np.float64))
    out = np.zeros((Xref.shape[0], len(B), dim))
    for j, phi in enumerate(B):
        for d in range(dim):
            for k, v in phi[d].items():
                term = float(v) * np.ones(Xref.shape[0])
                for dd in range(dim):
                    if k[dd]:
                        term = term * Xref[:, dd] ** k[dd]
                out[:, j, d] += term
    return out


def ndofs_per_cell(dim, order):
    ne = 6 if dim == 3 else 3
    if order == 1:
        return ne
    return 2 * ne + (8 if dim == 3 else 2)
