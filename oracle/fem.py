"""TEST INFRASTRUCTURE -- CPU oracle, never imported by the product path.

numpy restatement of what the reference's two hot-path lines compute
(/root/reference/sucemfem/SystemMatrices.py:54-55: ``dolfin.assemble(form, tensor=mat)`` followed by
``boundary_conditions.apply_essential(mat)``) for the three forms the reference feeds them:

  mass       eps_r*inner(v,u)*dx               /root/reference/sucemfem/Forms.py:71-77
  stiffness  dot(curl(v),curl(u))/mu_r*dx      /root/reference/sucemfem/Forms.py:79-85
  ABC        inner(cross(n,v),cross(n,u))*ds   /root/reference/sucemfem/BoundaryConditions/ABC.py:56-57

The arithmetic itself lives in DOLFIN/FFC/FIAT (un-vendored, no pinned version: the reference has no
setup.py / requirements file); this file restates their published algorithms under the conventions of
SURVEY.md section 8.1 and is pinned against the reference's own golden fixtures in
tests/test_oracle_golden.py (fixtures 1,2,3,5,7 of SURVEY.md section 8c).  What no reference fixture
pins is listed there too (2-D matrix values, the S_0 pattern, row-only ``ident``).

Also restated, because fixtures 2/3/5/7 need it to pin the dof map: the point-source RHS
(/root/reference/sucemfem/Sources/point_source.py:49-94) and the filament source
(/root/reference/sucemfem/Sources/fillament_current_source.py:62-82).
"""
import numpy as np
import scipy.sparse as sp

from . import refelem as R


# --------------------------------------------------------------------------------------
# X1: entity numbering by first appearance  
# --------------------------------------------------------------------------------------
def _first_appearance_numbering(keys2d):
    """keys2d [C, nloc, k] vertex tuples; returns (index [C, nloc], entity_vertices [n_ent, k])."""
    C, nloc, k = keys2d.shape
    flat = keys2d.reshape(C * nloc, k).astype(np.int64)
    uniq, first, inv = np.unique(flat, axis=0, return_index=True, return_inverse=True)
    order = np.argsort(first, kind='stable')          # unique-id -> rank by first appearance
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    return rank[inv.ravel()].reshape(C, nloc).astype(np.int64), uniq[order]


def number_entities(cells):
    """cells [C, dim+1] with ascending vertex ids per cell.  Returns dict with cell_edges, edges, and in
    3-D cell_faces, faces."""
    cells = np.asarray(cells)
    assert np.all(np.diff(cells, axis=1) > 0), "cells must list vertices in ascending order (UFC)"
    dim = cells.shape[1] - 1
    ledges = R.TET_EDGES if dim == 3 else R.TRI_EDGES
    ekeys = np.stack([cells[:, list(e)] for e in ledges], axis=1)
    cell_edges, edges = _first_appearance_numbering(ekeys)
    out = dict(dim=dim, cell_edges=cell_edges, edges=edges, E=len(edges))
    if dim == 3:
        fkeys = np.stack([cells[:, list(f)] for f in R.TET_FACES], axis=1)
        cell_faces, faces = _first_appearance_numbering(fkeys)
        out.update(cell_faces=cell_faces, faces=faces, F=len(faces))
    return out


# --------------------------------------------------------------------------------------
# X2: dof map
# --------------------------------------------------------------------------------------
def cell_dofs(num, order, ncells=None):
    ce = num['cell_edges']
    C = ce.shape[0]
    if order == 1:
        return ce.copy(), num['E']
    E = num['E']
    parts = [2 * ce[:, :, None] + np.arange(2)[None, None, :]]
    if num['dim'] == 3:
        parts.append(2 * E + 2 * num['cell_faces'][:, :, None] + np.arange(2)[None, None, :])
        ndof = 2 * E + 2 * num['F']
    else:
        parts.append(2 * E + 2 * np.arange(C)[:, None, None] + np.arange(2)[None, None, :])
        ndof = 2 * E + 2 * C
    return np.concatenate([p.reshape(C, -1) for p in parts], axis=1), ndof


# --------------------------------------------------------------------------------------
# X3: sparsity pattern
# --------------------------------------------------------------------------------------
def pattern(cd, ndof):
    C, k = cd.shape
    rows = np.repeat(cd, k, axis=1).ravel()
    cols = np.tile(cd, (1, k)).ravel()
    keys = np.unique(rows.astype(np.int64) * ndof + cols)
    r = keys // ndof
    rowptr = np.zeros(ndof + 1, np.int64)
    np.add.at(rowptr, r + 1, 1)
    return np.cumsum(rowptr), (keys % ndof).astype(np.int32)


# --------------------------------------------------------------------------------------
# X4: cell kernels
# --------------------------------------------------------------------------------------
def geometry(coords, cells):
    x = coords[cells]                                   # [C, dim+1, dim]
    J = np.swapaxes(x[:, 1:, :] - x[:, :1, :], 1, 2)    # columns = x_vi - x_v0
    det = np.linalg.det(J)
    Jinv = np.linalg.inv(J)
    Ginv = Jinv @ np.swapaxes(Jinv, 1, 2)
    G = np.swapaxes(J, 1, 2) @ J
    return J, det, Jinv, Ginv, G


def element_mass(coords, cells, order, eps_r=None):
    dim = cells.shape[1] - 1
    _, det, _, Ginv, _ = geometry(coords, cells)
    Mh = R.mass_tensor(dim, order)
    Ae = np.einsum('cab,abij->cij', Ginv, Mh) * np.abs(det)[:, None, None]
    if eps_r is not None:
        Ae = Ae * np.asarray(eps_r)[:, None, None]
    return Ae


def element_stiffness(coords, cells, order, mu_r=None):
    dim = cells.shape[1] - 1
    _, det, _, _, G = geometry(coords, cells)
    Ch = R.curl_tensor(dim, order)
    if dim == 3:
        Ae = np.einsum('cab,abij->cij', G, Ch) / np.abs(det)[:, None, None]
    else:
        Ae = Ch[0, 0][None, :, :] / np.abs(det)[:, None, None]
    if mu_r is not None:
        Ae = Ae / np.asarray(mu_r)[:, None, None]
    return Ae


def add_into_pattern(rowptr, cols, ndof, cd, Ae, vals=None):
    """A[dofs,dofs] += A_e in ascending cell order, into the slots of a prebuilt CSR pattern."""
    C, k = cd.shape
    keys_pat = np.repeat(np.arange(ndof, dtype=np.int64), np.diff(rowptr)) * ndof + cols
    rows = np.repeat(cd, k, axis=1).ravel().astype(np.int64)
    cls = np.tile(cd, (1, k)).ravel()
    slot = np.searchsorted(keys_pat, rows * ndof + cls)
    assert np.all(keys_pat[slot] == rows * ndof + cls)
    if vals is None:
        vals = np.zeros(len(cols))
    np.add.at(vals, slot, Ae.reshape(-1))
    return vals


# --------------------------------------------------------------------------------------
# X5: exterior facets and the ABC form
# --------------------------------------------------------------------------------------
def exterior_facets(cells, num):
    """(cell, local facet) pairs of facets that belong to exactly one cell, in ascending cell order."""
    dim = cells.shape[1] - 1
    cf = num['cell_faces'] if dim == 3 else num['cell_edges']
    nf = num['F'] if dim == 3 else num['E']
    cnt = np.bincount(cf.ravel(), minlength=nf)
    c, lf = np.nonzero(cnt[cf] == 1)
    return c, lf


def element_abc(coords, cells, order, ext_c, ext_lf):
    """Per exterior facet: full k x k cell block of int_f (n x phi_i).(n x phi_j) ds."""
    dim = cells.shape[1] - 1
    Bh = R.facet_tensor(dim, order)
    _, det, Jinv, Ginv, _ = geometry(coords, cells)
    facets = R.TET_FACES if dim == 3 else R.TRI_EDGES
    fv = np.array(facets)[ext_lf]                       # local vertex numbers of each facet
    x = coords[cells[ext_c][np.arange(len(ext_c))[:, None], fv]]   # [nf, dim, dim]
    if dim == 3:
        nrm = np.cross(x[:, 1] - x[:, 0], x[:, 2] - x[:, 0])
        meas2 = np.linalg.norm(nrm, axis=1)             # = 2 * area = area * (dim-1)!
    else:
        t = x[:, 1] - x[:, 0]
        nrm = np.stack([t[:, 1], -t[:, 0]], axis=1)
        meas2 = np.linalg.norm(nrm, axis=1)             # = length
    n = nrm / meas2[:, None]
    nu = np.einsum('fap,fp->fa', Jinv[ext_c], n)        # J^-1 n
    K = Ginv[ext_c] - nu[:, :, None] * nu[:, None, :]
    return np.einsum('fab,fabij->fij', K, Bh[ext_lf]) * meas2[:, None, None]


# --------------------------------------------------------------------------------------
# X6: PEC
# --------------------------------------------------------------------------------------
def facet_closure_dofs(cells, num, order, facet_mask):
    """Sorted unique dofs in the closure of the facets selected by facet_mask (indexed by the 8.1-N facet
    numbering: faces in 3-D, edges in 2-D)."""
    dim = num['dim']
    q = order
    if dim == 2:
        e = np.nonzero(facet_mask)[0]
        return np.unique((q * e[:, None] + np.arange(q)[None, :]).ravel())
    cf, ce = num['cell_faces'], num['cell_edges']
    sel = facet_mask[cf]                                 # [C,4]
    c, lf = np.nonzero(sel)
    # edges of local face lf: the local edges whose two vertices are both in the face
    face_edges = np.array([[le for le, ev in enumerate(R.TET_EDGES) if set(ev) <= set(fv)]
                           for fv in R.TET_FACES])
    ed = ce[c[:, None], face_edges[lf]]                  # [n,3]
    dofs = [(q * ed[:, :, None] + np.arange(q)[None, None, :]).ravel()]
    if order == 2:
        f = cf[c, lf]
        dofs.append((2 * num['E'] + 2 * f[:, None] + np.arange(2)[None, :]).ravel())
    return np.unique(np.concatenate(dofs))


def apply_ident(rowptr, cols, vals, rows):
    """Row-only ident: zero the row, 1.0 on the diagonal, pattern kept, columns untouched."""
    vals = vals.copy()
    for r in rows:
        s, e = rowptr[r], rowptr[r + 1]
        vals[s:e] = 0.0
        d = s + np.searchsorted(cols[s:e], r)
        assert cols[d] == r
        vals[d] = 1.0
    return vals


# --------------------------------------------------------------------------------------
# the whole path: what calc_system_matrices returns for {M, S, S_0}
# --------------------------------------------------------------------------------------
def assemble_system(coords, cells, order, eps_r=None, mu_r=None, abc=False, pec_facet_mask=None,
                    forms=('M', 'S')):
    coords = np.asarray(coords, np.float64)
    cells = np.asarray(cells)
    num = number_entities(cells)
    cd, ndof = cell_dofs(num, order)
    rowptr, cols = pattern(cd, ndof)
    out = dict(num=num, cell_dofs=cd, ndof=ndof, rowptr=rowptr, cols=cols)
    if 'M' in forms:
        out['M'] = add_into_pattern(rowptr, cols, ndof, cd, element_mass(coords, cells, order, eps_r))
    if 'S' in forms:
        out['S'] = add_into_pattern(rowptr, cols, ndof, cd, element_stiffness(coords, cells, order, mu_r))
    if abc:
        ec, elf = exterior_facets(cells, num)
        out['S_0'] = add_into_pattern(rowptr, cols, ndof, cd[ec], element_abc(coords, cells, order, ec, elf))
    if pec_facet_mask is not None:
        rows = facet_closure_dofs(cells, num, order, pec_facet_mask)
        out['pec_dofs'] = rows
        for k in ('M', 'S', 'S_0'):
            if k in out:
                out[k] = apply_ident(rowptr, cols, out[k], rows)
    return out


def exterior_facet_mask(cells, num):
    dim = num['dim']
    cf = num['cell_faces'] if dim == 3 else num['cell_edges']
    nf = num['F'] if dim == 3 else num['E']
    return np.bincount(cf.ravel(), minlength=nf) == 1


def to_scipy(rowptr, cols, vals, ndof):
    return sp.csr_matrix((vals, cols, rowptr), shape=(ndof, ndof))


# --------------------------------------------------------------------------------------
# sources (only to pin the dof map / basis against fixtures 2,3,5,7)
# --------------------------------------------------------------------------------------
def locate_cells(coords, cells, pt, tol=1e-12):
    """All cells containing pt (barycentric coordinates >= -tol), ascending."""
    x0 = coords[cells[:, 0]]
    J, det, Jinv, _, _ = geometry(coords, cells)
    X = np.einsum('cap,cp->ca', Jinv, pt[None, :] - x0)
    lam0 = 1 - X.sum(axis=1)
    ok = (X.min(axis=1) >= -tol) & (lam0 >= -tol)
    return np.nonzero(ok)[0], X


def point_source_contrib(coords, cells, cd, order, pt, value, cell_index=None):
    """dofnos, contribs of a point current source: contrib_i = phi_i(pt) . value."""
    dim = cells.shape[1] - 1
    pt = np.asarray(pt, np.float64)
    hits, X = locate_cells(coords, cells, pt)
    c = hits[0] if cell_index is None else cell_index
    _, _, Jinv, _, _ = geometry(coords, cells[c:c + 1])
    ref = R.eval_basis(dim, order, X[c])[0]              # [ndof, dim]  reference values
    phys = ref @ Jinv[0]                                  # phi = J^-T phihat  ->  row-vector times J^-1
    return cd[c], phys @ np.asarray(value)


def filament_source_contrib(coords, cells, cd, order, endpoints, I, npts, pick='first'):
    """fillament_current_source.py:62-82: N-point sum of point sources along the filament."""
    start, end = np.asarray(endpoints, np.float64)
    delta = end - start
    length = np.sqrt(np.sum(delta ** 2))
    vec = delta / length * I
    pts = [start + delta * 0.5] if npts == 1 else start + delta * np.linspace(0, 1, npts)[:, None]
    mag = vec * length / npts
    acc = {}
    for p in pts:
        dn, v = point_source_contrib(coords, cells, cd, order, p, mag)
        for d, x in zip(dn, v):
            acc[int(d)] = acc.get(int(d), 0.) + x
    dn = np.array(sorted(acc))
    return dn, np.array([acc[d] for d in dn])
