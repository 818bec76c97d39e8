"""Point current source (electric dipole) RHS contribution
(mirror of /root/reference/sucemfem/Sources/point_source.py:28-94).

The reference asks DOLFIN to locate the cell (CGAL intersection operator), tabulate its dofs and evaluate all basis
functions at the point.  Here the dofs come from the device dof map (sfem_space_cell_dofs); the basis is evaluated on
the host from the generated nodal-basis polynomials (sucemfem/_generated_basis.py, tools/gen_tables.py) with the
covariant Piola map phi = J^-T phihat.  Sources are O(#points) work and outside the assembly hot path.
"""
# Interface (module, class, method and argument names) after the file of cemagg/sucem-fem cited above:
# Copyright (C) 2011 Stellenbosch University, authors Neilen Marais; GNU GPL v3 or later
# (see NOTICE.md at the repository root).  The implementation below is this repository's own.
import numpy as N

from .current_source import CurrentSource
from .._generated_basis import BASIS


class PointCurrentSource(CurrentSource):
    def set_position(self, position):
        self.position = N.array(position, dtype=N.float64)

    def set_value(self, value):
        iscomplex = N.any(N.iscomplex(value))
        dtype = N.complex128 if iscomplex else N.float64
        self.value = N.array(value, dtype=dtype)

    def get_contribution(self):
        return calc_pointsource_contrib(self.function_space, self.position, self.value)


def _locator(V):
    """Per function space cache: cell vertex coordinates, inverse Jacobians, cell dofs (host)."""
    cache = getattr(V, '_point_locator', None)
    if cache is None:
        mesh = V.mesh()
        x = mesh.coordinates()[mesh.cells()]                           # [C, 4, 3]
        J = N.swapaxes(x[:, 1:, :] - x[:, :1, :], 1, 2)
        cache = V._point_locator = dict(x0=x[:, 0, :].copy(), Jinv=N.linalg.inv(J),
                                        cell_dofs=V.dofmap().cell_dofs_array())
    return cache


def locate_cell(V, point, tol=1e-12):
    """Index of the lowest-numbered cell containing the point, and the point's reference coordinates in it."""
    loc = _locator(V)
    X = N.einsum('cap,cp->ca', loc['Jinv'], point[None, :] - loc['x0'])
    inside = (X.min(axis=1) >= -tol) & (1 - X.sum(axis=1) >= -tol)
    hits = N.nonzero(inside)[0]
    if len(hits) == 0:
        raise ValueError('point %r is outside the mesh' % (point,))
    return int(hits[0]), X[hits[0]]


def evaluate_basis_all(V, cell_index, Xref):
    """Physical basis function values [ndof_cell, 3] of cell `cell_index` at reference point Xref."""
    exps, coef = BASIS[V.order]
    exps, coef = N.asarray(exps), N.asarray(coef)
    monos = N.prod(Xref[None, :] ** exps, axis=1)                      # [n_mono]
    ref = coef @ monos                                                 # [ndof, 3]
    return ref @ _locator(V)['Jinv'][cell_index]                       # phi = J^-T phihat


def calc_pointsource_contrib(V, source_coords, source_value):
    """(dofnos, rhs_contribs): RHS[dofnos] += rhs_contribs adds the point source to the system's RHS."""
    if V.mesh().geometry().dim() != 3:
        raise NotImplementedError('point sources are provided for tetrahedral meshes')
    source_coords = N.asarray(source_coords, dtype=N.float64)
    source_value = N.asarray(source_value)
    cell_index, Xref = locate_cell(V, source_coords)
    dofnos = _locator(V)['cell_dofs'][cell_index].astype(N.uintc)
    el_basis_vals = evaluate_basis_all(V, cell_index, Xref)
    rhs_contribs = N.sum(el_basis_vals * source_value, axis=1)
    return dofnos, rhs_contribs
