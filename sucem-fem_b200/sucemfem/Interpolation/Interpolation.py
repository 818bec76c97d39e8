"""Complex surface interpolant on a 3-D domain (mirror of /root/reference/sucemfem/Interpolation/Interpolation.py:26-83).

The reference applies dolfin.DirichletBC(V, expression, on_boundary) to a zero vector, i.e. it evaluates the element's
dof functionals on the expression for every dof in the closure of an exterior facet.  The functionals are the
un-normalised tangential point evaluations of SURVEY 8.1-B: edge (a<b), k = 0,1: u(x_a + (k+1)/3 (x_b - x_a)).(x_b - x_a)
(order 1: the mid point); face (a<b<c), k = 0,1: u(centroid).t_k, t_0 = x_b - x_a, t_1 = x_c - x_a.
"""
# Interface (module, class, method and argument names) after the file of cemagg/sucem-fem cited above:
# Copyright (C) 2011 Stellenbosch University, authors Neilen Marais; GNU GPL v3 or later
# (see NOTICE.md at the repository root).  The implementation below is this repository's own.
import numpy as np
import torch

__all__ = ['SurfaceInterpolant']

_TET_EDGES = ((2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1))
_TET_FACES = ((1, 2, 3), (0, 2, 3), (0, 1, 3), (0, 1, 2))


class _CallableExpression(object):
    def __init__(self, fn, part):
        self.fn, self.part = fn, part

    def __call__(self, pts):
        return np.array([getattr(np.asarray(self.fn(p)), self.part) for p in pts])


class SurfaceInterpolant(object):
    """Calculate a complex surface interpolant on a 3D domain"""

    def __init__(self, function_space):
        self.function_space = function_space
        self._topology = None

    def set_interpolant(self, interpolant):
        """Set interpolant f([x,y,z]) -> [f_x, f_y, f_z] (complex)"""
        self.interpolant = interpolant
        self.interpolant_expression_Re = _CallableExpression(interpolant, 'real')
        self.interpolant_expression_Im = _CallableExpression(interpolant, 'imag')

    def set_interpolant_expression(self, expr_r, expr_i):
        """Real and imaginary parts as vectorised expressions: expr(points [n,3]) -> values [n,3]."""
        self.interpolant_expression_Re = expr_r
        self.interpolant_expression_Im = expr_i

    def _get_topology(self):
        if self._topology is None:
            V = self.function_space
            sp = V.device_space
            mesh = V.mesh()
            if mesh.geometry().dim() != 3:
                raise NotImplementedError('SurfaceInterpolant is provided for tetrahedral meshes')
            co, ce = mesh.coordinates(), mesh.cells()
            cell_edges, cell_faces = (t.cpu().numpy() for t in sp.cell_entities())
            edges = np.zeros((sp.n_edges, 2), np.int64)
            for k, (a, b) in enumerate(_TET_EDGES):
                edges[cell_edges[:, k], 0] = ce[:, a]
                edges[cell_edges[:, k], 1] = ce[:, b]
            faces = np.zeros((sp.n_facets, 3), np.int64)
            for k, (a, b, c) in enumerate(_TET_FACES):
                faces[cell_faces[:, k]] = np.stack([ce[:, a], ce[:, b], ce[:, c]], axis=1)
            markers = torch.zeros(sp.n_facets, dtype=torch.int32, device=sp.mesh.device)
            sp.mark_exterior_facets(markers, 1)
            on_boundary = sp.pec_row_flags(markers, 1).cpu().numpy().astype(bool)
            self._topology = (co, edges, faces, on_boundary)
        return self._topology

    def calculate_interpolation(self):
        V = self.function_space
        sp = V.device_space
        q = V.order
        co, edges, faces, on_boundary = self._get_topology()
        fun = lambda pts: (np.asarray(self.interpolant_expression_Re(pts), dtype=np.float64)
                           + 1j * np.asarray(self.interpolant_expression_Im(pts), dtype=np.float64))
        x = np.zeros(sp.ndof, np.complex128)
        E = sp.n_edges
        e_b = np.nonzero(on_boundary[q * np.arange(E)])[0]
        xa, xb = co[edges[e_b, 0]], co[edges[e_b, 1]]
        for k in range(q):
            pts = xa + ((k + 1.0) / (q + 1.0)) * (xb - xa)
            x[q * e_b + k] = np.einsum('nd,nd->n', fun(pts), xb - xa)
        if q == 2:
            f_b = np.nonzero(on_boundary[2 * E + 2 * np.arange(len(faces))])[0]
            fa, fb, fc = co[faces[f_b, 0]], co[faces[f_b, 1]], co[faces[f_b, 2]]
            v = fun((fa + fb + fc) / 3.0)
            x[2 * E + 2 * f_b] = np.einsum('nd,nd->n', v, fb - fa)
            x[2 * E + 2 * f_b + 1] = np.einsum('nd,nd->n', v, fc - fa)
        return x
