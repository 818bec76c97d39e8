"""Field reconstruction and the EM functional (mirror of /root/reference/sucemfem/PostProcessing/PostProcessing.py:25-172).

CalcEMFunctional evaluates <curl E, curl g>/mu_r - k0^2 <eps_r E, g> over the whole mesh or a cell sub-domain.  The
reference assembles two scalar UFL forms; here the functional is g^T (S_dom - k0^2 M_dom) E with S_dom, M_dom from the
same device kernels as the system matrices (coefficients masked to the sub-domain) and the product taken by the fused
two-matrix SpMV (sfem_matvec) -- SURVEY 8.2 row f4.
"""
# Interface (module, class, method and argument names) after the file of cemagg/sucem-fem cited above:
# Copyright (C) 2011 Stellenbosch University, authors Neilen Marais; GNU GPL v3 or later
# (see NOTICE.md at the repository root).  The implementation below is this repository's own.
import numpy as N
import torch

from .. import _native as NV

__all__ = ['Reconstruct', 'CalcEMFunctional']


class Reconstruct(object):
    """Reconstruct field values, dealing with complex numbers as required (PostProcessing.py:27-60)."""

    def __init__(self, function_space):
        self.function_space = function_space

    def set_dof_values(self, x):
        self.dof_values = x

    def reconstruct_points(self, points):
        """Reconstruct the function at points (n x d array); host evaluation of the nodal basis in the cell that
        contains each point (the lowest-numbered one on shared faces)."""
        from ..Sources.point_source import locate_cell, evaluate_basis_all, _locator
        V = self.function_space
        x = N.asarray(self.dof_values)
        cell_dofs = _locator(V)['cell_dofs']
        out = N.zeros((len(points), 3), dtype=x.dtype if N.iscomplexobj(x) and N.any(N.iscomplex(x)) else N.float64)
        for i, fp in enumerate(N.asarray(points, dtype=N.float64)):
            c, Xref = locate_cell(V, fp)
            vals = x[cell_dofs[c]] @ evaluate_basis_all(V, c, Xref)
            out[i, :] = vals if N.iscomplexobj(out) else N.real(vals)
        return out


class CalcEMFunctional(object):
    """Evaluate <curl(E), curl(g)> - k0^2 <E, g> (PostProcessing.py:62-172); E the field solution, g an H(curl)
    testing function given by its dofs."""

    def __init__(self, function_space, testing_space=None):
        if testing_space is not None and testing_space is not function_space:
            raise NotImplementedError('a separate testing space is not provided')
        self.function_space = function_space
        self.testing_space = function_space
        self.cell_domains = None
        self.cell_mark_value = None
        self.epsr_function = None
        self.mur_function = None
        self.dirty = True
        self._matrices = None

    def set_k0(self, k0):
        self.k0 = k0

    def set_cell_domains(self, cell_domains, cell_mark_value):
        """Cell domain mesh function: the functional is then integrated over the marked cells only."""
        self.cell_domains = cell_domains
        self.cell_mark_value = cell_mark_value
        self.dirty = True

    def set_epsr_function(self, epsr_function):
        self.epsr_function = epsr_function
        self.dirty = True

    def set_mur_function(self, mur_function):
        self.mur_function = mur_function
        self.dirty = True

    def set_quadrature_degree(self, quadrature_degree):
        """Accepted for compatibility: the element matrices are integrated exactly."""
        self.quadrature_degree = quadrature_degree

    def set_E_dofs(self, E_dofs):
        self.E_dofs = N.asarray(E_dofs, dtype=N.complex128).copy()

    def set_g_dofs(self, g_dofs):
        self.g_dofs = N.asarray(g_dofs, dtype=N.complex128).copy()

    def _get_matrices(self):
        if self.dirty:
            space = self.function_space.device_space
            dev = space.mesh.device
            ncell = space.mesh.num_cells
            eps = None if self.epsr_function is None else self.epsr_function.device_values()
            # the reference only uses mu_r when eps_r is set as well (PostProcessing.py:130-136)
            mu = None if (self.epsr_function is None or self.mur_function is None) else self.mur_function.device_values()
            if self.cell_domains is not None:
                inside = torch.from_numpy(self.cell_domains.array().astype(N.int64) == int(self.cell_mark_value)).to(dev)
                one = torch.ones(ncell, dtype=torch.float64, device=dev)
                eps = torch.where(inside, one if eps is None else eps, torch.zeros_like(one))
                # 1/mu_r = 0 outside the sub-domain
                mu = torch.where(inside, one if mu is None else mu, torch.full_like(one, float('inf')))
            pat = space.pattern()
            S, M = pat.assemble_ST(eps, mu, None, kernel=NV.KERNEL_AUTO)
            self._matrices = (pat, S, M)
            self.dirty = False
        return self._matrices

    def calc_functional(self):
        """Calculate functional using given E, g and k0 values"""
        pat, S, M = self._get_matrices()
        dev = S.device
        E = torch.from_numpy(self.E_dofs).to(dev)
        g = torch.from_numpy(self.g_dofs).to(dev)
        y = pat.matvec([S, M], E, [1.0, -self.k0 ** 2])
        return complex(torch.sum(g * y).item())
