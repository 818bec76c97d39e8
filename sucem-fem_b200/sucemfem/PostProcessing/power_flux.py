"""Power flux through the bounding surface (mirror of /root/reference/sucemfem/PostProcessing/power_flux.py:11-133).

SurfaceFlux integrates the Poynting vector over the exterior facets (host, O(boundary) work, basis and curl from the
generated nodal polynomials); VariationalSurfaceFlux evaluates the same quantity through the EM functional over the skin
of cells connected to the boundary (device matrices + SpMV, see PostProcessing.CalcEMFunctional).
"""
# Interface (module, class, method and argument names) after the file of cemagg/sucem-fem cited above:
# Copyright (C) 2011 Stellenbosch University, authors Neilen Marais; GNU GPL v3 or later
# (see NOTICE.md at the repository root).  The implementation below is this repository's own.
import numpy as np
import torch

from ..Consts import Z0
from .. import Geometry
from .. import dolfin_compat as dolfin
from . import CalcEMFunctional

from ._facets import ExteriorFacets, tri_rule


class SurfaceFlux(object):
    """Real part of the power flux through the problem's bounding surface by a surface integral of the Poynting
    vector (power_flux.py:11-61)."""

    def __init__(self, function_space):
        self.function_space = function_space
        self.mur_function = None

    def set_dofs(self, dofs):
        self.dofs = np.asarray(dofs, dtype=np.complex128)

    def set_k0(self, k0):
        self.k0 = k0

    def set_mur_function(self, mur_function):
        self.mur_function = mur_function

    def calc_flux(self):
        fac = ExteriorFacets(self.function_space)
        mur = np.ones(len(fac.cell)) if self.mur_function is None else self.mur_function.vector()[fac.cell]
        pts, wts = tri_rule(3)          # exact to degree 5 (the integrand n.(E x curl E) has degree 3 at order 2)
        total = 0.0
        for sel, X, nrm, pos in fac.groups(pts):
            E, C = fac.fields(sel, X, self.dofs)
            Er, Ei, Cr, Ci = E.real, E.imag, C.real, C.imag
            integrand = np.einsum('nd,npd->np', nrm, np.cross(Er, -Ci) + np.cross(Ei, Cr)) / mur[sel][:, None]
            total += float(np.sum(integrand * wts[None, :]))
        return total / self.k0 / Z0


class VariationalSurfaceFlux(object):
    """The flux through the functional over the skin of boundary-connected cells (power_flux.py:63-133)."""

    def __init__(self, function_space):
        self.function_space = V = function_space
        self.mur_function = None
        self.epsr_function = None
        boundary_cells = Geometry.BoundaryEdgeCells(V.mesh())
        cell_domains = dolfin.CellFunction('uint', V.mesh())
        cell_domains.set_all(0)
        cell_region = 1
        boundary_cells.mark(cell_domains, cell_region)
        self.functional = CalcEMFunctional(V)
        self.functional.set_cell_domains(cell_domains, cell_region)
        self.k0_dirty = True

    def set_dofs(self, dofs):
        self.dofs = np.asarray(dofs, dtype=np.complex128)
        # the solution's dofs on the boundary (DirichletBC(V, E, DomainBoundary()) applied to a zero vector)
        sp = self.function_space.device_space
        markers = torch.zeros(sp.n_facets, dtype=torch.int32, device=sp.mesh.device)
        sp.mark_exterior_facets(markers, 1)
        on_boundary = sp.pec_row_flags(markers, 1).cpu().numpy().astype(bool)
        self.dirich_dofs = np.where(on_boundary, self.dofs, 0.0)
        self.functional.set_E_dofs(dofs)
        self.k0_dirty = True

    def set_k0(self, k0):
        self.k0 = k0
        self.functional.set_k0(k0)
        self.k0_dirty = True

    def set_epsr_function(self, epsr_function):
        self.epsr_function = epsr_function
        self.functional.set_epsr_function(epsr_function)

    def set_mur_function(self, mur_function):
        self.mur_function = mur_function
        self.functional.set_mur_function(mur_function)

    def calc_flux(self):
        if self.k0_dirty:
            self.functional.set_g_dofs(1j * self.dirich_dofs.conjugate() / self.k0 / Z0)
            self.k0_dirty = False
        return self.functional.calc_functional().conjugate()
