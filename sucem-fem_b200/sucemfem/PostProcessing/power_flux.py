"""Power flux through the bounding surface (mirror of /root/reference/sucemfem/PostProcessing/power_flux.py:11-133).

SurfaceFlux integrates the Poynting vector over the exterior facets (host, O(boundary) work, basis and curl from the
generated nodal polynomials); VariationalSurfaceFlux evaluates the same quantity through the EM functional over the skin
of cells connected to the boundary (device matrices + SpMV, see PostProcessing.CalcEMFunctional).
"""
import numpy as np
import torch

from ..Consts import Z0
from .. import Geometry
from .. import dolfin_compat as dolfin
from .._generated_basis import BASIS
from . import CalcEMFunctional

# Dunavant degree-4 rule on the reference triangle (exact for the degree-3 integrand n.(E x curl E) at order 2)
_TRI_PTS = np.array([[0.445948490915965, 0.445948490915965], [0.445948490915965, 0.108103018168070],
                     [0.108103018168070, 0.445948490915965], [0.091576213509771, 0.091576213509771],
                     [0.091576213509771, 0.816847572980459], [0.816847572980459, 0.091576213509771]])
_TRI_WTS = np.array([0.223381589678011] * 3 + [0.109951743655322] * 3) * 0.5
_FACE_VERTS = np.array([[1, 2, 3], [0, 2, 3], [0, 1, 3], [0, 1, 2]])
_REF_VERTS = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1.]])


def _ref_basis_and_curl(order, X):
    """Reference basis values [npts, ndof, 3] and reference curls [npts, ndof, 3] at reference points X [npts, 3]."""
    exps, coef = np.asarray(BASIS[order][0]), np.asarray(BASIS[order][1])      # [nm,3], [ndof,3,nm]
    mono = np.prod(X[:, None, :] ** exps[None], axis=2)                          # [npts, nm]
    val = np.einsum('jdm,pm->pjd', coef, mono)
    dm = np.zeros((3,) + mono.shape)                                              # d mono / d X_a
    for a in range(3):
        e = exps.copy()
        fac = e[:, a].astype(float)
        e[:, a] = np.maximum(e[:, a] - 1, 0)
        dm[a] = fac[None] * np.prod(X[:, None, :] ** e[None], axis=2)
    grad = np.einsum('jdm,apm->pjda', coef, dm)                                  # d phi_d / d X_a
    curl = np.stack([grad[..., 2, 1] - grad[..., 1, 2], grad[..., 0, 2] - grad[..., 2, 0],
                     grad[..., 1, 0] - grad[..., 0, 1]], axis=-1)
    return val, curl


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
        V = self.function_space
        sp = V.device_space
        mesh = V.mesh()
        if mesh.geometry().dim() != 3:
            raise NotImplementedError('SurfaceFlux is provided for tetrahedral meshes')
        co, ce = mesh.coordinates(), mesh.cells()
        cell_faces = sp.cell_entities()[1].cpu().numpy()
        markers = torch.zeros(sp.n_facets, dtype=torch.int32, device=sp.mesh.device)
        sp.mark_exterior_facets(markers, 1)
        ext = markers.cpu().numpy()[cell_faces] == 1                              # [C, 4]
        cidx, lf = np.nonzero(ext)
        x = co[ce[cidx]]                                                          # [nf, 4, 3]
        J = np.swapaxes(x[:, 1:, :] - x[:, :1, :], 1, 2)
        detJ = np.linalg.det(J)
        Jinv = np.linalg.inv(J)
        cd = V.dofmap().cell_dofs_array()[cidx]
        xr, xi = self.dofs.real[cd], self.dofs.imag[cd]                           # [nf, ndof]
        mur = np.ones(len(cidx)) if self.mur_function is None else self.mur_function.vector()[cidx]
        total = 0.0
        for f in range(4):
            sel = lf == f
            if not sel.any():
                continue
            fv = _REF_VERTS[_FACE_VERTS[f]]
            X = fv[0][None] + _TRI_PTS[:, :1] * (fv[1] - fv[0])[None] + _TRI_PTS[:, 1:] * (fv[2] - fv[0])[None]
            val, curl = _ref_basis_and_curl(V.order, X)                           # [p, ndof, 3]
            Js, Ji, dJ = J[sel], Jinv[sel], detJ[sel]
            # physical fields at the quadrature points: E = J^-T sum x_j phihat_j ; curl E = J sum x_j curlhat_j / detJ
            def field(xx):
                Eh = np.einsum('nj,pjd->npd', xx[sel], val)
                Ch = np.einsum('nj,pjd->npd', xx[sel], curl)
                return np.einsum('nda,npd->npa', Ji, Eh), np.einsum('nad,npd->npa', Js, Ch) / dJ[:, None, None]
            Er, Cr = field(xr)
            Ei, Ci = field(xi)
            px = x[sel][:, _FACE_VERTS[f], :]
            nrm = np.cross(px[:, 1] - px[:, 0], px[:, 2] - px[:, 0])              # |nrm| = 2 area
            outward = np.einsum('nd,nd->n', nrm, px[:, 0] - x[sel][:, f, :])
            nrm = nrm * np.sign(outward)[:, None]
            integrand = np.einsum('nd,npd->np', nrm, np.cross(Er, -Ci) + np.cross(Ei, Cr)) / mur[sel][:, None]
            total += float(np.sum(integrand * _TRI_WTS[None, :]))
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
