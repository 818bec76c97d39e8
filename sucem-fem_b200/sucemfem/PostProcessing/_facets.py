"""Host helpers shared by the surface functionals (power flux, NTFF): exterior-facet geometry from the device
numbering, reference basis / curl evaluation from the generated nodal polynomials, triangle quadrature.  O(boundary)
work: outside the assembly hot path."""
import numpy as np
import torch

from .._generated_basis import BASIS

FACE_VERTS = np.array([[1, 2, 3], [0, 2, 3], [0, 1, 3], [0, 1, 2]])
REF_VERTS = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1.]])


def tri_rule(m):
    """Collapsed Gauss-Jacobi rule with m x m points on the reference triangle (exact to degree 2m-1); weights sum
    to 1/2."""
    from numpy.polynomial.legendre import leggauss
    from scipy.special import roots_jacobi
    xa, wa = roots_jacobi(m, 1, 0)
    xb, wb = leggauss(m)
    u = (1 + xa) / 2
    v = (1 + xb) / 2
    pts = np.array([[u[i], v[j] * (1 - u[i])] for i in range(m) for j in range(m)])
    wts = np.array([wa[i] * wb[j] / 8 for i in range(m) for j in range(m)])
    return pts, wts


def ref_basis_and_curl(order, X):
    """Reference basis values and reference curls, both [npts, ndof, 3], at reference points X [npts, 3]."""
    exps, coef = np.asarray(BASIS[order][0]), np.asarray(BASIS[order][1])      # [nm,3], [ndof,3,nm]
    mono = np.prod(X[:, None, :] ** exps[None], axis=2)
    val = np.einsum('jdm,pm->pjd', coef, mono)
    dm = np.zeros((3,) + mono.shape)
    for a in range(3):
        e = exps.copy()
        fac = e[:, a].astype(float)
        e[:, a] = np.maximum(e[:, a] - 1, 0)
        dm[a] = fac[None] * np.prod(X[:, None, :] ** e[None], axis=2)
    grad = np.einsum('jdm,apm->pjda', coef, dm)                                  # d phi_d / d X_a
    curl = np.stack([grad[..., 2, 1] - grad[..., 1, 2], grad[..., 0, 2] - grad[..., 2, 0],
                     grad[..., 1, 0] - grad[..., 0, 1]], axis=-1)
    return val, curl


class ExteriorFacets(object):
    """(cell, local face) pairs of the exterior facets of a tetrahedral function space and their geometry."""

    def __init__(self, V):
        sp = V.device_space
        mesh = V.mesh()
        if mesh.geometry().dim() != 3:
            raise NotImplementedError('surface functionals are provided for tetrahedral meshes')
        co, ce = mesh.coordinates(), mesh.cells()
        cell_faces = sp.cell_entities()[1].cpu().numpy()
        markers = torch.zeros(sp.n_facets, dtype=torch.int32, device=sp.mesh.device)
        sp.mark_exterior_facets(markers, 1)
        ext = markers.cpu().numpy()[cell_faces] == 1
        self.cell, self.lf = np.nonzero(ext)
        self.x = co[ce[self.cell]]                                                # [nf, 4, 3]
        self.J = np.swapaxes(self.x[:, 1:, :] - self.x[:, :1, :], 1, 2)
        self.detJ = np.linalg.det(self.J)
        self.Jinv = np.linalg.inv(self.J)
        self.cell_dofs = V.dofmap().cell_dofs_array()[self.cell]
        self.order = V.order

    def groups(self, pts2d):
        """For each local face f: (selection mask, reference points on that face [p,3], outward normals scaled by
        2*area [n,3], physical points [n,p,3])."""
        for f in range(4):
            sel = self.lf == f
            if not sel.any():
                continue
            fv = REF_VERTS[FACE_VERTS[f]]
            X = fv[0][None] + pts2d[:, :1] * (fv[1] - fv[0])[None] + pts2d[:, 1:] * (fv[2] - fv[0])[None]
            xs = self.x[sel]
            px = xs[:, FACE_VERTS[f], :]
            nrm = np.cross(px[:, 1] - px[:, 0], px[:, 2] - px[:, 0])
            outward = np.einsum('nd,nd->n', nrm, px[:, 0] - xs[:, f, :])
            nrm = nrm * np.sign(outward)[:, None]
            pos = xs[:, 0, None, :] + np.einsum('nda,pa->npd', self.J[sel], X)
            yield sel, X, nrm, pos

    def fields(self, sel, X, dofs):
        """E [n,p,3] and curl E [n,p,3] of the function with (complex) dof vector `dofs` at the points X."""
        val, curl = ref_basis_and_curl(self.order, X)
        xe = dofs[self.cell_dofs[sel]]
        Eh = np.einsum('nj,pjd->npd', xe, val)
        Ch = np.einsum('nj,pjd->npd', xe, curl)
        E = np.einsum('nda,npd->npa', self.Jinv[sel], Eh)                        # E = J^-T Ehat
        C = np.einsum('nad,npd->npa', self.J[sel], Ch) / self.detJ[sel][:, None, None]
        return E, C
