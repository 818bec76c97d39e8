"""Near-field to far-field transform by surface integrals of the equivalent currents over the bounding surface
(mirror of /root/reference/sucemfem/PostProcessing/surface_ntff.py:29-155).

N = int n x H e^{j k0 r'.rhat} ds,  L = -int n x E e^{j k0 r'.rhat} ds,  H = j curl E / (k0 Z0).
Host evaluation over the exterior facets (O(boundary) work) with a collapsed Gauss-Jacobi rule; the integrand is not
polynomial, the default rule (6 x 6 points) reproduces the reference's shipped far fields to 2e-8.
"""
# Interface (module, class, method and argument names) after the file of cemagg/sucem-fem cited above:
# Copyright (C) 2011 Stellenbosch University, authors Neilen Marais; GNU GPL v3 or later
# (see NOTICE.md at the repository root).  The implementation below is this repository's own.
import numpy as np

from ..Consts import Z0, c0
from . import ntff_expressions
from ._facets import ExteriorFacets, tri_rule


class SurfaceNTFFForms(object):
    quadrature_points_1d = 6

    def __init__(self, function_space):
        self.function_space = function_space
        self._facets = None
        self.dofs = None

    def set_dofs(self, dofs):
        self.dofs = np.asarray(dofs, dtype=np.complex128)

    def set_parms(self, theta, phi, k0):
        self.k0 = k0
        self.r_hat = ntff_expressions.r_hat(theta, phi)
        self.theta_hat = ntff_expressions.theta_hat(theta, phi)
        self.phi_hat = ntff_expressions.phi_hat(theta, phi)

    def _integrate(self, current):
        """int current(n, E, curlE) e^{j phase} ds over the exterior facets -> complex 3-vector"""
        if self._facets is None:
            self._facets = ExteriorFacets(self.function_space)
        fac = self._facets
        pts, wts = tri_rule(self.quadrature_points_1d)
        total = np.zeros(3, np.complex128)
        for sel, X, nrm, pos in fac.groups(pts):
            E, C = fac.fields(sel, X, self.dofs)
            ej = np.exp(1j * ntff_expressions.phase(self.k0, pos, self.r_hat))
            total += np.einsum('npd,np,p->d', current(nrm[:, None, :], E, C), ej, wts)    # |nrm| = 2 area, sum(wts) = 1/2
        return total

    def assemble_N(self):
        N = self._integrate(lambda n, E, C: np.cross(n, 1j * C / (self.k0 * Z0)))
        return (np.dot(self.theta_hat, N), np.dot(self.phi_hat, N))

    def assemble_L(self):
        L = self._integrate(lambda n, E, C: -np.cross(n, E))
        return (np.dot(self.theta_hat, L), np.dot(self.phi_hat, L))


class NTFF(object):
    def __init__(self, function_space):
        self.function_space = function_space
        self.forms = SurfaceNTFFForms(function_space)
        self._L = []
        self._N = []
        self._r_fac = []

    def set_dofs(self, dofs):
        self.forms.set_dofs(dofs)

    def set_frequency(self, frequency):
        self.frequency = frequency
        self.k0 = self.frequency * 2 * np.pi / c0

    def calc_pt(self, theta_deg, phi_deg):
        theta = np.deg2rad(theta_deg)
        phi = np.deg2rad(phi_deg)
        self.forms.set_parms(theta, phi, self.k0)
        L_theta, L_phi = self.forms.assemble_L()
        N_theta, N_phi = self.forms.assemble_N()
        # far fields normalised to radius 1
        r_fac = 1j * self.k0 * np.exp(-1j * self.k0) / (4 * np.pi)
        E_theta = -r_fac * (L_phi + Z0 * N_theta)
        E_phi = r_fac * (L_theta - Z0 * N_phi)
        self._L.append([L_theta, L_phi])
        self._N.append([N_theta, N_phi])
        self._r_fac.append(r_fac)
        return (E_theta, E_phi)
