"""Variational near-field to far-field transform (mirror of /root/reference/sucemfem/PostProcessing/variational_ntff.py:30-131):
the H-field contribution comes from the EM functional g^T (S - k0^2 M) E over the skin of boundary cells (device
matrices + fused SpMV, see CalcEMFunctional) with g the boundary interpolant of 1j/k0 rhat x (ahat x rhat) e^{j k0 rhat.r'};
the E-field contribution (L) from the surface integral."""
# Interface (module, class, method and argument names) after the file of cemagg/sucem-fem cited above:
# Copyright (C) 2011 Stellenbosch University, authors Neilen Marais; GNU GPL v3 or later
# (see NOTICE.md at the repository root).  The implementation below is this repository's own.
import numpy as np

from ..Consts import c0
from .. import Geometry
from .. import dolfin_compat as dolfin
from ..Interpolation import SurfaceInterpolant
from . import CalcEMFunctional
from .surface_ntff import SurfaceNTFFForms


def _spherical_unit_vectors(theta_deg, phi_deg):
    """(rhat, theta_hat, phi_hat) of the spherical coordinate system at (theta, phi) given in degrees."""
    th, ph = np.deg2rad(theta_deg), np.deg2rad(phi_deg)
    st, ct, sp, cp = np.sin(th), np.cos(th), np.sin(ph), np.cos(ph)
    return (np.array([st * cp, st * sp, ct], dtype=np.float64), np.array([ct * cp, ct * sp, -st], dtype=np.float64),
            np.array([-sp, cp, 0.0], dtype=np.float64))


class _Expr(object):
    """Vector expression cf * trig(k0 rhat . x); callable on point arrays [n,3], and eval(value, x) like dolfin's."""

    def __init__(self, sign, trig):
        self.sign, self.trig = sign, trig
        self.cf = np.zeros(3)
        self.k0 = 0.0
        self.rhat = np.zeros(3)

    def __call__(self, pts):
        pts = np.atleast_2d(np.asarray(pts, dtype=np.float64))
        return self.sign * self.trig(self.k0 * pts @ self.rhat)[:, None] * self.cf[None, :]

    def eval(self, value, x):
        value[:] = self(x)[0]


class TransformTestingExpression(object):
    """1j/k0 rhat x (ahat x rhat) e^{j k0 rhat . r'}: real part -cf sin(.), imaginary part cf cos(.)"""

    def __init__(self):
        self.expr_r = _Expr(-1.0, np.sin)
        self.expr_i = _Expr(1.0, np.cos)

    def set_parms(self, rhat, ahat, k0):
        crossy = np.cross(rhat, np.cross(ahat, rhat))
        constfac = crossy / k0
        for expr in (self.expr_r, self.expr_i):
            expr.cf = np.asarray(constfac, dtype=np.float64)
            expr.k0 = k0
            expr.rhat = np.asarray(rhat, dtype=np.float64)

    def get_expression(self):
        return self.expr_r, self.expr_i


class NTFF(object):
    def __init__(self, function_space, testing_space=None):
        self.function_space = function_space
        if testing_space is None:
            testing_space = function_space
        self.testing_space = testing_space
        self.functional = CalcEMFunctional(function_space, testing_space)
        self.testing_interpolator = SurfaceInterpolant(testing_space)
        self.cell_domains = dolfin.CellFunction('uint', self.function_space.mesh())
        self.cell_domains.set_all(0)
        self.cell_region = 1
        boundary_cells = Geometry.BoundaryEdgeCells(self.function_space.mesh())
        boundary_cells.mark(self.cell_domains, self.cell_region)
        self.surface_forms = SurfaceNTFFForms(self.function_space)
        self.testing_expression_gen = TransformTestingExpression()
        self.functional.set_cell_domains(self.cell_domains, self.cell_region)

    def set_k0(self, k0):
        self.k0 = k0
        self.functional.set_k0(k0)

    def set_frequency(self, frequency):
        self.frequency = frequency
        self.set_k0(self.frequency * 2 * np.pi / c0)

    def set_dofs(self, dofs):
        self.functional.set_E_dofs(dofs)
        self.surface_forms.set_dofs(dofs)

    def calc_pt(self, theta_deg, phi_deg):
        """(E_theta, E_phi) of the far field in direction (theta, phi), normalised to radius 1: the H-field part from
        the variational functional, the E-field part (L) from the surface integral."""
        from_H = self.calc_pt_E_H(theta_deg, phi_deg)
        self.surface_forms.set_parms(np.deg2rad(theta_deg), np.deg2rad(phi_deg), self.k0)
        L_theta, L_phi = self.surface_forms.assemble_L()
        scale = 1j * self.k0 * np.exp(-1j * self.k0) / (4 * np.pi)
        return (scale * (from_H[0] - L_phi), scale * (from_H[1] + L_theta))

    def calc_pt_E_H(self, theta_deg, phi_deg):
        rhat, theta_hat, phi_hat = _spherical_unit_vectors(theta_deg, phi_deg)
        return self.calc_ff_func(rhat, theta_hat), self.calc_ff_func(rhat, phi_hat)

    def calc_ff_func(self, rhat, ahat):
        """Functional with the testing function interpolated from the expression for direction rhat, polarisation ahat."""
        self.testing_expression_gen.set_parms(rhat, ahat, self.k0)
        self.testing_interpolator.set_interpolant_expression(*self.testing_expression_gen.get_expression())
        self.functional.set_g_dofs(self.testing_interpolator.calculate_interpolation())
        return self.functional.calc_functional()
