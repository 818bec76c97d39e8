"""Driven problem terminated by a first-order ABC
(mirror of /root/reference/sucemfem/ProblemConfigurations/EMDrivenProblem.py:33-91):
A = S - k0**2*M + 1j*k0*S_0."""
# Interface (module, class, method and argument names) after the file of cemagg/sucem-fem cited above:
# Copyright (C) 2011 Stellenbosch University, authors Neilen Marais; GNU GPL v3 or later
# (see NOTICE.md at the repository root).  The implementation below is this repository's own.
import numpy as N

from .. import Forms
from .. import dolfin_compat as dolfin
from ..Consts import c0, Z0
from ..Utilities.Converters import dolfin_ublassparse_to_scipy_csr
from .EMProblem import EMProblem


class CombineForms(Forms.CombineGalerkinForms):
    def get_forms(self):
        m = self.interior_forms.get_mass_form()
        s = self.interior_forms.get_stiffness_form()
        ABC_form = self.boundary_conditions.get_bilinear_form()
        return dict(M=m, S=s, S_0=ABC_form)


class DrivenProblemABC(EMProblem):
    FormCombiner = CombineForms
    # The reference hands its matrices to scipy (uBLASSparseMatrix = host CSR, EMDrivenProblem.py:82-87).  Setting
    # this to dolfin.PETScMatrix (device CSR) keeps S, M, S_0 in HBM for get_LHS_operator() and the device solver.
    matrix_class = dolfin.uBLASSparseMatrix

    def set_matrix_class(self, matrix_class):
        self.matrix_class = matrix_class

    def set_sources(self, sources):
        self.sources = sources

    def set_frequency(self, frequency):
        """Set simulation frequency in Hz"""
        self.frequency = frequency

    def get_LHS_matrix(self):
        k0 = 2 * N.pi * self.frequency / c0
        M = dolfin_ublassparse_to_scipy_csr(self.system_matrices['M'])
        S = dolfin_ublassparse_to_scipy_csr(self.system_matrices['S'])
        S_0 = dolfin_ublassparse_to_scipy_csr(self.system_matrices['S_0'])
        return S - k0 ** 2 * M + 1j * k0 * S_0

    def get_LHS_operator(self):
        """The same matrix as a device operator (never formed): needs device-resident matrices
        (set_matrix_class(dolfin.PETScMatrix) before init_problem)."""
        k0 = 2 * N.pi * self.frequency / c0
        sm = self.system_matrices
        if sm['S_0'] is None:
            return dolfin.DeviceLinearOperator([sm['S'], sm['M']], [1.0, -k0 ** 2])
        return dolfin.DeviceLinearOperator([sm['S'], sm['M'], sm['S_0']], [1.0, -k0 ** 2, 1j * k0])

    def get_RHS(self):
        RHS = N.zeros(self.get_global_dimension(), N.complex128)
        dofnos, contribs = self._get_RHS_contributions()
        k0 = 2 * N.pi * self.frequency / c0
        contribs = -1j * k0 * Z0 * contribs
        RHS[dofnos] += contribs
        return RHS

    def _init_system_matrices(self):
        EMProblem._init_system_matrices(self, matrix_class=self.matrix_class)

    def _get_RHS_contributions(self):
        self.sources.set_function_space(self.function_space)
        self.sources.init_sources()
        return self.sources.get_source_contributions()
