"""Vector-wave eigenproblem S x = k^2 M x
(mirror of /root/reference/sucemfem/ProblemConfigurations/EMVectorWaveEigenproblem.py:38-105).

EigenProblem is the hot-path caller (forms {M, S, BC}).  DefaultEigenSolver is a consumer of the path ("next" row
f1 of SURVEY.md 8.2): it runs on the host with scipy, like the reference's, on the PEC-reduced pencil.  The
reference passes an already shift-inverted operator to ARPACK's shift-invert mode, which selects the modes with
k^2 > sigma closest to sigma and returns them shifted (its example scripts undo the shift by hand,
/root/reference/examples/resonant_cavity/driver.py:66); solve_problem() here returns k^2 itself, the convention of
/root/reference/functional_tests/test_EigenProblems.py:104."""
# Interface (module, class, method and argument names) after the file of cemagg/sucem-fem cited above:
# Copyright (C) 2011 Stellenbosch University, authors Neilen Marais; GNU GPL v3 or later
# (see NOTICE.md at the repository root).  The implementation below is this repository's own.
import numpy as np

from .. import Forms
from .EMProblem import EMProblem


class CombineForms(Forms.CombineGalerkinForms):
    def get_forms(self):
        m = self.interior_forms.get_mass_form()
        s = self.interior_forms.get_stiffness_form()
        BC_bilin_forms = self.boundary_conditions.get_bilinear_form()
        return dict(M=m, S=s, BC=BC_bilin_forms)


class EigenProblem(EMProblem):
    FormCombiner = CombineForms


class DefaultEigenSolver(object):
    def set_eigenproblem(self, eigenproblem):
        """Sets initialised instance of EigenProblem to solve"""
        self.eigenproblem = eigenproblem

    def set_sigma(self, sigma):
        """Spectrum shift (sigma) to apply to k^2"""
        self.sigma = sigma

    def solve_problem(self, nev, ncv=None):
        """Return (eigs_w, eigs_v): the nev eigenvalues k^2 > sigma closest to sigma and the eigenvectors as rows
        (n_eig, problem_dim), Dirichlet-constrained dofs zero."""
        import scipy.sparse.linalg as spla
        M = self.eigenproblem.system_matrices['M'].to_scipy()
        S = self.eigenproblem.system_matrices['S'].to_scipy()
        n = M.shape[0]
        ep = self.eigenproblem
        flags = ep.boundary_conditions.get_essential_row_flags(ep.function_space)
        free = np.arange(n) if flags is None else np.nonzero(flags.cpu().numpy() == 0)[0]
        Sf = S[free][:, free].tocsc()
        Mf = M[free][:, free].tocsc()
        lu = spla.splu((Sf - self.sigma * Mf).tocsc())
        # nu = 1/(k^2 - sigma): the physical modes above the shift are the largest positive nu; the curl-curl null
        # space (k^2 = 0) has nu < 0 and is never selected
        op = spla.LinearOperator(Sf.shape, matvec=lambda x: lu.solve(Mf @ x), dtype=np.float64)
        nu, v = spla.eigs(op, k=nev, which='LR', ncv=ncv)
        order = np.argsort(-nu.real)
        eigs_w = self.sigma + 1.0 / nu[order].real
        eigs_v = np.zeros((nev, n))
        eigs_v[:, free] = v[:, order].real.T
        return eigs_w, eigs_v
