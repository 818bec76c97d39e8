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


class DeviceEigenSolver(DefaultEigenSolver):
    """The same shift-invert eigen solve with the matrices left where the assembly kernels wrote them (HBM).

    The reference factorises S - sigma M with a sparse LU on the host and hands the solve to ARPACK
    (/root/reference/sucemfem/ProblemConfigurations/EMVectorWaveEigenproblem.py:76-103).  Here the shift-invert operator
    (S - sigma M)^-1 M is applied with a Jacobi-preconditioned MINRES whose matrix-vector products are sfem_matvec on the
    CSR value arrays (S - sigma M is never formed; one iteration of the recurrence is a CUDA graph, the residual estimate
    is read back every 25 iterations), inside a Lanczos iteration in the M inner product with full
    reorthogonalisation (sucemfem/Utilities/DeviceKrylov.py).  Dirichlet (PEC) dofs: the matrices carry row-only unit
    rows, so on vectors whose constrained entries are zero they act as the reduced pencil S_ff, M_ff.
    solve_problem() returns what DefaultEigenSolver returns: k^2 itself and the modes as rows; `self.residuals` holds
    ||S x - k^2 M x|| / ||k^2 M x|| of every returned pair, `self.info` the iteration counts."""

    def solve_problem(self, nev, ncv=None, tol=1e-11, inner_rtol=1e-13, inner_maxiter=50000, seed=0, return_device=False,
                      use_cuda_graph=True):
        import torch
        from ..Utilities import DeviceKrylov as K
        ep = self.eigenproblem
        M, S = ep.system_matrices['M'], ep.system_matrices['S']
        pat = M.pattern
        assert S.pattern is pat, 'S and M must share one pattern'
        n = pat.space.ndof
        dev = M.values.device
        flags = ep.boundary_conditions.get_essential_row_flags(ep.function_space)
        free = torch.ones(n, dtype=torch.float64, device=dev) if flags is None else (flags == 0).to(torch.float64)
        sigma = float(self.sigma)
        Mmul = lambda x: pat.matvec([M.values], x) * free
        Amul = lambda x: pat.matvec([S.values, M.values], x, [1.0, -sigma]) * free
        d = pat.diagonal([S.values, M.values], [1.0, -sigma]).real
        inv_diag = free / d.abs().clamp_min(1e-300)
        inner = {"solves": 0, "iterations": 0, "worst_residual": 0.0}
        # the MINRES recurrence is launch-bound: one iteration is captured in a CUDA graph and replayed
        graph = K.GraphMinres(Amul, inv_diag, n, dev) if use_cuda_graph else None

        def solve(w):
            x, its, res = K.solve_symmetric(Amul, w, inv_diag, inner_rtol, inner_maxiter, graph=graph)
            inner["solves"] += 1
            inner["iterations"] += its
            inner["worst_residual"] = max(inner["worst_residual"], res)
            return x * free

        g = torch.Generator(device=dev)
        g.manual_seed(seed)
        v0 = torch.rand(n, dtype=torch.float64, device=dev, generator=g) * free
        v0 = solve(Mmul(v0))             # one application of the operator: the start vector lies in its range
        nu, X, steps, est = K.lanczos_shift_invert(solve, Mmul, n, nev, v0, ncv=ncv, tol=tol)
        lam = sigma + 1.0 / nu
        order = torch.argsort(lam)
        lam, X = lam[order], X[order.to(X.device)]
        res = []
        for i in range(len(lam)):
            Mx = Mmul(X[i])
            r = pat.matvec([S.values], X[i]) * free - float(lam[i]) * Mx
            res.append(float(torch.linalg.vector_norm(r)) / (abs(float(lam[i])) * float(torch.linalg.vector_norm(Mx))))
        self.residuals = res
        self.info = dict(inner, lanczos_steps=steps)
        if return_device:
            return lam, X
        return lam.cpu().numpy(), X.cpu().numpy()
