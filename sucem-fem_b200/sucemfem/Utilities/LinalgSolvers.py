"""A collection of linear system solvers (mirror of /root/reference/sucemfem/Utilities/LinalgSolvers.py:29-322).

Same classes, methods and logging as the reference.  The system matrix may be
  * a sucemfem.dolfin_compat.DeviceLinearOperator (matrices resident in HBM, e.g. DrivenProblemABC.get_LHS_operator()):
    BiCGStabSolver then runs on the device (sfem_bicgstab: fused three-matrix SpMV over the shared pattern, no
    preconditioner or the reference's 'diagonal' one); or
  * a scipy sparse matrix (what get_LHS_matrix() returns): the solvers call scipy exactly as the reference does.
"""
# Interface (module, class, method and argument names) after the file of cemagg/sucem-fem cited above:
# Copyright (C) 2011 Stellenbosch University, authors Evan Lezar, Neilen Marais; GNU GPL v3 or later
# (see NOTICE.md at the repository root).  The implementation below is this repository's own.
from time import time

import numpy as np
import scipy.sparse
import scipy.sparse.linalg

from ..dolfin_compat import DeviceLinearOperator


class SystemSolverBase(object):
    """Base class with the reference's logging (LinalgSolvers.py:29-236); subclasses implement _call_solver."""

    def __init__(self, A, preconditioner_type=None):
        self._logging_data = {'time': [], 'id': [], 'res': []}
        self._user_callbacks = []
        self._timestamp('init')
        self._A = A
        self._on_device = isinstance(A, DeviceLinearOperator)
        self.set_preconditioner(preconditioner_type)
        self._callback_count = 0

    def _call_solver(self):
        raise Exception("Solver driver not implemented.")

    def add_user_callback(self, cb):
        """Add a user callback function cb(self, residual)"""
        self._user_callbacks.append(cb)

    def _callback(self, xk):
        self._callback_count += 1
        if np.ndim(xk) == 0:      # GMRES hands its callback the (preconditioned, relative) residual norm, not x_k
            res = float(xk)
        else:
            res = calculate_residual(self._A, xk, self._b)
        self._timestamp(self._callback_count, res=res)

    def _timestamp(self, id, res=np.nan):
        self._logging_data['time'].append(time())
        self._logging_data['id'].append(id)
        self._logging_data['res'].append(res)
        for cb in self._user_callbacks:
            cb(self, res)

    def set_b(self, b):
        self._b = b

    def set_preconditioner(self, M_type):
        self._timestamp('preconditioner::start')
        self._M_type = None if M_type is None else M_type.lower()
        if self._on_device:
            if self._M_type not in (None, 'diagonal'):
                raise NotImplementedError("device solver: preconditioner %r (None and 'diagonal' are available)" % M_type)
            M = None          # applied inside sfem_bicgstab
        elif M_type is None:
            M = None
        elif self._M_type == 'diagonal':
            M = scipy.sparse.spdiags(1. / self._A.diagonal(), 0, self._A.shape[0], self._A.shape[1])
        elif self._M_type == 'ilu':
            self._M_data = scipy.sparse.linalg.spilu(self._A.tocsc(), drop_tol=1e-8, fill_factor=10)
            M = scipy.sparse.linalg.LinearOperator(self._A.shape, self._M_data.solve)
        else:
            print("Warning: Preconditioner type '%s' not recognised. Not using preconditioner." % M_type)
            M = None
        self._M = M
        self._timestamp('preconditioner::end')

    def solve(self, b):
        """Solve the linear system Ax = b for x"""
        self._timestamp('solve::begin')
        self.set_b(b)
        x, info = self._call_solver()
        self.info = info
        self._timestamp('solve::end')
        return x

    def get_logging_data(self):
        return self._logging_data

    def _get_time_for_id(self, id):
        return self._logging_data['time'][self._logging_data['id'].index(id)]

    def _get_elapsed_time(self, id0, id1, force_total=False):
        if force_total:
            return self._logging_data['time'][-1] - self._logging_data['time'][0]
        return self._get_time_for_id(id1) - self._get_time_for_id(id0)

    def print_logging_data(self):
        for k in self._logging_data:
            print(k)
            print(self._logging_data[k])

    def plot_convergence(self, x_is_time=False, show_plot=False, label=None, style='-'):
        """Residual history against the iteration number or the elapsed time (LinalgSolvers.py:192-236); returns the
        (x, residual) arrays it plots, so it is usable without a display."""
        ids, res, t = self._logging_data['id'], self._logging_data['res'], self._logging_data['time']
        keep = [i for i, k in enumerate(ids) if isinstance(k, (int, np.integer))]
        x = np.array([t[i] - t[0] for i in keep]) if x_is_time else np.array([ids[i] for i in keep])
        y = np.array([res[i] for i in keep])
        try:
            import matplotlib.pyplot as plt
            plt.semilogy(x, y, style, label=label)
            plt.xlabel('time (s)' if x_is_time else 'iteration')
            plt.ylabel('residual')
            if label:
                plt.legend()
            if show_plot:
                plt.show()
        except ImportError:
            pass
        return x, y

    def print_timing_info(self):
        print('solve time:', self._get_elapsed_time('solve::begin', 'solve::end'))
        print('preconditioner time:', self._get_elapsed_time('preconditioner::start', 'preconditioner::end'))
        print('total time:', self._get_elapsed_time(None, None, True))


class BiCGStabSolver(SystemSolverBase):
    """Stabilised BiCG: on the device for a DeviceLinearOperator, scipy otherwise (LinalgSolvers.py:239-249)."""
    rtol = 1e-5          # scipy's default tolerance, which the reference relies on
    maxiter = None

    def _call_solver(self):
        if self._on_device:
            A = self._A
            x, info, its, relres = A.pattern.bicgstab(A.values, A.coefficients, self._b, preconditioner=self._M_type,
                                                      rtol=self.rtol, maxiter=self.maxiter)
            self.iterations = its
            self._timestamp(its, res=relres * float(np.linalg.norm(np.asarray(self._b))))
            return x.cpu().numpy(), info
        return scipy.sparse.linalg.bicgstab(self._A, self._b, M=self._M, callback=self._callback, rtol=self.rtol,
                                            maxiter=self.maxiter)


class GMRESSolver(SystemSolverBase):
    """GMRES through scipy (LinalgSolvers.py:252-262); host matrices only."""

    def _call_solver(self):
        A = self._A.to_scipy() if self._on_device else self._A
        return scipy.sparse.linalg.gmres(A, self._b, M=self._M, callback=self._callback, callback_type='pr_norm')


class UMFPACKSolver(SystemSolverBase):
    """Direct sparse solve through scipy (LinalgSolvers.py:264-279); host matrices only."""

    def _call_solver(self):
        A = self._A.to_scipy() if self._on_device else self._A
        return scipy.sparse.linalg.spsolve(scipy.sparse.csc_matrix(A), self._b), 0


def calculate_residual(A, x, b):
    """norm(A x - b) (LinalgSolvers.py:292-300)"""
    return np.linalg.norm(A * x - b.reshape(x.shape))


def solve_sparse_system(A, b, preconditioner_type='ilu'):
    """BiCGStab solve of A x = b (LinalgSolvers.py:303-322); A: scipy sparse matrix or DeviceLinearOperator (for which
    the reference default 'ilu' is replaced by 'diagonal', the preconditioner its drivers pass)."""
    if isinstance(A, DeviceLinearOperator) and preconditioner_type == 'ilu':
        preconditioner_type = 'diagonal'
    solver = BiCGStabSolver(A, preconditioner_type)
    return solver.solve(b)
