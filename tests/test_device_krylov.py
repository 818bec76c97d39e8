"""The Krylov layer of the device eigen solver (sucemfem/Utilities/DeviceKrylov.py) on CPU tensors: MINRES and the
shift-invert Lanczos iteration, with the oracle's matrices standing in for the device CSR arrays (the GPU test
tests/test_eigen_gpu.py runs the same code on sfem_matvec).  Criterion of SURVEY 8.2 / BASELINE north_star: cavity
eigenvalues to 1e-10 against a dense generalised eigen solve of the PEC-reduced pencil."""
import os
import sys

import numpy as np
import pytest
import scipy.linalg
import scipy.sparse as sps
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'sucem-fem_b200'))

from oracle import fem, meshes


def _load_krylov():
    import importlib.util
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'sucem-fem_b200', 'sucemfem',
                        'Utilities', 'DeviceKrylov.py')
    spec = importlib.util.spec_from_file_location('DeviceKrylov', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


K = _load_krylov()


def _pencil(n, order):
    co, ce = meshes.unit_cube(n)
    num = fem.number_entities(ce)
    pec = fem.exterior_facet_mask(ce, num)
    sysm = fem.assemble_system(co, ce, order, pec_facet_mask=pec)    # row-only ident applied, as the product does
    ndof = len(sysm['rowptr']) - 1
    S = sps.csr_matrix((sysm['S'], sysm['cols'], sysm['rowptr']), shape=(ndof, ndof))
    M = sps.csr_matrix((sysm['M'], sysm['cols'], sysm['rowptr']), shape=(ndof, ndof))
    rows = fem.facet_closure_dofs(ce, num, order, pec)
    free = np.ones(ndof)
    free[rows] = 0.0
    return S, M, free


def test_minres_indefinite_system():
    S, M, free = _pencil(3, 2)
    sigma = 12.0
    A = (S - sigma * M).tocsr()
    f = torch.from_numpy(free)
    Amul = lambda x: torch.from_numpy(A @ x.numpy()) * f
    rng = np.random.default_rng(0)
    b = torch.from_numpy(rng.standard_normal(len(free))) * f
    inv_diag = f / torch.from_numpy(np.abs(A.diagonal()))
    x, its, res = K.solve_symmetric(Amul, b, inv_diag, rtol=1e-12, maxiter=20000)
    assert res <= 1e-12, res
    idx = np.nonzero(free)[0]
    ref = np.linalg.solve(A[idx][:, idx].toarray(), b.numpy()[idx])
    assert np.abs(x.numpy()[idx] - ref).max() <= 1e-9 * np.abs(ref).max()


@pytest.mark.parametrize('n,order', [(4, 1), (3, 2)])
def test_shift_invert_lanczos_vs_dense(n, order):
    S, M, free = _pencil(n, order)
    idx = np.nonzero(free)[0]
    w = scipy.linalg.eigh(S[idx][:, idx].toarray(), M[idx][:, idx].toarray(), eigvals_only=True)
    sigma = 10.0
    want = np.sort(w[w > sigma])[:10]
    f = torch.from_numpy(free)
    A = (S - sigma * M).tocsr()
    Mmul = lambda x: torch.from_numpy(M @ x.numpy()) * f
    Amul = lambda x: torch.from_numpy(A @ x.numpy()) * f
    inv_diag = f / torch.from_numpy(np.abs(A.diagonal()))
    solve = lambda b: K.solve_symmetric(Amul, b, inv_diag, rtol=1e-13, maxiter=20000)[0] * f
    g = torch.Generator().manual_seed(0)
    v0 = solve(Mmul(torch.rand(len(free), dtype=torch.float64, generator=g) * f))
    nu, X, steps, est = K.lanczos_shift_invert(solve, Mmul, len(free), 10, v0, tol=1e-12)
    lam = np.sort(sigma + 1.0 / nu.numpy())
    assert len(lam) == 10
    assert np.abs(lam - want).max() <= 1e-10 * np.abs(want).max(), (lam, want)
    # Ritz vectors: M-orthonormal, residuals small, constrained entries zero
    Xn = X.numpy()
    G = Xn @ (M @ Xn.T)
    assert np.abs(G - np.eye(10)).max() <= 1e-9
    assert np.abs(Xn[:, free == 0]).max() == 0.0
