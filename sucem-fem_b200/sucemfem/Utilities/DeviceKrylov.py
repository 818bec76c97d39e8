"""Krylov solvers on device vectors for the consumers of the assembled matrices that must not leave HBM
(SURVEY.md 8.2 row f1): MINRES for the symmetric indefinite shift-invert system S - sigma M and a Lanczos iteration in the
M inner product with full reorthogonalisation.  The operators are callables on torch tensors (sfem_matvec underneath:
DevicePattern.matvec); the vector algebra is torch's, the matrices never exist in any other form than the CSR value
arrays the assembly kernels wrote.

What the reference does instead: a sparse LU of S - sigma M handed to ARPACK
(/root/reference/sucemfem/ProblemConfigurations/EMVectorWaveEigenproblem.py:76-103) -- a host-side direct solve that
needs the matrices in host memory and does not exist beyond a few 1e5 unknowns."""
import math

import torch


def minres(A, b, inv_diag=None, rtol=1e-12, maxiter=20000, x0=None):
    """Solve A x = b for symmetric (indefinite) A with the minimum-residual recurrence of Paige and Saunders, Jacobi
    preconditioned by the POSITIVE vector inv_diag (or None).  Returns (x, iterations, estimated relative residual in the
    preconditioner's norm).  A: callable tensor -> tensor."""
    Minv = (lambda r: r * inv_diag) if inv_diag is not None else (lambda r: r)
    x = torch.zeros_like(b) if x0 is None else x0.clone()
    r1 = b - A(x) if x0 is not None else b.clone()
    y = Minv(r1)
    beta1 = float(torch.dot(r1, y))
    if beta1 <= 0.0:
        return x, 0, 0.0
    beta1 = math.sqrt(beta1)
    oldb, beta, dbar, epsln, phibar = 0.0, beta1, 0.0, 0.0, beta1
    cs, sn = -1.0, 0.0
    w = torch.zeros_like(b)
    w2 = torch.zeros_like(b)
    r2 = r1
    itn = 0
    for itn in range(1, maxiter + 1):
        v = y * (1.0 / beta)
        y = A(v)
        if itn >= 2:
            y = y - (beta / oldb) * r1
        alfa = float(torch.dot(v, y))
        y = y - (alfa / beta) * r2
        r1, r2 = r2, y
        y = Minv(r2)
        oldb = beta
        bb = float(torch.dot(r2, y))
        if bb < 0.0:
            raise ArithmeticError('MINRES: the preconditioner is not positive definite')
        beta = math.sqrt(bb)
        oldeps = epsln
        delta = cs * dbar + sn * alfa
        gbar = sn * dbar - cs * alfa
        epsln = sn * beta
        dbar = -cs * beta
        gamma = max(math.sqrt(gbar * gbar + beta * beta), 1e-300)
        cs, sn = gbar / gamma, beta / gamma
        phi = cs * phibar
        phibar = sn * phibar
        w1, w2 = w2, w
        w = (v - oldeps * w1 - delta * w2) * (1.0 / gamma)
        x = x + phi * w
        if phibar <= rtol * beta1 or beta == 0.0:
            break
    return x, itn, phibar / beta1


class GraphMinres(object):
    """The same recurrence with every scalar kept on the device and ONE iteration captured in a CUDA graph: the loop is
    launch-bound (a matrix-vector product and a dozen vector kernels per iteration, two scalar read-backs in the eager
    version), so replaying the graph and reading the residual estimate back only every `check_every` iterations removes
    the host from the loop.  A: callable on device tensors that only launches kernels on the current stream (sfem_matvec
    + torch ops); n: vector length.  One instance per (A, inv_diag): buffers and the graph are reused across solves."""

    def __init__(self, A, inv_diag, n, device, check_every=25):
        self.A, self.inv_diag, self.n, self.check_every = A, inv_diag, n, check_every
        z = lambda: torch.zeros(n, dtype=torch.float64, device=device)
        self.x, self.y, self.r1, self.r2, self.v, self.w, self.w1, self.w2 = z(), z(), z(), z(), z(), z(), z(), z()
        # scalars: beta, oldb, dbar, epsln, phibar, cs, sn
        self.sc = torch.zeros(7, dtype=torch.float64, device=device)
        self.graph = None

    def _step(self):
        beta, oldb, dbar, epsln, phibar, cs, sn = (self.sc[i] for i in range(7))
        self.v.copy_(self.y / beta)
        t = self.A(self.v)
        t = t - (beta / oldb) * self.r1                    # r1 = 0 in the first iteration
        alfa = torch.dot(self.v, t)
        t = t - (alfa / beta) * self.r2
        self.r1.copy_(self.r2)
        self.r2.copy_(t)
        self.y.copy_(self.r2 * self.inv_diag)
        nb = torch.sqrt(torch.dot(self.r2, self.y).clamp_min(0.0))
        oldeps = epsln.clone()
        delta = cs * dbar + sn * alfa
        gbar = sn * dbar - cs * alfa
        n_eps = sn * nb
        n_dbar = -cs * nb
        gamma = torch.sqrt(gbar * gbar + nb * nb).clamp_min(1e-300)
        n_cs, n_sn = gbar / gamma, nb / gamma
        phi = n_cs * phibar
        n_phibar = n_sn * phibar
        self.w1.copy_(self.w2)
        self.w2.copy_(self.w)
        self.w.copy_((self.v - oldeps * self.w1 - delta * self.w2) / gamma)
        self.x.add_(phi * self.w)
        self.sc.copy_(torch.stack([nb, beta.clone(), n_dbar, n_eps, n_phibar, n_cs, n_sn]))

    def solve(self, b, rtol=1e-12, maxiter=20000):
        for t in (self.x, self.r1, self.w, self.w1, self.w2):
            t.zero_()
        self.r2.copy_(b)
        self.y.copy_(b * self.inv_diag)
        beta1 = float(torch.sqrt(torch.dot(self.r2, self.y)))
        if beta1 == 0.0:
            return self.x.clone(), 0, 0.0
        self.sc.copy_(torch.tensor([beta1, 1.0, 0.0, 0.0, beta1, -1.0, 0.0], dtype=torch.float64))
        if self.graph is None:
            state = [t.clone() for t in (self.x, self.y, self.r1, self.r2, self.v, self.w, self.w1, self.w2, self.sc)]
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):               # warm-up outside the capture
                self._step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._step()
            for t, s0 in zip((self.x, self.y, self.r1, self.r2, self.v, self.w, self.w1, self.w2, self.sc), state):
                t.copy_(s0)
        itn = 0
        while itn < maxiter:
            for _ in range(self.check_every):
                self.graph.replay()
            itn += self.check_every
            phibar, beta = float(self.sc[4]), float(self.sc[0])
            if not math.isfinite(phibar):
                raise ArithmeticError('MINRES: breakdown')
            if phibar <= rtol * beta1 or beta == 0.0:
                break
        return self.x.clone(), itn, float(self.sc[4]) / beta1


def solve_symmetric(A, b, inv_diag=None, rtol=1e-12, maxiter=20000, refinements=3, graph=None):
    """MINRES with iterative refinement on the TRUE residual: returns (x, total iterations, ||b - A x|| / ||b||).
    graph: a GraphMinres for (A, inv_diag) -- the CUDA-graph version of the same recurrence (device tensors only)."""
    bn = float(torch.linalg.vector_norm(b))
    if bn == 0.0:
        return torch.zeros_like(b), 0, 0.0
    run = (lambda rhs, tol: graph.solve(rhs, tol, maxiter)) if graph is not None else \
          (lambda rhs, tol: minres(A, rhs, inv_diag, tol, maxiter))
    x, its, _ = run(b, rtol)
    res = float(torch.linalg.vector_norm(b - A(x))) / bn
    for _ in range(refinements):
        if res <= rtol:
            break
        r = b - A(x)
        dx, it2, _ = run(r, max(rtol / max(res, 1e-300), 1e-6) * 0.1)
        x = x + dx
        its += it2
        res = float(torch.linalg.vector_norm(b - A(x))) / bn
    return x, its, res


def lanczos_shift_invert(solve, Mmul, n, nev, v0, ncv=None, tol=1e-12, max_steps=None):
    """Largest positive eigenvalues nu of OP = solve(M .) -- self-adjoint in the M inner product -- by Lanczos with full
    reorthogonalisation.  solve(w) = (S - sigma M)^-1 w, Mmul(x) = M x.  Returns (nu [k], X [k, n] M-orthonormal Ritz
    vectors, steps, residual estimates); k = nev if that many converged within max_steps, else fewer.  The basis grows
    until the nev wanted Ritz pairs satisfy |beta_m s_mi| <= tol |nu_i|."""
    ncv = max(2 * nev + 1, 20) if ncv is None else ncv
    max_steps = max(10 * nev + 60, ncv) if max_steps is None else max_steps
    V, W = [], []                       # v_j and M v_j
    alphas, betas = [], []
    v = v0.clone()
    w = Mmul(v)
    nrm = math.sqrt(float(torch.dot(v, w)))
    v, w = v / nrm, w / nrm
    nu = s = None
    est = None
    for j in range(max_steps):
        V.append(v)
        W.append(w)
        z = solve(w)
        alpha = float(torch.dot(z, w))
        alphas.append(alpha)
        # full reorthogonalisation against the whole basis in the M inner product (twice is enough)
        Vm, Wm = torch.stack(V), torch.stack(W)
        for _ in range(2):
            z = z - (Wm @ z) @ Vm
        wz = Mmul(z)
        b2 = float(torch.dot(z, wz))
        beta = math.sqrt(max(b2, 0.0))
        m = j + 1
        if m >= min(ncv, max_steps) or beta == 0.0:
            T = torch.diag(torch.tensor(alphas, dtype=torch.float64))
            if m > 1:
                off = torch.tensor(betas, dtype=torch.float64)
                T = T + torch.diag(off, 1) + torch.diag(off, -1)
            nu, s = torch.linalg.eigh(T)
            pos = torch.nonzero(nu > 0).flatten()
            pick = pos[torch.argsort(nu[pos], descending=True)][:nev]
            est = (beta * s[m - 1, pick].abs() / nu[pick].abs()) if len(pick) else torch.tensor([])
            if beta == 0.0 or (len(pick) == nev and float(est.max()) <= tol):
                break
        betas.append(beta)
        v, w = z / beta, wz / beta
    if nu is None:
        raise ArithmeticError('Lanczos: no steps taken')
    m = len(alphas)
    pos = torch.nonzero(nu > 0).flatten()
    pick = pos[torch.argsort(nu[pos], descending=True)][:nev]
    Vm = torch.stack(V[:m])
    X = s[:, pick].T.to(Vm.device) @ Vm
    return nu[pick], X, m, est
