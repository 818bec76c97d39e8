"""Parity at BASELINE full size (config 3: UnitCube(70), order 2, inhomogeneous eps_r; 2.06 M tets, 571 M non-zeros per
matrix) through size-independent properties -- the oracle cannot assemble this in seconds:

  * curl-curl annihilates gradients: S x_c = 0 for the interpolant x_c of any constant field c (edge dofs c.(x_b-x_a),
    face dofs c.t_0, c.t_1), up to round-off, on the un-constrained matrix;
  * mass is exact on constants: x_c^T M x_c = |c|^2 sum_cells eps_r vol;
  * symmetry: y^T (S x) = x^T (S y), same for M, for random x, y;
  * PEC rows are unit rows; all other rows are untouched by the fused ident;
  * the counts match the closed forms of BASELINE.md section 3; bit-reproducible; row blocks concatenate exactly.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N = 70


@pytest.fixture(scope='module')
def problem():
    import torch
    from sucemfem import device
    mesh = device.DeviceMesh.unit(N, N, N)
    space = device.DeviceSpace(mesh, 2)
    pat = space.pattern()
    cen_z = mesh.coords[mesh.cells.long()].mean(dim=1)[:, 2]
    eps = torch.where(cen_z < 0.5, 16.0, 1.0).to(torch.float64)       # config 3: eps_r = 16 below z = 0.5
    g = torch.Generator(device='cuda')
    g.manual_seed(7)
    mu = 1.0 + torch.rand(mesh.num_cells, dtype=torch.float64, device='cuda', generator=g)
    S, T = pat.assemble_ST(eps, mu)
    return dict(mesh=mesh, space=space, pat=pat, eps=eps, mu=mu, S=S, T=T)


def _csr(pat, vals):
    import torch
    rp, cl = pat.csr()
    n = pat.space.ndof
    return torch.sparse_csr_tensor(rp, cl.long(), vals, size=(n, n))


def _constant_field_dofs(p, c):
    """Interpolant of the constant field c: dof = c . tangent of its functional."""
    import torch
    mesh, space = p['mesh'], p['space']
    ce, cf = space.cell_entities()
    x = mesh.coords
    cells = mesh.cells.long()
    c = torch.tensor(c, dtype=torch.float64, device='cuda')
    dofs = torch.zeros(space.ndof, dtype=torch.float64, device='cuda')
    E = space.n_edges
    tet_edges = [(2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1)]
    tet_faces = [(1, 2, 3), (0, 2, 3), (0, 1, 3), (0, 1, 2)]
    for le, (a, b) in enumerate(tet_edges):
        val = (x[cells[:, b]] - x[cells[:, a]]) @ c
        e = ce[:, le].long()
        dofs[2 * e] = val
        dofs[2 * e + 1] = val
    for lf, (a, b, d) in enumerate(tet_faces):
        f = cf[:, lf].long()
        dofs[2 * E + 2 * f] = (x[cells[:, b]] - x[cells[:, a]]) @ c
        dofs[2 * E + 2 * f + 1] = (x[cells[:, d]] - x[cells[:, a]]) @ c
    return dofs


def test_counts_closed_forms(problem):
    sp, pat = problem['space'], problem['pat']
    n = N
    assert problem['mesh'].num_cells == 6 * n ** 3
    assert sp.n_edges == 7 * n ** 3 + 9 * n ** 2 + 3 * n and sp.n_faces == 12 * n ** 3 + 6 * n ** 2
    assert sp.ndof == 13181420 and pat.nnz == 1660 * n ** 3 + 348 * n ** 2 + 12 * n == 571086040


def test_stiffness_annihilates_constants_and_mass_integrates_them(problem):
    import torch
    S, T = _csr(problem['pat'], problem['S']), _csr(problem['pat'], problem['T'])
    vol = 1.0 / (6 * N ** 3)
    for c in ([1.0, 0.0, 0.0], [0.3, -1.1, 0.7]):
        x = _constant_field_dofs(problem, c)
        r = S @ x
        scale = float(problem['S'].abs().max()) * float(x.abs().max())
        assert float(r.abs().max()) <= 1e-12 * scale
        got = float(x @ (T @ x))
        want = float(np.dot(c, c)) * float(problem['eps'].sum()) * vol
        assert abs(got - want) <= 1e-11 * abs(want)


def test_symmetry(problem):
    import torch
    g = torch.Generator(device='cuda')
    g.manual_seed(3)
    n = problem['space'].ndof
    x = torch.rand(n, dtype=torch.float64, device='cuda', generator=g) - 0.5
    y = torch.rand(n, dtype=torch.float64, device='cuda', generator=g) - 0.5
    for key in ('S', 'T'):
        A = _csr(problem['pat'], problem[key])
        a, b = float(y @ (A @ x)), float(x @ (A @ y))
        assert abs(a - b) <= 1e-10 * max(abs(a), abs(b), float(problem[key].abs().max()))


def test_pec_unit_rows_determinism_and_row_blocks(problem):
    import torch
    from sucemfem import device
    sp, pat = problem['space'], problem['pat']
    markers = torch.zeros(sp.n_facets, dtype=torch.int32, device='cuda')
    sp.mark_exterior_facets(markers, 999)
    assert int((markers == 999).sum()) == 12 * N ** 2
    flags = sp.pec_row_flags(markers, 999)
    assert int(flags.sum()) == 2 * (18 * N ** 2) + 2 * (12 * N ** 2)      # boundary edges and faces, 2 dofs each
    Sp, Tp = pat.assemble_ST(problem['eps'], problem['mu'], flags)
    rp, cl = pat.csr()
    rows = torch.repeat_interleave(torch.arange(sp.ndof, device='cuda'), rp[1:] - rp[:-1])
    pec_entry = flags[rows].bool()
    del rows
    diag = pec_entry & (cl.long() == torch.repeat_interleave(torch.arange(sp.ndof, device='cuda'), rp[1:] - rp[:-1]))
    for A, A0 in ((Sp, problem['S']), (Tp, problem['T'])):
        assert torch.equal(A[~pec_entry], A0[~pec_entry])                 # other rows untouched, bit for bit
        assert bool((A[diag] == 1.0).all()) and float(A[pec_entry].sum()) == float(diag.sum())
    del pec_entry, diag
    S2, T2 = pat.assemble_ST(problem['eps'], problem['mu'], flags)
    assert torch.equal(S2, Sp) and torch.equal(T2, Tp)                    # no atomics: bit-reproducible
    del S2, T2
    b = sp.partition(8)
    off = 0
    for q in range(8):
        blk = device.DevicePattern(sp, b[q], b[q + 1])
        s, t = blk.assemble_ST(problem['eps'], problem['mu'], flags)
        assert torch.equal(s, Sp[off:off + blk.nnz]) and torch.equal(t, Tp[off:off + blk.nnz])
        off += blk.nnz
        del blk, s, t
    assert off == pat.nnz
