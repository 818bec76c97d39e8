"""Device consumers of the assembled matrices: the fused multi-matrix SpMV over the shared
entity-block pattern and the native BiCGStab, against scipy on the same matrices; and the reference's solver API
(sucemfem/Utilities/LinalgSolvers.py) driving them."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import meshes

pytestmark = pytest.mark.gpu

c0 = 299792458


@pytest.mark.parametrize('mesh,order', [('cube', 2), ('cube', 1), ('square', 2), ('square', 1)])
def test_matvec_matches_scipy(mesh, order):
    import torch
    from sucemfem import device as dev, _native as N
    from sucemfem import dolfin_compat as dolfin
    co, ce = meshes.unit_cube(3, 2, 4) if mesh == 'cube' else meshes.unit_square(5, 4)
    rng = np.random.default_rng(3)
    co = co + 0.03 * rng.standard_normal(co.shape)
    sp = dev.DeviceSpace(dev.DeviceMesh(co, ce), order)
    pat = sp.pattern()
    mats = [pat.assemble(f) for f in (N.FORM_STIFFNESS, N.FORM_MASS, N.FORM_ABC)]
    coefs = [1.0, -2.5, 0.7j]
    sc = []
    for v in mats:
        m = dolfin.DeviceCSRMatrix()
        m.init(pat, v)
        sc.append(m.to_scipy())
    A = sum(c * m for c, m in zip(coefs, sc))
    x = rng.standard_normal(sp.ndof) + 1j * rng.standard_normal(sp.ndof)
    y = pat.matvec(mats, torch.from_numpy(x).cuda(), coefs).cpu().numpy()
    ref = A @ x
    assert np.abs(y - ref).max() <= 1e-13 * np.abs(ref).max()
    # real vectors, real coefficients, a single matrix (M*x of the eigen solver)
    xr = rng.standard_normal(sp.ndof)
    yr = pat.matvec(mats[1:2], torch.from_numpy(xr).cuda()).cpu().numpy()
    assert np.abs(yr - sc[1] @ xr).max() <= 1e-13 * np.abs(sc[1] @ xr).max()
    # row blocks give slices of the product (multi-GPU contract)
    b = sp.partition(3)
    blk = dev.DevicePattern(sp, b[1], b[2])
    vb = [blk.assemble(f) for f in (N.FORM_STIFFNESS, N.FORM_MASS, N.FORM_ABC)]
    yb = blk.matvec(vb, torch.from_numpy(x).cuda(), coefs).cpu().numpy()
    assert np.abs(yb - ref[blk.row_begin:blk.row_begin + blk.n_rows]).max() <= 1e-13 * np.abs(ref).max()


@pytest.mark.parametrize('mesh,order', [('cube', 2), ('cube', 1), ('square', 2), ('square', 1)])
def test_compact_abc_matrix(mesh, order):
    """S_0 in compact boundary-row storage (sfem_assemble_abc): the stored rows are exactly the non-zero rows of the
    full-pattern S_0 (bit for bit, Dirichlet unit rows included), its CSR export rebuilds the same matrix, and the
    matvec / BiCGStab that take it as the last term agree with the full-pattern ones."""
    import scipy.sparse as sps
    import torch
    from sucemfem import device as dev, _native as N
    co, ce = meshes.unit_cube(3, 2, 4) if mesh == 'cube' else meshes.unit_square(5, 4)
    rng = np.random.default_rng(5)
    co = co + 0.03 * rng.standard_normal(co.shape)
    sp = dev.DeviceSpace(dev.DeviceMesh(co, ce), order)
    b = sp.partition(2)
    for lo, hi in [(0, -1), (b[1], b[2])]:
        pat = dev.DevicePattern(sp, lo, hi)
        rowptr, cols = pat.csr()
        full = pat.assemble(N.FORM_ABC)
        comp = pat.assemble_abc()
        rows, crp, ccols = pat.abc_csr()
        nr, nz = pat.abc_info()
        assert (len(rows), len(comp), int(crp[-1])) == (nr, nz, nz)
        # the listed rows, concatenated, are the compact values and columns; every other row of the full form is zero
        rl = rows - pat.row_begin
        ln = rowptr[rl + 1] - rowptr[rl]
        idx = torch.repeat_interleave(rowptr[rl], ln) + (torch.arange(int(ln.sum()), device='cuda')
                                                        - torch.repeat_interleave(torch.cumsum(ln, 0) - ln, ln))
        assert torch.equal(full[idx], comp) and torch.equal(cols[idx], ccols)
        assert torch.equal(crp[1:] - crp[:-1], ln)
        mask = torch.ones_like(full, dtype=torch.bool)
        mask[idx] = False
        assert float(full[mask].abs().max()) == 0.0 if mask.any() else True
    # the boundary-row kernel (AUTO) against the generic kernel, with a facet mask and Dirichlet rows too.  Not bit for
    # bit: the generic kernel also adds the exterior facets that do NOT contain the row entity, whose contraction
    # vanishes only up to round-off (1e-17), where the boundary-row kernel stores exact zeros
    pat = sp.pattern()
    fm = torch.from_numpy((rng.random(sp.n_facets) < 0.6).astype(np.uint8)).cuda()
    mk = torch.zeros(sp.n_facets, dtype=torch.int32, device='cuda')
    sp.mark_exterior_facets(mk, 999)
    pf = sp.pec_row_flags(mk, 999)
    for kw in (dict(), dict(facet_mask=fm), dict(pec_flags=pf), dict(facet_mask=fm, pec_flags=pf)):
        va, vb = pat.assemble(N.FORM_ABC, **kw), pat.assemble(N.FORM_ABC, kernel=N.KERNEL_GENERIC, **kw)
        assert bool(((va - vb).abs() <= 1e-12 * vb.abs() + 1e-15 * vb.abs().max()).all())
    # consumers (pattern of all rows)
    S, M = pat.assemble(N.FORM_STIFFNESS), pat.assemble(N.FORM_MASS)
    full, comp = pat.assemble(N.FORM_ABC), pat.assemble_abc()
    coefs = [1.0, -2.5, 0.7j]
    x = torch.from_numpy(rng.standard_normal(sp.ndof) + 1j * rng.standard_normal(sp.ndof)).cuda()
    y_full = pat.matvec([S, M, full], x, coefs)
    y_comp = pat.matvec([S, M, comp], x, coefs, last_is_abc_compact=True)
    assert float((y_full - y_comp).abs().max()) <= 1e-13 * float(y_full.abs().max())
    y_only = pat.matvec([comp], x, [0.7j], last_is_abc_compact=True)
    y_ref = pat.matvec([full], x, [0.7j])
    assert float((y_only - y_ref).abs().max()) <= 1e-13 * float(y_ref.abs().max())
    xr = torch.from_numpy(rng.standard_normal(sp.ndof)).cuda()
    assert float((pat.matvec([M, comp], xr, [1.0, 2.0], last_is_abc_compact=True)
                  - pat.matvec([M, full], xr, [1.0, 2.0])).abs().max()) <= 1e-12
    cs = [1.0, 1.0, 0.5j]     # S + M is positive definite: BiCGStab converges whatever the mesh
    rhs = torch.from_numpy(rng.standard_normal(sp.ndof) + 0j).cuda()
    if mesh == 'cube':
        xa, ia, _, _ = pat.bicgstab([S, M, full], cs, rhs, rtol=1e-8)
        xb, ib, _, _ = pat.bicgstab([S, M, comp], cs, rhs, rtol=1e-8, last_is_abc_compact=True)
        assert ia == 0 and ib == 0
        for xs in (xa, xb):     # true residual through the full-pattern operator
            r = rhs - pat.matvec([S, M, full], xs, cs)
            assert float(torch.linalg.vector_norm(r)) <= 1e-7 * float(torch.linalg.vector_norm(rhs))
    # Dirichlet rows inside the stored set become unit rows, as in the full form
    markers = torch.zeros(sp.n_facets, dtype=torch.int32, device='cuda')
    sp.mark_exterior_facets(markers, 999)
    flags = sp.pec_row_flags(markers, 999)
    rows, crp, _ = pat.abc_csr(with_cols=False)
    fullp, compp = pat.assemble(N.FORM_ABC, pec_flags=flags), pat.assemble_abc(pec_flags=flags)
    rowptr, _ = pat.csr()
    ln = rowptr[rows + 1] - rowptr[rows]
    idx = torch.repeat_interleave(rowptr[rows], ln) + (torch.arange(int(ln.sum()), device='cuda')
                                                      - torch.repeat_interleave(torch.cumsum(ln, 0) - ln, ln))
    assert torch.equal(fullp[idx], compp)


def _problem(device_resident):
    from sucemfem import dolfin_compat as dolfin
    from sucemfem.BoundaryConditions import ABCBoundaryCondition, BoundaryConditions
    from sucemfem.ProblemConfigurations.EMDrivenProblem import DrivenProblemABC
    from sucemfem.Utilities.MeshGenerators import get_centred_cube
    from sucemfem.Sources import current_source
    from sucemfem.Sources.fillament_current_source import FillamentCurrentSource
    g = np.load(os.path.join(GOLDEN, 'driven_order2_unitcube3.npz'))
    freq = float(g['freq'])
    lam = c0 / freq
    mesh = get_centred_cube(np.array([lam] * 3) * 0.25, lam / 6)
    srcs = current_source.CurrentSources()
    fil = FillamentCurrentSource()
    fil.set_source_endpoints(g['source_endpoints'])
    fil.set_value(float(g['I']))
    srcs.add_source(fil)
    material_mesh_func = dolfin.MeshFunction('uint', mesh, 3)
    material_mesh_func.set_all(0)
    abc = ABCBoundaryCondition()
    abc.set_region_number(1)
    bcs = BoundaryConditions()
    bcs.add_boundary_condition(abc)
    dp = DrivenProblemABC()
    if device_resident:
        dp.set_matrix_class(dolfin.PETScMatrix)
    dp.set_mesh(mesh)
    dp.set_basis_order(2)
    dp.set_material_regions({0: dict(eps_r=1, mu_r=1)})
    dp.set_region_meshfunction(material_mesh_func)
    dp.set_boundary_conditions(bcs)
    dp.set_sources(srcs)
    dp.init_problem()
    dp.set_frequency(freq)
    return dp, g


def test_device_bicgstab_reproduces_reference_solution():
    """The reference's shipped order-2 driven solution (fixture 5, generated with UMFPACK) from the DEVICE solver:
    S, M, S_0 stay in HBM, A is applied as a fused three-matrix SpMV, BiCGStab with the reference's 'diagonal'
    preconditioner (examples/infinitesimal_dipole/driver.py:83)."""
    from sucemfem.Utilities import LinalgSolvers
    dp, g = _problem(device_resident=True)
    A = dp.get_LHS_operator()
    b = dp.get_RHS()
    # the operator is the reference's LHS matrix
    As = A.to_scipy()
    x0 = np.random.default_rng(0).standard_normal(len(b)) + 0j
    assert np.abs(A * x0 - As @ x0).max() <= 1e-13 * np.abs(As @ x0).max()
    assert np.allclose(A.diagonal(), As.diagonal(), rtol=1e-13, atol=0)
    solver = LinalgSolvers.BiCGStabSolver(A, 'diagonal')
    # Jacobi-BiCGStab needs ~10 n iterations on this indefinite system (1314 dofs: 10 600 - 14 100 iterations to
    # 1e-11 depending on round-off in the last bit of the matrix); the default cap of 10 n sits right in that range
    solver.rtol = 1e-11
    solver.maxiter = 40 * len(b)
    x = solver.solve(b)
    assert solver.info == 0 and solver.iterations > 1
    assert np.linalg.norm(As @ x - b) <= 1e-10 * np.linalg.norm(b)
    assert np.linalg.norm(x - g['x']) / np.linalg.norm(g['x']) < 1e-8
    # un-preconditioned, default tolerance, through the reference's convenience function
    x2 = LinalgSolvers.solve_sparse_system(A, b, preconditioner_type=None)
    assert np.linalg.norm(As @ x2 - b) <= 1.01e-5 * np.linalg.norm(b)


def test_host_solver_path_matches_reference_usage():
    """solve_sparse_system on the scipy LHS (what the reference's drivers do) gives the same solution."""
    from sucemfem.Utilities import LinalgSolvers
    dp, g = _problem(device_resident=False)
    A, b = dp.get_LHS_matrix(), dp.get_RHS()
    s = LinalgSolvers.BiCGStabSolver(A, 'diagonal')
    s.rtol = 1e-11
    s.maxiter = 40 * len(b)
    x = s.solve(b)
    assert np.linalg.norm(x - g['x']) / np.linalg.norm(g['x']) < 1e-8
    x3 = LinalgSolvers.UMFPACKSolver(A).solve(b)
    assert np.linalg.norm(x3 - g['x']) / np.linalg.norm(g['x']) < 1e-11


def test_bicgstab_reports_non_convergence_and_zero_rhs():
    import torch
    from sucemfem import device as dev, _native as N
    co, ce = meshes.unit_cube(2)
    sp = dev.DeviceSpace(dev.DeviceMesh(co, ce), 2)
    pat = sp.pattern()
    S, M = pat.assemble_ST()
    b = torch.zeros(sp.ndof, dtype=torch.complex128, device='cuda')
    x, info, its, rr = pat.bicgstab([S, M], [1.0, 1.0], b)
    assert info == 0 and float(x.abs().max()) == 0.0
    b[3] = 1.0
    x, info, its, rr = pat.bicgstab([S, M], [1.0, 1.0], b, rtol=1e-14, maxiter=2)
    assert info > 0 and its == 2 and rr > 1e-14


def test_infinitesimal_dipole_driver_on_the_device():
    """/root/reference/examples/infinitesimal_dipole/driver.py:41-90 (BASELINE config 4 at the example's own size):
    2-lambda cube, lambda/6 edges, first-order ABC, z-directed point dipole at the origin, order 1;
    `solve_sparse_system(A, b, preconditioner_type='diagonal')` runs on the device (matrices never leave HBM), the field
    along the x axis is reconstructed and compared with the analytic infinitesimal-dipole field."""
    from sucemfem import dolfin_compat as dolfin
    from sucemfem.BoundaryConditions import ABCBoundaryCondition, BoundaryConditions
    from sucemfem.ProblemConfigurations.EMDrivenProblem import DrivenProblemABC
    from sucemfem.Utilities.MeshGenerators import get_centred_cube
    from sucemfem.Utilities.LinalgSolvers import solve_sparse_system, BiCGStabSolver
    from sucemfem.Sources import current_source, point_source
    from sucemfem.PostProcessing import Reconstruct
    from sucemfem.Consts import Z0
    freq = 1e9
    lam = c0 / freq
    source_coord = np.array([0, 0, 0.])
    I, l = 1.0, lam / 1000
    order = 1
    mesh = get_centred_cube(np.array([2 * lam] * 3), lam / 6, source_coord)
    material_mesh_func = dolfin.MeshFunction('uint', mesh, 3)
    material_mesh_func.set_all(0)
    abc = ABCBoundaryCondition()
    abc.set_region_number(1)
    bcs = BoundaryConditions()
    bcs.add_boundary_condition(abc)
    dp = DrivenProblemABC()
    dp.set_matrix_class(dolfin.PETScMatrix)          # keep S, M, S_0 on the device
    dp.set_mesh(mesh)
    dp.set_basis_order(order)
    dp.set_material_regions({0: dict(eps_r=1, mu_r=1)})
    dp.set_region_meshfunction(material_mesh_func)
    dp.set_boundary_conditions(bcs)
    srcs = current_source.CurrentSources()
    dipole = point_source.PointCurrentSource()
    dipole.set_position(source_coord)
    dipole.set_value(np.array([0, 0, 1.]) * I * l)
    srcs.add_source(dipole)
    dp.set_sources(srcs)
    dp.init_problem()
    dp.set_frequency(freq)
    assert mesh.num_cells() == 6 * 17 ** 3
    A, b = dp.get_LHS_operator(), dp.get_RHS()
    solver = BiCGStabSolver(A, 'diagonal')
    solver.rtol = 1e-8
    x = solver.solve(b)
    assert solver.info == 0
    As = A.to_scipy()
    assert np.linalg.norm(As @ x - b) <= 1e-7 * np.linalg.norm(b)
    # the reference's call, default tolerance
    x5 = solve_sparse_system(A, b, preconditioner_type='diagonal')
    assert np.linalg.norm(x5 - x) <= 1e-3 * np.linalg.norm(x)
    # field on the x axis against the analytic dipole: E_z = -E_theta(theta = 90 deg)
    r = lam * (np.arange(30) / 50 + 0.3)
    pts = np.array([1.0, 0, 0])[None, :] * r[:, None]
    rec = Reconstruct(dp.function_space)
    rec.set_dof_values(x)
    E = rec.reconstruct_points(pts)
    k0 = 2 * np.pi / lam
    kr = k0 * r
    E_theta = 1j * Z0 * k0 * I * l / (4 * np.pi * r) * (1 + 1 / (1j * kr) - 1 / kr ** 2) * np.exp(-1j * kr)
    err = np.abs(np.abs(E[:, 2]) - np.abs(E_theta)) / np.abs(E_theta)
    assert np.median(err) < 0.15 and err.max() < 0.5          # order 1 at lambda/6: the plot-level agreement of the example
