"""The reference's own API, end to end on the GPU: the test bodies follow the reference's tests/drivers
(file:line in each docstring) with `dolfin` replaced by sucemfem.dolfin_compat."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

c0 = 299792458


def test_get_LHS_matrix_reference_fixture():
    """/root/reference/sucemfem/ProblemConfigurations/tests/test_EMDrivenProblem.py:35-72: order-1 driven problem on
    InscribedTet with a first-order ABC at 1e8 Hz; the dense LHS must match LHS_matrix.npy with the reference's own
    tolerance rtol=1e-10, atol=3e-15."""
    from sucemfem import dolfin_compat as dolfin
    from oracle import meshes as omeshes
    from sucemfem.BoundaryConditions import ABCBoundaryCondition, BoundaryConditions
    from sucemfem.ProblemConfigurations import EMDrivenProblem
    mesh = dolfin.Mesh(coordinates=omeshes.inscribed_tet()[0], cells=omeshes.inscribed_tet()[1])
    material_mesh_func = dolfin.MeshFunction('uint', mesh, 3)
    material_mesh_func.set_all(0)
    materials = {0: dict(eps_r=1, mu_r=1)}
    abc = ABCBoundaryCondition()
    abc.set_region_number(1)
    bcs = BoundaryConditions()
    bcs.add_boundary_condition(abc)
    DUT = EMDrivenProblem.DrivenProblemABC()
    DUT.set_mesh(mesh)
    DUT.set_basis_order(1)
    DUT.set_material_regions(materials)
    DUT.set_region_meshfunction(material_mesh_func)
    DUT.set_boundary_conditions(bcs)
    DUT.set_frequency(1e8)
    DUT.init_problem()
    assert DUT.get_global_dimension() == 24
    actual = DUT.get_LHS_matrix().todense()
    desired = np.load(os.path.join(GOLDEN, 'driven_order1_inscribedtet.npz'))['LHS']
    assert np.allclose(actual, desired, rtol=1e-10, atol=3e-15)


@pytest.mark.parametrize('order,expected', [(1, [1.386086, 1.421522, 1.421522, 1.750142, 1.750142]),
                                            (2, [1.413965, 1.415131, 1.415131, 1.733798, 1.733798])])
def test_resonant_cavity_driver(order, expected):
    """/root/reference/examples/resonant_cavity/driver.py:34-66 (BASELINE config 1) at the orders this build provides:
    UnitCube(4), PEC walls, 10 lowest modes; k/pi against the regression anchors of SURVEY.md 8.3 and, loosely,
    against the analytic sqrt(2), sqrt(2), sqrt(2), sqrt(3), sqrt(3)."""
    from sucemfem import dolfin_compat as dol
    from sucemfem.ProblemConfigurations.EMVectorWaveEigenproblem import EigenProblem, DefaultEigenSolver
    from sucemfem.BoundaryConditions import PECWallsBoundaryCondition
    mesh = dol.UnitCube(4, 4, 4)
    pec_walls = PECWallsBoundaryCondition()
    pec_walls.init_with_mesh(mesh)
    ep = EigenProblem()
    ep.set_mesh(mesh)
    ep.set_basis_order(order)
    ep.set_boundary_conditions(pec_walls)
    ep.init_problem()
    assert ep.get_global_dimension() == {1: 604, 2: 2936}[order]
    assert ep.system_matrices['BC'] is None            # NullForm -> None (SystemMatrices.py:51-52)
    es = DefaultEigenSolver()
    es.set_eigenproblem(ep)
    es.set_sigma(1.1)
    eigs_w, eigs_v = es.solve_problem(10)
    res = np.sqrt(np.array(sorted(eigs_w))) / np.pi
    assert np.allclose(res[:5], expected, atol=2e-6)
    assert np.allclose(res[:5], [np.sqrt(2)] * 3 + [np.sqrt(3)] * 2, rtol=0.03)
    assert eigs_v.shape == (10, ep.get_global_dimension())


@pytest.mark.parametrize('order,pec,expected', [
    (1, True, [0.99027, 0.99882, 1.42117, 1.95559, 1.95692]),
    (2, True, [0.99997, 1.00003, 1.41447, 2.00012, 2.00012]),
    (1, False, [1.42167, 2.214, 2.25022, 2.86333]),
    (2, False, [1.41448, 2.23712, 2.23838, 2.83564])])
def test_2d_waveguide_driver(order, pec, expected):
    """/root/reference/examples/2D_waveguide/eigenmode/driver.py:38-72 (TE: PEC walls) and TM_driver.py (no essential
    condition) = BASELINE config 2 at the orders this build provides: UnitSquare(5,5) triangles, cutoff wavenumbers
    k/pi against the regression anchors of SURVEY.md 8.3 and, loosely, against the analytic values."""
    from sucemfem import dolfin_compat as dol
    from sucemfem.ProblemConfigurations.EMVectorWaveEigenproblem import EigenProblem, DefaultEigenSolver
    from sucemfem.BoundaryConditions import PECWallsBoundaryCondition
    mesh = dol.UnitSquare(5, 5)
    ep = EigenProblem()
    ep.set_mesh(mesh)
    ep.set_basis_order(order)
    if pec:
        pec_walls = PECWallsBoundaryCondition()
        pec_walls.init_with_mesh(mesh)
        ep.set_boundary_conditions(pec_walls)
    ep.init_problem()
    assert ep.get_global_dimension() == {1: 85, 2: 270}[order]
    es = DefaultEigenSolver()
    es.set_eigenproblem(ep)
    es.set_sigma(1.5)
    eigs_w, eigs_v = es.solve_problem(10)
    res = np.sqrt(np.array(sorted(w for w in eigs_w.real if w > 0.5))) / np.pi
    if pec:     # the identity rows of the PEC dofs put spurious modes at lambda = 1 (k/pi = 0.3183): drop them
        res = res[np.abs(res - 1 / np.pi) > 1e-6]
    assert np.allclose(res[:len(expected)], expected, atol=2e-5)
    analytic = [1, 1, np.sqrt(2), 2, 2] if pec else [np.sqrt(2), np.sqrt(5), np.sqrt(5), np.sqrt(8)]
    assert np.allclose(res[:len(analytic)], analytic, rtol=0.03)


def test_cavity_eigenvalues_match_oracle_matrices_to_1e10():
    """north_star: resonant-cavity eigenvalues computed from both matrix sets must agree to 1e-10 (relative)."""
    import scipy.linalg as sla
    import torch
    from oracle import fem, meshes
    from sucemfem import device
    co, ce = meshes.unit_cube(3)
    rng = np.random.default_rng(3)
    eps = np.where(co[ce].mean(axis=1)[:, 2] < 0.5, 16.0, 1.0)    # inhomogeneous eps_r as in BASELINE config 3
    num = fem.number_entities(ce)
    mask = fem.exterior_facet_mask(ce, num)
    ref = fem.assemble_system(co, ce, 2, eps_r=eps)
    pec = fem.facet_closure_dofs(ce, num, 2, mask)
    free = np.setdiff1d(np.arange(ref['ndof']), pec)
    sp = device.DeviceSpace(device.DeviceMesh(co, ce), 2)
    pat = sp.pattern()
    S, T = pat.assemble_ST(torch.from_numpy(eps).cuda(), None)
    def eig(Sv, Mv):
        A = fem.to_scipy(ref['rowptr'], ref['cols'], Sv, ref['ndof']).toarray()[np.ix_(free, free)]
        B = fem.to_scipy(ref['rowptr'], ref['cols'], Mv, ref['ndof']).toarray()[np.ix_(free, free)]
        w = sla.eigh(A, B, eigvals_only=True)
        return w[w > 1e-6][:10]
    w_ref, w_gpu = eig(ref['S'], ref['M']), eig(S.cpu().numpy(), T.cpu().numpy())
    assert np.allclose(w_gpu, w_ref, rtol=1e-10, atol=0)


def test_materials_errors_and_defaults():
    """/root/reference/sucemfem/Materials.py:40,84-87,147-148."""
    from sucemfem import Materials
    from sucemfem import dolfin_compat as dolfin
    with pytest.raises(ValueError, match='Unknown material property'):
        Materials.MaterialPropertiesFactory({0: dict(sigma=3)})
    mesh = dolfin.UnitCube(1, 1, 1)
    mf = dolfin.MeshFunction('uint', mesh, 3)
    mf.set_all(0)
    mf.array()[2] = 7
    props = Materials.MaterialPropertiesFactory({0: dict(eps_r=2)}).get_material_properties()
    with pytest.raises(ValueError, match='Material number 7 not found'):
        Materials.MaterialFunctionFactory(props, mf, mesh).get_material_functions('eps_r', 'mu_r')
    fns = Materials.MaterialFunctionFactory(Materials.MaterialPropertiesFactory(None).get_material_properties(),
                                            None, mesh).get_material_functions('eps_r', 'mu_r')
    assert np.all(fns['eps_r'].vector() == 1.0) and np.all(fns['mu_r'].vector() == 1.0)


def test_duplicate_bc_number_asserts():
    """/root/reference/sucemfem/BoundaryConditions/container.py:29-33."""
    from sucemfem.BoundaryConditions import ABCBoundaryCondition, BoundaryConditions
    bcs = BoundaryConditions()
    a, b = ABCBoundaryCondition(), ABCBoundaryCondition()
    a.set_region_number(1)
    b.set_region_number(1)
    bcs.add_boundary_condition(a)
    with pytest.raises(AssertionError):
        bcs.add_boundary_condition(b)


def test_get_centred_cube_matches_reference_fixture_mesh():
    """/root/reference/sucemfem/Utilities/MeshGenerators.py:26-53; fixture 7's recentring picks cell 372."""
    from oracle import meshes
    from sucemfem.Utilities.MeshGenerators import get_centred_cube
    lam = c0 / 1e9
    mesh = get_centred_cube(np.array([lam] * 3) * 0.5, lam / 6, np.zeros(3))
    co, ce = meshes.centred_cube(np.array([lam] * 3) * 0.5, lam / 6)
    assert np.array_equal(mesh.cells(), ce)
    assert np.allclose(mesh.coordinates(), co - co[ce[372]].mean(axis=0), rtol=0, atol=1e-17)


def test_gmsh_prism_pec_markers_and_cavity_modes():
    """/root/reference/functional_tests/test_EigenProblems.py:69-134 on the reference's gmsh mesh with its PEC
    facet-marker file (converted to tests/golden/rectangular_prism_gmsh.npz), at order 2 (the reference runs order 4
    to 4 decimals; order 2 on this coarse mesh gives the lowest modes of the 1 x 0.5 x 0.25 cavity to 0.3 %)."""
    import torch
    from oracle import fem
    from sucemfem import dolfin_compat as dolfin
    from sucemfem.BoundaryConditions import PECWallsBoundaryCondition
    from sucemfem.ProblemConfigurations.EMVectorWaveEigenproblem import EigenProblem, DefaultEigenSolver
    g = np.load(os.path.join(GOLDEN, 'rectangular_prism_gmsh.npz'))
    mesh = dolfin.Mesh(coordinates=g['coordinates'], cells=g['cells'])
    pec_mesh_function = dolfin.MeshFunction('uint', mesh, 2)
    assert len(pec_mesh_function.array()) == len(g['facet_markers'])      # same number of faces
    pec_mesh_function.array()[:] = g['facet_markers']
    # the device's exterior-facet marking reproduces the reference's marker file exactly
    chk = dolfin.MeshFunction('uint', mesh, 2)
    dolfin.DomainBoundary().mark(chk, 1)
    assert np.array_equal(chk.array(), g['facet_markers'])
    pec_walls = PECWallsBoundaryCondition()
    pec_walls.init_with_meshfunction(pec_mesh_function, 1)
    ep = EigenProblem()
    ep.set_mesh(mesh)
    ep.set_basis_order(2)
    ep.set_boundary_conditions(pec_walls)
    ep.init_problem()
    # matrices against the oracle on this unstructured mesh (canonical kernels, ring order on a real gmsh mesh)
    num = fem.number_entities(g['cells'])
    ref = fem.assemble_system(g['coordinates'], g['cells'], 2, pec_facet_mask=g['facet_markers'] == 1)
    ref0 = fem.assemble_system(g['coordinates'], g['cells'], 2)
    for name, key in (('S', 'S'), ('M', 'M')):
        rp, cl, v = ep.system_matrices[name].data()
        assert np.array_equal(rp, ref['rowptr']) and np.array_equal(cl, ref['cols'])
        assert np.all(np.abs(v - ref[key]) <= 1e-12 * np.abs(ref[key]) + 1e-15 * np.abs(ref0[key]).max())
    es = DefaultEigenSolver()
    es.set_eigenproblem(ep)
    es.set_sigma(1.1)
    eigs_w, _ = es.solve_problem(4)
    res = np.sqrt(np.array(sorted(eigs_w))) / np.pi
    a, b, d = 1.0, 0.5, 0.25
    k = lambda m, n, l: np.sqrt((m / a) ** 2 + (n / b) ** 2 + (l / d) ** 2)
    analytic = sorted([k(1, 1, 0), k(2, 1, 0), k(3, 1, 0), k(1, 0, 1), k(0, 1, 1), k(1, 2, 0)])[:4]
    assert np.allclose(res, analytic, rtol=5e-3)


def test_high_valence_edge_falls_back_to_generic_kernel():
    """Per-ENTITY fallback: an edge shared by 70 / 40 / 34 cells has more items than a warp has lanes (and a row longer than
    the shared-memory tile); the batch builder leaves exactly that edge out and the canonical path hands its rows to the
    generic kernel, every other row stays on the canonical kernels (5 launches: pack, edges, faces, cell geometry, generic
    rows).  Closed rings of 12, 20 and 30 cells go through the canonical kernels alone (ring-closing shuffle, batches limited
    by the tile)."""
    import torch
    from oracle import fem
    from sucemfem import device, _native as N
    for m in (70, 40, 34):
        ang = 2 * np.pi * np.arange(m) / m
        co = np.concatenate([[[0, 0, -1.0], [0, 0, 1.0]], np.stack([np.cos(ang), np.sin(ang), 0 * ang], 1)])
        ce = np.sort(np.array([[0, 1, 2 + i, 2 + (i + 1) % m] for i in range(m)]), axis=1)
        ref = fem.assemble_system(co, ce, 2)
        sp = device.DeviceSpace(device.DeviceMesh(co, ce), 2)
        pat = sp.pattern()
        S, T = pat.assemble_ST(kernel=N.KERNEL_AUTO)
        assert np.all(np.abs(S.cpu().numpy() - ref['S']) <= 1e-12 * np.abs(ref['S']) + 1e-15 * np.abs(ref['S']).max())
        assert np.all(np.abs(T.cpu().numpy() - ref['M']) <= 1e-12 * np.abs(ref['M']) + 1e-15 * np.abs(ref['M']).max())
        assert N.lib.sfem_launches_per_assemble_ST(pat.handle, N.KERNEL_AUTO) == 5
        Sc, Tc = pat.assemble_ST(kernel=N.KERNEL_CANONICAL)
        assert torch.equal(Sc, S) and torch.equal(Tc, T)
        Sg, Tg = pat.assemble_ST(kernel=N.KERNEL_GENERIC)
        assert not torch.equal(Sg, S)        # the other rows really come from the canonical kernels
        # order 1 on the same fan: AUTO / ROWS = the thread-per-edge kernel; the hub edge's row (1 + 3 m entries)
        # makes its warp's span exceed the shared-memory tile at m = 70, which takes the direct-write path
        ref1 = fem.assemble_system(co, ce, 1)
        pat1 = device.DeviceSpace(device.DeviceMesh(co, ce), 1).pattern()
        for kern in (N.KERNEL_AUTO, N.KERNEL_ROWS):
            S1, T1 = pat1.assemble_ST(kernel=kern)
            assert np.all(np.abs(S1.cpu().numpy() - ref1['S']) <= 1e-12 * np.abs(ref1['S']) + 1e-15 * np.abs(ref1['S']).max())
            assert np.all(np.abs(T1.cpu().numpy() - ref1['M']) <= 1e-12 * np.abs(ref1['M']) + 1e-15 * np.abs(ref1['M']).max())
    for m in (12, 20, 30):
        ang = 2 * np.pi * np.arange(m) / m
        co = np.concatenate([[[0, 0, -1.0], [0, 0, 1.0]], np.stack([np.cos(ang), np.sin(ang), 0 * ang], 1)])
        ce = np.sort(np.array([[0, 1, 2 + i, 2 + (i + 1) % m] for i in range(m)]), axis=1)
        ref = fem.assemble_system(co, ce, 2)
        pat = device.DeviceSpace(device.DeviceMesh(co, ce), 2).pattern()
        assert N.lib.sfem_launches_per_assemble_ST(pat.handle, N.KERNEL_AUTO) == 3
        S, T = pat.assemble_ST(kernel=N.KERNEL_CANONICAL)
        assert np.all(np.abs(S.cpu().numpy() - ref['S']) <= 1e-12 * np.abs(ref['S']) + 1e-15 * np.abs(ref['S']).max())
        assert np.all(np.abs(T.cpu().numpy() - ref['M']) <= 1e-12 * np.abs(ref['M']) + 1e-15 * np.abs(ref['M']).max())
    # an open fan (the hub edge is on the boundary): chain start without a predecessor, open end stores its next-side blocks
    m = 9
    ang = 1.5 * np.pi * np.arange(m + 1) / m
    co = np.concatenate([[[0, 0, -1.0], [0, 0, 1.0]], np.stack([np.cos(ang), np.sin(ang), 0 * ang], 1)])
    ce = np.sort(np.array([[0, 1, 2 + i, 3 + i] for i in range(m)]), axis=1)
    ref = fem.assemble_system(co, ce, 2)
    pat = device.DeviceSpace(device.DeviceMesh(co, ce), 2).pattern()
    S, T = pat.assemble_ST(kernel=N.KERNEL_CANONICAL)
    assert np.all(np.abs(S.cpu().numpy() - ref['S']) <= 1e-12 * np.abs(ref['S']) + 1e-15 * np.abs(ref['S']).max())
    assert np.all(np.abs(T.cpu().numpy() - ref['M']) <= 1e-12 * np.abs(ref['M']) + 1e-15 * np.abs(ref['M']).max())


def test_very_high_valence_edge_uses_the_large_row_pattern_launch():
    """An edge shared by 150 cells: 1500 candidate column entities for its row, more than the pattern kernel's normal
    launch holds in shared memory (1024).  The space switches to the one-warp-per-block launch with large arrays; pattern
    bit-exact, values by the per-entity fallback (hub edge -> generic kernel), orders 2 and 1."""
    import torch
    from oracle import fem
    from sucemfem import device, _native as N
    m = 150
    ang = 2 * np.pi * np.arange(m) / m
    co = np.concatenate([[[0, 0, -1.0], [0, 0, 1.0]], np.stack([np.cos(ang), np.sin(ang), 0 * ang], 1)])
    ce = np.sort(np.array([[0, 1, 2 + i, 2 + (i + 1) % m] for i in range(m)]), axis=1)
    for order in (2, 1):
        ref = fem.assemble_system(co, ce, order)
        sp = device.DeviceSpace(device.DeviceMesh(co, ce), order)
        pat = sp.pattern()
        rp, cl = pat.csr()
        assert np.array_equal(rp.cpu().numpy(), ref['rowptr']) and np.array_equal(cl.cpu().numpy(), ref['cols'])
        S, T = pat.assemble_ST(kernel=N.KERNEL_AUTO)
        assert np.all(np.abs(S.cpu().numpy() - ref['S']) <= 1e-12 * np.abs(ref['S']) + 1e-15 * np.abs(ref['S']).max())
        assert np.all(np.abs(T.cpu().numpy() - ref['M']) <= 1e-12 * np.abs(ref['M']) + 1e-15 * np.abs(ref['M']).max())


def _driven_problem(mesh, order, sources, freq):
    from sucemfem import dolfin_compat as dolfin
    from sucemfem.BoundaryConditions import ABCBoundaryCondition, BoundaryConditions
    from sucemfem.ProblemConfigurations.EMDrivenProblem import DrivenProblemABC
    material_mesh_func = dolfin.MeshFunction('uint', mesh, 3)
    material_mesh_func.set_all(0)
    abc = ABCBoundaryCondition()
    abc.set_region_number(1)
    bcs = BoundaryConditions()
    bcs.add_boundary_condition(abc)
    dp = DrivenProblemABC()
    dp.set_mesh(mesh)
    dp.set_basis_order(order)
    dp.set_material_regions({0: dict(eps_r=1, mu_r=1)})
    dp.set_region_meshfunction(material_mesh_func)
    dp.set_boundary_conditions(bcs)
    dp.set_sources(sources)
    dp.init_problem()
    dp.set_frequency(freq)
    return dp


def test_get_RHS_reference_fixture():
    """/root/reference/sucemfem/ProblemConfigurations/tests/test_EMDrivenProblem.py:74-82: point dipole [2,-1,3] at the
    origin of InscribedTet, order 1; RHS vs RHS_vector.pickle with the reference's rtol=1e-12, atol=1e-16."""
    from oracle import meshes as omeshes
    from sucemfem.Sources import point_source, current_source
    g = np.load(os.path.join(GOLDEN, 'driven_order1_inscribedtet.npz'))
    srcs = current_source.CurrentSources()
    dipole = point_source.PointCurrentSource()
    dipole.set_position(np.array([0, 0, 0.]))
    dipole.set_value(np.array([2, -1, 3.]))
    srcs.add_source(dipole)
    from sucemfem import dolfin_compat as dolfin
    co, ce = omeshes.inscribed_tet()
    dp = _driven_problem(dolfin.Mesh(coordinates=co, cells=ce), 1, srcs, 1e8)
    assert np.allclose(dp.get_RHS(), g['RHS'], rtol=1e-12, atol=1e-16)


def test_fillament_source_reference_fixture():
    """/root/reference/sucemfem/Sources/tests/test_fillament_source.py:45-82: order-2 dof numbers and values of a
    20-point current filament on get_centred_cube(lambda/2, lambda/3); the 26 significant entries of the fixture (the
    rest is round-off on a shared face where DOLFIN's point location picked the other cell)."""
    from sucemfem import dolfin_compat as dolfin
    from sucemfem.Utilities.MeshGenerators import get_centred_cube
    from sucemfem.Sources.fillament_current_source import FillamentCurrentSource
    g = np.load(os.path.join(GOLDEN, 'fillament_source_order2.npz'))
    mesh = get_centred_cube(g['domain_size'], float(g['max_edge_len']))
    V = dolfin.FunctionSpace(mesh, "Nedelec 1st kind H(curl)", int(g['order']))
    DUT = FillamentCurrentSource()
    DUT.set_function_space(V)
    DUT.set_no_integration_points(int(g['no_integration_points']))
    DUT.set_source_endpoints(g['source_endpoints'])
    DUT.set_value(float(g['I']))
    dofnos, rhs = DUT.get_contribution()
    sig = np.abs(g['rhs_contribs']) > 1e-12
    mine = dict(zip(dofnos.tolist(), rhs))
    assert sorted(k for k, v in mine.items() if abs(v) > 1e-12) == sorted(g['dofnos'][sig].tolist())
    got = np.array([mine[k] for k in g['dofnos'][sig].tolist()])
    assert np.allclose(got, g['rhs_contribs'][sig], rtol=1e-10, atol=1e-12)
    # test_negative_direction (:71-82)
    DUT.set_source_endpoints(-g['source_endpoints'])
    dn2, rhs2 = DUT.get_contribution()
    m2 = dict(zip(dn2.tolist(), rhs2))
    assert all(abs(mine[k] + m2[k]) < 1e-12 for k in mine if abs(mine[k]) > 1e-12)


def test_driven_solution_reference_fixture_order2():
    """The reference's shipped order-2 solution vector (power_flux_reference_dofs..., generated by
    sucemfem/PostProcessing/tests/data/power_flux_generate_dofs.py:44-92: lambda/4 cube, lambda/6 edges, 100-point
    filament, UMFPACK) is reproduced by the PRODUCT path end to end: device S, M, S_0 + host sources + sparse solve."""
    import scipy.sparse.linalg as spla
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
    dp = _driven_problem(mesh, 2, srcs, freq)
    assert dp.get_global_dimension() == 1314
    A, b = dp.get_LHS_matrix(), dp.get_RHS()
    x = spla.spsolve(A.tocsc(), b)
    assert np.linalg.norm(x - g['x']) / np.linalg.norm(g['x']) < 1e-11


def test_driven_residual_reference_fixture_unitcube5():
    """reference_dofs-2-... (sucemfem/PostProcessing/tests/data/ntff_generate_dofs.py:40-85: UnitCube(5) recentred on
    the dipole, order 2, ABC): the shipped solution satisfies the product's A x = b to 1e-11."""
    from sucemfem.Utilities.MeshGenerators import get_centred_cube
    from sucemfem.Sources import current_source, point_source
    g = np.load(os.path.join(GOLDEN, 'driven_order2_unitcube5.npz'))
    freq = float(g['freq'])
    lam = c0 / freq
    source_coord = np.array([0, 0, 0.])
    mesh = get_centred_cube(g['domain_size'], float(g['max_edge_len']), source_coord)
    srcs = current_source.CurrentSources()
    dipole = point_source.PointCurrentSource()
    dipole.set_position(source_coord)
    dipole.set_value(np.array([0, 0, 1.]) * lam / 1000)
    srcs.add_source(dipole)
    dp = _driven_problem(mesh, 2, srcs, freq)
    assert dp.get_global_dimension() == 5530
    A, b = dp.get_LHS_matrix(), dp.get_RHS()
    assert np.linalg.norm(A @ g['x'] - b) / np.linalg.norm(b) < 1e-11
