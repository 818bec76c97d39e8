"""Host-side mirrors around the path that need no GPU: mesh file formats, matrix IO, the solver classes on scipy
matrices.  Bodies follow the reference's tests (file:line in the docstrings)."""
import os

import numpy as np

from conftest import GOLDEN
import scipy.sparse


def test_sparse_identity_and_residual_shapes():
    """/root/reference/sucemfem/Utilities/tests/test_LinalgSolvers.py:35-100"""
    from sucemfem.Utilities.LinalgSolvers import solve_sparse_system, calculate_residual, BiCGStabSolver, UMFPACKSolver
    n = 1000
    A = scipy.sparse.eye(n, n, format='csr')
    b = np.random.default_rng(0).random(n)
    x = solve_sparse_system(A, b)
    np.testing.assert_array_equal(b, x)
    x = solve_sparse_system(A, b, preconditioner_type='diagonal')
    np.testing.assert_array_equal(b, x)
    for bb, xx in ((b, b.copy()), (b.reshape(1, n), b.copy()), (b.reshape(n, 1), b.copy()),
                   (b, b.copy().reshape(n, 1)), (b.reshape(1, n), b.copy().reshape(n, 1)),
                   (b.reshape(n, 1), b.copy().reshape(n, 1))):
        assert calculate_residual(A, xx, bb) == 0
    # a non-trivial complex system, logging as in the reference
    rng = np.random.default_rng(1)
    B = scipy.sparse.random(200, 200, density=0.05, random_state=2, format='csr') + 10 * scipy.sparse.eye(200)
    B = (B + 0.3j * scipy.sparse.eye(200)).tocsr()
    rhs = rng.standard_normal(200) + 1j * rng.standard_normal(200)
    s = BiCGStabSolver(B, 'diagonal')
    s.rtol = 1e-12
    seen = []
    s.add_user_callback(lambda solver, res: seen.append(res))
    x = s.solve(rhs)
    assert np.linalg.norm(B @ x - rhs) <= 1e-10 * np.linalg.norm(rhs)
    log = s.get_logging_data()
    assert log['id'][0] == 'init' and 'solve::begin' in log['id'] and log['id'][-1] == 'solve::end' and len(seen) > 4
    assert np.allclose(UMFPACKSolver(B).solve(rhs), x, rtol=1e-9, atol=1e-12)


def test_matrix_io_round_trip(tmp_path):
    """/root/reference/sucemfem/Utilities/tests/test_MatrixIO.py:40-75"""
    from sucemfem.Utilities.MatrixIO import load_scipy_matrix_from_mat, save_scipy_matrix_as_mat
    path = str(tmp_path / 'MatrixIO')
    A = scipy.sparse.rand(300, 300, density=0.02, format='csr', random_state=3)
    assert save_scipy_matrix_as_mat(path, 'A', A)
    L = load_scipy_matrix_from_mat(path, 'A')
    np.testing.assert_equal(L.shape, A.shape)
    np.testing.assert_equal(L.data, A.data)
    np.testing.assert_equal(L.indices, A.indices)
    np.testing.assert_equal(L.indptr, A.indptr)
    D = np.random.default_rng(4).random((7, 5))
    save_scipy_matrix_as_mat(path, 'D', D)
    np.testing.assert_equal(load_scipy_matrix_from_mat(path, 'D'), D)
    assert load_scipy_matrix_from_mat(path, 'missing') is None


def test_femmesh_reader_and_converters(tmp_path):
    """/root/reference/sucemfem/Utilities/MeshIO.py:38-117, MeshConverters.py:24-68: BLOCK nodes / tets with 1-based
    numbering -> mesh with UFC-ordered cells."""
    from sucemfem.Testing.Meshes import InscribedTet
    from sucemfem.Utilities import MeshIO, MeshConverters
    it = InscribedTet()
    f = tmp_path / 'inscribed.femmesh'
    rng = np.random.default_rng(5)
    with open(f, 'w') as fo:
        fo.write('BLOCK nodes\n%d\n' % len(it.coordinates))
        for i in rng.permutation(len(it.coordinates)):
            fo.write('%d %r %r %r\n' % ((i + 1,) + tuple(float(c) for c in it.coordinates[i])))
        fo.write('ENDBLOCK\nBLOCK tets\n%d\n' % len(it.element_nodes))
        for i, nodes in enumerate(it.element_nodes):
            fo.write('%d %d %s\n' % (i + 1, 7, ' '.join(str(int(v) + 1) for v in nodes[::-1])))
        fo.write('ENDBLOCK\nBLOCK tris\n0\nENDBLOCK\n')
    rd = MeshIO.FemmeshReader(str(f))
    try:
        rd.get_nodes()
        raise AssertionError('expected the not-read error')
    except Exception as e:
        assert 'read_meshfile' in str(e)
    mesh = MeshIO.femmesh_2_dolfin_mesh(str(f))
    np.testing.assert_array_equal(mesh.coordinates(), it.coordinates)
    np.testing.assert_array_equal(mesh.cells(), np.sort(it.element_nodes, axis=1))
    rd.read_meshfile()
    assert rd.get_mesh_filename() == 'inscribed.femmesh' and set(rd.get_tet_property_nos()) == {7}
    lm = MeshConverters.dolfin_mesh_2_listmesh(mesh)
    m2 = MeshConverters.listmesh_2_dolfin_mesh(lm)
    np.testing.assert_array_equal(m2.cells(), mesh.cells())
    np.testing.assert_array_equal(m2.coordinates(), mesh.coordinates())


def test_circuit_s11():
    from sucemfem.PostProcessing.circuit import S11
    assert S11(50.0, 50.0) == 0 and np.allclose(S11([0.0, 1e30], 50.0), [-1.0, 1.0])


class _HostSpace(object):
    """CPU stand-in for dolfin_compat.FunctionSpace with the attributes the host-side evaluators use (mesh(),
    dofmap().cell_dofs_array(), order); numbering from the oracle (this is a test of host logic only)."""

    def __init__(self, co, ce, order):
        from oracle import fem
        self._co, self._ce, self.order = co, ce, order
        self.num = fem.number_entities(ce)
        self._cd, self.ndof = fem.cell_dofs(self.num, order)

    def mesh(self):
        return self

    def coordinates(self):
        return self._co

    def cells(self):
        return self._ce

    def dofmap(self):
        return self

    def cell_dofs_array(self):
        return self._cd

    def geometry(self):
        return self

    def dim(self):
        return self._co.shape[1]

    # -- what the O(boundary) surface functionals ask of V.device_space (PostProcessing/_facets.py) ---------------
    @property
    def device_space(self):
        return _HostDeviceSpace(self)


class _HostDeviceSpace(object):
    def __init__(self, V):
        self._V = V
        self.n_facets = V.num['F']
        self.mesh = type('M', (), {'device': 'cpu'})()

    def cell_entities(self):
        import torch
        return torch.from_numpy(self._V.num['cell_edges']), torch.from_numpy(self._V.num['cell_faces'])

    def mark_exterior_facets(self, markers, value):
        import torch
        from oracle import fem
        markers[torch.from_numpy(fem.exterior_facet_mask(self._V._ce, self._V.num))] = value
        return markers


def _interpolate_order2(V, u):
    """Nodal dofs of a vector field u(x) in the order-2 N1curl space (SURVEY 8.1-B functionals)."""
    co, num = V._co, V.num
    x = np.zeros(V.ndof)
    E = num['E']
    for e, (a, b) in enumerate(num['edges']):
        t = co[b] - co[a]
        for k in range(2):
            x[2 * e + k] = u(co[a] + (k + 1) / 3.0 * t) @ t
    for f, (a, b, c) in enumerate(num['faces']):
        cen = (co[a] + co[b] + co[c]) / 3.0
        x[2 * E + 2 * f] = u(cen) @ (co[b] - co[a])
        x[2 * E + 2 * f + 1] = u(cen) @ (co[c] - co[a])
    return x


def test_voltage_along_line_reference_cases():
    """/root/reference/sucemfem/Sources/tests/test_postprocessing.py:30-77: E = (0, 0, 2z) on UnitCube(3), order 2;
    the reference's five paths give +1, +1, -1, -1 and 0 V.  (Its complex case interpolates -z^2 at order 3; here the
    imaginary part is -z, which the order-2 space holds exactly: -1/2.)"""
    from oracle import meshes
    from sucemfem.Sources.PostProcess import VoltageAlongLine, ComplexVoltageAlongLine, FieldFunction
    co, ce = meshes.unit_cube(3)
    V = _HostSpace(co, ce, 2)
    u_r = _interpolate_order2(V, lambda p: np.array([0, 0, 2 * p[2]]))
    u_i = _interpolate_order2(V, lambda p: np.array([0, 0, -p[2]]))
    p1v_pts = (np.array([[0.5, 0.35, 0], [0.5, 0.5, 1]]), np.array([[0, 0, 0], [1, 1, 1.]]))
    m1v_pts = (np.array([[0.25, 0.5, 1 - 3e-16], [0.5, 0.75, 3e-16]]), np.array([[1, 1, 1], [0, 0, 0.]]))
    p0v_pts = np.array([[0, 0, 1], [1, 1, 1.]])
    dut = VoltageAlongLine(FieldFunction(V, u_r))
    assert abs(dut.calculate_voltage(*p1v_pts[0]) - 1) < 1e-7 and abs(dut.calculate_voltage(*p1v_pts[1]) - 1) < 1e-7
    assert abs(dut.calculate_voltage(*m1v_pts[0]) + 1) < 1e-7 and abs(dut.calculate_voltage(*m1v_pts[1]) + 1) < 1e-7
    assert abs(dut.calculate_voltage(*p0v_pts)) < 1e-7
    cdut = ComplexVoltageAlongLine(V)
    cdut.set_dofs(u_r + 1j * u_i)
    for pts, want in ((p1v_pts[0], 1 - 0.5j), (p1v_pts[1], 1 - 0.5j), (m1v_pts[0], -(1 - 0.5j)), (m1v_pts[1], -(1 - 0.5j)),
                      (p0v_pts, 0)):
        assert abs(cdut.calculate_voltage(*pts) - want) < 1e-7


def test_small_example_helpers():
    """Testing.ErrorMeasures / Testing.Analytical / Testing.Paths / Utilities.Optimization: what the reference's example
    drivers import (examples/filament_source/driver.py:25-40)."""
    from sucemfem.Testing.ErrorMeasures import normalised_RMS, max_normalised_RMS
    from sucemfem.Testing.Analytical import current_fillament_farfield as ff
    from sucemfem.Testing import Paths
    from sucemfem.Utilities import Optimization
    from sucemfem.Consts import mu0, c0
    a, d = np.array([1.0, -2.0, 4.0]), np.array([1.1, -2.0, 3.0])
    assert np.isclose(normalised_RMS(a, d), 100 * (0.1 ** 2 + 0.25 ** 2))
    assert np.isclose(max_normalised_RMS(a, d, extra_error=np.array([0, 0.4, 0])), 100 * ((0.1 / 4) ** 2 + (1 / 4) ** 2 + 0.1 ** 2))
    # short dipole limit: |E_theta| -> w mu0 I L sin(theta) / (4 pi r); broadside value finite (the symbolic form is 0/0)
    freq, L, I = 1e9, 1e-4, 2.0
    th = np.deg2rad(np.array([30.0, 90.0, 140.0]))
    v = ff.eval_E_theta(freq, L, I, th, r_val=3.0)
    assert np.allclose(np.abs(v), 2 * np.pi * freq * mu0 * I * L * np.sin(th) / (4 * np.pi * 3.0), rtol=1e-6)
    lam = c0 / freq
    full = ff.eval_E_theta(freq, lam, 1.0, np.deg2rad(60.0))      # u = pi/2 * ... : pattern sin(60) sin(u)/u, u = pi cos(60)
    u = np.pi * np.cos(np.deg2rad(60.0))
    assert np.isclose(abs(full), 2 * np.pi * freq * mu0 * lam / (4 * np.pi) * np.sin(np.deg2rad(60.0)) * np.sin(u) / u)
    assert isinstance(ff.eval_E_theta(freq, lam / 4, 1.0, 0.3), complex)
    assert Paths.get_module_path_filename('x.dat', '/a/b/c.py') == '/a/b/x.dat'
    Optimization.set_dolfin_optimisation(False)
    assert Optimization.parameters['form_compiler'] == {'optimize': False, 'cpp_optimize': False}
    Optimization.set_dolfin_optimisation()


def test_sources_on_host_match_oracle_and_reference_fixture():
    """The product's sources (Sources.point_source / fillament_current_source: generated basis polynomials, covariant
    Piola map, lowest-cell point location) against the oracle's independent basis on a jittered mesh at orders 1 and 2,
    and against the reference's filament fixture (sucemfem/Sources/tests/test_fillament_source.py:62-69) -- on the CPU,
    with the dof map taken from the oracle instead of the device."""
    from oracle import fem, meshes
    from sucemfem.Sources.point_source import calc_pointsource_contrib
    from sucemfem.Sources.fillament_current_source import FillamentCurrentSource
    co, ce = meshes.unit_cube(3)
    rng = np.random.default_rng(2)
    co = co + 0.03 * rng.standard_normal(co.shape)
    for order in (1, 2):
        V = _HostSpace(co, ce, order)
        for pt in (np.array([0.31, 0.52, 0.47]), np.array([0.8, 0.15, 0.66])):
            val = np.array([0.3, -1.2, 2.0])
            dn, v = calc_pointsource_contrib(V, pt, val)
            dn_o, v_o = fem.point_source_contrib(co, ce, V._cd, order, pt, val)
            assert np.array_equal(dn, dn_o) and np.allclose(v, v_o, rtol=1e-12, atol=1e-14)
    # the reference's own filament fixture (order 2, get_centred_cube mesh)
    g = np.load(os.path.join(GOLDEN, 'fillament_source_order2.npz'))
    co, ce = meshes.centred_cube(g['domain_size'], float(g['max_edge_len']))
    V = _HostSpace(co, ce, int(g['order']))
    src = FillamentCurrentSource()
    src.set_function_space(V)
    src.set_no_integration_points(int(g['no_integration_points']))
    src.set_source_endpoints(np.asarray(g['source_endpoints']))
    src.set_value(float(g['I']))
    dn, v = src.get_contribution()
    big = np.abs(g['rhs_contribs']) > 1e-12        # the other 48 stored dofs carry round-off only
    got = dict(zip(np.asarray(dn).tolist(), v))
    assert big.sum() == 26
    assert all(np.isclose(got[int(d)], x, rtol=1e-10, atol=1e-12) for d, x in zip(g['dofnos'][big], g['rhs_contribs'][big]))


def test_surface_functionals_on_host_known_answers():
    """The O(boundary) host functionals on the CPU (numbering from the oracle instead of the device):
    SurfaceFlux = 26.3586862457 (/root/reference/sucemfem/PostProcessing/tests/test_power_flux.py:35-55) and the surface
    NTFF at the reference's six directions (test_ntff.py:80-110, rtol=1e-12, atol=1e-7)."""
    from oracle import meshes
    from sucemfem.PostProcessing import power_flux, surface_ntff
    c0 = 299792458
    g = np.load(os.path.join(GOLDEN, 'driven_order2_unitcube3.npz'))
    V = _HostSpace(g['mesh_coordinates'], g['mesh_cells'], int(g['order']))
    DUT = power_flux.SurfaceFlux(V)
    DUT.set_k0(2 * np.pi * float(g['freq']) / c0)
    DUT.set_dofs(g['x'])
    assert abs(DUT.calc_flux() - float(g['known_answer_surface_flux'])) < 5e-8
    g = np.load(os.path.join(GOLDEN, 'ntff_order2_unitcube5.npz'))
    x = np.load(os.path.join(GOLDEN, 'driven_order2_unitcube5.npz'))['x']
    co, ce = meshes.centred_cube(g['domain_size'], float(g['max_edge_len']))
    V = _HostSpace(co, ce, int(g['order']))
    DUT = surface_ntff.NTFF(V)
    DUT.set_frequency(float(g['freq']))
    DUT.set_dofs(x)
    actual = [DUT.calc_pt(th, ph) for th, ph in zip(g['surface_theta_deg'], g['surface_phi_deg'])]
    assert np.allclose(actual, g['surface_E_ff'], rtol=1e-12, atol=1e-7)
