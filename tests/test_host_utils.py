"""Host-side mirrors around the path that need no GPU: the solver classes on scipy matrices, sources and surface
functionals.  Bodies follow the reference's tests (file:line in the docstrings)."""
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
    # all three solver classes on the same system (GMRES hands its callback a scalar residual norm)
    from sucemfem.Utilities.LinalgSolvers import GMRESSolver
    for pre in (None, 'diagonal', 'ilu'):
        g = GMRESSolver(B, pre)
        xg = g.solve(rhs)
        assert g.info == 0 and np.linalg.norm(B @ xg - rhs) <= 1e-4 * np.linalg.norm(rhs)
        its, res = g.plot_convergence()
        assert len(its) >= 1 and np.all(np.isfinite(res)) and np.all(np.diff(its) == 1)


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
