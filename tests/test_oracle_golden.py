"""Pin the CPU oracle against the reference's own golden fixtures.

Tolerances are the reference's own where a reference test exists:
  fixture 1: rtol=1e-10, atol=3e-15  (sucemfem/ProblemConfigurations/tests/test_EMDrivenProblem.py:71-72)
  fixture 2: rtol=1e-12, atol=1e-16  (same file :81-82)
  fixture 3: rtol=1e-10, atol=1e-12  (sucemfem/Sources/tests/test_fillament_source.py:50-51,68-69)
Fixtures 5 and 7 are solution vectors of the order-2 driven problem; they pin S, M and S_0 at order 2
through the residual of the reference's own linear system.
"""
import os

import numpy as np
import scipy.sparse.linalg as spla

from oracle import fem, meshes
from conftest import GOLDEN

c0 = 299792458
mu0 = np.pi * 4 * 1e-7
Z0 = mu0 * c0


def _lhs(r, k0):
    n = r['ndof']
    mk = lambda k: fem.to_scipy(r['rowptr'], r['cols'], r[k], n)
    return mk('S') - k0 ** 2 * mk('M') + 1j * k0 * mk('S_0')


def test_fixture1_lhs_matrix_order1():
    g = np.load(os.path.join(GOLDEN, 'driven_order1_inscribedtet.npz'))
    co, ce = meshes.inscribed_tet()
    r = fem.assemble_system(co, ce, 1, abc=True)
    assert r['ndof'] == 24 and len(r['cols']) == 258
    k0 = 2 * np.pi * float(g['frequency']) / c0
    A = _lhs(r, k0).toarray()
    assert np.allclose(A, g['LHS'], rtol=1e-10, atol=3e-15)
    # every non-zero of the golden matrix lies inside the cell-block pattern
    pat = fem.to_scipy(r['rowptr'], r['cols'], np.ones(len(r['cols'])), 24).toarray() != 0
    assert np.all(pat[np.abs(g['LHS']) > 0])


def test_fixture2_rhs_vector_order1():
    g = np.load(os.path.join(GOLDEN, 'driven_order1_inscribedtet.npz'))
    co, ce = meshes.inscribed_tet()
    num = fem.number_entities(ce)
    cd, n = fem.cell_dofs(num, 1)
    k0 = 2 * np.pi * float(g['frequency']) / c0
    dn, v = fem.point_source_contrib(co, ce, cd, 1, g['source_coord'], g['source_value'])
    b = np.zeros(n, complex)
    b[dn] += -1j * k0 * Z0 * v
    assert np.allclose(b, g['RHS'], rtol=1e-12, atol=1e-16)


def test_fixture3_order2_dof_numbers_and_basis():
    g = np.load(os.path.join(GOLDEN, 'fillament_source_order2.npz'))
    co, ce = meshes.centred_cube(g['domain_size'], float(g['max_edge_len']))
    assert ce.shape[0] == 162
    num = fem.number_entities(ce)
    cd, n = fem.cell_dofs(num, int(g['order']))
    assert n == 1314
    dn, v = fem.filament_source_contrib(co, ce, cd, 2, g['source_endpoints'], float(g['I']),
                                        int(g['no_integration_points']))
    # 48 of the 74 stored dofs carry only round-off because sample points lie on a shared face and
    # DOLFIN/CGAL picked one of the two cells: compare the significant ones
    sig = np.abs(g['rhs_contribs']) > 1e-12
    mine = dict(zip(dn.tolist(), v))
    mine_sig = sorted(k for k, x in mine.items() if abs(x) > 1e-12)
    assert mine_sig == sorted(g['dofnos'][sig].tolist())          # exact dof numbers
    got = np.array([mine[k] for k in g['dofnos'][sig].tolist()])
    assert np.allclose(got, g['rhs_contribs'][sig], rtol=1e-10, atol=1e-12)


def test_fixture5_order2_driven_solution_unitcube3():
    g = np.load(os.path.join(GOLDEN, 'driven_order2_unitcube3.npz'))
    freq = float(g['freq'])
    lam = c0 / freq
    k0 = 2 * np.pi * freq / c0
    co, ce = meshes.centred_cube(np.array([lam] * 3) * 0.25, lam / 6)
    r = fem.assemble_system(co, ce, 2, abc=True)
    assert r['ndof'] == 1314 and len(r['cols']) == 47988
    A = _lhs(r, k0).tocsc()
    dn, v = fem.filament_source_contrib(co, ce, r['cell_dofs'], 2, g['source_endpoints'], float(g['I']), 100)
    b = np.zeros(r['ndof'], complex)
    b[dn] += -1j * k0 * Z0 * v
    x = spla.spsolve(A, b)
    assert np.linalg.norm(x - g['x']) / np.linalg.norm(g['x']) < 1e-11
    assert np.linalg.norm(A @ g['x'] - b) / np.linalg.norm(b) < 1e-11


def test_fixture7_order2_driven_residual_unitcube5():
    g = np.load(os.path.join(GOLDEN, 'driven_order2_unitcube5.npz'))
    freq = float(g['freq'])
    lam = c0 / freq
    k0 = 2 * np.pi * freq / c0
    co, ce = meshes.centred_cube(g['domain_size'], float(g['max_edge_len']))
    hits, _ = fem.locate_cells(co, ce, np.zeros(3))
    assert hits.tolist() == [372, 373, 374, 375, 376, 377]
    co = co - co[ce[hits[0]]].mean(axis=0)      # get_centred_cube recentring, MeshGenerators.py:37-52
    r = fem.assemble_system(co, ce, 2, abc=True)
    assert r['ndof'] == 5530
    A = _lhs(r, k0).tocsr()
    dn, v = fem.point_source_contrib(co, ce, r['cell_dofs'], 2, np.zeros(3), np.array([0, 0, 1.]) * lam / 1000,
                                     cell_index=hits[0])
    b = np.zeros(r['ndof'], complex)
    b[dn] += -1j * k0 * Z0 * v
    assert np.linalg.norm(A @ g['x'] - b) / np.linalg.norm(b) < 1e-11


def test_counts_unitcube():
    # SURVEY 8.1-M / BASELINE.md section 3 closed forms
    for n in (1, 2, 4):
        co, ce = meshes.unit_cube(n)
        num = fem.number_entities(ce)
        assert ce.shape[0] == 6 * n ** 3 and co.shape[0] == (n + 1) ** 3
        assert num['E'] == 7 * n ** 3 + 9 * n ** 2 + 3 * n and num['F'] == 12 * n ** 3 + 6 * n ** 2
        assert fem.exterior_facet_mask(ce, num).sum() == 12 * n ** 2
        for order, nnz in ((1, 115 * n ** 3 + 45 * n ** 2 + 3 * n), (2, 1660 * n ** 3 + 348 * n ** 2 + 12 * n)):
            cd, ndof = fem.cell_dofs(num, order)
            rp, cols = fem.pattern(cd, ndof)
            assert len(cols) == nnz


def test_pec_cavity_eigenvalues_order2():
    # SURVEY 8.3 anchor: UnitCube(4) order 2 PEC: 960 PEC dofs, 343-dim null space, lowest k/pi ~ sqrt(2) x3
    import scipy.linalg as sla
    co, ce = meshes.unit_cube(4)
    num = fem.number_entities(ce)
    r = fem.assemble_system(co, ce, 2, pec_facet_mask=fem.exterior_facet_mask(ce, num))
    assert r['ndof'] == 2936 and len(r['pec_dofs']) == 960
    r0 = fem.assemble_system(co, ce, 2)
    free = np.setdiff1d(np.arange(r['ndof']), r['pec_dofs'])
    S = fem.to_scipy(r0['rowptr'], r0['cols'], r0['S'], r['ndof']).toarray()[np.ix_(free, free)]
    M = fem.to_scipy(r0['rowptr'], r0['cols'], r0['M'], r['ndof']).toarray()[np.ix_(free, free)]
    w = sla.eigh(S, M, eigvals_only=True)
    assert np.sum(np.abs(w) < 1e-8) == 343
    k = np.sqrt(w[343:351]) / np.pi
    assert np.allclose(k[:3], [1.413965, 1.415131, 1.415131], atol=2e-6)
    # PEC rows of the ident-applied matrices are unit rows
    Sp = fem.to_scipy(r['rowptr'], r['cols'], r['S'], r['ndof']).toarray()
    assert np.all(Sp[r['pec_dofs']] == np.eye(r['ndof'])[r['pec_dofs']])


def test_fixture6_face_numbering_on_gmsh_prism():
    """The reference's facet-marker file for its gmsh prism (functional_tests/test_EigenProblems.py:79-84) is indexed
    by DOLFIN's face numbering: all 92 PEC markers must land exactly on the 92 boundary faces of OUR numbering."""
    g = np.load(os.path.join(GOLDEN, 'rectangular_prism_gmsh.npz'))
    co, ce = g['coordinates'], g['cells']
    assert np.all(np.diff(ce, axis=1) > 0)
    num = fem.number_entities(ce)
    assert num['F'] == len(g['facet_markers']) == 362
    assert np.array_equal(fem.exterior_facet_mask(ce, num), g['facet_markers'] == 1)
    assert (g['facet_markers'] == 1).sum() == 92


def test_fixture8_order2_filament_example_unitcube9():
    """examples/filament_source/data/dofs-2-...pickle (driver.py:48-98): 30186 dofs on UnitCube(9), lambda/4 filament,
    ABC.  The shipped vector is an ITERATIVE solution (bicgstab, scipy's default rtol 1e-5, driver.py:93-94): with the
    oracle's S, M, S_0 and right-hand side its residual is 9.97e-6, just under that tolerance -- i.e. the oracle's
    system is the one the reference solved, as far as this fixture can tell."""
    g = np.load(os.path.join(GOLDEN, 'driven_order2_unitcube9_bicgstab.npz'))
    freq = float(g['freq'])
    k0 = 2 * np.pi * freq / c0
    co, ce = meshes.centred_cube(g['domain_size'], float(g['max_edge_len']))
    assert len(ce) == 6 * 9 ** 3
    r = fem.assemble_system(co, ce, 2, abc=True)
    assert r['ndof'] == len(g['x']) == 30186
    A = _lhs(r, k0).tocsr()
    l = float(g['l'])
    ends = np.array([[0, 0, -l / 2], [0, 0, l / 2]])
    dn, v = fem.filament_source_contrib(co, ce, r['cell_dofs'], 2, ends, float(g['I']), int(g['source_integration_points']))
    b = np.zeros(r['ndof'], complex)
    b[dn] += -1j * k0 * Z0 * v
    res = np.linalg.norm(A @ g['x'] - b) / np.linalg.norm(b)
    assert res < 1.0e-5
    # sensitivity of the check: a 0.1 % error in S_0 alone would be seen
    A_bad = A + 1e-3 * 1j * k0 * fem.to_scipy(r['rowptr'], r['cols'], r['S_0'], r['ndof'])
    assert np.linalg.norm(A_bad @ g['x'] - b) / np.linalg.norm(b) > 2 * res
