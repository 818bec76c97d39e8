"""The bench-size oracle (oracle/rowcheck.py, oc_number_entities_hash, oc_assemble_rows) against the full oracle on
meshes the latter finishes in seconds: the O(C) hash numbering equals the sort-based one (and oracle/fem.py), sampled
rows are bit-identical to the rows of the fully assembled matrices, the geometric PEC rule equals the facet-closure
rule, the entity-level count pass gives the dof-level row pointer."""
import numpy as np
import pytest

from oracle import cport, fem, meshes, rowcheck


@pytest.mark.parametrize('n', [1, 2, 5, 8])
def test_hash_numbering_equals_sort_numbering(n):
    co, ce = meshes.unit_cube(n)
    A = cport.CPUAssembler(co, ce, 2)
    e, f, E, F = cport.number_entities_hash(ce, len(co))
    num = fem.number_entities(ce)
    assert (E, F) == (A.E, A.F) == (num['E'], num['F'])
    assert np.array_equal(e, A.cell_edges) and np.array_equal(f, A.cell_faces)
    assert np.array_equal(e, num['cell_edges']) and np.array_equal(f, num['cell_faces'])
    # tight table sizes (the closed forms) work too
    e2, f2, _, _ = cport.number_entities_hash(ce, len(co), 7 * n ** 3 + 9 * n * n + 3 * n, 12 * n ** 3 + 6 * n * n)
    assert np.array_equal(e2, e) and np.array_equal(f2, f)


def test_hash_numbering_unstructured():
    co, ce = meshes.inscribed_tet()
    e, f, E, F = cport.number_entities_hash(np.sort(ce, axis=1), len(co))
    num = fem.number_entities(np.sort(ce, axis=1))
    assert E == 24 and np.array_equal(e, num['cell_edges']) and np.array_equal(f, num['cell_faces'])


def test_sampled_rows_equal_full_assembly_and_geometric_pec():
    n = 6
    orc = rowcheck.CubeOracle(n)
    A = cport.CPUAssembler(orc.coords, orc.cells, 2)
    rng = np.random.default_rng(0)
    eps, mu = rng.uniform(1, 16, len(orc.cells)), rng.uniform(1, 2, len(orc.cells))
    S, M = A.assemble(eps, mu)
    num = fem.number_entities(orc.cells)
    pec_rows = fem.facet_closure_dofs(orc.cells, num, 2, fem.exterior_facet_mask(orc.cells, num))
    A.apply_ident(S, pec_rows)
    A.apply_ident(M, pec_rows)
    flag = np.zeros(A.ndof, np.uint8)
    flag[pec_rows] = 1
    assert np.array_equal(orc.pec_flags_of(np.arange(A.ndof)), flag)
    rows = rowcheck.pick_rows(lambda v: np.searchsorted(A.rowptr, v, side='right') - 1, A.ndof, len(A.cols), 300)
    want = orc.rows(rows, eps, mu, flag[rows])
    idx = np.concatenate([np.arange(A.rowptr[r], A.rowptr[r + 1]) for r in rows])
    info = rowcheck.compare(want, A.cols[idx], S[idx], M[idx])
    assert info['max_err_over_tol'] == 0.0 and info['rows'] == len(rows)
    with pytest.raises(AssertionError):
        bad = S[idx].copy()
        bad[len(bad) // 2] *= 1 + 1e-11
        rowcheck.compare(want, A.cols[idx], bad, M[idx])
    erp = orc.entity_rowptr(2)
    assert np.array_equal(4 * erp[:-1], A.rowptr[0:-1:2]) and 4 * erp[-1] == A.rowptr[-1]
    assert np.array_equal(4 * erp[:-1] + 2 * np.diff(erp), A.rowptr[1::2])
