"""Parity AT THE BENCHMARKED SIZES: the configurations bench.py times have more than 2^31 non-zeros
per matrix (UnitCube(119): 2.80e9; the scaling runs go to UnitCube(203) with global offsets above 2^33), where an
index-width bug would not show in any smaller test.

  * UnitCube(119), one GPU, the pattern of all rows (the BENCH line's workload):
      - entity numbering bit-exact against the oracle's O(C) first-appearance numbering (all 10.1 M cells);
      - the whole row pointer bit-exact against the oracle's count pass;
      - ~1e5 sampled rows -- head, tail, 16 rows around EVERY multiple of 2^28 in the value index (so around 2^31 and
        every 2 GiB byte offset) and random ones -- against the reference algorithm restricted to those rows
        (oracle/cpu_assemble.c: oc_assemble_rows): columns bit-exact, S and T to 1e-12 relative, PEC unit rows included
        (which rows are PEC is decided geometrically by the oracle, not taken from the device);
      - the size-independent properties of tests/test_gpu_fullsize.py through the product's own 64-bit SpMV;
      - bit-reproducibility of two assemblies.
  * UnitCube(140): 4.56e9 non-zeros, above 2^32: sampled rows around every multiple of 2^28 up to the end.
  * UnitCube(203) = BASELINE config 5, the two row blocks (edge range + face range) of ranks 0, 3 and 7 of the 8-way
    partition bench.py times at 8 GPUs: sampled rows of each block against the oracle.
Contract: /root/reference/sucemfem/ProblemConfigurations/tests/test_EMDrivenProblem.py:64-72 (allclose on the assembled
matrix), per sampled row.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free():
    import gc
    import torch
    from sucemfem import _native as N
    gc.collect()
    torch.cuda.empty_cache()
    N.check(N.lib.sfem_release_cached_memory())


def _coefficients(n_cells, seed=1234):
    """Per-cell eps_r ~ U[1,16), mu_r ~ U[1,2) on the host (SURVEY 8.2 row d), shared by both sides."""
    rng = np.random.default_rng(seed)
    return rng.uniform(1, 16, n_cells), rng.uniform(1, 2, n_cells)


def _check_block(orc, space, pat, eps_d, mu_d, flags, eps, mu, n_random, label):
    """Assemble the block on the device and compare a sample of its rows with the oracle."""
    from oracle import rowcheck
    S, T = pat.assemble_ST(eps_d, mu_d, flags)
    rowptr, cols = pat.csr()
    info = rowcheck.check_block(orc, rowptr, cols, S, T, pat.row_begin, pat.n_rows, pat.nnz, flags, eps, mu, n_random, label)
    return S, T, rowptr, cols, info


def _pec_flags(space):
    import torch
    markers = torch.zeros(space.n_facets, dtype=torch.int32, device='cuda')
    space.mark_exterior_facets(markers, 999)
    return space.pec_row_flags(markers, 999)


def test_unitcube119_all_rows_vs_oracle():
    """The BENCH workload: 10.1 M tets, 64.5 M dofs, 2.80e9 non-zeros per matrix (> 2^31)."""
    import torch
    from oracle import rowcheck
    from sucemfem import device
    n = 119
    orc = rowcheck.CubeOracle(n)
    mesh = device.DeviceMesh.unit(n, n, n)
    assert np.array_equal(mesh.cells.cpu().numpy(), orc.cells) and np.array_equal(mesh.coords.cpu().numpy(), orc.coords)
    space = device.DeviceSpace(mesh, 2)
    assert (space.n_edges, space.n_faces, space.ndof) == (orc.E, orc.F, 64461586)
    ce, cf = space.cell_entities()
    assert np.array_equal(ce.cpu().numpy(), orc.cell_edges) and np.array_equal(cf.cpu().numpy(), orc.cell_faces)
    del ce, cf
    pat = space.pattern()
    assert pat.nnz == 1660 * n ** 3 + 348 * n ** 2 + 12 * n == 2802293396 > 2 ** 31
    eps, mu = _coefficients(mesh.num_cells)
    eps_d, mu_d = torch.from_numpy(eps).cuda(), torch.from_numpy(mu).cuda()
    flags = _pec_flags(space)
    S, T, rowptr, cols, info = _check_block(orc, space, pat, eps_d, mu_d, flags, eps, mu, 50000, 'UnitCube(119)')
    assert info['rows'] > 9e4
    # the whole row pointer against the oracle's count pass (entity level -> dof level)
    erp = orc.entity_rowptr()
    ln = np.diff(erp)
    want_rp = np.empty(orc.ndof + 1, np.int64)
    want_rp[0:-1:2] = 4 * erp[:-1]
    want_rp[1::2] = 4 * erp[:-1] + 2 * ln
    want_rp[-1] = 4 * erp[-1]
    assert np.array_equal(rowptr.cpu().numpy(), want_rp)
    del erp, ln, want_rp
    # bit-reproducible at this size (no atomics anywhere on the path)
    S2, T2 = pat.assemble_ST(eps_d, mu_d, flags)
    assert torch.equal(S2, S) and torch.equal(T2, T)
    del S2, T2, cols
    # size-independent properties through the product's SpMV (64-bit offsets): un-constrained matrices
    S, T = pat.assemble_ST(eps_d, mu_d, None, S, T)
    x = _constant_field_dofs(mesh, space, [0.3, -1.1, 0.7])
    r = pat.matvec([S], x)
    assert float(r.abs().max()) <= 1e-12 * float(S.abs().max()) * float(x.abs().max())
    got = float(x @ pat.matvec([T], x))
    want = (0.3 ** 2 + 1.1 ** 2 + 0.7 ** 2) * float(eps.sum()) / (6 * n ** 3)
    assert abs(got - want) <= 1e-11 * abs(want)
    g = torch.Generator(device='cuda')
    g.manual_seed(3)
    u = torch.rand(space.ndof, dtype=torch.float64, device='cuda', generator=g) - 0.5
    v = torch.rand(space.ndof, dtype=torch.float64, device='cuda', generator=g) - 0.5
    for A in (S, T):
        a, b = float(v @ pat.matvec([A], u)), float(u @ pat.matvec([A], v))
        assert abs(a - b) <= 1e-10 * max(abs(a), abs(b), float(A.abs().max()))
    del S, T, pat, space, mesh, orc
    _free()


def _constant_field_dofs(mesh, space, c):
    """Interpolant of the constant field c: dof = c . tangent of its functional."""
    import torch
    ce, cf = space.cell_entities()
    x = mesh.coords
    cells = mesh.cells.long()
    c = torch.tensor(c, dtype=torch.float64, device='cuda')
    dofs = torch.zeros(space.ndof, dtype=torch.float64, device='cuda')
    E = space.n_edges
    for le, (a, b) in enumerate([(2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1)]):
        val = (x[cells[:, b]] - x[cells[:, a]]) @ c
        e = ce[:, le].long()
        dofs[2 * e] = val
        dofs[2 * e + 1] = val
    for lf, (a, b, d) in enumerate([(1, 2, 3), (0, 2, 3), (0, 1, 3), (0, 1, 2)]):
        f = cf[:, lf].long()
        dofs[2 * E + 2 * f] = (x[cells[:, b]] - x[cells[:, a]]) @ c
        dofs[2 * E + 2 * f + 1] = (x[cells[:, d]] - x[cells[:, a]]) @ c
    return dofs


def test_unitcube140_rows_beyond_2_to_32():
    """16.5 M tets, 4.56e9 non-zeros per matrix: value indices above 2^32 inside one pattern."""
    import torch
    from oracle import rowcheck
    from sucemfem import device
    n = 140
    orc = rowcheck.CubeOracle(n)
    mesh = device.DeviceMesh.unit(n, n, n)
    space = device.DeviceSpace(mesh, 2)
    pat = space.pattern()
    assert pat.nnz == 1660 * n ** 3 + 348 * n ** 2 + 12 * n > 2 ** 32
    eps, mu = _coefficients(mesh.num_cells, 99)
    S, T, rowptr, cols, info = _check_block(orc, space, pat, torch.from_numpy(eps).cuda(), torch.from_numpy(mu).cuda(),
                                            _pec_flags(space), eps, mu, 20000, 'UnitCube(140)')
    assert info['rows'] > 3e4
    del S, T, rowptr, cols, pat, space, mesh, orc
    _free()


def test_unitcube203_rank_blocks_of_the_8_way_partition():
    """BASELINE config 5 (50.2 M tets, 1.39e10 non-zeros per matrix) as one rank of the 8-GPU run sees it: the edge
    block and the face block of ranks 0, 3 and 7 (global dof ids up to 3.2e8, incidence offsets up to 5e8)."""
    import torch
    from oracle import rowcheck
    from sucemfem import device
    n = 203
    orc = rowcheck.CubeOracle(n)
    mesh = device.DeviceMesh.unit(n, n, n)
    space = device.DeviceSpace(mesh, 2)
    assert (space.n_edges, space.n_faces, space.ndof) == (58929479, 100632378, 319123714) == (orc.E, orc.F, orc.ndof)
    eps, mu = _coefficients(mesh.num_cells, 5)
    eps_d, mu_d = torch.from_numpy(eps).cuda(), torch.from_numpy(mu).cuda()
    flags = _pec_flags(space)
    eb, ob = space.partition_by_class(8)
    assert eb[0] == 0 and eb[-1] == space.n_edges == ob[0] and ob[-1] == space.n_entities
    total = 0
    for rank in (0, 3, 7):
        for lo, hi in ((eb[rank], eb[rank + 1]), (ob[rank], ob[rank + 1])):
            pat = device.DevicePattern(space, lo, hi)
            _, _, _, _, info = _check_block(orc, space, pat, eps_d, mu_d, flags, eps, mu, 6000,
                                            'UnitCube(203) rank %d entities [%d, %d)' % (rank, lo, hi))
            total += info['rows']
            del pat
            _free()
    assert total > 6e4
    del space, mesh, orc
    _free()
