"""GPU parity: libsucem_b200.so (through the C ABI / ctypes) against the CPU oracle on the same inputs.

Bar (BASELINE.json north_star): DOF maps, orientation (implicit: low->high) and the CSR pattern bit-exact;
matrix values to a relative entrywise tolerance of 1e-12 in fp64.  Entries of the reference's own golden
matrix that are pure round-off (|a| < 1e-14 at max|A| ~ 13) make a pure relative test meaningless, so the
test carries the absolute floor the reference's own test uses (atol = 3e-15 at max|A| ~ 13, i.e.
~2.3e-16 * max|A|; we allow 1e-15 * max|A|):   |a - b| <= 1e-12 |b| + 1e-15 max|B|.
"""
import numpy as np
import pytest

from oracle import fem, meshes

pytestmark = pytest.mark.gpu

RTOL, FLOOR = 1e-12, 1e-15


def close(a, b, scale=None, floor=FLOOR):
    """scale: magnitude the absolute floor refers to (default max|b|); for PEC-applied matrices pass the
    un-constrained matrix so that round-off entries are judged against the entries they came from."""
    a, b = np.asarray(a), np.asarray(b)
    scale = np.abs(b).max() if scale is None else max(np.abs(scale).max(), np.abs(b).max())
    return np.all(np.abs(a - b) <= RTOL * np.abs(b) + floor * scale)


def _meshes():
    rng = np.random.default_rng(7)
    out = {'inscribed_tet': meshes.inscribed_tet(), 'cube2': meshes.unit_cube(2), 'cube3x2x4': meshes.unit_cube(3, 2, 4),
           'square3': meshes.unit_square(3), 'square4x2': meshes.unit_square(4, 2)}
    co, ce = meshes.unit_cube(3)
    out['cube3_jitter'] = (co + 0.05 * rng.standard_normal(co.shape), ce)
    co, ce = meshes.unit_square(5)
    out['square5_jitter'] = (co + 0.02 * rng.standard_normal(co.shape), ce)
    co, ce = meshes.centred_cube(0.0749, 0.05)
    out['centred'] = (co, ce)
    # degenerate inputs: one cell; two cells that share nothing, with vertex ids no cell uses in between; an open
    # fan of tets around one edge and the closed ring (an interior edge whose row holds every entity of the mesh)
    out['single_tet'] = (np.array([[0, 0, 0], [1, .1, 0], [.2, 1, 0], [.1, .3, 1.0]]), np.array([[0, 1, 2, 3]], np.int32))
    out['single_tri'] = (np.array([[0, 0], [1, .2], [.3, 1.0]]), np.array([[0, 1, 2]], np.int32))
    unit = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1.0]])
    co = rng.standard_normal((12, 3))                       # vertices 1, 3, 4, 7 belong to no cell
    co[[0, 2, 5, 6]] = 0.7 * unit + 0.05 * rng.standard_normal((4, 3))
    co[[8, 9, 10, 11]] = 3.0 + 1.3 * unit[[1, 0, 2, 3]] + 0.05 * rng.standard_normal((4, 3))   # detJ < 0
    out['two_tets_apart'] = (co, np.array([[0, 2, 5, 6], [8, 9, 10, 11]], np.int32))
    ring = np.array([[np.cos(t), np.sin(t), 0.1 * np.sin(3 * t)] for t in 2 * np.pi * np.arange(5) / 5])
    co = np.vstack([[0, 0, -1.0], [0, 0, 1.0], ring]) + 0.03 * rng.standard_normal((7, 3))
    fan = np.array([[0, 1, 2 + k, 3 + k] for k in range(4)], np.int32)
    out['tet_fan_open'] = (co, fan)
    out['tet_ring_closed'] = (co, np.vstack([fan, [[0, 1, 2, 6]]]).astype(np.int32))
    return out


MESHES = _meshes()


@pytest.fixture(scope='module')
def dev():
    from sucemfem import device
    return device


@pytest.mark.parametrize('name', sorted(MESHES))
@pytest.mark.parametrize('order', [1, 2])
def test_numbering_dofmap_pattern_bit_exact(dev, name, order):
    co, ce = MESHES[name]
    num = fem.number_entities(ce)
    cd, ndof = fem.cell_dofs(num, order)
    rowptr, cols = fem.pattern(cd, ndof)
    m = dev.DeviceMesh(co, ce)
    sp = dev.DeviceSpace(m, order)
    assert sp.ndof == ndof and sp.n_edges == num['E'] and sp.n_faces == num.get('F', 0)
    cedges, cfaces = sp.cell_entities()
    assert np.array_equal(cedges.cpu().numpy(), num['cell_edges'])
    if cfaces is not None:
        assert np.array_equal(cfaces.cpu().numpy(), num['cell_faces'])
    assert np.array_equal(sp.cell_dofs().cpu().numpy(), cd)
    pat = sp.pattern()
    rp, cl = pat.csr()
    assert pat.nnz == len(cols)
    assert np.array_equal(rp.cpu().numpy(), rowptr)
    assert np.array_equal(cl.cpu().numpy(), cols)


@pytest.mark.parametrize('name', sorted(MESHES))
@pytest.mark.parametrize('order', [1, 2])
def test_values_generic_vs_oracle(dev, name, order):
    import torch
    from sucemfem import _native as N
    co, ce = MESHES[name]
    rng = np.random.default_rng(1234)
    eps = rng.uniform(1, 16, ce.shape[0])
    mu = rng.uniform(1, 2, ce.shape[0])
    num = fem.number_entities(ce)
    pecmask = fem.exterior_facet_mask(ce, num)
    ref = fem.assemble_system(co, ce, order, eps_r=eps, mu_r=mu, abc=True)
    refp = fem.assemble_system(co, ce, order, eps_r=eps, mu_r=mu, abc=True, pec_facet_mask=pecmask)
    m = dev.DeviceMesh(co, ce)
    sp = dev.DeviceSpace(m, order)
    pat = sp.pattern()
    d_eps, d_mu = torch.from_numpy(eps).cuda(), torch.from_numpy(mu).cuda()
    M = pat.assemble(N.FORM_MASS, coef=d_eps, kernel=N.KERNEL_GENERIC).cpu().numpy()
    S = pat.assemble(N.FORM_STIFFNESS, coef=d_mu, kernel=N.KERNEL_GENERIC).cpu().numpy()
    S0 = pat.assemble(N.FORM_ABC, kernel=N.KERNEL_GENERIC).cpu().numpy()
    assert close(M, ref['M']) and close(S, ref['S'])
    # S_0 contracts the tangential projector (J^-1 J^-T - nu nu^T), whose entries cancel to zero for basis
    # functions without tangential trace: round-off there is eps*|J^-1|^2*meas, not eps*max|S_0|
    assert close(S0, ref['S_0'], floor=2e-14)
    S2, T2 = pat.assemble_ST(d_eps, d_mu, kernel=N.KERNEL_GENERIC)
    assert np.array_equal(S2.cpu().numpy(), S) and np.array_equal(T2.cpu().numpy(), M)
    # PEC: exterior facets marked on the device, closure dofs, row-only ident (fused and as a separate step)
    markers = torch.zeros(sp.n_facets, dtype=torch.int32, device='cuda')
    sp.mark_exterior_facets(markers, 999)
    assert np.array_equal(markers.cpu().numpy() == 999, pecmask)
    flags = sp.pec_row_flags(markers, 999)
    assert np.array_equal(np.nonzero(flags.cpu().numpy())[0], refp['pec_dofs'])
    Sp, Tp = pat.assemble_ST(d_eps, d_mu, pec_flags=flags, kernel=N.KERNEL_GENERIC)
    assert close(Sp.cpu().numpy(), refp['S'], ref['S']) and close(Tp.cpu().numpy(), refp['M'], ref['M'])
    S0p = pat.apply_ident(flags, pat.assemble(N.FORM_ABC, kernel=N.KERNEL_GENERIC)).cpu().numpy()
    assert close(S0p, refp['S_0'], ref['S_0'], floor=2e-14)
    Sp2 = pat.apply_ident(flags, torch.from_numpy(S).cuda()).cpu().numpy()
    assert np.array_equal(Sp2, Sp.cpu().numpy())


def test_row_blocks_concatenate_to_full(dev):
    """Multi-GPU contract: blocks of rows assembled independently equal the P=1 result bit for bit."""
    from sucemfem import _native as N
    co, ce = meshes.unit_cube(3)
    m = dev.DeviceMesh(co, ce)
    sp = dev.DeviceSpace(m, 2)
    full = sp.pattern()
    S, T = full.assemble_ST(kernel=N.KERNEL_GENERIC)
    rp, cl = full.csr()
    for P in (2, 3, 8):
        b = sp.partition(P)
        assert b[0] == 0 and b[-1] == sp.n_entities and all(x <= y for x, y in zip(b, b[1:]))
        Ss, cs, off = [], [], 0
        for p in range(P):
            blk = dev.DevicePattern(sp, b[p], b[p + 1])
            assert blk.row_begin == 2 * b[p]
            s, t = blk.assemble_ST(kernel=N.KERNEL_GENERIC)
            brp, bcl = blk.csr()
            assert np.array_equal(brp.cpu().numpy() + off, rp.cpu().numpy()[blk.row_begin:blk.row_begin + blk.n_rows + 1])
            off += blk.nnz
            Ss.append(s.cpu().numpy()); cs.append(bcl.cpu().numpy())
        assert np.array_equal(np.concatenate(Ss), S.cpu().numpy())
        assert np.array_equal(np.concatenate(cs), cl.cpu().numpy())


def test_two_class_row_blocks_concatenate_to_full(dev):
    """The bench / multi-GPU partition: rank p owns one edge range and one face range (sfem_partition_rows_by_class).
    Edge blocks then face blocks, in rank order, are the full matrix bit for bit (canonical kernels)."""
    import torch
    from sucemfem import _native as N
    co, ce = meshes.unit_cube(4)
    m = dev.DeviceMesh(co, ce)
    sp = dev.DeviceSpace(m, 2)
    full = sp.pattern()
    S, T = full.assemble_ST(kernel=N.KERNEL_CANONICAL)
    rp, cl = full.csr()
    for P in (2, 5, 8):
        eb, ob = sp.partition_by_class(P)
        assert eb[0] == 0 and eb[-1] == sp.n_edges == ob[0] and ob[-1] == sp.n_entities
        assert all(x <= y for x, y in zip(eb, eb[1:])) and all(x <= y for x, y in zip(ob, ob[1:]))
        Ss, Ts, cs = [], [], []
        for b in (eb, ob):
            for p in range(P):
                blk = dev.DevicePattern(sp, b[p], b[p + 1])
                s, t = blk.assemble_ST(kernel=N.KERNEL_CANONICAL)
                Ss.append(s); Ts.append(t); cs.append(blk.csr()[1])
        assert torch.equal(torch.cat(Ss), S) and torch.equal(torch.cat(Ts), T) and torch.equal(torch.cat(cs), cl)
    # order 1 has no second class: the face ranges are empty
    sp1 = dev.DeviceSpace(m, 1)
    eb, ob = sp1.partition_by_class(3)
    assert eb[-1] == sp1.n_entities and ob == [sp1.n_entities] * 4


@pytest.mark.parametrize('name,order', [('cube3_jitter', 2), ('cube3x2x4', 1), ('square4x2', 2)])
def test_streamed_host_output_equals_device_matrices(dev, name, order):
    """sfem_host_job_* / sfem_space_assemble_to_host: host arrays in, host CSR out, streamed in row blocks; every block
    count gives the device path's S, T, S_0, rowptr and cols bit for bit (PEC from markers and 'exterior')."""
    import torch
    from sucemfem import _native as N
    co, ce = MESHES[name]
    rng = np.random.default_rng(11)
    eps, mu = rng.uniform(1, 16, len(ce)), rng.uniform(1, 2, len(ce))
    m = dev.DeviceMesh(co, ce)
    sp = dev.DeviceSpace(m, order)
    pat = sp.pattern()
    markers = torch.zeros(sp.n_facets, dtype=torch.int32, device='cuda')
    sp.mark_exterior_facets(markers, 999)
    flags = sp.pec_row_flags(markers, 999)
    d_eps, d_mu = torch.from_numpy(eps).cuda(), torch.from_numpy(mu).cuda()
    S, T = pat.assemble_ST(d_eps, d_mu, flags)
    S0 = pat.assemble(N.FORM_ABC, pec_flags=flags)
    rp, cl = pat.csr()
    want = {'rowptr': rp, 'cols': cl, 'S': S, 'T': T, 'S0': S0}
    assert sp.count_nnz() == pat.nnz
    for nb in (0, 1, 2, 7):
        out = sp.assemble_to_host(d_eps, d_mu, flags, want=('cols', 'S', 'T', 'S0'), n_blocks=nb)
        for k, v in want.items():
            assert out[k].is_pinned() and torch.equal(out[k], v.cpu()), (nb, k)
    # host arrays in: PEC walls by mode ('exterior') and by an explicit marker array
    for pec in ('exterior', (markers.cpu().numpy().astype(np.uint32), 999)):
        job = dev.HostAssemblyJob(co, ce, order, eps, mu, pec=pec)
        assert (job.ndof, job.nnz) == (sp.ndof, pat.nnz)
        out = job.run(want=('cols', 'S', 'T'), n_blocks=3)
        for k in ('rowptr', 'cols', 'S', 'T'):
            assert torch.equal(out[k], want[k].cpu()), k
        job.close()
    # multi-GPU use: a rank streams only its edge range and its face range; outputs are relative to the range
    job = dev.HostAssemblyJob(co, ce, order, eps, mu, pec='exterior')
    eb, ob = job.partition_by_class(3)
    assert (eb, ob) == sp.partition_by_class(3)
    Ss, cs, rps = [], [], []
    for b in (eb, ob):
        for r in range(3):
            if b[r + 1] == b[r]:
                continue
            o = job.run_rows(b[r], b[r + 1], want=('cols', 'S', 'T'), n_blocks=2)
            assert o['S'].numel() == job.rows_nnz(b[r], b[r + 1]) and int(o['rowptr'][0]) == 0
            Ss.append(o['S'].clone()); cs.append(o['cols'].clone()); rps.append(o['rowptr'].clone())
    assert torch.equal(torch.cat(Ss), S.cpu()) and torch.equal(torch.cat(cs), cl.cpu())
    full_rp, offs = [torch.zeros(1, dtype=torch.int64)], 0
    for t in rps:
        full_rp.append(t[1:] + offs)
        offs += int(t[-1])
    assert torch.equal(torch.cat(full_rp), rp.cpu())
    job.close()
    # no essential condition, default coefficients, only one matrix requested
    T1 = pat.assemble(N.FORM_MASS)       # a single matrix goes through the single-form kernel
    out = dev.assemble_host(co, ce, order, want=('T',), n_blocks=2)
    assert torch.equal(out['T'], T1.cpu()) and torch.equal(out['rowptr'], rp.cpu())


def test_host_job_rejects_bad_input(dev):
    """Error behaviour of the host-buffer entry points: message through sfem_last_error, no partial handle."""
    from sucemfem import _native as N
    co, ce = meshes.unit_cube(2)
    bad = ce.copy()
    bad[3, [0, 1]] = bad[3, [1, 0]]                      # not in UFC (ascending) order
    with pytest.raises(N.NativeError, match='ascending'):
        dev.HostAssemblyJob(co, bad, 2)
    with pytest.raises(N.NativeError, match='order 1 and 2'):
        dev.HostAssemblyJob(co, ce, 3)
    job = dev.HostAssemblyJob(co, ce, 1)
    with pytest.raises(N.NativeError, match='out of bounds'):
        job.run_rows(5, 10 ** 9, order=1)
    job.close()
    job.close()                                          # idempotent


def test_material_expansion(dev):
    import torch
    reg = torch.tensor([0, 1000, 1001, 1000, 0], dtype=torch.int32, device='cuda')
    out = dev.expand_material(reg, {0: 1.0, 1000: 16.0, 1001: 2.5}).cpu().numpy()
    assert out.tolist() == [1.0, 16.0, 2.5, 16.0, 1.0]
    with pytest.raises(ValueError, match='Material number 1001 not found'):
        dev.expand_material(reg, {0: 1.0, 1000: 16.0})


def test_unit_mesh_fill_matches_oracle(dev):
    for (nx, ny, nz) in ((2, 3, 4), (3, 3, 3)):
        co, ce = meshes.unit_cube(nx, ny, nz)
        m = dev.DeviceMesh.unit(nx, ny, nz)
        assert np.array_equal(m.cells.cpu().numpy(), ce)
        assert np.array_equal(m.coords.cpu().numpy(), co)
    co, ce = meshes.unit_square(4, 3)
    m = dev.DeviceMesh.unit(4, 3, 0)
    assert np.array_equal(m.cells.cpu().numpy(), ce) and np.array_equal(m.coords.cpu().numpy(), co)


def test_unsorted_cells_rejected(dev):
    import torch
    from sucemfem._native import NativeError
    co, ce = meshes.unit_cube(1)
    bad = torch.from_numpy(ce[:, ::-1].copy()).cuda()
    m = dev.DeviceMesh(torch.from_numpy(co).cuda(), bad)
    with pytest.raises(NativeError, match='ascending'):
        dev.DeviceSpace(m, 1)


TET_MESHES = [k for k in sorted(MESHES) if MESHES[k][0].shape[1] == 3]


@pytest.mark.parametrize('order', [1, 2])
@pytest.mark.parametrize('name', TET_MESHES)
def test_values_canonical_vs_oracle(dev, name, order):
    """The benchmarked kernels (canonical frame, tets of order 2 and 1) against the oracle and the generic kernel."""
    import torch
    from sucemfem import _native as N
    co, ce = MESHES[name]
    rng = np.random.default_rng(99)
    eps = rng.uniform(1, 16, ce.shape[0])
    mu = rng.uniform(1, 2, ce.shape[0])
    num = fem.number_entities(ce)
    pecmask = fem.exterior_facet_mask(ce, num)
    ref = fem.assemble_system(co, ce, order, eps_r=eps, mu_r=mu)
    refp = fem.assemble_system(co, ce, order, eps_r=eps, mu_r=mu, pec_facet_mask=pecmask)
    m = dev.DeviceMesh(co, ce)
    sp = dev.DeviceSpace(m, order)
    pat = sp.pattern()
    d_eps, d_mu = torch.from_numpy(eps).cuda(), torch.from_numpy(mu).cuda()
    S, T = pat.assemble_ST(d_eps, d_mu, kernel=N.KERNEL_CANONICAL)
    assert close(S.cpu().numpy(), ref['S']) and close(T.cpu().numpy(), ref['M'])
    Sg, Tg = pat.assemble_ST(d_eps, d_mu, kernel=N.KERNEL_GENERIC)
    assert close(S.cpu().numpy(), Sg.cpu().numpy()) and close(T.cpu().numpy(), Tg.cpu().numpy())
    # homogeneous default coefficients (NULL pointers) and fused PEC
    S1, T1 = pat.assemble_ST(kernel=N.KERNEL_CANONICAL)
    ref1 = fem.assemble_system(co, ce, order)
    assert close(S1.cpu().numpy(), ref1['S']) and close(T1.cpu().numpy(), ref1['M'])
    markers = torch.zeros(sp.n_facets, dtype=torch.int32, device='cuda')
    sp.mark_exterior_facets(markers, 999)
    flags = sp.pec_row_flags(markers, 999)
    Sp, Tp = pat.assemble_ST(d_eps, d_mu, pec_flags=flags, kernel=N.KERNEL_CANONICAL)
    assert close(Sp.cpu().numpy(), refp['S'], ref['S']) and close(Tp.cpu().numpy(), refp['M'], ref['M'])
    # bit-reproducible: no atomics, fixed summation order
    S2, T2 = pat.assemble_ST(d_eps, d_mu, pec_flags=flags, kernel=N.KERNEL_CANONICAL)
    assert torch.equal(S2, Sp) and torch.equal(T2, Tp)
    # row blocks (multi-GPU contract) are bit-identical slices of the full result
    b = sp.partition(3)
    Ss = []
    for q in range(3):
        blk = dev.DevicePattern(sp, b[q], b[q + 1])
        s, t = blk.assemble_ST(d_eps, d_mu, pec_flags=flags, kernel=N.KERNEL_CANONICAL)
        Ss.append(s)
    assert torch.equal(torch.cat(Ss), Sp)


@pytest.mark.parametrize('order', [1, 2])
def test_every_value_slot_is_written(dev, order):
    """The value kernels are store-only (no zero-fill of S / T beforehand): every CSR slot must be written by a kernel.
    Outputs pre-filled with NaN must come back NaN-free and equal to an assembly into fresh arrays -- for the whole
    pattern and for row blocks, with every kernel id.  (compute-sanitizer's initcheck cannot show this: it does not
    track the bulk-copy engine's global writes, profiles/r02_sanitizer.md.)"""
    import torch
    from sucemfem import _native as N
    co, ce = meshes.unit_cube(5, 4, 3)
    m = dev.DeviceMesh(co, ce)
    sp = dev.DeviceSpace(m, order)
    markers = torch.zeros(sp.n_facets, dtype=torch.int32, device='cuda')
    sp.mark_exterior_facets(markers, 999)
    flags = sp.pec_row_flags(markers, 999)
    b = sp.partition(3)
    for lo, hi in [(0, -1)] + [(b[q], b[q + 1]) for q in range(3)]:
        pat = dev.DevicePattern(sp, lo, hi)
        for kernel in (N.KERNEL_AUTO, N.KERNEL_CANONICAL, N.KERNEL_GENERIC):
            for pec in (None, flags):
                S0, T0 = pat.assemble_ST(pec_flags=pec, kernel=kernel)
                S = torch.full((pat.nnz,), float('nan'), dtype=torch.float64, device='cuda')
                T = torch.full((pat.nnz,), float('nan'), dtype=torch.float64, device='cuda')
                pat.assemble_ST(pec_flags=pec, S=S, T=T, kernel=kernel)
                assert not torch.isnan(S).any() and not torch.isnan(T).any()
                assert torch.equal(S, S0) and torch.equal(T, T0)


def test_lane_pair_edge_kernel_agrees_with_lane_per_item(dev, monkeypatch):
    """The round-1 edge kernel (a lane PAIR per edge, loop over the ring) is the fallback layout when the batch builder
    gives up; SFEM_EDGE_KERNEL=1 selects it at pattern creation.  Same rows, another summation order for the row edge's own
    block: equal to 1e-12, both against the oracle, PEC included."""
    import torch
    from sucemfem import _native as N
    co, ce = meshes.unit_cube(5, 4, 6)
    rng = np.random.default_rng(4)
    co = co + 0.02 * rng.standard_normal(co.shape)
    eps, mu = rng.uniform(1, 16, len(ce)), rng.uniform(1, 2, len(ce))
    num = fem.number_entities(ce)
    ref = fem.assemble_system(co, ce, 2, eps_r=eps, mu_r=mu, pec_facet_mask=fem.exterior_facet_mask(ce, num))
    base = fem.assemble_system(co, ce, 2, eps_r=eps, mu_r=mu)
    sp = dev.DeviceSpace(dev.DeviceMesh(co, ce), 2)
    markers = torch.zeros(sp.n_facets, dtype=torch.int32, device='cuda')
    sp.mark_exterior_facets(markers, 999)
    flags = sp.pec_row_flags(markers, 999)
    d_eps, d_mu = torch.from_numpy(eps).cuda(), torch.from_numpy(mu).cuda()
    out = {}
    for version in ('2', '1'):
        monkeypatch.setenv('SFEM_EDGE_KERNEL', version)
        pat = dev.DevicePattern(sp, 0, -1)
        assert N.lib.sfem_launches_per_assemble_ST(pat.handle, N.KERNEL_CANONICAL) == (3 if version == '2' else 2)
        out[version] = pat.assemble_ST(d_eps, d_mu, flags, kernel=N.KERNEL_CANONICAL)
        for got, want, scale in zip(out[version], (ref['S'], ref['M']), (base['S'], base['M'])):
            assert close(got.cpu().numpy(), want, scale)
    for a, b in zip(out['1'], out['2']):
        assert bool(((a - b).abs() <= 1e-12 * b.abs() + 1e-15 * b.abs().max()).all())


def test_canonical_midsize_vs_c_oracle(dev):
    """UnitCube(12) order 2 (10 368 tets, 3.0 M non-zeros) against the C restatement of the reference algorithm."""
    import torch
    from oracle import cport
    from sucemfem import _native as N
    co, ce = meshes.unit_cube(12)
    rng = np.random.default_rng(5)
    co = co + 0.01 * rng.standard_normal(co.shape)
    eps, mu = rng.uniform(1, 16, len(ce)), rng.uniform(1, 2, len(ce))
    A = cport.CPUAssembler(co, ce, 2)
    Sr, Mr = A.assemble(eps, mu)
    m = dev.DeviceMesh(co, ce)
    sp = dev.DeviceSpace(m, 2)
    pat = sp.pattern()
    rp, cl = pat.csr()
    assert np.array_equal(rp.cpu().numpy(), A.rowptr) and np.array_equal(cl.cpu().numpy(), A.cols)
    S, T = pat.assemble_ST(torch.from_numpy(eps).cuda(), torch.from_numpy(mu).cuda(), kernel=N.KERNEL_AUTO)
    assert close(S.cpu().numpy(), Sr) and close(T.cpu().numpy(), Mr)


TRI_MESHES = [k for k in sorted(MESHES) if MESHES[k][0].shape[1] == 2]


ROWS_CASES = [(k, o) for k in TRI_MESHES for o in (1, 2)] + [(k, 1) for k in TET_MESHES]


@pytest.mark.parametrize('name,order', ROWS_CASES)
def test_values_rows_kernel_vs_oracle(dev, name, order):
    """SFEM_KERNEL_ROWS (one thread per row entity, S and T in one sweep: what SFEM_KERNEL_AUTO takes on triangles and
    on order-1 tetrahedra) against the oracle and the generic kernel, with PEC, default coefficients, determinism and
    row blocks."""
    import torch
    from sucemfem import _native as N
    co, ce = MESHES[name]
    rng = np.random.default_rng(5)
    eps = rng.uniform(1, 16, ce.shape[0])
    mu = rng.uniform(1, 2, ce.shape[0])
    num = fem.number_entities(ce)
    pecmask = fem.exterior_facet_mask(ce, num)
    ref = fem.assemble_system(co, ce, order, eps_r=eps, mu_r=mu)
    refp = fem.assemble_system(co, ce, order, eps_r=eps, mu_r=mu, pec_facet_mask=pecmask)
    sp = dev.DeviceSpace(dev.DeviceMesh(co, ce), order)
    pat = sp.pattern()
    d_eps, d_mu = torch.from_numpy(eps).cuda(), torch.from_numpy(mu).cuda()
    S, T = pat.assemble_ST(d_eps, d_mu, kernel=N.KERNEL_ROWS)
    assert close(S.cpu().numpy(), ref['S']) and close(T.cpu().numpy(), ref['M'])
    Sg, Tg = pat.assemble_ST(d_eps, d_mu, kernel=N.KERNEL_GENERIC)
    assert close(S.cpu().numpy(), Sg.cpu().numpy()) and close(T.cpu().numpy(), Tg.cpu().numpy())
    S1, T1 = pat.assemble_ST(kernel=N.KERNEL_ROWS)
    ref1 = fem.assemble_system(co, ce, order)
    assert close(S1.cpu().numpy(), ref1['S']) and close(T1.cpu().numpy(), ref1['M'])
    markers = torch.zeros(sp.n_facets, dtype=torch.int32, device='cuda')
    sp.mark_exterior_facets(markers, 999)
    flags = sp.pec_row_flags(markers, 999)
    Sp, Tp = pat.assemble_ST(d_eps, d_mu, pec_flags=flags, kernel=N.KERNEL_ROWS)
    assert close(Sp.cpu().numpy(), refp['S'], ref['S']) and close(Tp.cpu().numpy(), refp['M'], ref['M'])
    # unit rows are exact
    rows = refp['pec_dofs']
    rp = ref['rowptr']
    for r_ in rows[:5]:
        seg = Sp.cpu().numpy()[rp[r_]:rp[r_ + 1]]
        assert seg.sum() == 1.0 and np.count_nonzero(seg) == 1
    S2, T2 = pat.assemble_ST(d_eps, d_mu, pec_flags=flags, kernel=N.KERNEL_ROWS)
    assert torch.equal(S2, Sp) and torch.equal(T2, Tp)
    if co.shape[1] == 2:      # AUTO = ROWS on triangles (tets of order 1: canonical where the pattern allows it)
        S3, T3 = pat.assemble_ST(d_eps, d_mu, pec_flags=flags, kernel=N.KERNEL_AUTO)
        assert torch.equal(S3, Sp) and torch.equal(T3, Tp)
    b = sp.partition(3)
    Ss, Ts = [], []
    for q in range(3):
        blk = dev.DevicePattern(sp, b[q], b[q + 1])
        s, t = blk.assemble_ST(d_eps, d_mu, pec_flags=flags, kernel=N.KERNEL_ROWS)
        Ss.append(s)
        Ts.append(t)
    assert torch.equal(torch.cat(Ss), Sp) and torch.equal(torch.cat(Ts), Tp)
