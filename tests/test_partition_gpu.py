"""Partitioned mesh on the device (sucemfem/partition.py + sfem_space_set_global_ids; SURVEY 8.2 row e): every logical
rank builds its space from its own cells + halo only, the phases of the global numbering run between them, and the two
row blocks each rank assembles -- pattern (global column ids), S, T, PEC unit rows, compact S_0 -- concatenate BIT FOR
BIT to what one space over the whole mesh gives.  (All logical ranks share one GPU here; tests/test_partition.py runs the
numbering phases through torch.distributed with gloo, and bench.py --gpus N through NCCL.)"""
import numpy as np
import pytest

from oracle import meshes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('order', [2, 1])
@pytest.mark.parametrize('world', [2, 3])
def test_rank_blocks_concatenate_to_the_global_assembly(order, world):
    import torch
    from sucemfem import device as dev, partition as P, _native as N
    co, ce = meshes.unit_cube(6, 5, 7)
    rng = np.random.default_rng(11)
    co = co + 0.02 * rng.standard_normal(co.shape)
    eps, mu = rng.uniform(1, 16, len(ce)), rng.uniform(1, 2, len(ce))
    # the whole mesh on one space
    sp = dev.DeviceSpace(dev.DeviceMesh(co, ce), order)
    markers = torch.zeros(sp.n_facets, dtype=torch.int32, device='cuda')
    sp.mark_exterior_facets(markers, 999)
    flags = sp.pec_row_flags(markers, 999)
    pat = sp.pattern()
    d_eps, d_mu = torch.from_numpy(eps).cuda(), torch.from_numpy(mu).cuda()
    S, T = pat.assemble_ST(d_eps, d_mu, flags)
    rowptr, cols = pat.csr()
    E = sp.n_edges
    # logical ranks
    ranks = [P.RankSpace(co, ce, order, p, world, distributed=False) for p in range(world)]
    P.number_logical_ranks([r.numbering for r in ranks])
    blocks = {0: [], 1: []}
    for r in ranks:
        r._attach()
        assert r.ndof_global == sp.ndof
        lm = torch.zeros(r.space.n_facets, dtype=torch.int32, device='cuda')
        r.space.mark_exterior_facets(lm, 999)
        lflags = r.space.pec_row_flags(lm, 999)
        le = torch.from_numpy(r.local_cell_values(eps)).cuda()
        lmu = torch.from_numpy(r.local_cell_values(mu)).cuda()
        for k, bp in enumerate(r.patterns()):
            s, t = bp.assemble_ST(le, lmu, lflags)
            rp, cl = bp.csr()
            blocks[k].append((bp.row_begin, bp.n_rows, rp, cl, s, t))
    # edge blocks tile the edge rows, face blocks the face rows, in rank order
    pos = 0
    for k in ((0, 1) if order == 2 else (0,)):
        for row_begin, n_rows, rp, cl, s, t in blocks[k]:
            assert row_begin == pos
            a, b = int(rowptr[pos]), int(rowptr[pos + n_rows])
            assert torch.equal(rp + a, rowptr[pos:pos + n_rows + 1])
            assert torch.equal(cl, cols[a:b])
            assert torch.equal(s, S[a:b]) and torch.equal(t, T[a:b])
            pos += n_rows
        if k == 0:
            assert pos == order * E
    assert pos == sp.ndof


def test_halo_is_thin_and_generic_kernel_agrees():
    import torch
    from sucemfem import device as dev, partition as P, _native as N
    co, ce = meshes.unit_cube(10)
    ranks = [P.RankSpace(co, ce, 2, p, 4, distributed=False) for p in range(4)]
    P.number_logical_ranks([r.numbering for r in ranks])
    for r in ranks:
        r._attach()
        assert len(r.rc.cell_ids) <= 0.5 * len(ce) + 1       # own quarter + at most one cube layer either side
        for bp in r.patterns():
            a = bp.assemble_ST(kernel=N.KERNEL_CANONICAL)
            b = bp.assemble_ST(kernel=N.KERNEL_GENERIC)
            for x, y in zip(a, b):
                assert float((x - y).abs().max()) <= 1e-12 * float(y.abs().max())


def test_rank_host_job_single_rank_group():
    """partition.RankHostJob through torch.distributed (NCCL, a group of one): the whole mesh as the only rank's share must
    give what the plain host job gives, bit for bit (plumbing of the host-buffer path of the partitioned mesh)."""
    import os
    import torch
    import torch.distributed as dist
    from sucemfem import device as dev, partition as P
    co, ce = meshes.unit_cube(5, 4, 3)
    rng = np.random.default_rng(2)
    eps, mu = rng.uniform(1, 16, len(ce)), rng.uniform(1, 2, len(ce))
    ref = dev.assemble_host(co, ce, 2, eps, mu, pec='exterior')
    from conftest import free_port
    dist.init_process_group('nccl', init_method='tcp://127.0.0.1:%d' % free_port(), rank=0, world_size=1)
    try:
        rc = P.RankCells(ce, 0, len(ce))
        assert len(rc.cell_ids) == len(ce) and rc.own_lo == 0
        job = P.RankHostJob(rc, co, ce, 2, eps, mu, pec='exterior')
        oe, of = job.run()
        job.close()
    finally:
        dist.destroy_process_group()
    ne = oe['S'].numel()
    for k in ('cols', 'S', 'T'):
        assert torch.equal(torch.cat([oe[k], of[k]]), ref[k]), k
    assert torch.equal(oe['rowptr'], ref['rowptr'][:oe['rowptr'].numel()])
    assert torch.equal(of['rowptr'] + ne, ref['rowptr'][oe['rowptr'].numel() - 1:])


def _host_job_worker(rank, world, port, ret):
    import os
    import sys
    import torch
    import torch.distributed as dist
    from conftest import ROOT
    sys.path.insert(0, os.path.join(ROOT, 'sucem-fem_b200'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)      # both ranks share cuda:0: gloo carries the exchange
    from sucemfem import device as dev, partition as P
    co, ce = meshes.unit_cube(6, 5, 4)
    rng = np.random.default_rng(5)
    eps, mu = rng.uniform(1, 16, len(ce)), rng.uniform(1, 2, len(ce))
    ref = dev.assemble_host(co, ce, 2, eps, mu, pec='exterior', want=('cols', 'S', 'T', 'S0'))
    lo, hi = P.cell_ranges(len(ce), world)[rank]
    rc = P.RankCells(torch.from_numpy(ce).cuda(), lo, hi)
    ids = rc.cell_ids.cpu().numpy()
    job = P.RankHostJob(rc, co[rc.vertex_ids.cpu().numpy()], rc.cells_local.cpu().numpy(), 2, eps[ids], mu[ids], pec='exterior')
    oe, of = job.run(want=('cols', 'S', 'T', 'S0'))
    rn = job.numbering
    job.close()
    ok = True
    E = rn.E_global
    for out, row0 in ((oe, 2 * rn.e_off), (of, 2 * (E + rn.f_off))):
        nrows = out['rowptr'].numel() - 1
        a, b = int(ref['rowptr'][row0]), int(ref['rowptr'][row0 + nrows])
        ok = ok and torch.equal(out['rowptr'] + a, ref['rowptr'][row0:row0 + nrows + 1])
        for k in ('cols', 'S', 'T', 'S0'):
            ok = ok and torch.equal(out[k], ref[k][a:b])
    ret[rank] = int(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_rank_host_jobs_two_processes_reproduce_the_serial_csr():
    """Two PROCESSES (gloo for the numbering exchange, one GPU shared) each run partition.RankHostJob on their own cells +
    halo: the rows they stream out -- pattern with global column ids, S, T and the full-pattern S_0, PEC unit rows included
    -- are bit for bit the rows of the serial host job."""
    import os
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    ret = ctx.Manager().dict()
    from conftest import free_port
    port = free_port()
    procs = [ctx.Process(target=_host_job_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
    assert all(p.exitcode == 0 for p in procs)
    assert [ret.get(r) for r in range(2)] == [1, 1]
