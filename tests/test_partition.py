"""Partitioned numbering (sucemfem/partition.py; SURVEY 8.2 row e) on the CPU: P logical ranks each number ONLY their own
cells + halo (the oracle's first-appearance numbering stands in for the device's K1), exchange two counts and the
interface tables, and must reproduce the numbering of the whole mesh BIT FOR BIT on every cell they hold; owned
ranges tile the global index ranges.  The same phases run through torch.distributed (gloo, world_size 2 and 3)."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, free_port
from oracle import fem, meshes


def _partition_module():
    spec = importlib.util.spec_from_file_location('partition', os.path.join(ROOT, 'sucem-fem_b200', 'sucemfem', 'partition.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _mesh(name):
    if name == 'cube':
        return meshes.unit_cube(5, 4, 6)
    if name == 'cube8':
        return meshes.unit_cube(8)
    co, ce = meshes.unit_cube(4, 5, 3)          # same topology, cells in a scrambled (non-slab) order
    perm = np.random.default_rng(7).permutation(len(ce))
    return co, ce[perm]


def _rank_numbering(P, ce, p, world, order):
    lo, hi = P.cell_ranges(len(ce), world)[p]
    rc = P.RankCells(ce, lo, hi)
    loc = fem.number_entities(rc.cells_local.numpy())        # numbering of the rank's sub-mesh only
    return P.RankNumbering(rc, torch.from_numpy(loc['cell_edges'].astype(np.int32)),
                           torch.from_numpy(loc['cell_faces'].astype(np.int32)), order)


def _check_rank(rn, num, order):
    ids = rn.rc.cell_ids.numpy()
    assert np.array_equal(rn.e_gid.numpy()[rn.ce.numpy()], num['cell_edges'][ids])
    if order == 2:
        assert np.array_equal(rn.f_gid.numpy()[rn.cf.numpy()], num['cell_faces'][ids])
    # owned = a contiguous local range that maps to a contiguous global range
    assert np.array_equal(rn.e_gid.numpy()[rn.e_lo:rn.e_hi], rn.e_off + np.arange(rn.e_hi - rn.e_lo))
    assert (rn.E_global, rn.F_global if order == 2 else 0) == (num['E'], num['F'] if order == 2 else 0)


@pytest.mark.parametrize('name', ['cube', 'scrambled'])
@pytest.mark.parametrize('world', [2, 3, 4, 8])
@pytest.mark.parametrize('order', [1, 2])
def test_logical_ranks_reproduce_global_numbering(name, world, order):
    P = _partition_module()
    co, ce = _mesh(name)
    num = fem.number_entities(ce)
    ranks = P.number_logical_ranks([_rank_numbering(P, ce, p, world, order) for p in range(world)])
    for rn in ranks:
        _check_rank(rn, num, order)
    # the owned ranges tile [0, E) and [0, F)
    assert sum(r.e_hi - r.e_lo for r in ranks) == num['E']
    assert [r.e_off for r in ranks] == list(np.cumsum([0] + [r.e_hi - r.e_lo for r in ranks[:-1]]))
    if order == 2:
        assert sum(r.f_hi - r.f_lo for r in ranks) == num['F']
    # every cell that contains an owned entity is held by the owner (no communication in assembly)
    for p, rn in enumerate(ranks):
        owned = np.arange(rn.e_off, rn.e_off + rn.e_hi - rn.e_lo)
        holders = np.nonzero(np.isin(num['cell_edges'], owned).any(axis=1))[0]
        assert np.isin(holders, rn.rc.cell_ids.numpy()).all()
    # the halo is thin on a slab partition of a structured cube
    if name == 'cube' and world == 2:
        assert all(len(r.rc.cell_ids) < 0.85 * len(ce) for r in ranks)


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    P = _partition_module()
    co, ce = _mesh('cube8')
    rn = P.number_distributed(_rank_numbering(P, ce, rank, world, 2))
    num = fem.number_entities(ce)
    try:
        _check_rank(rn, num, 2)
        ret[rank] = 1
    except AssertionError:
        ret[rank] = 0
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_gloo_ranks_reproduce_global_numbering(world):
    ctx = mp.get_context('spawn')
    ret = ctx.Manager().dict()
    port = free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
    assert all(p.exitcode == 0 for p in procs)
    assert [ret.get(r) for r in range(world)] == [1] * world
