"""N>1 host logic on CPU: world_size-2 gloo processes partition the rows, 'assemble' their block (sliced from the
oracle: the device path is exercised by the gpu tests) and gather the distributed CSR; the result must equal the
single-process matrix bit for bit (assembly is communication-free, so this is exact)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, free_port


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'sucem-fem_b200'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from oracle import fem, meshes
    # parallel.py is pure torch.distributed: import it without pulling in the CUDA library
    import importlib.util
    spec = importlib.util.spec_from_file_location('parallel', os.path.join(ROOT, 'sucem-fem_b200', 'sucemfem', 'parallel.py'))
    parallel = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(parallel)
    co, ce = meshes.unit_cube(3)
    ref = fem.assemble_system(co, ce, 2)
    num = ref['num']
    n_ent = num['E'] + num['F']
    # weight of an entity = number of incident cells (what sfem_partition_rows balances)
    ents = np.concatenate([num['cell_edges'].ravel(), num['E'] + num['cell_faces'].ravel()])
    prefix = np.concatenate([[0], np.cumsum(np.bincount(ents, minlength=n_ent))])
    bounds = parallel.split_by_weight(prefix, world)
    lo, hi = parallel.my_block(bounds)
    r0, r1 = 2 * lo, 2 * hi                                   # dof rows of the owned entities
    rp = torch.from_numpy(ref['rowptr'][r0:r1 + 1] - ref['rowptr'][r0])
    s, e = ref['rowptr'][r0], ref['rowptr'][r1]
    cols = torch.from_numpy(ref['cols'][s:e].copy())
    vals = torch.from_numpy(ref['S'][s:e].copy())
    full_rp, full_cols, full_vals = parallel.gather_csr_blocks(rp, cols, vals)
    ok = (np.array_equal(full_rp.numpy(), ref['rowptr']) and np.array_equal(full_cols.numpy(), ref['cols'])
          and np.array_equal(full_vals.numpy(), ref['S']) and bounds[0] == 0 and bounds[-1] == n_ent)
    # the two-class partition the bench uses: one edge range + one face range per rank
    eb, ob = parallel.split_by_class(prefix, num['E'], world)

    def block(b):
        a0, a1 = 2 * b[rank], 2 * b[rank + 1]
        s0, s1 = ref['rowptr'][a0], ref['rowptr'][a1]
        return (torch.from_numpy(ref['rowptr'][a0:a1 + 1] - s0), torch.from_numpy(ref['cols'][s0:s1].copy()),
                torch.from_numpy(ref['M'][s0:s1].copy()))
    rp2, cl2, vl2 = parallel.gather_csr_two_class(block(eb), block(ob))
    ok2 = (np.array_equal(rp2.numpy(), ref['rowptr']) and np.array_equal(cl2.numpy(), ref['cols'])
           and np.array_equal(vl2.numpy(), ref['M']) and eb[-1] == num['E'] == ob[0] and ob[-1] == n_ent)
    # several value arrays on one pattern (S and M): the pattern is gathered once
    def block2(b):
        a0, a1 = 2 * b[rank], 2 * b[rank + 1]
        s0, s1 = ref['rowptr'][a0], ref['rowptr'][a1]
        return (torch.from_numpy(ref['rowptr'][a0:a1 + 1] - s0), torch.from_numpy(ref['cols'][s0:s1].copy()),
                [torch.from_numpy(ref['S'][s0:s1].copy()), torch.from_numpy(ref['M'][s0:s1].copy())])
    rp3, cl3, (vS, vM) = parallel.gather_csr_classes([block2(eb), block2(ob)])
    ok3 = (np.array_equal(rp3.numpy(), ref['rowptr']) and np.array_equal(cl3.numpy(), ref['cols'])
           and np.array_equal(vS.numpy(), ref['S']) and np.array_equal(vM.numpy(), ref['M']))
    ret[rank] = bool(ok) and bool(ok2) and bool(ok3) and (hi - lo) > 0
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_row_blocks_gather_to_full_matrix(world):
    ctx = mp.get_context('spawn')
    ret = ctx.Manager().dict()
    port = free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    assert all(p.exitcode == 0 for p in procs)
    assert [ret.get(r) for r in range(world)] == [True] * world


def test_split_by_weight_balances():
    import importlib.util
    spec = importlib.util.spec_from_file_location('parallel', os.path.join(ROOT, 'sucem-fem_b200', 'sucemfem', 'parallel.py'))
    parallel = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(parallel)
    w = np.random.default_rng(0).integers(1, 9, 1000)
    prefix = np.concatenate([[0], np.cumsum(w)])
    for P in (1, 2, 4, 8):
        b = parallel.split_by_weight(prefix, P)
        assert b[0] == 0 and b[-1] == 1000 and all(x <= y for x, y in zip(b, b[1:]))
        loads = [prefix[b[p + 1]] - prefix[b[p]] for p in range(P)]
        assert max(loads) - min(loads) <= 2 * w.max()
