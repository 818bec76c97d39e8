"""SURVEY 8.2 row (e), second half: the ONE collective of the path -- gathering the distributed CSR for a consumer that
wants the whole matrix on every rank -- over NCCL / NVLink.  Not part of the assembly timing.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
      tools/gather_nccl.py [n] [reps]

Every rank numbers UnitCube(n), owns one edge range + one face range (sfem_partition_rows_by_class), assembles its
S and T blocks (order 2, PEC) and then all-gathers rowptr / cols / S / T of both classes
(sucemfem.parallel.gather_csr_two_class).  Rank 0 also assembles the whole matrix by itself and checks the gathered
arrays against it BIT FOR BIT.  Reports bytes received per rank, time (CUDA events, max over ranks) and GB/s per GPU.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'sucem-fem_b200'))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 70
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
    dist.init_process_group('nccl')
    from sucemfem import device, parallel
    mesh = device.DeviceMesh.unit(n, n, n)
    sp = device.DeviceSpace(mesh, 2)
    markers = torch.zeros(sp.n_facets, dtype=torch.int32, device='cuda')
    sp.mark_exterior_facets(markers, 999)
    flags = sp.pec_row_flags(markers, 999)
    g = torch.Generator(device='cuda').manual_seed(1234)
    eps = 1 + 15 * torch.rand(mesh.num_cells, dtype=torch.float64, device='cuda', generator=g)
    mu = 1 + torch.rand(mesh.num_cells, dtype=torch.float64, device='cuda', generator=g)
    eb, ob = sp.partition_by_class(world)
    blocks = []
    for b in (eb, ob):
        pat = device.DevicePattern(sp, b[rank], b[rank + 1])
        rp, cl = pat.csr()
        S, T = pat.assemble_ST(eps, mu, pec_flags=flags)
        blocks.append((rp, cl, [S, T]))
        del pat
    own = sum(x[1].numel() for x in blocks)
    times = []
    for it in range(reps + 1):
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rowptr, cols, (Sf, Tf) = parallel.gather_csr_two_class(blocks[0], blocks[1])
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if it > 0:
            times.append(float(t.item()))
        if it < reps:
            del rowptr, cols, Sf, Tf
    nnz = cols.numel()
    ok = None
    if rank == 0:
        del blocks
        full = sp.pattern()
        rp0, cl0 = full.csr()
        ok = bool(torch.equal(rowptr, rp0) and torch.equal(cols, cl0))
        del rp0, cl0, rowptr, cols
        S0, T0 = full.assemble_ST(eps, mu, pec_flags=flags)
        ok = ok and bool(torch.equal(Sf, S0) and torch.equal(Tf, T0))
        ms = sorted(times)[len(times) // 2]
        recv = (nnz - own) * 20 + (sp.ndof + 1) * 8          # bytes this rank received: cols + S + T (+ rowptr)
        print(json.dumps({"what": "NCCL gather of the distributed CSR (rowptr, cols, S, T), two-class row partition",
                          "workload": "UnitCube(%d) order 2" % n, "n_gpus": world, "ndof": sp.ndof, "nnz": nnz,
                          "bytes_out_per_rank": nnz * 20 + (sp.ndof + 1) * 8, "bytes_received_rank0": recv,
                          "ms": ms, "all_ms": [round(x, 2) for x in times],
                          "GBps_received_per_gpu": recv / ms / 1e6, "bit_exact_vs_single_rank": ok}))
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == '__main__':
    main()
