"""2-D (triangle) S+T assembly throughput: the thread-per-entity kernel (SFEM_KERNEL_AUTO) against the generic
warp-per-row kernel, UnitSquare(n), orders 1 and 2, with PEC.   python tools/tri_time.py [n]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'sucem-fem_b200')):
    sys.path.insert(0, p)
import torch
from sucemfem import device, _native as N
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
for order in (1, 2):
    mesh = device.DeviceMesh.unit(n, n, 0)
    sp = device.DeviceSpace(mesh, order)
    pat = sp.pattern()
    markers = torch.zeros(sp.n_facets, dtype=torch.int32, device='cuda')
    sp.mark_exterior_facets(markers, 999)
    flags = sp.pec_row_flags(markers, 999)
    g = torch.Generator(device='cuda').manual_seed(1)
    eps = 1 + 15 * torch.rand(mesh.num_cells, dtype=torch.float64, device='cuda', generator=g)
    mu = 1 + torch.rand(mesh.num_cells, dtype=torch.float64, device='cuda', generator=g)
    S, T = pat.new_values(), pat.new_values()
    cells, verts = mesh.num_cells, mesh.num_vertices
    b_alg = 16 * pat.nnz + 12 * cells + 16 * verts + 16 * cells       # values + connectivity + coordinates + coefficients
    for name, kern in (('thread/entity', N.KERNEL_AUTO), ('generic', N.KERNEL_GENERIC)):
        for _ in range(3):
            pat.assemble_ST(eps, mu, flags, S, T, kernel=kern)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            pat.assemble_ST(eps, mu, flags, S, T, kernel=kern)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print('UnitSquare(%d) order %d %-14s S+T %.3f ms = %.2f G cells/s, %.0f B/cell algorithmic -> %.2f TB/s' % (
            n, order, name, ms, cells / ms / 1e6, b_alg / cells, b_alg / ms / 1e9), flush=True)
    del S, T, pat, sp, mesh
    torch.cuda.empty_cache()
