"""Order-1 tetrahedra (BASELINE config 4: UnitCube(119), ABC): S+T with the thread-per-edge kernel (SFEM_KERNEL_ROWS),
the lane-pair canonical kernel and the generic kernel, and S_0.   python tools/tet1_time.py [n]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'sucem-fem_b200')):
    sys.path.insert(0, p)
import torch
from sucemfem import device, _native as N
n = int(sys.argv[1]) if len(sys.argv) > 1 else 119
mesh = device.DeviceMesh.unit(n, n, n)
sp = device.DeviceSpace(mesh, 1)
pat = sp.pattern()
g = torch.Generator(device='cuda').manual_seed(1)
eps = 1 + 15 * torch.rand(mesh.num_cells, dtype=torch.float64, device='cuda', generator=g)
mu = 1 + torch.rand(mesh.num_cells, dtype=torch.float64, device='cuda', generator=g)
S, T = pat.new_values(), pat.new_values()
cells, verts = mesh.num_cells, mesh.num_vertices
b_alg = 16 * pat.nnz + 16 * cells + 24 * verts + 16 * cells


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for name, kern in (('thread/edge', N.KERNEL_ROWS), ('canonical', N.KERNEL_CANONICAL), ('generic', N.KERNEL_GENERIC)):
    ms = timed(lambda: pat.assemble_ST(eps, mu, None, S, T, kernel=kern))
    print('UnitCube(%d) order 1 %-12s S+T %.3f ms = %.2f G tets/s, %.0f B/tet algorithmic -> %.2f TB/s' % (
        n, name, ms, cells / ms / 1e6, b_alg / cells, b_alg / ms / 1e9), flush=True)
ms = timed(lambda: pat.assemble(N.FORM_ABC, out=S))
print('UnitCube(%d) order 1 S_0 %.3f ms' % (n, ms))
