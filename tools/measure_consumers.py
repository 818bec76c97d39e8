"""Device-time measurements outside the headline metric: the fused SpMV (f1), one BiCGStab iteration's cost, the 2-D
(triangle) assembly throughput and order-1 tets.   python tools/measure_consumers.py [n3d] [n2d]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'sucem-fem_b200')):
    sys.path.insert(0, p)
import torch
from sucemfem import device, _native as N

n3 = int(sys.argv[1]) if len(sys.argv) > 1 else 70
n2 = int(sys.argv[2]) if len(sys.argv) > 2 else 4096


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


mesh = device.DeviceMesh.unit(n3, n3, n3)
sp = device.DeviceSpace(mesh, 2)
pat = sp.pattern()
S, M = pat.assemble_ST()
S0 = pat.assemble(N.FORM_ABC)
x = torch.randn(sp.ndof, dtype=torch.complex128, device='cuda')
y = torch.empty_like(x)
for mats, coefs, name in (([S, M, S0], [1.0, -3.0, 0.5j], 'S - k0^2 M + j k0 S_0'), ([M], [1.0], 'M')):
    ms = timed(lambda: pat.matvec(mats, x, coefs, out=y))
    b = 8 * len(mats) * pat.nnz + pat.nnz + 32 * sp.ndof + 8 * sp.n_entities
    print('matvec %-24s UnitCube(%d) order 2: %.3f ms, %.0f GB/s algorithmic (%d matrices, nnz %d)' % (
        name, n3, ms, b / ms / 1e6, len(mats), pat.nnz), flush=True)
xr = torch.randn(sp.ndof, dtype=torch.float64, device='cuda')
yr = torch.empty_like(xr)
ms = timed(lambda: pat.matvec([M], xr, out=yr))
print('matvec M (real vectors): %.3f ms, %.0f GB/s' % (ms, (8 * pat.nnz + pat.nnz + 16 * sp.ndof) / ms / 1e6), flush=True)
b = torch.zeros(sp.ndof, dtype=torch.complex128, device='cuda')
b[::1000] = 1.0
for rep in range(3):        # the first call grows the library's memory pool by the solver's work space
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    xs, info, its, rr = pat.bicgstab([S, M, S0], [1.0, -3.0, 0.5j], b, rtol=1e-30, maxiter=20)
    torch.cuda.synchronize()
    print('bicgstab 20 iterations (40 matvecs), call %d: %.1f ms per iteration' % (rep, (time.perf_counter() - t0) * 1e3 / 20), flush=True)
del S, M, S0, pat, sp, mesh, x, y, xs
torch.cuda.empty_cache()

# order-1 tets at the same size (config 4's order): canonical S+T in one sweep, and S_0
sp1 = device.DeviceSpace(device.DeviceMesh.unit(n3, n3, n3), 1)
p1 = sp1.pattern()
S, T = p1.new_values(), p1.new_values()
ms = timed(lambda: p1.assemble_ST(S=S, T=T))
print('order-1 tets UnitCube(%d): S+T %.3f ms = %.2f G tets/s' % (n3, ms, 6 * n3 ** 3 / ms / 1e6), flush=True)
S0 = p1.new_values()
ms = timed(lambda: p1.assemble(N.FORM_ABC, out=S0))
print('order-1 S_0: %.3f ms' % ms, flush=True)
del S, T, S0, p1, sp1
torch.cuda.empty_cache()

# triangles (generic kernels): UnitSquare(n2), orders 1 and 2
for order in (1, 2):
    sp2 = device.DeviceSpace(device.DeviceMesh.unit(n2, n2, 0), order)
    p2 = sp2.pattern()
    S, T = p2.new_values(), p2.new_values()
    ms = timed(lambda: p2.assemble_ST(S=S, T=T), reps=5)
    cells = 2 * n2 * n2
    print('triangles UnitSquare(%d) order %d: S+T %.3f ms = %.2f G cells/s (nnz %d, %.0f GB/s of values)' % (
        n2, order, ms, cells / ms / 1e6, p2.nnz, 16 * p2.nnz / ms / 1e6), flush=True)
    del S, T, p2, sp2
    torch.cuda.empty_cache()
