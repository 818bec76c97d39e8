"""Stage-by-stage timing of the host-buffer pipeline (H2D, numbering, pattern, values, export, D2H) and the raw
pinned-memory PCIe rates of the box.  Diagnostic only (not a bench line).   python tools/e2e_breakdown.py [n]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'sucem-fem_b200')):
    sys.path.insert(0, p)
import numpy as np
import torch
from sucemfem import device

n = int(sys.argv[1]) if len(sys.argv) > 1 else 119


def tick(label, t0):
    torch.cuda.synchronize()
    t = time.perf_counter()
    print('%-28s %9.1f ms' % (label, (t - t0) * 1e3), flush=True)
    return t


# raw PCIe
for gb in (1, 8):
    h = torch.empty(gb << 30, dtype=torch.uint8).pin_memory()
    d = torch.empty(gb << 30, dtype=torch.uint8, device='cuda')
    for name, fn in (('D2H', lambda: h.copy_(d, non_blocking=True)), ('H2D', lambda: d.copy_(h, non_blocking=True))):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        print('%s pinned %d GiB: %.1f GB/s' % (name, gb, (gb << 30) / (time.perf_counter() - t0) / 1e9), flush=True)
    del h, d
print('host cpus', os.cpu_count(), 'mem', os.sysconf('SC_PAGE_SIZE') * os.sysconf('SC_PHYS_PAGES') / 1e9, 'GB')

gen = device.DeviceMesh.unit(n, n, n)
co, ce = gen.coords.cpu().numpy(), gen.cells.cpu().numpy()
del gen
rng = np.random.default_rng(1234)
eps, mu = rng.uniform(1, 16, len(ce)), rng.uniform(1, 2, len(ce))
pin = lambda a: torch.from_numpy(a).pin_memory()
h_co, h_ce, h_eps, h_mu = pin(co), pin(ce), pin(eps), pin(mu)
out = None
for it in range(3):
    print('--- iteration', it)
    torch.cuda.synchronize()
    t0 = t = time.perf_counter()
    d_co, d_ce = h_co.cuda(non_blocking=True), h_ce.cuda(non_blocking=True)
    d_eps, d_mu = h_eps.cuda(non_blocking=True), h_mu.cuda(non_blocking=True)
    t = tick('H2D', t)
    mesh = device.DeviceMesh(d_co, d_ce)
    space = device.DeviceSpace(mesh, 2)
    t = tick('numbering (space)', t)
    pat = space.pattern()
    t = tick('pattern', t)
    markers = torch.zeros(space.n_facets, dtype=torch.int32, device='cuda')
    space.mark_exterior_facets(markers, 999)
    flags = space.pec_row_flags(markers, 999)
    t = tick('pec flags', t)
    S, T = pat.assemble_ST(d_eps, d_mu, flags)
    t = tick('values S+T (incl. alloc)', t)
    rp, cl = pat.csr()
    t = tick('export dof-level CSR', t)
    if out is None:
        out = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in
               (('S', S), ('T', T), ('rowptr', rp), ('cols', cl))}
        t = tick('pinned alloc (first only)', t)
    for k, v in (('S', S), ('T', T), ('rowptr', rp), ('cols', cl)):
        out[k].copy_(v, non_blocking=True)
        t = tick('D2H ' + k + ' %.1f GB' % (v.numel() * v.element_size() / 1e9), t)
    print('total %.1f ms' % ((t - t0) * 1e3))
    del S, T, rp, cl, pat, space, mesh
