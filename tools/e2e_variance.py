"""Why do some e2e steps take 2-3 x the median?  The bench's e2e loop with phase times, pool statistics and a
resident-benchmark-like prologue (torch tensors of S+T size allocated and dropped first).  Diagnostic only."""
import ctypes as C
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'sucem-fem_b200')):
    sys.path.insert(0, p)
import numpy as np
import torch
from sucemfem import device, _native as N
n = int(sys.argv[1]) if len(sys.argv) > 1 else 119
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 8
prologue = (sys.argv[3] if len(sys.argv) > 3 else 'resident') == 'resident'


def pool():
    r, u = C.c_int64(), C.c_int64()
    N.lib.sfem_cached_memory_info(C.byref(r), C.byref(u))
    return r.value / 1e9, u.value / 1e9


if prologue:   # what bench.py does before the e2e leg
    mesh = device.DeviceMesh.unit(n, n, n)
    sp = device.DeviceSpace(mesh, 2)
    pat = sp.pattern()
    S, T = pat.assemble_ST()
    torch.cuda.synchronize()
    del S, T, pat, sp, mesh
    torch.cuda.empty_cache()
    N.lib.sfem_release_cached_memory()
gen = device.DeviceMesh.unit(n, n, n)
co, ce = gen.coords.cpu().numpy(), gen.cells.cpu().numpy()
del gen
rng = np.random.default_rng(1234)
pin = lambda a: torch.from_numpy(a).pin_memory()
h_co, h_ce, h_eps, h_mu = pin(co), pin(ce), pin(rng.uniform(1, 16, len(ce))), pin(rng.uniform(1, 2, len(ce)))
out = {}
print('pool reserved/used GB at start: %.2f / %.2f; torch reserved %.2f GB' % (pool() + (torch.cuda.memory_reserved() / 1e9,)))
for it in range(iters):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    job = device.HostAssemblyJob(h_co, h_ce, 2, h_eps, h_mu, pec='exterior')
    t1 = time.perf_counter()
    torch.cuda.synchronize(); t1s = time.perf_counter()
    out = job.run(want=('cols', 'S', 'T'), out=out)
    t2 = time.perf_counter()
    job.close()
    torch.cuda.synchronize(); t3 = time.perf_counter()
    print('it %d: create %.1f (+sync %.1f) ms, run %.1f ms, close %.1f ms, total %.1f ms; pool reserved/used %.2f / %.2f GB'
          % ((it, (t1 - t0) * 1e3, (t1s - t1) * 1e3, (t2 - t1s) * 1e3, (t3 - t2) * 1e3, (t3 - t0) * 1e3) + pool()), flush=True)
