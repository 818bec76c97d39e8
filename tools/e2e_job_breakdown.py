"""Timing of the streamed host-buffer job (create / run) at several block counts.  Diagnostic only."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'sucem-fem_b200')):
    sys.path.insert(0, p)
import numpy as np
import torch
from sucemfem import device
n = int(sys.argv[1]) if len(sys.argv) > 1 else 119
gen = device.DeviceMesh.unit(n, n, n)
co, ce = gen.coords.cpu().numpy(), gen.cells.cpu().numpy()
del gen
rng = np.random.default_rng(1234)
pin = lambda a: torch.from_numpy(a).pin_memory()
h_co, h_ce, h_eps, h_mu = pin(co), pin(ce), pin(rng.uniform(1, 16, len(ce))), pin(rng.uniform(1, 2, len(ce)))
out = {}
for nb in (0, 0, 0, 1, 8, 64, 0):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    job = device.HostAssemblyJob(h_co, h_ce, 2, h_eps, h_mu, pec='exterior')
    torch.cuda.synchronize(); t1 = time.perf_counter()
    out = job.run(want=('cols', 'S', 'T'), out=out, n_blocks=nb)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    job.close()
    torch.cuda.synchronize(); t3 = time.perf_counter()
    print('n_blocks %3d: create %.1f ms, run %.1f ms, close %.1f ms, total %.1f ms -> %.2f M tets/s' % (
        nb, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t3 - t0) * 1e3, len(ce) / (t3 - t0) / 1e6), flush=True)
