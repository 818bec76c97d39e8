"""Device eigen solver beyond dense sizes: python tools/eigen_large.py [n] [nev] [order]
PEC cavity UnitCube(n) with eps_r = 4 in the lower half (BASELINE config 3 shape), DeviceEigenSolver (shift-invert Lanczos,
CUDA-graph MINRES on sfem_matvec).  Prints the eigenvalues, the eigenpair residuals, the inner-iteration counts and the time."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'sucem-fem_b200'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import numpy as np
import torch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
nev = int(sys.argv[2]) if len(sys.argv) > 2 else 4
order = int(sys.argv[3]) if len(sys.argv) > 3 else 2
from test_eigen_gpu import _eigenproblem
from sucemfem.ProblemConfigurations.EMVectorWaveEigenproblem import DeviceEigenSolver

t0 = time.perf_counter()
ep = _eigenproblem(n, order, eps_lower_half=4.0)
torch.cuda.synchronize()
t1 = time.perf_counter()
es = DeviceEigenSolver()
es.set_eigenproblem(ep)
es.set_sigma(2.0)
lam, X = es.solve_problem(nev, inner_rtol=1e-12, return_device=True)
torch.cuda.synchronize()
t2 = time.perf_counter()
print(json.dumps({"workload": "UnitCube(%d) order %d PEC cavity, eps_r = 4 below z = 0.5" % (n, order),
                  "ndof": ep.get_global_dimension(), "nev": nev, "k2": [float(x) for x in lam],
                  "residuals": es.residuals, "info": es.info, "assembly_s": round(t1 - t0, 3), "eigen_s": round(t2 - t1, 2)}))
