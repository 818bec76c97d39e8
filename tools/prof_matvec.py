import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'sucem-fem_b200')):
    sys.path.insert(0, p)
import torch
from sucemfem import device, _native as N
n3 = int(sys.argv[1]) if len(sys.argv) > 1 else 70
sp = device.DeviceSpace(device.DeviceMesh.unit(n3, n3, n3), 2)
pat = sp.pattern()
S, M = pat.assemble_ST()
S0 = pat.assemble(N.FORM_ABC)
x = torch.randn(sp.ndof, dtype=torch.complex128, device='cuda')
y = torch.empty_like(x)
for _ in range(3):
    pat.matvec([S, M, S0], x, [1.0, -3.0, 0.5j], out=y)
    pat.matvec([M], x, [1.0], out=y)
torch.cuda.synchronize()
