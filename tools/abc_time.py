import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'sucem-fem_b200')); sys.path.insert(0, ROOT)
import torch
from sucemfem import device, _native as N
n = int(sys.argv[1]) if len(sys.argv) > 1 else 119
for order in (1, 2):
    mesh = device.DeviceMesh.unit(n, n, n)
    sp = device.DeviceSpace(mesh, order)
    pat = sp.pattern()
    out = pat.new_values()
    for form, name in ((N.FORM_ABC, 'S_0'), (N.FORM_MASS, 'M generic'), (N.FORM_STIFFNESS, 'S generic')):
        pat.assemble(form, out=out, kernel=N.KERNEL_GENERIC)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            pat.assemble(form, out=out, kernel=N.KERNEL_GENERIC)
        e1.record(); torch.cuda.synchronize()
        print('n=%d order %d %-10s %.2f ms  (nnz %d)' % (n, order, name, e0.elapsed_time(e1) / 3, pat.nnz))
    del out, pat, sp, mesh
