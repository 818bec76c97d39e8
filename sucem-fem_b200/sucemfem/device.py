"""Device-side objects behind the sucemfem API: thin owners of C-ABI handles plus the torch tensors
(device buffers only) the handles point into.  All arithmetic happens in libsucem_b200.so.

  DeviceMesh      <- dolfin.Mesh            (coordinates / cells resident in HBM)
  DeviceSpace     <- dolfin.FunctionSpace   (/root/reference/sucemfem/ProblemConfigurations/EMProblem.py:100-105)
  DevicePattern   <- the tensor layout DOLFIN builds inside dolfin.assemble
                     (/root/reference/sucemfem/SystemMatrices.py:54)
"""
import ctypes as C

import numpy as np
import torch

from . import _native as N


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("sucemfem (B200 build) needs a CUDA device: there is no CPU fallback")


def pinned_empty(n, dtype):
    """Page-locked host tensor (cudaHostAlloc through torch)."""
    return torch.empty(int(n), dtype=dtype, pin_memory=True)


class DeviceMesh(object):
    def __init__(self, coords, cells, device=None):
        """coords [V, gdim] float64, cells [C, gdim+1] integer (host numpy arrays or device tensors)."""
        require_cuda()
        dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.device = dev
        with torch.cuda.device(dev):
            if isinstance(coords, torch.Tensor):
                self.coords = coords.to(dev, torch.float64).contiguous()
                self.cells = cells.to(dev, torch.int32).contiguous()
            else:
                self.coords = torch.from_numpy(np.ascontiguousarray(coords, dtype=np.float64)).to(dev)
                cells = np.sort(np.asarray(cells), axis=1)      # UFC ordering of the cell's vertices
                self.cells = torch.from_numpy(np.ascontiguousarray(cells, dtype=np.int32)).to(dev)
            self.gdim = int(self.coords.shape[1])
            assert self.cells.shape[1] == self.gdim + 1
            h = C.c_void_p()
            N.check(N.lib.sfem_mesh_create(self.gdim, self.coords.shape[0], _ptr(self.coords), self.cells.shape[0],
                                           _ptr(self.cells), _stream_ptr(), C.byref(h)))
            self.handle = h

    @classmethod
    def unit(cls, nx, ny, nz=0, scale=None, shift=None, device=None):
        """UnitCube(nx,ny,nz) / UnitSquare(nx,ny) generated directly in HBM (nz=0: 2-D)."""
        require_cuda()
        dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        gdim = 3 if nz > 0 else 2
        V = (nx + 1) * (ny + 1) * (nz + 1)
        ncell = 6 * nx * ny * nz if nz > 0 else 2 * nx * ny
        with torch.cuda.device(dev):
            coords = torch.empty((V, gdim), dtype=torch.float64, device=dev)
            cells = torch.empty((ncell, gdim + 1), dtype=torch.int32, device=dev)
            sc = (C.c_double * 3)(*(list(scale) + [1.0] * 3)[:3]) if scale is not None else None
            sh = (C.c_double * 3)(*(list(shift) + [0.0] * 3)[:3]) if shift is not None else None
            N.check(N.lib.sfem_fill_unit_mesh(nx, ny, nz, sc, sh, _ptr(coords), _ptr(cells), _stream_ptr()))
        return cls(coords, cells, device=dev)

    @property
    def num_cells(self):
        return int(self.cells.shape[0])

    @property
    def num_vertices(self):
        return int(self.coords.shape[0])

    def __del__(self):
        h = getattr(self, 'handle', None)
        if h:
            N.lib.sfem_mesh_destroy(h)
            self.handle = None


class DeviceSpace(object):
    def __init__(self, mesh, order):
        self.mesh = mesh
        self.order = int(order)
        h = C.c_void_p()
        with torch.cuda.device(mesh.device):
            N.check(N.lib.sfem_space_create(mesh.handle, self.order, _stream_ptr(), C.byref(h)))
        self.handle = h
        nd, ne, nf, ndc, nfa = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int(), C.c_int64()
        N.check(N.lib.sfem_space_info(h, C.byref(nd), C.byref(ne), C.byref(nf), C.byref(ndc)))
        N.check(N.lib.sfem_space_num_facets(h, C.byref(nfa)))
        self.ndof, self.n_edges, self.n_faces, self.ndof_cell = nd.value, ne.value, nf.value, ndc.value
        self.n_facets = nfa.value
        self.n_entities = self.ndof // self.order
        self._patterns = {}

    def cell_dofs(self):
        out = torch.empty((self.mesh.num_cells, self.ndof_cell), dtype=torch.int32, device=self.mesh.device)
        with torch.cuda.device(self.mesh.device):
            N.check(N.lib.sfem_space_cell_dofs(self.handle, _ptr(out), _stream_ptr()))
        return out

    def cell_entities(self):
        dev = self.mesh.device
        ce = torch.empty((self.mesh.num_cells, 6 if self.mesh.gdim == 3 else 3), dtype=torch.int32, device=dev)
        cf = torch.empty((self.mesh.num_cells, 4), dtype=torch.int32, device=dev) if self.mesh.gdim == 3 else None
        with torch.cuda.device(dev):
            N.check(N.lib.sfem_space_cell_entities(self.handle, _ptr(ce), _ptr(cf), _stream_ptr()))
        return ce, cf

    def mark_exterior_facets(self, markers, value):
        with torch.cuda.device(self.mesh.device):
            N.check(N.lib.sfem_mark_exterior_facets(self.handle, _ptr(markers), int(value), _stream_ptr()))
        return markers

    def pec_row_flags(self, markers, region, flags=None):
        if flags is None:
            flags = torch.zeros(self.ndof, dtype=torch.uint8, device=self.mesh.device)
        with torch.cuda.device(self.mesh.device):
            N.check(N.lib.sfem_pec_row_flags(self.handle, _ptr(markers), int(region), _ptr(flags), _stream_ptr()))
        return flags

    def partition(self, nparts):
        b = (C.c_int64 * (nparts + 1))()
        with torch.cuda.device(self.mesh.device):
            N.check(N.lib.sfem_partition_rows(self.handle, nparts, b, _stream_ptr()))
        return [int(x) for x in b]

    def partition_by_class(self, nparts):
        """(edge_bounds, other_bounds): rank p owns edges [eb[p], eb[p+1]) and faces/interiors [ob[p], ob[p+1])."""
        eb, ob = (C.c_int64 * (nparts + 1))(), (C.c_int64 * (nparts + 1))()
        with torch.cuda.device(self.mesh.device):
            N.check(N.lib.sfem_partition_rows_by_class(self.handle, nparts, eb, ob, _stream_ptr()))
        return [int(x) for x in eb], [int(x) for x in ob]

    def count_nnz(self):
        """nnz of the full dof-level pattern (one count pass over all rows, cached in the space handle)."""
        nnz = C.c_int64()
        with torch.cuda.device(self.mesh.device):
            N.check(N.lib.sfem_space_count_nnz(self.handle, _stream_ptr(), C.byref(nnz)))
        return nnz.value

    def assemble_to_host(self, eps_r=None, mu_r=None, pec_flags=None, want=('S', 'T'), out=None, n_blocks=0):
        """Stream the CSR of the requested matrices to page-locked HOST arrays, row block by row block (pattern fill,
        values and CSR expansion of block k overlap the device->host copy of block k-1; two blocks live in HBM).
        want: subset of ('cols', 'S', 'T', 'S0') ('rowptr' always).  out: dict of pinned CPU tensors to reuse
        (page-locking tens of GB is slow, so callers that repeat the call keep them).  Returns the dict."""
        nnz = self.count_nnz()
        out = {} if out is None else out
        spec = {'rowptr': (self.ndof + 1, torch.int64), 'cols': (nnz, torch.int32), 'S': (nnz, torch.float64),
                'T': (nnz, torch.float64), 'S0': (nnz, torch.float64)}
        for k in ('rowptr',) + tuple(want):
            n, dt = spec[k]
            if k not in out or out[k].numel() != n or out[k].dtype != dt:
                out[k] = pinned_empty(n, dt)
        hp = lambda k: C.c_void_p(out[k].data_ptr()) if k in want or k == 'rowptr' else C.c_void_p(0)
        with torch.cuda.device(self.mesh.device):
            N.check(N.lib.sfem_space_assemble_to_host(self.handle, _ptr(eps_r), _ptr(mu_r), _ptr(pec_flags), hp('rowptr'),
                                                      hp('cols'), hp('S'), hp('T'), hp('S0'), int(n_blocks),
                                                      _stream_ptr()))
        return out

    def pattern(self, ent_lo=0, ent_hi=-1):
        key = (ent_lo, ent_hi)
        if key not in self._patterns:
            self._patterns[key] = DevicePattern(self, ent_lo, ent_hi)
        return self._patterns[key]

    def __del__(self):
        self._patterns = {}
        h = getattr(self, 'handle', None)
        if h:
            N.lib.sfem_space_destroy(h)
            self.handle = None


class DevicePattern(object):
    def __init__(self, space, ent_lo=0, ent_hi=-1):
        self.space = space
        h = C.c_void_p()
        with torch.cuda.device(space.mesh.device):
            N.check(N.lib.sfem_pattern_create(space.handle, ent_lo, ent_hi, _stream_ptr(), C.byref(h)))
        self.handle = h
        rb, nr, nnz = C.c_int64(), C.c_int64(), C.c_int64()
        N.check(N.lib.sfem_pattern_info(h, C.byref(rb), C.byref(nr), C.byref(nnz)))
        self.row_begin, self.n_rows, self.nnz = rb.value, nr.value, nnz.value
        self._csr = None

    def csr(self):
        """(rowptr int64 [n_rows+1], cols int32 [nnz]) device tensors, built once."""
        if self._csr is None:
            dev = self.space.mesh.device
            rowptr = torch.empty(self.n_rows + 1, dtype=torch.int64, device=dev)
            cols = torch.empty(self.nnz, dtype=torch.int32, device=dev)
            with torch.cuda.device(dev):
                N.check(N.lib.sfem_pattern_export(self.handle, _ptr(rowptr), _ptr(cols), _stream_ptr()))
            self._csr = (rowptr, cols)
        return self._csr

    def new_values(self):
        return torch.empty(self.nnz, dtype=torch.float64, device=self.space.mesh.device)

    def assemble(self, form, coef=None, facet_mask=None, pec_flags=None, out=None, kernel=N.KERNEL_AUTO):
        out = self.new_values() if out is None else out
        with torch.cuda.device(self.space.mesh.device):
            N.check(N.lib.sfem_assemble(self.handle, form, _ptr(coef), _ptr(facet_mask), _ptr(pec_flags), _ptr(out),
                                        kernel, _stream_ptr()))
        return out

    def assemble_ST(self, eps_r=None, mu_r=None, pec_flags=None, S=None, T=None, kernel=N.KERNEL_AUTO):
        S = self.new_values() if S is None else S
        T = self.new_values() if T is None else T
        with torch.cuda.device(self.space.mesh.device):
            N.check(N.lib.sfem_assemble_ST(self.handle, _ptr(eps_r), _ptr(mu_r), _ptr(pec_flags), _ptr(S), _ptr(T),
                                           kernel, _stream_ptr()))
        return S, T

    # ---- compact ABC matrix S_0: only the rows of boundary entities are stored (sfem_abc_*) ----
    def abc_info(self):
        """(dof rows, non-zeros) of the compact S_0 storage."""
        nr, nz = C.c_int64(), C.c_int64()
        with torch.cuda.device(self.space.mesh.device):
            N.check(N.lib.sfem_abc_info(self.handle, _stream_ptr(), C.byref(nr), C.byref(nz)))
        return nr.value, nz.value

    def new_abc_values(self):
        return torch.empty(self.abc_info()[1], dtype=torch.float64, device=self.space.mesh.device)

    def assemble_abc(self, facet_mask=None, pec_flags=None, out=None):
        """S_0 in compact boundary-row storage: equal, bit for bit, to the boundary rows of assemble(FORM_ABC)."""
        out = self.new_abc_values() if out is None else out
        with torch.cuda.device(self.space.mesh.device):
            N.check(N.lib.sfem_assemble_abc(self.handle, _ptr(facet_mask), _ptr(pec_flags), _ptr(out), _stream_ptr()))
        return out

    def abc_csr(self, with_cols=True):
        """(rows int64 [n] global dof numbers, rowptr int64 [n+1], cols int32 [nnz] or None) of the compact storage."""
        dev = self.space.mesh.device
        nr, nz = self.abc_info()
        rows = torch.empty(nr, dtype=torch.int64, device=dev)
        rowptr = torch.empty(nr + 1, dtype=torch.int64, device=dev)
        cols = torch.empty(nz, dtype=torch.int32, device=dev) if with_cols else None
        with torch.cuda.device(dev):
            N.check(N.lib.sfem_abc_export(self.handle, _ptr(rows), _ptr(rowptr), _ptr(cols), _stream_ptr()))
        return rows, rowptr, cols

    def matvec(self, values, x, coefs=None, out=None, last_is_abc_compact=False):
        """y = sum_m coefs[m] * (A_m x) for value arrays that share this pattern (sfem_matvec).  x: device tensor
        [ndof], float64 or complex128; coefs: python complex / float per matrix (default 1).  last_is_abc_compact: the
        last array is a compact S_0 (assemble_abc)."""
        values = list(values)
        cplx = x.dtype == torch.complex128
        assert x.dtype in (torch.float64, torch.complex128) and x.is_contiguous()
        if out is None:
            out = torch.empty(self.n_rows, dtype=x.dtype, device=x.device)
        vp = (C.c_void_p * len(values))(*[v.data_ptr() for v in values])
        co = None
        if coefs is not None:
            flat = []
            for c in coefs:
                flat += [complex(c).real, complex(c).imag]
            co = (C.c_double * len(flat))(*flat)
        with torch.cuda.device(self.space.mesh.device):
            nm = len(values) | (N.MATS_LAST_IS_ABC_COMPACT if last_is_abc_compact else 0)
            N.check(N.lib.sfem_matvec(self.handle, nm, vp, co, int(cplx), _ptr(x), _ptr(out), _stream_ptr()))
        return out

    def diagonal(self, values, coefs=None):
        """Diagonal of sum_m coefs[m] A_m as a complex128 device tensor [n_rows] (sfem_diagonal)."""
        values = list(values)
        dev = self.space.mesh.device
        out = torch.empty(self.n_rows, dtype=torch.complex128, device=dev)
        vp = (C.c_void_p * len(values))(*[v.data_ptr() for v in values])
        co = None
        if coefs is not None:
            flat = []
            for c in coefs:
                flat += [complex(c).real, complex(c).imag]
            co = (C.c_double * len(flat))(*flat)
        with torch.cuda.device(dev):
            N.check(N.lib.sfem_diagonal(self.handle, len(values), vp, co, _ptr(out), _stream_ptr()))
        return out

    def bicgstab(self, values, coefs, b, x0=None, preconditioner='diagonal', rtol=1e-5, maxiter=None,
                 last_is_abc_compact=False):
        """Solve (sum_m coefs[m] A_m) x = b with BiCGStab on the device (sfem_bicgstab).  Returns (x, info, iterations,
        relres) with scipy's info convention (0 converged, > 0 = maxiter reached, < 0 breakdown)."""
        values = list(values)
        dev = self.space.mesh.device
        b = torch.as_tensor(b, device=dev).to(torch.complex128).contiguous()
        x = torch.zeros_like(b) if x0 is None else torch.as_tensor(x0, device=dev).to(torch.complex128).clone().contiguous()
        vp = (C.c_void_p * len(values))(*[v.data_ptr() for v in values])
        flat = []
        for c in coefs:
            flat += [complex(c).real, complex(c).imag]
        co = (C.c_double * len(flat))(*flat)
        pre = {None: 0, 'none': 0, 'diagonal': 1}[preconditioner.lower() if isinstance(preconditioner, str) else None]
        maxiter = min(10 * self.space.ndof if maxiter is None else int(maxiter), 2 ** 31 - 1)   # scipy's default; c_int
        it, rr = C.c_int(0), C.c_double(0.0)
        with torch.cuda.device(dev):
            nm = len(values) | (N.MATS_LAST_IS_ABC_COMPACT if last_is_abc_compact else 0)
            rc = N.lib.sfem_bicgstab(self.handle, nm, vp, co, _ptr(b), _ptr(x), pre, float(rtol), maxiter,
                                     C.byref(it), C.byref(rr), _stream_ptr())
        if rc == 6:
            return x, it.value, it.value, rr.value
        if rc == 5:
            return x, -10, it.value, rr.value
        N.check(rc)
        return x, 0, it.value, rr.value

    def apply_ident(self, flags, values):
        with torch.cuda.device(self.space.mesh.device):
            N.check(N.lib.sfem_apply_ident(self.handle, _ptr(flags), _ptr(values), _stream_ptr()))
        return values

    def __del__(self):
        h = getattr(self, 'handle', None)
        if h:
            N.lib.sfem_pattern_destroy(h)
            self.handle = None


def expand_material(regions, region_values, n_cells=None):
    """coef[c] = region_values[regions[c]] on the device; unknown region -> ValueError as Materials.py:147-148."""
    ids = np.array(sorted(region_values), dtype=np.uint32)
    vals = np.array([region_values[int(i)] for i in ids], dtype=np.float64)
    out = torch.empty(regions.shape[0], dtype=torch.float64, device=regions.device)
    bad = C.c_uint32(0)
    with torch.cuda.device(regions.device):
        rc = N.lib.sfem_expand_material(regions.shape[0], _ptr(regions), len(ids),
                                        ids.ctypes.data_as(C.POINTER(C.c_uint32)),
                                        vals.ctypes.data_as(C.POINTER(C.c_double)), _ptr(out), C.byref(bad),
                                        _stream_ptr())
    if rc == 4:
        raise ValueError('Material number %d not found' % bad.value)
    N.check(rc)
    return out


class HostAssemblyJob(object):
    """The whole path with HOST arrays in and out (sfem_host_job_*): numpy / CPU-tensor mesh, per-cell eps_r and mu_r
    and facet markers in; scipy-ready (rowptr, cols, S, T[, S0]) in page-locked host arrays out.  What a reference
    user gets from FunctionSpace + calc_system_matrices() + A.data()
    (/root/reference/sucemfem/SystemMatrices.py:46-58, /root/reference/sucemfem/Utilities/Converters.py:37).

    pec: None (no essential condition), 'exterior' (PEC walls: all exterior facets) or (markers uint32 [n_facets],
    region)."""

    def __init__(self, coords, cells, order, eps_r=None, mu_r=None, pec=None, pec_region=999):
        require_cuda()
        as_t = lambda a, dt: None if a is None else (a if isinstance(a, torch.Tensor) else torch.from_numpy(
            np.ascontiguousarray(a))).to(dt).contiguous()
        self._keep = [as_t(coords, torch.float64), as_t(cells, torch.int32), as_t(eps_r, torch.float64),
                      as_t(mu_r, torch.float64)]
        co, ce, ep, mu = self._keep
        mode, markers = 0, None
        if isinstance(pec, str):
            assert pec == 'exterior'
            mode = 1
        elif pec is not None:
            markers, pec_region = pec
            markers = torch.from_numpy(np.ascontiguousarray(np.asarray(markers, dtype=np.uint32)).view(np.int32))
            self._keep.append(markers)
            mode = 2
        hp = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
        h, nd, nnz = C.c_void_p(), C.c_int64(), C.c_int64()
        self.handle = None
        N.check(N.lib.sfem_host_job_create(int(co.shape[1]), co.shape[0], hp(co), ce.shape[0], hp(ce), int(order),
                                           hp(ep), hp(mu), mode, hp(markers), int(pec_region), _stream_ptr(),
                                           C.byref(nd), C.byref(nnz), C.byref(h)))
        self.handle, self.ndof, self.nnz = h, nd.value, nnz.value

    def run(self, want=('cols', 'S', 'T'), out=None, n_blocks=0):
        out = {} if out is None else out
        spec = {'rowptr': (self.ndof + 1, torch.int64), 'cols': (self.nnz, torch.int32), 'S': (self.nnz, torch.float64),
                'T': (self.nnz, torch.float64), 'S0': (self.nnz, torch.float64)}
        for k in ('rowptr',) + tuple(want):
            n, dt = spec[k]
            if k not in out or out[k].numel() != n or out[k].dtype != dt:
                out[k] = pinned_empty(n, dt)
        hp = lambda k: C.c_void_p(out[k].data_ptr()) if k in want or k == 'rowptr' else C.c_void_p(0)
        N.check(N.lib.sfem_host_job_run(self.handle, hp('rowptr'), hp('cols'), hp('S'), hp('T'), hp('S0'), int(n_blocks)))
        return out

    # ---- multi-GPU: every rank creates the job from the same mesh and streams only the rows it owns ----------------
    def partition_by_class(self, nparts):
        """(edge_bounds, other_bounds) of the job's space: see DeviceSpace.partition_by_class."""
        sp = C.c_void_p()
        N.check(N.lib.sfem_host_job_space(self.handle, C.byref(sp)))
        eb, ob = (C.c_int64 * (nparts + 1))(), (C.c_int64 * (nparts + 1))()
        N.check(N.lib.sfem_partition_rows_by_class(sp, nparts, eb, ob, _stream_ptr()))
        return [int(x) for x in eb], [int(x) for x in ob]

    def rows_nnz(self, ent_lo, ent_hi):
        sp, nnz = C.c_void_p(), C.c_int64()
        N.check(N.lib.sfem_host_job_space(self.handle, C.byref(sp)))
        N.check(N.lib.sfem_space_rows_nnz(sp, int(ent_lo), int(ent_hi), _stream_ptr(), C.byref(nnz)))
        return nnz.value

    def run_rows(self, ent_lo, ent_hi, want=('cols', 'S', 'T'), out=None, n_blocks=0, order=None):
        """The rows of the entities [ent_lo, ent_hi) only; rowptr is relative to the range."""
        out = {} if out is None else out
        q = self.ndof // self._n_entities() if order is None else order
        nnz = self.rows_nnz(ent_lo, ent_hi)
        spec = {'rowptr': (q * (ent_hi - ent_lo) + 1, torch.int64), 'cols': (nnz, torch.int32),
                'S': (nnz, torch.float64), 'T': (nnz, torch.float64), 'S0': (nnz, torch.float64)}
        for k in ('rowptr',) + tuple(want):
            n, dt = spec[k]
            if k not in out or out[k].numel() != n or out[k].dtype != dt:
                out[k] = pinned_empty(n, dt)
        hp = lambda k: C.c_void_p(out[k].data_ptr()) if k in want or k == 'rowptr' else C.c_void_p(0)
        N.check(N.lib.sfem_host_job_run_rows(self.handle, int(ent_lo), int(ent_hi), hp('rowptr'), hp('cols'), hp('S'),
                                             hp('T'), hp('S0'), int(n_blocks)))
        return out

    def _n_entities(self):
        sp = C.c_void_p()
        N.check(N.lib.sfem_host_job_space(self.handle, C.byref(sp)))
        nd, ne, nf, ndc = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int()
        N.check(N.lib.sfem_space_info(sp, C.byref(nd), C.byref(ne), C.byref(nf), C.byref(ndc)))
        eb, ob = (C.c_int64 * 2)(), (C.c_int64 * 2)()
        N.check(N.lib.sfem_partition_rows_by_class(sp, 1, eb, ob, _stream_ptr()))
        return int(ob[1])

    def close(self):
        if self.handle:
            N.lib.sfem_host_job_destroy(self.handle)
            self.handle = None

    __del__ = close


def assemble_host(coords, cells, order, eps_r=None, mu_r=None, pec=None, want=('cols', 'S', 'T'), out=None, n_blocks=0):
    """One call: host mesh + coefficients in -> host CSR out (see HostAssemblyJob)."""
    job = HostAssemblyJob(coords, cells, order, eps_r, mu_r, pec)
    try:
        return job.run(want, out, n_blocks)
    finally:
        job.close()
