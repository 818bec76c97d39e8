"""TEST INFRASTRUCTURE -- ctypes wrapper of oracle/cpu_assemble.c (the C restatement of the reference
algorithm; CPU baseline of bench.py and the oracle for mid-size meshes).  Never used by the product path."""
import ctypes as C
import os
import subprocess

import numpy as np

from . import refelem as R

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'liboracle_cpu.so')


def _lib():
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, 'cpu_assemble.c')):
        subprocess.check_call(['make', '-C', _HERE])
    lib = C.CDLL(_SO)
    lib.oc_pattern.restype = C.c_int64
    lib.oc_pattern_mt.restype = C.c_int64
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class CPUAssembler(object):
    """Numbering + pattern once (untimed set-up), then assemble() = the timed DOLFIN-style cell loop."""

    def __init__(self, coords, cells, order, threads=1):
        self.lib = _lib()
        self.coords = np.ascontiguousarray(coords, np.float64)
        self.cells = np.ascontiguousarray(cells, np.int32)
        self.order = order
        Cn, V = self.cells.shape[0], self.coords.shape[0]
        self.cell_edges = np.empty((Cn, 6), np.int32)
        self.cell_faces = np.empty((Cn, 4), np.int32)
        nE, nF = C.c_int64(), C.c_int64()
        self.lib.oc_number_entities(_p(self.cells), C.c_int64(Cn), C.c_int64(V), _p(self.cell_edges),
                                    _p(self.cell_faces), C.byref(nE), C.byref(nF))
        self.E, self.F = nE.value, nF.value
        self.nd = 6 if order == 1 else 20
        self.ndof = self.E if order == 1 else 2 * self.E + 2 * self.F
        self.cd = np.empty((Cn, self.nd), np.int32)
        self.lib.oc_cell_dofs(_p(self.cell_edges), _p(self.cell_faces), C.c_int64(Cn), C.c_int64(self.E), order,
                              _p(self.cd))
        self.rowptr = np.empty(self.ndof + 1, np.int64)
        nnz = self.lib.oc_pattern_mt(_p(self.cd), C.c_int64(Cn), self.nd, C.c_int64(self.ndof), _p(self.rowptr), None,
                                     threads)
        self.cols = np.empty(nnz, np.int32)
        self.lib.oc_pattern_mt(_p(self.cd), C.c_int64(Cn), self.nd, C.c_int64(self.ndof), _p(self.rowptr), _p(self.cols),
                               threads)
        self.Mhat = np.ascontiguousarray(R.mass_tensor(3, order))
        self.Chat = np.ascontiguousarray(R.curl_tensor(3, order))

    def max_threads(self):
        return self.lib.oc_max_threads()

    def assemble(self, eps=None, mu=None, threads=1, S=None, M=None):
        S = np.zeros(len(self.cols)) if S is None else S
        M = np.zeros(len(self.cols)) if M is None else M
        eps = None if eps is None else np.ascontiguousarray(eps, np.float64)
        mu = None if mu is None else np.ascontiguousarray(mu, np.float64)
        self.lib.oc_assemble_cells(_p(self.coords), _p(self.cells), C.c_int64(self.cells.shape[0]), _p(self.cd),
                                   self.nd, _p(self.Mhat), _p(self.Chat), _p(eps), _p(mu), _p(self.rowptr),
                                   _p(self.cols), _p(S), _p(M), threads)
        return S, M

    def apply_ident(self, vals, rows):
        rows = np.ascontiguousarray(rows, np.int64)
        self.lib.oc_apply_ident(_p(self.rowptr), _p(self.cols), _p(vals), _p(rows), C.c_int64(len(rows)))
        return vals


# ---- bench-size oracle: O(C) first-appearance numbering + the reference algorithm restricted to sampled rows --------
def number_entities_hash(cells, n_vertices, expected_edges=0, expected_faces=0):
    """(cell_edges [C,6], cell_faces [C,4], E, F): one serial pass over the cells with a hash map (what DOLFIN's
    mesh.init does); same numbering as CPUAssembler / oracle.fem.number_entities."""
    lib = _lib()
    cells = np.ascontiguousarray(cells, np.int32)
    Cn = cells.shape[0]
    ce, cf = np.empty((Cn, 6), np.int32), np.empty((Cn, 4), np.int32)
    nE, nF = C.c_int64(), C.c_int64()
    rc = lib.oc_number_entities_hash(_p(cells), C.c_int64(Cn), C.c_int64(int(n_vertices)), C.c_int64(int(expected_edges)),
                                     C.c_int64(int(expected_faces)), _p(ce), _p(cf), C.byref(nE), C.byref(nF))
    assert rc == 0, 'hash numbering failed (out of memory or too small a table)'
    return ce, cf, nE.value, nF.value


def assemble_rows(coords, cells, cell_edges, cell_faces, E, F, rows, eps=None, mu=None, pec=None):
    """Order-2 tets: pattern rows and S, M values of the dof rows `rows` (ascending global dof numbers) by the reference
    algorithm (sorted union of the incident cells' dofs; element tensors added in ascending cell order; unit rows where
    pec[i] != 0).  Returns (rowptr [n+1], cols int64, S, M) over the sample."""
    lib = _lib()
    lib.oc_assemble_rows.restype = C.c_int64
    coords = np.ascontiguousarray(coords, np.float64)
    cells = np.ascontiguousarray(cells, np.int32)
    rows = np.ascontiguousarray(rows, np.int64)
    assert np.all(np.diff(rows) > 0)
    Mhat = np.ascontiguousarray(R.mass_tensor(3, 2))
    Chat = np.ascontiguousarray(R.curl_tensor(3, 2))
    eps = None if eps is None else np.ascontiguousarray(eps, np.float64)
    mu = None if mu is None else np.ascontiguousarray(mu, np.float64)
    pec = None if pec is None else np.ascontiguousarray(pec, np.uint8)
    rp = np.empty(len(rows) + 1, np.int64)
    args = (_p(coords), _p(cells), C.c_int64(cells.shape[0]), _p(cell_edges), _p(cell_faces), C.c_int64(E), C.c_int64(F),
            _p(Mhat), _p(Chat), _p(eps), _p(mu), C.c_int64(len(rows)), _p(rows), _p(pec), _p(rp))
    nnz = lib.oc_assemble_rows(*args, None, None, None)
    cols, S, M = np.empty(nnz, np.int64), np.empty(nnz), np.empty(nnz)
    lib.oc_assemble_rows(*args, _p(cols), _p(S), _p(M))
    return rp, cols, S, M
