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
