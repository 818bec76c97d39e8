"""TEST INFRASTRUCTURE -- bench-size parity checker, never imported by the product path.

The oracle cannot assemble 10^7..10^8 tetrahedra in seconds, but it can (a) number the whole mesh in one O(C) pass
(cport.number_entities_hash: DOLFIN's first-appearance numbering) and (b) run the reference algorithm for a SAMPLE of
rows (cport.assemble_rows: sorted union of the incident cells' dofs, element tensors added in ascending cell order,
row-only ident).  This module holds the host half of the comparison used by tests/test_gpu_benchsize.py and by
`bench.py --check`: the structured mesh, the sample and the verdict; the caller supplies the rows the CUDA path
produced (gathered from the device or read from its host output arrays).

Contract checked (the same as /root/reference/sucemfem/ProblemConfigurations/tests/test_EMDrivenProblem.py:64-72 applies
to its 24 x 24 fixture, here per sampled row): pattern columns bit-exact, values |a - b| <= rtol |b| + atol with
rtol = 1e-12 (BASELINE.json north_star) and atol = 1e-15 max|B| (the reference test's own absolute floor, scaled).
"""
import numpy as np

from . import cport, meshes

RTOL = 1e-12
ATOL_REL = 1e-15


class CubeOracle(object):
    """UnitCube(n) order 2: mesh + first-appearance numbering on the host (seconds at n = 119, ~30 s at n = 203)."""

    def __init__(self, n):
        self.n = n
        self.coords, self.cells = meshes.unit_cube(n)
        self.cell_edges, self.cell_faces, self.E, self.F = cport.number_entities_hash(
            self.cells, len(self.coords), 7 * n ** 3 + 9 * n * n + 3 * n, 12 * n ** 3 + 6 * n * n)
        self.ndof = 2 * self.E + 2 * self.F

    def pec_flags_of(self, rows):
        """uint8 per sampled dof row: 1 if the dof's entity lies on the boundary of the cube (PEC on all exterior
        facets: the closure of every exterior face).  Decided geometrically, independent of the device's
        facet marking: an edge / face is on the boundary iff all its vertices share a coordinate plane x|y|z = 0|n."""
        rows = np.asarray(rows, np.int64)
        ent = rows // 2
        out = np.zeros(len(rows), np.uint8)
        is_edge = ent < self.E
        ev, fv = self._entity_vertices()
        n1 = self.n + 1

        def on_boundary(v):          # v [m, k] vertex ids
            ix, iy, iz = v % n1, (v // n1) % n1, v // (n1 * n1)
            res = np.zeros(len(v), bool)
            for c in (ix, iy, iz):
                res |= np.all(c == 0, axis=1) | np.all(c == self.n, axis=1)
            return res
        out[is_edge] = on_boundary(ev[ent[is_edge]])
        out[~is_edge] = on_boundary(fv[ent[~is_edge] - self.E])
        return out

    def _entity_vertices(self):
        if not hasattr(self, '_ev'):
            from . import refelem as R
            ev = np.zeros((self.E, 2), np.int64)
            fv = np.zeros((self.F, 3), np.int64)
            for l, e in enumerate(R.TET_EDGES):
                ev[self.cell_edges[:, l]] = self.cells[:, list(e)]
            for l, f in enumerate(R.TET_FACES):
                fv[self.cell_faces[:, l]] = self.cells[:, list(f)]
            self._ev, self._fv = ev, fv
        return self._ev, self._fv

    def rows(self, rows, eps=None, mu=None, pec=None):
        """(rowptr, cols, S, M) of the sampled dof rows by the reference algorithm."""
        return cport.assemble_rows(self.coords, self.cells, self.cell_edges, self.cell_faces, self.E, self.F, rows, eps,
                                   mu, pec)

    def entity_rowptr(self, threads=0):
        """Entity-level row pointer of ALL rows (count pass of the pattern builder over the [C][10] entity table)."""
        ent = np.concatenate([self.cell_edges, self.cell_faces + np.int32(self.E)], axis=1)
        ent = np.ascontiguousarray(ent, np.int32)
        n_ent = self.E + self.F
        rp = np.empty(n_ent + 1, np.int64)
        import ctypes as C
        lib = cport._lib()
        lib.oc_pattern_mt(cport._p(ent), C.c_int64(len(ent)), 10, C.c_int64(n_ent), cport._p(rp), None, int(threads))
        return rp


def pick_rows(rowptr_of, n_rows, nnz, n_random=50000, seed=0, row_lo=0):
    """Sample of dof rows [row_lo, row_lo + n_rows): head, tail, a window of rows either side of every multiple of 2^28
    in the value index (covers the 2^31 and 2^32 nnz offsets and every 2 GiB byte offset), and random entities (both of
    their rows).  rowptr_of(v) -> index of the row whose value range contains value index v (caller: searchsorted on the
    CUDA path's row pointer)."""
    rng = np.random.default_rng(seed)
    picks = [np.arange(0, min(64, n_rows)), np.arange(max(0, n_rows - 64), n_rows)]
    v = 1 << 28
    while v < nnz:
        r = int(rowptr_of(v))
        picks.append(np.arange(max(0, r - 8), min(n_rows, r + 8)))
        v += 1 << 28
    ent = rng.integers(0, n_rows // 2, n_random)
    picks += [2 * ent, 2 * ent + 1]
    return np.unique(np.concatenate(picks)) + row_lo


def compare(want, got_cols, got_S, got_T, label=''):
    """want = (rowptr, cols, S, M) from CubeOracle.rows; got_*: the same rows concatenated, from the CUDA path.
    Raises AssertionError with the first offending row; returns a small summary dict."""
    rp, cols, S, M = want
    assert len(got_cols) == len(cols), '%s: sampled rows hold %d entries, oracle %d' % (label, len(got_cols), len(cols))
    assert np.array_equal(np.asarray(got_cols, np.int64), cols), '%s: pattern columns differ' % label
    worst = 0.0
    for name, a, b in (('S', np.asarray(got_S), S), ('T', np.asarray(got_T), M)):
        tol = RTOL * np.abs(b) + ATOL_REL * np.abs(b).max()
        err = np.abs(a - b)
        bad = np.nonzero(err > tol)[0]
        assert len(bad) == 0, '%s: %s differs at sampled entry %d: %r vs %r' % (label, name, bad[0], a[bad[0]], b[bad[0]])
        worst = max(worst, float((err / tol).max()))
    # max_err_over_tol: worst |got - want| / (1e-12 |want| + 1e-15 max|want|) over the sample; parity holds iff <= 1
    return {"rows": int(len(rp) - 1), "entries": int(len(cols)), "max_err_over_tol": worst}


# ---- device half (torch tensors in, numpy out): shared by tests/test_gpu_benchsize.py and bench.py --check ----------
def gather_rows_torch(rowptr, rows_rel, arrays):
    """Concatenated entries of the block-relative dof rows `rows_rel` (ascending) of each device array in `arrays`;
    also returns the row lengths."""
    import torch
    r = torch.from_numpy(np.asarray(rows_rel, np.int64)).to(rowptr.device)
    a, b = rowptr[r], rowptr[r + 1]
    ln = b - a
    start = torch.repeat_interleave(a, ln)
    within = torch.arange(int(ln.sum()), device=rowptr.device) - torch.repeat_interleave(torch.cumsum(ln, 0) - ln, ln)
    idx = start + within
    return [x[idx].cpu().numpy() for x in arrays], ln.cpu().numpy()


def check_block(orc, rowptr, cols, S, T, row_begin, n_rows, nnz, flags, eps, mu, n_random, label, seed=0,
                flags_row_shift=0):
    """Compare a sample of the rows of one assembled row block (device tensors: block-relative rowptr, cols, S, T; flags =
    the device's uint8 PEC row flags or None) with the oracle.  eps, mu: host arrays.  Returns the summary dict."""
    import torch
    assert int(rowptr[0]) == 0 and int(rowptr[-1]) == nnz, '%s: row pointer ends do not match nnz' % label
    find = lambda v: int(torch.searchsorted(rowptr, torch.tensor([v], device=rowptr.device), right=True)) - 1
    rows = pick_rows(find, n_rows, nnz, n_random=n_random, seed=seed, row_lo=row_begin)
    (gc, gs, gt), ln = gather_rows_torch(rowptr, rows - row_begin, (cols, S, T))
    pec = None
    if flags is not None:
        pec = orc.pec_flags_of(rows)
        got = flags[torch.from_numpy(rows - flags_row_shift).to(flags.device)].cpu().numpy()
        assert np.array_equal(pec, got), '%s: PEC row flags differ from the geometric rule' % label
    want = orc.rows(rows, eps, mu, pec)
    assert np.array_equal(np.diff(want[0]), ln), '%s: row lengths differ' % label
    info = compare(want, gc, gs, gt, label)
    assert info['rows'] == len(rows)
    return info
