""".femmesh reader (same interface as /root/reference/sucemfem/Utilities/MeshIO.py:28-117).

File layout: `BLOCK <name>` / count line / one record per line / `ENDBLOCK`, 1-based numbering.  `nodes` records are
`number x y z`, `tets` records are `number property n1 n2 n3 n4`; `tris` blocks (and anything else) are skipped.  Records
may come in any order but every number 1..count must appear exactly once.  Tetrahedra are returned with ascending
node numbers (UFC order).  Each block is parsed as one numpy table.
"""
import os

import numpy as N

from . import MeshConverters


class UninitialisedException(Exception):
    pass


def _blocks(lines):
    """{block name: list of record lines} for every BLOCK ... ENDBLOCK section."""
    out, name, body = {}, None, []
    for raw in lines:
        l = raw.strip()
        if name is None:
            if l.startswith('BLOCK '):
                name, body = l[len('BLOCK '):].strip(), []
        elif l.split()[:1] == ['ENDBLOCK']:
            out[name] = body
            name = None
        elif l:
            body.append(l)
    if name is not None:
        raise ValueError('BLOCK %s is not closed by ENDBLOCK' % name)
    return out


def _table(body, n_columns, what):
    """The block's records as a float table ordered by record number (0-based row = number - 1)."""
    count = int(body[0].split()[0])
    tab = N.array([l.split() for l in body[1:]], dtype=N.float64).reshape(-1, n_columns)
    if len(tab) != count:
        raise AssertionError('%s block announces %d records and holds %d' % (what, count, len(tab)))
    number = tab[:, 0].astype(N.int64) - 1
    if count and not N.array_equal(N.sort(number), N.arange(count)):
        raise AssertionError('%s numbers must be exactly 1..%d' % (what, count))
    ordered = N.empty_like(tab)
    ordered[number] = tab
    return ordered


class FemmeshReader(object):
    read = False
    _DATA = ('nodes', 'tet_nodes', 'tet_property_nos')

    def __init__(self, mesh_filename):
        self.mesh_filename = mesh_filename

    def read_meshfile(self):
        with open(self.mesh_filename) as fo:
            blocks = _blocks(fo)
        if 'nodes' in blocks:
            t = _table(blocks['nodes'], 4, 'nodes')
            self.no_nodes, self.nodes = len(t), N.ascontiguousarray(t[:, 1:4])
        if 'tets' in blocks:
            t = _table(blocks['tets'], 6, 'tets')
            self.no_tets = len(t)
            self.tet_property_nos = t[:, 1].astype(N.int32)
            self.tet_nodes = N.sort(t[:, 2:6].astype(N.int32) - 1, axis=1)
        self.read = True

    def _need(self, name):
        if not hasattr(self, name):
            raise Exception("Call read_meshfile() method of %s first" % self.__class__)
        return getattr(self, name)

    def get_mesh_filename(self):
        self._need('nodes')
        return os.path.basename(self.mesh_filename)

    def get_mesh_dirname(self):
        self._need('nodes')
        return os.path.dirname(self.mesh_filename)

    def get_tet_nodes(self):
        return self._need('tet_nodes')

    def get_tet_property_nos(self):
        return self._need('tet_property_nos')

    def get_nodes(self):
        return self._need('nodes')


def femmesh_2_dolfin_mesh(femmesh_file):
    """Convert a femmesh file to a mesh object"""
    return MeshConverters.femmesh_reader_2_dolfin_mesh(FemmeshReader(femmesh_file))
