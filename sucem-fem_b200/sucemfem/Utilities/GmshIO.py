"""Gmsh .msh (ASCII format 2.x) reader and DOLFIN-XML writer.

The reference gets its unstructured meshes from gmsh through `dolfin-convert` / its own variant of it
(/root/reference/sandbox/pec_labels/pec_labels.py:27-290, `gmsh_to_dolfin_mesh`; facet marking `:371-420`,
`convert_and_create_facet_mesh_function`) and ships both the .msh and the converted .xml files
(functional_tests/data/meshes, sandbox/eig, sandbox/patch, sandbox/interior_integration).  Conventions reproduced
here (pinned by those file pairs, tests/test_gmsh_io.py):

  * only elements of the highest dimension present are cells (gmsh types 1 line, 2 triangle, 4 tetrahedron);
  * only the nodes those cells use become vertices, numbered in the order they appear in $Nodes;
  * cells keep the file order; every cell lists its vertices in ascending order (`mesh.order()`, pec_labels.py:377);
  * `physical_region` = first tag of each cell (written only if not all zero, pec_labels.py:262-268);
  * `facet_region` = first tag of every element of dimension dim-1, put on the facet with the same (sorted) vertex
    set; facets not mentioned keep 0 (pec_labels.py:389-412).  Facets are indexed by the first-appearance entity
    numbering of the mesh, which the function space builder computes on the device.
"""
import os

import numpy as np

GMSH_DIM = {1: 1, 2: 2, 4: 3}                    # gmsh element type -> topological dimension (pec_labels.py:50)
CELL_TYPE = {1: 'interval', 2: 'triangle', 3: 'tetrahedron'}


class GmshError(Exception):
    pass


class GmshMesh(object):
    """Result of read_gmsh(): host arrays in DOLFIN's conventions."""

    def __init__(self, dim, coordinates, cells, physical_regions, facet_nodes, facet_values):
        self.dim = dim
        self.coordinates = coordinates            # f64 [V, dim]
        self.cells = cells                        # i32 [C, dim+1], ascending per cell
        self.physical_regions = physical_regions  # u32 [C]
        self.facet_nodes = facet_nodes            # i32 [M, dim], ascending per facet (marked facets only)
        self.facet_values = facet_values          # u32 [M]

    def has_physical_regions(self):
        return bool(np.any(self.physical_regions != 0))

    def has_facet_regions(self):
        return bool(np.any(self.facet_values != 0))

    # -- the dolfin stand-ins of this package (device numbering behind mesh.num_facets()) ------------------------
    def mesh(self):
        from .. import dolfin_compat
        return dolfin_compat.Mesh(coordinates=self.coordinates, cells=self.cells)

    def physical_region_function(self, mesh):
        from .. import dolfin_compat
        mf = dolfin_compat.MeshFunction('uint', mesh, self.dim)
        mf.array()[:] = self.physical_regions
        return mf

    def facet_region_function(self, mesh):
        """MeshFunction('uint', mesh, dim-1) holding the gmsh physical numbers of the marked facets."""
        from .. import dolfin_compat
        mf = dolfin_compat.MeshFunction('uint', mesh, self.dim - 1)
        sp = mesh._numbering()
        ce, cf = sp.cell_entities()
        table = (cf if self.dim == 3 else ce).cpu().numpy()
        mf.array()[:] = facet_values_by_index(self.cells, table, self.facet_nodes, self.facet_values,
                                              len(mf.array()))
        return mf


def _pack(nodes, n_vertices):
    """One int64 key per ascending vertex tuple."""
    key = np.zeros(len(nodes), np.int64)
    for k in range(nodes.shape[1]):
        key = key * np.int64(n_vertices) + nodes[:, k].astype(np.int64)
    return key


def facet_values_by_index(cells, cell_facets, facet_nodes, facet_values, n_facets):
    """Scatter per-facet values given by vertex tuples into an array indexed by facet number.
    cell_facets[c, i] = number of the facet of cell c opposite its i-th vertex (UFC)."""
    cells = np.asarray(cells)
    nv = int(cells.max()) + 1 if cells.size else 1
    nloc = cells.shape[1]
    if nv ** (nloc - 1) >= 2 ** 63:
        raise GmshError('mesh too large for packed facet keys')
    keys, index = [], []
    for i in range(nloc):
        keys.append(_pack(np.delete(cells, i, axis=1), nv))
        index.append(np.asarray(cell_facets)[:, i])
    keys, index = np.concatenate(keys), np.concatenate(index)
    order = np.argsort(keys, kind='stable')
    keys, index = keys[order], index[order]
    out = np.zeros(n_facets, np.uint32)
    if len(facet_nodes) == 0:
        return out
    want = _pack(np.asarray(facet_nodes), nv)
    pos = np.searchsorted(keys, want)
    pos_c = np.minimum(pos, len(keys) - 1)
    found = (pos < len(keys)) & (keys[pos_c] == want)
    marked = np.asarray(facet_values) != 0
    if np.any(marked & ~found):
        i = int(np.nonzero(marked & ~found)[0][0])
        # same failure as pec_labels.py:408-409
        raise GmshError('The facet (%d) was not found to mark: %s' % (i, list(np.asarray(facet_nodes)[i])))
    sel = marked & found
    # later elements win, as repeated assignment in file order would
    out[index[pos_c[sel]]] = np.asarray(facet_values)[sel]
    return out


def _sections(path):
    with open(path, 'r') as f:
        lines = [l.strip() for l in f]
    sec, i = {}, 0
    while i < len(lines):
        l = lines[i]
        if l.startswith('$') and not l.startswith('$End'):
            name = l[1:]
            try:
                j = lines.index('$End' + name, i + 1)
            except ValueError:
                raise GmshError('section $%s is not closed' % name)
            sec.setdefault(name, lines[i + 1:j])
            i = j
        i += 1
    return sec


def read_gmsh(path):
    """Parse an ASCII gmsh 2.x file -> GmshMesh."""
    sec = _sections(path)
    if 'MeshFormat' not in sec or 'Nodes' not in sec or 'Elements' not in sec:
        raise GmshError('Missing data, unable to convert. Did you use version 2 of the gmsh file format?')
    version, file_type = sec['MeshFormat'][0].split()[:2]
    if not version.startswith('2') or int(file_type) != 0:
        raise GmshError('only ASCII gmsh format 2.x is supported (file says %s, type %s)' % (version, file_type))
    nodes = sec['Nodes']
    n_nodes = int(nodes[0])
    node_table = np.array([l.split() for l in nodes[1:1 + n_nodes] if l and l[0] != '#'], dtype=np.float64)
    if node_table.shape != (n_nodes, 4):
        raise GmshError('expected %d node lines' % n_nodes)
    node_no = node_table[:, 0].astype(np.int64)
    elems = sec['Elements']
    n_elems = int(elems[0])
    if n_elems == 0:
        raise GmshError('No elements found in gmsh file.')
    by_dim = {1: ([], []), 2: ([], []), 3: ([], [])}
    for l in elems[1:1 + n_elems]:
        e = l.split()
        etype, ntags = int(e[1]), int(e[2])
        if etype not in GMSH_DIM:
            continue                               # points, higher-order elements ...: ignored as in the reference
        d = GMSH_DIM[etype]
        by_dim[d][0].append([int(x) for x in e[3 + ntags:3 + ntags + d + 1]])
        by_dim[d][1].append(int(e[3]) if ntags > 0 else 0)
    dim = max([d for d in by_dim if by_dim[d][0]] or [0])
    if dim == 0:
        raise GmshError('Unable to find cells of supported type.')
    cell_nodes = np.array(by_dim[dim][0], np.int64)
    # vertices = nodes used by the cells, in $Nodes order
    lookup = -np.ones(int(max(node_no.max(), cell_nodes.max())) + 1, np.int64)
    used = np.zeros(len(lookup), bool)
    used[cell_nodes.ravel()] = True
    keep = used[node_no]
    lookup[node_no[keep]] = np.arange(int(keep.sum()))
    if np.any(lookup[cell_nodes] < 0):
        c = int(np.nonzero((lookup[cell_nodes] < 0).any(axis=1))[0][0])
        raise GmshError('Vertex of %s %d not previously defined.' % (CELL_TYPE[dim], c))
    coordinates = np.ascontiguousarray(node_table[keep][:, 1:1 + dim])
    cells = np.sort(lookup[cell_nodes], axis=1).astype(np.int32)
    regions = np.array(by_dim[dim][1], np.uint32)
    fn, fv = np.zeros((0, dim), np.int32), np.zeros(0, np.uint32)
    if dim > 1 and by_dim[dim - 1][0]:
        raw = np.array(by_dim[dim - 1][0], np.int64)
        val = np.array(by_dim[dim - 1][1], np.uint32)
        inside = (raw < len(lookup)).all(axis=1)
        mapped = np.where(inside[:, None], lookup[np.minimum(raw, len(lookup) - 1)], -1)
        ok = (mapped >= 0).all(axis=1)
        if np.any(~ok & (val != 0)):
            i = int(np.nonzero(~ok & (val != 0))[0][0])
            raise GmshError('The facet (%d) was not found to mark: %s' % (i, list(raw[i])))
        fn = np.sort(mapped[ok], axis=1).astype(np.int32)
        fv = val[ok]
    return GmshMesh(dim, coordinates, cells, regions, fn, fv)


# ---- DOLFIN XML writers (the files dolfin-convert leaves behind: x.xml, x_physical_region.xml, x_facet_region.xml) --
def write_dolfin_xml_mesh(path, coordinates, cells):
    dim = cells.shape[1] - 1
    with open(path, 'w') as f:
        f.write('<?xml version="1.0" encoding="UTF-8"?>\n\n<dolfin xmlns:dolfin="http://www.fenicsproject.org">\n')
        f.write('  <mesh celltype="%s" dim="%d">\n    <vertices size="%d">\n' % (CELL_TYPE[dim], dim, len(coordinates)))
        for i, x in enumerate(coordinates):
            f.write('      <vertex index="%d" %s/>\n' % (i, ' '.join('%s="%s"' % (a, repr(float(v)))
                                                                        for a, v in zip('xyz', x))))
        f.write('    </vertices>\n    <cells size="%d">\n' % len(cells))
        for i, c in enumerate(cells):
            f.write('      <%s index="%d" %s/>\n' % (CELL_TYPE[dim], i, ' '.join('v%d="%d"' % (k, v)
                                                                                  for k, v in enumerate(c))))
        f.write('    </cells>\n  </mesh>\n</dolfin>\n')


def write_dolfin_xml_meshfunction(path, dim, values):
    with open(path, 'w') as f:
        f.write('<?xml version="1.0" encoding="UTF-8"?>\n<dolfin xmlns:dolfin="http://www.fenics.org/dolfin/">\n')
        f.write('  <meshfunction type="uint" dim="%d" size="%d">\n' % (dim, len(values)))
        for i, v in enumerate(values):
            f.write('    <entity index="%d" value="%d"/>\n' % (i, int(v)))
        f.write('  </meshfunction>\n</dolfin>\n')


def convert(ifilename, ofilename):
    """gmsh -> DOLFIN XML, like `dolfin-convert x.msh x.xml`: writes ofilename and, when the gmsh file defines them,
    <base>_physical_region.xml and <base>_facet_region.xml.  Returns (GmshMesh, [files written])."""
    gm = read_gmsh(ifilename)
    base = os.path.splitext(ofilename)[0]
    written = [ofilename]
    write_dolfin_xml_mesh(ofilename, gm.coordinates, gm.cells)
    if gm.has_physical_regions():
        write_dolfin_xml_meshfunction(base + '_physical_region.xml', gm.dim, gm.physical_regions)
        written.append(base + '_physical_region.xml')
    if gm.has_facet_regions():
        mf = gm.facet_region_function(gm.mesh())
        write_dolfin_xml_meshfunction(base + '_facet_region.xml', gm.dim - 1, mf.array())
        written.append(base + '_facet_region.xml')
    return gm, written
