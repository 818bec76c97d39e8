"""TEST INFRASTRUCTURE -- CPU oracle, never imported by the product path.

Restatement of the mesh generators the reference's hot path is fed from (DOLFIN 1.0
``UnitCube`` / ``UnitSquare`` as used at /root/reference/examples/resonant_cavity/driver.py:34
and /root/reference/examples/2D_waveguide/eigenmode/driver.py:38), of
``get_centred_cube`` (/root/reference/sucemfem/Utilities/MeshGenerators.py:26-53) and of the
``InscribedTet`` test mesh (/root/reference/sucemfem/Testing/Meshes.py:25-48).
Conventions: SURVEY.md section 8.1-M (verified there against the shipped DOLFIN-XML meshes).
"""
import math
import xml.etree.ElementTree as ET

import numpy as np

# six tets per cube, as corner numbers of the cube (x fastest), before per-cell vertex sorting
_CUBE_TETS = ((0, 1, 3, 7), (0, 1, 7, 5), (0, 5, 7, 4), (0, 3, 2, 7), (0, 6, 4, 7), (0, 2, 6, 7))


def unit_cube(nx, ny=None, nz=None):
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    ix, iy, iz = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), np.arange(nz + 1), indexing='ij')
    # vertex id = ix + (nx+1) iy + (nx+1)(ny+1) iz  -> order 'F' over (ix,iy,iz)
    coords = np.stack([ix.ravel(order='F') / nx, iy.ravel(order='F') / ny, iz.ravel(order='F') / nz], axis=1)
    cx, cy, cz = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing='ij')
    v0 = (cx + (nx + 1) * cy + (nx + 1) * (ny + 1) * cz).ravel(order='F')
    sx, sy, sz = 1, nx + 1, (nx + 1) * (ny + 1)
    corner = np.stack([v0, v0 + sx, v0 + sy, v0 + sy + sx, v0 + sz, v0 + sz + sx, v0 + sz + sy,
                       v0 + sz + sy + sx], axis=1)
    cells = np.stack([corner[:, list(t)] for t in _CUBE_TETS], axis=1).reshape(-1, 4)
    cells = np.sort(cells, axis=1).astype(np.int32)
    return coords.astype(np.float64), cells


def unit_square(nx, ny=None):
    """Right-diagonal UnitSquare ([UNPINNED] in SURVEY 8.1-M: no reference fixture stores a 2-D mesh)."""
    ny = nx if ny is None else ny
    ix, iy = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), indexing='ij')
    coords = np.stack([ix.ravel(order='F') / nx, iy.ravel(order='F') / ny], axis=1)
    cx, cy = np.meshgrid(np.arange(nx), np.arange(ny), indexing='ij')
    v0 = (cx + (nx + 1) * cy).ravel(order='F')
    v1, v2, v3 = v0 + 1, v0 + nx + 1, v0 + nx + 2
    cells = np.stack([np.stack([v0, v1, v3], 1), np.stack([v0, v2, v3], 1)], axis=1).reshape(-1, 3)
    return coords.astype(np.float64), np.sort(cells, axis=1).astype(np.int32)


def centred_cube(domain_size, max_edge_len):
    """get_centred_cube without the optional recentring (MeshGenerators.py:32-36)."""
    domain_size = np.atleast_1d(np.asarray(domain_size, dtype=np.float64))
    if domain_size.size == 1:
        domain_size = np.repeat(domain_size, 3)
    n = np.ceil(np.sqrt(2) * domain_size / max_edge_len).astype(int)
    coords, cells = unit_cube(int(n[0]), int(n[1]), int(n[2]))
    coords = coords * domain_size
    coords = coords - domain_size / 2
    return coords, cells


def inscribed_tet():
    c = np.array([[-1. / 2, 1. / 2, -1. / 2], [1. / 2, 1. / 2, 1. / 2], [1. / 2, -1. / 2, -1. / 2],
                  [-1. / 2, -1. / 2, 1. / 2], [1. / 6, 1. / 6, -1. / 6], [-1. / 6, 1. / 6, 1. / 6],
                  [-1. / 6, -1. / 6, -1. / 6], [1. / 6, -1. / 6, 1. / 6]], np.float64)
    e = np.array([[1, 4, 6, 7], [1, 3, 5, 7], [3, 4, 7, 8], [1, 2, 5, 6], [2, 3, 5, 8], [2, 4, 6, 8],
                  [4, 6, 7, 8], [1, 5, 6, 7], [3, 5, 7, 8], [2, 5, 6, 8], [5, 6, 7, 8]], np.int32) - 1
    return c, e


def read_dolfin_xml(path):
    """Minimal DOLFIN-XML mesh reader (vertices + cells), enough for the shipped fixtures."""
    root = ET.parse(path).getroot()
    mesh = root.find('mesh')
    dim = int(mesh.get('dim'))
    vs = mesh.find('vertices')
    coords = np.zeros((int(vs.get('size')), dim))
    for v in vs:
        coords[int(v.get('index'))] = [float(v.get(a)) for a in ('x', 'y', 'z')[:dim]]
    cs = mesh.find('cells')
    cells = np.zeros((int(cs.get('size')), dim + 1), np.int32)
    for c in cs:
        cells[int(c.get('index'))] = [int(c.get('v%d' % i)) for i in range(dim + 1)]
    return coords, cells


def read_dolfin_xml_meshfunction(path):
    root = ET.parse(path).getroot()
    mf = root.find('meshfunction')
    if mf is None:  # newer "mesh_function" spelling
        mf = root.find('mesh_function')
    out = np.zeros(int(mf.get('size')), np.uint32)
    for e in mf:
        out[int(e.get('index'))] = int(e.get('value'))
    return int(mf.get('dim')), out
