"""Mesh containers <-> the list-mesh dict of the reference (mirror of /root/reference/sucemfem/Utilities/MeshConverters.py:24-68)."""
import numpy as N

from .. import dolfin_compat as dolfin


def listmesh_2_dolfin_mesh(listmesh, reorder=False):
    """Mesh from listmesh['Nodes'] / listmesh['ElementNodes'].  The reference's reorder=True calls mesh.order() (UFC
    ordering of each cell's vertices); the device path requires that ordering, so cells are always stored sorted."""
    return dolfin.Mesh(coordinates=N.asarray(listmesh['Nodes'], dtype=N.float64),
                       cells=N.asarray(listmesh['ElementNodes'], dtype=N.int32))


def dolfin_mesh_2_listmesh(dolfin_mesh):
    """Only 'Nodes' (vertex coordinates) and 'ElementNodes' (tetrahedron node numbers) are set, as in the reference."""
    listmesh = {}
    listmesh['Nodes'] = dolfin_mesh.coordinates().copy()
    listmesh['ElementNodes'] = N.array(dolfin_mesh.cells(), dtype=N.int32, copy=True)
    return listmesh


def femmesh_reader_2_dolfin_mesh(femmesh_reader, reorder=True):
    """Convert a MeshIO.FemmeshReader object to a mesh"""
    if not femmesh_reader.read:
        femmesh_reader.read_meshfile()
    listmesh = dict(
        FemmeshFilename=femmesh_reader.get_mesh_filename(),
        FemmeshDir=femmesh_reader.get_mesh_dirname(),
        Nodes=femmesh_reader.get_nodes(),
        ElementNodes=femmesh_reader.get_tet_nodes())
    return listmesh_2_dolfin_mesh(listmesh, reorder=reorder)
