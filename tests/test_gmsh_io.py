"""gmsh reader against the .msh / .xml pairs the reference ships: the XML files are the output of
the reference's converter (dolfin-convert, /root/reference/sandbox/pec_labels/pec_labels.py) for exactly these .msh
files, so vertex numbering, cells, physical regions and facet markers must come out bit-exact (coordinates to the
XML writer's print precision)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'sucem-fem_b200')):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLD = np.load(os.path.join(ROOT, 'tests', 'golden', 'gmsh_pairs.npz'))
NAMES = ['sample_mesh', 'lee_mittra92_fig6c', 'lee_mittra92_fig6b', 'rect_white', 'patch']


def _msh(tmp_path, name):
    p = os.path.join(str(tmp_path), name + '.msh')
    with open(p, 'wb') as f:
        f.write(GOLD[name + '__msh'].tobytes())
    return p


def _check_mesh(gm, name):
    co = GOLD[name + '__coordinates']
    assert gm.coordinates.shape == co.shape
    # dolfin's XML writer printed about 6 significant digits
    assert np.allclose(gm.coordinates, co, rtol=1e-5, atol=1e-5 * np.abs(co).max())
    assert np.array_equal(gm.cells, GOLD[name + '__cells'])
    if name + '__physical_region' in GOLD:
        assert np.array_equal(gm.physical_regions, GOLD[name + '__physical_region'])


@pytest.mark.parametrize('name', NAMES)
def test_reader_reproduces_reference_conversion(tmp_path, name):
    from sucemfem.Utilities import GmshIO
    from oracle import fem
    gm = GmshIO.read_gmsh(_msh(tmp_path, name))
    _check_mesh(gm, name)
    if name + '__facet_region' in GOLD:
        want = GOLD[name + '__facet_region']
        num = fem.number_entities(gm.cells)          # the checker's numbering (the product's is the device one)
        table = num['cell_faces'] if gm.dim == 3 else num['cell_edges']
        got = GmshIO.facet_values_by_index(gm.cells, table, gm.facet_nodes, gm.facet_values, len(want))
        assert np.array_equal(got, want)
    else:
        assert not gm.has_facet_regions()


def test_reader_errors(tmp_path):
    from sucemfem.Utilities import GmshIO
    p = os.path.join(str(tmp_path), 'bad.msh')
    with open(p, 'w') as f:
        f.write('$MeshFormat\n4.1 0 8\n$EndMeshFormat\n$Nodes\n0\n$EndNodes\n$Elements\n0\n$EndElements\n')
    with pytest.raises(GmshIO.GmshError):
        GmshIO.read_gmsh(p)
    with open(p, 'w') as f:
        f.write('$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n1\n1 0 0 0\n$EndNodes\n$Elements\n0\n$EndElements\n')
    with pytest.raises(GmshIO.GmshError, match='No elements'):
        GmshIO.read_gmsh(p)
    with open(p, 'w') as f:   # only a point element: no supported cell type
        f.write('$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n1\n1 0 0 0\n$EndNodes\n$Elements\n1\n1 15 2 0 1 1\n'
                '$EndElements\n')
    with pytest.raises(GmshIO.GmshError, match='supported type'):
        GmshIO.read_gmsh(p)
    with open(p, 'w') as f:   # a marked triangle with a node no cell uses
        f.write('$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n5\n1 0 0 0\n2 1 0 0\n3 0 1 0\n4 0 0 1\n5 1 1 1\n$EndNodes\n'
                '$Elements\n2\n1 2 2 7 1 1 2 5\n2 4 2 1 1 1 2 3 4\n$EndElements\n')
    with pytest.raises(GmshIO.GmshError, match='not found to mark'):
        GmshIO.read_gmsh(p)
    # a marked vertex triple that is no face of any cell fails when the markers are indexed
    cells = np.array([[0, 1, 2, 3], [1, 2, 3, 4]], np.int32)
    table = np.array([[0, 1, 2, 3], [4, 5, 6, 0]], np.int32)
    with pytest.raises(GmshIO.GmshError, match='not found to mark'):
        GmshIO.facet_values_by_index(cells, table, np.array([[0, 1, 4]], np.int32), np.array([7], np.uint32), 7)
    got = GmshIO.facet_values_by_index(cells, table, np.array([[1, 2, 3], [0, 1, 2]], np.int32),
                                       np.array([7, 9], np.uint32), 7)
    assert list(got) == [7, 0, 0, 9, 0, 0, 0]


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['sample_mesh', 'patch'])
def test_product_facet_functions_and_xml_round_trip(tmp_path, name):
    """Through the product API: device numbering behind the facet MeshFunction, and convert() writes files that
    read back as the reference's own XML files."""
    from sucemfem.Utilities import GmshIO
    from sucemfem import dolfin_compat as dolfin
    msh = _msh(tmp_path, name)
    xml = os.path.join(str(tmp_path), name + '.xml')
    gm, written = GmshIO.convert(msh, xml)
    mesh = dolfin.Mesh(xml)
    assert np.array_equal(mesh.cells(), GOLD[name + '__cells'])
    assert np.array_equal(mesh.coordinates(), gm.coordinates)          # repr() round-trips doubles exactly
    assert np.array_equal(dolfin.Mesh(msh).cells(), mesh.cells())
    base = os.path.splitext(xml)[0]
    mf = dolfin.MeshFunction('uint', mesh, base + '_facet_region.xml')
    assert mf.dim() == gm.dim - 1
    assert np.array_equal(mf.array(), GOLD[name + '__facet_region'])
    assert np.array_equal(gm.facet_region_function(mesh).array(), GOLD[name + '__facet_region'])
    if name == 'patch':
        assert base + '_physical_region.xml' in written
        pf = dolfin.MeshFunction('uint', mesh, base + '_physical_region.xml')
        assert np.array_equal(pf.array(), GOLD[name + '__physical_region'])
        assert np.array_equal(gm.physical_region_function(mesh).array(), GOLD[name + '__physical_region'])


@pytest.mark.gpu
@pytest.mark.parametrize('name,order', [('lee_mittra92_fig6b', 2), ('lee_mittra92_fig6c', 1)])
def test_dielectric_loaded_cavity_from_gmsh(tmp_path, name, order):
    """/root/reference/sandbox/eig/eig_driver.py:41-66 (the Albani-Bernardi / Lee-Mittra flow): gmsh mesh + its
    physical regions -> material regions {1000: eps_r=16, 1001: eps_r=1} -> EigenProblem.  S and M from the device
    kernels against the oracle on the same unstructured, inhomogeneous mesh: pattern bit-exact, values to 1e-12."""
    from oracle import fem
    from sucemfem.Utilities import GmshIO
    from sucemfem import dolfin_compat as dolfin
    from sucemfem.ProblemConfigurations.EMVectorWaveEigenproblem import EigenProblem
    gm = GmshIO.read_gmsh(_msh(tmp_path, name))
    mesh = dolfin.Mesh(_msh(tmp_path, name))
    material_mesh_func = gm.physical_region_function(mesh)
    materials = {1000: dict(eps_r=16), 1001: dict(eps_r=1)}
    ep = EigenProblem()
    ep.set_mesh(mesh)
    ep.set_basis_order(order)
    ep.set_material_regions(materials)
    ep.set_region_meshfunction(material_mesh_func)
    ep.init_problem()
    eps = np.where(gm.physical_regions == 1000, 16.0, 1.0)
    assert 0 < (eps == 16.0).sum() < len(eps)
    ref = fem.assemble_system(gm.coordinates, gm.cells, order, eps_r=eps)
    for key in ('S', 'M'):
        rp, cl, v = ep.system_matrices[key].data()
        assert np.array_equal(rp, ref['rowptr']) and np.array_equal(cl, ref['cols'])
        assert np.all(np.abs(v - ref[key]) <= 1e-12 * np.abs(ref[key]) + 1e-15 * np.abs(ref[key]).max())


class _OracleNumbering(object):
    """CPU stand-in for the device numbering behind dolfin_compat.Mesh._numbering() (test of host logic only)."""

    def __init__(self, cells):
        import torch
        from oracle import fem
        num = fem.number_entities(cells)
        self.n_edges = num['E']
        self.n_facets = num['F'] if num['dim'] == 3 else num['E']
        ce = torch.from_numpy(num['cell_edges'])
        self._ent = (ce, torch.from_numpy(num['cell_faces']) if num['dim'] == 3 else None)

    def cell_entities(self):
        return self._ent


@pytest.mark.parametrize('name', ['sample_mesh', 'patch'])
def test_convert_writes_the_reference_xml_files_cpu(tmp_path, monkeypatch, name):
    """GmshIO.convert() end to end on the CPU (the facet numbering comes from the oracle instead of the device): the
    written mesh / physical_region / facet_region XML files read back as the reference's own files."""
    from sucemfem.Utilities import GmshIO
    from sucemfem import dolfin_compat as dolfin
    monkeypatch.setattr(dolfin.Mesh, '_numbering', lambda self: _OracleNumbering(self.cells()))
    msh = _msh(tmp_path, name)
    xml = os.path.join(str(tmp_path), name + '.xml')
    gm, written = GmshIO.convert(msh, xml)
    base = os.path.splitext(xml)[0]
    assert xml in written and base + '_facet_region.xml' in written
    mesh = dolfin.Mesh(xml)
    assert np.array_equal(mesh.cells(), GOLD[name + '__cells']) and np.array_equal(mesh.coordinates(), gm.coordinates)
    mf = dolfin.MeshFunction('uint', mesh, base + '_facet_region.xml')
    assert mf.dim() == gm.dim - 1 and np.array_equal(mf.array(), GOLD[name + '__facet_region'])
    if name == 'patch':
        pf = dolfin.MeshFunction('uint', mesh, base + '_physical_region.xml')
        assert pf.dim() == 3 and np.array_equal(pf.array(), GOLD[name + '__physical_region'])
    else:
        assert base + '_physical_region.xml' not in written or gm.has_physical_regions()
