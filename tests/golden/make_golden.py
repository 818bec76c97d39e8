"""Convert the reference's own golden fixtures for the assembly path into small .npz files.

Run HERE (the build container), where /root/reference exists; the GPU box only sees the .npz files.
Sources (fixtures 1,2,3,5,7) -- Python-2 pickles holding only numpy arrays and
scalars, loaded with encoding='latin1':

  1,2  sucemfem/ProblemConfigurations/tests/LHS_matrix.npy, RHS_vector.pickle
       (test: sucemfem/ProblemConfigurations/tests/test_EMDrivenProblem.py:64-82)
  3    sucemfem/Sources/tests/fillament_source_test_data.pickle
       (test: sucemfem/Sources/tests/test_fillament_source.py:62-69)
  5    sucemfem/PostProcessing/tests/data/power_flux_reference_dofs_...pickle
       (generator: sucemfem/PostProcessing/tests/data/power_flux_generate_dofs.py:44-92)
  7    sucemfem/PostProcessing/tests/data/reference_dofs-2-0.149896229-0.0499654096667.pickle
       (generator: sucemfem/PostProcessing/tests/data/ntff_generate_dofs.py:40-85)
  6    functional_tests/data/meshes/rectangular_prism.xml + rectangular_prism_facet_region.xml (gmsh mesh of the
       1 x 0.5 x 0.25 cavity and its PEC facet markers, indexed by DOLFIN's face numbering;
       test: functional_tests/test_EigenProblems.py:69-134)
  f3   gmsh .msh files shipped together with their dolfin-convert XML output (sandbox/interior_integration/sample_mesh,
       sandbox/eig/{lee_mittra92_fig6b,lee_mittra92_fig6c,rect-white}, sandbox/patch/patch) -> gmsh_pairs.npz
"""
import os
import pickle

import numpy as np

REF = '/root/reference/sucemfem/'
HERE = os.path.dirname(os.path.abspath(__file__))


def load(p):
    with open(REF + p, 'rb') as f:
        return pickle.load(f, encoding='latin1')


def read_dolfin_xml_mesh(path):
    import xml.etree.ElementTree as ET
    mesh = ET.parse(path).getroot().find('mesh')
    vs, cs = mesh.find('vertices'), mesh.find('cells')
    co = np.zeros((int(vs.get('size')), 3))
    for v in vs:
        co[int(v.get('index'))] = [float(v.get(a)) for a in 'xyz']
    ce = np.zeros((int(cs.get('size')), 4), np.int32)
    for c in cs:
        ce[int(c.get('index'))] = [int(c.get('v%d' % i)) for i in range(4)]
    return co, ce


def main():
    lhs = np.load(REF + 'ProblemConfigurations/tests/LHS_matrix.npy')
    rhs = load('ProblemConfigurations/tests/RHS_vector.pickle')
    np.savez_compressed(os.path.join(HERE, 'driven_order1_inscribedtet.npz'), LHS=lhs, RHS=np.asarray(rhs),
                        frequency=1e8, source_coord=np.zeros(3), source_value=np.array([2, -1, 3.]))
    d = load('Sources/tests/fillament_source_test_data.pickle')
    np.savez_compressed(os.path.join(HERE, 'fillament_source_order2.npz'), **{k: np.asarray(v) for k, v in d.items()})
    d = load('PostProcessing/tests/data/power_flux_reference_dofs_f-1000000000.000000_o-2_s-0.074948_'
             'l-0.100000_h-0.166667.pickle')
    # the mesh the reference's flux tests read back (coordinates rounded by the XML writer: the known answers below
    # belong to THIS mesh, not to the exact get_centred_cube one the solution was computed on)
    mco, mce = read_dolfin_xml_mesh(REF + 'PostProcessing/tests/data/' + d['meshfile'])
    np.savez_compressed(os.path.join(HERE, 'driven_order2_unitcube3.npz'), x=d['x'], freq=d['freq'],
                        order=d['order'], source_endpoints=d['source_endpoints'], I=d['I'],
                        mesh_coordinates=mco, mesh_cells=mce,
                        known_answer_variational_flux=np.complex128(26.9015957282 - 7.85747240242e-08j),
                        known_answer_surface_flux=26.3586862457)
    d = load('PostProcessing/tests/data/reference_dofs-2-0.149896229-0.0499654096667.pickle')
    np.savez_compressed(os.path.join(HERE, 'driven_order2_unitcube5.npz'), **{k: np.asarray(v) for k, v in d.items()})
    # far-field fixtures of test_ntff.py:80-110 (the near-field input is the dof vector above, on the NOT recentred
    # get_centred_cube(domain_size, max_edge_len) mesh) and the testing-expression samples of test_ntff.py:33-61
    out = {}
    for kind in ('surface', 'variational'):
        t = load('PostProcessing/tests/data/reference_%s_ntff-2-0.149896229-0.0499654096667.pickle' % kind)
        ff, nf = t['ff_result_data'], t['nf_input_data']
        assert np.array_equal(nf['x'], d['x'])
        out.update({kind + '_E_ff': np.asarray(ff['E_ff']), kind + '_theta_deg': np.asarray(ff['theta_deg'], float),
                    kind + '_phi_deg': np.asarray(ff['phi_deg'], float)})
        out.update(dict(domain_size=np.asarray(nf['domain_size']), max_edge_len=float(nf['max_edge_len']),
                        freq=float(nf['freq']), order=int(nf['order'])))
    t = load('PostProcessing/tests/data/interpolant_test_data.pickle')
    out.update({'interp_' + k: np.asarray(v) for k, v in t.items()})
    np.savez_compressed(os.path.join(HERE, 'ntff_order2_unitcube5.npz'), **out)


def prism():
    import xml.etree.ElementTree as ET
    base = '/root/reference/functional_tests/data/meshes/rectangular_prism'
    mesh = ET.parse(base + '.xml').getroot().find('mesh')
    vs, cs = mesh.find('vertices'), mesh.find('cells')
    co = np.zeros((int(vs.get('size')), 3))
    for v in vs:
        co[int(v.get('index'))] = [float(v.get(a)) for a in 'xyz']
    ce = np.zeros((int(cs.get('size')), 4), np.int32)
    for c in cs:
        ce[int(c.get('index'))] = [int(c.get('v%d' % i)) for i in range(4)]
    mf = ET.parse(base + '_facet_region.xml').getroot().find('meshfunction')
    mk = np.zeros(int(mf.get('size')), np.uint32)
    for e in mf:
        mk[int(e.get('index'))] = int(e.get('value'))
    np.savez_compressed(os.path.join(HERE, 'rectangular_prism_gmsh.npz'), coordinates=co, cells=ce,
                        facet_markers=mk, facet_dim=int(mf.get('dim')))



def gmsh_pairs():
    """.msh files the reference ships TOGETHER WITH their dolfin-convert output (same mesh revision): the .msh text
    is the reader's input, the XML contents are the expected output."""
    import xml.etree.ElementTree as ET
    R = '/root/reference/'
    out = {}
    for name, base in (('sample_mesh', 'sandbox/interior_integration/sample_mesh'),
                       ('lee_mittra92_fig6c', 'sandbox/eig/lee_mittra92_fig6c'),
                       ('lee_mittra92_fig6b', 'sandbox/eig/lee_mittra92_fig6b'),
                       ('rect_white', 'sandbox/eig/rect-white'),
                       ('patch', 'sandbox/patch/patch')):
        with open(R + base + '.msh', 'rb') as f:
            out[name + '__msh'] = np.frombuffer(f.read(), np.uint8)
        mesh = ET.parse(R + base + '.xml').getroot().find('mesh')
        dim = int(mesh.get('dim'))
        vs, cs = mesh.find('vertices'), mesh.find('cells')
        co = np.zeros((int(vs.get('size')), dim))
        for v in vs:
            co[int(v.get('index'))] = [float(v.get(a)) for a in 'xyz'[:dim]]
        ce = np.zeros((int(cs.get('size')), dim + 1), np.int32)
        for c in cs:
            ce[int(c.get('index'))] = [int(c.get('v%d' % i)) for i in range(dim + 1)]
        out[name + '__coordinates'], out[name + '__cells'] = co, ce
        for kind in ('physical_region', 'facet_region'):
            p = R + base + '_' + kind + '.xml'
            if os.path.exists(p):
                mf = ET.parse(p).getroot().find('meshfunction')
                a = np.zeros(int(mf.get('size')), np.uint32)
                for e in mf:
                    a[int(e.get('index'))] = int(e.get('value'))
                assert int(mf.get('dim')) == (dim if kind == 'physical_region' else dim - 1)
                out[name + '__' + kind] = a
    np.savez_compressed(os.path.join(HERE, 'gmsh_pairs.npz'), **out)



def filament_example():
    """examples/filament_source/data/dofs-2-...-100.pickle: order-2 solution of examples/filament_source/driver.py:48-98
    on get_centred_cube(1 lambda, lambda/6) = UnitCube(9), 30186 dofs.  The stored vector satisfies the linear system
    to 9.97e-6 -- it came out of the driver's iterative branch (scipy bicgstab, default tolerance 1e-5,
    driver.py:93-94), so it pins the matrices and the right-hand side only to that tolerance."""
    with open('/root/reference/examples/filament_source/data/'
              'dofs-2-0.299792458-0.0499654096667-0.0749481145-100.pickle', 'rb') as f:
        d = pickle.load(f, encoding='latin1')
    np.savez_compressed(os.path.join(HERE, 'driven_order2_unitcube9_bicgstab.npz'), **{k: np.asarray(v) for k, v in d.items()})


if __name__ == '__main__':
    prism()
    gmsh_pairs()
    filament_example()
    main()
