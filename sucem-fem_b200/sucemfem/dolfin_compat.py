"""Stand-ins for the handful of DOLFIN 1.0 objects the reference's drivers and the assembly path touch
(SURVEY.md 8.2 row b): Mesh / UnitCube / UnitSquare, MeshFunction, DomainBoundary, FunctionSpace, the
matrix classes and assemble().  They hold host (numpy) arrays where DOLFIN exposes mutable arrays to the user
(`mesh.coordinates()[:] *= size`, /root/reference/sucemfem/Utilities/MeshGenerators.py:34-36) and delegate all
computation to libsucem_b200.so through sucemfem.device.
"""
import xml.etree.ElementTree as ET

import numpy as np
import torch

from . import _native as N
from . import device


class _Dim(object):
    def __init__(self, d):
        self._d = d

    def dim(self):
        return self._d


class Mesh(object):
    """dolfin.Mesh: coordinates()/cells() are host arrays; device copies are made when a FunctionSpace is built."""

    def __init__(self, filename=None, coordinates=None, cells=None):
        if filename is not None and str(filename).lower().endswith(('.msh', '.gmsh')):
            from .Utilities import GmshIO   # gmsh file: converted on the fly (dolfin-convert's conventions)
            gm = GmshIO.read_gmsh(filename)
            coordinates, cells = gm.coordinates, gm.cells
        elif filename is not None:
            coordinates, cells = _read_dolfin_xml(filename)
        if coordinates is None:
            coordinates, cells = np.zeros((0, 3)), np.zeros((0, 4), np.int32)
        self._coords = np.ascontiguousarray(coordinates, dtype=np.float64)
        self._cells = np.ascontiguousarray(np.sort(np.asarray(cells), axis=1), dtype=np.int32)
        self._topology_space = None

    def coordinates(self):
        return self._coords

    def cells(self):
        return self._cells

    def num_cells(self):
        return self._cells.shape[0]

    def num_vertices(self):
        return self._coords.shape[0]

    def topology(self):
        return _Dim(self._cells.shape[1] - 1)

    def geometry(self):
        return _Dim(self._coords.shape[1])

    def _numbering(self):
        """Order-1 space used only for its entity numbering (facet mesh functions are indexed by it)."""
        if self._topology_space is None:
            self._topology_space = device.DeviceSpace(device.DeviceMesh(self._coords, self._cells), 1)
        return self._topology_space

    def num_facets(self):
        return self._numbering().n_facets


def _read_dolfin_xml(path):
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


def UnitCube(nx, ny, nz):
    m = device.DeviceMesh.unit(int(nx), int(ny), int(nz))
    return Mesh(coordinates=m.coords.cpu().numpy(), cells=m.cells.cpu().numpy())


def UnitSquare(nx, ny):
    m = device.DeviceMesh.unit(int(nx), int(ny), 0)
    return Mesh(coordinates=m.coords.cpu().numpy(), cells=m.cells.cpu().numpy())


class MeshFunction(object):
    """dolfin.MeshFunction('uint', mesh, dim | filename): one uint per mesh entity of the given dimension, indexed
    by the first-appearance entity numbering (facets) or the cell index."""

    def __init__(self, value_type, mesh, dim_or_file):
        assert value_type == 'uint'
        self.mesh = mesh
        if isinstance(dim_or_file, str):
            root = ET.parse(dim_or_file).getroot()
            mf = root.find('meshfunction')
            if mf is None:
                mf = root.find('mesh_function')
            self._dim = int(mf.get('dim'))
            self._values = np.zeros(int(mf.get('size')), np.uint32)
            for e in mf:
                self._values[int(e.get('index'))] = int(e.get('value'))
        else:
            self._dim = int(dim_or_file)
            tdim = mesh.topology().dim()
            if self._dim == tdim:
                n = mesh.num_cells()
            elif self._dim == tdim - 1:
                n = mesh.num_facets()
            elif self._dim == 1:
                n = mesh._numbering().n_edges
            elif self._dim == 0:
                n = mesh.num_vertices()
            else:
                raise NotImplementedError('MeshFunction of dimension %d' % self._dim)
            self._values = np.zeros(n, np.uint32)

    def dim(self):
        return self._dim

    def set_all(self, value):
        self._values[:] = value

    def array(self):
        return self._values


def CellFunction(value_type, mesh):
    return MeshFunction(value_type, mesh, mesh.topology().dim())


def FaceFunction(value_type, mesh):
    return MeshFunction(value_type, mesh, 2)


def EdgeFunction(value_type, mesh):
    return MeshFunction(value_type, mesh, 1)


class DomainBoundary(object):
    """dolfin.DomainBoundary().mark(facet_function, value): all exterior facets."""

    def mark(self, mesh_function, value):
        sp = mesh_function.mesh._numbering()
        d = torch.from_numpy(mesh_function.array().astype(np.int32)).to(sp.mesh.device)
        sp.mark_exterior_facets(d, value)
        mesh_function.array()[:] = d.cpu().numpy().astype(np.uint32)


class _DofMap(object):
    def __init__(self, space):
        self._space = space

    def global_dimension(self):
        return self._space.ndof

    def max_cell_dimension(self):
        return self._space.ndof_cell

    def cell_dofs_array(self):
        return self._space.cell_dofs().cpu().numpy()


class FunctionSpace(object):
    """dolfin.FunctionSpace(mesh, "Nedelec 1st kind H(curl)", order)
    (/root/reference/sucemfem/ProblemConfigurations/EMProblem.py:36,100-105)."""
    FAMILIES = ("Nedelec 1st kind H(curl)", "N1curl")

    def __init__(self, mesh, family, order):
        if family not in self.FAMILIES:
            raise NotImplementedError("only the %r element is provided by the B200 assembly path" % (self.FAMILIES[0],))
        if order not in (1, 2):
            raise NotImplementedError("N1curl order %d: only orders 1 and 2 are provided" % order)
        self._mesh = mesh
        self.order = order
        self.device_mesh = device.DeviceMesh(mesh.coordinates(), mesh.cells())
        self.device_space = device.DeviceSpace(self.device_mesh, order)

    def mesh(self):
        return self._mesh

    def dofmap(self):
        return _DofMap(self.device_space)

    def dim(self):
        return self.device_space.ndof


class DeviceCSRMatrix(object):
    """The tensor class handed to assemble() -- stands in for dolfin.PETScMatrix / dolfin.uBLASSparseMatrix
    (/root/reference/sucemfem/SystemMatrices.py:27, /root/reference/sucemfem/ProblemConfigurations/
    EMDrivenProblem.py:82-87).  Values live in HBM; data() is the hand-off used by
    /root/reference/sucemfem/Utilities/Converters.py:37."""

    def __init__(self):
        self.pattern = None
        self.values = None

    def init(self, pattern, values):
        self.pattern, self.values = pattern, values

    def size(self, dim):
        return self.pattern.space.ndof

    def data(self):
        rp, cl = self.pattern.csr()
        return rp.cpu().numpy(), cl.cpu().numpy(), self.values.cpu().numpy()

    def ident(self, row_flags):
        self.pattern.apply_ident(row_flags, self.values)

    def to_scipy(self):
        import scipy.sparse
        rp, cl, v = self.data()
        n = self.size(0)
        return scipy.sparse.csr_matrix((v, cl, rp), shape=(n, n))

    # linear algebra used by consumers of the path (solvers are out of scope: done on the host by scipy)
    def __mul__(self, other):
        return self.to_scipy() * other

    def __rmul__(self, scalar):
        return scalar * self.to_scipy()

    def __sub__(self, other):
        return self.to_scipy() - (other.to_scipy() if isinstance(other, DeviceCSRMatrix) else other)

    def __add__(self, other):
        return self.to_scipy() + (other.to_scipy() if isinstance(other, DeviceCSRMatrix) else other)


class DeviceLinearOperator(object):
    """sum_m c_m A_m of device matrices that share one pattern, applied without being formed (sfem_matvec) -- e.g. the
    driven system matrix S - k0^2 M + j k0 S_0 (/root/reference/sucemfem/ProblemConfigurations/EMDrivenProblem.py:67-72).
    Quacks enough like a scipy sparse matrix for solver code: shape, dtype, matvec, `A * x`, diagonal()."""

    def __init__(self, matrices, coefficients):
        mats = [m for m in matrices]
        assert len(mats) == len(coefficients) and len(mats) >= 1
        assert all(m.pattern is mats[0].pattern for m in mats), 'the matrices must share one pattern'
        self.pattern = mats[0].pattern
        self.values = [m.values for m in mats]
        self.coefficients = [complex(c) for c in coefficients]
        n = self.pattern.space.ndof
        self.shape = (n, n)
        self.dtype = np.dtype(np.complex128)

    def matvec(self, x):
        host = not isinstance(x, torch.Tensor)
        xd = torch.as_tensor(np.ascontiguousarray(x), device=self.pattern.space.mesh.device) if host else x
        y = self.pattern.matvec(self.values, xd.to(torch.complex128).contiguous(), self.coefficients)
        return y.cpu().numpy() if host else y

    __mul__ = matvec
    dot = matvec

    def diagonal(self):
        n = self.shape[0]
        rp, cl = self.pattern.csr()
        rows = torch.repeat_interleave(torch.arange(n, device=rp.device), rp[1:] - rp[:-1])
        sel = cl.long() == rows
        d = torch.zeros(n, dtype=torch.complex128, device=rp.device)
        for v, c in zip(self.values, self.coefficients):
            d[rows[sel]] += c * v[sel]
        return d.cpu().numpy()

    def to_scipy(self):
        out = None
        for v, c in zip(self.values, self.coefficients):
            m = DeviceCSRMatrix()
            m.init(self.pattern, v)
            out = c * m.to_scipy() if out is None else out + c * m.to_scipy()
        return out


class HostCSRMatrix(DeviceCSRMatrix):
    """A matrix whose CSR arrays live in (page-locked) HOST memory, like the reference's dolfin.uBLASSparseMatrix that
    EMDrivenProblem selects for the scipy hand-off (/root/reference/sucemfem/ProblemConfigurations/
    EMDrivenProblem.py:82-87).  SystemMatrices fills it by streaming row blocks out of the device while they are
    being assembled (DeviceSpace.assemble_to_host); data() then costs nothing.  assemble(form, tensor=HostCSRMatrix())
    of a single form takes the device path and copies."""

    def __init__(self):
        DeviceCSRMatrix.__init__(self)
        self._host = None
        self._ndof = None

    def init_host(self, ndof, rowptr, cols, values):
        self._ndof, self._host = ndof, (rowptr, cols, values)

    def size(self, dim):
        return self._ndof if self._host is not None else DeviceCSRMatrix.size(self, dim)

    def data(self):
        if self._host is None:
            return DeviceCSRMatrix.data(self)
        return tuple(t.numpy() for t in self._host)

    def ident(self, row_flags):
        if self._host is None:
            return DeviceCSRMatrix.ident(self, row_flags)
        rp, cl, v = self.data()
        for r in np.nonzero(row_flags.cpu().numpy())[0]:
            a, b = rp[r], rp[r + 1]
            v[a:b] = (cl[a:b] == r).astype(np.float64)


PETScMatrix = DeviceCSRMatrix
uBLASSparseMatrix = HostCSRMatrix


def assemble(form, tensor=None, pec_row_flags=None, kernel=N.KERNEL_AUTO):
    """dolfin.assemble(form, tensor=mat) for the three bilinear forms of the path
    (/root/reference/sucemfem/SystemMatrices.py:54)."""
    from . import Forms
    if not isinstance(form, Forms.BilinearForm):
        raise NotImplementedError('assemble() handles the N1curl mass, stiffness and ABC bilinear forms')
    mat = DeviceCSRMatrix() if tensor is None else tensor
    space = form.function_space.device_space
    pat = space.pattern()
    coef = None if form.coefficient is None else form.coefficient.device_values()
    code = {'mass': N.FORM_MASS, 'stiffness': N.FORM_STIFFNESS, 'abc': N.FORM_ABC}[form.kind]
    if kernel == N.KERNEL_AUTO and form.kind != 'abc':   # single forms: the generic kernel; S_0: the boundary-row kernel
        kernel = N.KERNEL_GENERIC
    vals = pat.assemble(code, coef=coef, pec_flags=pec_row_flags, kernel=kernel)
    mat.init(pat, vals)
    return mat
