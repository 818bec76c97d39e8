"""Mesh partition for the multi-GPU path (BASELINE.json north_star: "the mesh is partitioned across the GPUs by DOF-row
ownership, with ghost elements duplicated so assembly needs no communication"; SURVEY.md 8.2 row e).  The reference is
serial (/root/reference/sucemfem/ProblemConfigurations/EMDrivenProblem.py:82-87): there is nothing to mirror, only the
numbering contract to keep -- the result must be, bit for bit, the rows of the single-mesh assembly.

  * cells are cut into P contiguous ranges of the global cell list; rank p loads its range plus the HALO = every other
    cell that shares a vertex with it (so every cell containing one of its entities is there), in global cell order,
    with locally renumbered vertices (a monotone map: cells stay vertex-sorted);
  * the device numbers that sub-mesh by first appearance as usual (K1).  An entity is OWNED by the rank in whose range
    it appears first; because the lower halo precedes the own cells and the upper halo follows them, the owned edges
    (faces) are ONE contiguous local index range, in the global order: global index = local index + shift, the shift
    coming from an all-gather of two integers per rank (the counts);
  * columns reference entities owned by other ranks too.  Every rank publishes (key, global index) of its owned
    entities that lie in a boundary cell (an own cell with a vertex shared with another rank); keys are global vertex
    pairs for edges and (global index of the face's lowest edge, third vertex) for faces.  One variable-size all-gather
    per class; ranks look their foreign entities up by key;
  * sfem_space_set_global_ids attaches the maps: patterns of the rank's space then carry GLOBAL column ids, sorted by
    them, and global row numbers; the value kernels run unchanged on local data.  No collective in the data path.

Everything here is set-up plumbing on torch tensors (CPU or CUDA): the same functions run in the gloo / logical-rank
tests on the CPU with the oracle's numbering standing in for the device's."""
import numpy as np
import torch

TET_EDGES = ((2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1))
TET_FACES = ((1, 2, 3), (0, 2, 3), (0, 1, 3), (0, 1, 2))
_EDGE_OF = {e: i for i, e in enumerate(TET_EDGES)}
# local edge that joins the two LOWEST vertices of local face f
FACE_LOW_EDGE = tuple(_EDGE_OF[(f[0], f[1])] for f in TET_FACES)


def cell_ranges(n_cells, world):
    """Contiguous, balanced ranges of the global cell list: [(lo, hi)] * world."""
    b = [(n_cells * p) // world for p in range(world + 1)]
    return [(b[p], b[p + 1]) for p in range(world)]


class RankCells(object):
    """One rank's share of the mesh: own cells [lo, hi) of the global cell list + halo.  `cells` may be a numpy array or
    a torch tensor on any device (the selection runs where the cell list lives: one pass over it)."""

    def __init__(self, cells, lo, hi):
        cells = cells if isinstance(cells, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(cells))
        cells = cells.long()
        dev = cells.device
        n_vertices = int(cells.max()) + 1
        own_v = torch.zeros(n_vertices, dtype=torch.bool, device=dev)
        own_v[cells[lo:hi].reshape(-1)] = True
        touch = own_v[cells].any(dim=1)
        touch[lo:hi] = True
        self.cell_ids = torch.nonzero(touch).flatten()                 # global cell indices, ascending
        self.own_lo = int(torch.searchsorted(self.cell_ids, torch.tensor([lo], device=dev)))   # local range of the own cells
        self.own_hi = self.own_lo + (hi - lo)
        loc = cells[self.cell_ids]
        self.vertex_ids = torch.unique(loc)                            # global vertex ids, ascending
        self.cells_local = torch.searchsorted(self.vertex_ids, loc.reshape(-1)).reshape(loc.shape).to(torch.int32)
        halo = torch.ones(len(self.cell_ids), dtype=torch.bool, device=dev)
        halo[self.own_lo:self.own_hi] = False
        halo_v = torch.zeros(n_vertices, dtype=torch.bool, device=dev)
        halo_v[loc[halo].reshape(-1)] = True
        shared = own_v & halo_v
        # own cells that other ranks hold as halo: exactly those with a shared vertex
        self.boundary_cells = self.own_lo + torch.nonzero(shared[cells[lo:hi]].any(dim=1)).flatten()
        self.global_range = (lo, hi)


def _first_cell(table, n_ent):
    """First (lowest) local cell of every entity of a [C, k] cell -> entity table."""
    C, k = table.shape
    idx = torch.arange(C, device=table.device).repeat_interleave(k)
    out = torch.full((n_ent,), C, dtype=torch.int64, device=table.device)
    return out.scatter_reduce(0, table.reshape(-1).long(), idx, 'amin')


class RankNumbering(object):
    """The phases of the global numbering of one rank; the gathers in between are the caller's (torch.distributed or a
    loop over logical ranks)."""

    def __init__(self, rc, cell_edges, cell_faces, order):
        self.rc, self.order = rc, order
        dev = cell_edges.device
        self.ce = cell_edges.long()
        self.cf = cell_faces.long() if (order == 2 and cell_faces is not None) else None
        self.E = int(self.ce.max()) + 1
        self.F = int(self.cf.max()) + 1 if self.cf is not None else 0
        fe = _first_cell(self.ce, self.E)
        self.e_lo, self.e_hi = int((fe < rc.own_lo).sum()), int((fe < rc.own_hi).sum())
        # local numbering is by first appearance: the owned entities must be the contiguous range [e_lo, e_hi)
        assert bool((fe[self.e_lo:self.e_hi] >= rc.own_lo).all()) and bool((fe[self.e_lo:self.e_hi] < rc.own_hi).all())
        self.f_lo = self.f_hi = 0
        if self.cf is not None:
            ff = _first_cell(self.cf, self.F)
            self.f_lo, self.f_hi = int((ff < rc.own_lo).sum()), int((ff < rc.own_hi).sum())
        self.cells_g = rc.vertex_ids.to(dev)[rc.cells_local.to(dev).long()]   # [C, 4] global vertex ids
        self.bcells = rc.boundary_cells.to(dev).long()

    def counts(self):
        return [self.e_hi - self.e_lo, self.f_hi - self.f_lo]

    # ---- edges
    def edge_table(self, all_counts, rank):
        """all_counts: [world][2].  Sets the owned global edge ids; returns (keys, gids) this rank publishes."""
        self.rank = rank
        self.E_global = sum(c[0] for c in all_counts)
        self.F_global = sum(c[1] for c in all_counts)
        self.e_off = sum(c[0] for c in all_counts[:rank])
        self.f_off = sum(c[1] for c in all_counts[:rank])
        dev = self.ce.device
        key = torch.zeros(self.E, dtype=torch.int64, device=dev)
        for l, (a, b) in enumerate(TET_EDGES):
            key[self.ce[:, l]] = (self.cells_g[:, a] << 32) | self.cells_g[:, b]
        self.e_key = key
        self.e_gid = torch.full((self.E,), -1, dtype=torch.int64, device=dev)
        self.e_gid[self.e_lo:self.e_hi] = self.e_off + torch.arange(self.e_hi - self.e_lo, device=dev)
        pub = torch.zeros(self.E, dtype=torch.bool, device=dev)
        pub[self.ce[self.bcells].reshape(-1)] = True
        pub[:self.e_lo] = False
        pub[self.e_hi:] = False
        return key[pub], self.e_gid[pub]

    @staticmethod
    def _lookup(keys, gids, q):
        order = torch.argsort(keys)
        ks, gs = keys[order], gids[order]
        pos = torch.searchsorted(ks, q).clamp_max(max(len(ks) - 1, 0))
        assert len(ks) > 0 and bool((ks[pos] == q).all()), 'an entity of the halo is missing from the owners\' tables'
        return gs[pos]

    def face_table(self, all_edge_keys, all_edge_gids):
        """Resolves the foreign edges; returns the (keys, gids) of the faces this rank publishes (order 2)."""
        foreign = self.e_gid < 0
        if bool(foreign.any()):
            self.e_gid[foreign] = self._lookup(all_edge_keys, all_edge_gids, self.e_key[foreign])
        if self.cf is None:
            return None
        dev = self.ce.device
        key = torch.zeros(self.F, dtype=torch.int64, device=dev)
        for l, f in enumerate(TET_FACES):
            low_edge = self.e_gid[self.ce[:, FACE_LOW_EDGE[l]]]
            key[self.cf[:, l]] = (low_edge << 32) | self.cells_g[:, f[2]]
        self.f_key = key
        self.f_gid = torch.full((self.F,), -1, dtype=torch.int64, device=dev)
        self.f_gid[self.f_lo:self.f_hi] = self.f_off + torch.arange(self.f_hi - self.f_lo, device=dev)
        pub = torch.zeros(self.F, dtype=torch.bool, device=dev)
        pub[self.cf[self.bcells].reshape(-1)] = True
        pub[:self.f_lo] = False
        pub[self.f_hi:] = False
        return key[pub], self.f_gid[pub]

    def finish(self, all_face_keys, all_face_gids):
        if self.cf is not None:
            foreign = self.f_gid < 0
            if bool(foreign.any()):
                self.f_gid[foreign] = self._lookup(all_face_keys, all_face_gids, self.f_key[foreign])
        return self


def number_logical_ranks(ranks):
    """Drive the phases for a list of RankNumbering objects living in ONE process (tests, single-GPU emulation)."""
    counts = [r.counts() for r in ranks]
    tabs = [r.edge_table(counts, p) for p, r in enumerate(ranks)]
    ek, eg = torch.cat([t[0] for t in tabs]), torch.cat([t[1] for t in tabs])
    tabs = [r.face_table(ek.to(r.ce.device), eg.to(r.ce.device)) for r in ranks]
    if tabs[0] is not None:
        fk, fg = torch.cat([t[0] for t in tabs]), torch.cat([t[1] for t in tabs])
        for r in ranks:
            r.finish(fk.to(r.ce.device), fg.to(r.ce.device))
    return ranks


def _all_gather_var(t, group=None):
    """Variable-length all-gather of a 1-D int64 tensor (concatenated in rank order)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = t.device
    if dist.get_backend(group) == 'gloo':       # CPU tests / single-GPU multi-process tests: gloo gathers host tensors
        t = t.cpu()
    n = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    cap = max(max(sizes), 1)
    buf = torch.zeros(cap, dtype=torch.int64, device=t.device)
    buf[:t.numel()] = t
    out = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return torch.cat([o[:s] for o, s in zip(out, sizes)]).to(dev)


def number_distributed(rn, group=None):
    """The same phases with torch.distributed collectives (nccl on the GPUs, gloo in the CPU tests): one all-gather of
    two integers per rank, then one variable-size all-gather of (key, id) per entity class."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = rn.ce.device if dist.get_backend(group) != 'gloo' else torch.device('cpu')
    c = torch.tensor(rn.counts(), dtype=torch.int64, device=dev)
    allc = [torch.zeros_like(c) for _ in range(world)]
    dist.all_gather(allc, c, group=group)
    counts = [[int(x[0]), int(x[1])] for x in allc]
    k, g = rn.edge_table(counts, rank)
    tab = rn.face_table(_all_gather_var(k, group), _all_gather_var(g, group))
    if tab is not None:
        rn.finish(_all_gather_var(tab[0], group), _all_gather_var(tab[1], group))
    return rn


class RankSpace(object):
    """One rank's device space with the global numbering attached, and its two row blocks."""

    def __init__(self, coords, cells, order, rank, world, distributed=True, group=None, _logical=None):
        from . import device, _native as N
        import ctypes as C
        lo, hi = cell_ranges(len(cells), world)[rank]
        self.rc = rc = RankCells(cells, lo, hi)
        self.rank, self.world = rank, world
        coords = coords if isinstance(coords, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(coords))
        self.mesh = device.DeviceMesh(coords.to(rc.vertex_ids.device)[rc.vertex_ids].cuda(), rc.cells_local.cuda())
        self.space = device.DeviceSpace(self.mesh, order)
        ce, cf = self.space.cell_entities()
        self.numbering = rn = RankNumbering(rc, ce, cf if order == 2 else None, order)
        self._attach = lambda: self._set_ids(N, C)
        if distributed:
            number_distributed(rn, group)
            self._attach()

    def _set_ids(self, N, C):
        rn, sp = self.numbering, self.space
        eg = rn.e_gid.to(torch.int32).contiguous()
        fg = rn.f_gid.to(torch.int32).contiguous() if rn.cf is not None else None
        assert int(rn.e_gid.min()) >= 0 and (fg is None or int(rn.f_gid.min()) >= 0)
        with torch.cuda.device(self.mesh.device):
            N.check(N.lib.sfem_space_set_global_ids(sp.handle, eg.data_ptr(), fg.data_ptr() if fg is not None else None,
                                                    rn.E_global, rn.F_global, rn.e_lo, rn.e_off, rn.f_lo, rn.f_off,
                                                    None))
            torch.cuda.synchronize()
        self.ndof_global = sp.order * (rn.E_global + (rn.F_global if rn.cf is not None else 0))

    def patterns(self):
        """[edge-row block, face-row block] of this rank (order 1: the edge block only)."""
        from . import device
        rn, E = self.numbering, self.space.n_edges
        out = [device.DevicePattern(self.space, rn.e_lo, rn.e_hi)]
        if rn.cf is not None:
            out.append(device.DevicePattern(self.space, E + rn.f_lo, E + rn.f_hi))
        return out

    def local_cell_values(self, per_cell_global):
        """Slice of a per-cell array of the WHOLE mesh (eps_r, mu_r, regions) for this rank's cells (own + halo)."""
        ids = self.rc.cell_ids
        if isinstance(per_cell_global, torch.Tensor):
            return per_cell_global[ids.to(per_cell_global.device)]
        return per_cell_global[ids.cpu().numpy()]


class RankHostJob(object):
    """The host-buffer call (device.HostAssemblyJob = sfem_host_job_*) for ONE RANK of a partitioned mesh: the rank's cells +
    halo come in as HOST arrays (local vertex numbering, from RankCells), the global numbering is exchanged over the process
    group, and the rows the rank owns go out to its page-locked host arrays with GLOBAL column ids.  What the ranks write,
    concatenated (edge rows in rank order, then face rows), is the CSR of the whole mesh."""

    def __init__(self, rc, coords_local, cells_local, order, eps_local=None, mu_local=None, pec='exterior', group=None):
        import ctypes as C
        from . import device, _native as N
        self.job = device.HostAssemblyJob(coords_local, cells_local, order, eps_local, mu_local, pec=pec)
        sp = C.c_void_p()
        N.check(N.lib.sfem_host_job_space(self.job.handle, C.byref(sp)))
        n_cells = int(cells_local.shape[0])
        ce = torch.empty((n_cells, 6), dtype=torch.int32, device='cuda')
        cf = torch.empty((n_cells, 4), dtype=torch.int32, device='cuda') if order == 2 else None
        N.check(N.lib.sfem_space_cell_entities(sp, ce.data_ptr(), cf.data_ptr() if cf is not None else None, None))
        nd, ne, nf, ndc = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int()
        N.check(N.lib.sfem_space_info(sp, C.byref(nd), C.byref(ne), C.byref(nf), C.byref(ndc)))
        self.n_edges_local, self.order = ne.value, order
        torch.cuda.synchronize()
        self.numbering = rn = number_distributed(RankNumbering(rc, ce, cf, order), group)
        eg = rn.e_gid.to(torch.int32).contiguous()
        fg = rn.f_gid.to(torch.int32).contiguous() if cf is not None else None
        N.check(N.lib.sfem_space_set_global_ids(sp, eg.data_ptr(), fg.data_ptr() if fg is not None else None, rn.E_global,
                                                rn.F_global, rn.e_lo, rn.e_off, rn.f_lo, rn.f_off, None))
        torch.cuda.synchronize()

    def run(self, want=('cols', 'S', 'T'), out_edges=None, out_faces=None):
        rn, E = self.numbering, self.n_edges_local
        oe = self.job.run_rows(rn.e_lo, rn.e_hi, want=want, out=out_edges, order=self.order)
        of = None
        if self.order == 2:
            of = self.job.run_rows(E + rn.f_lo, E + rn.f_hi, want=want, out=out_faces, order=self.order)
        return oe, of

    def close(self):
        self.job.close()
