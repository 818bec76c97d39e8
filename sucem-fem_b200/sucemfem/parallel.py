"""Multi-GPU host logic: row-block ownership and the (optional) gather of the distributed CSR.

The path shards by DOF-row ownership: rank p owns the rows of the entities [bounds[p], bounds[p+1]) of the combined
entity numbering and assembles them from the cells incident to those entities (ghost cells duplicated), so assembly
itself needs NO communication.  The only collective is the gather of row blocks for a consumer that wants the whole
matrix on every rank (a host solver); it is plain torch.distributed (NCCL over NVLink on the GPUs, gloo in the CPU
tests) and is never part of the assembly timing.
"""
import torch
import torch.distributed as dist


def split_by_weight(weights_prefix, nparts):
    """Contiguous ranges of (nearly) equal weight: weights_prefix = exclusive prefix sums [n+1] (host tensor/list).
    Returns nparts+1 boundaries; mirrors sfem_partition_rows (which does this on the device over the entity->cell
    incidence counts)."""
    wp = torch.as_tensor(weights_prefix, dtype=torch.int64)
    n = wp.numel() - 1
    total = int(wp[-1] - wp[0])
    targets = torch.tensor([int(wp[0]) + (total * p) // nparts for p in range(nparts + 1)], dtype=torch.int64)
    b = torch.searchsorted(wp, targets, right=False)
    b[0], b[-1] = 0, n
    return [int(x) for x in b]


def split_by_class(weights_prefix, n_edges, nparts):
    """One range of edge entities and one range of face / interior entities per rank, each class balanced on its own
    (mirrors sfem_partition_rows_by_class): every rank then runs the same mix of edge-row and face-row work.
    Returns (edge_bounds, other_bounds), nparts+1 boundaries each, in the combined entity numbering."""
    wp = torch.as_tensor(weights_prefix, dtype=torch.int64)
    eb = split_by_weight(wp[:n_edges + 1], nparts)
    ob = [n_edges + x for x in split_by_weight(wp[n_edges:], nparts)]
    return eb, ob


def my_block(bounds, rank=None):
    rank = dist.get_rank() if rank is None else rank
    return bounds[rank], bounds[rank + 1]


def _all_gather_var(t, counts, group=None):
    """All-gather of blocks of different lengths straight into ONE output array (rank order).  NCCL: torch issues one
    grouped broadcast per rank into the output views (no padding, no staging copies); gloo (CPU tests) only knows
    equal sizes, so blocks are padded to the longest one and compacted afterwards."""
    world = dist.get_world_size(group)
    out = torch.empty(int(sum(counts)), dtype=t.dtype, device=t.device)
    views = list(out.split([int(c) for c in counts]))
    if dist.get_backend(group) == 'nccl' or len(set(int(c) for c in counts)) == 1:
        dist.all_gather(views, t.contiguous(), group=group)
        return out
    n = int(max(counts))
    pad = torch.zeros(n, dtype=t.dtype, device=t.device)
    pad[:t.numel()] = t
    bufs = [torch.empty(n, dtype=t.dtype, device=t.device) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    for p in range(world):
        views[p].copy_(bufs[p][:int(counts[p])])
    return out


def gather_csr_blocks(rowptr_block, cols_block, values_block, group=None):
    """All-gather row blocks (block-relative rowptr [nrows_p+1], cols [nnz_p], values [nnz_p]) into the full CSR on
    every rank (rank order = row order).  values_block may also be a list / tuple of value arrays sharing the pattern
    (S, T, S_0): the pattern is then gathered once and a list of value arrays is returned."""
    many = isinstance(values_block, (list, tuple))
    vals_in = list(values_block) if many else [values_block]
    dev = vals_in[0].device
    sizes = torch.tensor([rowptr_block.numel() - 1, cols_block.numel()], dtype=torch.int64, device=dev)
    all_sizes = [torch.zeros_like(sizes) for _ in range(dist.get_world_size(group))]
    dist.all_gather(all_sizes, sizes, group=group)
    all_sizes = torch.stack(all_sizes).cpu()
    n_rows, n_nnz = [int(x) for x in all_sizes[:, 0]], [int(x) for x in all_sizes[:, 1]]
    ends = _all_gather_var(rowptr_block[1:] - rowptr_block[0], n_rows, group)    # row end offsets, block-relative
    nnz_before = torch.tensor([0] + n_nnz[:-1], dtype=torch.int64).cumsum(0).to(device=dev, dtype=ends.dtype)
    ends += torch.repeat_interleave(nnz_before, torch.tensor(n_rows, device=dev))
    rowptr = torch.cat([torch.zeros(1, dtype=ends.dtype, device=dev), ends])
    cols = _all_gather_var(cols_block, n_nnz, group)
    vals = [_all_gather_var(v, n_nnz, group) for v in vals_in]
    return rowptr, cols, (vals if many else vals[0])


def gather_csr_two_class(edge_block, other_block, group=None):
    """Full CSR from the two-class partition: edge_block / other_block = (rowptr, cols, values) of this rank's edge
    rows and face rows.  Global row order is all edge rows (rank order) followed by all face rows (rank order)."""
    rp_e, cl_e, vl_e = gather_csr_blocks(*edge_block, group=group)
    rp_o, cl_o, vl_o = gather_csr_blocks(*other_block, group=group)
    rowptr, cols = torch.cat([rp_e, rp_o[1:] + rp_e[-1]]), torch.cat([cl_e, cl_o])
    if isinstance(vl_e, list):
        return rowptr, cols, [torch.cat([a, b]) for a, b in zip(vl_e, vl_o)]
    return rowptr, cols, torch.cat([vl_e, vl_o])
