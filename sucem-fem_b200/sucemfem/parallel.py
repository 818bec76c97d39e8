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


def _all_gather_var(t, counts, out, group=None):
    """All-gather of blocks of different lengths straight into `out` (rank order, sum(counts) elements, usually a
    slice of the final array).  NCCL: torch issues one grouped broadcast per rank into the output views (no padding,
    no staging copies); gloo (CPU tests) only knows equal sizes, so blocks are padded to the longest one and
    compacted afterwards."""
    world = dist.get_world_size(group)
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


def gather_csr_classes(blocks, group=None):
    """Full CSR on every rank from row blocks of several row CLASSES.  blocks = [(rowptr, cols, values), ...], one
    entry per class, each this rank's block of that class: block-relative rowptr [nrows+1], cols [nnz], values [nnz]
    or a list of value arrays sharing the pattern (S, T, S_0).  Global row order = class 0 (rank order), class 1
    (rank order), ...  One small all_gather of the block sizes, then every array is gathered once, directly into its
    slice of the final array."""
    world = dist.get_world_size(group)
    many = isinstance(blocks[0][2], (list, tuple))
    vals_in = [list(b[2]) if many else [b[2]] for b in blocks]
    dev = vals_in[0][0].device
    sizes = torch.tensor([x for b in blocks for x in (b[0].numel() - 1, b[1].numel())], dtype=torch.int64, device=dev)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    all_sizes = torch.stack(all_sizes).cpu().reshape(world, len(blocks), 2)
    n_rows = [[int(x) for x in all_sizes[:, k, 0]] for k in range(len(blocks))]
    n_nnz = [[int(x) for x in all_sizes[:, k, 1]] for k in range(len(blocks))]
    tot_rows, tot_nnz = sum(map(sum, n_rows)), sum(map(sum, n_nnz))
    rdt = blocks[0][0].dtype
    rowptr = torch.empty(tot_rows + 1, dtype=rdt, device=dev)
    rowptr[0] = 0
    cols = torch.empty(tot_nnz, dtype=blocks[0][1].dtype, device=dev)
    vals = [torch.empty(tot_nnz, dtype=v.dtype, device=dev) for v in vals_in[0]]
    r0 = z0 = 0
    for k, (rp, cl, _) in enumerate(blocks):
        nr, nz = sum(n_rows[k]), sum(n_nnz[k])
        ends = _all_gather_var(rp[1:] - rp[0], n_rows[k], rowptr[1 + r0:1 + r0 + nr], group)   # block-relative ends
        before = torch.tensor([z0] + n_nnz[k][:-1], dtype=torch.int64).cumsum(0).to(device=dev, dtype=rdt)
        ends += torch.repeat_interleave(before, torch.tensor(n_rows[k], device=dev))
        _all_gather_var(cl, n_nnz[k], cols[z0:z0 + nz], group)
        for v_out, v_in in zip(vals, vals_in[k]):
            _all_gather_var(v_in, n_nnz[k], v_out[z0:z0 + nz], group)
        r0, z0 = r0 + nr, z0 + nz
    return rowptr, cols, (vals if many else vals[0])


def gather_csr_blocks(rowptr_block, cols_block, values_block, group=None):
    """All-gather row blocks (block-relative rowptr [nrows_p+1], cols [nnz_p], values [nnz_p] or a list of value
    arrays) of ONE contiguous range per rank into the full CSR on every rank (rank order = row order)."""
    return gather_csr_classes([(rowptr_block, cols_block, values_block)], group=group)


def gather_csr_two_class(edge_block, other_block, group=None):
    """Full CSR from the two-class partition: edge_block / other_block = (rowptr, cols, values) of this rank's edge
    rows and face rows.  Global row order is all edge rows (rank order) followed by all face rows (rank order)."""
    return gather_csr_classes([edge_block, other_block], group=group)
