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


def gather_csr_blocks(rowptr_block, cols_block, values_block, group=None):
    """All-gather row blocks (block-relative rowptr [nrows_p+1], cols [nnz_p], values [nnz_p]) into the full CSR on
    every rank.  Blocks are padded to the largest block for the fixed-size all_gather."""
    world = dist.get_world_size(group)
    dev = values_block.device
    sizes = torch.tensor([rowptr_block.numel() - 1, values_block.numel()], dtype=torch.int64, device=dev)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    all_sizes = torch.stack(all_sizes).cpu()
    max_rows, max_nnz = int(all_sizes[:, 0].max()), int(all_sizes[:, 1].max())

    def padded(t, n):
        out = torch.zeros(n, dtype=t.dtype, device=dev)
        out[:t.numel()] = t
        return out

    def gather(t, n):
        bufs = [torch.zeros(n, dtype=t.dtype, device=dev) for _ in range(world)]
        dist.all_gather(bufs, padded(t, n), group=group)
        return bufs

    rps = gather(rowptr_block[1:] - rowptr_block[0], max_rows)   # row end offsets, block-relative
    cls = gather(cols_block, max_nnz)
    vls = gather(values_block, max_nnz)
    rowptr, cols, vals, off = [torch.zeros(1, dtype=rowptr_block.dtype, device=dev)], [], [], 0
    for p in range(world):
        nr, nz = int(all_sizes[p, 0]), int(all_sizes[p, 1])
        rowptr.append(rps[p][:nr] + off)
        cols.append(cls[p][:nz])
        vals.append(vls[p][:nz])
        off += nz
    return torch.cat(rowptr), torch.cat(cols), torch.cat(vals)


def gather_csr_two_class(edge_block, other_block, group=None):
    """Full CSR from the two-class partition: edge_block / other_block = (rowptr, cols, values) of this rank's edge
    rows and face rows.  Global row order is all edge rows (rank order) followed by all face rows (rank order)."""
    rp_e, cl_e, vl_e = gather_csr_blocks(*edge_block, group=group)
    rp_o, cl_o, vl_o = gather_csr_blocks(*other_block, group=group)
    return torch.cat([rp_e, rp_o[1:] + rp_e[-1]]), torch.cat([cl_e, cl_o]), torch.cat([vl_e, vl_o])
