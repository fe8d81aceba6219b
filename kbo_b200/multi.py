"""Multi-GPU driver for the query path: the index is replicated on every GPU, the CSR batch is cut
into contiguous query ranges of (nearly) equal base counts, every rank runs the hot path on its own
range, and the per-base results are gathered once at the end.

One process per GPU (torchrun); `torch.distributed` is plumbing only: NCCL on GPUs, gloo in the
CPU tests.  There is no collective on the data path.
"""
import numpy as np


def partition_by_bases(offsets, world):
    """Contiguous query ranges [(q0, q1), ...], one per rank, balanced by cumulative bases.

    offsets: CSR offsets (n_queries + 1).  Every query belongs to exactly one rank; ranks may be
    empty when there are fewer queries than ranks."""
    offsets = np.asarray(offsets, dtype=np.uint64)
    nq = len(offsets) - 1
    total = int(offsets[nq] - offsets[0])
    cuts = [0]
    for r in range(1, world):
        target = int(offsets[0]) + (total * r) // world
        q = int(np.searchsorted(offsets, np.uint64(target), side="left"))
        q = min(max(q, cuts[-1]), nq)
        cuts.append(q)
    cuts.append(nq)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def local_slice(concat, offsets, rank, world):
    """This rank's (concat, offsets rebased to 0, (q0, q1))."""
    q0, q1 = partition_by_bases(offsets, world)[rank]
    offsets = np.asarray(offsets, dtype=np.uint64)
    a, b = int(offsets[q0]), int(offsets[q1])
    return np.ascontiguousarray(concat[a:b]), (offsets[q0:q1 + 1] - offsets[q0]).astype(np.uint64), (q0, q1)


def gather_bytes(local, total_len, rank, world, device=None, dst=0):
    """Gathers per-rank uint8 results (contiguous slices of one array of `total_len` bytes, in rank
    order) on rank `dst`.  Works on gloo (CPU tensors) and NCCL (`device` = this rank's cuda device)."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(local, dtype=np.uint8))
    n_local = torch.tensor([t.numel()], dtype=torch.int64)
    if device is not None:
        t, n_local = t.to(device), n_local.to(device)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local)
    sizes = [int(s.item()) for s in sizes]
    width = max(max(sizes), 1)
    padded = torch.zeros(width, dtype=torch.uint8, device=t.device)
    padded[:t.numel()] = t
    bufs = [torch.zeros(width, dtype=torch.uint8, device=t.device) for _ in range(world)]
    dist.all_gather(bufs, padded)  # one collective, after all compute (NCCL has no gather to host)
    if rank != dst:
        return None
    out = np.empty(total_len, dtype=np.uint8)
    pos = 0
    for r in range(world):
        out[pos:pos + sizes[r]] = bufs[r][:sizes[r]].cpu().numpy()
        pos += sizes[r]
    assert pos == total_len
    return out


def matches_sharded(concat, offsets, compute_fn, rank, world, device=None):
    """kbo::matches over a CSR batch on `world` ranks.  compute_fn(concat, offsets) -> uint8 array runs
    this rank's slice (on a GPU box: lambda c, o: api.matches_csr(c, o, index)[:len(c)]).
    Returns the full alignment on rank 0, None elsewhere."""
    c, o, _ = local_slice(concat, offsets, rank, world)
    local = compute_fn(c, o) if len(o) > 1 else np.zeros(0, dtype=np.uint8)
    total = int(np.asarray(offsets, dtype=np.uint64)[-1] - np.asarray(offsets, dtype=np.uint64)[0])
    return gather_bytes(local, total, rank, world, device=device)
