"""Column sharding over ranks.

Columns are independent (radsurf_interface.F90:105; the reference's own
OpenMP decomposition relies on it), so the solve needs no collective: every
rank owns a contiguous block of columns and the matching contiguous slice of
the packed (ntotlay) arrays.  For ragged inputs the blocks are balanced on the
prefix sum of nlay.  The only communication is the optional final gather of
the output slices (torch.distributed: NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np


def shard_columns(nlay, world_size):
    """Split columns 0..ncol into `world_size` contiguous ranges of nearly equal
    total layer count.  Returns [(col0, col1)] with col1 exclusive."""
    nlay = np.asarray(nlay, dtype=np.int64)
    ncol = int(nlay.size)
    work = np.cumsum(np.maximum(nlay, 1))  # Flat tiles still cost one thread
    total = int(work[-1]) if ncol else 0
    bounds = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        bounds.append(int(np.searchsorted(work, target, side="right")))
    bounds.append(ncol)
    for i in range(1, len(bounds)):
        bounds[i] = max(bounds[i], bounds[i - 1])
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


def gather_columns(local, dist, dst=0):
    """Gather per-rank (ncol_local, ...) arrays (numpy or torch) on rank `dst`,
    concatenated in rank order; other ranks get None."""
    import torch
    t = torch.as_tensor(local)
    world = dist.get_world_size()
    sizes = [torch.zeros(1, dtype=torch.int64, device=t.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device))
    nmax = int(max(int(s.item()) for s in sizes))
    pad = torch.zeros((nmax,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    parts = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    if dist.get_rank() != dst:
        return None
    return torch.cat([p[: int(s.item())] for p, s in zip(parts, sizes)], dim=0)
