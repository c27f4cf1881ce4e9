"""Point-range sharded MSM across ranks (one process per GPU).

The reference already splits an MSM into contiguous chunks and folds the partial sums
(halo2_proofs @6b43b6b src/arithmetic.rs:152-176); here a chunk is a GPU.  Rank g of G
owns points [g*n/G, (g+1)*n/G): its slice of the SRS stays resident, its slice of the
scalars arrives over its own PCIe link, it runs the single-GPU MSM and emits one
projective point.  The only exchange is an all-gather of G x 96 bytes (group addition is
not an NCCL reduction op), after which every rank folds the G partials on its GPU.
"""
from __future__ import annotations

from typing import Tuple


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) owned by ``rank``; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_partials(partial, group=None):
    """all_gather of one (12,) int64 partial per rank -> (world, 12) tensor on the same device."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    out = torch.empty((world, 12), dtype=partial.dtype, device=partial.device)
    dist.all_gather_into_tensor(out.view(-1), partial.contiguous().view(-1), group=group)
    return out


def sharded_multiexp(coeffs_local, bases_local, group=None, stream=None):
    """MSM over the union of all ranks' (coeffs_local, bases_local) slices.

    Inputs are cuda int64 tensors ((m,4) and (m,8)) already resident on this rank's GPU.
    Returns a (12,) int64 cuda tensor holding the folded projective sum (same on every rank).
    """
    import torch
    import torch.distributed as dist

    from .arithmetic import dev_g1_fold, dev_msm

    dev = coeffs_local.device
    partial = torch.empty(12, dtype=torch.int64, device=dev)
    dev_msm(coeffs_local, bases_local, partial, stream=stream)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return partial
    parts = gather_partials(partial, group)
    out = torch.empty(12, dtype=torch.int64, device=dev)
    dev_g1_fold(parts, out, stream=stream)
    return out


def sharded_commit(params, coeffs_local, which: str = "g", group=None, stream=None):
    """ParamsKZG::commit over an SRS sharded by point range: ``params`` holds THIS rank's slice of the bases
    (registered once, with its precomputed window table), ``coeffs_local`` the matching slice of the
    polynomial ((m,4) int64 cuda tensor).  Returns the folded (12,) projective sum (same on every rank)."""
    import torch
    import torch.distributed as dist

    from .arithmetic import dev_g1_fold

    dev = coeffs_local.device
    partial = torch.empty(12, dtype=torch.int64, device=dev)
    params.dev_commit(coeffs_local, partial, which=which, stream=stream)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return partial
    parts = gather_partials(partial, group)
    out = torch.empty(12, dtype=torch.int64, device=dev)
    dev_g1_fold(parts, out, stream=stream)
    return out


def column_owner(q: int, world: int) -> int:
    """Rank that commits / transforms column ``q`` when independent columns are spread over GPUs."""
    return q % world


def commit_columns(commit_many_fn, cols, group=None, device=None):
    """Independent per-column commitments scheduled across ranks (north_star: small MSMs do not shard, whole
    columns are dealt round-robin instead).  ``commit_many_fn(list_of_columns) -> (m, 12) uint64 array`` is this
    rank's batched commit (``ParamsKZG.commit_many`` with the full SRS registered on every rank); every rank
    returns all ``len(cols)`` results in column order.  The only exchange is one all_gather of the padded
    per-rank results."""
    import numpy as np
    import torch
    import torch.distributed as dist

    m = len(cols)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    mine = [q for q in range(m) if column_owner(q, world) == rank]
    out_local = np.asarray(commit_many_fn([cols[q] for q in mine]), dtype=np.uint64).reshape(len(mine), 12)
    if world == 1:
        return out_local
    per = (m + world - 1) // world
    pad = np.zeros((per, 12), dtype=np.uint64)
    pad[: len(mine)] = out_local
    t = torch.from_numpy(pad.view(np.int64))
    if device is not None:
        t = t.to(device)
    gathered = torch.empty((world, per, 12), dtype=torch.int64, device=t.device)
    dist.all_gather_into_tensor(gathered.view(-1), t.contiguous().view(-1), group=group)
    g = gathered.cpu().numpy().view(np.uint64)
    res = np.zeros((m, 12), dtype=np.uint64)
    for q in range(m):
        res[q] = g[column_owner(q, world), q // world]
    return res
