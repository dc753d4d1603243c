"""Sharding of independent q-batches across the GPUs of one box.

Each of the `b` q-batches of a raw-sample sweep (optim/initializers.py:449-460) or of an L-BFGS-B round is
independent given the replicated model state, so rank `k` of `G` evaluates the contiguous slice
`X[lo_k:hi_k]` with no data-path collective.  The only exchange is an all-gather of the per-shard
acquisition values (8 bytes per q-batch) so that initial-condition selection and the final argmax run on
the full vector with the reference's own host code (bit-exact indices when values agree).
One process per GPU, `torch.distributed` (NCCL on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist
from torch import Tensor


def shard_bounds(b: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous split of range(b) into `world` nearly equal slices (first `b % world` slices get one more)."""
    base, rem = divmod(b, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_values(local: Tensor, b: int, group=None) -> Tensor:
    """All-gather ragged per-shard value vectors into the full length-b vector (on every rank)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_bounds(b, r, world)[1] - shard_bounds(b, r, world)[0] for r in range(world)]
    mx = max(sizes)
    padded = torch.zeros(mx, dtype=local.dtype, device=local.device)
    padded[: local.numel()] = local
    out = torch.empty(world * mx, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return torch.cat([out[r * mx: r * mx + sizes[r]] for r in range(world)])


def sharded_evaluate(acq_function, X: Tensor, group=None) -> Tensor:
    """Evaluate `acq_function` on this rank's slice of the t-batch and return the full value vector."""
    b = X.shape[0]
    if not (dist.is_available() and dist.is_initialized()):
        return acq_function(X)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_bounds(b, rank, world)
    local = acq_function(X[lo:hi]) if hi > lo else X.new_zeros(0)
    return all_gather_values(local.detach(), b, group=group)


def global_argmax(values: Tensor) -> tuple[float, int]:
    """(max, first argmax) of the gathered vector -- `torch.argmax` semantics as in optim/optimize.py:605."""
    idx = int(torch.argmax(values))
    return float(values[idx]), idx
