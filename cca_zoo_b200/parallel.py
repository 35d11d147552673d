"""Sample-sharded fitting: one process per GPU, each holding a row shard of every view.

The only exchange step of the path: the additive moment buffer
``[M (Dp x Dp) | s (Dp) | n]`` is summed over ranks with ONE all-reduce; everything after it
(covariance finalisation, eigensolves) is replicated and bit-identical on every rank.
Backend-agnostic (NCCL on GPUs; the host logic is exercised with gloo on CPU in tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def is_distributed(group=None) -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def pack_moments(moments: torch.Tensor, n_local: int) -> torch.Tensor:
    """Append the local row count so that it rides in the same message."""
    tail = torch.tensor([float(n_local)], dtype=moments.dtype, device=moments.device)
    return torch.cat([moments, tail])


def allreduce_moments(moments: torch.Tensor, n_local: int, group=None, dims=None):
    """Sum (moments, n) over the ranks of ``group``.  Returns (moments_total, n_total) -- one host read-back (n)."""
    if not is_distributed(group):
        return moments, int(n_local)
    mom, _, n_dev = allreduce_moments_lazy(moments, n_local, group, dims)
    return mom, int(round(float(n_dev.item())))


def allreduce_moments_lazy(moments: torch.Tensor, n_local: int, group=None, dims=None):
    """Like ``allreduce_moments`` but WITHOUT reading the sample count back: returns
    ``(moments_total, n_host, n_dev)`` where exactly one of ``n_host`` (int, single process) and ``n_dev`` (1-element
    float64 device tensor, sharded fit) is not None.  The device-side fit (``ops.rcca_fit``) takes ``n_dev`` as it is,
    so a sharded fit has no host synchronisation between the moment pass and the final copy of the weights."""
    if not is_distributed(group):
        return moments, int(n_local), None
    if moments.is_cuda and dims is not None:
        # the message carries only the upper triangle of 128 x 128 blocks, the column sums and n (half the bytes of the
        # square buffer, no torch.cat copy); packing / unpacking are two small kernels around the ONE all-reduce
        from . import ops

        packed = ops.moments_pack(moments, dims, n_local)
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
        mom, n_dev = ops.moments_unpack(packed, dims, out=moments)
        return mom, None, n_dev
    packed = pack_moments(moments, n_local)
    dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    return packed[:-1], None, packed[-1:]


def shard_rows(n_rows: int, rank: int, world: int):
    """Contiguous row block [lo, hi) of rank ``rank``."""
    base, rem = divmod(n_rows, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
