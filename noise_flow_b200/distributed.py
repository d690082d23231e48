"""Multi-GPU plumbing: the patch batch shards embarrassingly (one process per GPU, contiguous N/G slices);
the only data-path collective is ONE all-reduce of ``[sum nll, sum sd_z, n]`` (3 doubles) for the global
mean NLL / sd_z (reference ``tf.reduce_mean``, noise_flow_model.py:478,484).  ``torch.distributed`` is the
transport: NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of a batch of n patches owned by ``rank`` (remainder spread over low ranks)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_sums(sums: torch.Tensor, group=None) -> torch.Tensor:
    """In-place SUM all-reduce of the fp64 ``[sum nll, sum sd_z, n]`` vector."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return sums


def global_means(sums: torch.Tensor, group=None):
    """(mean nll per patch, mean sd_z) over every rank's shard."""
    s = allreduce_sums(sums.clone(), group)
    n = torch.clamp(s[2], min=1.0)
    return s[0] / n, s[1] / n


def sharded_loss(nf, x, y, nlf0=None, nlf1=None, iso=None, cam=None, group=None):
    """``NoiseFlow.loss`` over this rank's shard followed by the single all-reduce: every rank returns the
    global ``(mean NLL, mean sd_z)``."""
    nf._loss(x, y, nlf0, nlf1, iso, cam)
    mean_nll, sd_z = global_means(nf._tls.last_sums, group)      # this thread's sums (many threads may share the model)
    return mean_nll.to(torch.float32), sd_z.to(torch.float32)
