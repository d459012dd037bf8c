"""Sharding of the (parameter vector × series) batch over the GPUs of one box — the only part of the path that
shards.  One process per GPU (torchrun); every rank evaluates a contiguous slice of the batch on
its own device with no data-path collective, then ONE all-gather (NCCL over NVLink; gloo on CPU in the tests)
returns the log-likelihood vector to every rank, i.e. to the sampler rank.

The reference's own parallelism is process-level too (MPI ranks in examples/ultranest/single_pl.jl:19-21, pmap
workers in examples/turing_distributed/single_pl.jl:70-80): each rank evaluates its points, the sampler gathers.
"""
import numpy as np


def shard_bounds(n, world):
    """Contiguous near-equal split of range(n) into `world` shards: returns world+1 offsets."""
    base, extra = divmod(int(n), int(world))
    sizes = [base + (1 if r < extra else 0) for r in range(world)]
    return np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)


def shard_series(lengths, world):
    """Longest-processing-time assignment of whole series to ranks (ragged multi-source batches, config C3):
    returns a list of index arrays, one per rank, balancing Σ N_s."""
    order = np.argsort(-np.asarray(lengths), kind="stable")
    loads = np.zeros(world)
    out = [[] for _ in range(world)]
    for s in order:
        r = int(np.argmin(loads))
        out[r].append(int(s))
        loads[r] += lengths[s]
    return [np.array(sorted(x), dtype=np.int64) for x in out]


class ShardedEvaluator:
    """Wraps a per-rank evaluator `f(theta_local) -> logl_local` (torch tensors on this rank's device).

    __call__(theta) takes the FULL batch [B × P] (identical on every rank, as ultranest's MPI mode provides it),
    evaluates this rank's slice, and all-gathers so that every rank returns the full logL [B]."""

    def __init__(self, evaluator, group=None):
        import torch.distributed as dist
        self.f = evaluator
        self.group = group
        self.dist = dist
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def local_slice(self, B):
        off = shard_bounds(B, self.world)
        return int(off[self.rank]), int(off[self.rank + 1])

    def __call__(self, theta):
        import torch
        B = theta.shape[0]
        lo, hi = self.local_slice(B)
        local = self.f(theta[lo:hi])
        if self.world == 1:
            return local
        # equal-sized shards for all_gather_into_tensor: pad to the largest shard
        off = shard_bounds(B, self.world)
        m = int(np.max(np.diff(off)))
        buf = torch.full((m,), float("nan"), dtype=local.dtype, device=local.device)
        buf[: hi - lo] = local
        gathered = torch.empty((self.world * m,), dtype=local.dtype, device=local.device)
        self.dist.all_gather_into_tensor(gathered, buf, group=self.group)
        parts = [gathered[r * m: r * m + int(off[r + 1] - off[r])] for r in range(self.world)]
        return torch.cat(parts)
