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


def scan_bounds(n, world):
    """Step ranges of the time-axis split (scan_logl_sharded): the near-equal split with the inner bounds at multiples of 8
    steps — the block grid of the tensor-pipe fold and re-filter (csrc/scan_blocked.cuh); an even count is what the warp
    kernel's even/odd step alternation needs when the self-check sweeps on across a hand-over."""
    off = shard_bounds(n, world)
    off[1:-1] &= ~np.int64(7)
    return off


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


def scan_check_total(checks):
    """Deviation estimate (log L units) of a time-axis-sharded scan from the ranks' self-check rows [world × 8]
    (Context.scan_range_check): the inner estimates plus, per hand-over r-1 → r, the difference between the sums rank r-1 got
    by sweeping on into rank r's first steps and the sums rank r got for them from the chained composites.  NaN propagates."""
    checks = np.asarray(checks, dtype=np.float64).reshape(-1, 8)
    est = float(np.sum(checks[:, 0]))
    for r in range(1, checks.shape[0]):
        prev, cur = checks[r - 1], checks[r]
        if prev[6] > 0 and prev[6] == cur[5]:
            est += 0.5 * cur[7] * (abs(prev[3] - cur[1]) + abs(prev[4] - cur[2]))
        else:
            est = float("nan")      # a range too short to be checked
    return est


def scan_logl_sharded(range_begin, range_end, N, rank=0, world=1, all_gather=None, all_reduce_sum=None, range_check=None,
                      sequential=None, tol=1e-10, info=None):
    """Log-likelihood of ONE long series with the time axis split across `world` ranks.

    Rank r owns steps scan_bounds(N, world)[r : r + 2] (near-equal, inner bounds even).  `range_begin(n_lo, n_hi)` folds them and returns the range's composite
    scan element (Context.scan_range_begin); `all_gather(x) -> [world × len(x)]` exchanges the composites — the only
    collective on the data path besides the final 2-value sum; `range_end(prev)` re-filters the range from the state the
    `prev` earlier composites leave behind and returns (Σ log|D_n|, Σ z_n²/D_n); `all_reduce_sum` adds those over the ranks.
    With world == 1 (or no collectives given) it reduces to the single-GPU scan.

    Self-check (the composites lose accuracy on ill-conditioned covariances): with `range_check` (Context.scan_range_check)
    the ranks also gather their check rows; when the estimated deviation exceeds tol·max(1, |log L|) and `sequential` — a
    callable returning the sequential sweep's value of the whole series — is given, every rank returns that instead (all
    ranks see the same gathered rows and take the same decision).  `info`, a dict, receives 'estimate' and 'fallback'."""
    off = scan_bounds(N, world)
    comp = np.asarray(range_begin(int(off[rank]), int(off[rank + 1])), dtype=np.float64)
    if world > 1:
        gathered = np.asarray(all_gather(comp), dtype=np.float64).reshape(world, -1)
        sums = np.asarray(range_end(gathered[:rank]), dtype=np.float64)
        sums = np.asarray(all_reduce_sum(sums), dtype=np.float64)
    else:
        sums = np.asarray(range_end(None), dtype=np.float64)
    value = float(-0.5 * sums[0] - 0.5 * sums[1] - 0.5 * N * np.log(2.0 * np.pi))   # celerite_solver.jl:333
    if range_check is not None:
        row = np.asarray(range_check(), dtype=np.float64)
        rows = np.asarray(all_gather(row), dtype=np.float64).reshape(world, 8) if world > 1 else row.reshape(1, 8)
        rel = scan_check_total(rows) / max(1.0, abs(value))
        fallback = sequential is not None and not (rel <= tol)
        if info is not None:
            info["estimate"], info["fallback"] = rel, fallback
        if fallback:
            value = float(sequential())
    return value


def torch_collectives(device=None, group=None):
    """all_gather / all_reduce_sum for scan_logl_sharded on top of torch.distributed (NCCL on GPUs, gloo on CPU)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)

    def all_gather(x):
        t = torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64, device=device)
        out = torch.empty((world * t.numel(),), dtype=torch.float64, device=device)
        dist.all_gather_into_tensor(out, t, group=group)
        return out.cpu().numpy().reshape(world, -1)

    def all_reduce_sum(x):
        t = torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return t.cpu().numpy()

    return all_gather, all_reduce_sum
