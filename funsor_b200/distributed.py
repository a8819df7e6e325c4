"""Batch- and time-sharding of the chain path across the GPUs of one box.

One process per GPU (torch.distributed, backend "nccl" on GPUs; "gloo" in the CPU tests).

* Batch sharding: chains are independent (the batch dims are plain leading dims everywhere,
  funsor/tensor.py:138-150), so rank r takes chains [r*B/G, (r+1)*B/G) and nothing is exchanged on the data
  path; `gather_batch` collects the per-chain results at the end.
* Time sharding (B small, T long; BASELINE config 5): the (x)-product over time is associative -- all that
  sequential_sum_product relies on (funsor/sum_product.py:652-701) -- so rank r reduces its contiguous chunk
  to one S x S summary, ONE all-gather moves the G summaries (mixed_sequential_sum_product with segments =
  GPUs, sum_product.py:766-795), and every rank folds the summaries on its left into its incoming boundary
  message.  The fold is G-1 vector-matrix products of size S: negligible.

The numerical work is done by callables so that the host logic can run under gloo on CPU with the oracle
standing in for the kernels (tests/test_distributed_cpu.py); the defaults are the CUDA entry points.
"""
import torch
import torch.distributed as dist


def shard_bounds(n, rank, world):
    """Contiguous, balanced split of range(n): the first n % world shards get one extra element."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _world(group=None):
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


# ------------------------------------------------------------------------------------------------ batch
def gather_batch(local, total, group=None):
    """All-gather per-chain results of a batch-sharded call back into [total, ...] on every rank."""
    rank, world = _world(group)
    if world == 1:
        return local
    sizes = [shard_bounds(total, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)], 0)


def hmm_forward_batch_sharded(init, A, obs, sum_op="logaddexp", group=None, forward=None):
    """Each rank holds the FULL inputs here only for convenience of the caller; it computes its own slice of
    the batch and the slices are gathered.  (bench.py generates per-rank inputs directly instead.)"""
    if forward is None:
        from .ops import hmm_forward as forward
    rank, world = _world(group)
    B = obs.shape[0]
    lo, hi = shard_bounds(B, rank, world)
    init_l = init[lo:hi] if init is not None and init.dim() == 2 else init
    A_l = A[lo:hi] if A.dim() > 2 else A
    z = forward(init_l, A_l, obs[lo:hi].contiguous(), sum_op=sum_op) if hi > lo else obs.new_empty((0,))
    return gather_batch(z, B, group)


# ------------------------------------------------------------------------------------------------ time
def _default_matmul(sum_op):
    from .ops import semiring_matmul

    return lambda x, y: semiring_matmul(sum_op, "add", x, y)


def exchange_summaries(summary, group=None):
    """The one collective of the time-sharded path: all-gather of the per-chunk summaries.
    summary [B,S,S] -> [G,B,S,S] (chunk order = rank order)."""
    rank, world = _world(group)
    if world == 1:
        return summary.unsqueeze(0)
    flat = torch.empty((world * summary.shape[0],) + tuple(summary.shape[1:]), dtype=summary.dtype, device=summary.device)
    dist.all_gather_into_tensor(flat, summary.contiguous(), group=group)  # concatenates along dim 0 in rank order
    return flat.reshape((world,) + tuple(summary.shape))


def fold_left(init, summaries, upto, matmul):
    """Boundary message entering chunk `upto`: init (x) M_0 (x) ... (x) M_{upto-1}.  init [B,S] -> [B,S]."""
    v = init.unsqueeze(-2)  # [B,1,S]
    for r in range(upto):
        v = matmul(v, summaries[r])
    return v.squeeze(-2)


def fold_right(final, summaries, after, matmul):
    """Backward boundary message leaving chunk `after`: M_{after+1} (x) ... (x) M_{G-1} (x) final.  [B,S]."""
    v = final.unsqueeze(-1)  # [B,S,1]
    for r in range(summaries.shape[0] - 1, after, -1):
        v = matmul(summaries[r], v)
    return v.squeeze(-1)


def hmm_forward_time_sharded(init, A, obs_chunk, sum_op="logaddexp", group=None, summary_fn=None, matmul=None,
                             reduce_fn=None):
    """Log-likelihood (or max-plus value) of chains whose TIME axis is split over the ranks.

    init [B,S] (used by every rank), A [S,S] | [B,S,S], obs_chunk [B,T_r,S] = this rank's contiguous slice of
    time, rank order = time order.  Returns (value[B], boundary[B,S]): value on every rank, and the message
    entering this rank's chunk (the hand-over a local alpha / marginals pass starts from)."""
    if summary_fn is None:
        from .ops import hmm_chunk_summary

        summary_fn = lambda A_, obs_: hmm_chunk_summary(A_, obs_, sum_op=sum_op)  # noqa: E731
    if matmul is None:
        matmul = _default_matmul(sum_op)
    if reduce_fn is None:
        reduce_fn = (lambda v: torch.logsumexp(v, -1)) if sum_op in ("logaddexp", "log") else (lambda v: v.amax(-1))
    rank, world = _world(group)
    mine = summary_fn(A, obs_chunk)  # [B,S,S]
    allm = exchange_summaries(mine, group)  # [G,B,S,S]
    boundary = fold_left(init, allm, rank, matmul)
    last = fold_left(boundary, allm[rank:], allm.shape[0] - rank, matmul)
    return reduce_fn(last), boundary
