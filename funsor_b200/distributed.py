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
# The folds below are G-1 vector-matrix products of size S on every rank: they run in float64 with torch (device of the
# inputs).  Summaries and boundary messages carry log-magnitudes of ~1e5 for long chunks; in fp32 that would leave ~1e-2
# of absolute resolution on quantities whose differences matter, hence (i) summaries travel as fp32 rows + float64 row
# offsets, (ii) the folds are float64, (iii) boundary messages are re-centred before they are handed to the fp32 kernels.
def _reduce_op(sum_op):
    if sum_op in ("logaddexp", "log", "logsumexp"):
        return lambda z, dim: torch.logsumexp(z, dim)
    return lambda z, dim: z.amax(dim)


def _as_double_summary(summary):
    """summary: tensor [B,S,S] (absolute) or (rows[B,S,S], row_offset[B,S]) -> float64 [B,S,S]"""
    if isinstance(summary, (tuple, list)):
        rows, off = summary
        return rows.double() + off.double().unsqueeze(-1)
    return summary.double()


def exchange_summaries(summary, group=None):
    """The one collective of the time-sharded path: all-gather of the per-chunk summaries.
    summary [B,S,S] float64 -> [G,B,S,S] (chunk order = rank order)."""
    rank, world = _world(group)
    if world == 1:
        return summary.unsqueeze(0)
    flat = torch.empty((world * summary.shape[0],) + tuple(summary.shape[1:]), dtype=summary.dtype, device=summary.device)
    dist.all_gather_into_tensor(flat, summary.contiguous(), group=group)  # concatenates along dim 0 in rank order
    return flat.reshape((world,) + tuple(summary.shape))


def fold_left(init, summaries, upto, sum_op="logaddexp"):
    """Boundary message entering chunk `upto`: init (x) M_0 (x) ... (x) M_{upto-1}.  init [B,S] -> float64 [B,S]."""
    red = _reduce_op(sum_op)
    v = init.double()
    for r in range(upto):
        v = red(v.unsqueeze(-1) + summaries[r], -2)
    return v


def fold_right(final, summaries, after, sum_op="logaddexp"):
    """Backward boundary message leaving chunk `after`: M_{after+1} (x) ... (x) M_{G-1} (x) final.  float64 [B,S]."""
    red = _reduce_op(sum_op)
    v = final.double()
    for r in range(summaries.shape[0] - 1, after, -1):
        v = red(summaries[r] + v.unsqueeze(-2), -1)
    return v


def _default_summary_fn(sum_op):
    from .ops import hmm_chunk_summary

    return lambda A_, obs_: hmm_chunk_summary(A_, obs_, sum_op=sum_op, return_offsets=True)


def hmm_forward_time_sharded(init, A, obs_chunk, sum_op="logaddexp", group=None, summary_fn=None):
    """Log-likelihood (or max-plus value) of chains whose TIME axis is split over the ranks.

    init [B,S] (used by every rank), A [S,S] | [B,S,S], obs_chunk [B,T_r,S] = this rank's contiguous slice of
    time, rank order = time order.  Returns (value[B] float64, boundary[B,S] float64): value on every rank, and the message
    entering this rank's chunk (the hand-over a local alpha / marginals pass starts from)."""
    summary_fn = summary_fn or _default_summary_fn(sum_op)
    rank, world = _world(group)
    allm = exchange_summaries(_as_double_summary(summary_fn(A, obs_chunk)), group)  # [G,B,S,S]
    boundary = fold_left(init, allm, rank, sum_op)
    last = fold_left(boundary, allm[rank:], allm.shape[0] - rank, sum_op)
    return _reduce_op(sum_op)(last, -1), boundary


def hmm_marginals_time_sharded(init, A, obs_chunk, group=None, summary_fn=None, marginals_fn=None):
    """Forward-backward posteriors with the time axis split over the ranks (BASELINE config 5).

    Each rank reduces its chunk to a summary, ONE all-gather moves the summaries, every rank folds the ones on its left
    into the message entering its chunk and the ones on its right into the backward message leaving it, then runs the
    ordinary local forward-backward on its chunk: the left boundary is the (re-centred) `init`, the right boundary is
    folded into the last observation row (obs[T_r-1] + log beta_out is exactly what the backward recursion starts from).
    Returns (logZ[B] float64 -- the GLOBAL log-likelihood, identical on all ranks -- and gamma[B,T_r,S] for this chunk)."""
    summary_fn = summary_fn or _default_summary_fn("logaddexp")
    if marginals_fn is None:
        from .ops import hmm_marginals

        marginals_fn = lambda i_, A_, o_: hmm_marginals(i_, A_, o_, want_xi=False)[:2]  # noqa: E731
    rank, world = _world(group)
    allm = exchange_summaries(_as_double_summary(summary_fn(A, obs_chunk)), group)
    alpha_in = fold_left(init, allm, rank)
    beta_out = fold_right(torch.zeros_like(init, dtype=torch.float64), allm, rank)
    a_off = alpha_in.amax(-1, keepdim=True)
    b_off = beta_out.amax(-1, keepdim=True)
    a_off = torch.where(torch.isfinite(a_off), a_off, torch.zeros_like(a_off))  # an all -inf boundary (impossible evidence) stays -inf
    b_off = torch.where(torch.isfinite(b_off), b_off, torch.zeros_like(b_off))
    obs2 = obs_chunk.clone()
    obs2[:, -1, :] += (beta_out - b_off).to(obs_chunk.dtype)
    logZ_local, gamma = marginals_fn((alpha_in - a_off).to(obs_chunk.dtype), A, obs2)
    return logZ_local.double() + a_off[:, 0] + b_off[:, 0], gamma
