"""Host-side mirror of the reference interface for the Markov-chain sum-product path, on raw
torch CUDA tensors.  Every function here is a thin marshalling layer over the C ABI in
include/funsor_b200.h (ctypes, raw device pointers + strides, launch on the current torch stream).
PyTorch is used for device memory and streams only.  There is no CPU fallback: CPU tensors, non-fp32
dtypes or a missing shared library raise.

Reference functions mirrored (paths relative to the funsor tree):
  sequential_sum_product / naive_ / mixed_   funsor/sum_product.py:652-701, 625-649, 704-795
  Contraction rules (log, max-plus)          funsor/cnf.py:335-379, 312-324
  DiscreteHMM.log_prob                       funsor/pyro/hmm.py:119-138
  AdjointTape.adjoint over a scan            funsor/adjoint.py:72-131
"""
import os

import torch

from . import _lib
from ._lib import SEMIRING_LOG, SEMIRING_MAX, check, strides4


def semiring_id(sum_op, prod_op="add"):
    """Map (sum_op, prod_op) -- strings or funsor.ops objects -- to the C-ABI semiring id.
    Only the two semirings of the path are accepted (funsor/einsum/__init__.py:13-28)."""
    s = getattr(sum_op, "__name__", None) or str(sum_op)
    p = getattr(prod_op, "__name__", None) or str(prod_op)
    s, p = s.replace("ops.", ""), p.replace("ops.", "")
    if p != "add":
        raise NotImplementedError("funsor_b200 supports prod_op=add only, got {!r}".format(prod_op))
    if s in ("logaddexp", "log", "logsumexp"):
        return SEMIRING_LOG
    if s in ("max", "amax"):
        return SEMIRING_MAX
    raise NotImplementedError("funsor_b200 supports sum_op in {{logaddexp, max}}, got {!r}".format(sum_op))


def _dev_f32(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError("{} must be a torch.Tensor".format(name))
    if not t.is_cuda:
        raise ValueError("{} must live on a CUDA device: funsor_b200 has no CPU path".format(name))
    if t.dtype != torch.float32:
        raise ValueError("{} must be float32, got {}".format(name, t.dtype))
    return t


def _stream():
    return torch.cuda.current_stream().cuda_stream


_GUARD = os.environ.get("FB2_GUARD") == "1"  # debugging aid: a sentinel band behind every workspace, verified by check_guards()
_GUARD_BYTES = 1 << 16
_guards = []


def _workspace(nbytes, device):
    n = max(int(nbytes), 256)
    if not _GUARD:
        return torch.empty(n, dtype=torch.uint8, device=device)
    buf = torch.empty(n + _GUARD_BYTES, dtype=torch.uint8, device=device)
    buf[n:] = 0xA5
    _guards.append((buf, n))
    return buf[:n]


def check_guards():
    """With FB2_GUARD=1: synchronise and verify that no kernel wrote past the end of a workspace handed out since the last check."""
    bad = 0
    for buf, n in _guards:
        if not bool((buf[n:] == 0xA5).all()):
            bad += 1
    checked = len(_guards)
    _guards.clear()
    if bad:
        raise RuntimeError("{} of {} workspaces were written past their end".format(bad, checked))
    return checked


def _ptr(t):
    return 0 if t is None else t.data_ptr()


# ----------------------------------------------------------------------------------------- a6 / a7
def semiring_matmul(sum_op, prod_op, x, y, want_argmax=False):
    """out[..., i, j] = (+)_k x[..., i, k] (x) y[..., k, j]  for strided views x, y.

    Mirrors the eager Contraction of two Tensors that one scan level performs (cnf.py:335-379 for
    (logaddexp, add); cnf.py:312-324 for (max, add)).  With want_argmax (max only) also returns the
    first maximising k as int32 (approximations.py:115-127).  Differentiable (autograd.SemiringMatmul) when an
    operand requires grad."""
    from . import autograd as ag

    if not want_argmax and ag.needs_grad(x, y):
        semiring_id(sum_op, prod_op)
        return ag.SemiringMatmul.apply(sum_op, x, y)
    return _semiring_matmul_raw(sum_op, prod_op, x.detach(), y.detach(), want_argmax)


def _semiring_matmul_raw(sum_op, prod_op, x, y, want_argmax=False):
    semi = semiring_id(sum_op, prod_op)
    _dev_f32(x, "x"), _dev_f32(y, "y")
    if x.dim() < 2 or y.dim() < 2 or x.shape[-1] != y.shape[-2]:
        raise ValueError("shape mismatch: x {} y {}".format(tuple(x.shape), tuple(y.shape)))
    I, K, J = x.shape[-2], x.shape[-1], y.shape[-1]
    if K < 1:
        raise ValueError("contracted dimension must be non-empty")
    batch = torch.broadcast_shapes(x.shape[:-2], y.shape[:-2])
    x = x.expand(batch + (I, K))
    y = y.expand(batch + (K, J))
    if len(batch) > 2:  # collapse leading dims (may copy for exotic layouts)
        x = x.reshape((-1, batch[-1], I, K))
        y = y.reshape((-1, batch[-1], K, J))
    while x.dim() < 4:
        x, y = x.unsqueeze(0), y.unsqueeze(0)
    n0, n1 = x.shape[0], x.shape[1]
    out = torch.empty((n0, n1, I, J), dtype=torch.float32, device=x.device)
    arg = torch.empty((n0, n1, I, J), dtype=torch.int32, device=x.device) if want_argmax else None
    lib = _lib.load()
    ws = _workspace(lib.fb2_semiring_matmul_workspace_bytes(n0, n1, I, K, J), x.device)
    with torch.cuda.device(x.device):
        check(lib.fb2_semiring_matmul_f32(semi, x.data_ptr(), strides4(x.stride()), y.data_ptr(), strides4(y.stride()),
                                          out.data_ptr(), _ptr(arg), n0, n1, I, K, J, ws.data_ptr(), ws.numel(), _stream()))
    out = out.reshape(batch + (I, J))
    return (out, arg.reshape(batch + (I, J))) if want_argmax else out


# ----------------------------------------------------------------------------------------- a1 - a4
def sequential_sum_product(sum_op, prod_op, trans, time=-3):
    """Chain product over the `time` dim of trans[..., T, S, S] (prev = dim -2, curr = dim -1), in the
    reference's balanced-tree order (sum_product.py:690-701).  Returns [..., S, S].  When `trans` requires grad the
    same tree is evaluated level by level on differentiable semiring matmuls (autograd.chain_tree)."""
    semi = semiring_id(sum_op, prod_op)
    _dev_f32(trans, "trans")
    if trans.dim() < 3 or trans.shape[-1] != trans.shape[-2]:
        raise ValueError("trans must be [..., T, S, S], got {}".format(tuple(trans.shape)))
    if time not in (-3, trans.dim() - 3):
        trans = trans.movedim(time, -3)
    from . import autograd as ag

    if ag.needs_grad(trans):
        if trans.shape[-3] < 1:
            raise ValueError("time dimension must be non-empty")
        return ag.chain_tree(sum_op, trans)
    trans = trans.detach()
    batch = trans.shape[:-3]
    T, S = trans.shape[-3], trans.shape[-1]
    if T < 1:
        raise ValueError("time dimension must be non-empty")
    t4 = trans.reshape((-1, T, S, S)) if len(batch) != 1 else trans
    B = t4.shape[0]
    out = torch.empty((B, S, S), dtype=torch.float32, device=trans.device)
    if B == 0:  # empty batch (torch hands out a null data pointer for empty tensors)
        return out.reshape(batch + (S, S))
    lib = _lib.load()
    nbytes = lib.fb2_chain_product_workspace_bytes(B, T, S)
    if semi == SEMIRING_LOG and t4.is_contiguous():
        # the linear-domain GEMM tree (csrc/chain_gemm_tc.cu) needs room for the TF32 planes of every level: take it when it is there
        big = lib.fb2_chain_product_gemm_workspace_bytes(B, T, S)
        if big and big < 0.8 * torch.cuda.mem_get_info(trans.device)[0]:
            nbytes = max(nbytes, big)
    ws = _workspace(nbytes, trans.device)
    with torch.cuda.device(trans.device):
        check(lib.fb2_chain_product_f32(semi, t4.data_ptr(), strides4(t4.stride()), B, T, S, out.data_ptr(),
                                        ws.data_ptr(), ws.numel(), _stream()))
    return out.reshape(batch + (S, S))


def naive_sequential_sum_product(sum_op, prod_op, trans, time=-3):
    """Same value as sequential_sum_product (sum_product.py:625-649 is the serial right fold of the
    same associative product); on the GPU both run the tree."""
    return sequential_sum_product(sum_op, prod_op, trans, time)


def mixed_sequential_sum_product(sum_op, prod_op, trans, time=-3, num_segments=None):
    """sum_product.py:704-795: serial scan inside num_segments equal segments, parallel scan across.
    The segment products are themselves chain products, so this is two calls of the same kernel;
    a ragged tail is folded on afterwards exactly like the reference's recursion (731-758)."""
    if time not in (-3, trans.dim() - 3):
        trans = trans.movedim(time, -3)
    T = trans.shape[-3]
    num_segments = T if num_segments is None else num_segments
    if num_segments <= 1 or num_segments >= T:
        return sequential_sum_product(sum_op, prod_op, trans)
    rem = T % num_segments
    head = trans[..., : T - rem, :, :]
    L = (T - rem) // num_segments
    seg = head.reshape(head.shape[:-3] + (num_segments, L) + head.shape[-2:])
    first = sequential_sum_product(sum_op, prod_op, seg)  # [..., num_segments, S, S]
    result = sequential_sum_product(sum_op, prod_op, first)
    if rem:
        tail = torch.cat([result.unsqueeze(-3), trans[..., T - rem :, :, :]], dim=-3)
        result = sequential_sum_product(sum_op, prod_op, tail)
    return result


# ----------------------------------------------------------------------------------------- factored HMM
def _hmm_args(init, A, obs):
    _dev_f32(obs, "obs"), _dev_f32(A, "A")
    if obs.dim() != 3:
        raise ValueError("obs must be [B, T, S], got {}".format(tuple(obs.shape)))
    B, T, S = obs.shape
    obs = obs.contiguous()
    if A.shape[-2:] != (S, S):
        raise ValueError("A must end in [S, S] with S={}, got {}".format(S, tuple(A.shape)))
    A = A.contiguous()
    if A.dim() == 2:
        a_sb, a_st = 0, 0
    elif A.dim() == 3 and A.shape[0] == B:
        a_sb, a_st = S * S, 0
    elif A.dim() == 4 and A.shape[:2] == (B, T):
        a_sb, a_st = T * S * S, S * S
    elif A.dim() == 3 and A.shape[0] == T and B != T:
        raise ValueError("time-varying A must be given as [B, T, S, S]")
    else:
        raise ValueError("A must be [S,S], [B,S,S] or [B,T,S,S]; got {}".format(tuple(A.shape)))
    if init is None:
        init_t, init_sb = None, 0
    else:
        _dev_f32(init, "init")
        init_t = init.contiguous()
        if init_t.shape == (S,):
            init_sb = 0
        elif init_t.shape == (B, S):
            init_sb = S
        else:
            raise ValueError("init must be [S] or [B, S]; got {}".format(tuple(init.shape)))
    return init_t, init_sb, A, a_sb, a_st, obs, B, T, S


def hmm_forward(init, A, obs, sum_op="logaddexp", return_alpha=False):
    """DiscreteHMM.log_prob (hmm.py:129-135) without materialising trans + obs:
    returns (+)_{paths} init[z0] (x) prod_t (A[z_{t-1}, z_t] (x) obs[t, z_t])  as [B].
    Differentiable in (init, A, obs): the backward pass is one forward-backward pass (autograd.HmmForward)."""
    from . import autograd as ag

    if not return_alpha and ag.needs_grad(init, A, obs):
        _hmm_args(init, A, obs)  # argument validation before autograd sees them
        return ag.HmmForward.apply(sum_op, init, A, obs)
    return _hmm_forward_raw(None if init is None else init.detach(), A.detach(), obs.detach(), sum_op, return_alpha)


def _hmm_forward_raw(init, A, obs, sum_op="logaddexp", return_alpha=False):
    semi = semiring_id(sum_op)
    init_t, init_sb, A, a_sb, a_st, obs, B, T, S = _hmm_args(init, A, obs)
    out = torch.empty((B,), dtype=torch.float32, device=obs.device)
    alpha = torch.empty((B, T, S), dtype=torch.float32, device=obs.device) if return_alpha else None
    if B == 0:
        return (out, alpha) if return_alpha else out
    lib = _lib.load()
    ws = _workspace(lib.fb2_hmm_workspace_bytes(3 if return_alpha else 0, B, T, S), obs.device)
    with torch.cuda.device(obs.device):
        check(lib.fb2_hmm_forward_f32(semi, _ptr(init_t), init_sb, A.data_ptr(), a_sb, a_st, obs.data_ptr(), B, T, S,
                                      out.data_ptr(), _ptr(alpha), ws.data_ptr(), ws.numel(), _stream()))
    return (out, alpha) if return_alpha else out


def hmm_viterbi(init, A, obs):
    """Max-plus forward + back-pointers + backtrace.  Returns (best[B], path[B, T+1] int64) with
    path[:, 0] the state before the first transition; first-index tie-break."""
    init_t, init_sb, A, a_sb, a_st, obs, B, T, S = _hmm_args(init, A, obs)
    best = torch.empty((B,), dtype=torch.float32, device=obs.device)
    path = torch.empty((B, T + 1), dtype=torch.int64, device=obs.device)
    if B == 0:
        return best, path
    lib = _lib.load()
    ws = _workspace(lib.fb2_hmm_workspace_bytes(1, B, T, S), obs.device)
    with torch.cuda.device(obs.device):
        check(lib.fb2_hmm_viterbi_f32(_ptr(init_t), init_sb, A.data_ptr(), a_sb, a_st, obs.data_ptr(), B, T, S,
                                      best.data_ptr(), path.data_ptr(), ws.data_ptr(), ws.numel(), _stream()))
    return best, path


def hmm_marginals(init, A, obs, want_xi=True):
    """Forward-backward posteriors (examples/forward_backward.py:58-107,162):
    logZ[B], gamma[B,T,S] = log p(z_t | obs), xi_sum[B,S,S] = sum_t p(z_{t-1}, z_t | obs)."""
    init_t, init_sb, A, a_sb, a_st, obs, B, T, S = _hmm_args(init, A, obs)
    logZ = torch.empty((B,), dtype=torch.float32, device=obs.device)
    gamma = torch.empty((B, T, S), dtype=torch.float32, device=obs.device)
    xi = torch.empty((B, S, S), dtype=torch.float32, device=obs.device) if want_xi else None
    if B == 0:
        return logZ, gamma, xi
    lib = _lib.load()
    ws = _workspace(lib.fb2_hmm_workspace_bytes(2, B, T, S), obs.device)
    with torch.cuda.device(obs.device):
        check(lib.fb2_hmm_marginals_f32(_ptr(init_t), init_sb, A.data_ptr(), a_sb, a_st, obs.data_ptr(), B, T, S,
                                        logZ.data_ptr(), gamma.data_ptr(), _ptr(xi), ws.data_ptr(), ws.numel(), _stream()))
    return logZ, gamma, xi


def hmm_posterior_sample(init, A, obs, num_samples=1, uniforms=None, generator=None):
    """Forward-filter backward-sample: exact draws of state paths from p(z | obs) (recipes.py:16-90 specialised to a chain;
    the discrete draw is Tensor._sample tensor.py:332-427).  `uniforms[B,N,T+1]` in [0,1) may be given to make the draws a
    deterministic function of the inputs.  Returns (paths[B,N,T+1] int64, log_q[B,N] float64 = log p(path | obs), logZ[B])."""
    init_t, init_sb, A, a_sb, a_st, obs, B, T, S = _hmm_args(init, A, obs)
    if uniforms is None:
        uniforms = torch.rand((B, num_samples, T + 1), dtype=torch.float32, device=obs.device, generator=generator)
    _dev_f32(uniforms, "uniforms")
    if uniforms.dim() != 3 or uniforms.shape[0] != B or uniforms.shape[2] != T + 1:
        raise ValueError("uniforms must be [B, N, T+1], got {}".format(tuple(uniforms.shape)))
    uniforms = uniforms.contiguous()
    N = uniforms.shape[1]
    paths = torch.empty((B, N, T + 1), dtype=torch.int64, device=obs.device)
    score = torch.empty((B, N), dtype=torch.float64, device=obs.device)
    if B == 0 or N == 0:
        return paths, score, torch.empty((B,), dtype=torch.float32, device=obs.device)
    logZ, alpha = hmm_forward(init, A, obs, return_alpha=True)
    lib = _lib.load()
    ws = _workspace(lib.fb2_hmm_backward_sample_workspace_bytes(B, T, S, a_sb, a_st), obs.device)
    with torch.cuda.device(obs.device):
        check(lib.fb2_hmm_backward_sample_f32(_ptr(init_t), init_sb, A.data_ptr(), a_sb, a_st, obs.data_ptr(), alpha.data_ptr(),
                                              uniforms.data_ptr(), B, T, S, N, paths.data_ptr(), score.data_ptr(), ws.data_ptr(),
                                              ws.numel(), _stream()))
    return paths, score - logZ.double()[:, None], logZ


def _recentre(M):
    """M[..., S, S] -> (M - max entry, max entry as float64): keeps the fp32 entries O(1)-O(100) whatever the chain length."""
    m = M.amax((-1, -2))
    m = torch.where(torch.isfinite(m), m, torch.zeros_like(m))  # an all -inf product (impossible observation) stays -inf, not NaN
    return M - m[..., None, None], m.double()


def hmm_chunk_summary_tree(A, obs, sum_op="logaddexp", block_bytes=2 << 30, leaf_steps=128):
    """The same summary as a blocked balanced tree of dense semiring matmuls -- literally sequential_sum_product
    (sum_product.py:652-701) over trans[t] = A + obs[t] (hmm.py:130).  Every level is a batch of independent [S x S]
    products, which for the log semiring run on the tcgen05 tensor cores (log_matmul_tc): ~2 S^3 useful flops per time step at
    ~100 TFLOP/s instead of a latency-bound recurrence with S chains.

    fp32 log-domain entries lose absolute resolution as their magnitude grows with the number of multiplied steps, so the
    tree is re-centred: observations are shifted by their per-step maximum, leaves of `leaf_steps` steps are reduced by the
    chain-product kernel, and from there on every product is followed by subtracting its largest entry; all removed shifts
    accumulate in float64.  Time is processed in super-blocks whose dense factor tensor fits in `block_bytes`.
    Returns (rows[B,S,S] fp32 with max entry 0, row_offset[B,S] float64)."""
    _dev_f32(obs, "obs"), _dev_f32(A, "A")
    B, T, S = obs.shape
    if A.dim() != 2:
        raise ValueError("hmm_chunk_summary_tree needs a shared transition matrix A[S,S]")
    shift = obs.amax(-1)  # [B,T]
    shift = torch.where(torch.isfinite(shift), shift, torch.zeros_like(shift))
    off = shift.double().sum(-1)
    L = max(1, min(leaf_steps, T))
    steps_per_block = max(L, int(block_bytes // (B * S * S * 4)) // L * L)
    acc = None
    t0 = 0
    while t0 < T:
        tb = min(steps_per_block, T - t0)
        if tb >= L:
            tb = tb // L * L
            leaf = L
        else:
            leaf = tb  # ragged tail shorter than a leaf
        nb = tb // leaf
        ob = (obs[:, t0 : t0 + tb] - shift[:, t0 : t0 + tb, None]).reshape(B * nb, leaf, 1, S)
        trans = A[None, None] + ob  # [B*nb, leaf, S, S]: the reference's own materialisation, one block at a time
        M = sequential_sum_product(sum_op, "add", trans).reshape(B, nb, S, S)
        del trans, ob
        M, m = _recentre(M)
        off += m.sum(-1)
        while nb > 1:  # balanced tree over the leaves of this block, re-centred after every level
            half = nb // 2
            P2 = semiring_matmul(sum_op, "add", M[:, 0 : 2 * half : 2], M[:, 1 : 2 * half : 2])
            P2, m = _recentre(P2)
            off += m.sum(-1)
            M = P2 if nb == 2 * half else torch.cat([P2, M[:, nb - 1 :]], 1)
            nb = half + (nb - 2 * half)
        M = M[:, 0]
        if acc is None:
            acc = M
        else:
            acc, m = _recentre(semiring_matmul(sum_op, "add", acc, M))
            off += m
        t0 += tb
    return acc.contiguous(), off[:, None].expand(B, S).contiguous()


def hmm_chunk_summary_gemm(A, obs, block_steps=16384):
    """The chunk summary as a linear-domain matrix-chain product on the tensor cores (csrc/chain_gemm_tc.cu: TMA-fed 3xTF32 tcgen05
    GEMMs, hi / lo planes written by the epilogue of the level below, the reference's tree order sum_product.py:690-701 inside blocks
    of `block_steps` steps and across blocks).  Shared log-semiring A[S,S], S a multiple of 128 in [256, 1024].
    Returns (rows[B,S,S] fp32 with max entry 0, row_offset[B,S] float64)."""
    _dev_f32(obs, "obs"), _dev_f32(A, "A")
    B, T, S = obs.shape
    if A.dim() != 2 or A.shape != (S, S):
        raise ValueError("hmm_chunk_summary_gemm needs a shared transition matrix A[S,S]")
    A, obs = A.contiguous(), obs.contiguous()
    rows = torch.empty((B, S, S), dtype=torch.float32, device=obs.device)
    off = torch.empty((B, S), dtype=torch.float64, device=obs.device)
    lib = _lib.load()
    ws = _workspace(lib.fb2_hmm_chunk_summary_gemm_workspace_bytes(T, S, block_steps), obs.device)
    with torch.cuda.device(obs.device):
        for b in range(B):
            check(lib.fb2_hmm_chunk_summary_gemm_f32(A.data_ptr(), obs[b].data_ptr(), T, S, block_steps, rows[b].data_ptr(), off[b].data_ptr(),
                                                     ws.data_ptr(), ws.numel(), _stream()))
    return rows, off


def _gemm_summary_ok(A, obs, sum_op):
    S = obs.shape[-1]
    return (semiring_id(sum_op) == SEMIRING_LOG and A.dim() == 2 and S % 128 == 0 and 256 <= S <= 1024 and obs.shape[1] >= 2
            and os.environ.get("FB2_SUMMARY_GEMM", "1") != "0")


def hmm_chunk_summary(A, obs, sum_op="logaddexp", return_offsets=False, method="auto"):
    """Transfer matrix of a chunk of time: out[b] = (x)_{t in chunk} (A + obs[b,t]) as [B,S,S] -- the summary
    that time-sharded ranks all-gather (mixed_sequential_sum_product sum_product.py:766-795).
    With return_offsets the rows come back relative to a float64 offset [B,S] (true entry = out + offset[..., None]):
    the log-magnitude of a long chunk (~1e5) would otherwise eat the fp32 resolution of the entries.
    method: "gemm" = linear-domain TMA + tcgen05 matrix-chain tree (default with offsets, log semiring, shared A, S a multiple of 128
    in [256, 1024]), "tree" = blocked dense log-domain tree on the tensor cores (shared A, S >= 64), "scan" = S chains per
    batch entry on the recurrence kernels (dataflow kernel with offsets, generic cooperative kernel without)."""
    if method == "auto":
        if return_offsets and _gemm_summary_ok(A, obs, sum_op):
            method = "gemm"
        else:
            method = "tree" if (return_offsets and A.dim() == 2 and obs.shape[-1] >= 64) else "scan"
    if method == "gemm":
        rows, off = hmm_chunk_summary_gemm(A, obs)
        return (rows, off) if return_offsets else (rows.double() + off[..., None]).float()
    if method == "tree":
        rows, off = hmm_chunk_summary_tree(A, obs, sum_op)
        return (rows, off) if return_offsets else (rows.double() + off[..., None]).float()
    semi = semiring_id(sum_op)
    _, _, A, a_sb, a_st, obs, B, T, S = _hmm_args(None, A, obs)
    out = torch.empty((B, S, S), dtype=torch.float32, device=obs.device)
    off = torch.zeros((B, S), dtype=torch.float64, device=obs.device) if return_offsets else None
    lib = _lib.load()
    ws = _workspace(lib.fb2_hmm_chunk_summary_workspace_bytes(B, T, S), obs.device)
    with torch.cuda.device(obs.device):
        check(lib.fb2_hmm_chunk_summary_f32(semi, A.data_ptr(), a_sb, a_st, obs.data_ptr(), B, T, S, out.data_ptr(), _ptr(off),
                                            ws.data_ptr(), ws.numel(), _stream()))
    return (out, off) if return_offsets else out


def chain_adjoint(sum_op, prod_op, trans, init=None, final=None):
    """Backward pass of the scan (adjoint.py:72-131 specialised to a chain):
    returns (Z[B], adj[B,T,S,S]) with adj[b,t,i,j] = alpha_{t-1}[i] (x) beta_t[j].
    Posterior pairwise marginal = adj + trans - Z (examples/forward_backward.py:162)."""
    semi = semiring_id(sum_op, prod_op)
    _dev_f32(trans, "trans")
    if trans.dim() != 4 or trans.shape[-1] != trans.shape[-2]:
        raise ValueError("trans must be [B, T, S, S]")
    trans = trans.contiguous()
    B, T, S, _ = trans.shape
    init_t = None if init is None else _dev_f32(init, "init").expand(B, S).contiguous()
    final_t = None if final is None else _dev_f32(final, "final").expand(B, S).contiguous()
    Z = torch.empty((B,), dtype=torch.float32, device=trans.device)
    adj = torch.empty((B, T, S, S), dtype=torch.float32, device=trans.device)
    lib = _lib.load()
    ws = _workspace(lib.fb2_chain_adjoint_workspace_bytes(B, T, S), trans.device)
    with torch.cuda.device(trans.device):
        check(lib.fb2_chain_adjoint_f32(semi, trans.data_ptr(), strides4(trans.stride()), _ptr(init_t), _ptr(final_t), B, T, S,
                                        Z.data_ptr(), adj.data_ptr(), ws.data_ptr(), ws.numel(), _stream()))
    return Z, adj


def async_status(clear=True, raise_on_abort=True):
    """After synchronising: did a dataflow pass on the current device give up while running (its outputs are then NaN)?
    Returns the condition code (0 = none); with raise_on_abort raises _lib.Fb2Aborted instead of returning non-zero."""
    lib = _lib.load()
    kind = int(lib.fb2_async_status(1 if clear else 0))
    if kind and raise_on_abort:
        raise _lib.Fb2Aborted(_lib.ERR_ABORTED, "a dataflow pass aborted while running (kind {}): its outputs are NaN".format(kind))
    return kind


def launch_count(reset=False):
    lib = _lib.load()
    n = int(lib.fb2_launch_count())
    if reset:
        lib.fb2_launch_count_reset()
    return n
