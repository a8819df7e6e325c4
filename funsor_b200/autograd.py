"""Backward passes of the path's entry points, so that `loss.backward()` through a funsor expression keeps working when
the eager rules of funsor_plugin.py (or the classes of hmm.py) put the CUDA kernels on the forward path.

The reference gets its gradients from torch autograd through `torch.logsumexp` / `einsum` (examples/discrete_hmm.py:76-78;
the max shift is detached at einsum/numpy_log.py:27).  Kernel outputs have no grad_fn, so each differentiable entry point is
wrapped in a torch.autograd.Function whose backward again runs on the kernels:

  SemiringMatmul   d out / d x, d out / d y of one scan level (cnf.py:335-379 / 312-324)
                   log: two log-semiring matmuls of W = log|g| - out with y', x'   (the softmax weight of every (i,k,j) term)
                   max: scatter of g to the first maximising k (approximations.py:115-127)
  HmmForward       d logZ / d (init, A, obs) of the factored forward pass (pyro/hmm.py:129-135) = the posterior marginals
                   (examples/forward_backward.py:58-107): one forward-backward pass of fb2_hmm_marginals_f32
                   max: one-hot along the Viterbi path (fb2_hmm_viterbi_f32)

A chain product whose input requires grad is evaluated as the reference's own tree (sum_product.py:690-701) of
SemiringMatmul nodes, so autograd walks the same levels back.  There is no CPU path here either.
"""
import torch

from . import ops


def _neg_inf_like(t):
    return torch.full_like(t, float("-inf"))


class SemiringMatmul(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sum_op, x, y):
        semi = ops.semiring_id(sum_op)
        ctx.semi = semi
        if semi == ops.SEMIRING_MAX:
            out, argk = ops._semiring_matmul_raw(sum_op, "add", x, y, want_argmax=True)
            ctx.save_for_backward(argk)
            ctx.shapes = (x.shape, y.shape)
        else:
            out = ops._semiring_matmul_raw(sum_op, "add", x, y)
            ctx.save_for_backward(x, y, out)
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        if ctx.semi == ops.SEMIRING_MAX:
            (argk,) = ctx.saved_tensors
            xs, ys = ctx.shapes
            I, K, J = xs[-2], xs[-1], ys[-1]
            batch = g.shape[:-2]
            idx = argk.long()
            gx = gy = None
            if ctx.needs_input_grad[1]:
                gx = torch.zeros(batch + (I, K), dtype=g.dtype, device=g.device).scatter_add_(-1, idx, g).sum_to_size(xs)
            if ctx.needs_input_grad[2]:
                gy = torch.zeros(batch + (K, J), dtype=g.dtype, device=g.device).scatter_add_(-2, idx, g).sum_to_size(ys)
            return None, gx, gy
        x, y, out = ctx.saved_tensors
        xt = x.transpose(-1, -2).contiguous()
        yt = y.transpose(-1, -2).contiguous()
        live = torch.isfinite(out)
        parts = [(g > 0) & live]
        if bool((g < 0).any()):
            parts.append((g < 0) & live)
        gx = gy = None
        for sign, mask in zip((1.0, -1.0), parts):
            # W = log|g| - out on the entries of this sign; -inf elsewhere (they contribute exp(-inf) = 0)
            W = torch.where(mask, (g * sign).clamp_min(1e-38).log() - torch.where(live, out, torch.zeros_like(out)), _neg_inf_like(out))
            if ctx.needs_input_grad[1]:
                t = (x + ops._semiring_matmul_raw("logaddexp", "add", W, yt)).exp()  # [.., I, K]
                gx = sign * t if gx is None else gx + sign * t
            if ctx.needs_input_grad[2]:
                t = (y + ops._semiring_matmul_raw("logaddexp", "add", xt, W)).exp()  # [.., K, J]
                gy = sign * t if gy is None else gy + sign * t
        if gx is not None:
            gx = torch.nan_to_num(gx, nan=0.0).sum_to_size(x.shape)  # x = -inf meets W-product = +inf only when out = -inf (masked)
        if gy is not None:
            gy = torch.nan_to_num(gy, nan=0.0).sum_to_size(y.shape)
        return None, gx, gy


def chain_tree(sum_op, trans):
    """sequential_sum_product as the reference's balanced even/odd tree (sum_product.py:690-701) of differentiable levels:
    trans[..., T, S, S] -> [..., S, S]."""
    T = trans.shape[-3]
    while T > 1:
        even = T // 2 * 2
        z = SemiringMatmul.apply(sum_op, trans[..., 0:even:2, :, :], trans[..., 1:even:2, :, :])
        trans = z if even == T else torch.cat([z, trans[..., T - 1 :, :, :]], dim=-3)
        T = trans.shape[-3]
    return trans[..., 0, :, :]


def _lse(t, dim):
    """logsumexp that returns -inf (not nan) for all -inf slices"""
    m = t.amax(dim, keepdim=True)
    m = torch.where(torch.isfinite(m), m, torch.zeros_like(m))
    return (t - m).exp().sum(dim).log() + m.squeeze(dim)


class HmmForward(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sum_op, init, A, obs):
        ctx.semi = ops.semiring_id(sum_op)
        ctx.save_for_backward(init, A, obs)
        return ops._hmm_forward_raw(init, A, obs, sum_op=sum_op)

    @staticmethod
    def backward(ctx, g):
        init, A, obs = ctx.saved_tensors
        B, T, S = obs.shape
        need_init, need_A, need_obs = ctx.needs_input_grad[1:4]
        g_init = g_A = g_obs = None
        if ctx.semi == ops.SEMIRING_MAX:
            _, path = ops.hmm_viterbi(init, A, obs)  # [B, T+1]
            b = torch.arange(B, device=obs.device)[:, None].expand(B, T)
            t = torch.arange(T, device=obs.device)[None, :].expand(B, T)
            gb = g[:, None].expand(B, T)
            if need_obs:
                g_obs = torch.zeros_like(obs).index_put_((b, t, path[:, 1:]), gb, accumulate=True)
            if need_A:
                g_A = torch.zeros_like(A)
                if A.dim() == 2:
                    g_A.index_put_((path[:, :-1], path[:, 1:]), gb, accumulate=True)
                elif A.dim() == 3:
                    g_A.index_put_((b, path[:, :-1], path[:, 1:]), gb, accumulate=True)
                else:
                    g_A.index_put_((b, t, path[:, :-1], path[:, 1:]), gb, accumulate=True)
            if need_init and init is not None:
                g_init = torch.zeros(B, S, device=obs.device).index_put_((b[:, 0], path[:, 0]), g, accumulate=True).sum_to_size(init.shape)
            return None, g_init, g_A, g_obs
        if A.dim() == 4:
            # time-varying transitions: the pairwise posteriors of every step are needed (d logZ / d A[b,t]); dense adjoint of the
            # chain (adjoint.py:72-131) over trans = A + obs, which has the size of A itself
            trans = A + obs[:, :, None, :]
            i0 = None if init is None else init.expand(B, S)
            Z, adj = ops.chain_adjoint("logaddexp", "add", trans, init=i0)
            xi = torch.nan_to_num((adj + trans - Z[:, None, None, None]).exp(), nan=0.0)
            if need_A:
                g_A = g[:, None, None, None] * xi
            if need_obs:
                g_obs = g[:, None, None] * xi.sum(-2)
            if need_init and init is not None:
                g_init = (g[:, None] * xi[:, 0].sum(-1)).sum_to_size(init.shape)
            return None, g_init, g_A, g_obs
        logZ, gamma, xi_sum = ops.hmm_marginals(init, A, obs, want_xi=bool(need_A))
        if need_obs:
            g_obs = g[:, None, None] * gamma.exp()
        if need_A:
            g_A = g[:, None, None] * xi_sum
            if A.dim() == 2:
                g_A = g_A.sum(0)
        if need_init and init is not None:
            # posterior of the state before the first transition: p(z0=i | obs) ~ init[i] sum_j A[i,j] obs[0,j] beta_1[j], with
            # obs[0,j] beta_1[j] = gamma_1[j] Z / (sum_i init[i] A[i,j])
            A0 = A if A.dim() == 2 else A  # [S,S] or [B,S,S]
            i0 = init.expand(B, S)
            pred = _lse(i0[:, :, None] + A0, 1)  # [B,S]: log sum_i init[i] A[i,j]
            tmp = torch.where(torch.isfinite(pred), gamma[:, 0] + logZ[:, None] - pred, _neg_inf_like(pred))
            post0 = (i0 + _lse(A0 + tmp[:, None, :], 2) - logZ[:, None]).exp()
            g_init = torch.nan_to_num(g[:, None] * post0, nan=0.0).sum_to_size(init.shape)
        return None, g_init, g_A, g_obs


def needs_grad(*tensors):
    return torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors)
