"""ORACLE -- test infrastructure / CPU baseline, not product code.

Torch-CPU port of what the reference executes for the discrete HMM path, used ONLY by bench.py's
`cpu_baseline` leg and `--impl reference` arm (the reference is pure Python and cannot travel to the
GPU box, so its algorithm is restated here with the same library calls on all host threads):

  hmm.py:129-130   result = trans + obs                      -> dense [B,T,S,S] materialisation
  sum_product.py:690-701  even/odd pairwise tree             -> ceil(log2 T) levels
  cnf.py:335-379 -> einsum/numpy_log.py:24-47                -> amax / sub / exp / bmm / log / add per level
  hmm.py:134-135   init + reduce(curr), reduce(state)        -> logsumexp

Validated against the numpy oracle in tests/test_oracle_golden.py::test_cpu_port_matches_oracle.
"""
import torch


def log_matmul(x, y):
    finfo_min = torch.finfo(x.dtype).min
    sx = x.detach().amax(-1, keepdim=True).clamp(min=finfo_min)
    sy = y.detach().amax(-2, keepdim=True).clamp(min=finfo_min)
    return (sx + sy) + torch.matmul((x - sx).exp(), (y - sy).exp()).log()


def max_matmul(x, y, chunk=64):
    # tensor.py:700-726 broadcast add + amax, chunked over rows so the [.., S, S, S] tensor fits in memory
    outs = []
    for i0 in range(0, x.shape[-2], chunk):
        outs.append((x[..., i0 : i0 + chunk, :, None] + y[..., None, :, :]).amax(-2))
    return torch.cat(outs, -2)


def chain_product(trans, semiring="log"):
    mm = log_matmul if semiring == "log" else max_matmul
    T = trans.shape[-3]
    while T > 1:
        even = T // 2 * 2
        z = mm(trans[..., 0:even:2, :, :], trans[..., 1:even:2, :, :])
        if T > even:
            z = torch.cat([z, trans[..., T - 1 :, :, :]], -3)
        trans = z
        T = (T + 1) // 2
    return trans[..., 0, :, :]


def hmm_log_prob(init, A, obs, semiring="log"):
    """The reference's DiscreteHMM.log_prob data flow on CPU tensors. init[B,S], A[S,S], obs[B,T,S]."""
    trans = A + obs[..., None, :]  # [B,T,S,S]
    chain = chain_product(trans, semiring)
    if semiring == "log":
        return torch.logsumexp(init[..., None] + chain, dim=(-1, -2))
    return (init[..., None] + chain).amax(dim=(-1, -2))
