"""GPU: the linear-domain matrix-chain kernel (csrc/chain_gemm_tc.cu: TMA + tcgen05 3xTF32 GEMM tree) behind
ops.hmm_chunk_summary(method="gemm") against a float64 evaluation of the same chain product and against the log-domain tree."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def fb():
    import funsor_b200

    return funsor_b200


def _fp64_summary(A, obs):
    """(x)_t (A + obs[t]) in float64, linear domain with per-step renormalisation: log-matrix [S,S]"""
    P = A.double().exp()
    T, S = obs.shape
    M = torch.eye(S, dtype=torch.float64, device=obs.device)
    acc = torch.zeros((), dtype=torch.float64, device=obs.device)
    for t in range(T):
        M = (M @ P) * obs[t].double().exp()[None, :]
        s = M.max()
        M, acc = M / s, acc + s.log()
    return M.log() + acc


@pytest.mark.parametrize("T,S,block", [(2, 256, 4096), (3, 256, 4096), (7, 256, 4), (16, 512, 4096), (37, 256, 8), (130, 384, 32),
                                       (64, 1024, 4096), (257, 512, 64), (1000, 256, 128)])
def test_chunk_summary_gemm_against_float64(fb, T, S, block):
    from funsor_b200.ops import hmm_chunk_summary_gemm

    g = torch.Generator(device=DEV).manual_seed(T + S)
    A = torch.randn(S, S, device=DEV, generator=g).log_softmax(-1)
    obs = torch.randn(1, T, S, device=DEV, generator=g) * 2.0
    rows, off = hmm_chunk_summary_gemm(A, obs, block_steps=block)
    torch.cuda.synchronize()
    got = rows[0].double() + off[0][:, None]
    want = _fp64_summary(A, obs[0])
    assert bool(torch.isfinite(got).all())
    err = (got - want).abs().max().item()
    assert err <= 2e-5 + 1e-6 * T, (T, S, block, err)  # absolute, on log entries of magnitude ~T
    assert abs(rows.max().item()) < 1e-6


def test_chunk_summary_gemm_matches_log_tree_and_feeds_time_sharding(fb):
    from funsor_b200 import distributed as fd

    T, S = 4096, 512
    g = torch.Generator(device=DEV).manual_seed(5)
    A = torch.randn(S, S, device=DEV, generator=g).log_softmax(-1)
    init = torch.randn(1, S, device=DEV, generator=g).log_softmax(-1)
    obs = torch.randn(1, T, S, device=DEV, generator=g)
    r1, o1 = fb.hmm_chunk_summary(A, obs, return_offsets=True, method="gemm")
    r2, o2 = fb.hmm_chunk_summary(A, obs, return_offsets=True, method="tree")
    d = ((r1.double() + o1[..., None]) - (r2.double() + o2[..., None])).abs().max().item()
    assert d < 5e-3, d  # both fp32 trees over 4096 steps; each is checked against float64 separately
    z = fb.hmm_forward(init, A, obs).double()
    val, _ = fd.hmm_forward_time_sharded(init, A, obs, summary_fn=lambda a, o: (r1, o1))
    assert abs(val.item() - z.item()) <= 1e-6 * abs(z.item()) + 1e-3, (val.item(), z.item())


def test_chunk_summary_gemm_sparse_transitions(fb):
    """left-to-right chain: most transition entries are exactly zero (log = -inf); zeros must stay exact zeros"""
    from funsor_b200.ops import hmm_chunk_summary_gemm

    T, S = 24, 256
    g = torch.Generator(device=DEV).manual_seed(9)
    A = torch.full((S, S), -float("inf"), device=DEV)
    idx = torch.arange(S, device=DEV)
    A[idx, idx] = np.log(0.6)
    A[idx[:-1], idx[1:]] = np.log(0.4)
    A[-1, -1] = 0.0
    obs = torch.randn(1, T, S, device=DEV, generator=g)
    rows, off = hmm_chunk_summary_gemm(A, obs, block_steps=8)
    got = rows[0].double() + off[0][:, None]
    want = _fp64_summary(A, obs[0])
    fin = torch.isfinite(want)
    assert bool((torch.isfinite(got) == fin).all())
    assert (got[fin] - want[fin]).abs().max().item() < 1e-4


@pytest.mark.parametrize("B,T,S", [(1, 4, 256), (2, 5, 256), (3, 16, 256), (2, 37, 384), (1, 8, 1024), (4, 7, 512)])
def test_dense_chain_on_the_gemm_engine(fb, B, T, S):
    """sequential_sum_product (sum_product.py:652-701) of a dense log-semiring chain at large S: linear-domain GEMM tree against a
    float64 evaluation of the same products, and against the level-by-level log-domain kernels (FB2_CHAIN_GEMM=0 path)."""
    import os

    g = torch.Generator(device=DEV).manual_seed(B * 100 + T)
    trans = torch.randn(B, T, S, S, device=DEV, generator=g) * 3.0 - 5.0
    got = fb.sequential_sum_product("logaddexp", "add", trans)
    want = trans[:, 0].double()
    for t in range(1, T):
        want = torch.logsumexp(want.unsqueeze(-1) + trans[:, t].double().unsqueeze(-3), -2)
    err = ((got.double() - want).abs() / (1.0 + want.abs())).max().item()
    assert err <= 1e-5, err
    os.environ["FB2_CHAIN_GEMM"] = "0"
    try:
        old = fb.sequential_sum_product("logaddexp", "add", trans)
    finally:
        del os.environ["FB2_CHAIN_GEMM"]
    assert ((got - old).abs() / (1.0 + old.abs())).max().item() <= 2e-5
