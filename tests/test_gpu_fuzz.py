"""Randomised parity sweep over shapes (fixed seeds): every discrete entry point against the float64 / fp32 oracle at shapes
chosen to cross the dispatch boundaries (small-S warp kernel <= 64 < dataflow kernels <= 512 < 1024 < streaming kernels;
chain groups of 16; odd T; batched and time-varying transitions)."""
import numpy as np
import pytest
import torch

from conftest import rel_close
from oracle import semiring_ref as R

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _workspace_guards():
    """FB2_GUARD=1 puts a sentinel band behind every workspace; any kernel writing past the size the library asked for fails here."""
    yield
    from funsor_b200 import ops

    ops.check_guards()


@pytest.fixture(scope="module")
def fb():
    import funsor_b200 as fb

    assert torch.cuda.is_available()
    return fb


def _case(seed):
    rng = np.random.default_rng(seed)
    S = int(rng.choice([3, 17, 64, 65, 100, 128, 129, 200, 256, 257, 384, 512, 513, 700, 1024, 1025, 1300]))
    B = int(rng.choice([1, 2, 5, 16, 17, 33, 130, 260]))
    T = int(rng.choice([1, 2, 3, 7, 31, 64, 65, 130]))
    if S * S * T * B > 6e8:
        T = max(1, int(6e8 // (S * S * B)))
    amode = rng.choice(["shared", "shared", "shared", "batched", "timevarying"]) if S <= 200 else "shared"
    shape = {"shared": (S, S), "batched": (B, S, S), "timevarying": (B, T, S, S)}[amode]
    A = torch.tensor(rng.standard_normal(shape), dtype=torch.float32)
    if rng.random() < 0.5:
        A[..., rng.integers(0, S), :] = -float("inf")
        A[..., rng.integers(0, S), rng.integers(0, S)] = 0.0  # keep every row alive somewhere
    A = torch.log_softmax(torch.nan_to_num(A, neginf=-1e30), -1)
    A = torch.where(A < -1e20, torch.full_like(A, -float("inf")), A)
    init = torch.log_softmax(torch.tensor(rng.standard_normal((B, S)), dtype=torch.float32), -1)
    obs = torch.tensor(rng.standard_normal((B, T, S)) * rng.choice([0.5, 1.0, 5.0]) + rng.choice([0.0, -30.0, 20.0]), dtype=torch.float32)
    if rng.random() < 0.3:
        obs[rng.integers(0, B), rng.integers(0, T), rng.integers(0, S)] = -float("inf")
    return B, T, S, amode, init, A, obs


@pytest.mark.parametrize("seed", range(64))
def test_fuzz_forward_viterbi_marginals(fb, seed):
    B, T, S, amode, init, A, obs = _case(seed)
    i64, A64, o64 = init.double().numpy(), A.double().numpy(), obs.double().numpy()
    want = R.hmm_forward(R.LOG, i64, A64, o64)
    got = fb.hmm_forward(init.cuda(), A.cuda(), obs.cuda()).cpu().numpy()
    rel_close(got, want, rtol=1e-5, atol=1.0)
    # Viterbi: bit-exact value and path against the fp32 evaluation of the same recurrence
    best, path = fb.hmm_viterbi(init.cuda(), A.cuda(), obs.cuda())
    wbest, wpath = R.viterbi(init.numpy(), A.numpy(), obs.numpy())
    assert np.array_equal(best.cpu().numpy(), wbest.astype(np.float32)), (B, T, S, amode)
    assert np.array_equal(path.cpu().numpy(), wpath), (B, T, S, amode)
    # posteriors: log-likelihood and normalisation of the marginals
    logZ, gamma, _ = fb.hmm_marginals(init.cuda(), A.cuda(), obs.cuda(), want_xi=False)
    rel_close(logZ.cpu().numpy(), want, rtol=1e-5, atol=1.0)
    live = np.isfinite(want)
    if live.any():
        tot = np.exp(gamma.cpu().numpy().astype(np.float64)).sum(-1)[live]
        assert np.abs(tot - 1).max() < 2e-3, (B, T, S, amode, np.abs(tot - 1).max())


@pytest.mark.parametrize("seed", range(24))
def test_fuzz_gaussian_chain(seed):
    """Random (B, T, D): general scan and homogeneous operator path against each other and against the float64 oracle."""
    import funsor_b200.gaussian as fg
    from oracle import gaussian_ref as G

    rng = np.random.default_rng(1000 + seed)
    D = int(rng.choice([1, 2, 3, 4, 5, 8, 9, 16, 17, 32]))
    O = int(rng.integers(max(1, D // 2), D + 1))
    B = int(rng.choice([1, 2, 7, 64, 300]))
    T = int(rng.choice([1, 2, 3, 4, 5, 9, 33, 64, 100, 257]))
    if B * T * D * D > 4e6:
        B = max(1, int(4e6 // (T * D * D)))
    F = 0.9 * np.linalg.qr(rng.standard_normal((B, D, D)))[0]
    H = rng.standard_normal((B, D, O)) / np.sqrt(D)

    def spd_tril(n):
        a = rng.standard_normal((B, n, n))
        return np.linalg.cholesky(a @ a.transpose(0, 2, 1) / n + 0.5 * np.eye(n))

    args = (F, 0.1 * rng.standard_normal((B, D)), spd_tril(D), H, 0.1 * rng.standard_normal((B, O)), spd_tril(O), rng.standard_normal((B, T, O)))
    dev = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).cuda()  # noqa: E731
    dargs = [dev(a) for a in args]
    want = G.chain_info(*G.kalman_elements_info(*args), D)
    homog = fg.kalman_chain(*dargs)
    w, P, s = fg.kalman_elements_sqrt(*dargs)
    general = fg.gaussian_sequential_sum_product(*fg.sqrt_to_info(w, P, s))
    for got in (homog, general):
        for a, b in zip(got, want):
            a = a.cpu().numpy().astype(np.float64)
            scale = max(np.abs(b).max(), 1.0)
            tol = 1e-5 * scale * (T * D if a.ndim == 1 else 1)  # the scalar sums T terms of magnitude ~D
            assert np.abs(a - b).max() <= max(tol, 2e-5 * scale), (B, T, D, O, np.abs(a - b).max(), scale)


@pytest.mark.parametrize("seed", range(32))
def test_fuzz_dense_products(fb, seed):
    """Random shapes for the dense operators: pair products (incl. strided even/odd time views, the tcgen05 size classes and
    their edges) and whole chain products (fused tail, level-by-level, tensor-core levels), both semirings."""
    rng = np.random.default_rng(2000 + seed)
    # pair product on even/odd views of a [B, T, I, K] / [B, T, K, J] pair
    B, T = int(rng.integers(1, 4)), int(rng.integers(2, 7))
    I, K, J = (int(rng.choice([1, 5, 31, 32, 63, 64, 65, 127, 128, 129, 200, 257])) for _ in range(3))
    x = torch.tensor(rng.standard_normal((B, T, I, K)) * 3, dtype=torch.float32)
    y = torch.tensor(rng.standard_normal((B, T, K, J)) * 3, dtype=torch.float32)
    if rng.random() < 0.4:
        x[:, :, rng.integers(0, I), :] = -float("inf")
        y[:, :, :, rng.integers(0, J)] = -float("inf")
    xv, yv = x[:, 0::2][:, : T // 2], y[:, 1::2][:, : T // 2]
    for op, sem in (("logaddexp", R.LOG), ("max", R.MAX)):
        got = fb.semiring_matmul(op, "add", xv.cuda(), yv.cuda()).cpu().numpy()
        want = R.pair_product(sem, xv.double().numpy(), yv.double().numpy())
        if op == "max":
            want32 = R.pair_product(sem, xv.numpy(), yv.numpy())
            assert np.array_equal(got, want32.astype(np.float32)), (B, T, I, K, J)
        else:
            rel_close(got, want, rtol=1e-5, atol=1.0)
    # chain product
    S = int(rng.choice([2, 3, 7, 16, 33, 64, 96, 130, 200]))
    Tc = int(rng.choice([1, 2, 3, 5, 8, 13, 31, 64, 100, 257]))
    Bc = int(rng.choice([1, 2, 5]))
    if Bc * Tc * S ** 3 > 3e9:
        Tc = max(1, int(3e9 // (Bc * S ** 3)))
    trans = torch.tensor(rng.standard_normal((Bc, Tc, S, S)), dtype=torch.float32)
    for op, sem in (("logaddexp", R.LOG), ("max", R.MAX)):
        got = fb.sequential_sum_product(op, "add", trans.cuda()).cpu().numpy()
        if op == "max":
            want32 = R.chain_product(sem, trans.numpy())
            assert np.array_equal(got, want32.astype(np.float32)), (Bc, Tc, S)
        else:
            want = R.chain_product(sem, trans.double().numpy())
            rel_close(got, want, rtol=1e-5, atol=1.0)
