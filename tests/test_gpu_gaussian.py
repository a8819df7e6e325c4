"""GPU parity tests for the Gaussian-factor chain path against golden fixtures from the unmodified
reference and the float64 oracle.

Tolerance (north star): 1e-5 relative on Gaussian natural parameters, measured per tensor against its
largest entry (the reference's own comparator with atol = max|expected|): entries that are exactly or
nearly zero by structure (e.g. the prev-curr coupling after many contracting steps) carry no relative
information."""
import numpy as np
import pytest
import torch

from conftest import golden_cases, load_golden
from oracle import gaussian_ref as G

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fg():
    import funsor_b200.gaussian as fg

    assert torch.cuda.is_available()
    return fg


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).cuda()


def host(t):
    return t.detach().cpu().numpy().astype(np.float64)


def close_norm(actual, expected, rtol):
    scale = max(np.abs(expected).max(), 1.0)
    err = np.abs(actual - expected).max()
    assert err <= rtol * scale, "max err {:.3e} vs {:.3e} (scale {:.3e})".format(err, rtol * scale, scale)


def test_sqrt_to_info(fg):
    rng = np.random.default_rng(0)
    for batch, n, r in [((3, 5), 4, 4), ((7,), 6, 9), ((2,), 64, 48), ((1,), 64, 192), ((4,), 2, 1)]:
        w = rng.standard_normal(batch + (r,))
        P = rng.standard_normal(batch + (n, r))
        s = rng.standard_normal(batch)
        got = fg.sqrt_to_info(dev(w), dev(P), dev(s))
        want = G.sqrt_to_info(w, P, s)
        for a, b in zip(got, want):
            close_norm(host(a), b, 2e-6 * max(r, 8))


def test_pair_combine_vs_oracle(fg):
    rng = np.random.default_rng(1)
    for D in (1, 2, 3, 8, 17, 32):
        N, n = 6, 2 * D

        def rand_elem():
            P = rng.standard_normal((N, n, n + 3)) / np.sqrt(n)
            Lam = P @ np.swapaxes(P, -1, -2) + 0.5 * np.eye(n)
            return Lam, rng.standard_normal((N, n)), rng.standard_normal(N)

        x, y = rand_elem(), rand_elem()
        Lo, eo, co, flag = fg.gaussian_pair_combine(tuple(dev(t) for t in x), tuple(dev(t) for t in y))
        want = G.info_pair_combine(x, y, D)
        assert int(flag.item()) == 0
        for a, b in zip((Lo, eo, co), want):
            close_norm(host(a), b, 1e-5)


def test_chain_golden(fg):
    for case, d in golden_cases(load_golden("gaussian_chain")).items():
        D = int(case.split("_D")[1])
        Lam, eta, c = fg.sqrt_to_info(dev(d["w"]), dev(d["P"]))
        got = fg.gaussian_sequential_sum_product(Lam, eta, c)
        for a, key in zip(got, ("Lam", "eta", "c")):
            close_norm(host(a), d[key], 1e-5 * max(1, d["w"].shape[1] // 4))


def test_kalman_golden(fg):
    for case, d in golden_cases(load_golden("kalman")).items():
        args = [dev(d[k]) for k in ("F", "q_loc", "q_tril", "H", "r_loc", "r_tril", "y")]
        chain = fg.kalman_chain(*args)
        for a, key in zip(chain, ("chain_Lam", "chain_eta", "chain_c")):
            close_norm(host(a), d[key], 1e-5)
        lp = fg.gaussian_hmm_log_prob(dev(d["m0"]), dev(d["p0_tril"]), chain)
        close_norm(host(lp), d["logprob"], 1e-5)


@pytest.mark.parametrize("B,T,D,O", [(4, 257, 8, 4), (2, 1000, 32, 16), (64, 100, 32, 16), (3, 33, 5, 3)])
def test_kalman_random_vs_oracle(fg, B, T, D, O):
    rng = np.random.default_rng(2)

    def spd_tril(n):
        a = rng.standard_normal((B, n, n))
        return np.linalg.cholesky(a @ np.swapaxes(a, -1, -2) / n + 0.5 * np.eye(n))

    F = 0.9 * np.linalg.qr(rng.standard_normal((B, D, D)))[0]
    H = rng.standard_normal((B, D, O)) / np.sqrt(D)
    q_loc, r_loc = 0.1 * rng.standard_normal((B, D)), 0.1 * rng.standard_normal((B, O))
    q_tril, r_tril = spd_tril(D), spd_tril(O)
    m0, p0 = rng.standard_normal((B, D)), spd_tril(D)
    y = rng.standard_normal((B, T, O))
    args = (F, q_loc, q_tril, H, r_loc, r_tril, y)
    chain = fg.kalman_chain(*[dev(a) for a in args])
    want = G.chain_info(*G.kalman_elements_info(*args), D)
    for a, b in zip(chain, want):
        close_norm(host(a), b, 1e-5)
    lp = fg.gaussian_hmm_log_prob(dev(m0), dev(p0), chain)
    wlp = G.hmm_logprob_from_chain(G.mvn_info(m0, p0), want, D)
    close_norm(host(lp), wlp, 1e-5)
    if T <= 300:
        close_norm(host(lp), G.kalman_filter_logprob(m0, p0, *args), 1e-5)


def test_too_little_information_raises(fg):
    D = 2
    Lam = torch.zeros(1, 2, 4, 4, device="cuda")  # no information on the shared block
    eta = torch.zeros(1, 2, 4, device="cuda")
    c = torch.zeros(1, 2, device="cuda")
    with pytest.raises(ValueError, match="Too little information"):
        fg.gaussian_sequential_sum_product(Lam, eta, c)


@pytest.mark.parametrize("B,T,D,O", [(3, 37, 4, 2), (2, 64, 32, 16), (5, 1, 3, 2), (1, 9, 8, 8)] + [(2, T, 3, 2) for T in range(2, 36)]
                         + [(2, 100, 16, 8), (1, 1023, 32, 16), (3, 1000, 5, 3), (2, 77, 32, 20)])
def test_kalman_homogeneous_path_matches_general_path(B, T, D, O):
    """kalman_chain (time-invariant prec_sqrt, Lam formed once per chain, level 1 reads it with stride 0) against the general
    path that materialises every element."""
    import funsor_b200.gaussian as fg

    g = torch.Generator().manual_seed(51)
    F = 0.9 * torch.linalg.qr(torch.randn(B, D, D, generator=g))[0]
    H = torch.randn(B, D, O, generator=g) / D ** 0.5

    def tril(n):
        a = torch.randn(B, n, n, generator=g) * 0.2
        return torch.tril(a, -1) + torch.diag_embed(0.5 + torch.rand(B, n, generator=g))

    args = [F, torch.randn(B, D, generator=g) * 0.1, tril(D), H, torch.randn(B, O, generator=g) * 0.1, tril(O), torch.randn(B, T, O, generator=g)]
    args = [a.cuda() for a in args]
    got = fg.kalman_chain(*args)
    w, P, s = fg.kalman_elements_sqrt(*args)
    want = fg.gaussian_sequential_sum_product(*fg.sqrt_to_info(w, P, s))
    for a, b in zip(got, want):
        close_norm(host(a), host(b).astype(np.float64), 2e-5)


@pytest.mark.parametrize("T", [4, 5, 6, 7, 11, 12, 13, 29, 64, 65, 1000, 4097])
@pytest.mark.parametrize("D,O", [(2, 1), (32, 16)])
def test_kalman_homogeneous_operator_path_against_float64_filter(fg, T, D, O):
    """The per-level-operator path for time-homogeneous chains (eta-only sweeps; every tail / odd-leftover combination of the
    reference's tree, sum_product.py:690-701) against the float64 oracle: chain elements and, through them, the log-likelihood."""
    rng = np.random.default_rng(T * 100 + D)
    B = 2
    F = 0.9 * np.linalg.qr(rng.standard_normal((B, D, D)))[0]
    H = rng.standard_normal((B, D, O)) / np.sqrt(D)

    def spd_tril(n):
        a = rng.standard_normal((B, n, n))
        return np.linalg.cholesky(a @ a.transpose(0, 2, 1) / n + 0.5 * np.eye(n))

    args = (F, 0.1 * rng.standard_normal((B, D)), spd_tril(D), H, 0.1 * rng.standard_normal((B, O)), spd_tril(O), rng.standard_normal((B, T, O)))
    chain = fg.kalman_chain(*[dev(a) for a in args])
    want = G.chain_info(*G.kalman_elements_info(*args), D)
    for a, b in zip(chain, want):
        close_norm(host(a), b, 1e-5)
    m0, p0 = rng.standard_normal((B, D)), spd_tril(D)
    lp = fg.gaussian_hmm_log_prob(dev(m0), dev(p0), chain)
    close_norm(host(lp), G.hmm_logprob_from_chain(G.mvn_info(m0, p0), want, D), 1e-5)


def _kalman_args(B, T, D, O, seed):
    g = torch.Generator().manual_seed(seed)
    F = 0.9 * torch.linalg.qr(torch.randn(B, D, D, generator=g))[0]
    H = torch.randn(B, D, O, generator=g) / D ** 0.5

    def tril(n):
        a = torch.randn(B, n, n, generator=g) * 0.2
        return torch.tril(a, -1) + torch.diag_embed(0.5 + torch.rand(B, n, generator=g))

    args = [F, torch.randn(B, D, generator=g) * 0.1, tril(D), H, torch.randn(B, O, generator=g) * 0.1, tril(O), torch.randn(B, T, O, generator=g)]
    return [a.cuda() for a in args]


@pytest.mark.parametrize("B,D,O", [(3, 4, 2), (2, 32, 16), (5, 3, 1), (1, 8, 8), (2, 17, 5), (64, 32, 16)])
def test_kalman_constants_kernel_matches_the_torch_statement(fg, B, D, O):
    """fb2_gaussian_kalman_constants_f32 against _kalman_constants + sqrt_to_info evaluated in float64 by torch."""
    args = _kalman_args(B, 4, D, O, 7 + D)
    got = fg.kalman_constants(*args[:6])
    a64 = [a.double() for a in args[:6]]
    P, Pr, w_t, w_o, s = fg._kalman_constants(*a64)
    want = (P @ P.transpose(-1, -2), Pr, w_o, (P[:, :, :D] @ w_t[:, :, None])[:, :, 0], P[:, :, D:], s - 0.5 * (w_t * w_t).sum(-1))
    for a, b in zip(got, want):
        assert a.shape == b.shape
        close_norm(host(a), host(b), 1e-5)


@pytest.mark.parametrize("K", [1, 2, 3, 4, 5, 6])
@pytest.mark.parametrize("D,O", [(3, 2), (32, 16), (8, 5)])
def test_kalman_fused_tiles_are_bit_identical_to_the_level_by_level_path(fg, monkeypatch, K, D, O):
    """gaussian_kalman_tile (elements + the first K tree levels per tile of 2^K steps in registers, remainder tile level by level) must
    walk the reference's tree (sum_product.py:690-701) exactly: every T from 2 to 70 plus a few long ones, against kalman_eta +
    gaussian_homog_eta on the same constants, compared with torch.equal."""
    monkeypatch.setenv("FB2_KALMAN_K", str(K))
    monkeypatch.setenv("FB2_KALMAN_TC", "0")  # the FFMA tile kernel (the tensor-core one is not bit-identical: next test)
    B = 2
    lengths = list(range(2, 71)) + [127, 128, 129, 1000, 1023, 4097] if D <= 8 else [2, 3, 5, 16, 17, 31, 33, 47, 64, 65, 100, 1023, 4097]
    for T in lengths:
        args = _kalman_args(B, T, D, O, 100 * T + D)
        consts = fg.kalman_constants(*args[:6])
        got = fg.kalman_chain_fused(consts, args[6])
        if T >= 4:
            elems = fg.kalman_eta_from_constants(consts, args[6])
            monkeypatch.setenv("FB2_GAUSS_HOMOG_TILES", "0")  # the level-by-level path
            want = fg.gaussian_chain_info_homog(*elems)
            monkeypatch.setenv("FB2_GAUSS_HOMOG_TILES", "1")  # the tile walk over given elements (what the funsor rule reaches)
            given = fg.gaussian_chain_info_homog(*elems)
            for a, b, c in zip(got, want, given):
                assert torch.equal(a, b), "T={} K={}: max diff {}".format(T, K, (a - b).abs().max().item())
                assert torch.equal(c, b), "given elements, T={} K={}: max diff {}".format(T, K, (c - b).abs().max().item())
        w, P, s = fg.kalman_elements_sqrt(*args)
        ref = fg.gaussian_sequential_sum_product(*fg.sqrt_to_info(w, P, s))
        for a, b in zip(got, ref):
            close_norm(host(a), host(b), 2e-5)


@pytest.mark.parametrize("K", [1, 3, 4, 6])
@pytest.mark.parametrize("D,O", [(32, 16), (20, 7), (17, 16)])
def test_kalman_tensor_core_tiles_match_the_level_by_level_path(fg, monkeypatch, K, D, O):
    """gaussian_kalman_tile_tc (D > 16: the tile walk as 3xTF32 mma.sync GEMMs against the shared operators, 16 tiles per warp) against
    the level-by-level path on the same constants (5e-6 of each tensor's largest entry) and against the general scan."""
    monkeypatch.setenv("FB2_KALMAN_K", str(K))
    monkeypatch.setenv("FB2_KALMAN_TC", "1")
    for B, T in [(2, 2), (1, 3), (2, 17), (3, 64), (2, 65), (1, 100), (2, 1023), (1, 1024), (2, 1040), (1, 4097), (5, 1087)]:
        args = _kalman_args(B, T, D, O, 200 * T + D)
        consts = fg.kalman_constants(*args[:6])
        got = fg.kalman_chain_fused(consts, args[6])
        if T >= 4:
            elems = fg.kalman_eta_from_constants(consts, args[6])
            monkeypatch.setenv("FB2_GAUSS_HOMOG_TILES", "0")  # the level-by-level path
            want = fg.gaussian_chain_info_homog(*elems)
            monkeypatch.setenv("FB2_GAUSS_HOMOG_TILES", "1")  # the tensor-core tile walk over given elements
            given = fg.gaussian_chain_info_homog(*elems)
            for a, b, c in zip(got, want, given):
                close_norm(host(a), host(b), 5e-6)
                close_norm(host(c), host(b), 5e-6)
        w, P, s = fg.kalman_elements_sqrt(*args)
        ref = fg.gaussian_sequential_sum_product(*fg.sqrt_to_info(w, P, s))
        for a, b in zip(got, ref):
            close_norm(host(a), host(b), 2e-5)


@pytest.mark.parametrize("D,O", [(3, 2), (8, 5), (16, 16), (32, 16), (21, 9)])
def test_cta_wide_level_operators_are_bit_identical_to_the_warp_version(fg, monkeypatch, D, O):
    """gaussian_homog_levels_cta (256 threads per chain, level precisions in shared memory) against gaussian_homog_levels (one warp per
    pair contraction): same operation order per output, so the whole scan must agree bit for bit -- every tail / leftover combination."""
    monkeypatch.setenv("FB2_KALMAN_TC", "0")
    for T in list(range(4, 40)) + [64, 65, 127, 1000, 1023, 4097]:
        args = _kalman_args(2, T, D, O, 300 * T + D)
        consts = fg.kalman_constants(*args[:6])
        elems = fg.kalman_eta_from_constants(consts, args[6])
        monkeypatch.setenv("FB2_GAUSS_LEVELS_WARP", "1")
        want = fg.gaussian_chain_info_homog(*elems)
        want_fused = fg.kalman_chain_fused(consts, args[6])
        monkeypatch.setenv("FB2_GAUSS_LEVELS_WARP", "0")
        got = fg.gaussian_chain_info_homog(*elems)
        got_fused = fg.kalman_chain_fused(consts, args[6])
        for a, b in zip(got + got_fused, want + want_fused):
            assert torch.equal(a, b), "T={}: max diff {}".format(T, (a - b).abs().max().item())


# ------------------------------------------------------------------------------------------------ (f2) Gaussian backward sampling
@pytest.mark.gpu
@pytest.mark.parametrize("B,T,D,O", [(2, 2, 1, 1), (2, 7, 3, 2), (1, 8, 2, 1), (2, 13, 4, 2), (1, 33, 2, 2)])
def test_kalman_posterior_sample_has_the_exact_joint_law(B, T, D, O):
    """Forward-filter backward-sample (recipes.py:16-90; Gaussian._sample gaussian.py:946-1059) on the scan tree.  The sampler is an
    affine map of its white noise, x = x(0) + A eps, so the joint law is verified EXACTLY against the dense float64 posterior of all
    states (oracle/gaussian_ref.kalman_posterior_dense, no scan): x(0) must be the posterior mean and A A' the posterior covariance;
    the reported log density must be the dense Gaussian log density of every draw."""
    import funsor_b200.gaussian as fg
    from oracle import gaussian_ref as G

    rng = np.random.default_rng(100 * T + D)
    n = (T + 1) * D

    def spd_tril(k):
        a = rng.standard_normal((B, k, k))
        return np.linalg.cholesky(a @ a.transpose(0, 2, 1) / k + 0.5 * np.eye(k))

    F = 0.9 * np.linalg.qr(rng.standard_normal((B, D, D)))[0]
    H = rng.standard_normal((B, D, O))
    args64 = (rng.standard_normal((B, D)), spd_tril(D), F, 0.1 * rng.standard_normal((B, D)), spd_tril(D), H,
              0.1 * rng.standard_normal((B, O)), spd_tril(O), rng.standard_normal((B, T, O)))
    mean, cov, logZ = G.kalman_posterior_dense(*args64)
    args = [torch.tensor(a, dtype=torch.float32, device="cuda") for a in args64]
    # noise: sample 0 = zeros, samples 1..n = unit vectors, then a few random draws
    extra = 5
    noise = np.zeros((B, 1 + n + extra, T + 1, D), np.float32)
    for k in range(n):
        noise[:, 1 + k].reshape(B, n)[:, k] = 1.0
    noise[:, 1 + n:] = rng.standard_normal((B, extra, T + 1, D))
    x, log_q, log_prob = fg.kalman_posterior_sample(*args, noise=torch.tensor(noise, device="cuda"))
    x = x.cpu().double().numpy().reshape(B, 1 + n + extra, n)
    scale = np.abs(mean).max() + np.sqrt(np.abs(cov).max())
    assert np.abs(x[:, 0] - mean).max() <= 2e-4 * scale, np.abs(x[:, 0] - mean).max()
    A = (x[:, 1 : 1 + n] - x[:, :1]).transpose(0, 2, 1)  # column k = response to unit noise k
    assert np.abs(A @ A.transpose(0, 2, 1) - cov).max() <= 5e-4 * np.abs(cov).max(), np.abs(A @ A.transpose(0, 2, 1) - cov).max()
    np.testing.assert_allclose(log_prob.cpu().numpy(), logZ, rtol=1e-4, atol=1e-3)
    # affine in the noise, and the density of every draw
    sign, logdet = np.linalg.slogdet(cov)
    for k in range(extra):
        xs = x[:, 1 + n + k]
        want_x = mean + (A @ noise[:, 1 + n + k].reshape(B, n, 1).astype(np.float64))[..., 0]
        assert np.abs(xs - want_x).max() <= 1e-3 * scale
        d = xs - mean
        want_q = -0.5 * (d * np.linalg.solve(cov, d[..., None])[..., 0]).sum(-1) - 0.5 * logdet - 0.5 * n * np.log(2 * np.pi)
        np.testing.assert_allclose(log_q[:, 1 + n + k].cpu().numpy(), want_q, rtol=2e-3, atol=2e-2)


@pytest.mark.gpu
def test_kalman_posterior_sample_moments_at_scale():
    """C4-like shape (D=32, O=16) on a longer chain: sample mean over many draws against the smoothed mean of the dense oracle at a
    few time points (T small enough for the dense float64 solve)."""
    import funsor_b200.gaussian as fg
    from oracle import gaussian_ref as G

    rng = np.random.default_rng(7)
    B, T, D, O, N = 1, 40, 32, 16, 4096
    F = 0.9 * np.linalg.qr(rng.standard_normal((B, D, D)))[0]
    H = rng.standard_normal((B, D, O)) / np.sqrt(D)
    eye = lambda k: np.broadcast_to(np.eye(k), (B, k, k)).copy()  # noqa: E731
    args64 = (np.zeros((B, D)), eye(D), F, np.zeros((B, D)), eye(D), H, np.zeros((B, O)), eye(O), rng.standard_normal((B, T, O)))
    mean, cov, _ = G.kalman_posterior_dense(*args64)
    args = [torch.tensor(a, dtype=torch.float32, device="cuda") for a in args64]
    g = torch.Generator(device="cuda").manual_seed(0)
    x, log_q, _ = fg.kalman_posterior_sample(*args, num_samples=N, generator=g)
    xm = x.double().mean(1).cpu().numpy().reshape(B, -1)
    sd = np.sqrt(np.diagonal(cov, axis1=-2, axis2=-1))
    assert np.abs((xm - mean) / sd).max() < 6.0 / np.sqrt(N) * 1.2  # 6 sigma of the Monte-Carlo error of a mean
    xv = x.double().var(1).cpu().numpy().reshape(B, -1)
    assert np.abs(xv / sd ** 2 - 1).max() < 0.2
    assert bool(torch.isfinite(log_q).all())


# ------------------------------------------------------------------------------------------------ (f3) moment matching
@pytest.mark.gpu
def test_moment_match_kernel_golden():
    """fb2_gaussian_moment_match_f32 against tests/golden/moment_matching.npz (the reference's `moment_matching` interpretation,
    joint.py:102-150, run by oracle/make_golden.py)"""
    import funsor_b200.gaussian as fg

    for name, d in golden_cases(load_golden("moment_matching")).items():
        w, P = torch.tensor(d["w"], dtype=torch.float32, device="cuda"), torch.tensor(d["P"], dtype=torch.float32, device="cuda")
        Lam, eta, c = fg.sqrt_to_info(w, P)
        Lo, eo, co = fg.gaussian_moment_match(torch.tensor(d["logits"], dtype=torch.float32, device="cuda"), Lam, eta, c)
        for label, a, b in (("Lam", Lo, d["Lam"]), ("eta", eo, d["eta"]), ("c", co, d["c"])):
            a = a.cpu().double().numpy()
            err = np.abs(a - b).max() / max(1.0, np.abs(b).max())
            assert err <= (5e-4 if name == "m4" else 5e-5), (name, label, err)  # m4: 64 reals, random (ill-conditioned) precisions


@pytest.mark.gpu
@pytest.mark.parametrize("N,K,n", [(5, 1, 3), (7, 2, 1), (100, 4, 8), (33, 6, 32), (9, 3, 64), (3, 17, 5)])
def test_moment_match_kernel_random(N, K, n):
    import funsor_b200.gaussian as fg

    rng = np.random.default_rng(N + K + n)
    A = rng.standard_normal((N, K, n, 2 * n))
    Lam = A @ A.transpose(0, 1, 3, 2) / (2 * n) + 0.5 * np.eye(n)  # well-conditioned component precisions
    mu = rng.standard_normal((N, K, n))
    eta = (Lam @ mu[..., None])[..., 0]
    c = rng.standard_normal((N, K))
    logits = rng.standard_normal((N, K))
    if K > 1:
        logits[0, 0] = -np.inf
        logits[-1, :] = -np.inf  # an empty mixture: mass -inf, unit covariance (joint.py:141-143)
    want = G.moment_match(logits, Lam, eta, c)
    got = fg.gaussian_moment_match(*[torch.tensor(a, dtype=torch.float32, device="cuda") for a in (logits, Lam, eta, c)])
    for label, a, b in zip(("Lam", "eta", "c"), got, want):
        a = a.cpu().double().numpy()
        fin = np.isfinite(b)
        assert np.array_equal(np.isfinite(a), fin), label
        assert np.abs(a[fin] - b[fin]).max() <= 2e-5 * max(1.0, np.abs(b[fin]).max()), (label, np.abs(a[fin] - b[fin]).max())
