"""GPU: ABSOLUTE parity at BASELINE.json's full sizes, against a float64 evaluation of the same recurrences on the device
(the reference's own data flow cannot exist at these sizes: C2's dense trans is 4.4 TB).

  C2  forward log-likelihood T=65536, S=1024, B=16      vs float64 recurrence, 1e-5 relative per chain (north_star tolerance)
  C5  forward-backward T=1,048,576, S=512               logZ 1e-5 relative; gamma rows at sampled times vs float64 alpha/beta
  C4  Kalman scan D=32, O=16 (chains of 1e3 .. 1e5)     (Lam, eta, c) vs a float64 evaluation of the same tree

The float64 references are plain torch on the GPU (allowed for floating-point kernels); they follow the recurrences of
examples/forward_backward.py:58-107 and gaussian.py:841-906, 1061-1090 in information form.
"""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def fb():
    import funsor_b200

    return funsor_b200


def _fp64_forward(init, A, obs, block=4096, keep=None, reverse=False):
    """float64 scaled forward recursion v_t = (v_{t-1} P) * e_t on the device; returns logZ[B] and {t: log-message[B,S]} for t in keep.
    With reverse: the backward recursion b_{t-1} = P (e_t * b_t) started from ones, messages stored BEFORE absorbing e_t."""
    B, T, S = obs.shape
    P = A.double().exp()
    if reverse:
        P = P.t().contiguous()
    v = torch.ones(B, S, dtype=torch.float64, device=obs.device) if reverse else init.double().exp().expand(B, S).contiguous()
    acc = torch.zeros(B, dtype=torch.float64, device=obs.device)
    kept = {}
    keep = set() if keep is None else set(int(k) for k in keep)
    order = range(T - 1, -1, -1) if reverse else range(T)
    blocks = [list(order)[i : i + block] for i in range(0, T, block)]
    for ts in blocks:
        lo, hi = min(ts), max(ts) + 1
        E = obs[:, lo:hi].double().exp()
        for t in ts:
            if reverse:
                if t in keep:
                    kept[t] = v.log() + acc[:, None]
                v = (v * E[:, t - lo]) @ P
            else:
                v = (v @ P) * E[:, t - lo]
                if t in keep:
                    kept[t] = v.log() + acc[:, None]
            if (t & 7) == 0:
                s = v.sum(-1)
                v = v / s[:, None]
                acc = acc + s.log()
    if reverse:
        return acc, kept, v
    return acc + v.sum(-1).log(), kept


def test_c2_full_size_absolute_against_float64(fb):
    T, S, B = 65536, 1024, 16
    g = torch.Generator(device=DEV).manual_seed(0)
    A = torch.randn(S, S, device=DEV, generator=g).log_softmax(-1)
    init = torch.randn(B, S, device=DEV, generator=g).log_softmax(-1)
    obs = torch.randn(B, T, S, device=DEV, generator=g)
    z = fb.hmm_forward(init, A, obs).double()
    ref, _ = _fp64_forward(init, A, obs)
    rel = ((z - ref).abs() / ref.abs()).max().item()
    print("C2 full size: max relative error of logZ over 16 chains = {:.3e} (logZ ~ {:.1f})".format(rel, ref.mean().item()))
    assert rel <= 1e-5, (z, ref)
    assert rel <= 2e-6  # the achieved level (documented in DESIGN.md); a regression beyond it is a bug even inside 1e-5


def test_c5_full_size_absolute_against_float64(fb):
    T, S = 1 << 20, 512
    g = torch.Generator(device=DEV).manual_seed(0)
    A = torch.randn(S, S, device=DEV, generator=g).log_softmax(-1)
    init = torch.randn(1, S, device=DEV, generator=g).log_softmax(-1)
    obs = torch.randn(1, T, S, device=DEV, generator=g)
    logZ, gamma, xi = fb.hmm_marginals(init, A, obs, want_xi=True)
    rng = np.random.default_rng(5)
    keep = sorted(set([0, 1, 2, T // 2, T - 2, T - 1] + [int(t) for t in rng.integers(0, T, 58)]))
    ref, alpha = _fp64_forward(init, A, obs, keep=keep)
    _, beta, _ = _fp64_forward(init, A, obs, keep=keep, reverse=True)
    rel = ((logZ.double() - ref).abs() / ref.abs()).item()
    print("C5 full size: relative error of logZ = {:.3e} (logZ = {:.2f})".format(rel, ref.item()))
    assert rel <= 1e-5
    worst = 0.0
    for t in keep:
        want = alpha[t] + beta[t] - ref[:, None]  # log p(z_t | obs)
        got = gamma[:, t].double()
        live = want > want.max() - 60.0  # entries below e^-60 of the row maximum are outside fp32 range of the linear-domain kernels
        d = (got - want)[live].abs().max().item()
        worst = max(worst, d)
        assert (got[~live] <= want[~live] + 1e-3).all()
        assert abs(torch.logsumexp(got, -1).item()) < 1e-3, t
    print("C5 full size: max |log gamma - float64| over {} sampled rows = {:.3e}".format(len(keep), worst))
    assert worst <= 2e-3
    assert abs(xi.double().sum().item() / T - 1.0) < 1e-3


# ------------------------------------------------------------------------------------------------ C4: Kalman scan vs float64
def _pair_combine_f64(X, Y, D):
    """Schur complement on the shared block of info-form elements (gaussian.py:841-906, 1061-1090 restated): float64, batched"""
    Lx, ex, cx = X
    Ly, ey, cy = Y
    M = Lx[..., D:, D:] + Ly[..., :D, :D]
    G = torch.cat([Lx[..., :D, D:], Ly[..., D:, :D]], -2)  # [2D, D]
    z = ex[..., D:] + ey[..., :D]
    L = torch.linalg.cholesky(M)
    U = torch.linalg.solve_triangular(L, G.transpose(-1, -2), upper=False)  # [D, 2D]
    v = torch.linalg.solve_triangular(L, z[..., None], upper=False)[..., 0]
    base = torch.zeros_like(Lx)
    base[..., :D, :D] = Lx[..., :D, :D]
    base[..., D:, D:] = Ly[..., D:, D:]
    Lo = base - U.transpose(-1, -2) @ U
    eo = torch.cat([ex[..., :D], ey[..., D:]], -1) - (U.transpose(-1, -2) @ v[..., None])[..., 0]
    co = cx + cy + 0.5 * D * math.log(2 * math.pi) - L.diagonal(dim1=-1, dim2=-2).log().sum(-1) + 0.5 * (v * v).sum(-1)
    return Lo, eo, co


def _chain_f64(Lam, eta, c, D):
    """the reference's tree (sum_product.py:690-701) over elements [B, T, ...] in float64"""
    T = eta.shape[1]
    while T > 1:
        even = T // 2 * 2
        X = (Lam[:, 0:even:2], eta[:, 0:even:2], c[:, 0:even:2])
        Y = (Lam[:, 1:even:2], eta[:, 1:even:2], c[:, 1:even:2])
        Lo, eo, co = _pair_combine_f64(X, Y, D)
        if even < T:
            Lo, eo, co = torch.cat([Lo, Lam[:, T - 1 :]], 1), torch.cat([eo, eta[:, T - 1 :]], 1), torch.cat([co, c[:, T - 1 :]], 1)
        Lam, eta, c = Lo, eo, co
        T = eta.shape[1]
    return Lam[:, 0], eta[:, 0], c[:, 0]


def _kalman_inputs(B, T, D, O, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    F = 0.9 * torch.linalg.qr(torch.randn(B, D, D, device=DEV, generator=g))[0]
    H = torch.randn(B, D, O, device=DEV, generator=g) / D ** 0.5

    def spd_tril(n):
        a = torch.randn(B, n, n, device=DEV, generator=g)
        return torch.linalg.cholesky(a @ a.transpose(-1, -2) / n + 0.5 * torch.eye(n, device=DEV))

    y = torch.randn(B, T, O, device=DEV, generator=g)
    return F, 0.1 * torch.randn(B, D, device=DEV, generator=g), spd_tril(D), H, 0.1 * torch.randn(B, O, device=DEV, generator=g), spd_tril(O), y


@pytest.mark.parametrize("T", [1000, 10000, 100000])
def test_c4_kalman_scan_elementwise_against_float64(T):
    """C4's shape (D=32, O=16, SPD noise covariances) at chain lengths up to the full 1e5: the homogeneous-dynamics scan in fp32
    against the same tree in float64.  Information form carries Lam = P P' (condition number squared w.r.t. the reference's
    square-root form), so the bound is stated per quantity: the reference comparator |d| <= rtol (1 + |expected|)
    (funsor/testing.py:176-178 with atol = 1) at rtol = 1e-5 for Lam and eta, and 1e-5 RELATIVE for the log-normaliser c."""
    import funsor_b200.gaussian as fg

    B, D, O = 4, 32, 16
    args = _kalman_inputs(B, T, D, O, seed=T)
    got = fg.kalman_chain(*args)
    w, P, s = fg.kalman_elements_homog(*args)
    Lam1 = (P.double() @ P.double().transpose(-1, -2))  # [B, 2D, 2D]
    eta = (P.double()[:, None] @ w.double()[..., None])[..., 0]  # [B, T, 2D]
    c = s.double() - 0.5 * (w.double() ** 2).sum(-1)
    want = _chain_f64(Lam1[:, None].expand(B, T, 2 * D, 2 * D), eta, c, D)
    report = {}
    for name, a, b in zip(("Lam", "eta", "c"), got, want):
        d = (a.double() - b).abs()
        report[name] = (float((d / (1.0 + b.abs())).max()), float(d.max() / b.abs().max()))
    print("C4 T={}: comparator / max-norm relative errors: {}".format(T, {k: ("%.2e" % v[0], "%.2e" % v[1]) for k, v in report.items()}))
    assert report["Lam"][0] <= 1e-5 and report["eta"][0] <= 1e-5, report
    assert report["c"][1] <= 1e-5, report


def test_general_gaussian_scan_elementwise_against_float64():
    """time-VARYING precision (the general pair-combine kernel) at D=32: same statement as above"""
    import funsor_b200.gaussian as fg

    B, T, D, O = 2, 4097, 32, 16
    args = _kalman_inputs(B, T, D, O, seed=3)
    w, P, s = fg.kalman_elements_sqrt(*args)
    Lam, eta, c = fg.sqrt_to_info(w, P, s)
    got = fg.gaussian_sequential_sum_product(Lam, eta, c)
    Pd, wd = P.double(), w.double()
    want = _chain_f64(Pd @ Pd.transpose(-1, -2), (Pd @ wd[..., None])[..., 0], s.double() - 0.5 * (wd ** 2).sum(-1), D)
    for name, a, b in zip(("Lam", "eta", "c"), got, want):
        d = (a.double() - b).abs()
        if name == "c":
            assert float(d.max() / b.abs().max()) <= 1e-5, name
        else:
            assert float((d / (1.0 + b.abs())).max()) <= 1e-5, (name, float((d / (1.0 + b.abs())).max()))
