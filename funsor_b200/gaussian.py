"""Host-side mirror of the reference's Gaussian-factor chain path on raw torch CUDA tensors.

Chain elements are carried in information form (Lam[..., 2D, 2D], eta[..., 2D], c[...]) over rows
(prev; curr): f(x) = c - 1/2 x Lam x' + x eta.  This is the representation-independent content of the
reference's (white_vec, prec_sqrt, scalar Tensor) triple (funsor/gaussian.py:419-498, 545, 569;
funsor/testing.py:149-153 compares exactly these).

Mirrors: eager_add_gaussian_gaussian gaussian.py:1117-1132, Gaussian.eager_reduce 841-906,
_marginalize_after_split 1061-1090, eager_contraction_gaussian cnf.py:385-389,
sequential_sum_product sum_product.py:652-701, matrix_and_mvn_to_funsor pyro/convert.py:278-298,
GaussianHMM pyro/hmm.py:258-274.
"""
import math
import os

import torch

from . import _lib
from ._lib import check
from .ops import _dev_f32, _ptr, _stream, _workspace


def sqrt_to_info(white_vec, prec_sqrt, scalar=None):
    """(w[..., r], P[..., n, r], s[...]) -> (Lam[..., n, n], eta[..., n], c[...]) with one fused kernel."""
    _dev_f32(white_vec, "white_vec"), _dev_f32(prec_sqrt, "prec_sqrt")
    batch = prec_sqrt.shape[:-2]
    n, r = prec_sqrt.shape[-2:]
    w = white_vec.expand(batch + (r,)).contiguous()
    P = prec_sqrt.contiguous()
    s = None if scalar is None else _dev_f32(scalar, "scalar").expand(batch).contiguous()
    N = int(torch.Size(batch).numel())
    Lam = torch.empty(batch + (n, n), dtype=torch.float32, device=P.device)
    eta = torch.empty(batch + (n,), dtype=torch.float32, device=P.device)
    c = torch.empty(batch, dtype=torch.float32, device=P.device)
    lib = _lib.load()
    with torch.cuda.device(P.device):
        check(lib.fb2_gaussian_sqrt_to_info_f32(w.data_ptr(), P.data_ptr(), _ptr(s), N, n, r, Lam.data_ptr(), eta.data_ptr(),
                                                c.data_ptr(), _stream()))
    return Lam, eta, c


def info_to_sqrt(Lam, eta, c, rel_tol=0.0, return_rank=False):
    """(Lam[..., n, n], eta[..., n], c[...]) -> (white_vec[..., n], prec_sqrt[..., n, n], scalar[...]) with P P' = Lam, P w = eta,
    s - |w|^2/2 = c: the inverse of sqrt_to_info, so that results can be handed back as the reference's Gaussian(white_vec,
    prec_sqrt) + Tensor (gaussian.py:419-498).  Lam may be rank deficient: diagonally pivoted Cholesky, one warp per element."""
    for t in (Lam, eta, c):
        _dev_f32(t, "gaussian element")
    batch = Lam.shape[:-2]
    n = Lam.shape[-1]
    Lam, eta, c = Lam.contiguous(), eta.expand(batch + (n,)).contiguous(), c.expand(batch).contiguous()
    N = int(torch.Size(batch).numel())
    w = torch.empty(batch + (n,), dtype=torch.float32, device=Lam.device)
    P = torch.empty(batch + (n, n), dtype=torch.float32, device=Lam.device)
    s = torch.empty(batch, dtype=torch.float32, device=Lam.device)
    rank = torch.empty(batch, dtype=torch.int32, device=Lam.device) if return_rank else None
    lib = _lib.load()
    with torch.cuda.device(Lam.device):
        check(lib.fb2_gaussian_info_to_sqrt_f32(Lam.data_ptr(), eta.data_ptr(), c.data_ptr(), N, n, float(rel_tol), w.data_ptr(),
                                                P.data_ptr(), s.data_ptr(), _ptr(rank), _stream()))
    return (w, P, s, rank) if return_rank else (w, P, s)


def gaussian_moment_match(logits, Lam, eta, c, validate=True):
    """Collapse mixtures of K Gaussians into single Gaussians with the mixture's mass, mean and covariance -- the
    `moment_matching` interpretation's Contraction[Logaddexp, Add, {discrete}, Tensor, Gaussian] (joint.py:102-150), used by the
    SwitchingLinearHMM scan (pyro/hmm.py:527-551).  logits[..., K], Lam[..., K, n, n], eta[..., K, n], c[..., K]
    -> (Lam[..., n, n], eta[..., n], c[...])."""
    for t_ in (logits, Lam, eta, c):
        _dev_f32(t_, "mixture component")
    batch = Lam.shape[:-3]
    K, n = Lam.shape[-3], Lam.shape[-1]
    if K < 1 or n < 1 or n > 64:
        raise ValueError("gaussian_moment_match: K >= 1 components over at most 64 real dimensions, got K={} n={}".format(K, n))
    Lam = Lam.contiguous()
    eta = eta.expand(batch + (K, n)).contiguous()
    c = c.expand(batch + (K,)).contiguous()
    logits = logits.expand(batch + (K,)).contiguous()
    N = int(torch.Size(batch).numel())
    Lo = torch.empty(batch + (n, n), dtype=torch.float32, device=Lam.device)
    eo = torch.empty(batch + (n,), dtype=torch.float32, device=Lam.device)
    co = torch.empty(batch, dtype=torch.float32, device=Lam.device)
    flag = torch.zeros(1, dtype=torch.int32, device=Lam.device)
    lib = _lib.load()
    ws = _workspace(lib.fb2_gaussian_moment_match_workspace_bytes(N, K, n), Lam.device)
    with torch.cuda.device(Lam.device):
        check(lib.fb2_gaussian_moment_match_f32(logits.data_ptr(), Lam.data_ptr(), eta.data_ptr(), c.data_ptr(), N, K, n, Lo.data_ptr(),
                                                eo.data_ptr(), co.data_ptr(), flag.data_ptr(), ws.data_ptr(), ws.numel(), _stream()))
    if validate and bool(flag.any()):
        raise ValueError("gaussian_moment_match: a component precision (or the matched covariance) is not positive definite")
    return Lo, eo, co


def gaussian_pair_combine(x, y):
    """Contraction[logaddexp, add] of two chain elements x over (a;b) and y over (b;c) (cnf.py:385-389).
    x, y: tuples (Lam[..., 2D, 2D], eta[..., 2D], c[...]) with identical batch shapes."""
    Lx, ex, cx = (t.contiguous() for t in x)
    Ly, ey, cy = (t.contiguous() for t in y)
    for t in (Lx, ex, cx, Ly, ey, cy):
        _dev_f32(t, "gaussian element")
    batch = Lx.shape[:-2]
    n = Lx.shape[-1]
    D = n // 2
    N = int(torch.Size(batch).numel())
    Lo, eo, co = torch.empty_like(Lx), torch.empty_like(ex), torch.empty_like(cx)
    flag = torch.zeros(1, dtype=torch.int32, device=Lx.device)
    lib = _lib.load()
    with torch.cuda.device(Lx.device):
        check(lib.fb2_gaussian_pair_combine_f32(Lx.data_ptr(), ex.data_ptr(), cx.data_ptr(), 1, Ly.data_ptr(), ey.data_ptr(),
                                                cy.data_ptr(), 1, 1, N, 0, 0, D, Lo.data_ptr(), eo.data_ptr(), co.data_ptr(),
                                                flag.data_ptr(), _stream()))
    return Lo, eo, co, flag


def gaussian_sequential_sum_product(Lam, eta, c, validate=True):
    """sequential_sum_product(logaddexp, add, ...) over Gaussian chain elements
    Lam[B, T, 2D, 2D], eta[B, T, 2D], c[B, T] -> (Lam[B, 2D, 2D], eta[B, 2D], c[B])."""
    for t in (Lam, eta, c):
        _dev_f32(t, "gaussian element")
    if Lam.dim() != 4 or Lam.shape[-1] != Lam.shape[-2] or Lam.shape[-1] % 2:
        raise ValueError("Lam must be [B, T, 2D, 2D], got {}".format(tuple(Lam.shape)))
    B, T, n, _ = Lam.shape
    D = n // 2
    Lam, eta, c = Lam.contiguous(), eta.expand(B, T, n).contiguous(), c.expand(B, T).contiguous()
    Lo = torch.empty((B, n, n), dtype=torch.float32, device=Lam.device)
    eo = torch.empty((B, n), dtype=torch.float32, device=Lam.device)
    co = torch.empty((B,), dtype=torch.float32, device=Lam.device)
    lib = _lib.load()
    ws = _workspace(lib.fb2_gaussian_chain_workspace_bytes(B, T, D), Lam.device)
    with torch.cuda.device(Lam.device):
        check(lib.fb2_gaussian_chain_f32(Lam.data_ptr(), eta.data_ptr(), c.data_ptr(), B, T, D, Lo.data_ptr(), eo.data_ptr(),
                                         co.data_ptr(), ws.data_ptr(), ws.numel(), _stream()))
    if validate and bool(torch.isnan(co).any()):
        raise ValueError("Too little information to marginalize over the shared state. Consider adding a prior.")
    return Lo, eo, co


def kalman_elements_homog(F, q_loc, q_tril, H, r_loc, r_tril, y):
    """GaussianHMM chain elements (pyro/hmm.py:258-270: trans + obs(value=y_t)) in the reference's sqrt form for time-homogeneous
    parameters, without the T-sized prec_sqrt: only white_vec depends on t, so P is built once per chain (_kalman_constants).
    Returns w[B,T,D+O], P[B,2D,D+O], s[B,T].  (kalman_eta goes one step further and never forms w either.)"""
    B, T, O = y.shape
    D = F.shape[-1]
    P, Pr, w_t, w_o, s = _kalman_constants(F, q_loc, q_tril, H, r_loc, r_tril)
    w = torch.cat([w_t[:, None, :].expand(B, T, D), w_o[:, None, :] + y @ Pr], dim=-1)
    return w.contiguous(), P, s[:, None].expand(B, T).contiguous()


def kalman_elements_sqrt(F, q_loc, q_tril, H, r_loc, r_tril, y):
    """The same elements with the prec_sqrt materialised per step, P[B,T,2D,D+O] -- the reference's own data flow (the Gaussian sum
    broadcasts over `time`); used by the tests of the general (time-varying) scan."""
    w, P, s = kalman_elements_homog(F, q_loc, q_tril, H, r_loc, r_tril, y)
    return w, P[:, None].expand(P.shape[0], y.shape[1], P.shape[1], P.shape[2]).contiguous(), s


def kalman_elements_timevarying(F, q_loc, q_tril, H, r_loc, r_tril, y):
    """Chain elements when any parameter depends on time (pyro/hmm.py:258-270 with heterogeneous inputs): every parameter is
    [B, T|1, ...] and is broadcast over the time axis.  Returns w[B,T,D+O], P[B,T,2D,D+O], s[B,T] (sqrt form)."""
    B, T, O = y.shape
    D = F.shape[-1]
    eyeD = torch.eye(D, device=F.device, dtype=F.dtype)
    eyeO = torch.eye(O, device=F.device, dtype=F.dtype)
    Pq = torch.linalg.solve_triangular(q_tril, eyeD.expand(q_tril.shape), upper=False).transpose(-1, -2)  # [B,Tq,D,D]
    Pr = torch.linalg.solve_triangular(r_tril, eyeO.expand(r_tril.shape), upper=False).transpose(-1, -2)  # [B,Tr,O,O]
    P = torch.zeros(B, T, 2 * D, D + O, device=F.device, dtype=F.dtype)
    P[:, :, :D, :D] = F @ Pq
    P[:, :, D:, :D] = -Pq
    P[:, :, D:, D:] = H @ Pr
    w_t = -(q_loc[..., None, :] @ Pq)[..., 0, :]  # [B,T|1,D]
    w_o = -(r_loc[..., None, :] @ Pr)[..., 0, :] + (y[..., None, :] @ Pr)[..., 0, :]  # [B,T,O]
    w = torch.cat([w_t.expand(B, T, D), w_o.expand(B, T, O)], dim=-1)
    s = (-0.5 * (D + O) * math.log(2 * math.pi) - q_tril.diagonal(dim1=-1, dim2=-2).log().sum(-1)
         - r_tril.diagonal(dim1=-1, dim2=-2).log().sum(-1))
    return w.contiguous(), P, s.expand(B, T).contiguous()


def gaussian_chain_info_homog(Lam, eta, c, validate=True):
    """Scan over info-form chain elements that share one precision per chain: Lam[B,2D,2D], eta[B,T,2D], c[B,T]."""
    for t in (Lam, eta, c):
        _dev_f32(t, "gaussian element")
    B, T, n = eta.shape
    D = n // 2
    Lam, eta, c = Lam.contiguous(), eta.contiguous(), c.expand(B, T).contiguous()
    Lo = torch.empty((B, n, n), dtype=torch.float32, device=Lam.device)
    eo = torch.empty((B, n), dtype=torch.float32, device=Lam.device)
    co = torch.empty((B,), dtype=torch.float32, device=Lam.device)
    lib = _lib.load()
    ws = _workspace(lib.fb2_gaussian_chain_workspace_bytes(B, T, D), Lam.device)
    with torch.cuda.device(Lam.device):
        check(lib.fb2_gaussian_chain_homog_f32(Lam.data_ptr(), eta.data_ptr(), c.data_ptr(), B, T, D, Lo.data_ptr(), eo.data_ptr(),
                                               co.data_ptr(), ws.data_ptr(), ws.numel(), _stream()))
    if validate and bool(torch.isnan(co).any()):
        raise ValueError("Too little information to marginalize over the shared state. Consider adding a prior.")
    return Lo, eo, co


def gaussian_chain_homog(white_vec, prec_sqrt, scalar, validate=True):
    """Scan over chain elements that share one prec_sqrt per chain: white_vec[B,T,r], prec_sqrt[B,2D,r], scalar[B,T].
    Lam[B] = P P' once per chain, eta / c per step; the T-sized precision array is never formed."""
    _dev_f32(white_vec, "white_vec"), _dev_f32(prec_sqrt, "prec_sqrt")
    B, T, r = white_vec.shape
    n = prec_sqrt.shape[-2]
    D = n // 2
    w = white_vec.contiguous()
    P = prec_sqrt.contiguous()
    s = None if scalar is None else _dev_f32(scalar, "scalar").expand(B, T).contiguous()
    Lam, _, _ = sqrt_to_info(torch.zeros(B, r, device=P.device), P)
    eta = torch.empty((B, T, n), dtype=torch.float32, device=P.device)
    c = torch.empty((B, T), dtype=torch.float32, device=P.device)
    Lo = torch.empty((B, n, n), dtype=torch.float32, device=P.device)
    eo = torch.empty((B, n), dtype=torch.float32, device=P.device)
    co = torch.empty((B,), dtype=torch.float32, device=P.device)
    lib = _lib.load()
    ws = _workspace(lib.fb2_gaussian_chain_workspace_bytes(B, T, D), P.device)
    with torch.cuda.device(P.device):
        check(lib.fb2_gaussian_eta_f32(w.data_ptr(), P.data_ptr(), _ptr(s), B, T, n, r, eta.data_ptr(), c.data_ptr(), _stream()))
        check(lib.fb2_gaussian_chain_homog_f32(Lam.data_ptr(), eta.data_ptr(), c.data_ptr(), B, T, D, Lo.data_ptr(), eo.data_ptr(),
                                               co.data_ptr(), ws.data_ptr(), ws.numel(), _stream()))
    if validate and bool(torch.isnan(co).any()):
        raise ValueError("Too little information to marginalize over the shared state. Consider adding a prior.")
    return Lo, eo, co


def _kalman_constants(F, q_loc, q_tril, H, r_loc, r_tril):
    """Per-chain constants of the GaussianHMM chain element (pyro/hmm.py:258-270; matrix_and_mvn_to_funsor pyro/convert.py:278-298):
    prec_sqrt P[B,2D,D+O] over rows (state; state'), the observation's Pr[B,O,O] and white_vec offset wo[B,O], w_trans[B,D] and the
    scalar s[B].  [B]-sized linear algebra in torch; nothing here depends on T."""
    B, D = F.shape[0], F.shape[-1]
    O = H.shape[-1]
    eyeD = torch.eye(D, device=F.device, dtype=F.dtype)
    eyeO = torch.eye(O, device=F.device, dtype=F.dtype)
    Pq = torch.linalg.solve_triangular(q_tril, eyeD.expand(B, D, D), upper=False).transpose(-1, -2)
    Pr = torch.linalg.solve_triangular(r_tril, eyeO.expand(B, O, O), upper=False).transpose(-1, -2)
    P = torch.zeros(B, 2 * D, D + O, device=F.device, dtype=F.dtype)
    P[:, :D, :D] = F @ Pq
    P[:, D:, :D] = -Pq
    P[:, D:, D:] = H @ Pr
    w_t = -(q_loc[:, None, :] @ Pq)[:, 0]
    w_o = -(r_loc[:, None, :] @ Pr)[:, 0]
    s = (-0.5 * (D + O) * math.log(2 * math.pi) - q_tril.diagonal(dim1=-1, dim2=-2).log().sum(-1)
         - r_tril.diagonal(dim1=-1, dim2=-2).log().sum(-1))
    return P, Pr.contiguous(), w_t, w_o.contiguous(), s


def kalman_eta(F, q_loc, q_tril, H, r_loc, r_tril, y):
    """Information-form chain elements of a GaussianHMM with fixed dynamics straight from the observations: (Lam[B,2D,2D] shared by all
    steps of a chain, eta[B,T,2D], c[B,T]).  One kernel reads y once and writes eta / c once (fb2_gaussian_kalman_eta_f32)."""
    _dev_f32(y, "y")
    B, T, O = y.shape
    D = F.shape[-1]
    P, Pr, w_t, w_o, s = _kalman_constants(F, q_loc, q_tril, H, r_loc, r_tril)
    P = P.contiguous()
    Lam, _, _ = sqrt_to_info(torch.zeros(B, D + O, device=P.device), P)
    a0 = (P[:, :, :D] @ w_t[:, :, None])[:, :, 0].contiguous()
    Po = P[:, :, D:].contiguous()
    c0 = (s - 0.5 * (w_t * w_t).sum(-1)).contiguous()
    y = y.contiguous()
    eta = torch.empty((B, T, 2 * D), dtype=torch.float32, device=P.device)
    c = torch.empty((B, T), dtype=torch.float32, device=P.device)
    lib = _lib.load()
    with torch.cuda.device(P.device):
        check(lib.fb2_gaussian_kalman_eta_f32(y.data_ptr(), Pr.data_ptr(), w_o.data_ptr(), a0.data_ptr(), Po.data_ptr(), c0.data_ptr(), B, T, D, O,
                                              eta.data_ptr(), c.data_ptr(), _stream()))
    return Lam, eta, c


def kalman_constants(F, q_loc, q_tril, H, r_loc, r_tril):
    """The t-independent part of the GaussianHMM chain element in one launch (fb2_gaussian_kalman_constants_f32; the same quantities as
    _kalman_constants + sqrt_to_info, which remain as the torch statement of them): (Lam[B,2D,2D], Pr[B,O,O], wo[B,O], a0[B,2D],
    Po[B,2D,O], c0[B])."""
    B, D = F.shape[0], F.shape[-1]
    O = H.shape[-1]
    args = [_dev_f32(t, "kalman parameter").expand(shape).contiguous() for t, shape in (
        (F, (B, D, D)), (q_loc, (B, D)), (q_tril, (B, D, D)), (H, (B, D, O)), (r_loc, (B, O)), (r_tril, (B, O, O)))]
    dev = F.device
    out = [torch.empty(shape, dtype=torch.float32, device=dev) for shape in ((B, 2 * D, 2 * D), (B, O, O), (B, O), (B, 2 * D), (B, 2 * D, O), (B,))]
    lib = _lib.load()
    with torch.cuda.device(dev):
        check(lib.fb2_gaussian_kalman_constants_f32(*[t.data_ptr() for t in args], B, D, O, *[t.data_ptr() for t in out], _stream()))
    return tuple(out)


def kalman_eta_from_constants(constants, y):
    """Per-step chain elements (Lam[B,2D,2D], eta[B,T,2D], c[B,T]) from kalman_constants and the observations (the level-by-level
    counterpart of kalman_chain_fused)."""
    Lam, Pr, wo, a0, Po, c0 = constants
    B, T, O = y.shape
    n = Lam.shape[-1]
    y = _dev_f32(y, "y").contiguous()
    eta = torch.empty((B, T, n), dtype=torch.float32, device=y.device)
    c = torch.empty((B, T), dtype=torch.float32, device=y.device)
    lib = _lib.load()
    with torch.cuda.device(y.device):
        check(lib.fb2_gaussian_kalman_eta_f32(y.data_ptr(), Pr.data_ptr(), wo.data_ptr(), a0.data_ptr(), Po.data_ptr(), c0.data_ptr(), B, T, n // 2, O,
                                              eta.data_ptr(), c.data_ptr(), _stream()))
    return Lam, eta, c


def kalman_chain_fused(constants, y, validate=True):
    """kalman_eta + gaussian_chain_info_homog in one library call (fb2_gaussian_kalman_chain_f32): per tile of 2^K steps the elements
    and the first K tree levels are evaluated in registers (D <= 16: fp32 FMA, bit-identical to the level-by-level path; D > 16: 3xTF32
    tensor-core GEMMs against the shared per-level operators, within 5e-6 of it)."""
    Lam, Pr, wo, a0, Po, c0 = constants
    _dev_f32(y, "y")
    B, T, O = y.shape
    n = Lam.shape[-1]
    D = n // 2
    y = y.contiguous()
    Lo = torch.empty((B, n, n), dtype=torch.float32, device=y.device)
    eo = torch.empty((B, n), dtype=torch.float32, device=y.device)
    co = torch.empty((B,), dtype=torch.float32, device=y.device)
    lib = _lib.load()
    ws = _workspace(lib.fb2_gaussian_kalman_chain_workspace_bytes(B, T, D), y.device)
    with torch.cuda.device(y.device):
        check(lib.fb2_gaussian_kalman_chain_f32(y.data_ptr(), Lam.data_ptr(), Pr.data_ptr(), wo.data_ptr(), a0.data_ptr(), Po.data_ptr(),
                                                c0.data_ptr(), B, T, D, O, Lo.data_ptr(), eo.data_ptr(), co.data_ptr(), ws.data_ptr(),
                                                ws.numel(), _stream()))
    if validate and bool(torch.isnan(co).any()):
        raise ValueError("Too little information to marginalize over the shared state. Consider adding a prior.")
    return Lo, eo, co


def kalman_chain(F, q_loc, q_tril, H, r_loc, r_tril, y, validate=True):
    """Parallel-scan Kalman filter for time-homogeneous parameters (examples/kalman_filter.py, GaussianHMM pyro/hmm.py:258-274):
    Returns (Lam[B,2D,2D], eta[B,2D], c[B]) over (state_{-1}; state_{T-1}), the value of the reference's MarkovProduct
    (pyro/hmm.py:269).  Time-varying elements go through kalman_elements_sqrt + sqrt_to_info + gaussian_sequential_sum_product."""
    if F.shape[-1] <= 32 and y.shape[-1] <= 16 and y.shape[1] >= 2 and os.environ.get("FB2_KALMAN_FUSED", "1") != "0":
        return kalman_chain_fused(kalman_constants(F, q_loc, q_tril, H, r_loc, r_tril), y, validate=validate)
    if F.shape[-1] <= 32 and y.shape[-1] <= 32:
        Lam, eta, c = kalman_eta(F, q_loc, q_tril, H, r_loc, r_tril, y)
        return gaussian_chain_info_homog(Lam, eta, c, validate=validate)
    w, P, s = kalman_elements_homog(F, q_loc, q_tril, H, r_loc, r_tril, y)
    return gaussian_chain_homog(w, P, s, validate=validate)


def kalman_posterior_sample(m0, p0_tril, F, q_loc, q_tril, H, r_loc, r_tril, y, num_samples=1, noise=None, generator=None):
    """Forward-filter backward-sample for a time-homogeneous linear-Gaussian state-space model: exact joint draws of the state path
    x_0 .. x_T from p(x | y) (recipes.forward_filter_backward_rsample recipes.py:16-90 specialised to a chain; Gaussian._sample
    gaussian.py:946-1059 draws every eliminated state given its two neighbours in the scan tree, sum_product.py:690-701).

    The forward sweep is the parallel-scan Kalman filter (kalman_chain) that keeps its levels; the two end states are drawn from the
    joint of the chain result and N(m0, p0 p0') (a [B, 2D, 2D] problem, torch.linalg in float64), then the tree is walked top-down on
    the GPU, all pairs and samples of a level in parallel.  `noise[B, N, T+1, D]` standard normal makes the draws a deterministic
    function of the inputs.  Returns (x[B,N,T+1,D], log_q[B,N] = log p(x | y), log_prob[B] = log p(y))."""
    B, T, O = y.shape
    D = F.shape[-1]
    if T < 2:
        raise ValueError("kalman_posterior_sample needs at least two time steps")
    if noise is None:
        noise = torch.randn((B, num_samples, T + 1, D), dtype=torch.float32, device=y.device, generator=generator)
    _dev_f32(noise, "noise")
    if noise.dim() != 4 or noise.shape[0] != B or noise.shape[2:] != (T + 1, D):
        raise ValueError("noise must be [B, N, T+1, D], got {}".format(tuple(noise.shape)))
    noise = noise.contiguous()
    N = noise.shape[1]
    w, P, s = kalman_elements_homog(F, q_loc, q_tril, H, r_loc, r_tril, y)
    n, r = 2 * D, w.shape[-1]
    P = P.contiguous()
    Lam, _, _ = sqrt_to_info(torch.zeros(B, r, device=P.device), P)
    eta = torch.empty((B, T, n), dtype=torch.float32, device=P.device)
    c = torch.empty((B, T), dtype=torch.float32, device=P.device)
    Lo = torch.empty((B, n, n), dtype=torch.float32, device=P.device)
    eo = torch.empty((B, n), dtype=torch.float32, device=P.device)
    co = torch.empty((B,), dtype=torch.float32, device=P.device)
    x = torch.empty((B, N, T + 1, D), dtype=torch.float32, device=P.device)
    lib = _lib.load()
    ws = _workspace(lib.fb2_gaussian_chain_sample_workspace_bytes(B, T, D), P.device)
    with torch.cuda.device(P.device):
        check(lib.fb2_gaussian_eta_f32(w.data_ptr(), P.data_ptr(), s.data_ptr(), B, T, n, r, eta.data_ptr(), c.data_ptr(), _stream()))
        check(lib.fb2_gaussian_chain_homog_keep_f32(Lam.data_ptr(), eta.data_ptr(), c.data_ptr(), B, T, D, Lo.data_ptr(), eo.data_ptr(),
                                                    co.data_ptr(), ws.data_ptr(), ws.numel(), _stream()))
    if bool(torch.isnan(co).any()):
        raise ValueError("Too little information to marginalize over the shared state. Consider adding a prior.")
    # the two end states: joint of the chain result over (x_0; x_T) and the initial distribution on x_0
    Lj, ej = Lo.double().clone(), eo.double().clone()
    eye = torch.eye(D, device=P.device, dtype=torch.float64)
    Li = torch.linalg.solve_triangular(p0_tril.double(), eye.expand(B, D, D), upper=False)
    L0 = Li.transpose(-1, -2) @ Li
    Lj[:, :D, :D] += L0
    ej[:, :D] += (L0 @ m0.double()[..., None])[..., 0]
    Lc = torch.linalg.cholesky(Lj)
    v = torch.linalg.solve_triangular(Lc, ej[..., None], upper=False)[..., 0]  # [B, 2D]
    eps = torch.cat([noise[:, :, 0], noise[:, :, T]], -1).double()  # [B, N, 2D]
    ends = torch.linalg.solve_triangular(Lc.transpose(-1, -2)[:, None], (eps + v[:, None])[..., None], upper=True)[..., 0]
    x[:, :, 0] = ends[..., :D].float()
    x[:, :, T] = ends[..., D:].float()
    with torch.cuda.device(P.device):
        check(lib.fb2_gaussian_chain_homog_backward_sample_f32(eta.data_ptr(), B, T, D, N, noise.data_ptr(), x.data_ptr(), ws.data_ptr(),
                                                               ws.numel(), _stream()))
    log_prob = gaussian_hmm_log_prob(m0, p0_tril, (Lo, eo, co))
    # log density of the draws: sum of the factors at the sample minus the normaliser (recipes.py:80-88)
    pairs = torch.cat([x[:, :, :-1], x[:, :, 1:]], -1)  # [B, N, T, 2D]
    resid = pairs.reshape(B, N * T, n) @ P - w[:, None].expand(B, N, T, r).reshape(B, N * T, r)
    elem = s[:, None] - 0.5 * (resid * resid).sum(-1).reshape(B, N, T)
    d0 = (x[:, :, 0].double() - m0.double()[:, None]) @ Li.transpose(-1, -2)
    init_lp = -0.5 * D * math.log(2 * math.pi) - p0_tril.double().diagonal(dim1=-1, dim2=-2).log().sum(-1)[:, None] - 0.5 * (d0 * d0).sum(-1)
    log_q = elem.double().sum(-1) + init_lp - log_prob[:, None]
    return x, log_q, log_prob


def gaussian_hmm_log_prob(m0, p0_tril, chain):
    """hmm.py:271-272: init + chain.reduce(curr) then reduce(state), for init = N(m0, p0 p0').
    A [B, 2D, 2D] problem once per call (not T-sized): done with torch.linalg in float64."""
    Lam, eta, c = (t.double() for t in chain)
    D = Lam.shape[-1] // 2
    B = Lam.shape[0]
    eye = torch.eye(D, device=Lam.device, dtype=torch.float64)
    Li = torch.linalg.solve_triangular(p0_tril.double(), eye.expand(B, D, D), upper=False)
    L0 = Li.transpose(-1, -2) @ Li
    e0 = (L0 @ m0.double()[..., None])[..., 0]
    c0 = -0.5 * D * math.log(2 * math.pi) - p0_tril.double().diagonal(dim1=-1, dim2=-2).log().sum(-1) - 0.5 * (m0.double() * e0).sum(-1)
    Lam = Lam.clone()
    eta = eta.clone()
    Lam[:, :D, :D] += L0
    eta[:, :D] += e0
    Lc = torch.linalg.cholesky(Lam)
    v = torch.linalg.solve_triangular(Lc, eta[..., None], upper=False)[..., 0]
    n = 2 * D
    return c + c0 + 0.5 * n * math.log(2 * math.pi) - Lc.diagonal(dim1=-1, dim2=-2).log().sum(-1) + 0.5 * (v * v).sum(-1)
