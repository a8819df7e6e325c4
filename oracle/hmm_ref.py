"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): float64 CPU restatements of the reference's chain distributions
(funsor/pyro/hmm.py) that do NOT share the scan structure of the product path.

  discrete_hmm_log_prob   hmm.py:91-93 (normalisation), 117-138 (log_prob): a plain forward recursion
  gaussian_hmm_log_prob   hmm.py:162-274: the generative model written out (hmm.py:170-176) -> the DENSE joint Gaussian over all
                          observations, evaluated with one Cholesky
  gaussian_mrf_log_prob   hmm.py:277-383: the dense precision matrix of the whole random field; log Z(x) - log Z by
                          Gaussian integrals (hmm.py:357-377)

Pinned by tests/golden/distributions.npz, which oracle/make_golden.py generates by evaluating the same hmm.py
expressions with the unmodified reference's funsor machinery (tests/test_oracle_golden.py).
"""
import math

import numpy as np


def _lse(x, axis):
    m = np.max(x, axis=axis, keepdims=True)
    m = np.where(np.isfinite(m), m, 0.0)
    return np.squeeze(m, axis) + np.log(np.sum(np.exp(x - m), axis=axis))


def discrete_hmm_log_prob(initial_logits, transition_logits, obs_logp):
    """initial_logits [..., S]; transition_logits [..., T|1, S, S] or [S, S]; obs_logp [..., T, S] =
    observation_dist.log_prob(value[..., None]).  Returns [...] (broadcast batch)."""
    init = np.asarray(initial_logits, np.float64)
    trans = np.asarray(transition_logits, np.float64)
    obs = np.asarray(obs_logp, np.float64)
    init = init - _lse(init, -1)[..., None]
    trans = trans - _lse(trans, -1)[..., None]
    if trans.ndim == 2:
        trans = trans[None]
    T = obs.shape[-2]
    alpha = init
    for t in range(T):
        A = trans[..., t if trans.shape[-3] > 1 else 0, :, :]
        alpha = _lse(alpha[..., :, None] + A, -2) + obs[..., t, :]
    return _lse(alpha, -1)


def _mvn_logpdf(x, mean, cov):
    n = mean.shape[-1]
    L = np.linalg.cholesky(cov)
    r = np.linalg.solve(L, x - mean)
    return -0.5 * n * math.log(2 * math.pi) - np.sum(np.log(np.diag(L))) - 0.5 * r @ r


def gaussian_hmm_log_prob(m0, P0, F, q_loc, Q, H, r_loc, R, y):
    """One chain (no batch dims): F[T|1,D,D], q_loc[T|1,D], Q[T|1,D,D] (covariances), H[T|1,D,O], r_loc[T|1,O], R[T|1,O,O], y[T,O].
    z_t = z_{t-1} F_t + q_t,  x_t = z_t H_t + r_t  (row-vector convention, hmm.py:170-176)."""
    T, O = y.shape
    D = m0.shape[-1]
    pick = lambda a, t: a[t if a.shape[0] > 1 else 0]  # noqa: E731
    # stack the state as a linear function of the independent sources (z0, q_1..q_T): z_t = z0 Phi_t + sum_s q_s Psi_{s,t}
    mean_z, maps = m0.copy(), [np.eye(D)] + [None] * T  # maps[s] : source s -> current z (row-vector maps)
    covs = [P0] + [pick(Q, t) for t in range(T)]
    mean_x = np.zeros((T, O))
    G = np.zeros((T + 1, T, D, O))  # G[s, t] maps source s into x_t
    for t in range(T):
        Ft, Ht = pick(F, t), pick(H, t)
        mean_z = mean_z @ Ft + pick(q_loc, t)
        for s in range(t + 1):
            maps[s] = maps[s] @ Ft
        maps[t + 1] = np.eye(D)
        mean_x[t] = mean_z @ Ht + pick(r_loc, t)
        for s in range(t + 2):
            G[s, t] = maps[s] @ Ht
    cov = np.zeros((T * O, T * O))
    for s in range(T + 1):
        Gs = np.concatenate([G[s, t] for t in range(T)], axis=1)  # [D, T*O]
        cov += Gs.T @ covs[s] @ Gs
    for t in range(T):
        cov[t * O:(t + 1) * O, t * O:(t + 1) * O] += pick(R, t)
    return _mvn_logpdf(y.reshape(-1), mean_x.reshape(-1), cov)


def _log_gauss_integral(Lam, eta, c):
    """log of the integral over all x of exp(c - 1/2 x Lam x + x eta)"""
    n = eta.shape[-1]
    L = np.linalg.cholesky(Lam)
    v = np.linalg.solve(L, eta)
    return c + 0.5 * n * math.log(2 * math.pi) - np.sum(np.log(np.diag(L))) + 0.5 * v @ v


def _info(mean, cov):
    Lam = np.linalg.inv(cov)
    eta = Lam @ mean
    n = mean.shape[-1]
    c = -0.5 * n * math.log(2 * math.pi) - 0.5 * np.linalg.slogdet(cov)[1] - 0.5 * mean @ eta
    return Lam, eta, c


def gaussian_mrf_log_prob(m0, P0, t_loc, t_cov, o_loc, o_cov, y):
    """One chain: t_loc[T|1,2D], t_cov[T|1,2D,2D] over (old; new); o_loc[T|1,D+O], o_cov[T|1,D+O,D+O] over (state; obs); y[T,O]."""
    T, O = y.shape
    D = m0.shape[-1]
    pick = lambda a, t: a[t if a.shape[0] > 1 else 0]  # noqa: E731
    nz = (T + 1) * D
    N = nz + T * O  # variables: z_0 .. z_T, x_1 .. x_T
    Lam, eta, c = np.zeros((N, N)), np.zeros(N), 0.0

    def add(idx, L_, e_, c_):
        nonlocal c
        Lam[np.ix_(idx, idx)] += L_
        eta[idx] += e_
        c += c_

    add(np.arange(D), *_info(m0, P0))
    for t in range(T):
        zi = np.arange(t * D, (t + 2) * D)
        add(zi, *_info(pick(t_loc, t), pick(t_cov, t)))
        oi = np.concatenate([np.arange((t + 1) * D, (t + 2) * D), nz + np.arange(t * O, (t + 1) * O)])
        add(oi, *_info(pick(o_loc, t), pick(o_cov, t)))
    log_z_all = _log_gauss_integral(Lam, eta, c)  # hmm.py:370-376 (observations integrated out)
    # condition on x = y: integrate over z only (hmm.py:359-368)
    x = y.reshape(-1)
    Lzz, Lzx, Lxx = Lam[:nz, :nz], Lam[:nz, nz:], Lam[nz:, nz:]
    log_z_obs = _log_gauss_integral(Lzz, eta[:nz] - Lzx @ x, c + x @ eta[nz:] - 0.5 * x @ Lxx @ x)
    return log_z_obs - log_z_all
