"""Host-side mirror of the reference's user-facing chain distributions (funsor/pyro/hmm.py) on the CUDA path.

Same constructor arguments, shapes and broadcasting rules as the reference classes; `log_prob` does not build funsors
but hands the factored inputs straight to the kernels, which is what removes the only O(T S^2) materialisation of
DiscreteHMM.log_prob (`self._trans + obs`, pyro/hmm.py:130) and the per-step Python dispatch of the Gaussian scans:

  DiscreteHMM   pyro/hmm.py:28-138   -> fb2_hmm_forward_f32 (init, A, obs[B,T,S] fused)
  GaussianHMM   pyro/hmm.py:162-274  -> Kalman chain elements (pyro/convert.py:278-298) + Gaussian scan
  GaussianMRF   pyro/hmm.py:277-383  -> two Gaussian scans (with / without the observation) and their difference

pyro itself is not a dependency: parameters are torch.distributions objects, exactly what the reference asserts on
(pyro/hmm.py:77, 234-238, 331-333).  Tensors must live on a CUDA device in float32; there is no CPU path.
Extra methods beyond the reference class (named in the north star): DiscreteHMM.viterbi / .marginals.
"""
import math

import torch

from . import gaussian as G
from . import ops
from ._lib import check  # noqa: F401


def _check(cond, what):
    """Argument validation in the reference's style: a failed check is an AssertionError (pyro/hmm.py:75-80, 233-243, 331-336)."""
    if not cond:
        raise AssertionError(what)


def _is_mvn(d):
    return isinstance(d, torch.distributions.MultivariateNormal)


def _broadcast_shape(*shapes):
    return tuple(torch.broadcast_shapes(*[tuple(s) for s in shapes]))


def _flat(t, batch_shape, tail):
    """expand `t[..., *tail]` to batch_shape + tail and flatten the batch dims to one"""
    t = t.expand(tuple(batch_shape) + tuple(tail))
    return t.reshape((-1,) + tuple(tail))


class DiscreteHMM:
    """Hidden Markov model with discrete latent state and an arbitrary observation distribution (pyro/hmm.py:28-138).

    initial_logits     [..., S]                      broadcastable to batch_shape + (S,)
    transition_logits  [..., T|1, S, S] (old, new)   broadcastable to batch_shape + (T, S, S)
    observation_dist   batch_shape [..., T|1, S], any event_shape
    event_shape = (T|1,) + observation_dist.event_shape; the homogeneous case (T = 1) accepts data of any length.
    """

    def __init__(self, initial_logits, transition_logits, observation_dist, validate_args=None):
        _check(torch.is_tensor(initial_logits) and torch.is_tensor(transition_logits), "logits must be torch tensors")
        _check(isinstance(observation_dist, torch.distributions.Distribution), "observation_dist must be a torch distribution")
        _check(initial_logits.dim() >= 1 and transition_logits.dim() >= 2, "logits need a state dim (init) / two state dims (trans)")
        _check(len(observation_dist.batch_shape) >= 1, "observation_dist needs a state dim in its batch_shape")
        shape = _broadcast_shape(initial_logits.shape[:-1] + (1,), transition_logits.shape[:-2], observation_dist.batch_shape[:-1])
        self.batch_shape, time_shape = torch.Size(shape[:-1]), shape[-1:]
        self.event_shape = torch.Size(time_shape + tuple(observation_dist.event_shape))
        self.has_rsample = observation_dist.has_rsample
        self.state_dim = initial_logits.shape[-1]
        # Normalize (pyro/hmm.py:91-93).
        self.initial_logits = initial_logits - initial_logits.logsumexp(-1, True)
        self.transition_logits = transition_logits - transition_logits.logsumexp(-1, True)
        self.observation_dist = observation_dist
        self._validate_args = bool(validate_args)

    @property
    def event_dim(self):
        return len(self.event_shape)

    def expand(self, batch_shape):
        new = object.__new__(DiscreteHMM)
        new.__dict__.update(self.__dict__)
        new.batch_shape = torch.Size(batch_shape)
        return new

    def _factors(self, value):
        """(init[B,S], A[S,S] | [B,S,S] | [B,T,S,S], obs[B,T,S], batch shape of the result)"""
        S = self.state_dim
        obs_event_dim = len(self.observation_dist.event_shape)
        if value.dim() < 1 + obs_event_dim:
            raise ValueError("value must have shape [..., num_steps] + observation event_shape")
        T = value.shape[-1 - obs_event_dim]
        if self.event_shape[0] not in (1, T):
            raise ValueError("expected {} time steps, value has {}".format(self.event_shape[0], T))
        # observation log-likelihood per (t, state): hmm.py:127 `self._obs(value=value)`
        obs = self.observation_dist.log_prob(value.unsqueeze(-1 - obs_event_dim))  # [..., T, S]
        batch = _broadcast_shape(self.batch_shape, obs.shape[:-2], self.transition_logits.shape[:-3], self.initial_logits.shape[:-1])
        obs = _flat(obs.float(), batch, (T, S)).contiguous()
        init = _flat(self.initial_logits, batch, (S,)).contiguous()
        A = self.transition_logits
        lead, Ta = A.shape[:-3], (A.shape[-3] if A.dim() >= 3 else 1)
        if A.dim() == 2:
            pass
        elif Ta == 1 and all(d == 1 for d in lead):
            A = A.reshape(S, S)
        elif Ta == 1:
            A = _flat(A.squeeze(-3), batch, (S, S)).contiguous()
        else:
            A = _flat(A, batch, (T, S, S)).contiguous()
        return init, A.contiguous(), obs, batch

    def log_prob(self, value):
        """hmm.py:117-138: logsumexp over state paths of init + sum_t (trans + obs)."""
        init, A, obs, batch = self._factors(value)
        return ops.hmm_forward(init, A, obs, sum_op="logaddexp").reshape(batch)

    def viterbi(self, value):
        """MAP state path: (score[batch], path[batch, T+1] int64); path[..., 0] is the state before the first transition."""
        init, A, obs, batch = self._factors(value)
        best, path = ops.hmm_viterbi(init, A, obs)
        return best.reshape(batch), path.reshape(tuple(batch) + (path.shape[-1],))

    def sample_posterior(self, value, num_samples=1, uniforms=None, generator=None):
        """Forward-filter backward-sample: (paths[batch, N, T+1] int64, log p(path | value)[batch, N])."""
        init, A, obs, batch = self._factors(value)
        paths, log_q, _ = ops.hmm_posterior_sample(init, A, obs, num_samples, uniforms, generator)
        return paths.reshape(tuple(batch) + tuple(paths.shape[-2:])), log_q.reshape(tuple(batch) + (paths.shape[-2],))

    def marginals(self, value):
        """Forward-backward: (log_prob[batch], log p(z_t | value)[batch, T, S])."""
        init, A, obs, batch = self._factors(value)
        logZ, gamma, _ = ops.hmm_marginals(init, A, obs, want_xi=False)
        return logZ.reshape(batch), gamma.reshape(tuple(batch) + tuple(gamma.shape[-2:]))


# ------------------------------------------------------------------------------------------------- Gaussian
def _mvn_parts(d):
    """(loc, scale_tril) of a MultivariateNormal, or of an Independent(Normal, 1) (pyro/hmm.py:228-229)"""
    if isinstance(d, torch.distributions.MultivariateNormal):
        return d.loc, d.scale_tril
    if isinstance(d, torch.distributions.Independent) and isinstance(d.base_dist, torch.distributions.Normal) and d.reinterpreted_batch_ndims == 1:
        return d.base_dist.loc, torch.diag_embed(d.base_dist.scale)
    raise TypeError("expected MultivariateNormal or Independent(Normal, 1), got {}".format(type(d).__name__))


def _mvn_info(loc, tril):
    """N(x; loc, tril tril') as (Lam, eta, c):  log density = c - 1/2 x Lam x + x eta"""
    n = loc.shape[-1]
    Lam = torch.cholesky_inverse(tril)
    eta = (Lam @ loc[..., None])[..., 0]
    c = -0.5 * n * math.log(2 * math.pi) - tril.diagonal(dim1=-1, dim2=-2).log().sum(-1) - 0.5 * (loc * eta).sum(-1)
    return Lam, eta, c


def _info_chain(Lam, eta, c):
    """Scan over info-form chain elements; Lam may be [B,n,n] (homogeneous) or [B,T,n,n]."""
    if Lam.dim() == 4:
        return G.gaussian_sequential_sum_product(Lam, eta, c)
    return G.gaussian_chain_info_homog(Lam, eta, c)


class GaussianHMM:
    """Linear-Gaussian state space model (pyro/hmm.py:162-274):
        z_t = z_{t-1} @ transition_matrix + transition_dist noise;  x_t = z_t @ observation_matrix + observation_dist noise.
    Shapes as in the reference: matrices [..., T|1, D, D] / [..., T|1, D, O], noise batch shapes [..., T|1]."""

    has_rsample = True

    def __init__(self, initial_dist, transition_matrix, transition_dist, observation_matrix, observation_dist, validate_args=None):
        _check(_is_mvn(initial_dist) and _is_mvn(transition_dist), "initial_dist / transition_dist must be MultivariateNormal")
        _check(torch.is_tensor(transition_matrix) and torch.is_tensor(observation_matrix), "matrices must be torch tensors")
        _check(_is_mvn(observation_dist) or isinstance(observation_dist, torch.distributions.Independent),
               "observation_dist must be MultivariateNormal or Independent(Normal, 1)")
        D_, O_ = observation_matrix.shape[-2], observation_matrix.shape[-1]
        hidden_dim, obs_dim = int(D_), int(O_)
        _check(2 * obs_dim >= hidden_dim - 1, "obs_dim must be at least half of hidden_dim")
        _check(tuple(initial_dist.event_shape) == (hidden_dim,) and tuple(transition_dist.event_shape) == (hidden_dim,),
               "initial_dist / transition_dist must have event_shape (hidden_dim,)")
        _check(tuple(transition_matrix.shape[-2:]) == (hidden_dim, hidden_dim), "transition_matrix must end in (hidden_dim, hidden_dim)")
        _check(tuple(observation_dist.event_shape) == (obs_dim,), "observation_dist must have event_shape (obs_dim,)")
        shape = _broadcast_shape(initial_dist.batch_shape + (1,), transition_matrix.shape[:-2], transition_dist.batch_shape,
                                 observation_matrix.shape[:-2], observation_dist.batch_shape)
        self.batch_shape, time_shape = torch.Size(shape[:-1]), shape[-1:]
        self.event_shape = torch.Size(time_shape + (obs_dim,))
        self.hidden_dim, self.obs_dim = hidden_dim, obs_dim
        self._init = (initial_dist.loc, initial_dist.scale_tril)
        self._F, self._H = transition_matrix, observation_matrix
        self._q = _mvn_parts(transition_dist)
        self._r = _mvn_parts(observation_dist)

    def log_prob(self, value):
        D, O = self.hidden_dim, self.obs_dim
        if value.dim() < 2 or value.shape[-1] != O:
            raise ValueError("value must have shape [..., num_steps, obs_dim]")
        T = value.shape[-2]
        if self.event_shape[0] not in (1, T):
            raise ValueError("expected {} time steps, value has {}".format(self.event_shape[0], T))
        batch = _broadcast_shape(self.batch_shape, value.shape[:-2])
        y = _flat(value.float(), batch, (T, O)).contiguous()
        B = y.shape[0]
        params = {"F": (self._F, (D, D)), "q_loc": (self._q[0], (D,)), "q_tril": (self._q[1], (D, D)),
                  "H": (self._H, (D, O)), "r_loc": (self._r[0], (O,)), "r_tril": (self._r[1], (O, O))}
        homog = all(p.dim() <= len(tail) or p.shape[-1 - len(tail)] == 1 for p, tail in params.values())
        if homog:
            flat = {}
            for k, (p, tail) in params.items():
                p = p if p.dim() <= len(tail) else p.squeeze(-1 - len(tail))  # drop the size-1 time dim
                flat[k] = _flat(p, batch, tail).contiguous()
            chain = G.kalman_chain(flat["F"], flat["q_loc"], flat["q_tril"], flat["H"], flat["r_loc"], flat["r_tril"], y)
        else:
            flat = {}
            for k, (p, tail) in params.items():
                p = p if p.dim() > len(tail) else p.unsqueeze(-1 - len(tail))  # add a time dim
                tp = p.shape[-1 - len(tail)]
                flat[k] = _flat(p, batch, (tp,) + tail)
            w, P, s = G.kalman_elements_timevarying(flat["F"], flat["q_loc"], flat["q_tril"], flat["H"], flat["r_loc"], flat["r_tril"], y)
            chain = G.gaussian_sequential_sum_product(*G.sqrt_to_info(w, P, s))
        m0 = _flat(self._init[0], batch, (D,))
        p0 = _flat(self._init[1], batch, (D, D))
        return G.gaussian_hmm_log_prob(m0, p0, chain).float().reshape(batch)


class GaussianMRF:
    """Temporal Markov random field with joint Gaussian factors (pyro/hmm.py:277-383):
    transition_dist over (old; new) states [..., T|1] x (2D), observation_dist over (state; obs) [..., T|1] x (D + O).
    log_prob = log Z(with observations) - log Z(observations marginalised out)  (hmm.py:357-377)."""

    has_rsample = True

    def __init__(self, initial_dist, transition_dist, observation_dist, validate_args=None):
        _check(_is_mvn(initial_dist) and _is_mvn(transition_dist) and _is_mvn(observation_dist), "all three factors must be MultivariateNormal")
        hidden_dim = int(initial_dist.event_shape[-1])
        _check(int(transition_dist.event_shape[-1]) == 2 * hidden_dim, "transition_dist must be over (old; new) states: event size 2 * hidden_dim")
        obs_dim = int(observation_dist.event_shape[-1]) - hidden_dim
        _check(obs_dim >= 1, "observation_dist must be over (state; observation)")
        shape = _broadcast_shape(initial_dist.batch_shape + (1,), transition_dist.batch_shape, observation_dist.batch_shape)
        self.batch_shape, time_shape = torch.Size(shape[:-1]), shape[-1:]
        self.event_shape = torch.Size(time_shape + (obs_dim,))
        self.hidden_dim, self.obs_dim = hidden_dim, obs_dim
        self._init = (initial_dist.loc, initial_dist.scale_tril)
        self._trans = (transition_dist.loc, transition_dist.scale_tril)
        self._obs = (observation_dist.loc, observation_dist.scale_tril)

    def log_prob(self, value):
        D, O = self.hidden_dim, self.obs_dim
        if value.dim() < 2 or value.shape[-1] != O:
            raise ValueError("value must have shape [..., num_steps, obs_dim]")
        T = value.shape[-2]
        if self.event_shape[0] not in (1, T):
            raise ValueError("expected {} time steps, value has {}".format(self.event_shape[0], T))
        batch = _broadcast_shape(self.batch_shape, value.shape[:-2])
        v = _flat(value.float(), batch, (T, O))
        B = v.shape[0]

        def timed(p, tail):  # -> [B, T|1, *tail]
            p = p if p.dim() > len(tail) else p.unsqueeze(-1 - len(tail))
            return _flat(p, batch, (p.shape[-1 - len(tail)],) + tail)

        t_loc, t_tril = timed(self._trans[0], (2 * D,)), timed(self._trans[1], (2 * D, 2 * D))
        o_loc, o_tril = timed(self._obs[0], (D + O,)), timed(self._obs[1], (D + O, D + O))
        Lt, et, ct = _mvn_info(t_loc, t_tril)  # [B,Tt,2D,2D], [B,Tt,2D], [B,Tt]
        Lo, eo, co = _mvn_info(o_loc, o_tril)  # over (state; obs)
        # substitute the observed value into the joint factor (gaussian.py:723-744): a factor on the new state only
        Loss, Losv, Lovv = Lo[..., :D, :D], Lo[..., :D, D:], Lo[..., D:, D:]
        e_s = eo[..., :D] - (Losv @ v[..., None])[..., 0]  # [B,T,D]
        c_s = co + (v * eo[..., D:]).sum(-1) - 0.5 * (v * (Lovv @ v[..., None])[..., 0]).sum(-1)  # [B,T]
        # marginalise the observation out of the joint factor instead (hmm.py:370): N(state; loc_s, tril_ss)
        Lm, em, cm = _mvn_info(o_loc[..., :D], o_tril[..., :D, :D])

        def chain(L_add, e_add, c_add):
            homog = Lt.shape[1] == 1 and L_add.shape[1] == 1
            Tl = 1 if homog else T
            Lam = Lt.expand(B, Tl, 2 * D, 2 * D).clone()
            Lam[..., D:, D:] += L_add.expand(B, Tl, D, D)
            eta = et.expand(B, T, 2 * D).clone()
            eta[..., D:] += e_add.expand(B, T, D)
            c = (ct + c_add).expand(B, T).contiguous()
            Lam = Lam[:, 0].contiguous() if homog else Lam.contiguous()
            out = _info_chain(Lam, eta.contiguous(), c)
            return G.gaussian_hmm_log_prob(_flat(self._init[0], batch, (D,)), _flat(self._init[1], batch, (D, D)), out)

        logp_oh = chain(Loss, e_s, c_s)
        logp_h = chain(Lm, em, cm)
        return (logp_oh - logp_h).float().reshape(batch)
