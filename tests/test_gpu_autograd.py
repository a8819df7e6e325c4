"""GPU: backward passes of the differentiable entry points (funsor_b200/autograd.py) against torch autograd through a float64
evaluation of the same expression -- the way the reference gets its gradients (examples/discrete_hmm.py:76-78 call
loss.backward() through torch.logsumexp / einsum; the max shift is detached, einsum/numpy_log.py:27)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def fb():
    import funsor_b200

    return funsor_b200


def _ref_matmul(sname, x, y):
    z = x.unsqueeze(-1) + y.unsqueeze(-3)
    return torch.logsumexp(z, -2) if sname == "logaddexp" else z.amax(-2)


def _ref_forward(sname, init, A, obs):
    B, T, S = obs.shape
    v = init.expand(B, S)
    for t in range(T):
        At = A if A.dim() == 2 else (A if A.dim() == 3 else A[:, t])
        z = v[:, :, None] + At + obs[:, t, None, :]
        v = torch.logsumexp(z, 1) if sname == "logaddexp" else z.amax(1)
    return torch.logsumexp(v, -1) if sname == "logaddexp" else v.amax(-1)


def _check(got, want, name, rtol=2e-5):
    a, b = got.detach().cpu().double().numpy(), want.detach().numpy()
    assert a.shape == b.shape, (name, a.shape, b.shape)
    assert np.abs(a - b).max() <= rtol * max(1.0, np.abs(b).max()), (name, np.abs(a - b).max(), np.abs(b).max())


@pytest.mark.parametrize("sname", ["logaddexp", "max"])
@pytest.mark.parametrize("xs,ys", [((3, 5, 40, 70), (3, 5, 70, 33)), ((2, 1, 130, 64), (1, 4, 64, 129)), ((7, 9), (9, 4))])
def test_semiring_matmul_backward(fb, sname, xs, ys):
    g = torch.Generator().manual_seed(1)
    x0, y0 = torch.randn(*xs, generator=g), torch.randn(*ys, generator=g)
    cot = torch.randn(torch.broadcast_shapes(xs[:-2], ys[:-2]) + (xs[-2], ys[-1]), generator=g)  # signed cotangent
    dead = sname == "logaddexp"
    if dead:  # a dead row (all -inf): zero gradient, no NaN.  torch's own logsumexp backward yields NaN there, so the float64
        cot[..., 0, :] = 0.0  # reference sees a large negative finite value instead and the row is kept out of the loss
    xr, yr = x0.double(), y0.double()
    if dead:
        xr[..., 0, :] = -1e5
        x0[..., 0, :] = -float("inf")
    xr, yr = xr.requires_grad_(), yr.requires_grad_()
    out_r = _ref_matmul(sname, xr, yr)
    (out_r * cot.double()).sum().backward()
    x, y = x0.to(DEV).requires_grad_(), y0.to(DEV).requires_grad_()
    out = fb.semiring_matmul(sname, "add", x, y)
    assert out.grad_fn is not None
    live = torch.isfinite(out)
    if dead:
        assert not bool(live[..., 0, :].any()) and bool(live[..., 1:, :].all())
    (torch.where(live, out, torch.zeros_like(out)) * cot.to(DEV)).sum().backward()
    _check(out[..., 1:, :], out_r[..., 1:, :], "out")
    _check(x.grad, xr.grad, "grad x")
    _check(y.grad, yr.grad, "grad y")
    assert bool(torch.isfinite(x.grad).all()) and bool(torch.isfinite(y.grad).all())


@pytest.mark.parametrize("sname", ["logaddexp", "max"])
@pytest.mark.parametrize("B,T,S", [(2, 9, 16), (1, 12, 70), (3, 1, 5), (2, 33, 128)])
def test_chain_product_backward(fb, sname, B, T, S):
    g = torch.Generator().manual_seed(2)
    t0 = torch.randn(B, T, S, S, generator=g)
    cot = torch.randn(B, S, S, generator=g)
    tr = t0.double().requires_grad_()
    out_r = tr[:, 0]
    for t in range(1, T):
        out_r = _ref_matmul(sname, out_r, tr[:, t])
    (out_r * cot.double()).sum().backward()
    tg = t0.to(DEV).requires_grad_()
    out = fb.sequential_sum_product(sname, "add", tg)
    (out * cot.to(DEV)).sum().backward()
    _check(out, out_r, "out", rtol=1e-5)
    if sname == "max":  # association order differs from the serial fold: compare supports loosely through the value only
        assert tg.grad is not None and tg.grad.shape == t0.shape
        assert abs(float(tg.grad.sum()) - float(tr.grad.sum())) <= 1e-3 * (1 + abs(float(tr.grad.sum())))
    else:
        _check(tg.grad, tr.grad, "grad trans", rtol=5e-5)


@pytest.mark.parametrize("sname", ["logaddexp", "max"])
@pytest.mark.parametrize("B,T,S,amode", [(2, 17, 16, "shared"), (3, 40, 70, "shared"), (2, 25, 33, "batched"), (2, 12, 6, "timevarying"),
                                         (1, 300, 257, "shared"), (4, 64, 128, "shared"), (2, 1, 5, "shared")])
def test_hmm_forward_backward_pass(fb, sname, B, T, S, amode):
    """d logZ / d (init, A, obs) = posterior marginals (one forward-backward pass on the kernels) vs torch autograd in float64"""
    g = torch.Generator().manual_seed(3)
    shape = {"shared": (S, S), "batched": (B, S, S), "timevarying": (B, T, S, S)}[amode]
    A0, obs0, init0 = torch.randn(*shape, generator=g), torch.randn(B, T, S, generator=g), torch.randn(B, S, generator=g)
    cot = torch.randn(B, generator=g)

    def run(put, fwd):
        A, obs, init = (put(t).requires_grad_() for t in (A0, obs0, init0))
        z = fwd(init.log_softmax(-1), A.log_softmax(-1), obs)
        (z * put(cot)).sum().backward()
        return z, A.grad, obs.grad, init.grad

    want = run(lambda t: t.double(), lambda i, a, o: _ref_forward(sname, i, a, o))
    got = run(lambda t: t.to(DEV), lambda i, a, o: fb.hmm_forward(i, a, o, sum_op=sname))
    for name, a, b in zip(("logZ", "grad A", "grad obs", "grad init"), got, want):
        _check(a, b, name, rtol=1e-5 if name == "logZ" else 1e-4)


def test_hmm_forward_backward_with_forbidden_states(fb):
    """-inf entries (forbidden transitions / impossible observations): finite gradients, zero on the forbidden entries.  The
    float64 reference sees -1e5 in their place (torch's logsumexp backward is NaN-prone at -inf; exp(-1e5) is exactly 0)."""
    g = torch.Generator().manual_seed(4)
    B, T, S = 2, 20, 48
    A0 = torch.randn(S, S, generator=g)
    A0[0, 1] = A0[5, :3] = -float("inf")
    obs0 = torch.randn(B, T, S, generator=g)
    obs0[0, 3, 2] = obs0[1, 0, 7] = -float("inf")
    init0 = torch.randn(B, S, generator=g)
    init0[:, 4] = -float("inf")

    def run(put, fwd):
        A, obs, init = (put(t).requires_grad_() for t in (A0, obs0, init0))
        z = fwd(init, A, obs)
        z.sum().backward()
        return z, A.grad, obs.grad, init.grad

    want = run(lambda t: torch.nan_to_num(t.double(), neginf=-1e5), lambda i, a, o: _ref_forward("logaddexp", i, a, o))
    got = run(lambda t: t.to(DEV), lambda i, a, o: fb.hmm_forward(i, a, o))
    for name, a, b in zip(("logZ", "grad A", "grad obs", "grad init"), got, want):
        assert bool(torch.isfinite(a).all()), name
        _check(a, b, name, rtol=1e-4)
    assert float(got[1][0, 1]) == 0.0 and float(got[2][0, 3, 2]) == 0.0 and float(got[3][0, 4]) == 0.0


def test_discrete_hmm_class_trains(fb):
    """the user-facing mirror of funsor.pyro.hmm.DiscreteHMM: a few Adam steps on logits must increase the likelihood
    (examples/discrete_hmm.py:60-80 training loop shape)"""
    from funsor_b200.hmm import DiscreteHMM

    torch.manual_seed(5)
    S, T, B = 16, 100, 8
    init_logits = torch.zeros(S, device=DEV, requires_grad=True)
    trans_logits = torch.randn(S, S, device=DEV, requires_grad=True)
    loc = torch.randn(S, device=DEV, requires_grad=True)
    data = torch.randn(B, T, device=DEV)
    opt = torch.optim.Adam([init_logits, trans_logits, loc], lr=0.1)
    losses = []
    for _ in range(8):
        opt.zero_grad()
        d = DiscreteHMM(init_logits, trans_logits, torch.distributions.Normal(loc, torch.ones(S, device=DEV)))
        loss = -d.log_prob(data).sum()
        loss.backward()
        assert all(p.grad is not None and bool(torch.isfinite(p.grad).all()) for p in (init_logits, trans_logits, loc))
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < losses[0], losses
