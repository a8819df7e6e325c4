"""GPU: the UNMODIFIED reference (funsor 0.4.8, imported from baseline/_ref on the GPU box) evaluating its own expressions on
CUDA float32 tensors with funsor_b200.funsor_plugin registered.  Every case asserts that the rule took its kernel branch
(`stats()["*_fast"] > 0`) and compares with the same expression evaluated by the stock rules on CPU float64 tensors in the same
process (CPU operands make every rule delegate), and with the committed golden fixtures.

Covers SURVEY 8 rows a4 (MarkovProduct rule), a5 (planner -> optimizer -> Contraction), a6/a7 (Contraction rules), a8 (einsum front
door), a11 (AdjointTape over a scan), a12 (torch op registrations on CUDA), (b) the reference-side binding, and autograd through the
rules (examples/discrete_hmm.py:76-78).
"""
from collections import OrderedDict

import numpy as np
import pytest
import torch

from conftest import golden_cases, load_golden, rel_close
from oracle import ref_boot

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_boot.available(), reason="reference package not present (baseline/_ref)")]

DEV = "cuda"


@pytest.fixture(scope="module")
def funsor():
    f = ref_boot.boot("float32")
    from funsor_b200 import funsor_plugin

    funsor_plugin.set_hooks()
    funsor_plugin.register()
    return f


@pytest.fixture
def plugin(funsor):
    from funsor_b200 import funsor_plugin

    funsor_plugin.stats(reset=True)
    return funsor_plugin


def _ops(funsor, sname):
    import funsor.ops as ops

    return ops.logaddexp if sname == "logaddexp" else ops.max


def _close(got, want, rtol=1e-5, atol=1.0):
    """funsor Tensor (CUDA fp32) against funsor Tensor (CPU fp64): same inputs in the same order, reference comparator"""
    assert list(got.inputs.items()) == list(want.inputs.items()), (got.inputs, want.inputs)
    assert got.data.is_cuda and got.data.dtype == torch.float32
    rel_close(got.data.detach().cpu().numpy(), want.data.detach().numpy(), rtol=rtol, atol=atol)


@pytest.mark.parametrize("sname", ["logaddexp", "max"])
@pytest.mark.parametrize("default_device", [False, True])
def test_reference_expressions_run_on_the_kernels(funsor, plugin, sname, default_device):
    """default_device=True is the reference's own way of running on a GPU (examples/mixed_hmm/experiment.py:72-73 sets the default
    tensor type to CUDA): every internal allocation then lands on the GPU.  With explicit CUDA operands only (False) the reference's
    AdjointTape seeds its backward pass from a default-device tensor (tensor.py:635 via adjoint.py:74) and cannot run; everything else
    works thanks to the device-correct op registrations of the plugin (a12)."""
    import contextlib

    import funsor.ops as ops
    from funsor.adjoint import AdjointTape
    from funsor.cnf import Contraction
    from funsor.domains import Bint
    from funsor.einsum import einsum
    from funsor.interpretations import lazy
    from funsor.optimizer import apply_optimizer
    from funsor.sum_product import MarkovProduct, mixed_sequential_sum_product, sequential_sum_product, sum_product
    from funsor.tensor import Tensor
    from funsor.terms import Variable

    op = _ops(funsor, sname)
    g = torch.Generator().manual_seed(7)
    B, T, S = 3, 37, 48
    d_trans, d_init = torch.randn(B, T, S, S, generator=g), torch.randn(B, S, generator=g)
    d_x, d_y, d_z = torch.randn(B, 40, 70, generator=g), torch.randn(70, 33, B, generator=g), torch.randn(B, 33, generator=g)
    tv = Variable("time", Bint[T])

    def run(build):
        """build(put) with put = to CPU float64 (stock rules) and put = to CUDA float32 (kernel branch)"""
        want = build(lambda t: t.double())
        plugin.stats(reset=True)
        with (torch.device(DEV) if default_device else contextlib.nullcontext()):
            got = build(lambda t: t.to(DEV))
        return want, got, plugin.stats()

    def trans_of(put):
        return Tensor(put(d_trans), OrderedDict(b=Bint[B], time=Bint[T], prev=Bint[S], curr=Bint[S]))

    def contraction(put):
        x = Tensor(put(d_x), OrderedDict(b=Bint[B], i=Bint[40], k=Bint[70]))
        y = Tensor(put(d_y), OrderedDict(k=Bint[70], j=Bint[33], b=Bint[B]))
        return Contraction(op, ops.add, frozenset({Variable("k", Bint[70])}), x, y)

    def markov(put):
        return MarkovProduct(op, ops.add, trans_of(put), tv, {"prev": "curr"})

    def sequential(put):
        return sequential_sum_product(op, ops.add, trans_of(put), tv, {"prev": "curr"})

    def mixed(put):
        return mixed_sequential_sum_product(op, ops.add, trans_of(put), tv, {"prev": "curr"}, num_segments=5)

    def hmm_expression(put):  # pyro/hmm.py:128-137
        init = Tensor(put(d_init), OrderedDict(b=Bint[B], prev=Bint[S]))
        r = sequential_sum_product(op, ops.add, trans_of(put), tv, {"prev": "curr"})
        return (init + r.reduce(op, "curr")).reduce(op, "prev")

    def front_door(put):  # einsum/__init__.py:13-28: the backend NAME selects the semiring
        backend = "pyro.ops.einsum.torch_log" if sname == "logaddexp" else "pyro.ops.einsum.torch_map"
        xs = Tensor(put(d_x[0]), OrderedDict(a=Bint[40], b=Bint[70]))
        ys = Tensor(put(d_y[..., 0]), OrderedDict(b=Bint[70], c=Bint[33]))
        return einsum("ab,bc->ac", xs, ys, backend=backend)

    def planner(put):  # sum_product builds lazily, the optimizer emits binary Contractions (einsum/__init__.py:126-128)
        x = Tensor(put(d_x), OrderedDict(b=Bint[B], i=Bint[40], k=Bint[70]))
        y = Tensor(put(d_y), OrderedDict(k=Bint[70], j=Bint[33], b=Bint[B]))
        z = Tensor(put(d_z), OrderedDict(b=Bint[B], j=Bint[33]))
        with lazy:
            expr = sum_product(op, ops.add, [x, y, z], frozenset({"k", "j", "b"}), frozenset({"b"}))
        return apply_optimizer(expr)

    def adjoint(put):  # adjoint.py:72-131 over the scan; test_adjoint.py:187-261 body
        trans = trans_of(put)
        with AdjointTape() as tape:
            z = sequential_sum_product(op, ops.add, trans, tv, {"prev": "curr"})
            z = z.reduce(op, frozenset({"prev", "curr"}))
        return tape.adjoint(op, ops.add, z, (trans,), batch_vars=frozenset({Variable("b", Bint[B])}))[trans]

    cases = {"contraction": (contraction, "contraction_fast"), "markov": (markov, "markov_fast"), "sequential": (sequential, "contraction_fast"),
             "mixed": (mixed, "contraction_fast"), "hmm": (hmm_expression, "contraction_fast"), "einsum": (front_door, "contraction_fast"),
             "sum_product": (planner, "contraction_fast"), "adjoint": (adjoint, "contraction_fast")}
    for name, (build, counter) in cases.items():
        if name == "adjoint" and not default_device:
            continue
        want, got, st = run(build)
        assert st[counter] > 0, (name, st)
        assert isinstance(got, Tensor), (name, type(got))
        try:
            _close(got, want)
        except AssertionError as e:
            raise AssertionError("{}: {}".format(name, e))
        if sname == "max" and name in ("contraction", "markov", "sequential"):  # max-plus: bit-exact against an fp32 evaluation
            want32 = build(lambda t: t.clone())
            assert np.array_equal(got.data.cpu().numpy(), want32.data.numpy()), name


@pytest.mark.parametrize("sname", ["logaddexp", "max"])
def test_reference_scan_on_golden_fixtures(funsor, plugin, sname):
    """funsor's own sequential_sum_product / MarkovProduct on CUDA against tests/golden/chain_product.npz (written by the same
    reference on CPU float64, oracle/make_golden.py)."""
    import funsor.ops as ops
    from funsor.domains import Bint
    from funsor.sum_product import MarkovProduct, sequential_sum_product
    from funsor.tensor import Tensor
    from funsor.terms import Variable

    op = _ops(funsor, sname)
    key = "log" if sname == "logaddexp" else "max"
    n = 0
    for case, d in golden_cases(load_golden("chain_product")).items():
        trans = torch.from_numpy(d["trans"]).float()
        B, T, S, _ = trans.shape
        f = Tensor(trans.to(DEV), OrderedDict(b=Bint[B], time=Bint[T], prev=Bint[S], curr=Bint[S]))
        tv = Variable("time", Bint[T])
        for fn, suffix in ((sequential_sum_product, "_seq"), (MarkovProduct, "_markov")):
            want = d.get(key + suffix)
            if want is None:
                continue
            got = fn(op, ops.add, f, tv, {"prev": "curr"})
            got = got.align(("b", "prev", "curr")).data.cpu().numpy()
            rel_close(got, want, rtol=1e-5, atol=1.0)
            n += 1
    assert n > 0
    st = plugin.stats()
    assert st["contraction_fast"] > 0 and st["markov_fast"] > 0, st


def test_c1_through_the_reference_api(funsor, plugin):
    """BASELINE config 1 (examples/discrete_hmm.py-style forward log-likelihood, T=100, hidden_dim=16) exactly as SURVEY 8(d) C1,
    evaluated by funsor on CUDA through the rules, against the fixture the reference wrote on CPU."""
    import funsor.ops as ops
    from funsor.domains import Bint
    from funsor.sum_product import sequential_sum_product
    from funsor.tensor import Tensor
    from funsor.terms import Variable

    d = golden_cases(load_golden("hmm_logprob"))["c1"]
    trans, init = torch.from_numpy(d["trans"]).to(DEV), torch.from_numpy(d["init"]).to(DEV)
    f = Tensor(trans, OrderedDict([("b", Bint[1]), ("time", Bint[100]), ("state", Bint[16]), ("state(time=1)", Bint[16])]))
    r = sequential_sum_product(ops.logaddexp, ops.add, f, Variable("time", Bint[100]), {"state": "state(time=1)"})
    r = Tensor(init, OrderedDict([("b", Bint[1]), ("state", Bint[16])])) + r.reduce(ops.logaddexp, "state(time=1)")
    r = r.reduce(ops.logaddexp, "state")
    assert plugin.stats()["contraction_fast"] >= 7  # one per tree level
    rel_close(r.data.cpu().numpy(), d["logZ_fp64"], rtol=1e-5, atol=1.0)


# ------------------------------------------------------------------------------------------------ autograd through the rules
@pytest.mark.parametrize("sname", ["logaddexp", "max"])
@pytest.mark.parametrize("via", ["sequential", "markov"])
def test_gradients_through_the_rules_match_the_reference(funsor, plugin, sname, via):
    """loss.backward() through funsor expressions (examples/discrete_hmm.py:76-78): gradients of the plugin path on CUDA against
    torch autograd through the reference's own CPU float64 evaluation."""
    import funsor.ops as ops
    from funsor.domains import Bint
    from funsor.sum_product import MarkovProduct, sequential_sum_product
    from funsor.tensor import Tensor
    from funsor.terms import Variable

    op = _ops(funsor, sname)
    g = torch.Generator().manual_seed(11)
    B, T, S = 2, 21, 40
    d_A, d_obs, d_init = torch.randn(S, S, generator=g), torch.randn(B, T, S, generator=g), torch.randn(B, S, generator=g)
    tv = Variable("time", Bint[T])

    def loss(put):
        A, obs, init = (put(t).requires_grad_() for t in (d_A, d_obs, d_init))
        fA = Tensor(A.log_softmax(-1), OrderedDict([("state", Bint[S]), ("state(time=1)", Bint[S])]))
        fobs = Tensor(obs, OrderedDict([("b", Bint[B]), ("time", Bint[T]), ("state(time=1)", Bint[S])]))
        finit = Tensor(init.log_softmax(-1), OrderedDict([("b", Bint[B]), ("state", Bint[S])]))
        trans = fA + fobs  # hmm.py:130
        if via == "sequential":
            r = sequential_sum_product(op, ops.add, trans, tv, {"state": "state(time=1)"})
        else:
            r = MarkovProduct(op, ops.add, trans, tv, {"state": "state(time=1)"})
        r = (finit + r.reduce(op, "state(time=1)")).reduce(op, "state")
        weights = put(torch.tensor([1.0, -0.5]))  # a negative cotangent exercises the signed branch of the log backward
        total = (r.data * weights).sum()
        total.backward()
        return total.detach(), [A.grad, obs.grad, init.grad]

    want, wgrads = loss(lambda t: t.double())
    plugin.stats(reset=True)
    got, ggrads = loss(lambda t: t.to(DEV))
    st = plugin.stats()
    assert st["contraction_fast" if via == "sequential" else "markov_fast"] > 0, st
    assert abs(float(got) - float(want)) <= 1e-5 * (1 + abs(float(want)))
    for name, a, b in zip(("A", "obs", "init"), ggrads, wgrads):
        assert a is not None and a.is_cuda, name
        a, b = a.cpu().double().numpy(), b.numpy()
        if sname == "max":  # one-hot along the arg-max path: identical support
            assert np.array_equal(a != 0, b != 0), name
        assert np.abs(a - b).max() <= 2e-5 * max(1.0, np.abs(b).max()), (name, np.abs(a - b).max())


def test_gaussian_rules_delegate_when_gradients_are_needed(funsor, plugin):
    import funsor.ops as ops
    from funsor.domains import Bint, Reals
    from funsor.gaussian import Gaussian
    from funsor.sum_product import sequential_sum_product
    from funsor.terms import Variable
    from funsor.testing import random_gaussian

    torch.manual_seed(12)
    proto = random_gaussian(OrderedDict(time=Bint[6], prev=Reals[2], curr=Reals[2]))
    w = proto.white_vec.to(DEV).requires_grad_()
    g = Gaussian(white_vec=w, prec_sqrt=proto.prec_sqrt.to(DEV), inputs=proto.inputs)
    r = sequential_sum_product(ops.logaddexp, ops.add, g, Variable("time", Bint[6]), {"prev": "curr"})
    st = plugin.stats()
    assert st["gaussian_contraction_fast"] == 0 and st["gaussian_contraction_delegated"] > 0, st
    z = r.reduce(ops.logaddexp, frozenset({"prev", "curr"}))
    z.data.sum().backward()
    assert w.grad is not None and bool(torch.isfinite(w.grad).all())


# ------------------------------------------------------------------------------------------------ Gaussian chain elements
def _info_of(funsor, result, names, D):
    from funsor.cnf import Contraction
    from funsor.domains import Reals
    from funsor.gaussian import Gaussian
    from funsor.tensor import Tensor

    terms = result.terms if isinstance(result, Contraction) else (result,)
    (g,) = [t for t in terms if isinstance(t, Gaussian)]
    reals = [k for k, d in g.inputs.items() if d.dtype == "real"]
    assert sorted(reals) == sorted(names)
    perm = torch.cat([torch.arange(D) + D * reals.index(n) for n in names]).to(g.white_vec.device)
    zero = {n: funsor.to_funsor(torch.zeros(D, dtype=g.white_vec.dtype, device=g.white_vec.device), Reals[D]) for n in names}
    c = result(**zero)
    assert isinstance(c, Tensor)
    bnames = [k for k, d in g.inputs.items() if d.dtype != "real"]
    Lam = g._precision[..., perm, :][..., :, perm]
    eta = g._info_vec[..., perm]
    return Lam, eta, c.align(tuple(bnames)).data


def _same_function(funsor, got, want, names, D, rtol):
    assert list(got.inputs.items()) == list(want.inputs.items()), (got.inputs, want.inputs)
    assert [type(t) for t in got.terms] == [type(t) for t in want.terms]
    for label, a, b in zip(("Lam", "eta", "c"), _info_of(funsor, got, names, D), _info_of(funsor, want, names, D)):
        a, b = a.detach().cpu().double(), b.detach().cpu().double()
        scale = max(1.0, float(b.abs().max()))
        assert float((a - b).abs().max()) <= rtol * scale, (label, float((a - b).abs().max()), scale)


@pytest.mark.parametrize("T,D", [(2, 1), (5, 2), (8, 3), (11, 2), (64, 4), (37, 8)])
@pytest.mark.parametrize("via", ["sequential", "markov"])
def test_gaussian_scan_inside_reference(funsor, plugin, T, D, via):
    """sum_product.py:652-701 over Gaussian chain elements, called directly (pyro/hmm.py:269: every level reaches the
    Contraction[.., Gaussian|GaussianMixture x2] rule, cnf.py:385-389) and as a MarkovProduct (whole-scan kernel)."""
    import funsor.ops as ops
    from funsor.domains import Bint, Reals
    from funsor.gaussian import Gaussian
    from funsor.sum_product import MarkovProduct, sequential_sum_product
    from funsor.terms import Variable
    from funsor.testing import random_gaussian

    torch.manual_seed(8)
    torch.set_default_dtype(torch.float64)
    try:
        proto = random_gaussian(OrderedDict(b=Bint[2], time=Bint[T], prev=Reals[D], curr=Reals[D]))
    finally:
        torch.set_default_dtype(torch.float32)
    fn = sequential_sum_product if via == "sequential" else MarkovProduct

    def build(put):
        g = Gaussian(white_vec=put(proto.white_vec), prec_sqrt=put(proto.prec_sqrt), inputs=proto.inputs)
        return fn(ops.logaddexp, ops.add, g, Variable("time", Bint[T]), {"prev": "curr"})

    torch.set_default_dtype(torch.float64)
    try:
        want = build(lambda t: t.clone())
    finally:
        torch.set_default_dtype(torch.float32)
    plugin.stats(reset=True)
    got = build(lambda t: t.float().to(DEV))
    st = plugin.stats()
    if via == "sequential":
        assert st["gaussian_contraction_fast"] > 0 and st["gaussian_contraction_delegated"] == 0, st
    else:
        assert st["markov_fast"] > 0, st
    _same_function(funsor, got, want, ["prev", "curr"], D, 2e-4)


def test_gaussian_hmm_expression_inside_reference(funsor, plugin):
    """pyro/hmm.py:258-272 built with the reference's own converters (pyro/convert.py:210-298) on CUDA: rank-deficient Kalman
    elements (rank D + O < 2D) with a Tensor part, scanned level by level through the Gaussian contraction rule."""
    import importlib.util
    import os

    import funsor.ops as ops
    from funsor.domains import Bint
    from funsor.sum_product import sequential_sum_product
    from funsor.tensor import Tensor
    from funsor.terms import Variable

    spec = importlib.util.spec_from_file_location("_ref_convert", os.path.join(ref_boot.REFERENCE_ROOT, "funsor/pyro/convert.py"))
    conv = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(conv)
    MVN = torch.distributions.MultivariateNormal
    torch.manual_seed(9)
    B, T, D, O = 2, 33, 6, 2
    F = (0.9 * torch.linalg.qr(torch.randn(B, D, D))[0]).double()
    H = torch.randn(B, D, O).double()
    y = torch.randn(B, T, O).double()
    ex = lambda t: t.unsqueeze(1).expand((B, T) + t.shape[1:])  # noqa: E731

    def build(put):
        eyeD, eyeO = put(torch.eye(D).double()), put(torch.eye(O).double())
        zD, zO = put(torch.zeros(B, T, D).double()), put(torch.zeros(B, T, O).double())
        trans = conv.matrix_and_mvn_to_funsor(put(ex(F).contiguous()), MVN(zD, scale_tril=eyeD.expand(B, T, D, D)), ("time",), "state", "state(time=1)")
        obs = conv.matrix_and_mvn_to_funsor(put(ex(H).contiguous()), MVN(zO, scale_tril=eyeO.expand(B, T, O, O)), ("time",), "state(time=1)", "value")
        value = Tensor(put(y), OrderedDict([("_pyro_dim_-1", Bint[B]), ("time", Bint[T])]))
        elem = trans + obs(value=value)
        return sequential_sum_product(ops.logaddexp, ops.add, elem, Variable("time", Bint[T]), {"state": "state(time=1)"})

    torch.set_default_dtype(torch.float64)
    try:
        want = build(lambda t: t.clone())
    finally:
        torch.set_default_dtype(torch.float32)
    plugin.stats(reset=True)
    got = build(lambda t: t.float().to(DEV))
    st = plugin.stats()
    assert st["gaussian_contraction_fast"] > 0, st
    _same_function(funsor, got, want, ["state", "state(time=1)"], D, 2e-4)


def test_info_to_sqrt_kernel_round_trip():
    """fb2_gaussian_info_to_sqrt_f32: P P' = Lam, P w = eta, s - |w|^2/2 = c, for full-rank and rank-deficient precisions"""
    import funsor_b200.gaussian as fg

    torch.manual_seed(4)
    for n, r in [(2, 2), (6, 2), (16, 16), (16, 9), (64, 64), (64, 48), (33, 20)]:
        A = torch.randn(50, n, r, dtype=torch.float64)
        Lam = A @ A.transpose(-1, -2)
        eta = (A @ torch.randn(50, r, 1, dtype=torch.float64))[..., 0]  # in the range of Lam
        if n == r:  # full rank: a square Gaussian factor is nearly singular (condition ~ n^2), which tests the matrix, not the kernel
            Lam = Lam + n * torch.eye(n, dtype=torch.float64)
        c = torch.randn(50, dtype=torch.float64)
        w, P, s, rank = fg.info_to_sqrt(Lam.float().to(DEV), eta.float().to(DEV), c.float().to(DEV), return_rank=True)
        w, P, s = w.cpu().double(), P.cpu().double(), s.cpu().double()
        scale = float(Lam.abs().max())
        assert float((P @ P.transpose(-1, -2) - Lam).abs().max()) <= 3e-6 * scale, (n, r)
        assert float(((P @ w[..., None])[..., 0] - eta).abs().max()) <= 1e-4 * float(eta.abs().max()), (n, r)
        assert float((s - 0.5 * (w * w).sum(-1) - c).abs().max()) <= 1e-5 * float(1 + (0.5 * (w * w).sum(-1)).abs().max()), (n, r)
        assert bool((rank.cpu() == r).all()), (n, r, rank.unique())


# ------------------------------------------------------------------------------------------------ a12: torch op registrations on CUDA
def test_torch_backend_ops_allocate_on_the_operand_device(funsor):
    """funsor/torch/ops.py registrations used on the path (new_arange 300-306, new_eye 309-311, new_zeros/new_full 314-321, einsum,
    logsumexp, cholesky, triangular_solve, cat/stack) must follow the device of their prototype tensor, not the default device."""
    import funsor.ops as ops
    from funsor.domains import Bint
    from funsor.tensor import Tensor
    from funsor.terms import Slice, Variable

    x = torch.randn(4, 5, device=DEV)
    for out in (ops.new_arange(x, 3), ops.new_arange(x, 1, 7, 2), ops.new_eye(x, (3,)), ops.new_zeros(x, (2, 3)), ops.new_full(x, (2,), 1.5),
                ops.logsumexp(x, -1), ops.cat([x, x], 0), ops.stack((x, x), 0), ops.einsum([x, x], "ab,cb->ac"),
                ops.cholesky(x @ x.T + 5 * torch.eye(4, device=DEV)), ops.amax(x, -1), ops.exp(x), ops.expand(x[None], (2, 4, 5))):
        assert out.is_cuda, out
    L = ops.cholesky(x @ x.T + 5 * torch.eye(4, device=DEV))
    assert ops.triangular_solve(x, L).is_cuda
    # the substitutions the scan performs on CUDA data: Slice (even/odd views), renaming, Cat of the odd leftover, integer indexing
    f = Tensor(torch.randn(7, 3, 3, device=DEV), OrderedDict(time=Bint[7], prev=Bint[3], curr=Bint[3]))
    even = f(time=Slice("time", 0, 6, 2, 7), curr="drop")
    assert even.data.is_cuda and even.data.shape == (3, 3, 3)
    assert f(time=Variable("t2", Bint[7])).data.is_cuda and f(time=6).data.is_cuda
    idx = Tensor(torch.tensor([0, 2], device=DEV), OrderedDict(k=Bint[2]), 7)
    assert f(time=idx).data.is_cuda


# ------------------------------------------------------------------------------------------------ (f4) higher-order / plated planners
SARKKA_BILMES = {  # test_sum_product.py:2658-2760: inputs of `trans`, durations
    "lag1": (OrderedDict([("a", 3), ("b", 2), ("_PREV_b", 2)]), [2, 3, 6, 33]),
    "lag1_and_2": (OrderedDict([("a", 4), ("b", 3), ("_PREV_b", 3), ("c", 2), ("_PREV__PREV_c", 2)]), [2, 5, 8, 32]),
    "lag2": (OrderedDict([("a", 4), ("c", 2), ("_PREV__PREV_c", 2)]), [2, 7, 8, 40]),
    "lag1_and_3": (OrderedDict([("a", 2), ("_PREV_a", 2), ("_PREV__PREV__PREV_a", 2)]), [3, 6, 9, 30]),
    "wide": (OrderedDict([("a", 48), ("_PREV_a", 48), ("x", 2)]), [4, 17, 64]),
}


@pytest.mark.parametrize("sname", ["logaddexp", "max"])
@pytest.mark.parametrize("case", sorted(SARKKA_BILMES))
def test_sarkka_bilmes_product_through_the_rules(funsor, plugin, sname, case):
    """sum_product.py:856-937: the higher-order Markov planner re-blocks `trans` into period-sized steps and calls
    mixed_sequential_sum_product with SEVERAL (prev, curr) pairs per step; every level reaches the Contraction rule with several
    reduced names (contract_named folds them into one k)."""
    import funsor.ops as ops
    from funsor.domains import Bint
    from funsor.sum_product import sarkka_bilmes_product
    from funsor.tensor import Tensor
    from funsor.terms import Variable

    op = _ops(funsor, sname)
    sizes, durations = SARKKA_BILMES[case]
    for duration in durations:
        g = torch.Generator().manual_seed(duration)
        data = torch.randn(duration, *sizes.values(), generator=g)
        inputs = OrderedDict([("time", Bint[duration])] + [(k, Bint[v]) for k, v in sizes.items()])

        def build(put):
            return sarkka_bilmes_product(op, ops.add, Tensor(put(data), inputs), Variable("time", Bint[duration]), frozenset())

        want = build(lambda t: t.double())
        plugin.stats(reset=True)
        got = build(lambda t: t.to(DEV))
        st = plugin.stats()
        period = max(k.count("_PREV_") for k in sizes)
        if duration // period >= 2:
            assert st["contraction_fast"] > 0, (case, duration, st)
        assert set(got.inputs) == set(want.inputs)
        got = got.align(tuple(want.inputs))
        rel_close(got.data.cpu().numpy(), want.data.numpy(), rtol=1e-5, atol=1.0)


@pytest.mark.parametrize("sname", ["logaddexp", "max"])
@pytest.mark.parametrize("impl", ["modified_partial_sum_product", "dynamic_partial_sum_product"])
@pytest.mark.parametrize("x_dim,y_dim,time", [(2, 3, 5), (48, 3, 37), (33, 40, 16)])
def test_plated_markov_planners_through_the_rules(funsor, plugin, sname, impl, x_dim, y_dim, time):
    """sum_product.py:486-600 (modified_partial_sum_product) and 344-483 (dynamic_partial_sum_product) on the factor graph of
    test_sum_product.py:413-467: a Markov chain in a `time` plate with an emission per step; the time plate is eliminated by
    sequential_sum_product inside the planner."""
    from functools import reduce

    import funsor.ops as ops
    import funsor.sum_product as sp
    from funsor.domains import Bint
    from funsor.tensor import Tensor

    op = _ops(funsor, sname)
    g = torch.Generator().manual_seed(x_dim + time)
    datas = [torch.randn((), generator=g), torch.randn(x_dim, generator=g), torch.randn(time, x_dim, x_dim, generator=g),
             torch.randn(x_dim, y_dim, generator=g), torch.randn(time, x_dim, y_dim, generator=g)]
    names = [OrderedDict(), OrderedDict(x_0=Bint[x_dim]), OrderedDict(time=Bint[time], x_prev=Bint[x_dim], x_curr=Bint[x_dim]),
             OrderedDict(x_0=Bint[x_dim], y_0=Bint[y_dim]), OrderedDict(time=Bint[time], x_curr=Bint[x_dim], y_curr=Bint[y_dim])]
    plate_to_step = {"time": frozenset({("x_0", "x_prev", "x_curr")})}
    vars1, vars2 = frozenset({"y_0", "y_curr"}), frozenset({"time", "x_0", "x_prev", "x_curr"})

    def build(put):
        factors = [Tensor(put(d), n) for d, n in zip(datas, names)]
        f1 = getattr(sp, impl)(op, ops.add, factors, vars1, plate_to_step)
        f2 = getattr(sp, impl)(op, ops.add, f1, vars2, plate_to_step)
        return reduce(ops.add, f2)

    want = build(lambda t: t.double())
    plugin.stats(reset=True)
    got = build(lambda t: t.to(DEV))
    st = plugin.stats()
    assert st["contraction_fast"] + st["markov_fast"] > 0, st
    assert isinstance(got, Tensor) and not got.inputs
    rel_close(got.data.cpu().numpy(), want.data.numpy(), rtol=1e-5, atol=1.0)


# ------------------------------------------------------------------------------------------------ (f3) moment matching / SwitchingLinearHMM
def test_moment_matching_rule_inside_reference(funsor, plugin):
    import funsor.ops as ops
    from funsor.cnf import Contraction
    from funsor.domains import Bint, Reals
    from funsor.gaussian import Gaussian
    from funsor.interpretations import moment_matching
    from funsor.tensor import Tensor
    from funsor.terms import Variable
    from funsor.testing import random_gaussian

    torch.manual_seed(30)
    torch.set_default_dtype(torch.float64)
    try:
        proto = random_gaussian(OrderedDict([("b", Bint[3]), ("k", Bint[4]), ("x", Reals[6])]))
    finally:
        torch.set_default_dtype(torch.float32)
    logits = torch.randn(4, 3, dtype=torch.float64)

    def build(put):
        g = Gaussian(white_vec=put(proto.white_vec), prec_sqrt=put(proto.prec_sqrt), inputs=proto.inputs)
        d = Tensor(put(logits), OrderedDict(k=Bint[4], b=Bint[3]))
        with moment_matching:
            return Contraction(ops.logaddexp, ops.add, frozenset({Variable("k", Bint[4])}), d, g)

    torch.set_default_dtype(torch.float64)
    try:
        want = build(lambda t: t.clone())
    finally:
        torch.set_default_dtype(torch.float32)
    plugin.stats(reset=True)
    got = build(lambda t: t.float().to(DEV))
    assert plugin.stats()["moment_matching_fast"] == 1, plugin.stats()
    _same_function(funsor, got, want, ["x"], 6, 2e-4)


@pytest.mark.parametrize("T,C,D", [(6, 2, 2), (16, 3, 4), (9, 2, 8)])
def test_switching_linear_hmm_expression_inside_reference(funsor, plugin, T, C, D):
    """pyro/hmm.py:527-551 (SwitchingLinearHMM.log_prob, moment-matching mode) evaluated by funsor on CUDA: the scan over
    (class, state) elements under `moment_matching`, with the mixture collapse and the Gaussian pair contraction on the kernels."""
    import funsor.ops as ops
    from funsor.domains import Bint, Reals
    from funsor.interpretations import moment_matching
    from funsor.sum_product import sequential_sum_product
    from funsor.tensor import Tensor
    from funsor.terms import Variable
    from funsor.testing import random_gaussian

    torch.manual_seed(31)
    torch.set_default_dtype(torch.float64)
    try:
        gt = random_gaussian(OrderedDict([("time", Bint[T]), ("class(time=1)", Bint[C]), ("state", Reals[D]), ("state(time=1)", Reals[D])]))
        gi = random_gaussian(OrderedDict([("state", Reals[D])]))
        tl = torch.randn(T, C, C).log_softmax(-1)
        il = torch.randn(C).log_softmax(-1)
    finally:
        torch.set_default_dtype(torch.float32)

    def build(put):
        trans = (Tensor(put(tl), OrderedDict([("time", Bint[T]), ("class", Bint[C]), ("class(time=1)", Bint[C])]))
                 + type(gt)(white_vec=put(gt.white_vec), prec_sqrt=put(gt.prec_sqrt), inputs=gt.inputs))
        init = Tensor(put(il), OrderedDict([("class", Bint[C])])) + type(gi)(white_vec=put(gi.white_vec), prec_sqrt=put(gi.prec_sqrt), inputs=gi.inputs)
        with moment_matching:
            r = sequential_sum_product(ops.logaddexp, ops.add, trans, Variable("time", Bint[T]),
                                       {"class": "class(time=1)", "state": "state(time=1)"})
            r = r + init
            return r.reduce(ops.logaddexp, frozenset(["class", "state", "class(time=1)", "state(time=1)"]))

    torch.set_default_dtype(torch.float64)
    try:
        want = build(lambda t: t.clone())
    finally:
        torch.set_default_dtype(torch.float32)
    plugin.stats(reset=True)
    got = build(lambda t: t.float().to(DEV))
    assert plugin.stats()["moment_matching_fast"] > 0, plugin.stats()
    assert isinstance(got, Tensor) and got.data.is_cuda
    assert abs(float(got.data) - float(want.data)) <= 1e-4 * (1 + abs(float(want.data))), (float(got.data), float(want.data))
