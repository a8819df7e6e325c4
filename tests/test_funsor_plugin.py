"""CPU: the funsor-side half of the drop-in boundary (funsor_b200/funsor_plugin.py) against the UNMODIFIED
reference, imported from /root/reference (build container only; skipped where the reference is absent).

The kernels cannot run here, so the rules' named-dimension logic (`contract_named`, `markov_named`) is
driven with torch stand-ins for the two kernel entry points and compared with the reference's own
eager evaluation; tests/test_gpu_plugin.py drives the same functions with the real kernels."""
from collections import OrderedDict

import pytest
import torch

from oracle import ref_boot

pytestmark = pytest.mark.skipif(not ref_boot.available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def funsor():
    return ref_boot.boot("float64")


def torch_matmul(semiring, x, y):
    z = x.unsqueeze(-1) + y.unsqueeze(-3)  # [..., I, K, J]
    return torch.logsumexp(z, -2) if semiring == "logaddexp" else z.amax(-2)


def torch_chain(semiring, t):
    T = t.shape[-3]
    out = t[..., 0, :, :]
    for i in range(1, T):
        out = torch_matmul(semiring, out, t[..., i, :, :])
    return out


CASES = [
    # (x inputs, y inputs, reduced)
    (("b", "t", "i", "k"), ("b", "t", "k", "j"), ("k",)),
    (("k", "i"), ("j", "k"), ("k",)),
    (("t", "i", "k"), ("k", "j", "t"), ("k",)),
    (("i", "k"), ("k",), ("k",)),
    (("k",), ("k", "j"), ("k",)),
    (("b", "k", "l", "i"), ("l", "b", "k", "j"), ("k", "l")),
    (("b", "i", "k"), ("k", "j"), ("k",)),
]
SIZES = dict(b=2, t=3, i=4, j=5, k=6, l=2)


@pytest.mark.parametrize("sname", ["logaddexp", "max"])
@pytest.mark.parametrize("xn,yn,red", CASES)
def test_contract_named_matches_reference_contraction(funsor, sname, xn, yn, red):
    import funsor.ops as ops
    from funsor.cnf import Contraction
    from funsor.domains import Bint
    from funsor.tensor import Tensor
    from funsor.terms import Variable

    from funsor_b200.funsor_plugin import contract_named

    torch.manual_seed(0)
    x = torch.randn(*[SIZES[n] for n in xn])
    y = torch.randn(*[SIZES[n] for n in yn])
    fx = Tensor(x, OrderedDict((n, Bint[SIZES[n]]) for n in xn))
    fy = Tensor(y, OrderedDict((n, Bint[SIZES[n]]) for n in yn))
    op = ops.logaddexp if sname == "logaddexp" else ops.max
    want = Contraction(op, ops.add, frozenset(Variable(r, Bint[SIZES[r]]) for r in red), fx, fy)
    assert isinstance(want, Tensor)
    data, names = contract_named(x, xn, y, yn, set(red), sname, torch_matmul)
    assert tuple(names) == tuple(want.inputs)
    assert torch.allclose(data, want.data, rtol=1e-10, atol=1e-10)


def test_contract_named_declines_non_matrix_products(funsor):
    from funsor_b200.funsor_plugin import contract_named

    x, y = torch.zeros(2, 3), torch.zeros(3, 4)
    assert contract_named(x, ("a", "b"), y, ("b", "c"), {"a"}, "max", torch_matmul) is None  # 'a' only in x
    assert contract_named(x, ("a", "b"), y, ("b", "c"), set(), "max", torch_matmul) is None


@pytest.mark.parametrize("sname", ["logaddexp", "max"])
@pytest.mark.parametrize("order", [("b", "time", "prev", "curr"), ("time", "curr", "b", "prev"), ("prev", "curr", "time")])
def test_markov_named_matches_reference_scan(funsor, sname, order):
    import funsor.ops as ops
    from funsor.domains import Bint
    from funsor.sum_product import sequential_sum_product
    from funsor.tensor import Tensor
    from funsor.terms import Variable

    from funsor_b200.funsor_plugin import markov_named

    sizes = dict(b=2, time=7, prev=3, curr=3)
    torch.manual_seed(1)
    data = torch.randn(*[sizes[n] for n in order])
    trans = Tensor(data, OrderedDict((n, Bint[sizes[n]]) for n in order))
    op = ops.logaddexp if sname == "logaddexp" else ops.max
    want = sequential_sum_product(op, ops.add, trans, Variable("time", Bint[7]), {"prev": "curr"})
    got, names = markov_named(data, order, "time", "prev", "curr", sname, torch_chain)
    assert set(names) == set(want.inputs)
    got = got.permute([names.index(n) for n in want.inputs])
    assert torch.allclose(got, want.data, rtol=1e-9, atol=1e-9)


def test_rules_register_and_delegate_on_cpu_tensors(funsor):
    """With CPU tensors every rule must delegate to the handler it replaced: same values as before
    registration, never a lazy term (SURVEY 8(b) decline convention)."""
    import funsor.ops as ops
    from funsor.cnf import Contraction
    from funsor.domains import Bint
    from funsor.sum_product import MarkovProduct
    from funsor.tensor import Tensor
    from funsor.terms import Variable

    from funsor_b200 import funsor_plugin

    torch.manual_seed(2)
    x = Tensor(torch.randn(2, 4, 5), OrderedDict(b=Bint[2], i=Bint[4], k=Bint[5]))
    y = Tensor(torch.randn(2, 5, 3), OrderedDict(b=Bint[2], k=Bint[5], j=Bint[3]))
    rv = frozenset({Variable("k", Bint[5])})
    trans = Tensor(torch.randn(6, 3, 3), OrderedDict(time=Bint[6], prev=Bint[3], curr=Bint[3]))
    tv = Variable("time", Bint[6])
    before = {
        "log": Contraction(ops.logaddexp, ops.add, rv, x, y),
        "max": Contraction(ops.max, ops.add, rv, x, y),
        "mlog": MarkovProduct(ops.logaddexp, ops.add, trans, tv, {"prev": "curr"}),
        "mmax": MarkovProduct(ops.max, ops.add, trans, tv, {"prev": "curr"}),
    }
    funsor_plugin.register()
    funsor_plugin.register()  # idempotent
    funsor_plugin.stats(reset=True)
    # fresh data objects: funsor cons-hashes terms by tensor identity, new tensors force re-evaluation
    x2 = Tensor(x.data.clone(), x.inputs)
    y2 = Tensor(y.data.clone(), y.inputs)
    t2 = Tensor(trans.data.clone(), trans.inputs)
    after = {
        "log": Contraction(ops.logaddexp, ops.add, rv, x2, y2),
        "max": Contraction(ops.max, ops.add, rv, x2, y2),
        "mlog": MarkovProduct(ops.logaddexp, ops.add, t2, tv, {"prev": "curr"}),
        "mmax": MarkovProduct(ops.max, ops.add, t2, tv, {"prev": "curr"}),
    }
    st = funsor_plugin.stats()
    assert st["contraction_delegated"] >= 2 and st["markov_delegated"] >= 2, st
    assert st["contraction_fast"] == 0 and st["markov_fast"] == 0
    for k in before:
        assert isinstance(after[k], Tensor), (k, type(after[k]))
        assert tuple(after[k].inputs) == tuple(before[k].inputs)
        assert torch.allclose(after[k].data, before[k].data, rtol=1e-12, atol=1e-12)


# ------------------------------------------------------------------ Gaussian chain elements
def _oracle_gaussian_chain(w, P, s):
    from oracle import gaussian_ref as G

    Lam, eta, c = G.sqrt_to_info(w.double().numpy(), P.double().numpy(), s.double().numpy())
    D = P.shape[-2] // 2
    Lo, eo, co = G.chain_info(Lam, eta, c, D)
    return torch.from_numpy(Lo), torch.from_numpy(eo), torch.from_numpy(co)


def _info_of(funsor, result, names, D):
    """(Lam, eta, c) of a `Tensor + Gaussian` funsor over real inputs `names` (as oracle/make_golden.py does)."""
    from funsor.cnf import Contraction
    from funsor.domains import Reals
    from funsor.gaussian import Gaussian
    from funsor.tensor import Tensor

    terms = result.terms if isinstance(result, Contraction) else (result,)
    (g,) = [t for t in terms if isinstance(t, Gaussian)]
    reals = [k for k, d in g.inputs.items() if d.dtype == "real"]
    assert sorted(reals) == sorted(names)
    perm = torch.cat([torch.arange(D) + D * reals.index(n) for n in names])
    zero = {n: funsor.to_funsor(torch.zeros(D), Reals[D]) for n in names}
    c = result(**zero)
    assert isinstance(c, Tensor)
    bnames = [k for k, d in g.inputs.items() if d.dtype != "real"]
    Lam = g._precision[..., perm, :][..., :, perm]
    eta = g._info_vec[..., perm]
    return Lam, eta, c.align(tuple(bnames)).data


@pytest.mark.parametrize("order", [("b", "time", "prev", "curr"), ("time", "b", "curr", "prev"), ("time", "prev", "curr")])
@pytest.mark.parametrize("T,D", [(1, 1), (3, 2), (8, 2), (11, 3)])
def test_gaussian_markov_named_matches_reference_scan(funsor, order, T, D):
    import funsor.ops as ops
    from funsor.domains import Bint, Reals
    from funsor.sum_product import sequential_sum_product
    from funsor.terms import Variable
    from funsor.testing import random_gaussian

    from funsor_b200.funsor_plugin import gaussian_markov_named

    torch.manual_seed(3)
    doms = dict(b=Bint[2], time=Bint[T], prev=Reals[D], curr=Reals[D])
    g = random_gaussian(OrderedDict((n, doms[n]) for n in order))
    want = sequential_sum_product(ops.logaddexp, ops.add, g, Variable("time", Bint[T]), {"prev": "curr"})
    wLam, weta, wc = _info_of(funsor, want, ["prev", "curr"], D)
    bints = [n for n in order if n in ("b", "time")]
    reals = [(n, D) for n in order if n in ("prev", "curr")]
    Lam, eta, c, batch = gaussian_markov_named(g.white_vec, g.prec_sqrt, None, bints, reals, "time", "prev", "curr", _oracle_gaussian_chain)
    assert batch == [n for n in bints if n != "time"]
    assert torch.allclose(Lam, wLam, rtol=1e-8, atol=1e-8)
    assert torch.allclose(eta, weta, rtol=1e-8, atol=1e-8)
    assert torch.allclose(c, wc, rtol=1e-8, atol=1e-8)
    # and back to the reference's representation: the (w, P, s) triple must describe the same function
    w, P, s = _np_info_to_sqrt(Lam, eta, c)  # the oracle's twin of fb2_gaussian_info_to_sqrt_f32
    assert torch.allclose(P @ P.transpose(-1, -2), wLam, rtol=1e-7, atol=1e-7)
    assert torch.allclose((P @ w[..., None])[..., 0], weta, rtol=1e-6, atol=1e-6)
    assert torch.allclose(s - 0.5 * (w * w).sum(-1), wc, rtol=1e-7, atol=1e-7)


def test_info_to_sqrt_handles_singular_precision(funsor):
    info_to_sqrt = _np_info_to_sqrt

    torch.manual_seed(4)
    A = torch.randn(3, 6, 2, dtype=torch.float64)  # rank 2 of 6
    Lam = A @ A.transpose(-1, -2)
    eta = (A @ torch.randn(3, 2, 1, dtype=torch.float64))[..., 0]  # in the range of Lam
    c = torch.randn(3, dtype=torch.float64)
    w, P, s = info_to_sqrt(Lam, eta, c)
    assert torch.allclose(P @ P.transpose(-1, -2), Lam, atol=1e-10)
    assert torch.allclose((P @ w[..., None])[..., 0], eta, atol=1e-9)
    assert torch.allclose(s - 0.5 * (w * w).sum(-1), c, atol=1e-10)


def test_gaussian_rule_delegates_on_cpu(funsor):
    import funsor.ops as ops
    from funsor.domains import Bint, Reals
    from funsor.sum_product import MarkovProduct
    from funsor.terms import Variable
    from funsor.testing import random_gaussian

    from funsor_b200 import funsor_plugin

    torch.manual_seed(5)
    inputs = OrderedDict(b=Bint[2], time=Bint[5], prev=Reals[2], curr=Reals[2])
    g1, g2 = random_gaussian(inputs), None
    tv = Variable("time", Bint[5])
    before = MarkovProduct(ops.logaddexp, ops.add, g1, tv, {"prev": "curr"})
    funsor_plugin.register()
    funsor_plugin.stats(reset=True)
    from funsor.gaussian import Gaussian

    g2 = Gaussian(white_vec=g1.white_vec.clone(), prec_sqrt=g1.prec_sqrt.clone(), inputs=g1.inputs)
    after = MarkovProduct(ops.logaddexp, ops.add, g2, tv, {"prev": "curr"})
    assert funsor_plugin.stats()["markov_delegated"] >= 1 and funsor_plugin.stats()["markov_fast"] == 0
    for x, y in zip(_info_of(funsor, before, ["prev", "curr"], 2), _info_of(funsor, after, ["prev", "curr"], 2)):
        assert torch.allclose(x, y, rtol=1e-10, atol=1e-10)


def test_plugin_sequential_sum_product_function(funsor):
    """funsor_plugin.sequential_sum_product (what hmm.py-style direct callers would import) == the reference function, for
    Tensor and Gaussian chain elements (CPU tensors here, so the rules delegate)."""
    import funsor.ops as ops
    from funsor.domains import Bint, Reals
    from funsor.sum_product import sequential_sum_product as ref_ssp
    from funsor.tensor import Tensor
    from funsor.terms import Variable
    from funsor.testing import random_gaussian

    from funsor_b200 import funsor_plugin

    funsor_plugin.register()
    torch.manual_seed(6)
    tv = Variable("time", Bint[6])
    trans = Tensor(torch.randn(2, 6, 3, 3), OrderedDict(b=Bint[2], time=Bint[6], prev=Bint[3], curr=Bint[3]))
    for op in (ops.logaddexp, ops.max):
        want = ref_ssp(op, ops.add, trans, tv, {"prev": "curr"})
        got = funsor_plugin.sequential_sum_product(op, ops.add, trans, tv, {"prev": "curr"})
        assert isinstance(got, Tensor) and set(got.inputs) == set(want.inputs)
        assert torch.allclose(got.align(tuple(want.inputs)).data, want.data, rtol=1e-10, atol=1e-10)
    g = random_gaussian(OrderedDict(b=Bint[2], time=Bint[6], prev=Reals[2], curr=Reals[2]))
    want = ref_ssp(ops.logaddexp, ops.add, g, tv, {"prev": "curr"})
    got = funsor_plugin.sequential_sum_product(ops.logaddexp, ops.add, g, tv, {"prev": "curr"})
    for x, y in zip(_info_of(funsor, want, ["prev", "curr"], 2), _info_of(funsor, got, ["prev", "curr"], 2)):
        assert torch.allclose(x, y, rtol=1e-9, atol=1e-9)


# ------------------------------------------------------------------ the rules' FAST branches inside the unmodified reference
# The kernels cannot run in the CPU container, so the entry points the rules call are replaced by oracle stand-ins
# (funsor_plugin.set_hooks); everything else -- pattern matching, named-dimension bookkeeping, input order, Subs, cons-hashing,
# AdjointTape -- is the code that runs on the GPU box, where tests/test_gpu_funsor.py repeats these cases with the real kernels.
def _np_sqrt_to_info(w, P, s):
    from oracle import gaussian_ref as G

    return tuple(torch.from_numpy(np.asarray(a)) for a in G.sqrt_to_info(w.double().numpy(), P.double().numpy(), s.double().numpy()))


def _np_pair(X, Y):
    from oracle import gaussian_ref as G

    D = X[0].shape[-1] // 2
    out = G.info_pair_combine(tuple(t.numpy() for t in X), tuple(t.numpy() for t in Y), D)
    return tuple(torch.from_numpy(np.asarray(a)) for a in out) + (None,)


def _np_info_to_sqrt(Lam, eta, c):
    from oracle import gaussian_ref as G

    return tuple(torch.from_numpy(np.asarray(a)) for a in G.info_to_sqrt(Lam.numpy(), eta.numpy(), c.numpy()))


import numpy as np  # noqa: E402


@pytest.fixture
def fast(funsor):
    from funsor_b200 import funsor_plugin

    funsor_plugin.register()
    funsor_plugin.set_hooks(fast_ok=lambda *d: all(isinstance(t, torch.Tensor) for t in d), matmul=torch_matmul, chain=torch_chain,
                            gaussian_chain=_oracle_gaussian_chain, sqrt_to_info=_np_sqrt_to_info, gaussian_pair=_np_pair,
                            info_to_sqrt=_np_info_to_sqrt)
    funsor_plugin.stats(reset=True)
    yield funsor_plugin
    funsor_plugin.set_hooks()


def _both(plugin, build):
    """build() evaluated by the stock rules (hooks off: CPU tensors delegate) and by the fast branches; fresh tensors each time
    because funsor cons-hashes terms by tensor identity."""
    saved = dict(plugin._HOOKS)
    plugin.set_hooks()
    want = build()
    plugin.set_hooks(**saved)
    plugin.stats(reset=True)
    got = build()
    return want, got, plugin.stats()


@pytest.mark.parametrize("sname", ["logaddexp", "max"])
def test_fast_tensor_rules_inside_reference(funsor, fast, sname):
    import funsor.ops as ops
    from funsor.adjoint import AdjointTape
    from funsor.cnf import Contraction
    from funsor.domains import Bint
    from funsor.einsum import einsum
    from funsor.sum_product import MarkovProduct, mixed_sequential_sum_product, naive_sequential_sum_product, sequential_sum_product, sum_product
    from funsor.tensor import Tensor
    from funsor.terms import Variable
    from funsor.testing import assert_close

    op = ops.logaddexp if sname == "logaddexp" else ops.max
    g = torch.Generator().manual_seed(7)
    B, T, S = 2, 7, 3
    d_trans, d_init = torch.randn(B, T, S, S, generator=g), torch.randn(B, S, generator=g)
    d_x, d_y = torch.randn(B, 4, 5, generator=g), torch.randn(5, 3, B, generator=g)
    tv = Variable("time", Bint[T])

    def tensors():
        trans = Tensor(d_trans.clone(), OrderedDict(b=Bint[B], time=Bint[T], prev=Bint[S], curr=Bint[S]))
        init = Tensor(d_init.clone(), OrderedDict(b=Bint[B], prev=Bint[S]))
        x = Tensor(d_x.clone(), OrderedDict(b=Bint[B], i=Bint[4], k=Bint[5]))
        y = Tensor(d_y.clone(), OrderedDict(k=Bint[5], j=Bint[3], b=Bint[B]))
        return trans, init, x, y

    def contraction():
        _, _, x, y = tensors()
        return Contraction(op, ops.add, frozenset({Variable("k", Bint[5])}), x, y)

    def markov():
        trans = tensors()[0]
        return MarkovProduct(op, ops.add, trans, tv, {"prev": "curr"})

    def seq(fn, **kw):
        def build():
            trans = tensors()[0]
            return fn(op, ops.add, trans, tv, {"prev": "curr"}, **kw)
        return build

    def hmm_expression():  # pyro/hmm.py:128-137
        trans, init, _, _ = tensors()
        r = sequential_sum_product(op, ops.add, trans, tv, {"prev": "curr"})
        return (init + r.reduce(op, "curr")).reduce(op, "prev")

    def front_door():  # funsor.einsum.einsum routes by backend NAME to the semiring (einsum/__init__.py:13-28)
        _, _, x, y = tensors()
        backend = "pyro.ops.einsum.torch_log" if sname == "logaddexp" else "pyro.ops.einsum.torch_map"
        xs = Tensor(x.data[0], OrderedDict(a=Bint[4], b=Bint[5]))
        ys = Tensor(y.data[..., 0], OrderedDict(b=Bint[5], c=Bint[3]))
        return einsum("ab,bc->ac", xs, ys, backend=backend)

    def planner():  # the planner builds the expression lazily; the optimizer turns it into binary Contractions (einsum/__init__.py:126-128)
        from funsor.interpretations import lazy
        from funsor.optimizer import apply_optimizer

        _, _, x, y = tensors()
        z = Tensor(d_init.clone()[:, :3], OrderedDict(b=Bint[B], j=Bint[3]))
        with lazy:
            expr = sum_product(op, ops.add, [x, y, z], frozenset({"k", "j", "b"}), frozenset({"b"}))
        return apply_optimizer(expr)

    def adjoint():
        trans = tensors()[0]
        with AdjointTape() as tape:
            z = sequential_sum_product(op, ops.add, trans, tv, {"prev": "curr"})
            z = z.reduce(op, frozenset({"prev", "curr"}))
        return tape.adjoint(op, ops.add, z, (trans,), batch_vars=frozenset({Variable("b", Bint[B])}))[trans]

    cases = {"contraction": (contraction, "contraction_fast"), "markov": (markov, "markov_fast"),
             "sequential": (seq(sequential_sum_product), "contraction_fast"), "naive": (seq(naive_sequential_sum_product), None),  # Binary + Reduce under eager (sum_product.py:644-648): never a Contraction
             "mixed": (seq(mixed_sequential_sum_product, num_segments=3), "contraction_fast"), "hmm": (hmm_expression, "contraction_fast"),
             "einsum": (front_door, "contraction_fast"), "sum_product": (planner, "contraction_fast"), "adjoint": (adjoint, "contraction_fast")}
    for name, (build, counter) in cases.items():
        want, got, st = _both(fast, build)
        assert counter is None or st[counter] > 0, (name, st)
        assert isinstance(got, Tensor), (name, type(got))
        assert_close(got, want, atol=1e-9, rtol=1e-9)


def _same_gaussian_function(funsor, got, want, names, D, tol):
    """`Tensor + Gaussian` results describe the same function: same inputs in the same order, same (Lam, eta) and same value at
    x = 0.  (white_vec / prec_sqrt / the Tensor term are representation dependent: a rank-2D pivoted-Cholesky root here, the
    uncompressed rank-r root in the reference, whose |w|^2 differs; testing.py:149-153 compares info_vec and precision.)"""
    assert type(got) is type(want) and list(got.inputs.items()) == list(want.inputs.items()), (got.inputs, want.inputs)
    assert [type(t) for t in got.terms] == [type(t) for t in want.terms]
    assert [list(t.inputs.items()) for t in got.terms] == [list(t.inputs.items()) for t in want.terms]
    for a, b in zip(_info_of(funsor, got, names, D), _info_of(funsor, want, names, D)):
        assert torch.allclose(a, b, rtol=tol, atol=tol), (a - b).abs().max()


@pytest.mark.parametrize("T,D", [(2, 1), (5, 2), (8, 3), (11, 2)])
def test_fast_gaussian_contraction_rule_inside_reference(funsor, fast, T, D):
    """sequential_sum_product over Gaussian chain elements, called directly as pyro/hmm.py:269 does: every level reaches the
    Contraction[.., Gaussian|GaussianMixture x 2] rule; values AND input order must be the reference's (assert_close compares
    the OrderedDicts, funsor/testing.py:132)."""
    import funsor.ops as ops
    from funsor.domains import Bint, Reals
    from funsor.gaussian import Gaussian
    from funsor.sum_product import sequential_sum_product
    from funsor.terms import Variable
    from funsor.testing import assert_close, random_gaussian

    torch.manual_seed(8)
    proto = random_gaussian(OrderedDict(b=Bint[2], time=Bint[T], prev=Reals[D], curr=Reals[D]))

    def build():
        g = Gaussian(white_vec=proto.white_vec.clone(), prec_sqrt=proto.prec_sqrt.clone(), inputs=proto.inputs)
        return sequential_sum_product(ops.logaddexp, ops.add, g, Variable("time", Bint[T]), {"prev": "curr"})

    want, got, st = _both(fast, build)
    assert st["gaussian_contraction_fast"] > 0 and st["gaussian_contraction_delegated"] == 0, st
    _same_gaussian_function(funsor, got, want, ["prev", "curr"], D, 1e-7)


def test_fast_gaussian_hmm_expression_inside_reference(funsor, fast):
    """pyro/hmm.py:258-272 with rank-deficient Kalman elements (rank D + O < 2D) and a Tensor part on the elements."""
    import importlib.util
    import os

    import funsor.ops as ops
    from funsor.domains import Bint
    from funsor.sum_product import sequential_sum_product
    from funsor.tensor import Tensor
    from funsor.terms import Variable
    from funsor.testing import assert_close

    spec = importlib.util.spec_from_file_location("_ref_convert", os.path.join(ref_boot.REFERENCE_ROOT, "funsor/pyro/convert.py"))
    conv = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(conv)
    MVN = torch.distributions.MultivariateNormal
    torch.manual_seed(9)
    B, T, D, O = 2, 9, 4, 1
    F = 0.9 * torch.linalg.qr(torch.randn(B, D, D))[0]
    H = torch.randn(B, D, O)
    y = torch.randn(B, T, O)
    ex = lambda t: t.unsqueeze(1).expand((B, T) + t.shape[1:])  # noqa: E731

    def build():
        trans = conv.matrix_and_mvn_to_funsor(ex(F).clone(), MVN(torch.zeros(B, T, D), scale_tril=torch.eye(D).expand(B, T, D, D).clone()),
                                              ("time",), "state", "state(time=1)")
        obs = conv.matrix_and_mvn_to_funsor(ex(H).clone(), MVN(torch.zeros(B, T, O), scale_tril=torch.eye(O).expand(B, T, O, O).clone()),
                                            ("time",), "state(time=1)", "value")
        value = Tensor(y.clone(), OrderedDict([("_pyro_dim_-1", Bint[B]), ("time", Bint[T])]))
        elem = trans + obs(value=value)
        return sequential_sum_product(ops.logaddexp, ops.add, elem, Variable("time", Bint[T]), {"state": "state(time=1)"})

    want, got, st = _both(fast, build)
    assert st["gaussian_contraction_fast"] > 0, st
    _same_gaussian_function(funsor, got, want, ["state", "state(time=1)"], D, 1e-6)


@pytest.mark.parametrize("T", [1, 2, 3, 8])
@pytest.mark.parametrize("with_tensor", [False, True])
@pytest.mark.parametrize("order", [("b", "time", "prev", "curr"), ("time", "b", "curr", "prev")])
def test_fast_gaussian_markov_rule_inside_reference(funsor, fast, T, with_tensor, order):
    """MarkovProduct over Gaussian / Tensor + Gaussian elements through the whole-scan rule: same function AND the same input order
    as the reference's level-by-level evaluation."""
    import funsor.ops as ops
    from funsor.domains import Bint, Reals
    from funsor.gaussian import Gaussian
    from funsor.sum_product import MarkovProduct
    from funsor.tensor import Tensor
    from funsor.terms import Variable
    from funsor.testing import random_gaussian

    torch.manual_seed(10)
    D = 2
    doms = dict(b=Bint[2], time=Bint[T], prev=Reals[D], curr=Reals[D])
    proto = random_gaussian(OrderedDict((n, doms[n]) for n in order))
    tdata = torch.randn(T, 2)

    def build():
        g = Gaussian(white_vec=proto.white_vec.clone(), prec_sqrt=proto.prec_sqrt.clone(), inputs=proto.inputs)
        elem = g + Tensor(tdata.clone(), OrderedDict(time=Bint[T], b=Bint[2])) if with_tensor else g
        return MarkovProduct(ops.logaddexp, ops.add, elem, Variable("time", Bint[T]), {"prev": "curr"})

    want, got, st = _both(fast, build)
    assert st["markov_fast"] > 0, st
    if T == 1 and not with_tensor:
        assert list(got.inputs.items()) == list(want.inputs.items())
        return
    _same_gaussian_function(funsor, got, want, ["prev", "curr"], D, 1e-7)


@pytest.mark.parametrize("sname", ["logaddexp", "max"])
def test_fast_planners_inside_reference(funsor, fast, sname):
    """(f4) sarkka_bilmes_product (sum_product.py:856-937) and the plated Markov planners (486-600, 344-483) with the rules' fast
    branches on: several (prev, curr) pairs per step reach contract_named as several reduced names."""
    from functools import reduce

    import funsor.ops as ops
    import funsor.sum_product as sp
    from funsor.domains import Bint
    from funsor.tensor import Tensor
    from funsor.terms import Variable
    from funsor.testing import assert_close

    op = ops.logaddexp if sname == "logaddexp" else ops.max
    g = torch.Generator().manual_seed(21)
    sizes = OrderedDict([("a", 4), ("b", 3), ("_PREV_b", 3), ("c", 2), ("_PREV__PREV_c", 2)])
    for duration in (2, 5, 8):
        data = torch.randn(duration, *sizes.values(), generator=g)
        inputs = OrderedDict([("time", Bint[duration])] + [(k, Bint[v]) for k, v in sizes.items()])

        def build():
            return sp.sarkka_bilmes_product(op, ops.add, Tensor(data.clone(), inputs), Variable("time", Bint[duration]), frozenset())

        want, got, st = _both(fast, build)
        if duration >= 4:
            assert st["contraction_fast"] > 0, st
        assert_close(got.align(tuple(want.inputs)), want, atol=1e-9, rtol=1e-9)
    x_dim, y_dim, time = 2, 3, 5
    datas = [torch.randn((), generator=g), torch.randn(x_dim, generator=g), torch.randn(time, x_dim, x_dim, generator=g),
             torch.randn(x_dim, y_dim, generator=g), torch.randn(time, x_dim, y_dim, generator=g)]
    names = [OrderedDict(), OrderedDict(x_0=Bint[x_dim]), OrderedDict(time=Bint[time], x_prev=Bint[x_dim], x_curr=Bint[x_dim]),
             OrderedDict(x_0=Bint[x_dim], y_0=Bint[y_dim]), OrderedDict(time=Bint[time], x_curr=Bint[x_dim], y_curr=Bint[y_dim])]
    plate_to_step = {"time": frozenset({("x_0", "x_prev", "x_curr")})}
    for impl in (sp.modified_partial_sum_product, sp.dynamic_partial_sum_product):
        def build():
            factors = [Tensor(d.clone(), n) for d, n in zip(datas, names)]
            f1 = impl(op, ops.add, factors, frozenset({"y_0", "y_curr"}), plate_to_step)
            f2 = impl(op, ops.add, f1, frozenset({"time", "x_0", "x_prev", "x_curr"}), plate_to_step)
            return reduce(ops.add, f2)

        want, got, st = _both(fast, build)
        assert st["contraction_fast"] + st["markov_fast"] > 0, (impl.__name__, st)
        assert_close(got, want, atol=1e-9, rtol=1e-9)


def _np_moment_match(logits, Lam, eta, c):
    from oracle import gaussian_ref as G

    return tuple(torch.from_numpy(np.asarray(a)) for a in G.moment_match(logits.numpy(), Lam.numpy(), eta.numpy(), c.numpy()))


@pytest.mark.parametrize("order", [("b", "k", "x"), ("k", "b", "x")])
def test_fast_moment_matching_rule_inside_reference(funsor, fast, order):
    """(f3) the `moment_matching` interpretation's mixture collapse (joint.py:102-150) through the plugin rule: same function and the
    same input order as the reference."""
    import funsor.ops as ops
    from funsor.cnf import Contraction
    from funsor.domains import Bint, Reals
    from funsor.gaussian import Gaussian
    from funsor.interpretations import moment_matching
    from funsor.tensor import Tensor
    from funsor.terms import Variable
    from funsor.testing import random_gaussian

    fast.set_hooks(moment_match=_np_moment_match)
    torch.manual_seed(30)
    doms = dict(b=Bint[3], k=Bint[4], x=Reals[2])
    proto = random_gaussian(OrderedDict((n, doms[n]) for n in order))
    logits = torch.randn(4, 3)

    def build():
        g = Gaussian(white_vec=proto.white_vec.clone(), prec_sqrt=proto.prec_sqrt.clone(), inputs=proto.inputs)
        d = Tensor(logits.clone(), OrderedDict(k=Bint[4], b=Bint[3]))
        with moment_matching:
            return Contraction(ops.logaddexp, ops.add, frozenset({Variable("k", Bint[4])}), d, g)

    want, got, st = _both(fast, build)
    assert st["moment_matching_fast"] == 1, st
    _same_gaussian_function(funsor, got, want, ["x"], 2, 1e-8)


def test_fast_switching_linear_hmm_expression(funsor, fast):
    """pyro/hmm.py:527-551 (SwitchingLinearHMM.log_prob, moment-matching mode): the scan over (class, state) elements under the
    `moment_matching` interpretation with the plugin's collapse rule and Gaussian contraction rule on."""
    import funsor.ops as ops
    from funsor.domains import Bint, Reals
    from funsor.interpretations import moment_matching
    from funsor.sum_product import sequential_sum_product
    from funsor.tensor import Tensor
    from funsor.terms import Variable
    from funsor.testing import random_gaussian

    fast.set_hooks(moment_match=_np_moment_match)
    torch.manual_seed(31)
    T, C, D = 6, 2, 2
    gt = random_gaussian(OrderedDict([("time", Bint[T]), ("class(time=1)", Bint[C]), ("state", Reals[D]), ("state(time=1)", Reals[D])]))
    tl = torch.randn(T, C, C).log_softmax(-1)
    gi = random_gaussian(OrderedDict([("state", Reals[D])]))
    il = torch.randn(C).log_softmax(-1)

    def build():
        trans = (Tensor(tl.clone(), OrderedDict([("time", Bint[T]), ("class", Bint[C]), ("class(time=1)", Bint[C])]))
                 + type(gt)(white_vec=gt.white_vec.clone(), prec_sqrt=gt.prec_sqrt.clone(), inputs=gt.inputs))
        init = Tensor(il.clone(), OrderedDict([("class", Bint[C])])) + type(gi)(white_vec=gi.white_vec.clone(), prec_sqrt=gi.prec_sqrt.clone(), inputs=gi.inputs)
        with moment_matching:
            r = sequential_sum_product(ops.logaddexp, ops.add, trans, Variable("time", Bint[T]),
                                       {"class": "class(time=1)", "state": "state(time=1)"})
            r = r + init
            return r.reduce(ops.logaddexp, frozenset(["class", "state", "class(time=1)", "state(time=1)"]))

    want, got, st = _both(fast, build)
    assert st["moment_matching_fast"] > 0, st
    assert isinstance(got, Tensor)
    assert torch.allclose(got.data, want.data, rtol=1e-7, atol=1e-7), (got.data, want.data)
