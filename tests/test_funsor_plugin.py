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

    from funsor_b200.funsor_plugin import gaussian_markov_named, info_to_sqrt

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
    w, P, s = info_to_sqrt(Lam, eta, c)
    assert torch.allclose(P @ P.transpose(-1, -2), wLam, rtol=1e-7, atol=1e-7)
    assert torch.allclose((P @ w[..., None])[..., 0], weta, rtol=1e-6, atol=1e-6)
    assert torch.allclose(s - 0.5 * (w * w).sum(-1), wc, rtol=1e-7, atol=1e-7)


def test_info_to_sqrt_handles_singular_precision(funsor):
    from funsor_b200.funsor_plugin import info_to_sqrt

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
