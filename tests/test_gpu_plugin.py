"""GPU: the named-dimension logic of the funsor rules (funsor_b200/funsor_plugin.py) driven with the real
kernels, against a torch fp64 evaluation of the same contraction.  funsor itself is absent on the GPU box;
tests/test_funsor_plugin.py pins the same functions against the unmodified reference on the CPU side."""
import numpy as np
import pytest
import torch

from conftest import rel_close

pytestmark = pytest.mark.gpu

SIZES = dict(b=3, t=5, i=40, j=33, k=70, l=2)
CASES = [
    (("b", "t", "i", "k"), ("b", "t", "k", "j"), ("k",)),
    (("k", "i"), ("j", "k"), ("k",)),
    (("t", "i", "k"), ("k", "j", "t"), ("k",)),
    (("i", "k"), ("k",), ("k",)),
    (("b", "k", "l", "i"), ("l", "b", "k", "j"), ("k", "l")),
    (("b", "i", "k"), ("k", "j"), ("k",)),
]


def ref_contract(sname, x, xn, y, yn, red, out_names):
    names = list(dict.fromkeys(list(xn) + list(yn)))
    xe = x.double().permute([xn.index(n) for n in names if n in xn]).reshape([SIZES[n] if n in xn else 1 for n in names])
    ye = y.double().permute([yn.index(n) for n in names if n in yn]).reshape([SIZES[n] if n in yn else 1 for n in names])
    z = xe + ye
    dims = [names.index(r) for r in red]
    z = torch.logsumexp(z, dims) if sname == "logaddexp" else z.amax(dims)
    kept = [n for n in names if n not in red]
    return z.permute([kept.index(n) for n in out_names])


@pytest.mark.parametrize("sname", ["logaddexp", "max"])
@pytest.mark.parametrize("xn,yn,red", CASES)
def test_contract_named_kernel(sname, xn, yn, red):
    from funsor_b200.funsor_plugin import _kernel_matmul, contract_named

    torch.manual_seed(0)
    x = torch.randn(*[SIZES[n] for n in xn], device="cuda")
    y = torch.randn(*[SIZES[n] for n in yn], device="cuda")
    data, names = contract_named(x, xn, y, yn, set(red), sname, _kernel_matmul)
    want = ref_contract(sname, x, xn, y, yn, red, names)
    rel_close(data.cpu().numpy(), want.cpu().numpy(), rtol=1e-5, atol=1.0)


@pytest.mark.parametrize("sname", ["logaddexp", "max"])
@pytest.mark.parametrize("order", [("b", "time", "prev", "curr"), ("time", "curr", "b", "prev"), ("prev", "curr", "time")])
def test_markov_named_kernel(sname, order):
    from funsor_b200.funsor_plugin import _kernel_chain, markov_named
    from oracle import semiring_ref as R

    sizes = dict(b=2, time=37, prev=48, curr=48)
    torch.manual_seed(1)
    data = torch.randn(*[sizes[n] for n in order], device="cuda")
    got, names = markov_named(data, order, "time", "prev", "curr", sname, _kernel_chain)
    batch = [n for n in order if n == "b"]
    canon = data.permute([order.index(n) for n in batch + ["time", "prev", "curr"]]).double().cpu().numpy()
    want = R.chain_product(R.LOG if sname == "logaddexp" else R.MAX, canon)
    got = got.permute([names.index(n) for n in batch + ["prev", "curr"]])
    rel_close(got.cpu().numpy(), want, rtol=1e-5, atol=1.0)


@pytest.mark.parametrize("order", [("b", "time", "prev", "curr"), ("time", "b", "curr", "prev"), ("time", "prev", "curr")])
@pytest.mark.parametrize("T,D", [(1, 2), (9, 3), (40, 8), (17, 32)])
def test_gaussian_markov_named_kernel(order, T, D):
    """The Gaussian MarkovProduct rule's tensor logic with the real kernels (sqrt->info, scan) against the numpy oracle,
    then back to the reference's sqrt representation."""
    from funsor_b200.funsor_plugin import _kernel_gaussian_chain, gaussian_markov_named, info_to_sqrt
    from oracle import gaussian_ref as G

    g = torch.Generator().manual_seed(7)
    sizes = dict(b=3, time=T)
    bints = [n for n in order if n in sizes]
    reals = [(n, D) for n in order if n in ("prev", "curr")]
    bshape = [sizes[n] for n in bints]
    n = 2 * D
    P = torch.randn(*bshape, n, n, generator=g) / n ** 0.5 + 0.7 * torch.eye(n)
    w = torch.randn(*bshape, n, generator=g)
    s = torch.randn(*bshape, generator=g)
    Lam, eta, c, batch = gaussian_markov_named(w.cuda(), P.cuda(), s.cuda(), bints, reals, "time", "prev", "curr", _kernel_gaussian_chain)
    # oracle on canonical layout [B, T, ...], rows (prev; curr)
    perm = [bints.index(x) for x in batch + ["time"]]
    wc = w.permute(perm + [len(bints)]).double().numpy()
    Pc = P.permute(perm + [len(bints), len(bints) + 1]).double()
    if reals[0][0] != "prev":
        Pc = torch.cat([Pc[..., D:, :], Pc[..., :D, :]], dim=-2)
    sc = s.permute(perm).double().numpy()
    wL, we, wcc = G.sqrt_to_info(wc, Pc.numpy(), sc)
    wL, we, wcc = G.chain_info(wL.reshape(-1, T, n, n), we.reshape(-1, T, n), wcc.reshape(-1, T), D)
    shape = tuple(sizes[x] for x in batch)
    rel_close(Lam.cpu().numpy(), wL.reshape(shape + (n, n)), rtol=2e-4, atol=1.0)
    rel_close(eta.cpu().numpy(), we.reshape(shape + (n,)), rtol=2e-4, atol=1.0)
    rel_close(c.cpu().numpy(), wcc.reshape(shape), rtol=1e-5, atol=float(T * D))  # c sums T terms of magnitude ~D
    w2, P2, s2 = info_to_sqrt(Lam, eta, c)
    assert torch.allclose(P2 @ P2.transpose(-1, -2), Lam, rtol=1e-4, atol=1e-4)
    assert torch.allclose(s2 - 0.5 * (w2 * w2).sum(-1), c, rtol=1e-5, atol=1e-3)
