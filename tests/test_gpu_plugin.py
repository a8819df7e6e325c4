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


@pytest.mark.parametrize("T,D,r", [(8, 2, 3), (37, 4, 6), (200, 16, 24)])
def test_gaussian_markov_rule_detects_time_invariant_prec_sqrt(T, D, r):
    """GaussianHMM with fixed dynamics hands the rule a dense prec_sqrt whose time slices are equal: the kernel entry detects
    it and takes the per-level-operator path; the result must match the general scan of the same (materialised) elements."""
    import funsor_b200.gaussian as fg
    from funsor_b200 import funsor_plugin as fp

    g = torch.Generator().manual_seed(T + D)
    B, n = 3, 2 * D
    P1 = torch.randn(B, 1, n, r, generator=g) / n ** 0.5
    P1[:, 0, :D, :D] += torch.eye(D)  # enough information on the shared block
    P1[:, 0, D:, :D] -= torch.eye(D)
    P = P1.expand(B, T, n, r).contiguous().cuda()
    w = torch.randn(B, T, r, generator=g).cuda()
    s = torch.randn(B, T, generator=g).cuda()
    fp.stats(reset=True)
    got = fp._kernel_gaussian_chain(w, P, s)
    assert fp.stats()["gaussian_homogeneous"] == 1
    want = fg.gaussian_sequential_sum_product(*fg.sqrt_to_info(w, P, s))
    for a, b in zip(got, want):
        assert torch.allclose(a, b, rtol=2e-5, atol=2e-5 * float(b.abs().max())), (a - b).abs().max()
    P[:, T // 2] += 0.01  # one deviating step -> general path
    fp.stats(reset=True)
    fp._kernel_gaussian_chain(w, P, s)
    assert fp.stats().get("gaussian_homogeneous", 0) == 0
