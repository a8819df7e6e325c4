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
