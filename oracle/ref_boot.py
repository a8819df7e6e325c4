"""Test infrastructure: import the UNMODIFIED reference (funsor 0.4.8) from
/root/reference in the build container, to pin the oracle and to generate the
golden fixtures under tests/golden/.  /root/reference does not exist on the GPU
box; there the same unmodified package is found under baseline/_ref (installed with pip --target,
git-ignored) and is used by the `-m gpu` plugin tests and by `bench.py --impl reference` only.

What is stubbed and why:
  * multipledispatch / makefun / opt_einsum: absent third-party deps -> oracle/shims/.
  * funsor.torch.distributions hard-imports pyro (torch/distributions.py:8-13) -> empty module.
  * log-semiring backend: pyro.ops.einsum.torch_log (pyro-ppl >= 1.8.0, unpinned,
    pyproject.toml:38-42) is absent; the reference's own in-tree twin
    funsor/einsum/numpy_log.py is substituted (cnf.py:603-609).
"""
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIMS = os.path.join(_HERE, "shims")
# Where the unmodified reference lives: the source tree in the build container, or the copy that
# `pip install --no-deps --target baseline/_ref` made of it (git-ignored; it travels to the GPU box).
_CANDIDATES = [os.environ.get("FUNSOR_REFERENCE_ROOT"), "/root/reference", os.path.join(os.path.dirname(_HERE), "baseline", "_ref")]
REFERENCE_ROOT = next((p for p in _CANDIDATES if p and os.path.isdir(os.path.join(p, "funsor"))), "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "funsor"))


def boot(dtype="float64"):
    """Returns the imported reference `funsor` module with the torch backend set."""
    if not available():
        raise RuntimeError("reference tree not present at " + REFERENCE_ROOT)
    for p in (_SHIMS, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch

    sys.modules.setdefault("funsor.torch.distributions", types.ModuleType("funsor.torch.distributions"))
    import funsor
    import funsor.cnf

    funsor.cnf.BACKEND_TO_LOGSUMEXP_BACKEND["torch"] = "funsor.einsum.numpy_log"
    funsor.cnf.BACKEND_TO_MAP_BACKEND["torch"] = "funsor.einsum.numpy_map"
    funsor.set_backend("torch")
    torch.set_default_dtype(getattr(torch, dtype))
    return funsor
