"""Test-infrastructure shim (oracle only) for the three `opt_einsum` entry
points the reference touches: `get_symbol`, `contract` (cnf.py:377,
tensor.py:1177) and `paths.greedy` (optimizer.py:6).  `contract` forwards to the
named backend module's `einsum`, which is how the reference reaches its
log-semiring backend (cnf.py:335-340)."""
import importlib
import string

from . import paths  # noqa: F401

_LETTERS = string.ascii_lowercase + string.ascii_uppercase


def get_symbol(i):
    return _LETTERS[i] if i < len(_LETTERS) else chr(i + 140)


def contract(equation, *operands, backend="numpy", **_unused):
    if backend == "numpy":
        import numpy

        return numpy.einsum(equation, *operands)
    if backend == "torch":
        import torch

        return torch.einsum(equation, *operands)
    return importlib.import_module(backend).einsum(equation, *operands)
