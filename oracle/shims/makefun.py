"""Test-infrastructure shim (oracle only) for `makefun.with_signature`, the one
symbol the reference uses (funsor/factory.py:218, funsor/distribution.py:386).
It wraps a function so that it exposes a given signature."""
import functools
import inspect


def _parse(sig):
    if isinstance(sig, inspect.Signature):
        return sig
    inner = sig[sig.index("(") + 1 : sig.rindex(")")]
    return inspect.signature(eval("lambda " + inner + ": None"))


def with_signature(func_signature, **attrs):
    def decorate(fn):
        sig = _parse(func_signature)
        call = ", ".join("{0}={0}".format(n) for n in sig.parameters)
        scope = {}
        exec("def wrapper{}:\n    return _target({})".format(sig, call), {"_target": fn}, scope)
        wrapper = functools.update_wrapper(scope["wrapper"], fn)
        wrapper.__signature__ = sig
        for key, val in attrs.items():
            setattr(wrapper, key, val)
        return wrapper

    return decorate
