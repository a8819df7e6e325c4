"""Test-infrastructure shim (oracle only): the reference hard-depends on the
`multipledispatch` package (pyproject.toml:29-35), which is not installed in
this image.  PyTorch vendors the same library, so this module re-exports it
under the expected name.  Never imported by the product path."""
import sys as _sys

from torch.fx.experimental.unification import multipledispatch as _vendored
from torch.fx.experimental.unification.multipledispatch import (  # noqa: F401
    Dispatcher,
    dispatch,
)
from torch.fx.experimental.unification.multipledispatch import (  # noqa: F401
    conflict,
    core,
    dispatcher,
    utils,
    variadic,
)

for _sub in ("conflict", "core", "dispatcher", "utils", "variadic"):
    _sys.modules[__name__ + "." + _sub] = getattr(_vendored, _sub)
