"""CPU: the C-ABI library loads, exports every symbol include/funsor_b200.h declares, and the product
path refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from funsor_b200 import _lib, build

    build.build()
    return _lib.load()


def _declared():
    text = open(os.path.join(ROOT, "include", "funsor_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fb2_\w+)\s*\(", text)))


def test_header_symbols_exported(lib):
    from funsor_b200 import _lib

    names = _declared()
    assert len(names) >= 20
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), "declared in the header but not exported: " + n
    assert sorted(_lib.SIGNATURES) == names, set(_lib.SIGNATURES) ^ set(names)


def test_version_and_status_strings(lib):
    assert lib.fb2_version() == 100
    for code in range(7):
        assert lib.fb2_status_string(code)


def test_no_cpu_fallback():
    import funsor_b200 as fb

    x = torch.randn(2, 3, 4, 4)
    with pytest.raises(ValueError, match="CUDA"):
        fb.sequential_sum_product("logaddexp", "add", x)
    with pytest.raises(ValueError, match="CUDA"):
        fb.hmm_forward(torch.zeros(4), torch.zeros(4, 4), torch.zeros(1, 5, 4))
    with pytest.raises(NotImplementedError):
        fb.ops.semiring_id("add", "mul")
    with pytest.raises(ValueError):
        fb.ops._dev_f32(torch.zeros(1, dtype=torch.float64), "x")


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_device_entry_points_fail_loudly_without_gpu(lib):
    from funsor_b200 import _lib

    arr = (ctypes.c_int64 * 4)(1, 1, 1, 1)
    st = lib.fb2_semiring_matmul_f32(0, 8, arr, 8, arr, 8, None, 1, 1, 1, 1, 1, None, 0, None)
    assert st == _lib.ERR_NO_DEVICE
    with pytest.raises(_lib.Fb2Error):
        _lib.check(st)
