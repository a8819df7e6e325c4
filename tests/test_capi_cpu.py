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


def test_chain_distribution_shapes_on_cpu():
    """Constructor-side logic of the funsor.pyro.hmm mirrors needs no GPU: batch_shape / event_shape follow the reference's
    broadcasting rules (pyro/hmm.py:82-88, 240-248, 335-341) for its own shape lists (test/pyro/test_hmm.py:26-44, 157-173)."""
    import numpy as np

    from funsor_b200.hmm import DiscreteHMM, GaussianHMM, GaussianMRF

    MVN = torch.distributions.MultivariateNormal
    S = 3
    for ish, tsh, osh in [((), (1,), ()), ((), (), (1,)), ((), (2,), ()), ((), (7,), (1,)), ((), (7,), (5, 7)), ((5,), (5, 7), (7,)),
                          ((4, 1, 1), (3, 1, 7), (2, 7))]:
        d = DiscreteHMM(torch.randn(ish + (S,)), torch.randn(tsh + (S, S)), torch.distributions.Normal(torch.randn(osh + (S,)), torch.ones(osh + (S,))))
        shape = np.broadcast_shapes(ish + (1,), tsh, osh)
        assert tuple(d.batch_shape) == tuple(shape[:-1]) and tuple(d.event_shape) == (shape[-1],)
        assert torch.allclose(d.initial_logits.logsumexp(-1), torch.zeros(ish), atol=1e-6)
        assert torch.allclose(d.transition_logits.logsumexp(-1), torch.zeros(tsh + (S,)), atol=1e-6)
        with pytest.raises(ValueError, match="CUDA"):
            d.log_prob(torch.randn(tuple(shape[:-1]) + (max(shape[-1], 2),)))
    D, O = 2, 3
    eye = lambda sh, n: torch.eye(n).expand(sh + (n, n))  # noqa: E731
    for ish, tmsh, tvsh, omsh, ovsh in [((), (), (), (), ()), ((), (6,), (), (), ()), ((5,), (6,), (), (), ()), ((), (5, 1), (6,), (), ()),
                                        ((5,), (5, 6), (5, 6), (5, 6), (5, 6))]:
        g = GaussianHMM(MVN(torch.zeros(ish + (D,)), scale_tril=eye(ish, D)), torch.randn(tmsh + (D, D)), MVN(torch.zeros(tvsh + (D,)), scale_tril=eye(tvsh, D)),
                        torch.randn(omsh + (D, O)), MVN(torch.zeros(ovsh + (O,)), scale_tril=eye(ovsh, O)))
        shape = np.broadcast_shapes(ish + (1,), tmsh, tvsh, omsh, ovsh)
        assert tuple(g.batch_shape) == tuple(shape[:-1]) and tuple(g.event_shape) == (shape[-1], O)
    m = GaussianMRF(MVN(torch.zeros(5, D), scale_tril=eye((5,), D)), MVN(torch.zeros(7, 2 * D), scale_tril=eye((7,), 2 * D)),
                    MVN(torch.zeros(5, 7, D + O), scale_tril=eye((5, 7), D + O)))
    assert tuple(m.batch_shape) == (5,) and tuple(m.event_shape) == (7, O) and m.hidden_dim == D and m.obs_dim == O


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` runs on host cores only (no GPU): one JSON line with the contract's keys."""
    import json
    import subprocess
    import sys

    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-T", "2"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] in ("reference", "port")  # the real funsor when baseline/_ref is installed
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["config"]["workload"] == "c2"
