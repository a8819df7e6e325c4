import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))


def golden_cases(npz):
    """Group 'case/key' entries of a fixture into {case: {key: array}}."""
    cases = {}
    for k, v in npz.items():
        case, key = k.split("/", 1)
        cases.setdefault(case, {})[key] = v
    return cases


@pytest.fixture(scope="session")
def golden():
    return lambda name: golden_cases(load_golden(name))


def rel_close(actual, expected, rtol, atol=0.0):
    """The reference comparator (funsor/testing.py:176-178): |d| / (atol + |expected|) < rtol,
    with infinities required to match exactly."""
    actual = np.asarray(actual, dtype=np.float64)
    expected = np.asarray(expected, dtype=np.float64)
    assert actual.shape == expected.shape, (actual.shape, expected.shape)
    inf = ~np.isfinite(expected)
    assert np.array_equal(actual[inf], expected[inf]), "non-finite entries differ"
    d = np.abs(actual[~inf] - expected[~inf])
    bound = rtol * (atol + np.abs(expected[~inf]))
    if d.size and not (d <= bound).all():
        i = np.argmax(d - bound)
        raise AssertionError("max violation: |d|={:.3e} bound={:.3e} expected={:.6e}".format(d[i], bound[i], expected[~inf][i]))
