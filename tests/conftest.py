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


# ---------------------------------------------------------------------- chain distributions (tests/golden/distributions.npz)
def dist_cases(kind):
    """{case: {key: array}} for kind in ('discrete', 'ghmm', 'mrf')"""
    cases = {}
    for k, v in load_golden("distributions").items():
        knd, case, key = k.split("/", 2)
        if knd == kind:
            cases.setdefault(case, {})[key] = v
    return cases


def _timed(a, tail_ndim, batch):
    """param [..., (T|1), *tail] or [*tail] -> broadcast to batch + (T|1,) + tail"""
    a = np.asarray(a, np.float64)
    if a.ndim == tail_ndim:
        a = a[None]
    tp = a.shape[-1 - tail_ndim]
    return np.broadcast_to(a, tuple(batch) + (tp,) + a.shape[a.ndim - tail_ndim:])


def ghmm_oracle(d):
    """dense float64 evaluation of GaussianHMM.log_prob for a fixture dict (oracle/hmm_ref.py), looped over the batch"""
    from oracle import hmm_ref as H

    y = np.asarray(d["data"], np.float64)
    batch = np.broadcast_shapes(d["init_loc"].shape[:-1], d["trans_mat"].shape[:-3], d["trans_loc"].shape[:-2], d["obs_mat"].shape[:-3],
                                d["obs_loc"].shape[:-2], y.shape[:-2])
    cov = lambda L: L @ np.swapaxes(L, -1, -2)  # noqa: E731
    m0 = np.broadcast_to(d["init_loc"], batch + d["init_loc"].shape[-1:])
    P0 = np.broadcast_to(cov(d["init_tril"].astype(np.float64)), batch + d["init_tril"].shape[-2:])
    F, ql, Q = _timed(d["trans_mat"], 2, batch), _timed(d["trans_loc"], 1, batch), _timed(cov(d["trans_tril"].astype(np.float64)), 2, batch)
    Hm, rl, R = _timed(d["obs_mat"], 2, batch), _timed(d["obs_loc"], 1, batch), _timed(cov(d["obs_tril"].astype(np.float64)), 2, batch)
    y = np.broadcast_to(y, batch + y.shape[-2:])
    out = np.zeros(batch)
    for i in np.ndindex(*batch):
        out[i] = H.gaussian_hmm_log_prob(m0[i].astype(np.float64), P0[i], F[i], ql[i], Q[i], Hm[i], rl[i], R[i], y[i])
    return out


def mrf_oracle(d):
    from oracle import hmm_ref as H

    y = np.asarray(d["data"], np.float64)
    batch = np.broadcast_shapes(d["init_loc"].shape[:-1], d["trans_loc"].shape[:-2], d["obs_loc"].shape[:-2], y.shape[:-2])
    cov = lambda L: L @ np.swapaxes(L, -1, -2)  # noqa: E731
    m0 = np.broadcast_to(d["init_loc"], batch + d["init_loc"].shape[-1:])
    P0 = np.broadcast_to(cov(d["init_tril"].astype(np.float64)), batch + d["init_tril"].shape[-2:])
    tl, tc = _timed(d["trans_loc"], 1, batch), _timed(cov(d["trans_tril"].astype(np.float64)), 2, batch)
    ol, oc = _timed(d["obs_loc"], 1, batch), _timed(cov(d["obs_tril"].astype(np.float64)), 2, batch)
    y = np.broadcast_to(y, batch + y.shape[-2:])
    out = np.zeros(batch)
    for i in np.ndindex(*batch):
        out[i] = H.gaussian_mrf_log_prob(m0[i].astype(np.float64), P0[i], tl[i], tc[i], ol[i], oc[i], y[i])
    return out
