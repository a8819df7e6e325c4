"""CPU: pin the oracle (oracle/*.py restatements) against fixtures produced by the unmodified
reference (oracle/make_golden.py).  These are the checks that make parity 'pinned'."""
import numpy as np
import pytest

from conftest import golden_cases, load_golden, rel_close
from oracle import gaussian_ref as G
from oracle import semiring_ref as R

SEMI = {"log": R.LOG, "max": R.MAX}


@pytest.mark.parametrize("sname", ["log", "max"])
def test_pair_product(sname):
    for case, d in golden_cases(load_golden("pair_product")).items():
        got = R.pair_product(SEMI[sname], d["x"], d["y"])
        rel_close(got, d[sname], rtol=1e-12, atol=1.0)


def test_max_argmax_first_index():
    x = np.zeros((1, 2, 4))
    y = np.zeros((1, 4, 3))
    y[0, 2, 1] = 1.0
    y[0, 3, 1] = 1.0
    out, arg = R.max_matmul(x, y, want_arg=True)
    assert arg[0, 0, 0] == 0 and arg[0, 0, 1] == 2  # ties -> first index (approximations.py:115-127)


@pytest.mark.parametrize("sname", ["log", "max"])
def test_chain_product(sname):
    s = SEMI[sname]
    for case, d in golden_cases(load_golden("chain_product")).items():
        T = d["trans"].shape[1]
        rel_close(R.chain_product(s, d["trans"]), d[sname + "_seq"], rtol=1e-11 * T, atol=1.0)
        if sname + "_naive" in d:
            rel_close(R.chain_product_naive(s, d["trans"]), d[sname + "_naive"], rtol=1e-11 * T, atol=1.0)
            rel_close(R.chain_product(s, d["trans"]), d[sname + "_markov"], rtol=1e-11 * T, atol=1.0)
            for seg in (2, 3):
                rel_close(R.chain_product_mixed(s, d["trans"], seg), d[sname + "_mixed%d" % seg], rtol=1e-11 * T, atol=1.0)


def test_max_chain_bit_exact_fp64():
    # max-plus in a fixed association order is exactly reproducible
    for case, d in golden_cases(load_golden("chain_product")).items():
        assert np.array_equal(R.chain_product(R.MAX, d["trans"]), d["max_seq"]), case


def test_hmm_logprob():
    g = golden_cases(load_golden("hmm_logprob"))
    c1 = g.pop("c1")
    z = R.reduce_last(R.LOG, c1["init"].astype(np.float64) + R.reduce_last(R.LOG, R.chain_product(R.LOG, c1["trans"].astype(np.float64))), -1)
    rel_close(z, c1["logZ_fp64"], rtol=1e-12)
    rel_close(z, c1["logZ_fp32"], rtol=1e-5)
    z32 = R.hmm_forward(R.LOG, c1["init"], c1["trans"], np.zeros((1, 100, 16), np.float32))
    rel_close(z32, c1["logZ_fp64"], rtol=1e-5)
    for case, d in g.items():
        for sname, s in SEMI.items():
            rel_close(R.hmm_logprob_via_chain(s, d["init"], d["A"], d["obs"]), d[sname + "_Z"], rtol=1e-11, atol=1.0)
            rel_close(R.hmm_forward(s, d["init"], d["A"], d["obs"]), d[sname + "_Z"], rtol=1e-11, atol=1.0)


def test_viterbi_matches_bruteforce_and_golden():
    import itertools

    g = golden_cases(load_golden("hmm_logprob"))
    d = g["h2"]
    best, path = R.viterbi(d["init"], d["A"], d["obs"])
    rel_close(best, d["max_Z"], rtol=1e-12, atol=1.0)
    B, T, S = d["obs"].shape
    for b in range(B):
        scores = {}
        for p in itertools.product(range(S), repeat=T + 1):
            pp = np.array([p])
            scores[p] = R.path_score(d["init"][b : b + 1], d["A"][b : b + 1], d["obs"][b : b + 1], pp)[0]
        bf = max(scores, key=scores.get)
        assert tuple(path[b]) == bf
    rel_close(R.path_score(d["init"], d["A"], d["obs"], path), best, rtol=1e-12, atol=1.0)


@pytest.mark.parametrize("sname", ["log", "max"])
def test_adjoint(sname):
    for case, d in golden_cases(load_golden("adjoint")).items():
        Z, adj = R.forward_backward_dense(SEMI[sname], d["trans"])
        rel_close(Z, d[sname + "_Z"], rtol=1e-11, atol=1.0)
        rel_close(adj, d[sname + "_adj"], rtol=1e-10, atol=1.0)


def test_hmm_marginals_consistency():
    d = golden_cases(load_golden("hmm_logprob"))["h0"]
    logZ, gamma, xi = R.hmm_marginals(d["init"], d["A"], d["obs"])
    rel_close(logZ, d["log_Z"], rtol=1e-11, atol=1.0)
    np.testing.assert_allclose(np.exp(gamma).sum(-1), 1.0, atol=1e-10)
    np.testing.assert_allclose(xi.sum((-1, -2)), d["obs"].shape[1], rtol=1e-10)
    # pairwise marginals agree with the dense adjoint: xi_t = adj + trans - Z (forward_backward.py:162)
    M = R.hmm_dense_trans(d["A"], d["obs"])
    Z, adj = R.forward_backward_dense(R.LOG, M, init=d["init"])
    np.testing.assert_allclose(np.exp(adj + M - Z[:, None, None, None]).sum(1), xi, rtol=1e-9)


def _chain_inputs(d, D):
    return G.sqrt_to_info(d["w"], d["P"], np.zeros(d["w"].shape[:-1]))


def test_gaussian_chain_sqrt_and_info():
    for case, d in golden_cases(load_golden("gaussian_chain")).items():
        D = int(case.split("_D")[1])
        T = d["w"].shape[1]
        w, P, s = G.chain_sqrt(d["w"], d["P"], np.zeros(d["w"].shape[:-1]), D)
        Lam, eta, c = G.sqrt_to_info(w, P, s)
        for got, key in ((Lam, "Lam"), (eta, "eta"), (c, "c")):
            rel_close(got, d[key], rtol=1e-8 * T, atol=1.0)
        Lam, eta, c = G.chain_info(*_chain_inputs(d, D), D)
        for got, key in ((Lam, "Lam"), (eta, "eta"), (c, "c")):
            rel_close(got, d[key], rtol=1e-8 * T, atol=1.0)
        if "Lam_naive" in d:
            rel_close(Lam, d["Lam_naive"], rtol=1e-7 * T, atol=1.0)
            rel_close(c, d["c_naive"], rtol=1e-7 * T, atol=1.0)


def test_kalman_elements_chain_logprob():
    for case, d in golden_cases(load_golden("kalman")).items():
        D = d["F"].shape[-1]
        args = (d["F"], d["q_loc"], d["q_tril"], d["H"], d["r_loc"], d["r_tril"], d["y"])
        Lam, eta, c = G.kalman_elements_info(*args)
        chain = G.chain_info(Lam, eta, c, D)
        for got, key in zip(chain, ("chain_Lam", "chain_eta", "chain_c")):
            rel_close(got, d[key], rtol=1e-8, atol=1.0)
        w, P, s = G.chain_sqrt(*G.kalman_elements_sqrt(*args), D)
        for got, key in zip(G.sqrt_to_info(w, P, s), ("chain_Lam", "chain_eta", "chain_c")):
            rel_close(got, d[key], rtol=1e-8, atol=1.0)
        lp = G.hmm_logprob_from_chain(G.mvn_info(d["m0"], d["p0_tril"]), chain, D)
        rel_close(lp, d["logprob"], rtol=1e-9, atol=1.0)
        # independent absolute check: textbook Kalman filter
        kf = G.kalman_filter_logprob(d["m0"], d["p0_tril"], *args)
        rel_close(kf, d["logprob"], rtol=1e-9, atol=1.0)


def test_cpu_port_matches_oracle():
    import torch

    from oracle import cpu_port

    d = golden_cases(load_golden("hmm_logprob"))["h0"]
    for sname, s in SEMI.items():
        got = cpu_port.hmm_log_prob(torch.tensor(d["init"]), torch.tensor(d["A"]), torch.tensor(d["obs"]), sname)
        rel_close(got.numpy(), d[sname + "_Z"], rtol=1e-11, atol=1.0)


# ---------------------------------------------------------------------- chain distributions (funsor/pyro/hmm.py)
def _normal_logp(x, loc, scale):
    return -0.5 * ((x - loc) / scale) ** 2 - np.log(scale) - 0.5 * np.log(2 * np.pi)


def test_discrete_hmm_oracle_matches_reference():
    from conftest import dist_cases
    from oracle import hmm_ref as H

    cases = dist_cases("discrete")
    assert len(cases) >= 6
    for name, d in cases.items():
        obs = _normal_logp(d["data"][..., None].astype(np.float64), d["loc"], d["scale"])
        got = H.discrete_hmm_log_prob(d["init_logits"], d["trans_logits"], obs)
        want = d["log_prob"]
        rel_close(np.broadcast_to(got, want.shape) if got.size == want.size else got, want.reshape(np.shape(got)), rtol=1e-10, atol=1.0)


def test_gaussian_hmm_oracle_matches_reference():
    from conftest import dist_cases, ghmm_oracle

    cases = dist_cases("ghmm")
    assert len(cases) >= 6
    for name, d in cases.items():
        got = ghmm_oracle(d)
        rel_close(got, d["log_prob"].reshape(got.shape), rtol=1e-8, atol=1.0)


def test_gaussian_mrf_oracle_matches_reference():
    from conftest import dist_cases, mrf_oracle

    cases = dist_cases("mrf")
    assert len(cases) >= 5
    for name, d in cases.items():
        got = mrf_oracle(d)
        rel_close(got, d["log_prob"].reshape(got.shape), rtol=1e-8, atol=1.0)


def test_dense_posterior_oracle_agrees_with_the_sequential_filter():
    """oracle/gaussian_ref.kalman_posterior_dense (the law forward-filter backward-sample draws from) against the textbook
    sequential Kalman filter and against the golden log-probabilities written by the reference (tests/golden/kalman.npz)."""
    from oracle import gaussian_ref as G

    for name, d in golden_cases(load_golden("kalman")).items():
        args = [d[k].astype(np.float64) for k in ("m0", "p0_tril", "F", "q_loc", "q_tril", "H", "r_loc", "r_tril", "y")]
        mean, cov, logZ = G.kalman_posterior_dense(*args)
        np.testing.assert_allclose(logZ, d["logprob"], rtol=1e-8, atol=1e-8)
        np.testing.assert_allclose(logZ, G.kalman_filter_logprob(*args), rtol=1e-9, atol=1e-9)
        assert np.all(np.linalg.eigvalsh(cov) > 0)


def test_moment_matching_oracle_matches_reference_fixture():
    """oracle/gaussian_ref.moment_match (joint.py:102-150 restated) against tests/golden/moment_matching.npz, which
    oracle/make_golden.py wrote by running the reference's `moment_matching` interpretation."""
    from oracle import gaussian_ref as G

    for name, d in golden_cases(load_golden("moment_matching")).items():
        Lam, eta, c = G.sqrt_to_info(d["w"], d["P"], np.zeros(d["w"].shape[:-1]))
        Lo, eo, co = G.moment_match(d["logits"], Lam, eta, c)
        np.testing.assert_allclose(Lo, d["Lam"], rtol=1e-9, atol=1e-9, err_msg=name)
        np.testing.assert_allclose(eo, d["eta"], rtol=1e-9, atol=1e-9, err_msg=name)
        np.testing.assert_allclose(co, d["c"], rtol=1e-9, atol=1e-9, err_msg=name)
