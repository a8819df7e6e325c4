"""GPU parity tests for the discrete path: CUDA kernels (through the C ABI) vs the numpy oracle and the
golden fixtures produced by the unmodified reference.

Tolerances: the north star asks for 1e-5 relative (fp32) on log-normalisers; we use the reference
comparator |d| <= rtol * (atol + |expected|) (funsor/testing.py:176-178) with rtol = 1e-5 and atol = 1.
Max-plus values and all argmax / path indices are compared bit-exactly against an fp32 evaluation of
the same recurrence."""
import numpy as np
import pytest
import torch

from conftest import golden_cases, load_golden, rel_close
from oracle import semiring_ref as R

pytestmark = pytest.mark.gpu
RTOL = 1e-5
SEMI = {"log": ("logaddexp", R.LOG), "max": ("max", R.MAX)}


@pytest.fixture(scope="module")
def fb():
    import funsor_b200

    assert torch.cuda.is_available()
    return funsor_b200


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).cuda()


def host(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("sname", ["log", "max"])
def test_pair_product_golden(fb, sname):
    op, _ = SEMI[sname]
    for case, d in golden_cases(load_golden("pair_product")).items():
        got = fb.semiring_matmul(op, "add", dev(d["x"]), dev(d["y"]))
        rel_close(host(got), d[sname], rtol=RTOL, atol=1.0)


@pytest.mark.parametrize("sname", ["log", "max"])
@pytest.mark.parametrize("shape", [(3, 5, 70, 100, 33), (1, 2, 129, 64, 257), (2, 1, 64, 64, 64), (1, 1, 5, 300, 7)])
def test_pair_product_random_strided(fb, sname, shape):
    op, sid = SEMI[sname]
    n0, n1, I, K, J = shape
    g = torch.Generator().manual_seed(1)
    xx = torch.randn(n0, 2 * n1, I, K, generator=g) * 3
    yy = torch.randn(n0, 2 * n1, K, J, generator=g) * 3
    x, y = xx.cuda()[:, 0::2], yy.cuda()[:, 1::2]  # even/odd views as in sum_product.py:692-693
    got = fb.semiring_matmul(op, "add", x, y)
    want = R.pair_product(sid, xx[:, 0::2].double().numpy(), yy[:, 1::2].double().numpy())
    rel_close(host(got), want, rtol=RTOL, atol=1.0)
    # transposed (non unit inner stride) and broadcast operands
    got_t = fb.semiring_matmul(op, "add", yy.cuda()[:, 1::2].transpose(-1, -2), xx.cuda()[:, 0::2].transpose(-1, -2))
    rel_close(host(got_t), np.swapaxes(want, -1, -2), rtol=RTOL, atol=1.0)
    got_b = fb.semiring_matmul(op, "add", x[:1, :1], y)
    want_b = R.pair_product(sid, xx[:1, 0:1].double().numpy(), yy[:, 1::2].double().numpy())
    rel_close(host(got_b), want_b, rtol=RTOL, atol=1.0)


def test_pair_product_max_bit_exact_and_argmax(fb):
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 3, 65, 90, generator=g)
    y = torch.randn(2, 3, 90, 70, generator=g)
    y[0, 0, 5] = y[0, 0, 3]  # duplicate rows -> exact ties, first index must win
    x[0, 0, :, 5] = x[0, 0, :, 3]
    out, arg = fb.semiring_matmul("max", "add", x.cuda(), y.cuda(), want_argmax=True)
    want, warg = R.max_matmul(x.numpy(), y.numpy(), want_arg=True)
    assert np.array_equal(host(out), want)
    assert np.array_equal(host(arg).astype(np.int64), warg)
    out_s, arg_s = fb.semiring_matmul("max", "add", x[..., :9, :].cuda(), y[..., :7].cuda(), want_argmax=True)
    want, warg = R.max_matmul(x[..., :9, :].numpy(), y[..., :7].numpy(), want_arg=True)
    assert np.array_equal(host(out_s), want) and np.array_equal(host(arg_s).astype(np.int64), warg)


def test_pair_product_neg_inf(fb):
    x = torch.randn(1, 1, 40, 40)
    y = torch.randn(1, 1, 40, 40)
    x[0, 0, 3, :] = -float("inf")
    y[0, 0, :, 7] = -float("inf")
    x[0, 0, 10, 5] = -float("inf")
    for sname, (op, sid) in SEMI.items():
        got = host(fb.semiring_matmul(op, "add", x.cuda(), y.cuda()))
        want = R.pair_product(sid, x.double().numpy(), y.double().numpy())
        assert np.isneginf(got[0, 0, 3]).all() and np.isneginf(got[0, 0, :, 7]).all()
        assert not np.isnan(got).any()
        rel_close(got, want, rtol=RTOL, atol=1.0)


@pytest.mark.parametrize("sname", ["log", "max"])
def test_chain_product_golden(fb, sname):
    op, sid = SEMI[sname]
    for case, d in golden_cases(load_golden("chain_product")).items():
        T = d["trans"].shape[1]
        got = host(fb.sequential_sum_product(op, "add", dev(d["trans"])))
        rel_close(got, d[sname + "_seq"], rtol=RTOL, atol=1.0)
        if sname == "max":  # same association order as the reference -> bit-identical to its fp32 evaluation
            assert np.array_equal(got, R.chain_product(R.MAX, d["trans"].astype(np.float32))), case
        if sname + "_mixed2" in d and T >= 2:
            for seg in (2, 3):
                got = host(fb.mixed_sequential_sum_product(op, "add", dev(d["trans"]), num_segments=seg))
                rel_close(got, d[sname + "_mixed%d" % seg], rtol=RTOL, atol=1.0)


@pytest.mark.parametrize("sname", ["log", "max"])
@pytest.mark.parametrize("B,T,S", [(2, 33, 64), (1, 7, 130), (5, 1, 9), (3, 1000, 4), (1, 2, 256)])
def test_chain_product_random(fb, sname, B, T, S):
    op, sid = SEMI[sname]
    g = torch.Generator().manual_seed(3)
    trans = torch.randn(B, T, S, S, generator=g)
    got = host(fb.sequential_sum_product(op, "add", trans.cuda()))
    want = R.chain_product(sid, trans.double().numpy())
    rel_close(got, want, rtol=RTOL, atol=1.0)
    # strided input: time in a different position (Cat moves time to dim 0 in the reference, tensor.py:942-962)
    tt = trans.permute(1, 0, 2, 3).contiguous().cuda()
    got2 = host(fb.sequential_sum_product(op, "add", tt, time=0))
    assert np.array_equal(got, got2)


def test_c1_config(fb):
    d = golden_cases(load_golden("hmm_logprob"))["c1"]
    chain = fb.sequential_sum_product("logaddexp", "add", dev(d["trans"]))
    z = torch.logsumexp(dev(d["init"])[:, :, None] + chain, dim=(-1, -2))
    rel_close(host(z), d["logZ_fp64"], rtol=RTOL)
    z2 = fb.hmm_forward(dev(d["init"]), dev(d["trans"]), torch.zeros(1, 100, 16, device="cuda"))
    rel_close(host(z2), d["logZ_fp64"], rtol=RTOL)


@pytest.mark.parametrize("sname", ["log", "max"])
def test_hmm_forward_golden(fb, sname):
    op, sid = SEMI[sname]
    g = golden_cases(load_golden("hmm_logprob"))
    g.pop("c1")
    for case, d in g.items():
        got = fb.hmm_forward(dev(d["init"]), dev(d["A"]), dev(d["obs"]), sum_op=op)
        rel_close(host(got), d[sname + "_Z"], rtol=RTOL, atol=1.0)


@pytest.mark.parametrize("sname", ["log", "max"])
@pytest.mark.parametrize("B,T,S,amode", [(5, 200, 64, "shared"), (16, 40, 1024, "shared"), (3, 30, 100, "batched"),
                                         (2, 12, 37, "timevarying"), (40, 25, 16, "shared"), (1, 64, 700, "shared")])
def test_hmm_forward_random(fb, sname, B, T, S, amode):
    op, sid = SEMI[sname]
    g = torch.Generator().manual_seed(4)
    init = torch.randn(B, S, generator=g).log_softmax(-1)
    shape = {"shared": (S, S), "batched": (B, S, S), "timevarying": (B, T, S, S)}[amode]
    A = torch.randn(*shape, generator=g).log_softmax(-1)
    obs = torch.randn(B, T, S, generator=g)
    got, alpha = fb.hmm_forward(init.cuda(), A.cuda(), obs.cuda(), sum_op=op, return_alpha=True)
    want, walpha = R.hmm_forward(sid, init.double().numpy(), A.double().numpy(), obs.double().numpy(), want_alpha=True)
    rel_close(host(got), want, rtol=RTOL, atol=1.0)
    # filtered messages hover around 0 in the log domain while their error accumulates with t: the reference's own fp32
    # scan differs from its serial fp32 scan by 3.1e-5 at T=100, so the absolute part scales with T
    rel_close(host(alpha), walpha, rtol=RTOL, atol=max(1.0, T / 10.0))


def test_hmm_forward_neg_inf(fb):
    g = torch.Generator().manual_seed(5)
    B, T, S = 3, 20, 12
    init = torch.randn(B, S, generator=g)
    A = torch.randn(S, S, generator=g)
    A[0, 1] = -float("inf")
    A[:, 4] = -float("inf")  # unreachable state
    obs = torch.randn(B, T, S, generator=g)
    obs[1, 3, 2] = -float("inf")
    obs[2, 5, :] = -float("inf")  # impossible observation -> logZ = -inf, not NaN
    for sname, (op, sid) in SEMI.items():
        got = host(fb.hmm_forward(init.cuda(), A.cuda(), obs.cuda(), sum_op=op))
        want = R.hmm_forward(sid, init.double().numpy(), A.double().numpy(), obs.double().numpy())
        assert np.isneginf(got[2]) and not np.isnan(got).any()
        rel_close(got, want, rtol=RTOL, atol=1.0)


@pytest.mark.parametrize("B,T,S,amode", [(2, 50, 16, "shared"), (1, 300, 257, "shared"), (3, 20, 40, "batched"),
                                         (2, 9, 5, "timevarying"), (17, 11, 8, "shared")])
def test_viterbi_bit_exact(fb, B, T, S, amode):
    g = torch.Generator().manual_seed(6)
    init = torch.randn(B, S, generator=g)
    shape = {"shared": (S, S), "batched": (B, S, S), "timevarying": (B, T, S, S)}[amode]
    A = torch.randn(*shape, generator=g)
    obs = torch.randn(B, T, S, generator=g)
    best, path = fb.hmm_viterbi(init.cuda(), A.cuda(), obs.cuda())
    wbest, wpath = R.viterbi(init.numpy(), A.numpy(), obs.numpy())  # fp32, same recurrence
    assert np.array_equal(host(best), wbest)
    assert np.array_equal(host(path), wpath)
    # value parity with the reference's max-plus scan semantics in fp64
    want = R.hmm_forward(R.MAX, init.double().numpy(), A.double().numpy(), obs.double().numpy())
    rel_close(host(best), want, rtol=RTOL, atol=1.0)


def test_viterbi_ties_first_index(fb):
    B, T, S = 2, 6, 5
    init = torch.zeros(B, S)
    A = torch.zeros(S, S)
    obs = torch.zeros(B, T, S)
    best, path = fb.hmm_viterbi(init.cuda(), A.cuda(), obs.cuda())
    wbest, wpath = R.viterbi(init.numpy(), A.numpy(), obs.numpy())
    assert np.array_equal(host(path), wpath) and (host(path) == 0).all()


def test_viterbi_planted_path(fb):
    # SURVEY 8(d) C3 at reduced size: planted path with a margin must be recovered exactly
    g = torch.Generator().manual_seed(7)
    T, S = 400, 512
    A = torch.randn(S, S, generator=g)
    obs = torch.randn(1, T, S, generator=g)
    init = torch.randn(1, S, generator=g)
    z = torch.randint(0, S, (T + 1,), generator=g)
    init[0, z[0]] += 30.0
    obs[0, torch.arange(T), z[1:]] += 30.0
    best, path = fb.hmm_viterbi(init.cuda(), A.cuda(), obs.cuda())
    assert np.array_equal(host(path)[0], z.numpy())
    score = R.path_score(init.numpy(), A.numpy(), obs.numpy(), z.numpy()[None])
    rel_close(host(best), score, rtol=RTOL, atol=1.0)


@pytest.mark.parametrize("B,T,S,amode", [(2, 30, 16, "shared"), (1, 100, 130, "shared"), (3, 12, 20, "batched"), (2, 9, 6, "timevarying")])
def test_hmm_marginals(fb, B, T, S, amode):
    g = torch.Generator().manual_seed(8)
    init = torch.randn(B, S, generator=g).log_softmax(-1)
    shape = {"shared": (S, S), "batched": (B, S, S), "timevarying": (B, T, S, S)}[amode]
    A = torch.randn(*shape, generator=g).log_softmax(-1)
    obs = torch.randn(B, T, S, generator=g)
    logZ, gamma, xi = fb.hmm_marginals(init.cuda(), A.cuda(), obs.cuda())
    wZ, wgamma, wxi = R.hmm_marginals(init.double().numpy(), A.double().numpy(), obs.double().numpy())
    rel_close(host(logZ), wZ, rtol=RTOL, atol=1.0)
    rel_close(host(gamma), wgamma, rtol=1e-4, atol=1.0)
    np.testing.assert_allclose(host(xi), wxi, rtol=1e-4, atol=1e-5 * T)
    np.testing.assert_allclose(np.exp(host(gamma)).sum(-1), 1.0, atol=1e-4)


@pytest.mark.parametrize("sname", ["log", "max"])
def test_chain_adjoint_golden(fb, sname):
    op, sid = SEMI[sname]
    for case, d in golden_cases(load_golden("adjoint")).items():
        Z, adj = fb.chain_adjoint(op, "add", dev(d["trans"]))
        rel_close(host(Z), d[sname + "_Z"], rtol=RTOL, atol=1.0)
        rel_close(host(adj), d[sname + "_adj"], rtol=RTOL, atol=1.0)


def test_chain_adjoint_viterbi_argmax(fb):
    # per-step argmax of adj + trans under (max, add) is the Viterbi path when it is unique
    g = torch.Generator().manual_seed(9)
    B, T, S = 2, 15, 6
    trans = torch.randn(B, T, S, S, generator=g)
    init = torch.randn(B, S, generator=g)
    Z, adj = fb.chain_adjoint("max", "add", trans.cuda(), init=init.cuda())
    m = host(adj) + trans.numpy()
    _, path = R.viterbi(init.numpy(), trans.numpy(), np.zeros((B, T, S), np.float32))
    for b in range(B):
        for t in range(T):
            i, j = np.unravel_index(np.argmax(m[b, t]), (S, S))
            assert (i, j) == (path[b, t], path[b, t + 1])


# ------------------------------------------------------------------ size-independent properties at BASELINE sizes
def test_c2_full_size_uniform_transition_property(fb):
    """C2 (T=65536, S=1024, B=16): with a uniform transition matrix the log-likelihood factorises:
    logZ = lse(init) + sum_t (lse_j obs[t, j] - log S).  Checked in float64 on the device."""
    B, T, S = 16, 65536, 1024
    g = torch.Generator(device="cuda").manual_seed(0)
    obs = torch.randn(B, T, S, device="cuda", generator=g)
    init = torch.randn(B, S, device="cuda", generator=g).log_softmax(-1)
    A = torch.full((S, S), -float(np.log(S)), device="cuda")
    got = fb.hmm_forward(init, A, obs)
    want = torch.logsumexp(init.double(), -1) + (torch.logsumexp(obs.double(), -1) - np.log(S)).sum(-1)
    rel_close(host(got), host(want), rtol=RTOL)


def test_c2_full_size_shift_and_split_properties(fb):
    """C2 size, random A: (i) adding a constant per step to obs adds its sum to logZ; (ii) the chain
    splits: logZ(T) = lse_j(alpha_{T/2}[j] + rest) is checked through alpha hand-over at T/2."""
    B, T, S = 16, 65536, 1024
    g = torch.Generator(device="cuda").manual_seed(1)
    obs = torch.randn(B, T, S, device="cuda", generator=g)
    init = torch.randn(B, S, device="cuda", generator=g).log_softmax(-1)
    A = torch.randn(S, S, device="cuda", generator=g).log_softmax(-1)
    z = fb.hmm_forward(init, A, obs)
    shift = torch.randn(B, T, 1, device="cuda", generator=g)
    z_shift = fb.hmm_forward(init, A, obs + shift)
    rel_close(host(z_shift), host(z.double() + shift.double().sum((1, 2))), rtol=RTOL)
    half = T // 2
    z1, alpha = fb.hmm_forward(init, A, obs[:, :half].contiguous(), return_alpha=True)
    z2 = fb.hmm_forward(alpha[:, -1].contiguous(), A, obs[:, half:].contiguous())
    rel_close(host(z2), host(z), rtol=RTOL)


# ------------------------------------------------------------------ dataflow forward kernel (hmm_flow.cu)
def _flow_case(B, T, S, seed, obs_scale=1.0, obs_shift=0.0):
    g = torch.Generator().manual_seed(seed)
    init = torch.randn(B, S, generator=g).log_softmax(-1)
    A = torch.randn(S, S, generator=g).log_softmax(-1)
    obs = torch.randn(B, T, S, generator=g) * obs_scale + obs_shift
    return init, A, obs


@pytest.mark.parametrize("B,T,S", [(16, 40, 1024), (1, 64, 700), (5, 33, 65), (3, 1, 128), (40, 25, 257), (16, 2, 512),
                                   (2, 300, 100), (17, 50, 1000), (1, 3, 1024)])
def test_hmm_forward_flow_random(fb, B, T, S):
    """Homogeneous log-semiring chains with 64 < S <= 1024 and no alpha output run the dataflow kernel."""
    init, A, obs = _flow_case(B, T, S, seed=11)
    n0 = fb.launch_count(reset=True)
    got = fb.hmm_forward(init.cuda(), A.cuda(), obs.cuda())
    assert fb.launch_count() == 3, "expected colmax + flow + finalise launches, i.e. the dataflow path"
    want = R.hmm_forward(R.LOG, init.double().numpy(), A.double().numpy(), obs.double().numpy())
    rel_close(host(got), want, rtol=RTOL, atol=1.0)


def test_hmm_forward_flow_tight_short_chain(fb):
    """Short chains have small |logZ|: the 3xTF32 contraction must hold fp32-class accuracy (1e-5 of 1 + |logZ|)."""
    init, A, obs = _flow_case(16, 8, 256, seed=12)
    got = host(fb.hmm_forward(init.cuda(), A.cuda(), obs.cuda()))
    want = R.hmm_forward(R.LOG, init.double().numpy(), A.double().numpy(), obs.double().numpy())
    assert np.max(np.abs(got - want)) < 2e-5 * (1 + np.max(np.abs(want)))


def test_hmm_forward_flow_matches_generic_kernel(fb):
    import os

    init, A, obs = _flow_case(7, 120, 384, seed=13)
    z_flow = host(fb.hmm_forward(init.cuda(), A.cuda(), obs.cuda()))
    os.environ["FB2_DISABLE_FLOW"] = "1"
    try:
        z_gen = host(fb.hmm_forward(init.cuda(), A.cuda(), obs.cuda()))
    finally:
        os.environ.pop("FB2_DISABLE_FLOW")
    rel_close(z_flow, z_gen, rtol=RTOL, atol=1.0)


@pytest.mark.parametrize("B,T,S", [(16, 40, 1024), (5, 33, 65), (40, 25, 257), (3, 90, 160), (2, 300, 100), (4, 20, 130)])
def test_hmm_forward_flow_tcgen05_variant(fb, B, T, S):
    """The tcgen05 / tensor-memory implementation of the same dataflow (FB2_FLOW_TC=1), incl. a sparse banded chain."""
    import os

    init, A, obs = _flow_case(B, T, S, seed=21)
    if S == 160:  # left-to-right chain with a wide dynamic range
        A = torch.full((S, S), -float("inf"))
        for i in range(S):
            A[i, i] = np.log(0.6)
            A[i, (i + 1) % S] = np.log(0.4)
        init = torch.full((B, S), -float("inf"))
        init[:, 0] = 0.0
        obs = obs * 3
    if S == 130:
        A[:, 64:72] = -float("inf")
        obs[2, 5, :] = -float("inf")
    os.environ["FB2_FLOW_TC"] = "1"
    os.environ["FB2_FLOW_DIRECT"] = "0"
    try:
        got = host(fb.hmm_forward(init.cuda(), A.cuda(), obs.cuda()))
    finally:
        os.environ.pop("FB2_FLOW_TC")
        os.environ.pop("FB2_FLOW_DIRECT")
    want = R.hmm_forward(R.LOG, init.double().numpy(), A.double().numpy(), obs.double().numpy())
    assert not np.isnan(got).any()
    rel_close(got, want, rtol=RTOL, atol=1.0)


def test_hmm_forward_flow_extreme_observations(fb):
    """Observation log-likelihoods of magnitude ~1e3 with a spread of hundreds of nats between states
    (continuous-emission HMMs): block exponents absorb them; nothing over- or underflows."""
    init, A, obs = _flow_case(4, 60, 200, seed=14, obs_scale=150.0, obs_shift=-900.0)
    got = host(fb.hmm_forward(init.cuda(), A.cuda(), obs.cuda()))
    want = R.hmm_forward(R.LOG, init.double().numpy(), A.double().numpy(), obs.double().numpy())
    rel_close(got, want, rtol=RTOL, atol=1.0)


def test_hmm_forward_flow_neg_inf(fb):
    g = torch.Generator().manual_seed(15)
    B, T, S = 4, 20, 130
    init = torch.randn(B, S, generator=g)
    A = torch.randn(S, S, generator=g)
    A[0, 1] = -float("inf")
    A[:, 4] = -float("inf")  # unreachable state
    A[:, 64:72] = -float("inf")  # a whole column group unreachable
    obs = torch.randn(B, T, S, generator=g)
    obs[1, 3, 2] = -float("inf")
    obs[2, 5, :] = -float("inf")  # impossible observation -> logZ = -inf, not NaN
    init[3, :] = -float("inf")
    init[3, 7] = 0.0  # deterministic start
    got = host(fb.hmm_forward(init.cuda(), A.cuda(), obs.cuda()))
    want = R.hmm_forward(R.LOG, init.double().numpy(), A.double().numpy(), obs.double().numpy())
    assert np.isneginf(got[2]) and not np.isnan(got).any()
    rel_close(got, want, rtol=RTOL, atol=1.0)


def test_hmm_forward_flow_sparse_left_to_right(fb):
    """Banded (left-to-right) transition matrix: most of P is exactly zero and the message is sparse."""
    B, T, S = 3, 90, 160
    g = torch.Generator().manual_seed(16)
    A = torch.full((S, S), -float("inf"))
    for i in range(S):
        A[i, i] = np.log(0.6)
        A[i, (i + 1) % S] = np.log(0.4)
    init = torch.full((B, S), -float("inf"))
    init[:, 0] = 0.0
    obs = torch.randn(B, T, S, generator=g) * 3
    got = host(fb.hmm_forward(init.cuda(), A.cuda(), obs.cuda()))
    want = R.hmm_forward(R.LOG, init.double().numpy(), A.double().numpy(), obs.double().numpy())
    rel_close(got, want, rtol=RTOL, atol=1.0)


def test_hmm_forward_flow_long_chain_stamp_wrap(fb):
    """T > 65536 wraps the 16-bit packet stamp; uniform transitions give a closed form for logZ."""
    B, T, S = 2, 70001, 128
    g = torch.Generator(device="cuda").manual_seed(17)
    obs = torch.randn(B, T, S, device="cuda", generator=g)
    init = torch.randn(B, S, device="cuda", generator=g).log_softmax(-1)
    A = torch.full((S, S), -float(np.log(S)), device="cuda")
    got = fb.hmm_forward(init, A, obs)
    want = torch.logsumexp(init.double(), -1) + (torch.logsumexp(obs.double(), -1) - np.log(S)).sum(-1)
    rel_close(host(got), host(want), rtol=RTOL)


def test_hmm_forward_flow_large_magnitude_init(fb):
    """Hand-over of a message from an earlier chunk: |init| ~ 1e4-1e5 must not hit the block-exponent range."""
    init, A, obs = _flow_case(5, 40, 300, seed=18)
    init = init - 65000.0 + torch.arange(5)[:, None] * 30000.0
    got = host(fb.hmm_forward(init.cuda(), A.cuda(), obs.cuda()))
    want = R.hmm_forward(R.LOG, init.double().numpy(), A.double().numpy(), obs.double().numpy())
    rel_close(got, want, rtol=RTOL, atol=1.0)


def test_hmm_forward_flow_unaligned_state_count(fb):
    """S not a multiple of 4 takes the scalar observation prefetch."""
    init, A, obs = _flow_case(3, 30, 301, seed=19)
    got = host(fb.hmm_forward(init.cuda(), A.cuda(), obs.cuda()))
    want = R.hmm_forward(R.LOG, init.double().numpy(), A.double().numpy(), obs.double().numpy())
    rel_close(got, want, rtol=RTOL, atol=1.0)


# ------------------------------------------------------------------ two-hop kernel, one chain group, 512 < S <= 1024 (BASELINE config 2's shape):
# 4-byte self-validating words in fragment order + TF32 main product with bf16 side products (hmm_flow_forward<8,2,8,true>)
def _flow4_scenario(name):
    g = torch.Generator().manual_seed(sum(map(ord, name)))
    if name == "extreme_observations":
        return _flow_case(4, 60, 640, seed=24, obs_scale=150.0, obs_shift=-900.0)
    if name == "neg_inf":
        B, T, S = 4, 20, 1001
        init = torch.randn(B, S, generator=g)
        A = torch.randn(S, S, generator=g)
        A[0, 1] = -float("inf")
        A[:, 4] = -float("inf")  # unreachable state
        A[:, 64:72] = -float("inf")  # a whole column group unreachable
        A[:, 512:640] = -float("inf")  # a whole K-slice of the message stays zero
        obs = torch.randn(B, T, S, generator=g)
        obs[1, 3, 2] = -float("inf")
        obs[2, 5, :] = -float("inf")  # impossible observation -> logZ = -inf, not NaN
        init[3, :] = -float("inf")
        init[3, 7] = 0.0  # deterministic start
        return init, A, obs
    if name == "sparse_left_to_right":
        B, T, S = 3, 90, 800
        A = torch.full((S, S), -float("inf"))
        for i in range(S):
            A[i, i] = np.log(0.6)
            A[i, (i + 1) % S] = np.log(0.4)
        init = torch.full((B, S), -float("inf"))
        init[:, 0] = 0.0
        return init, A, torch.randn(B, T, S, generator=g) * 3
    if name == "large_magnitude_init":
        init, A, obs = _flow_case(5, 40, 520, seed=28)
        return init - 65000.0 + torch.arange(5)[:, None] * 30000.0, A, obs
    if name == "unaligned_state_count":
        return _flow_case(3, 30, 901, seed=29)
    if name == "peaked_messages":  # near-deterministic emissions: one or two states carry the message, no averaging over the side products
        init, A, obs = _flow_case(16, 50, 1024, seed=30)
        obs = obs * 0.1
        hot = torch.randint(0, 1024, (16, 50), generator=g)
        obs.scatter_(2, hot[..., None], 40.0)
        return init, A, obs
    raise KeyError(name)


@pytest.mark.parametrize("name", ["extreme_observations", "neg_inf", "sparse_left_to_right", "large_magnitude_init", "unaligned_state_count",
                                  "peaked_messages"])
def test_hmm_forward_flow_four_byte_words(fb, name):
    import os

    init, A, obs = _flow4_scenario(name)
    n0 = fb.launch_count(reset=True)
    got = host(fb.hmm_forward(init.cuda(), A.cuda(), obs.cuda()))
    assert fb.launch_count() == 3, "expected colmax + flow + finalise launches, i.e. the dataflow path"
    want = R.hmm_forward(R.LOG, init.double().numpy(), A.double().numpy(), obs.double().numpy())
    assert not np.isnan(got).any()
    rel_close(got, want, rtol=RTOL, atol=1.0)
    os.environ["FB2_FLOW_PK4"] = "0"  # the 8-byte packet / 3xTF32 form of the same kernel
    try:
        old = host(fb.hmm_forward(init.cuda(), A.cuda(), obs.cuda()))
    finally:
        os.environ.pop("FB2_FLOW_PK4")
    rel_close(got, old, rtol=RTOL, atol=1.0)


def test_hmm_forward_flow_four_byte_words_tight_short_chain(fb):
    """Short chains have small |logZ|: TF32 main product + bf16 side products must hold fp32-class accuracy (2e-5 of 1 + |logZ|
    is the bound the 3xTF32 kernels are held to; measured here per step through the filtered messages as well)."""
    init, A, obs = _flow_case(16, 8, 1024, seed=32)
    got, alpha = fb.hmm_forward(init.cuda(), A.cuda(), obs.cuda(), return_alpha=True)
    want = R.hmm_forward(R.LOG, init.double().numpy(), A.double().numpy(), obs.double().numpy())
    assert np.max(np.abs(host(got) - want)) < 2e-5 * (1 + np.max(np.abs(want)))
    # filtered messages against a float64 recurrence: every entry within 60 nats of its row maximum to 2e-5 absolute in the log
    a = init.double()
    for t in range(8):
        a = torch.logsumexp(a[:, :, None] + A.double()[None], 1) + obs[:, t].double()
        d = (alpha[:, t].double().cpu() - a).abs()
        live = a > a.max(-1, keepdim=True).values - 60.0
        assert d[live].max().item() < 2e-5, (t, d[live].max().item())


def test_hmm_forward_flow_four_byte_words_stamp_wrap(fb):
    """T > 65536 wraps the 16-bit stamp of the exponent words (the data words carry one stamp bit); closed form for uniform transitions."""
    B, T, S = 2, 70001, 640
    g = torch.Generator(device="cuda").manual_seed(37)
    obs = torch.randn(B, T, S, device="cuda", generator=g)
    init = torch.randn(B, S, device="cuda", generator=g).log_softmax(-1)
    A = torch.full((S, S), -float(np.log(S)), device="cuda")
    got = fb.hmm_forward(init, A, obs)
    want = torch.logsumexp(init.double(), -1) + (torch.logsumexp(obs.double(), -1) - np.log(S)).sum(-1)
    rel_close(host(got), host(want), rtol=RTOL)


@pytest.mark.parametrize("B,T,S", [(1, 7, 5), (16, 40, 16), (33, 25, 8), (100, 60, 33), (70, 30, 64), (17, 5, 50), (8200, 12, 16)])
def test_hmm_forward_small_state_space_tensor_core_batches(fb, monkeypatch, B, T, S):
    """hmm_small_forward_tc (16 chains per warp as mma.sync rows, P as split-TF32 B fragments; the default for S <= 16 and B >= 8192,
    forced here): log Z against the float64 oracle incl. -inf transitions, a dead chain and an impossible observation."""
    monkeypatch.setenv("FB2_SMALL_TC", "1")
    g = torch.Generator().manual_seed(B * 1000 + S)
    init = torch.randn(B, S, generator=g).log_softmax(-1)
    A = torch.randn(S, S, generator=g)
    if S > 2:
        A[0, 1] = -float("inf")
        A[:, 2] = -float("inf")  # unreachable state
    A = A.log_softmax(-1)
    obs = torch.randn(B, T, S, generator=g) * 3 - 20.0
    if B > 2 and T > 3:
        obs[1, 3, :] = -float("inf")  # impossible observation -> logZ = -inf, not NaN
        init[2, :] = -float("inf")
        init[2, S - 1] = 0.0  # deterministic start
    n0 = fb.launch_count(reset=True)
    got = host(fb.hmm_forward(init.cuda(), A.cuda(), obs.cuda()))
    want = R.hmm_forward(R.LOG, init.double().numpy(), A.double().numpy(), obs.double().numpy())
    assert not np.isnan(got).any()
    if B > 2 and T > 3:
        assert np.isneginf(got[1])
    rel_close(got, want, rtol=RTOL, atol=1.0)


# ------------------------------------------------------------------ streaming max-plus kernel (Viterbi, S >= 256, shared A)
@pytest.mark.parametrize("B,T,S", [(1, 60, 4096), (5, 40, 1000), (4, 25, 257), (2, 33, 300), (9, 7, 512), (1, 30, 2500), (2, 20, 3000), (3, 9, 2401)])
def test_viterbi_stream_bit_exact(fb, B, T, S):
    g = torch.Generator().manual_seed(22)
    init = torch.randn(B, S, generator=g)
    A = torch.randn(S, S, generator=g)
    obs = torch.randn(B, T, S, generator=g)
    if S == 300:  # unreachable states, an impossible observation and exact ties
        A[:, 7] = -float("inf")
        A[3, :] = -float("inf")
        obs[1, 5, 17] = -float("inf")
        A[:, 40:44] = 0.25
        obs[:, :, 40:44] = 0.5
    best, path = fb.hmm_viterbi(init.cuda(), A.cuda(), obs.cuda())
    wbest, wpath = R.viterbi(init.numpy(), A.numpy(), obs.numpy())  # fp32, same recurrence, first-index ties
    assert np.array_equal(host(best), wbest)
    assert np.array_equal(host(path), wpath)
    z, alpha = fb.hmm_forward(init.cuda(), A.cuda(), obs.cuda(), sum_op="max", return_alpha=True)
    assert np.array_equal(host(z), wbest)


def test_viterbi_stream_matches_generic_kernel(fb):
    import os

    g = torch.Generator().manual_seed(23)
    B, T, S = 3, 50, 640
    init, A, obs = torch.randn(B, S, generator=g), torch.randn(S, S, generator=g), torch.randn(B, T, S, generator=g)
    best, path = fb.hmm_viterbi(init.cuda(), A.cuda(), obs.cuda())
    os.environ["FB2_DISABLE_STREAM"] = "1"
    try:
        gbest, gpath = fb.hmm_viterbi(init.cuda(), A.cuda(), obs.cuda())
    finally:
        os.environ.pop("FB2_DISABLE_STREAM")
    assert np.array_equal(host(best), host(gbest)) and np.array_equal(host(path), host(gpath))


# ------------------------------------------------------------------ tcgen05 log-semiring matmul (log_matmul_tc.cu)
@pytest.mark.parametrize("shape", [(1, 1, 128, 128, 128), (2, 3, 128, 64, 128), (1, 2, 200, 100, 130), (3, 1, 64, 33, 64),
                                   (1, 1, 257, 260, 129), (1, 1, 512, 1000, 384)])
def test_log_matmul_tensor_core_path(fb, shape):
    """Contiguous operands with I, J >= 64 and K >= 32 take the tcgen05 kernel (3 launches: row max, column max, GEMM);
    values must hold the fp32 tolerance, including -inf rows / columns."""
    n0, n1, I, K, J = shape
    g = torch.Generator().manual_seed(31)
    x = torch.randn(n0, n1, I, K, generator=g) * 3
    y = torch.randn(n0, n1, K, J, generator=g) * 3
    x[0, 0, 3, :] = -float("inf")  # a row that contributes nothing
    y[0, 0, :, 5] = -float("inf")
    x[0, 0, 7, ::2] = -float("inf")
    fb.launch_count(reset=True)
    got = fb.semiring_matmul("logaddexp", "add", x.cuda(), y.cuda())
    assert fb.launch_count() == 3
    want = R.pair_product(R.LOG, x.double().numpy(), y.double().numpy())
    assert not np.isnan(host(got)).any()
    rel_close(host(got), want, rtol=RTOL, atol=1.0)


def test_log_matmul_tensor_core_strided_views(fb):
    """Even/odd time views (what one scan level passes, sum_product.py:692-694) stay zero-copy on the tcgen05 path."""
    g = torch.Generator().manual_seed(32)
    trans = torch.randn(2, 9, 128, 128, generator=g).cuda()
    x, y = trans[:, 0:8:2], trans[:, 1:8:2]
    fb.launch_count(reset=True)
    got = fb.semiring_matmul("logaddexp", "add", x, y)
    assert fb.launch_count() == 3
    want = R.pair_product(R.LOG, host(x).astype(np.float64), host(y).astype(np.float64))
    rel_close(host(got), want, rtol=RTOL, atol=1.0)


# ------------------------------------------------------------------ forward-backward on the dataflow kernel
@pytest.mark.parametrize("B,T,S", [(1, 100, 130), (3, 40, 512), (17, 12, 65), (2, 2, 300), (16, 25, 1024), (2, 300, 96)])
def test_hmm_marginals_flow(fb, B, T, S):
    """want_xi=False + homogeneous A + 64 < S <= 1024: both passes run on hmm_flow_forward (the backward pass is the same
    kernel on A^T and time-reversed observations)."""
    g = torch.Generator().manual_seed(41)
    init = torch.randn(B, S, generator=g).log_softmax(-1)
    A = torch.randn(S, S, generator=g).log_softmax(-1)
    obs = torch.randn(B, T, S, generator=g) * 2
    if S == 300:
        A[:, 7] = -float("inf")
        obs[1, 1, 17] = -float("inf")
    fb.launch_count(reset=True)
    logZ, gamma, xi = fb.hmm_marginals(init.cuda(), A.cuda(), obs.cuda(), want_xi=False)
    assert xi is None and fb.launch_count() == 7  # 2 x (colmax, flow, finalise) + gamma
    wZ, wgamma, _ = R.hmm_marginals(init.double().numpy(), A.double().numpy(), obs.double().numpy())
    rel_close(host(logZ), wZ, rtol=RTOL, atol=1.0)
    rel_close(host(gamma), wgamma, rtol=1e-4, atol=1.0)
    np.testing.assert_allclose(np.exp(host(gamma)).sum(-1), 1.0, atol=1e-4)


def test_hmm_marginals_flow_matches_generic(fb):
    import os

    g = torch.Generator().manual_seed(42)
    B, T, S = 2, 60, 200
    init = torch.randn(B, S, generator=g).log_softmax(-1)
    A = torch.randn(S, S, generator=g).log_softmax(-1)
    obs = torch.randn(B, T, S, generator=g)
    z1, g1, _ = fb.hmm_marginals(init.cuda(), A.cuda(), obs.cuda(), want_xi=False)
    os.environ["FB2_DISABLE_FLOW"] = "1"
    try:
        z2, g2, _ = fb.hmm_marginals(init.cuda(), A.cuda(), obs.cuda(), want_xi=False)
    finally:
        os.environ.pop("FB2_DISABLE_FLOW")
    rel_close(host(z1), host(z2), rtol=RTOL, atol=1.0)
    rel_close(host(g1), host(g2), rtol=1e-4, atol=1.0)


def test_host_entry_point_overlapped_copy(fb):
    """fb2_hmm_forward_f32_host with T >= 4096 cuts the observation upload into time chunks and overlaps it with the pass
    (chunk-ready flags raised by stream memory operations); the result must equal the device-resident call, also for
    pageable host memory and for a T that is not a multiple of the chunk count."""
    from funsor_b200 import _lib

    lib = _lib.load()
    for B, T, S, pin in [(3, 5003, 128, True), (16, 4096, 256, False), (4, 4100, 1024, True)]:
        g = torch.Generator().manual_seed(61)
        init = torch.randn(B, S, generator=g).log_softmax(-1)
        A = torch.randn(S, S, generator=g).log_softmax(-1)
        obs = torch.randn(B, T, S, generator=g)
        if pin:
            obs = obs.pin_memory()
        out = torch.empty(B)
        _lib.check(lib.fb2_hmm_forward_f32_host(0, init.data_ptr(), S, A.data_ptr(), 0, 0, obs.data_ptr(), B, T, S, out.data_ptr()))
        want = host(fb.hmm_forward(init.cuda(), A.cuda(), obs.cuda()))
        np.testing.assert_allclose(out.numpy(), want, rtol=1e-6)
    lib.fb2_host_arena_release()


@pytest.mark.parametrize("B,T,S", [(1, 100, 130), (3, 40, 512), (2, 2, 300), (2, 3, 96), (1, 5000, 128), (17, 9, 65)])
def test_hmm_marginals_flow_with_pairwise_sum(fb, B, T, S):
    """xi_sum on the dataflow path: a rescaled outer-product accumulation over time (split-K GEMM) plus the two end steps."""
    g = torch.Generator().manual_seed(43)
    init = torch.randn(B, S, generator=g).log_softmax(-1)
    A = torch.randn(S, S, generator=g).log_softmax(-1)
    obs = torch.randn(B, T, S, generator=g) * 2
    if S == 300:
        A[:, 7] = -float("inf")
        obs[1, 1, 17] = -float("inf")
    logZ, gamma, xi = fb.hmm_marginals(init.cuda(), A.cuda(), obs.cuda(), want_xi=True)
    wZ, wgamma, wxi = R.hmm_marginals(init.double().numpy(), A.double().numpy(), obs.double().numpy())
    rel_close(host(logZ), wZ, rtol=RTOL, atol=1.0)
    rel_close(host(gamma), wgamma, rtol=1e-4, atol=1.0)
    np.testing.assert_allclose(host(xi), wxi, rtol=2e-4, atol=1e-5 * T)
    np.testing.assert_allclose(host(xi).sum((-1, -2)), float(T), rtol=1e-4)  # T pairwise distributions, each sums to one


# ------------------------------------------------------------------ edge cases across the ABI
def test_edge_cases_empty_batch_and_single_step(fb):
    S = 130
    g = torch.Generator().manual_seed(71)
    A = torch.randn(S, S, generator=g).log_softmax(-1).cuda()
    # empty batch: every entry point returns an empty result without launching
    z = fb.hmm_forward(torch.empty(0, S).cuda(), A, torch.empty(0, 5, S).cuda())
    assert z.shape == (0,)
    best, path = fb.hmm_viterbi(torch.empty(0, S).cuda(), A, torch.empty(0, 5, S).cuda())
    assert best.shape == (0,) and path.shape == (0, 6)
    out = fb.sequential_sum_product("logaddexp", "add", torch.empty(0, 4, 8, 8).cuda())
    assert out.shape == (0, 8, 8)
    # a single time step on every path
    init = torch.randn(2, S, generator=g).log_softmax(-1)
    obs = torch.randn(2, 1, S, generator=g)
    for op, sid in (("logaddexp", R.LOG), ("max", R.MAX)):
        got = host(fb.hmm_forward(init.cuda(), A, obs.cuda(), sum_op=op))
        want = R.hmm_forward(sid, init.double().numpy(), host(A).astype(np.float64), obs.double().numpy())
        rel_close(got, want, rtol=RTOL, atol=1.0)
    logZ, gamma, xi = fb.hmm_marginals(init.cuda(), A, obs.cuda())
    wZ, wgamma, wxi = R.hmm_marginals(init.double().numpy(), host(A).astype(np.float64), obs.double().numpy())
    rel_close(host(logZ), wZ, rtol=RTOL, atol=1.0)
    rel_close(host(gamma), wgamma, rtol=1e-4, atol=1.0)
    np.testing.assert_allclose(host(xi), wxi, rtol=1e-4, atol=1e-5)
    best, path = fb.hmm_viterbi(init.cuda(), A, obs.cuda())
    wbest, wpath = R.viterbi(init.numpy(), host(A), obs.numpy())
    assert np.array_equal(host(best), wbest) and np.array_equal(host(path), wpath)


def test_rejects_cpu_tensors_and_wrong_dtype(fb):
    """No CPU fallback: CPU tensors and non-fp32 dtypes raise instead of computing somewhere else."""
    A = torch.zeros(8, 8)
    with pytest.raises(ValueError):
        fb.hmm_forward(torch.zeros(1, 8), A, torch.zeros(1, 3, 8))
    with pytest.raises(ValueError):
        fb.hmm_forward(torch.zeros(1, 8).cuda(), A.cuda().double(), torch.zeros(1, 3, 8).cuda())
    with pytest.raises(NotImplementedError):
        fb.sequential_sum_product("add", "mul", torch.zeros(1, 2, 4, 4).cuda())


# ------------------------------------------------------------------ small state spaces: one warp per chain (hmm_small.cu)
@pytest.mark.parametrize("B,T,S,amode", [(1, 100, 16, "shared"), (40, 25, 16, "shared"), (7, 300, 5, "shared"), (3, 50, 33, "shared"),
                                         (9, 40, 64, "shared"), (5, 30, 12, "batched"), (2, 1, 8, "shared"), (130, 17, 32, "batched")])
def test_hmm_forward_small_state_space(fb, B, T, S, amode):
    g = torch.Generator().manual_seed(81)
    init = torch.randn(B, S, generator=g).log_softmax(-1)
    A = torch.randn(*((S, S) if amode == "shared" else (B, S, S)), generator=g).log_softmax(-1)
    obs = torch.randn(B, T, S, generator=g) * 2
    if S == 12:
        A[:, :, 3] = -float("inf")
        obs[1, 2, 5] = -float("inf")
    fb.launch_count(reset=True)
    got, alpha = fb.hmm_forward(init.cuda(), A.cuda(), obs.cuda(), return_alpha=True)
    assert fb.launch_count() == 1  # a single kernel, no barrier / packet machinery
    want, walpha = R.hmm_forward(R.LOG, init.double().numpy(), A.double().numpy(), obs.double().numpy(), want_alpha=True)
    rel_close(host(got), want, rtol=RTOL, atol=1.0)
    rel_close(host(alpha), walpha, rtol=RTOL, atol=max(1.0, T / 10.0))


def test_hmm_forward_small_long_chain_and_dead_chain(fb):
    B, T, S = 3, 20000, 16
    g = torch.Generator(device="cuda").manual_seed(82)
    obs = torch.randn(B, T, S, device="cuda", generator=g) - 40.0  # log Z ~ -8e5: the integer exponent carries it
    init = torch.randn(B, S, device="cuda", generator=g).log_softmax(-1)
    A = torch.full((S, S), -float(np.log(S)), device="cuda")
    obs[2, 7, :] = -float("inf")  # impossible observation: -inf, not NaN
    got = host(fb.hmm_forward(init, A, obs))
    want = host(torch.logsumexp(init.double(), -1) + (torch.logsumexp(obs.double(), -1) - np.log(S)).sum(-1))
    assert np.isneginf(got[2])
    rel_close(got[:2], want[:2], rtol=RTOL)


@pytest.mark.parametrize("B,T,S,amode", [(2, 50, 16, "shared"), (40, 25, 5, "shared"), (3, 60, 33, "batched"), (130, 17, 64, "shared"), (1, 1, 8, "shared")])
def test_viterbi_small_state_space_bit_exact(fb, B, T, S, amode):
    g = torch.Generator().manual_seed(83)
    init = torch.randn(B, S, generator=g)
    A = torch.randn(*((S, S) if amode == "shared" else (B, S, S)), generator=g)
    obs = torch.randn(B, T, S, generator=g)
    if S == 5:  # ties and unreachable states
        A[:, 2] = -float("inf")
        A[:, 3:] = 0.5
        obs[:, :, 3:] = 0.25
    fb.launch_count(reset=True)
    best, path = fb.hmm_viterbi(init.cuda(), A.cuda(), obs.cuda())
    assert fb.launch_count() == 2  # warp-per-chain forward + backtrace
    wbest, wpath = R.viterbi(init.numpy(), A.numpy(), obs.numpy())
    assert np.array_equal(host(best), wbest)
    assert np.array_equal(host(path), wpath)
    z, delta = fb.hmm_forward(init.cuda(), A.cuda(), obs.cuda(), sum_op="max", return_alpha=True)
    assert np.array_equal(host(z), wbest)


def test_time_sharded_marginals_emulated_on_one_gpu(fb):
    """The time-sharded forward-backward (funsor_b200/distributed.py) with the real kernels, the G ranks emulated as G
    consecutive chunks on one GPU: chunk summaries (fp32 rows + float64 offsets) -> folds -> local passes with boundaries.
    The collective itself is covered by tests/test_distributed_cpu.py (gloo, world size 2)."""
    from funsor_b200 import distributed as D

    g = torch.Generator().manual_seed(91)
    B, T, S, G = 2, 1200, 96, 3
    init = torch.randn(B, S, generator=g).log_softmax(-1).cuda()
    A = torch.randn(S, S, generator=g).log_softmax(-1).cuda()
    obs = (torch.randn(B, T, S, generator=g) - 50.0).cuda()  # log Z ~ -6e4: offsets and re-centring matter
    bounds = [D.shard_bounds(T, r, G) for r in range(G)]
    summaries = torch.stack([D._as_double_summary(fb.hmm_chunk_summary(A, obs[:, lo:hi].contiguous(), return_offsets=True)) for lo, hi in bounds])
    wZ, wgamma, _ = R.hmm_marginals(init.double().cpu().numpy(), A.double().cpu().numpy(), obs.double().cpu().numpy())
    for r, (lo, hi) in enumerate(bounds):
        alpha_in = D.fold_left(init, summaries, r)
        beta_out = D.fold_right(torch.zeros_like(init, dtype=torch.float64), summaries, r)
        a_off, b_off = alpha_in.amax(-1, keepdim=True), beta_out.amax(-1, keepdim=True)
        obs2 = obs[:, lo:hi].clone()
        obs2[:, -1, :] += (beta_out - b_off).float()
        z_loc, gamma, _ = fb.hmm_marginals((alpha_in - a_off).float(), A, obs2, want_xi=False)
        zg = z_loc.double() + a_off[:, 0] + b_off[:, 0]
        rel_close(host(zg), wZ, rtol=RTOL)
        rel_close(host(gamma), wgamma[:, lo:hi], rtol=1e-4, atol=1.0)


@pytest.mark.parametrize("B,T,S", [(2, 37, 96), (1, 200, 130), (3, 5, 65)])
def test_chunk_summary_on_dataflow_kernel(fb, B, T, S):
    """hmm_chunk_summary(return_offsets=True) = the chain product of (A + obs[t]) over the chunk, computed as S chains per
    batch entry started from the identity rows (dataflow kernel), returned as fp32 rows + float64 row offsets."""
    g = torch.Generator().manual_seed(92)
    A = torch.randn(S, S, generator=g).log_softmax(-1)
    obs = torch.randn(B, T, S, generator=g) - 20.0
    trans = A.double().numpy()[None, None] + obs.double().numpy()[:, :, None, :]
    want = R.chain_product(R.LOG, trans)
    for method in ("scan", "tree"):
        rows, off = fb.hmm_chunk_summary(A.cuda(), obs.cuda(), return_offsets=True, method=method)
        got = rows.double() + off[..., None]
        rel_close(host(got), want, rtol=RTOL, atol=1.0)
        assert float(rows.abs().max()) < 200.0  # entries stay O(1)-O(100): the magnitude lives in the offsets
    # the blocked tree with blocks much shorter than the chunk (several re-centred block products chained)
    rows, off = fb.hmm_chunk_summary_tree(A.cuda(), obs.cuda(), block_bytes=7 * B * S * S * 4)
    rel_close(host(rows.double() + off[..., None]), want, rtol=RTOL, atol=1.0)
    plain = fb.hmm_chunk_summary(A.cuda(), obs.cuda())  # generic kernel, absolute fp32
    rel_close(host(plain), want, rtol=RTOL, atol=1.0)


@pytest.mark.gpu
@pytest.mark.parametrize("K", [64, 512, 1024, 2048])
def test_log_matmul_tensor_core_has_no_accumulation_bias(fb, K):
    # tcgen05 accumulation truncates toward zero (one-sided loss ~2^-24 per MMA, -2.2e-8 K before the fix); the kernel splits the
    # accumulators and compensates the expected loss.  Signed mean error against float64 must stay at the fp32 noise level.
    g = torch.Generator(device="cuda").manual_seed(K)
    x = torch.randn(2, 256, K, device="cuda", generator=g)
    y = torch.randn(2, K, 256, device="cuda", generator=g)
    ref = torch.logsumexp(x.double().unsqueeze(-1) + y.double().unsqueeze(-3), -2)
    d = fb.semiring_matmul("logaddexp", "add", x, y).double() - ref
    assert abs(d.mean().item()) < 4e-7, d.mean().item()
    assert d.abs().max().item() < 1e-5


@pytest.mark.gpu
def test_long_chain_log_normaliser_drift_against_float64(fb):
    # a systematic per-step error adds up linearly along a chain: compare both fp32 routes (dataflow recurrence, tensor-core
    # summary tree) with a float64 recurrence.  Bound: 1e-6 relative on logZ (north_star tolerance is 1e-5).
    from funsor_b200 import distributed as fd

    T, S = 16384, 256
    g = torch.Generator(device="cuda").manual_seed(11)
    A = torch.log_softmax(torch.randn(S, S, device="cuda", generator=g), -1)
    init = torch.log_softmax(torch.randn(1, S, device="cuda", generator=g), -1)
    obs = torch.randn(1, T, S, device="cuda", generator=g)
    P, E, v = A.double().exp(), obs[0].double().exp(), init[0].double().exp()
    acc = torch.zeros((), dtype=torch.float64, device="cuda")
    for t in range(T):
        v = (v @ P) * E[t]
        s = v.sum()
        v, acc = v / s, acc + s.log()
    ref = acc.item()
    z = fb.hmm_forward(init, A, obs).item()
    assert abs(z - ref) < 1e-6 * abs(ref) + 4e-3, (z, ref)  # + half an fp32 ulp of the returned value
    summ = fb.hmm_chunk_summary(A, obs, return_offsets=True, method="tree")
    val, _ = fd.hmm_forward_time_sharded(init, A, obs, summary_fn=lambda a, o: summ)
    assert abs(val.item() - ref) < 1e-6 * abs(ref), (val.item(), ref)


@pytest.mark.gpu
@pytest.mark.parametrize("B,T,S,N,amode", [(3, 40, 70, 5, "shared"), (2, 25, 33, 9, "batched"), (2, 12, 6, 4, "timevarying"),
                                            (1, 300, 257, 3, "shared"), (2, 1, 5, 17, "shared")])
def test_posterior_sampling_every_conditional_draw(fb, B, T, S, N, amode):
    # forward-filter backward-sample (recipes.py:16-90 on a chain): every draw of every path is checked against the exact
    # float64 conditional CDF given the path's own next state; forbidden (-inf) states must never be chosen
    rng = np.random.default_rng(B * 1000 + T + S)
    init = torch.log_softmax(torch.tensor(rng.standard_normal((B, S)), dtype=torch.float32), -1)
    shape = {"shared": (S, S), "batched": (B, S, S), "timevarying": (B, T, S, S)}[amode]
    A = torch.tensor(rng.standard_normal(shape), dtype=torch.float32)
    A[..., 0, 1] = -float("inf")
    A = torch.log_softmax(A, -1)
    obs = torch.tensor(rng.standard_normal((B, T, S)), dtype=torch.float32)
    obs[0, T // 2, 2] = -float("inf")
    u = torch.tensor(rng.random((B, N, T + 1)), dtype=torch.float32)
    paths, log_q, logZ = fb.hmm_posterior_sample(init.cuda(), A.cuda(), obs.cuda(), uniforms=u.cuda())
    assert paths.shape == (B, N, T + 1) and paths.dtype == torch.int64
    checked, want_q = R.backward_sample_check(init.numpy(), A.numpy(), obs.numpy(), u.numpy(), paths.cpu().numpy())
    assert checked == B * N * (T + 1)
    rel_close(log_q.cpu().numpy(), want_q, rtol=1e-5, atol=max(1.0, T / 10))
    # same uniforms -> same paths (deterministic), different uniforms -> different paths
    paths2, _, _ = fb.hmm_posterior_sample(init.cuda(), A.cuda(), obs.cuda(), uniforms=u.cuda())
    assert torch.equal(paths, paths2)


@pytest.mark.gpu
def test_posterior_sampling_matches_exact_path_distribution(fb):
    # small chain: enumerate all S^(T+1) paths, compare the empirical path frequencies of 40000 draws with the exact posterior
    import itertools

    S, T, N = 3, 4, 40000
    g = torch.Generator().manual_seed(21)
    init = torch.log_softmax(torch.randn(1, S, generator=g), -1)
    A = torch.log_softmax(torch.randn(S, S, generator=g), -1)
    obs = torch.randn(1, T, S, generator=g)
    gen = torch.Generator(device="cuda").manual_seed(3)
    paths, log_q, logZ = fb.hmm_posterior_sample(init.cuda(), A.cuda(), obs.cuda(), num_samples=N, generator=gen)
    i64, A64, o64 = init.double().numpy()[0], A.double().numpy(), obs.double().numpy()[0]
    allp = list(itertools.product(range(S), repeat=T + 1))
    score = np.array([i64[p[0]] + sum(A64[p[t], p[t + 1]] + o64[t, p[t + 1]] for t in range(T)) for p in allp])
    post = np.exp(score - np.log(np.exp(score).sum()))
    code = (paths[0].cpu().numpy() * (S ** np.arange(T, -1, -1))).sum(-1)
    freq = np.bincount(code, minlength=len(allp)) / N
    sigma = np.sqrt(post * (1 - post) / N)
    assert (np.abs(freq - post) <= 5 * sigma + 1e-4).all(), np.abs(freq - post).max()
    # reported log-probabilities are the exact posterior log-probabilities of the drawn paths
    rel_close(log_q[0].cpu().numpy(), np.log(post[code]), rtol=1e-5, atol=1.0)
    assert abs(logZ.item() - np.log(np.exp(score).sum())) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("direct", ["0", "1"])
@pytest.mark.parametrize("B,T,S", [(16, 40, 1024), (5, 33, 65), (40, 25, 257), (3, 90, 160), (2, 300, 100), (4, 20, 130), (1, 64, 700),
                                   (33, 17, 512), (130, 9, 128)])
def test_hmm_forward_both_dataflow_kernels(fb, B, T, S, direct):
    """The two-hop kernel (DSMEM reduce-scatter + L2 packets) and the single-hop kernel (full-K column owners, fragments through
    L2 with the step stamp in the mantissa LSB) forced at every size, incl. a sparse left-to-right chain, -inf blocks, several
    chain groups per set and more groups than sets."""
    import os

    init, A, obs = _flow_case(B, T, S, seed=33)
    if S == 160:
        A = torch.full((S, S), -float("inf"))
        for i in range(S):
            A[i, i] = np.log(0.6)
            A[i, (i + 1) % S] = np.log(0.4)
        init = torch.full((B, S), -float("inf"))
        init[:, 0] = 0.0
        obs = obs * 3
    if S == 130:
        A[:, 64:72] = -float("inf")
        obs[2, 5, :] = -float("inf")
    os.environ["FB2_FLOW_DIRECT"] = direct
    try:
        got = host(fb.hmm_forward(init.cuda(), A.cuda(), obs.cuda()))
        z2, alpha = fb.hmm_forward(init.cuda(), A.cuda(), obs.cuda(), return_alpha=True)
        zm, gamma, _ = fb.hmm_marginals(init.cuda(), A.cuda(), obs.cuda(), want_xi=False)
    finally:
        os.environ.pop("FB2_FLOW_DIRECT")
    want, walpha = R.hmm_forward(R.LOG, init.double().numpy(), A.double().numpy(), obs.double().numpy(), want_alpha=True)
    assert not np.isnan(got).any()
    rel_close(got, want, rtol=RTOL, atol=1.0)
    rel_close(host(z2), want, rtol=RTOL, atol=1.0)
    rel_close(host(zm), want, rtol=RTOL, atol=1.0)
    # Filtered messages travel in the linear domain with one exponent per block of 8 states, and a contraction aligns the blocks
    # of a K-range to their largest exponent: entries more than ~87 nats (2^-126) below the largest entry of their K-range are
    # flushed to zero (log = -inf) where the reference's log-space arithmetic keeps e.g. -120.  Such states carry < 1e-37 of the
    # mass; compare the log-messages wherever they are within 60 nats of the step's maximum; the rest may only be smaller (up to -inf; +0.7 nat of denormal rounding).
    a = host(alpha)
    near = walpha >= walpha.max(-1, keepdims=True) - 60.0
    rel_close(np.where(near, a, 0.0), np.where(near, walpha, 0.0), rtol=RTOL, atol=max(1.0, T / 10))
    far, afar = walpha[~near], a[~near]
    # flushed entirely (-inf), partially (too small), or sitting in the denormal range of its block (few mantissa bits, +-0.7 nat)
    assert not np.isnan(afar).any() and not np.isposinf(afar).any()
    fin = np.isfinite(far)
    assert (afar[fin] <= far[fin] + 1.0).all() and np.isneginf(afar[~fin]).all()
    if S != 160 and S != 130:
        assert np.abs(np.exp(host(gamma)).sum(-1) - 1).max() < 1e-3


@pytest.mark.gpu
def test_single_hop_kernel_dead_chain_and_exact_zeros(fb):
    """Stamp bits in the mantissa LSB must not leak mass into forbidden states: a chain whose probability is exactly zero stays
    at -inf, next to live chains in the same group."""
    import os

    B, T, S = 3, 50, 96
    init, A, obs = _flow_case(B, T, S, seed=5)
    A[:, :48] = -float("inf")  # states 0..47 unreachable
    obs[1, 20, 48:] = -float("inf")  # chain 1 dies at t = 20
    os.environ["FB2_FLOW_DIRECT"] = "1"
    try:
        got = host(fb.hmm_forward(init.cuda(), A.cuda(), obs.cuda()))
        _, alpha = fb.hmm_forward(init.cuda(), A.cuda(), obs.cuda(), return_alpha=True)
    finally:
        os.environ.pop("FB2_FLOW_DIRECT")
    want = R.hmm_forward(R.LOG, init.double().numpy(), A.double().numpy(), obs.double().numpy())
    assert got[1] == -np.inf and np.isfinite(got[[0, 2]]).all()
    rel_close(got, want, rtol=RTOL, atol=1.0)
    assert (host(alpha)[:, :, :48] == -np.inf).all()


@pytest.mark.gpu
@pytest.mark.parametrize("B,T,S", [(1, 40, 1500), (5, 30, 2048), (2, 20, 4096), (4, 25, 1100), (64, 6, 1030)])
def test_hmm_forward_log_stream_kernel_large_state_space(fb, B, T, S):
    """S > 1024 (beyond the dataflow kernels): linear-domain streaming GEMV with P = exp(A - colmax) formed once.  logZ and the
    filtered messages against the float64 oracle, with forbidden transitions, a -inf observation and a chain that dies."""
    init, A, obs = _flow_case(B, T, S, seed=S + B)
    A[:, 100:140] = -float("inf")
    A = torch.log_softmax(A, -1)
    obs[0, T // 2, 7] = -float("inf")
    if B >= 4:
        obs[3, T // 3, :] = -float("inf")  # chain 3 dies
    z, alpha = fb.hmm_forward(init.cuda(), A.cuda(), obs.cuda(), return_alpha=True)
    z0 = fb.hmm_forward(init.cuda(), A.cuda(), obs.cuda())
    want, walpha = R.hmm_forward(R.LOG, init.double().numpy(), A.double().numpy(), obs.double().numpy(), want_alpha=True)
    rel_close(host(z), want, rtol=RTOL, atol=1.0)
    rel_close(host(z0), want, rtol=RTOL, atol=1.0)
    a = host(alpha)
    assert (a[:, :, 100:140] == -np.inf).all()
    near = walpha >= walpha.max(-1, keepdims=True) - 60.0
    rel_close(np.where(near, a, 0.0), np.where(near, walpha, 0.0), rtol=RTOL, atol=max(1.0, T / 10))
    if B >= 4:
        assert host(z)[3] == -np.inf


@pytest.mark.gpu
def test_hmm_forward_log_stream_matches_generic_kernel(fb):
    import os

    init, A, obs = _flow_case(3, 40, 1200, seed=77, obs_scale=4.0, obs_shift=-20.0)
    got = host(fb.hmm_forward(init.cuda(), A.cuda(), obs.cuda()))
    os.environ["FB2_DISABLE_STREAM"] = "1"
    try:
        ref = host(fb.hmm_forward(init.cuda(), A.cuda(), obs.cuda()))
    finally:
        os.environ.pop("FB2_DISABLE_STREAM")
    rel_close(got, ref, rtol=RTOL, atol=1.0)


@pytest.mark.gpu
@pytest.mark.parametrize("B,T,S", [(1, 30, 1500), (5, 12, 2048), (2, 40, 1100)])
def test_hmm_marginals_large_state_space_on_stream_kernels(fb, B, T, S):
    """Forward-backward for S > 1024: both recursions on the streaming log kernel (backward = forward on A^T over reversed rows)."""
    init, A, obs = _flow_case(B, T, S, seed=S + T)
    A[:, 40:60] = -float("inf")
    A = torch.log_softmax(A, -1)
    obs[0, T // 2, 7] = -float("inf")
    logZ, gamma, xi = fb.hmm_marginals(init.cuda(), A.cuda(), obs.cuda(), want_xi=(S <= 1500))
    wZ, wgamma, wxi = R.hmm_marginals(init.double().numpy(), A.double().numpy(), obs.double().numpy())
    rel_close(host(logZ), wZ, rtol=RTOL, atol=1.0)
    g = host(gamma)
    assert (g[:, :, 40:60] == -np.inf).all() and g[0, T // 2, 7] == -np.inf
    near = wgamma >= -60.0
    with np.errstate(invalid="ignore"):
        assert np.abs(np.where(near, g - wgamma, 0.0)).max() < 2e-4
    assert np.abs(np.exp(g).sum(-1) - 1).max() < 1e-4
    if xi is not None:
        assert np.abs(host(xi) - wxi).max() < 1e-3 * max(1.0, np.abs(wxi).max())


@pytest.mark.gpu
@pytest.mark.parametrize("B,T,S", [(16, 40, 1024), (3, 33, 700), (33, 9, 1000)])
def test_hmm_forward_two_hop_clusters_of_four(fb, B, T, S):
    """The experimental cluster-of-4 / half-K variant of the two-hop kernel (FB2_FLOW_CL4=1) must agree with the oracle too."""
    import os

    init, A, obs = _flow_case(B, T, S, seed=91)
    os.environ["FB2_FLOW_CL4"] = "1"
    os.environ["FB2_FLOW_DIRECT"] = "0"
    try:
        got = host(fb.hmm_forward(init.cuda(), A.cuda(), obs.cuda()))
        z2, alpha = fb.hmm_forward(init.cuda(), A.cuda(), obs.cuda(), return_alpha=True)
    finally:
        os.environ.pop("FB2_FLOW_CL4")
        os.environ.pop("FB2_FLOW_DIRECT")
    want, walpha = R.hmm_forward(R.LOG, init.double().numpy(), A.double().numpy(), obs.double().numpy(), want_alpha=True)
    rel_close(got, want, rtol=RTOL, atol=1.0)
    rel_close(host(alpha), walpha, rtol=RTOL, atol=max(1.0, T / 10))


@pytest.mark.gpu
@pytest.mark.parametrize("B,T,S,amode", [(256, 12, 200, "shared"), (320, 8, 130, "batched"), (200, 6, 260, "batched"), (100, 5, 600, "batched")])
def test_generic_kernel_many_groups_wide_slices(fb, B, T, S, amode):
    """Generic cooperative kernel with 16 chains per group AND more than 16 columns per CTA (S * B > ~38000): the per-column
    finalisation has to cover 16 x 32 (chain, column) pairs with 256 threads.  (A plain `if (tid < R * LW)` left half of the
    chains of every group without results; found by the dispatch-boundary scan, scripts/cliff_time.py.)"""
    rng = np.random.default_rng(B + S)
    init = torch.log_softmax(torch.tensor(rng.standard_normal((B, S)), dtype=torch.float32), -1)
    shape = (S, S) if amode == "shared" else (B, S, S)
    A = torch.log_softmax(torch.tensor(rng.standard_normal(shape), dtype=torch.float32), -1)
    obs = torch.tensor(rng.standard_normal((B, T, S)), dtype=torch.float32)
    best, path = fb.hmm_viterbi(init.cuda(), A.cuda(), obs.cuda())
    wbest, wpath = R.viterbi(init.numpy(), A.numpy(), obs.numpy())
    assert np.array_equal(path.cpu().numpy(), wpath)
    assert np.array_equal(best.cpu().numpy(), wbest.astype(np.float32))
    if amode == "batched":  # per-chain transition matrices: the log-semiring forward runs on the generic kernel as well
        want, walpha = R.hmm_forward(R.LOG, init.double().numpy(), A.double().numpy(), obs.double().numpy(), want_alpha=True)
        z, alpha = fb.hmm_forward(init.cuda(), A.cuda(), obs.cuda(), return_alpha=True)
        rel_close(host(z), want, rtol=RTOL, atol=1.0)
        rel_close(host(alpha), walpha, rtol=RTOL, atol=1.0)
        zm, gamma, _ = fb.hmm_marginals(init.cuda(), A.cuda(), obs.cuda(), want_xi=False)
        rel_close(host(zm), want, rtol=RTOL, atol=1.0)
        assert np.abs(np.exp(host(gamma)).sum(-1) - 1).max() < 1e-3


@pytest.mark.gpu
def test_aborted_dataflow_pass_is_loud_and_recoverable(fb):
    """A dataflow pass that gives up while running (lost co-residency, an input chunk that never landed) poisons its outputs with NaN
    AND raises a sticky per-device status (fb2_async_status): the next dataflow launch refuses to run until the condition has been
    read; the host-buffer entry point, which synchronises anyway, retries on the cooperative kernel.  FB2_FLOW_FORCE_ABORT=1 is the
    test hook that makes a pass behave as if a wait had timed out."""
    import os

    from funsor_b200 import _lib

    g = torch.Generator(device="cuda").manual_seed(3)
    B, T, S = 2, 4096, 128
    A = torch.randn(S, S, device="cuda", generator=g).log_softmax(-1)
    init = torch.randn(B, S, device="cuda", generator=g).log_softmax(-1)
    obs = torch.randn(B, T, S, device="cuda", generator=g)
    good = fb.hmm_forward(init, A, obs)
    torch.cuda.synchronize()
    assert fb.async_status(raise_on_abort=False) == 0
    os.environ["FB2_FLOW_FORCE_ABORT"] = "1"
    try:
        bad = fb.hmm_forward(init, A, obs)
        torch.cuda.synchronize()
        assert bool(torch.isnan(bad).all())
        with pytest.raises(_lib.Fb2Aborted):  # the next dataflow launch on this device reports the pending condition ...
            fb.hmm_forward(init, A, obs)
        assert fb.async_status(raise_on_abort=False) == 0  # ... once; reading it cleared it
        bad = fb.hmm_forward(init, A, obs)
        torch.cuda.synchronize()
        with pytest.raises(_lib.Fb2Aborted):
            fb.async_status()
        # host-buffer entry point: synchronises, sees the abort, recomputes on the cooperative kernel
        lib = _lib.load()
        h_init, h_A, h_obs = init.cpu().contiguous(), A.cpu().contiguous(), obs.cpu().contiguous()
        h_out = torch.empty(B, dtype=torch.float32)
        _lib.check(lib.fb2_hmm_forward_f32_host(0, h_init.data_ptr(), S, h_A.data_ptr(), 0, 0, h_obs.data_ptr(), B, T, S, h_out.data_ptr()))
        assert torch.allclose(h_out, good.cpu(), rtol=1e-5, atol=1e-2), (h_out, good)
    finally:
        del os.environ["FB2_FLOW_FORCE_ABORT"]
        fb.async_status(raise_on_abort=False)
    again = fb.hmm_forward(init, A, obs)
    assert torch.allclose(again, good)


@pytest.mark.gpu
def test_time_sharded_with_impossible_observation(fb):
    """a step at which every state is impossible (all -inf): the reference's scan returns -inf, the re-centring of the time-sharded
    path must not turn it into NaN (summary tree, folds, boundary offsets)"""
    from funsor_b200 import distributed as fd

    g = torch.Generator(device="cuda").manual_seed(4)
    T, S = 64, 128
    A = torch.randn(S, S, device="cuda", generator=g).log_softmax(-1)
    init = torch.randn(1, S, device="cuda", generator=g).log_softmax(-1)
    obs = torch.randn(1, T, S, device="cuda", generator=g)
    obs[0, 17, :] = -float("inf")
    for method in ("tree", "scan"):
        summ = fb.hmm_chunk_summary(A, obs, return_offsets=True, method=method)
        val, _ = fd.hmm_forward_time_sharded(init, A, obs, summary_fn=lambda a, o: summ)
        assert float(val) == -float("inf"), (method, float(val))
