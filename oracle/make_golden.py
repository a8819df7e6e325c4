"""Test infrastructure: generate tests/golden/*.npz by running the UNMODIFIED reference
(funsor 0.4.8 from /root/reference, imported through oracle/ref_boot.py) on seeded
inputs.  Run in the build container only:

    python -m oracle.make_golden

The fixtures are committed; the GPU box never runs this script.  Every array is stored
with its inputs so that tests can replay the same inputs through the oracle
(tests/test_oracle_golden.py, CPU) and through the CUDA path (tests/test_*_gpu.py).
"""
import importlib.util
import os
import sys
from collections import OrderedDict

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _np(t):
    return t.detach().cpu().numpy()


def main():
    sys.path.insert(0, ROOT)
    from oracle import ref_boot

    funsor = ref_boot.boot("float64")
    import torch

    import funsor.ops as ops
    from funsor.adjoint import AdjointTape
    from funsor.cnf import Contraction
    from funsor.domains import Bint, Reals
    from funsor.gaussian import Gaussian
    from funsor.sum_product import (
        MarkovProduct,
        mixed_sequential_sum_product,
        naive_sequential_sum_product,
        sequential_sum_product,
    )
    from funsor.tensor import Tensor
    from funsor.terms import Variable
    from funsor.testing import random_gaussian

    os.makedirs(GOLDEN, exist_ok=True)
    SEMI = {"log": (ops.logaddexp, ops.add), "max": (ops.max, ops.add)}

    def tfun(data, names):
        return Tensor(data, OrderedDict((n, Bint[s]) for n, s in zip(names, data.shape)))

    def to_bij(f, names):
        return _np(f.align(tuple(names)).data)

    # ------------------------------------------------------------------ pair products (a6 / a7)
    out = {}
    torch.manual_seed(11)
    for name, (B, T, S) in {"p0": (1, 1, 2), "p1": (2, 3, 3), "p2": (3, 5, 16), "p3": (2, 2, 33)}.items():
        x = torch.randn(B, T, S, S)
        y = torch.randn(B, T, S, S)
        if name == "p1":  # -inf handling incl. a fully masked row / column (A.2)
            x[0, 0, 1, :] = -float("inf")
            y[0, 0, :, 2] = -float("inf")
            x[1, 1, 0, 1] = -float("inf")
            y[1, 2, 1, 1] = -float("inf")
        fx = tfun(x, ["b", "t", "i", "k"])
        fy = tfun(y, ["b", "t", "k", "j"])
        out[name + "/x"] = _np(x)
        out[name + "/y"] = _np(y)
        for sname, (sop, pop) in SEMI.items():
            z = Contraction(sop, pop, frozenset({Variable("k", Bint[S])}), fx, fy)
            assert isinstance(z, Tensor), type(z)
            out[name + "/" + sname] = to_bij(z, ["b", "t", "i", "j"])
    # non-square contraction (I,K,J differ) through the same rules
    x = torch.randn(2, 4, 5)
    y = torch.randn(2, 5, 3)
    out["rect/x"], out["rect/y"] = _np(x), _np(y)
    for sname, (sop, pop) in SEMI.items():
        z = Contraction(sop, pop, frozenset({Variable("k", Bint[5])}), tfun(x, ["b", "i", "k"]), tfun(y, ["b", "k", "j"]))
        out["rect/" + sname] = to_bij(z, ["b", "i", "j"])
    np.savez_compressed(os.path.join(GOLDEN, "pair_product.npz"), **out)

    # ------------------------------------------------------------------ scans (a1-a4)
    out = {}
    torch.manual_seed(12)
    cases = [(B, T, S) for B in (1, 3) for T in (1, 2, 3, 4, 5, 7, 8, 11, 12) for S in (2, 3)]
    cases += [(2, 37, 5), (1, 100, 16), (2, 64, 8)]
    for B, T, S in cases:
        key = "B{}_T{}_S{}".format(B, T, S)
        trans = torch.randn(B, T, S, S)
        if (B, T, S) == (3, 7, 3):
            trans[1, 2, :, 1] = -float("inf")
            trans[2, 4, 0, :] = -float("inf")
        f = tfun(trans, ["b", "time", "prev", "curr"])
        time = Variable("time", Bint[T])
        out[key + "/trans"] = _np(trans)
        for sname, (sop, pop) in SEMI.items():
            r = sequential_sum_product(sop, pop, f, time, {"prev": "curr"})
            out[key + "/" + sname + "_seq"] = to_bij(r, ["b", "prev", "curr"])
            if T <= 12:
                r = naive_sequential_sum_product(sop, pop, f, time, {"prev": "curr"})
                out[key + "/" + sname + "_naive"] = to_bij(r, ["b", "prev", "curr"])
                for seg in (2, 3):
                    r = mixed_sequential_sum_product(sop, pop, f, time, {"prev": "curr"}, num_segments=seg)
                    out[key + "/" + sname + "_mixed{}".format(seg)] = to_bij(r, ["b", "prev", "curr"])
                r = MarkovProduct(sop, pop, f, time, {"prev": "curr"})
                out[key + "/" + sname + "_markov"] = to_bij(r, ["b", "prev", "curr"])
    np.savez_compressed(os.path.join(GOLDEN, "chain_product.npz"), **out)

    # ------------------------------------------------------------------ HMM log_prob (hmm.py:129-135) incl. C1
    out = {}
    torch.manual_seed(0)  # C1 exactly as SURVEY 8(d): seed 0, randn(1,100,16,16), randn(1,16), fp32
    torch.set_default_dtype(torch.float32)
    trans = torch.randn(1, 100, 16, 16)
    init = torch.randn(1, 16)
    f = tfun(trans, ["b", "time", "state", "state(time=1)"])
    r = sequential_sum_product(ops.logaddexp, ops.add, f, Variable("time", Bint[100]), {"state": "state(time=1)"})
    r = tfun(init, ["b", "state"]) + r.reduce(ops.logaddexp, "state(time=1)")
    r = r.reduce(ops.logaddexp, "state")
    out["c1/trans"], out["c1/init"], out["c1/logZ_fp32"] = _np(trans), _np(init), _np(r.data)
    torch.set_default_dtype(torch.float64)
    f = tfun(trans.double(), ["b", "time", "state", "state(time=1)"])
    r = sequential_sum_product(ops.logaddexp, ops.add, f, Variable("time", Bint[100]), {"state": "state(time=1)"})
    r = (tfun(init.double(), ["b", "state"]) + r.reduce(ops.logaddexp, "state(time=1)")).reduce(ops.logaddexp, "state")
    out["c1/logZ_fp64"] = _np(r.data)

    torch.manual_seed(13)
    for name, (B, T, S, amode) in {
        "h0": (2, 9, 4, "shared"),
        "h1": (3, 16, 7, "batched"),
        "h2": (2, 6, 3, "timevarying"),
        "h3": (1, 50, 32, "shared"),
    }.items():
        init = torch.randn(B, S).log_softmax(-1)
        if amode == "shared":
            A = torch.randn(S, S).log_softmax(-1)
            fA = tfun(A, ["state", "state(time=1)"])
        elif amode == "batched":
            A = torch.randn(B, S, S).log_softmax(-1)
            fA = tfun(A, ["b", "state", "state(time=1)"])
        else:
            A = torch.randn(B, T, S, S).log_softmax(-1)
            fA = tfun(A, ["b", "time", "state", "state(time=1)"])
        obs = torch.randn(B, T, S)
        if name == "h1":
            A[..., 0, 1] = -float("inf")  # forbidden transition (test_hmm.py:387-396 style)
            obs[0, 3, 2] = -float("inf")
        fobs = tfun(obs, ["b", "time", "state(time=1)"])
        finit = tfun(init, ["b", "state"])
        out[name + "/init"], out[name + "/A"], out[name + "/obs"] = _np(init), _np(A), _np(obs)
        for sname, (sop, pop) in SEMI.items():
            with AdjointTape() as tape:
                result = fA + fobs
                result = sequential_sum_product(sop, pop, result, Variable("time", Bint[T]), {"state": "state(time=1)"})
                result = finit + result.reduce(sop, "state(time=1)")
                result = result.reduce(sop, "state")
            out[name + "/" + sname + "_Z"] = to_bij(result, ["b"])
    np.savez_compressed(os.path.join(GOLDEN, "hmm_logprob.npz"), **out)

    # ------------------------------------------------------------------ adjoint (a11), test_adjoint.py:187-261 body
    out = {}
    torch.manual_seed(14)
    for B, T, S in [(1, 3, 2), (2, 7, 3), (2, 12, 3), (1, 9, 5)]:
        key = "B{}_T{}_S{}".format(B, T, S)
        trans = torch.randn(B, T, S, S)
        f = tfun(trans, ["b", "time", "prev", "curr"])
        out[key + "/trans"] = _np(trans)
        for sname, (sop, pop) in SEMI.items():
            with AdjointTape() as tape:
                z = sequential_sum_product(sop, pop, f, Variable("time", Bint[T]), {"prev": "curr"})
                z = z.reduce(sop, frozenset({"prev", "curr"}))
            adj = tape.adjoint(sop, pop, z, (f,), batch_vars=frozenset({Variable("b", Bint[B])}))[f]
            out[key + "/" + sname + "_Z"] = to_bij(z, ["b"])
            out[key + "/" + sname + "_adj"] = to_bij(adj, ["b", "time", "prev", "curr"])
    np.savez_compressed(os.path.join(GOLDEN, "adjoint.npz"), **out)

    # ------------------------------------------------------------------ Gaussian scans (a10), test_sum_product.py:2427-2493
    out = {}
    torch.manual_seed(15)

    def info_of(result, names, D):
        """(Lam, eta, c) of a `Tensor + Gaussian` funsor over real inputs `names`."""
        terms = result.terms if isinstance(result, Contraction) else (result,)
        g = [t for t in terms if isinstance(t, Gaussian)]
        assert len(g) == 1
        g = g[0]
        reals = [k for k, d in g.inputs.items() if d.dtype == "real"]
        assert sorted(reals) == sorted(names), list(g.inputs)
        perm = np.concatenate([np.arange(D) + D * reals.index(n) for n in names])  # reorder real blocks
        zero = {n: funsor.to_funsor(torch.zeros(D), Reals[D]) for n in names}
        c = result(**zero)
        assert isinstance(c, Tensor)
        bnames = [k for k, d in g.inputs.items() if d.dtype != "real"]
        Lam = _np(g._precision)[..., perm, :][..., :, perm]
        eta = _np(g._info_vec)[..., perm]
        return Lam, eta, _np(c.align(tuple(bnames)).data), bnames

    for B, T, D in [(1, 1, 1), (2, 2, 1), (2, 3, 2), (1, 5, 3), (3, 8, 2), (2, 11, 3), (1, 12, 4)]:
        key = "B{}_T{}_D{}".format(B, T, D)
        inputs = OrderedDict([("b", Bint[B]), ("time", Bint[T]), ("prev", Reals[D]), ("curr", Reals[D])])
        g = random_gaussian(inputs)
        out[key + "/w"], out[key + "/P"] = _np(g.white_vec), _np(g.prec_sqrt)
        r = sequential_sum_product(ops.logaddexp, ops.add, g, Variable("time", Bint[T]), {"prev": "curr"})
        Lam, eta, c, bn = info_of(r, ["prev", "curr"], D)
        assert bn == ["b"], bn
        out[key + "/Lam"], out[key + "/eta"], out[key + "/c"] = Lam, eta, c
        if T <= 8:
            r2 = naive_sequential_sum_product(ops.logaddexp, ops.add, g, Variable("time", Bint[T]), {"prev": "curr"})
            Lam2, eta2, c2, _ = info_of(r2, ["prev", "curr"], D)
            out[key + "/Lam_naive"], out[key + "/eta_naive"], out[key + "/c_naive"] = Lam2, eta2, c2
    # one explicit pair contraction with rank-deficient (Kalman-like) operands
    np.savez_compressed(os.path.join(GOLDEN, "gaussian_chain.npz"), **out)

    # ------------------------------------------------------------------ Kalman / GaussianHMM element + log_prob
    out = {}
    torch.manual_seed(16)
    spec = importlib.util.spec_from_file_location("_ref_convert", os.path.join(ref_boot.REFERENCE_ROOT, "funsor/pyro/convert.py"))
    conv = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(conv)
    MVN = torch.distributions.MultivariateNormal

    def spd_tril(*shape):
        a = torch.randn(*shape, shape[-1])
        return torch.linalg.cholesky(a @ a.transpose(-1, -2) / shape[-1] + 0.5 * torch.eye(shape[-1]))

    for name, (B, T, D, O) in {"k0": (2, 6, 2, 1), "k1": (2, 9, 4, 2), "k2": (1, 16, 6, 3)}.items():
        F = 0.9 * torch.linalg.qr(torch.randn(B, D, D))[0]
        H = torch.randn(B, D, O)
        q_loc, r_loc = 0.1 * torch.randn(B, D), 0.1 * torch.randn(B, O)
        q_tril, r_tril = spd_tril(B, D), spd_tril(B, O)
        m0, p0_tril = torch.randn(B, D), spd_tril(B, D)
        y = torch.randn(B, T, O)
        for k, v in dict(F=F, H=H, q_loc=q_loc, r_loc=r_loc, q_tril=q_tril, r_tril=r_tril, m0=m0, p0_tril=p0_tril, y=y).items():
            out[name + "/" + k] = _np(v)
        # hmm.py:258-270 with time-homogeneous parameters expanded over a `time` batch dim
        ex = lambda t: t.unsqueeze(1).expand((B, T) + t.shape[1:])  # noqa: E731
        trans = conv.matrix_and_mvn_to_funsor(ex(F), MVN(ex(q_loc), scale_tril=ex(q_tril)), ("time",), "state", "state(time=1)")
        obs = conv.matrix_and_mvn_to_funsor(ex(H), MVN(ex(r_loc), scale_tril=ex(r_tril)), ("time",), "state(time=1)", "value")
        value = Tensor(y, OrderedDict([("_pyro_dim_-1", Bint[B]), ("time", Bint[T])]))
        elem = trans + obs(value=value)
        chain = sequential_sum_product(ops.logaddexp, ops.add, elem, Variable("time", Bint[T]), {"state": "state(time=1)"})
        Lam, eta, c, bn = info_of(chain, ["state", "state(time=1)"], D)
        out[name + "/chain_Lam"], out[name + "/chain_eta"], out[name + "/chain_c"] = Lam, eta, c
        # initial N(m0, L L') as a funsor; dist_to_funsor needs pyro, so build the same density directly
        P0 = torch.linalg.inv(p0_tril).transpose(-1, -2)
        w0 = (m0.unsqueeze(-2) @ P0).squeeze(-2)
        s0 = -0.5 * D * np.log(2 * np.pi) - p0_tril.diagonal(dim1=-1, dim2=-2).log().sum(-1)
        binp = OrderedDict([("_pyro_dim_-1", Bint[B])])
        ginp = binp.copy()
        ginp["state"] = Reals[D]
        init = Gaussian(white_vec=w0, prec_sqrt=P0, inputs=ginp) + Tensor(s0, binp)
        lp = (init + chain.reduce(ops.logaddexp, "state(time=1)")).reduce(ops.logaddexp, "state")
        out[name + "/logprob"] = _np(lp.data)
    np.savez_compressed(os.path.join(GOLDEN, "kalman.npz"), **out)
    # ------------------------------------------------------------------ the chain distributions of funsor/pyro/hmm.py
    # The classes themselves subclass pyro's TorchDistribution and convert torch distributions through
    # funsor.torch.distributions (both need pyro, absent here), so their log_prob EXPRESSIONS (hmm.py:117-138, 258-274,
    # 339-378) are evaluated below with the reference's own funsor machinery and converters (tensor_to_funsor,
    # funsor_to_tensor, matrix_and_mvn_to_funsor); the one converter that needs pyro (dist_to_funsor / mvn_to_funsor for
    # an MVN) is restated by `mvn_funsor`.
    out = {}
    torch.manual_seed(17)
    from funsor.util import broadcast_shape

    def rand_mvn(shape, dim):
        loc = torch.randn(shape + (dim,))
        a = torch.randn(shape + (dim, 2 * dim))
        return MVN(loc, scale_tril=torch.linalg.cholesky(a @ a.transpose(-1, -2)))

    def mvn_funsor(d, event_dims, real_inputs):
        n = d.event_shape[0]
        Psq = torch.linalg.inv(d.scale_tril).transpose(-1, -2)
        wv = (d.loc[..., None, :] @ Psq)[..., 0, :]
        lp = -0.5 * n * np.log(2 * np.pi) - d.scale_tril.diagonal(dim1=-1, dim2=-2).log().sum(-1)
        lp = conv.tensor_to_funsor(lp, event_dims)
        wv_f = conv.tensor_to_funsor(wv.expand(Psq.shape[:-2] + (-1,)), event_dims, 1)
        Ps_f = conv.tensor_to_funsor(Psq, event_dims, 2)
        inputs = Ps_f.inputs.copy()
        for k, v in wv_f.inputs.items():
            inputs.setdefault(k, v)
        inputs.update(real_inputs)
        bshape = tuple(v.size for v in inputs.values() if v.dtype != "real")
        names = [k for k, v in inputs.items() if v.dtype != "real"]
        wv_d = wv_f.align(tuple(k for k in names if k in wv_f.inputs)).data
        Ps_d = Ps_f.align(tuple(k for k in names if k in Ps_f.inputs)).data
        assert wv_d.shape[:-1] == bshape and Ps_d.shape[:-2] == bshape, (wv_d.shape, Ps_d.shape, bshape)
        return Gaussian(white_vec=wv_d, prec_sqrt=Ps_d, inputs=inputs) + lp

    S = 3
    DISCRETE = {"d0": ((), (1,), ()), "d1": ((), (), (7,)), "d2": ((), (7,), (5, 7)), "d3": ((5,), (5, 7), (7,)),
                "d4": ((4, 1, 1), (3, 1, 7), (2, 7)), "d5": ((2,), (6,), (2, 6))}
    for name, (ish, tsh, osh) in DISCRETE.items():
        init_logits = torch.randn(ish + (S,))
        trans_logits = torch.randn(tsh + (S, S))
        loc, scale = torch.randn(osh + (S,)), torch.randn(osh + (S,)).exp()
        obs_dist = torch.distributions.Normal(loc, scale)
        shape = broadcast_shape(ish + (1,), tsh, osh)
        batch_shape, T = shape[:-1], shape[-1]
        data = obs_dist.expand(shape + (S,)).sample()[..., 0]
        # hmm.py:91-99
        il = init_logits - init_logits.logsumexp(-1, True)
        tl = trans_logits - trans_logits.logsumexp(-1, True)
        init = conv.tensor_to_funsor(il, ("state",))
        trans = conv.tensor_to_funsor(tl, ("time", "state", "state(time=1)"))
        obs = conv.tensor_to_funsor(obs_dist.log_prob(data.unsqueeze(-1)), ("time", "state(time=1)"))  # = self._obs(value=value)
        # hmm.py:128-137
        result = trans + obs
        result = sequential_sum_product(ops.logaddexp, ops.add, result, Variable("time", Bint[T]), {"state": "state(time=1)"})
        result = init + result.reduce(ops.logaddexp, "state(time=1)")
        result = result.reduce(ops.logaddexp, "state")
        lp = conv.funsor_to_tensor(result, ndims=max(len(batch_shape), data.dim() - 1))
        for k, v in dict(init_logits=init_logits, trans_logits=trans_logits, loc=loc, scale=scale, data=data, log_prob=lp).items():
            out["discrete/" + name + "/" + k] = _np(v)

    GH = {"g0": ((), (), (), (), (), 4), "g1": ((), (6,), (), (), (), 6), "g2": ((), (), (), (), (6,), 6),
          "g3": ((5,), (6,), (), (), (), 6), "g4": ((), (), (5, 1), (6,), (), 6), "g5": ((5,), (5, 6), (5, 6), (5, 6), (5, 6), 6)}
    for name, (ish, tmsh, tvsh, omsh, ovsh, Tdata) in GH.items():
        D, O = (2, 3) if name != "g5" else (3, 2)
        init_dist, trans_dist, obs_dist = rand_mvn(ish, D), rand_mvn(tvsh, D), rand_mvn(ovsh, O)
        trans_mat, obs_mat = 0.7 * torch.randn(tmsh + (D, D)), torch.randn(omsh + (D, O))
        shape = broadcast_shape(ish + (1,), tmsh, tvsh, omsh, ovsh)
        T = shape[-1]
        if T == 1:  # homogeneous: the reference's event_shape has num_steps = 1 (hmm.py:181-186); use T = 1 data here
            Tdata = 1
        data = obs_dist.expand(shape).sample()
        assert data.shape[-2] == Tdata
        init = mvn_funsor(init_dist, (), OrderedDict(state=Reals[D]))
        trans = conv.matrix_and_mvn_to_funsor(trans_mat, trans_dist, ("time",), "state", "state(time=1)")
        obs = conv.matrix_and_mvn_to_funsor(obs_mat, obs_dist, ("time",), "state(time=1)", "value")
        value = conv.tensor_to_funsor(data, ("time",), 1)
        result = trans + obs(value=value)
        result = sequential_sum_product(ops.logaddexp, ops.add, result, Variable("time", Bint[T]), {"state": "state(time=1)"})
        result = init + result.reduce(ops.logaddexp, "state(time=1)")
        result = result.reduce(ops.logaddexp, "state")
        lp = conv.funsor_to_tensor(result, ndims=len(shape) - 1)
        for k, v in dict(init_loc=init_dist.loc, init_tril=init_dist.scale_tril, trans_mat=trans_mat, trans_loc=trans_dist.loc,
                         trans_tril=trans_dist.scale_tril, obs_mat=obs_mat, obs_loc=obs_dist.loc, obs_tril=obs_dist.scale_tril,
                         data=data, log_prob=lp).items():
            out["ghmm/" + name + "/" + k] = _np(v)

    MRF = {"m0": ((), (1,), ()), "m1": ((), (7,), ()), "m2": ((), (), (7,)), "m3": ((5,), (7,), (5, 7)), "m4": ((2, 1, 1), (3, 1, 7), (2, 7))}
    for name, (ish, tsh, osh) in MRF.items():
        D, O = 2, 2
        init_dist, trans_dist, obs_dist = rand_mvn(ish, D), rand_mvn(tsh, 2 * D), rand_mvn(osh, D + O)
        shape = broadcast_shape(ish + (1,), tsh, osh)
        T = shape[-1]
        data = obs_dist.expand(shape).sample()[..., D:]
        init = mvn_funsor(init_dist, (), OrderedDict(state=Reals[D]))
        trans = mvn_funsor(trans_dist, ("time",), OrderedDict([("state", Reals[D]), ("state(time=1)", Reals[D])]))
        obs = mvn_funsor(obs_dist, ("time",), OrderedDict([("state(time=1)", Reals[D]), ("value", Reals[O])]))
        time = Variable("time", Bint[T])
        value = conv.tensor_to_funsor(data, ("time",), 1)
        both = frozenset({"state", "state(time=1)"})
        logp_oh = trans + obs(value=value)
        logp_oh = sequential_sum_product(ops.logaddexp, ops.add, logp_oh, time, {"state": "state(time=1)"})
        logp_oh = (logp_oh + init).reduce(ops.logaddexp, both)
        logp_h = trans + obs.reduce(ops.logaddexp, "value")
        logp_h = sequential_sum_product(ops.logaddexp, ops.add, logp_h, time, {"state": "state(time=1)"})
        logp_h = (logp_h + init).reduce(ops.logaddexp, both)
        lp = conv.funsor_to_tensor(logp_oh - logp_h, ndims=len(shape) - 1)
        for k, v in dict(init_loc=init_dist.loc, init_tril=init_dist.scale_tril, trans_loc=trans_dist.loc, trans_tril=trans_dist.scale_tril,
                         obs_loc=obs_dist.loc, obs_tril=obs_dist.scale_tril, data=data, log_prob=lp).items():
            out["mrf/" + name + "/" + k] = _np(v)
    np.savez_compressed(os.path.join(GOLDEN, "distributions.npz"), **out)
    moment_matching_golden()
    for fn in sorted(os.listdir(GOLDEN)):
        print(fn, os.path.getsize(os.path.join(GOLDEN, fn)))


def moment_matching_golden():
    """tests/golden/moment_matching.npz: the `moment_matching` interpretation of the reference (joint.py:102-150) collapsing
    Gaussian mixtures, and the SwitchingLinearHMM log_prob expression (pyro/hmm.py:527-551) under it.
    `python -m oracle.make_golden moment` writes only this file."""
    sys.path.insert(0, ROOT)
    from oracle import ref_boot

    funsor = ref_boot.boot("float64")
    import torch

    import funsor.ops as ops
    from funsor.cnf import Contraction
    from funsor.domains import Bint, Reals
    from funsor.gaussian import Gaussian
    from funsor.interpretations import moment_matching
    from funsor.tensor import Tensor
    from funsor.terms import Variable
    from funsor.testing import random_gaussian

    out = {}
    torch.manual_seed(18)
    for name, (B, K, n) in {"m0": (1, 2, 1), "m1": (3, 4, 2), "m2": (2, 3, 6), "m3": (2, 5, 16), "m4": (1, 7, 64)}.items():
        g = random_gaussian(OrderedDict([("b", Bint[B]), ("k", Bint[K]), ("x", Reals[n])]))
        logits = torch.randn(B, K)
        if name == "m1":
            logits[0, 1] = -float("inf")  # a component without mass
        d = Tensor(logits, OrderedDict([("b", Bint[B]), ("k", Bint[K])]))
        with moment_matching:
            r = Contraction(ops.logaddexp, ops.add, frozenset({Variable("k", Bint[K])}), d, g)
        (gr,) = [t for t in r.terms if isinstance(t, Gaussian)]
        assert list(gr.inputs) == ["b", "x"], gr.inputs
        c = r(x=funsor.to_funsor(torch.zeros(n), Reals[n]))
        out[name + "/logits"], out[name + "/w"], out[name + "/P"] = _np(logits), _np(g.white_vec), _np(g.prec_sqrt)
        out[name + "/Lam"], out[name + "/eta"], out[name + "/c"] = _np(gr._precision), _np(gr._info_vec), _np(c.data)
    np.savez_compressed(os.path.join(GOLDEN, "moment_matching.npz"), **out)


if __name__ == "__main__":
    if "moment" in sys.argv[1:]:
        moment_matching_golden()
    else:
        main()
