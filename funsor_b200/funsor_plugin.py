"""Eager patterns that route funsor's Markov-chain sum-product path to the sm_100a kernels.

This is the drop-in boundary on the reference side.  `register()` adds these
more-specific rules to funsor's `eager` interpretation for the torch backend:

  Contraction[Logaddexp, Add, frozenset, Tensor, Tensor]   replaces funsor/cnf.py:335-340
  Contraction[Max,       Add, frozenset, Tensor, Tensor]   beats the generic funsor/cnf.py:312-324
  MarkovProduct[Logaddexp|Max, Add, Tensor, Variable, frozenset, frozenset]
                                                           beats funsor/sum_product.py:1049-1064
  MarkovProduct[Logaddexp, Add, Gaussian | GaussianMixture, Variable, frozenset, frozenset]
                                                           the Gaussian scan (gaussian.py:841-1132 per level) as ONE
                                                           information-form chain kernel; the result is converted back
                                                           to the reference's (white_vec, prec_sqrt) + Tensor form
  moment_matching: Contraction[Logaddexp, Add, frozenset, Tensor, Gaussian]
                                                           the mixture collapse of funsor/joint.py:102-150 (SwitchingLinearHMM,
                                                           pyro/hmm.py:527-551) as ONE kernel per contraction
  Contraction[Logaddexp, Add, frozenset, Gaussian | GaussianMixture, Gaussian | GaussianMixture]
                                                           one level of that scan when sequential_sum_product is called
                                                           directly (pyro/hmm.py:269): replaces cnf.py:385-389
                                                           ((x + y).reduce) by sqrt->info, the Schur-complement pair kernel
                                                           and a pivoted-Cholesky info->sqrt

Gradients: the Tensor rules stay differentiable (ops.semiring_matmul / sequential_sum_product switch to the autograd
Functions of autograd.py when an operand requires grad, examples/discrete_hmm.py:76-78); the Gaussian rules delegate to
the reference whenever an operand requires grad.

A rule takes the CUDA path only when its operands are float32 CUDA tensors with scalar output shape and
the contraction has matrix-product structure; in every other case it calls the handler that was
registered BEFORE ours (captured with eager.dispatch, funsor/registry.py:107-108).  It never returns
None: that would leave a lazy term instead of trying the next rule (probe).

The named-dimension bookkeeping is kept in two pure functions (`contract_named`, `markov_named`) that
work on raw tensors and name lists and take the kernel entry point as an argument.  That is what lets
the path be tested on both sides of the gap between the two machines: against the unmodified reference
with a torch stand-in for the kernel (CPU container, tests/test_funsor_plugin.py) and against torch with
the real kernels (GPU box, tests/test_gpu_plugin.py).

funsor itself is not a dependency of this package: nothing here imports it until register() is called.
"""
from collections import OrderedDict

import torch

_STATE = {"registered": False, "calls": {"contraction_fast": 0, "contraction_delegated": 0,
                                         "markov_fast": 0, "markov_delegated": 0, "gaussian_homogeneous": 0,
                                         "gaussian_contraction_fast": 0, "gaussian_contraction_delegated": 0,
                                         "moment_matching_fast": 0, "moment_matching_delegated": 0}}


def stats(reset=False):
    out = dict(_STATE["calls"])
    if reset:
        for k in _STATE["calls"]:
            _STATE["calls"][k] = 0
    return out


# ------------------------------------------------------------------------------------------------
# pure named-tensor logic (no funsor types)
def contract_named(x, x_names, y, y_names, reduced, semiring, matmul):
    """out[names] = (+)_{reduced} x[x_names] (x) y[y_names] as one batched semiring matmul.

    x, y: tensors whose dims are exactly the named dims (funsor Tensor.data with scalar output,
    funsor/tensor.py:138-150).  reduced: set of names present in BOTH operands.  Returns
    (data, names) with names in funsor's output order: first-seen over x then y, minus reduced
    (funsor/cnf.py:348-375).  Returns None when the contraction is not a matrix product
    (a reduced name missing from one side)."""
    x_names, y_names = list(x_names), list(y_names)
    reduced = set(reduced)
    if not reduced or not all(r in x_names or r in y_names for r in reduced):
        return None
    # a reduced name that only ONE operand carries (sarkka_bilmes_product's block factors, sum_product.py:905-921) is summed out of
    # that operand first -- a unary reduction over its own elements, exact for max and within rounding for logaddexp
    for own, other in ((x_names, y_names), (y_names, x_names)):
        dims = [i for i, n in enumerate(own) if n in reduced and n not in other]
        if dims:
            t = x if own is x_names else y
            t = torch.logsumexp(t, dims) if semiring == "logaddexp" else t.amax(dims)
            kept = [n for i, n in enumerate(own) if i not in dims]
            if own is x_names:
                x, x_names = t, kept
            else:
                y, y_names = t, kept
    reduced = {r for r in reduced if r in x_names and r in y_names}
    if not reduced:
        return None
    shared = [n for n in x_names if n in y_names and n not in reduced]  # batch dims
    ks = [n for n in x_names if n in reduced]
    x_only = [n for n in x_names if n not in y_names]
    y_only = [n for n in y_names if n not in x_names]
    size = {n: x.shape[i] for i, n in enumerate(x_names)}
    size.update({n: y.shape[i] for i, n in enumerate(y_names)})

    def prod(names):
        p = 1
        for n in names:
            p *= size[n]
        return p

    xb = x.permute([x_names.index(n) for n in shared + x_only + ks])
    yb = y.permute([y_names.index(n) for n in shared + ks + y_only])
    bshape = tuple(size[n] for n in shared)
    I, K, J = prod(x_only), prod(ks), prod(y_only)
    xb = xb.reshape(bshape + (I, K))  # a view whenever the grouped dims are mergeable, else one copy
    yb = yb.reshape(bshape + (K, J))
    out = matmul(semiring, xb, yb)  # [batch..., I, J]
    out = out.reshape(bshape + tuple(size[n] for n in x_only) + tuple(size[n] for n in y_only))
    cur = shared + x_only + y_only
    want = [n for n in OrderedDict.fromkeys(x_names + y_names) if n not in reduced]
    out = out.permute([cur.index(n) for n in want])
    return out, want


def markov_named(trans, names, time, prev, curr, semiring, chain):
    """Chain product over `time` of trans[names] with (prev -> curr) steps.

    Returns (data, out_names): out_names = names without `time`, in the original order -- what
    sequential_sum_product returns (funsor/sum_product.py:652-701) before step_names are substituted."""
    names = list(names)
    batch = [n for n in names if n not in (time, prev, curr)]
    t = trans.permute([names.index(n) for n in batch + [time, prev, curr]])
    out = chain(semiring, t)  # [batch..., S, S]
    cur = batch + [prev, curr]
    want = [n for n in names if n != time]
    return out.permute([cur.index(n) for n in want]), want


def gaussian_markov_named(white_vec, prec_sqrt, scalar, bint_names, real_blocks, time, prev, curr, chain):
    """Scan of Gaussian chain elements in the reference's sqrt form (funsor/gaussian.py:419-498).

    white_vec[bints..., rank], prec_sqrt[bints..., dim, rank]; the leading dims are the bint inputs `bint_names`
    (one of them `time`), the rows of prec_sqrt are the real inputs concatenated in `real_blocks` order
    [(name, size), ...] (gaussian.py:133-151), which must be exactly {prev, curr} with equal sizes D.
    scalar: None or a tensor over the same bint dims (the Tensor half of a `Tensor + Gaussian` element).
    chain(w[B,T,r], P[B,T,2D,r], s[B,T]) -> (Lam[B,2D,2D], eta[B,2D], c[B]) over rows (prev; curr).
    Returns (Lam, eta, c, batch_names) with the batch dims in their original order."""
    bint_names = list(bint_names)
    names = [n for n, _ in real_blocks]
    sizes = dict(real_blocks)
    if sorted(names) != sorted([prev, curr]) or sizes[prev] != sizes[curr]:
        return None
    D = sizes[prev]
    batch = [n for n in bint_names if n != time]
    perm = [bint_names.index(n) for n in batch + [time]]
    nb = len(bint_names)
    w = white_vec.permute(perm + [nb])
    P = prec_sqrt.permute(perm + [nb, nb + 1])
    if names[0] != prev:  # rows are (curr; prev): swap the two blocks
        P = torch.cat([P[..., D:, :], P[..., :D, :]], dim=-2)
    bshape = tuple(w.shape[:-2])
    T, r = w.shape[-2], w.shape[-1]
    if scalar is None:
        s = w.new_zeros(bshape + (T,))
    else:
        s = scalar.permute(perm).expand(bshape + (T,))
    Bflat = 1
    for d in bshape:
        Bflat *= d
    Lam, eta, c = chain(w.reshape(Bflat, T, r), P.reshape(Bflat, T, 2 * D, r), s.reshape(Bflat, T))
    return Lam.reshape(bshape + (2 * D, 2 * D)), eta.reshape(bshape + (2 * D,)), c.reshape(bshape), batch


def gaussian_contract_named(x, y, reduced, bint_sizes, out_bints, out_reals, to_info, combine, to_sqrt):
    """One Gaussian pair contraction of the scan (cnf.py:385-389: (x + y).reduce(logaddexp, reduced)) on raw tensors.

    x, y: dicts with `w` [bints..., r], `P` [bints..., dim, r], `s` (None or a tensor over `s_bints`), `bints` (names of the
    leading dims of w / P), `s_bints`, `reals` [(name, size), ...] in row order of P (gaussian.py:133-151).
    x must be over real inputs {a, b}, y over {b, c} with reduced == {b} and all three of size D; otherwise returns None.
    bint_sizes: {name: size}; out_bints: batch names of the result in order; out_reals: [a, c] in the order the result lists them.
    to_info(w[N,r], P[N,2D,r], s[N]) -> (Lam, eta, c); combine(X, Y) -> (Lam, eta, c, flag); to_sqrt(Lam, eta, c) -> (w, P, s).
    Returns (w[bints..., 2D], P[bints..., 2D, 2D], s[bints...])."""
    if len(reduced) != 1:
        return None
    (b,) = tuple(reduced)
    xr, yr = dict(x["reals"]), dict(y["reals"])
    if b not in xr or b not in yr or len(xr) != 2 or len(yr) != 2:
        return None
    (a,) = [k for k in xr if k != b]
    (c,) = [k for k in yr if k != b]
    D = xr[b]
    if a == c or not (xr[a] == yr[c] == yr[b] == D) or sorted(out_reals) != sorted([a, c]):
        return None
    bshape = tuple(bint_sizes[n] for n in out_bints)
    N = 1
    for d in bshape:
        N *= d

    def aligned(t, names, tail):
        """t[names..., *tail dims] -> broadcast over out_bints, flattened to [N, *tail]"""
        names = list(names)
        perm = [names.index(n) for n in out_bints if n in names]
        t = t.permute(perm + list(range(len(names), t.dim())))
        shape = [bint_sizes[n] if n in names else 1 for n in out_bints]
        t = t.reshape(shape + list(t.shape[len(perm):]))
        return t.expand(list(bshape) + list(t.shape[len(out_bints):])).reshape((N,) + tuple(tail))

    def info(g, first, second):
        r = g["w"].shape[-1]
        P = g["P"]
        if [k for k, _ in g["reals"]] != [first, second]:  # rows are (second; first): swap the two blocks
            P = torch.cat([P[..., D:, :], P[..., :D, :]], dim=-2)
        w = aligned(g["w"], g["bints"], (r,))
        P = aligned(P, g["bints"], (2 * D, r))
        sc = w.new_zeros(N) if g["s"] is None else aligned(g["s"], g["s_bints"], ())
        return to_info(w, P, sc)

    Lo, eo, co, flag = combine(info(x, a, b), info(y, b, c))
    if flag is not None and bool(flag.any()):
        raise ValueError("Too little information to marginalize over {}. Consider adding a prior.".format({b}))  # gaussian.py:897-901
    w, P, sc = to_sqrt(Lo, eo, co)
    if list(out_reals) != [a, c]:
        P = torch.cat([P[..., D:, :], P[..., :D, :]], dim=-2)
    return w.reshape(bshape + (2 * D,)), P.reshape(bshape + (2 * D, 2 * D)), sc.reshape(bshape)


def info_to_sqrt(Lam, eta, c):
    """(Lam, eta, c) -> the reference's (white_vec, prec_sqrt, scalar) triple with P P' = Lam, P w = eta, s - |w|^2/2 = c.
    The precision of a chain product is positive SEMI-definite in general (transition factors are rank deficient,
    SURVEY section 7): diagonally pivoted Cholesky on the GPU (fb2_gaussian_info_to_sqrt_f32), one warp per element."""
    return _HOOKS["info_to_sqrt"](Lam, eta, c)


# ------------------------------------------------------------------------------------------------
# kernel entry points used by the rules
def _kernel_matmul(semiring, x, y):
    from . import ops

    return ops.semiring_matmul(semiring, "add", x, y)


def _kernel_chain(semiring, t):
    from . import ops

    return ops.sequential_sum_product(semiring, "add", t)


def _kernel_gaussian_chain(w, P, s):
    from . import gaussian as fg

    T = w.shape[1]
    # GaussianHMM with fixed dynamics reaches this rule with a dense prec_sqrt[B, T, 2D, r] whose time slices are all equal
    # (pyro/hmm.py:258-270: only white_vec depends on t; the Gaussian product materialises the broadcast).  A probe of four
    # time slices rules the heterogeneous case out cheaply; only a chain that passes it pays the full comparison pass.  The scan
    # then runs on per-level operators + eta-only sweeps (gaussian.gaussian_chain_homog).
    if T >= 8:
        probe = P[:, [1, T // 2, T - 2, T - 1]]
        if P.stride(1) == 0 or (bool((probe == P[:, :1]).all()) and bool((P[:, 1:] == P[:, :1]).all())):
            _STATE["calls"]["gaussian_homogeneous"] += 1
            return fg.gaussian_chain_homog(w.contiguous(), P[:, 0].contiguous(), s.contiguous())
    Lam, eta, c = fg.sqrt_to_info(w.contiguous(), P.contiguous(), s.contiguous())
    return fg.gaussian_sequential_sum_product(Lam, eta, c)


def _kernel_sqrt_to_info(w, P, s):
    from . import gaussian as fg

    return fg.sqrt_to_info(w, P, s)


def _kernel_gaussian_pair(X, Y):
    from . import gaussian as fg

    return fg.gaussian_pair_combine(X, Y)


def _kernel_info_to_sqrt(Lam, eta, c):
    from . import gaussian as fg

    return fg.info_to_sqrt(Lam, eta, c)


def _kernel_moment_match(logits, Lam, eta, c):
    from . import gaussian as fg

    return fg.gaussian_moment_match(logits, Lam, eta, c)


def _fast_ok(*datas):
    return all(isinstance(d, torch.Tensor) and d.is_cuda and d.dtype == torch.float32 for d in datas)


def _no_grad_needed(*datas):
    return not (torch.is_grad_enabled() and any(isinstance(d, torch.Tensor) and d.requires_grad for d in datas))


# The kernel entry points and the admission test, looked up at call time: the CPU tests replace them with oracle stand-ins to
# drive the rules' fast branches inside the unmodified reference without a GPU (tests/test_funsor_plugin.py); on the GPU box
# the defaults below are what runs (tests/test_gpu_funsor.py).
_HOOKS = {"fast_ok": _fast_ok, "matmul": _kernel_matmul, "chain": _kernel_chain, "gaussian_chain": _kernel_gaussian_chain,
          "sqrt_to_info": _kernel_sqrt_to_info, "gaussian_pair": _kernel_gaussian_pair, "info_to_sqrt": _kernel_info_to_sqrt,
          "moment_match": _kernel_moment_match}
_DEFAULT_HOOKS = dict(_HOOKS)


def set_hooks(**hooks):
    """Test aid: override kernel entry points / the admission test; set_hooks() with no arguments restores the defaults."""
    if not hooks:
        _HOOKS.update(_DEFAULT_HOOKS)
    for k, v in hooks.items():
        if k not in _HOOKS:
            raise KeyError(k)
        _HOOKS[k] = v


# ------------------------------------------------------------------------------------------------
def register():
    """Register the rules into the imported funsor (torch backend).  Idempotent."""
    if _STATE["registered"]:
        return
    import funsor
    import funsor.ops as fops
    from funsor.cnf import Contraction
    from funsor.domains import Bint
    from funsor.interpretations import eager
    from funsor.sum_product import MarkovProduct
    from funsor.tensor import Tensor
    from funsor.terms import Subs, Variable

    if funsor.get_backend() != "torch":
        raise RuntimeError("funsor_b200 plugs into the torch backend: call funsor.set_backend('torch') first")

    # ---- a12: the torch op registrations that allocate on the DEFAULT device (funsor/torch/ops.py:194, 201, 303-306, 311) break every
    # CUDA expression that touches them: Slice / arange substitution in the scan (sum_product.py:692-698), AdjointTape.adjoint
    # (adjoint.py:250), matrix_and_mvn_to_funsor (pyro/convert.py:284).  Same semantics, allocated on the operand's device.
    @fops.new_arange.register(torch.Tensor)
    def _new_arange(x, start, stop, step):
        if step is not None:
            return torch.arange(start, stop, step, device=x.device)
        if stop is not None:
            return torch.arange(start, stop, device=x.device)
        return torch.arange(start, device=x.device)

    @fops.new_eye.register(torch.Tensor)
    def _new_eye(x, shape):
        return torch.eye(shape[-1], dtype=x.dtype, device=x.device).expand(shape + (-1,))

    @fops.triangular_inv.register(torch.Tensor)
    def _triangular_inv(x, upper=False):
        if x.size(-1) == 1:
            return x.reciprocal()
        return torch.linalg.solve_triangular(x, torch.eye(x.size(-1), dtype=x.dtype, device=x.device), upper=upper)

    @fops.cholesky_inverse.register(torch.Tensor)
    def _cholesky_inverse(x):
        if x.dim() == 2:
            return x.cholesky_inverse()
        return torch.eye(x.size(-1), dtype=x.dtype, device=x.device).cholesky_solve(x)

    # ---- capture the handlers that are in force now, by dispatching on tiny sample terms
    sx = Tensor(torch.zeros(2, 2), OrderedDict(a=Bint[2], b=Bint[2]))
    sy = Tensor(torch.zeros(2, 2), OrderedDict(b=Bint[2], c=Bint[2]))
    rv = frozenset({Variable("b", Bint[2])})
    prev_contract = {}
    for name, op in (("logaddexp", fops.logaddexp), ("max", fops.max)):
        prev_contract[name] = eager.dispatch(Contraction, op, fops.add, rv, sx, sy)
    st = Tensor(torch.zeros(2, 2, 2), OrderedDict(time=Bint[2], prev=Bint[2], curr=Bint[2]))
    tv = Variable("time", Bint[2])
    step = frozenset({("prev", "curr")})
    names = frozenset({("curr", "prev")})
    prev_markov = eager.dispatch(MarkovProduct, fops.logaddexp, fops.add, st, tv, step, names)

    def _semiring(op):
        return "logaddexp" if op is fops.logaddexp else "max"

    def contraction_rule(red_op, bin_op, reduced_vars, x, y):
        sname = _semiring(red_op)
        fast = (
            x.dtype == "real" and y.dtype == "real" and not x.shape and not y.shape
            and _HOOKS["fast_ok"](x.data, y.data) and len(reduced_vars) >= 1
        )
        if fast:
            reduced = {v.name for v in reduced_vars}
            res = contract_named(x.data, x.inputs, y.data, y.inputs, reduced, sname, _HOOKS["matmul"])
            if res is not None:
                data, out_names = res
                inputs = OrderedDict()
                for term in (x, y):
                    inputs.update(term.inputs)
                _STATE["calls"]["contraction_fast"] += 1
                return Tensor(data, OrderedDict((n, inputs[n]) for n in out_names))
        _STATE["calls"]["contraction_delegated"] += 1
        return prev_contract[sname](red_op, bin_op, reduced_vars, x, y)

    eager.register(Contraction, fops.LogaddexpOp, fops.AddOp, frozenset, Tensor, Tensor)(contraction_rule)
    eager.register(Contraction, fops.MaxOp, fops.AddOp, frozenset, Tensor, Tensor)(contraction_rule)

    def markov_rule(sum_op, prod_op, trans, time, step, step_names):
        fast = (
            len(step) == 1 and trans.dtype == "real" and not trans.shape and _HOOKS["fast_ok"](trans.data)
            and time.name in trans.inputs
        )
        if fast:
            ((prev, curr),) = tuple(step)
            if prev in trans.inputs and curr in trans.inputs and trans.inputs[prev] == trans.inputs[curr]:
                data, out_names = markov_named(trans.data, trans.inputs, time.name, prev, curr, _semiring(sum_op), _HOOKS["chain"])
                result = Tensor(data, OrderedDict((n, trans.inputs[n]) for n in out_names))
                _STATE["calls"]["markov_fast"] += 1
                return Subs(result, step_names)  # funsor/sum_product.py:1064
        _STATE["calls"]["markov_delegated"] += 1
        return prev_markov(sum_op, prod_op, trans, time, step, step_names)

    for op_t in (fops.LogaddexpOp, fops.MaxOp):
        eager.register(MarkovProduct, op_t, fops.AddOp, Tensor, Variable, frozenset, frozenset)(markov_rule)

    # ---- Gaussian chain elements: MarkovProduct[Logaddexp, Add, Gaussian | Tensor + Gaussian, ...]
    from funsor.cnf import GaussianMixture
    from funsor.domains import Reals
    from funsor.gaussian import Gaussian

    def _split_gaussian(trans):
        """(scalar Tensor or None, Gaussian) of a chain element, or None if it is neither form."""
        if isinstance(trans, Gaussian):
            return None, trans
        if isinstance(trans, Contraction) and len(trans.terms) == 2 and not trans.reduced_vars:
            t, g = trans.terms
            if isinstance(t, Tensor) and isinstance(g, Gaussian) and t.dtype == "real" and not t.shape:
                return t, g
        return None

    def gaussian_markov_rule(sum_op, prod_op, trans, time, step, step_names):
        parts = _split_gaussian(trans)
        if parts is not None and len(step) == 1:
            t, g = parts
            ((prev, curr),) = tuple(step)
            bints = [(k, d) for k, d in g.inputs.items() if d.dtype != "real"]
            reals = [(k, d.num_elements) for k, d in g.inputs.items() if d.dtype == "real"]
            ok = (_HOOKS["fast_ok"](g.white_vec, g.prec_sqrt) and _no_grad_needed(g.white_vec, g.prec_sqrt, None if t is None else t.data)
                  and time.name in dict(bints)
                  and all(len(g.inputs[k].shape) == 1 for k, _ in reals) and max(n for _, n in reals) <= 32)
            scalar = None
            if ok and t is not None:
                ok = _HOOKS["fast_ok"](t.data) and set(t.inputs) <= set(k for k, _ in bints)
                if ok:  # broadcast the Tensor half over the Gaussian's bint dims
                    tdims = list(t.inputs)
                    data = t.data.permute([tdims.index(k) for k, _ in bints if k in tdims])
                    shape = [d.size if k in tdims else 1 for k, d in bints]
                    scalar = data.reshape(shape).expand([d.size for _, d in bints])
            if ok:
                res = gaussian_markov_named(g.white_vec, g.prec_sqrt, scalar, [k for k, _ in bints], reals, time.name, prev, curr,
                                            _HOOKS["gaussian_chain"])
                if res is not None:
                    Lam, eta, c, batch = res
                    w, P, s = info_to_sqrt(Lam, eta, c)
                    D = dict(reals)[prev]
                    binputs = OrderedDict((k, g.inputs[k]) for k in batch)
                    ginputs = binputs.copy()
                    # order of the real inputs as the reference's scan leaves it: untouched for a single step; (prev, curr) after one
                    # level of bare Gaussians (the Gaussian sum lists its left operand first, gaussian.py:1123-1124); (curr, prev) as
                    # soon as the last level combined `Tensor + Gaussian` elements (see gaussian_contraction_rule)
                    T = g.inputs[time.name].size
                    if T == 1:
                        real_order = [k for k, _ in reals]
                    elif T == 2 and t is None:
                        real_order = [prev, curr]
                    else:
                        real_order = [curr, prev]
                    if real_order[0] == curr:
                        P = torch.cat([P[..., D:, :], P[..., :D, :]], dim=-2)
                    for k in real_order:
                        ginputs[k] = Reals[D]
                    result = Tensor(s, binputs) + Gaussian(white_vec=w, prec_sqrt=P, inputs=ginputs)
                    _STATE["calls"]["markov_fast"] += 1
                    return Subs(result, step_names)
        _STATE["calls"]["markov_delegated"] += 1
        return prev_markov_gauss(sum_op, prod_op, trans, time, step, step_names)

    sg = Gaussian(white_vec=torch.zeros(2, 2), prec_sqrt=torch.eye(2).expand(2, 2, 2).contiguous(),
                  inputs=OrderedDict(time=Bint[2], prev=Reals[1], curr=Reals[1]))
    prev_markov_gauss = eager.dispatch(MarkovProduct, fops.logaddexp, fops.add, sg, tv, step, names)
    eager.register(MarkovProduct, fops.LogaddexpOp, fops.AddOp, Gaussian, Variable, frozenset, frozenset)(gaussian_markov_rule)
    eager.register(MarkovProduct, fops.LogaddexpOp, fops.AddOp, GaussianMixture, Variable, frozenset, frozenset)(gaussian_markov_rule)

    # ---- one level of the Gaussian scan: Contraction[Logaddexp, Add, frozenset, Gaussian | GaussianMixture x 2] (cnf.py:385-389)
    def _parts_dict(t, g):
        bints = [k for k, d in g.inputs.items() if d.dtype != "real"]
        reals = [(k, d.num_elements) for k, d in g.inputs.items() if d.dtype == "real"]
        return dict(w=g.white_vec, P=g.prec_sqrt, s=None if t is None else t.data, bints=bints,
                    s_bints=[] if t is None else list(t.inputs), reals=reals)

    def gaussian_contraction_rule(red_op, bin_op, reduced_vars, x, y):
        px, py = _split_gaussian(x), _split_gaussian(y)
        if px is not None and py is not None and len(reduced_vars) == 1:
            (tx, gx), (ty, gy) = px, py
            datas = [gx.white_vec, gx.prec_sqrt, gy.white_vec, gy.prec_sqrt] + [t.data for t in (tx, ty) if t is not None]
            (rv,) = tuple(reduced_vars)
            ok = (_HOOKS["fast_ok"](*datas) and _no_grad_needed(*datas) and rv.dtype == "real" and len(rv.output.shape) == 1
                  and rv.output.num_elements <= 32
                  and all(len(d.shape) == 1 for g in (gx, gy) for d in g.inputs.values() if d.dtype == "real")
                  and all(t is None or all(d.dtype != "real" for d in t.inputs.values()) for t in (tx, ty)))
            if ok:
                # input order of the reference's result: the Gaussian sum lists the inputs of its left operand first
                # (gaussian.py:1123-1124), and normalisation of (Tx + Gx) + (Ty + Gy) puts Gy on the left whenever a Tensor part is
                # present (verified against the reference in tests/test_funsor_plugin.py); the scalar part follows the same batch
                order = (gx, gy) if (tx is None and ty is None) else (gy, gx)
                ginputs = OrderedDict()
                for g in order:
                    ginputs.update(g.inputs)
                for t in (tx, ty):
                    if t is not None:
                        for k, d in t.inputs.items():
                            ginputs.setdefault(k, d)
                ginputs.pop(rv.name, None)
                binputs = OrderedDict((k, d) for k, d in ginputs.items() if d.dtype != "real")
                out_reals = [k for k, d in ginputs.items() if d.dtype == "real"]
                ginputs = OrderedDict(list(binputs.items()) + [(k, ginputs[k]) for k in out_reals])
                res = gaussian_contract_named(_parts_dict(tx, gx), _parts_dict(ty, gy), {rv.name}, {k: d.size for k, d in binputs.items()},
                                              list(binputs), out_reals, _HOOKS["sqrt_to_info"], _HOOKS["gaussian_pair"], _HOOKS["info_to_sqrt"])
                if res is not None:
                    w, P, sc = res
                    _STATE["calls"]["gaussian_contraction_fast"] += 1
                    return Tensor(sc, binputs) + Gaussian(white_vec=w, prec_sqrt=P, inputs=ginputs)
        _STATE["calls"]["gaussian_contraction_delegated"] += 1
        return prev_gauss_contract[(type(x) is Gaussian, type(y) is Gaussian)](red_op, bin_op, reduced_vars, x, y)

    sgx = Gaussian(white_vec=torch.zeros(2), prec_sqrt=torch.eye(2), inputs=OrderedDict(a=Reals[1], b=Reals[1]))
    sgy = Gaussian(white_vec=torch.zeros(2), prec_sqrt=torch.eye(2), inputs=OrderedDict(b=Reals[1], c=Reals[1]))
    stz = Tensor(torch.zeros(()), OrderedDict())
    rvb = frozenset({Variable("b", Reals[1])})
    with funsor.interpretations.lazy:
        smx, smy = stz + sgx, stz + sgy
    prev_gauss_contract = {}
    for fx, sx_ in ((True, sgx), (False, smx)):
        for fy, sy_ in ((True, sgy), (False, smy)):
            prev_gauss_contract[(fx, fy)] = eager.dispatch(Contraction, fops.logaddexp, fops.add, rvb, sx_, sy_)
    for tx_ in (Gaussian, GaussianMixture):
        for ty_ in (Gaussian, GaussianMixture):
            eager.register(Contraction, fops.LogaddexpOp, fops.AddOp, frozenset, tx_, ty_)(gaussian_contraction_rule)

    # ---- moment matching (joint.py:102-150): collapse the mixture components indexed by the reduced discrete inputs
    from funsor.interpretations import moment_matching

    def moment_matching_rule(red_op, bin_op, reduced_vars, discrete, gaussian):
        approx_vars = frozenset(v for v in reduced_vars & gaussian.input_vars if v.dtype != "real")
        datas = (discrete.data, gaussian.white_vec, gaussian.prec_sqrt)
        reals = [(k, d) for k, d in gaussian.inputs.items() if d.dtype == "real"]
        n = sum(d.num_elements for _, d in reals)
        fast = (approx_vars and approx_vars == reduced_vars and isinstance(discrete, Tensor) and discrete.dtype == "real" and not discrete.shape
                and all(v.name in discrete.inputs for v in approx_vars) and all(d.dtype != "real" for d in discrete.inputs.values())
                and _HOOKS["fast_ok"](*datas) and _no_grad_needed(*datas) and 1 <= n <= 64 and gaussian.rank >= n)
        if fast:
            approx = [k for k in gaussian.inputs if k in {v.name for v in approx_vars}]
            all_ints = OrderedDict(discrete.inputs)
            all_ints.update((k, d) for k, d in gaussian.inputs.items() if d.dtype != "real")
            binputs = OrderedDict((k, d) for k, d in all_ints.items() if k not in approx)  # order of new_loc.inputs (joint.py:127-128)
            order = list(binputs) + approx
            bshape = tuple(d.size for d in binputs.values())
            K = 1
            for k in approx:
                K *= all_ints[k].size
            N = 1
            for d_ in bshape:
                N *= d_

            def aligned(t, names, tail):
                names = list(names)
                perm = [names.index(k) for k in order if k in names]
                t = t.permute(perm + list(range(len(names), t.dim())))
                t = t.reshape([all_ints[k].size if k in names else 1 for k in order] + list(tail))
                return t.expand([all_ints[k].size for k in order] + list(tail)).reshape((N * K,) + tuple(tail))

            gints = [k for k, d in gaussian.inputs.items() if d.dtype != "real"]
            r = gaussian.rank
            w = aligned(gaussian.white_vec, gints, (r,))
            P = aligned(gaussian.prec_sqrt, gints, (n, r))
            logits = aligned(discrete.data, list(discrete.inputs), ())
            Lam, eta, c = _HOOKS["sqrt_to_info"](w, P, w.new_zeros(N * K))
            Lo, eo, co = _HOOKS["moment_match"](logits.reshape(N, K), Lam.reshape(N, K, n, n), eta.reshape(N, K, n), c.reshape(N, K))
            wv, Ps, sc = _HOOKS["info_to_sqrt"](Lo, eo, co)
            ginputs = binputs.copy()
            ginputs.update(reals)
            _STATE["calls"]["moment_matching_fast"] += 1
            return Tensor(sc.reshape(bshape), binputs) + Gaussian(white_vec=wv.reshape(bshape + (n,)), prec_sqrt=Ps.reshape(bshape + (n, n)),
                                                                  inputs=ginputs)
        _STATE["calls"]["moment_matching_delegated"] += 1
        return prev_mm(red_op, bin_op, reduced_vars, discrete, gaussian)

    smd = Tensor(torch.zeros(2), OrderedDict(k=Bint[2]))
    smg = Gaussian(white_vec=torch.zeros(2, 1), prec_sqrt=torch.ones(2, 1, 1), inputs=OrderedDict(k=Bint[2], x=Reals[1]))
    prev_mm = moment_matching.dispatch(Contraction, fops.logaddexp, fops.add, frozenset({Variable("k", Bint[2])}), smd, smg)
    moment_matching.register(Contraction, fops.LogaddexpOp, fops.AddOp, frozenset, Tensor, Gaussian)(moment_matching_rule)

    _STATE["registered"] = True


def sequential_sum_product(sum_op, prod_op, trans, time, step):
    """Same signature as funsor.sum_product.sequential_sum_product (652-701).  funsor/pyro/hmm.py:131, 269, 537 call that
    function directly (not through MarkovProduct), so callers that want the fused scan import this one: it builds the
    MarkovProduct term, which the rules registered above evaluate on the GPU (Tensor and Gaussian chain elements alike) or
    hand to the stock rule; the step renaming of MarkovProduct is undone so that the result has the scan's own inputs."""
    from funsor.sum_product import MarkovProduct
    from funsor.sum_product import sequential_sum_product as ref_ssp

    if not _STATE["registered"] or len(step) != 1:
        return ref_ssp(sum_op, prod_op, trans, time, step)
    # MarkovProduct with identity step_names == the raw scan (sum_product.py:1052-1064 with no renaming)
    return MarkovProduct(sum_op, prod_op, trans, time, dict(step))
