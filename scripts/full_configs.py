"""One pass of every BASELINE.json configuration at (or as near as memory allows to) its full size, with the
size-independent checks that stand in for an oracle run.  Output: one JSON line per configuration."""
import json, math, sys, time
sys.path.insert(0, "/root/repo")
import torch
import funsor_b200 as fb
import funsor_b200.gaussian as fg
dev = "cuda"
def timed(fn):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
    return out, e0.elapsed_time(e1)
g = torch.Generator(device=dev).manual_seed(0)
# C1: T=100, S=16, B=1 (dense trans, the reference's own CPU config)
trans = torch.randn(1, 100, 16, 16, device=dev, generator=g); init = torch.randn(1, 16, device=dev, generator=g)
fb.sequential_sum_product("logaddexp", "add", trans)
chain, ms = timed(lambda: fb.sequential_sum_product("logaddexp", "add", trans))
z = torch.logsumexp((init[..., None] + chain).reshape(1, -1), -1)
print(json.dumps({"config": "C1", "T": 100, "S": 16, "B": 1, "ms": ms, "units_per_s": 100 * 256 / ms * 1e3, "logZ": float(z)}), flush=True)
# C2: T=65536, S=1024, B=16
S, T, B = 1024, 65536, 16
A = torch.randn(S, S, device=dev, generator=g).log_softmax(-1); obs = torch.randn(B, T, S, device=dev, generator=g); init = torch.randn(B, S, device=dev, generator=g).log_softmax(-1)
fb.hmm_forward(init, A, obs[:, :64].contiguous())
z, ms = timed(lambda: fb.hmm_forward(init, A, obs))
print(json.dumps({"config": "C2", "T": T, "S": S, "B": B, "ms": ms, "units_per_s": T * S * S * B / ms * 1e3, "us_per_step": ms * 1e3 / T, "logZ0": float(z[0])}), flush=True)
del obs
# C3: Viterbi T=16384, S=4096, planted path
S, T = 4096, 16384
A = torch.randn(S, S, device=dev, generator=g); obs = torch.randn(1, T, S, device=dev, generator=g); init = torch.randn(1, S, device=dev, generator=g)
zpath = torch.randint(0, S, (T + 1,), device=dev, generator=g)
init[0, zpath[0]] += 30.0; obs[0, torch.arange(T, device=dev), zpath[1:]] += 30.0
fb.hmm_viterbi(init, A, obs[:, :32].contiguous())
(best, path), ms = timed(lambda: fb.hmm_viterbi(init, A, obs))
ok = bool((path[0] == zpath).all())
print(json.dumps({"config": "C3", "T": T, "S": S, "B": 1, "ms": ms, "units_per_s": T * S * S / ms * 1e3, "us_per_step": ms * 1e3 / T, "planted_path_recovered": ok}), flush=True)
del A, obs
# C4: Kalman D=32, O=16, B=64, T=100000, homogeneous dynamics (examples/kalman_filter.py shape): full size
Bk, Tk, D, O = 64, 100000, 32, 16
F = 0.9 * torch.linalg.qr(torch.randn(Bk, D, D, device=dev, generator=g))[0]
H = torch.randn(Bk, D, O, device=dev, generator=g) / D ** 0.5
eye = lambda n: torch.eye(n, device=dev).expand(Bk, n, n).contiguous()
y = torch.randn(Bk, Tk, O, device=dev, generator=g)
kargs = (F, torch.zeros(Bk, D, device=dev), eye(D), H, torch.zeros(Bk, O, device=dev), eye(O), y)
fg.kalman_chain(*[a[:, :64].contiguous() if a.dim() == 3 and a.shape[1] == Tk else a for a in kargs])
fg.kalman_chain(*kargs, validate=False)  # first full-size call pays the workspace allocation
out, ms = timed(lambda: fg.kalman_chain(*kargs, validate=False))
lp = fg.gaussian_hmm_log_prob(torch.zeros(Bk, D, device=dev), eye(D), out)
# split-and-recombine property: the chain over [0,T) equals the pair contraction of the chains over [0,T/2) and [T/2,T)
h1 = fg.kalman_chain(*[a[:, :Tk // 2].contiguous() if a.dim() == 3 and a.shape[1] == Tk else a for a in kargs])
h2 = fg.kalman_chain(*[a[:, Tk // 2:].contiguous() if a.dim() == 3 and a.shape[1] == Tk else a for a in kargs])
comb = fg.gaussian_pair_combine(h1, h2)
lp2 = fg.gaussian_hmm_log_prob(torch.zeros(Bk, D, device=dev), eye(D), comb[:3])
print(json.dumps({"config": "C4", "T": Tk, "B": Bk, "D": D, "O": O, "ms": ms, "kalman_steps_per_s": Bk * Tk / ms * 1e3,
                  "log_prob0": float(lp[0]), "finite": bool(torch.isfinite(lp).all()),
                  "split_recombine_rel_err": float(((lp - lp2).abs() / lp.abs()).max())}), flush=True)
del out, y
# C5: forward-backward T=1,048,576, S=512, B=1 (single GPU)
S, T = 512, 1048576
A = torch.randn(S, S, device=dev, generator=g).log_softmax(-1); obs = torch.randn(1, T, S, device=dev, generator=g); init = torch.randn(1, S, device=dev, generator=g).log_softmax(-1)
fb.hmm_marginals(init, A, obs[:, :64].contiguous(), want_xi=True)
(logZ, gamma, xi), ms = timed(lambda: fb.hmm_marginals(init, A, obs, want_xi=True))
norm_err = float((torch.logsumexp(gamma[0, ::4096], -1)).abs().max())
zf = fb.hmm_forward(init, A, obs)
print(json.dumps({"config": "C5", "T": T, "S": S, "B": 1, "ms": ms, "units_per_s": 2 * T * S * S / ms * 1e3, "us_per_step": ms * 1e3 / T,
                  "logZ": float(logZ[0]), "logZ_forward_only": float(zf[0]), "max_abs_log_sum_gamma": norm_err,
                  "xi_sum_total_over_T": float(xi.double().sum() / T)}), flush=True)
