import torch, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import funsor_b200 as fb
def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
out = {}
for S in (256, 512):
    g = torch.Generator(device="cuda").manual_seed(0)
    T = 16384
    A = torch.randn(S, S, device="cuda", generator=g).log_softmax(-1)
    obs = torch.randn(1, T, S, device="cuda", generator=g)
    init = torch.randn(1, S, device="cuda", generator=g).log_softmax(-1)
    z = fb.hmm_forward(init, A, obs)
    out["fwd_S%d_us" % S] = timeit(lambda: fb.hmm_forward(init, A, obs)) * 1e3 / T
    out["fb_S%d_us" % S] = timeit(lambda: fb.hmm_marginals(init, A, obs, want_xi=False)) * 1e3 / T
    out["logZ_S%d" % S] = float(z)
print(json.dumps(out))
