"""S in (256, 512]: single-hop kernel vs two-hop kernel with 8-byte packets vs two-hop with 4-byte words (forward and forward-backward)."""
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
variants = {"single-hop": {"FB2_FLOW_DIRECT": "1"}, "two-hop 8-byte": {"FB2_FLOW_DIRECT": "0", "FB2_FLOW_PK4": "0"}, "two-hop 4-byte": {"FB2_FLOW_DIRECT": "0", "FB2_FLOW_PK4": "1"}}
for (B, S) in ((1, 512), (16, 512), (1, 384)):
    g = torch.Generator(device="cuda").manual_seed(0)
    T = 16384
    A = torch.randn(S, S, device="cuda", generator=g).log_softmax(-1)
    obs = torch.randn(B, T, S, device="cuda", generator=g)
    init = torch.randn(B, S, device="cuda", generator=g).log_softmax(-1)
    for name, env in variants.items():
        for k in ("FB2_FLOW_DIRECT", "FB2_FLOW_PK4"):
            os.environ.pop(k, None)
        os.environ.update(env)
        z = fb.hmm_forward(init, A, obs)
        fwd = timeit(lambda: fb.hmm_forward(init, A, obs)) * 1e3 / T
        zz, gam, _ = fb.hmm_marginals(init, A, obs, want_xi=False)
        fbk = timeit(lambda: fb.hmm_marginals(init, A, obs, want_xi=False)) * 1e3 / T
        print("B=%2d S=%d %-16s forward %.3f us/step   forward-backward %.3f us/step   logZ %.4f / %.4f  gamma row sum %.6f" % (
            B, S, name, fwd, fbk, float(z[0]), float(zz[0]), float(gam[0, T // 2].exp().sum())), flush=True)
