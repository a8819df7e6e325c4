import sys, os
sys.path.insert(0, "/root/repo")
import torch
import funsor_b200 as fb
def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for (B, T, S) in [(1, 100, 16), (16, 1000, 16), (1024, 1000, 16), (4096, 1000, 32), (4096, 1000, 64)]:
    init = torch.randn(B, S, device="cuda").log_softmax(-1); A = torch.randn(S, S, device="cuda").log_softmax(-1); obs = torch.randn(B, T, S, device="cuda")
    for name, env in (("small", None), ("generic", "1")):
        if env: os.environ["FB2_DISABLE_FLOW"] = env
        else: os.environ.pop("FB2_DISABLE_FLOW", None)
        if env and B > 1024: continue
        ms = timeit(lambda: fb.hmm_forward(init, A, obs))
        print(name, (B, T, S), "%.3f ms  %.3g units/s" % (ms, B * T * S * S / ms * 1e3), flush=True)
os.environ.pop("FB2_DISABLE_FLOW", None)
