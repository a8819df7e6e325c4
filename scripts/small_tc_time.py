"""Batched small-state HMM forward (the reference's examples: hidden_dim 16..64, many chains): warp-per-chain kernel vs 16 chains per warp on the tensor pipe."""
import sys, os
sys.path.insert(0, "/root/repo")
import torch
import funsor_b200 as fb
def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
g = torch.Generator(device="cuda").manual_seed(0)
for (B, T, S) in [(16, 1000, 16), (1024, 1000, 16), (16384, 1000, 16), (4096, 1000, 32), (4096, 1000, 64), (65536, 200, 8), (2048, 1000, 50)]:
    init = torch.randn(B, S, device="cuda", generator=g).log_softmax(-1); A = torch.randn(S, S, device="cuda", generator=g).log_softmax(-1)
    obs = torch.randn(B, T, S, device="cuda", generator=g)
    res = {}
    for tc in ("0", "1"):
        os.environ["FB2_SMALL_TC"] = tc
        z = fb.hmm_forward(init, A, obs)
        res[tc] = (timeit(lambda: fb.hmm_forward(init, A, obs)), z.double())
    ref = torch.zeros(B, dtype=torch.float64, device="cuda")
    a = init.double()
    for t in range(T):
        a = torch.logsumexp(a[:, :, None] + A.double()[None], 1) + obs[:, t].double()
    ref = torch.logsumexp(a, -1)
    e0 = ((res["0"][1] - ref).abs() / ref.abs()).max().item(); e1 = ((res["1"][1] - ref).abs() / ref.abs()).max().item()
    gbs = B * T * S * 4 / (res["1"][0] * 1e-3) / 1e9
    print("B=%d T=%d S=%d: warp per chain %.3f ms (%.3g units/s, rel err %.1e)   16 chains per warp, tensor pipe %.3f ms (%.3g units/s, %.0f GB/s of obs, rel err %.1e)" % (
        B, T, S, res["0"][0], B * T * S * S / res["0"][0] * 1e3, e0, res["1"][0], B * T * S * S / res["1"][0] * 1e3, gbs, e1), flush=True)
