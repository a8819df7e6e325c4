"""Forward-backward timing at small batch: sequential passes vs the two passes side by side (FB2_FLOW_DUAL)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import funsor_b200 as fb

dev = torch.device("cuda")
for (B, T, S) in [(1, 131072, 512), (1, 131072, 256), (16, 32768, 512), (1, 65536, 128), (1, 32768, 1024)]:
    g = torch.Generator(device=dev); g.manual_seed(1)
    A = torch.log_softmax(torch.randn(S, S, device=dev, generator=g), -1)
    init = torch.log_softmax(torch.randn(B, S, device=dev, generator=g), -1)
    obs = torch.randn(B, T, S, device=dev, generator=g)
    res = {}
    for dual in ("0", "1"):
        os.environ["FB2_FLOW_DUAL"] = dual
        for _ in range(2):
            z, gam, _ = fb.hmm_marginals(init, A, obs, want_xi=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            z, gam, _ = fb.hmm_marginals(init, A, obs, want_xi=False)
        e1.record(); torch.cuda.synchronize()
        res[dual] = (e0.elapsed_time(e1) / 3, z.clone(), gam[:, ::997].clone())
    d = (res["0"][2] - res["1"][2]).abs().max().item()
    print("B=%d T=%d S=%d: sequential %.1f ms (%.2f us/step)  side-by-side %.1f ms (%.2f us/step)  logZ %s/%s  max|dgamma| %.2e"
          % (B, T, S, res["0"][0], res["0"][0] * 1e3 / T, res["1"][0], res["1"][0] * 1e3 / T, res["0"][1][0].item(), res["1"][1][0].item(), d))
