"""Two-hop kernel: clusters of 8 (default) vs clusters of 4 with half-K warps (FB2_FLOW_CL4=1) at S in (512, 1024]."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import funsor_b200 as fb
dev = torch.device("cuda")
os.environ["FB2_FLOW_DIRECT"] = "0"
for (B, T, S) in [(16, 8192, 1024), (3, 4096, 700), (40, 2048, 1000)]:
    g = torch.Generator(device=dev); g.manual_seed(1)
    A = torch.log_softmax(torch.randn(S, S, device=dev, generator=g), -1)
    init = torch.log_softmax(torch.randn(B, S, device=dev, generator=g), -1)
    obs = torch.randn(B, T, S, device=dev, generator=g)
    res = {}
    for mode in ("0", "1"):
        os.environ["FB2_FLOW_CL4"] = mode
        for _ in range(2):
            z = fb.hmm_forward(init, A, obs)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            z = fb.hmm_forward(init, A, obs)
        e1.record(); torch.cuda.synchronize()
        res[mode] = (e0.elapsed_time(e1) / 3, z.clone())
    groups = (B + 15) // 16
    d = ((res["0"][1] - res["1"][1]).abs() / res["0"][1].abs()).max().item()
    print("B=%d T=%d S=%d: clusters of 8 %.2f ms (%.2f us/step/group)  clusters of 4 %.2f ms (%.2f us/step/group)  max rel diff %.2e  logZ0 %.4f / %.4f"
          % (B, T, S, res["0"][0], res["0"][0] * 1e3 / T / groups, res["1"][0], res["1"][0] * 1e3 / T / groups, d, res["0"][1][0].item(), res["1"][1][0].item()), flush=True)
