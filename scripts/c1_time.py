import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import funsor_b200 as fb
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
for (B, T, S) in [(1, 100, 16), (64, 100, 16), (1024, 128, 16), (1, 1000, 8), (8, 4096, 4), (4, 33, 32)]:
    trans = torch.randn(B, T, S, S, device=dev, generator=g)
    for op in ("logaddexp", "max"):
        res = {}
        for tail in ("1", "0"):
            os.environ["FB2_DISABLE_CHAIN_TAIL"] = tail
            for _ in range(3):
                out = fb.sequential_sum_product(op, "add", trans)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                out = fb.sequential_sum_product(op, "add", trans)
            e1.record(); torch.cuda.synchronize()
            res[tail] = (e0.elapsed_time(e1) / 20, out.clone())
        d = (res["0"][1] - res["1"][1]).abs().max().item()
        print("B=%d T=%d S=%d %s: level-by-level %.3f ms, fused tail %.3f ms, max|diff| %.2e" % (B, T, S, op, res["1"][0], res["0"][0], d))
