"""Per-step time across the dispatch boundaries (looking for cliffs)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import funsor_b200 as fb
dev = torch.device("cuda")
T = 2048
for S in (32, 64, 65, 128, 200, 255, 256, 512, 1024, 1025):
    for B in (1, 16, 64, 256):
        g = torch.Generator(device=dev); g.manual_seed(1)
        A = torch.log_softmax(torch.randn(S, S, device=dev, generator=g), -1)
        init = torch.log_softmax(torch.randn(B, S, device=dev, generator=g), -1)
        Tt = T if S * B <= 65536 else 256
        obs = torch.randn(B, Tt, S, device=dev, generator=g)
        row = []
        for name, fn in (("fwd", lambda: fb.hmm_forward(init, A, obs)), ("vit", lambda: fb.hmm_viterbi(init, A, obs)), ("marg", lambda: fb.hmm_marginals(init, A, obs, want_xi=False))):
            fn(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            row.append("%s %7.2f" % (name, e0.elapsed_time(e1) * 1e3 / Tt))
        print("S=%4d B=%3d  us/step: %s" % (S, B, "  ".join(row)), flush=True)
