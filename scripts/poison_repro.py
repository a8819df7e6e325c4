"""Run entry points on workspaces that contain garbage (the caching allocator hands back poisoned blocks)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import funsor_b200 as fb
dev = torch.device("cuda")

def poison(nbytes=6 << 30):
    t = torch.empty(nbytes // 4, dtype=torch.int32, device=dev)
    t.fill_(0x7f7f7f7f)
    del t  # back to the caching allocator, contents stay

for (S, B, T) in [(200, 256, 512), (200, 64, 512), (128, 256, 256), (65, 17, 300), (300, 70, 200), (1100, 3, 100)]:
    g = torch.Generator(device=dev); g.manual_seed(1)
    A = torch.log_softmax(torch.randn(S, S, device=dev, generator=g), -1)
    init = torch.log_softmax(torch.randn(B, S, device=dev, generator=g), -1)
    obs = torch.randn(B, T, S, device=dev, generator=g)
    for name, fn in (("fwd", lambda: fb.hmm_forward(init, A, obs)), ("vit", lambda: fb.hmm_viterbi(init, A, obs)[0]), ("marg", lambda: fb.hmm_marginals(init, A, obs, want_xi=False)[0])):
        torch.cuda.empty_cache()
        ref = fn().clone(); torch.cuda.synchronize()
        poison()
        got = fn(); torch.cuda.synchronize()
        print(S, B, T, name, "ok" if torch.equal(ref, got) else "MISMATCH %g" % (ref - got).abs().max().item(), flush=True)
