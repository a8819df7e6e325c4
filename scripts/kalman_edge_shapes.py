"""Small Kalman / homogeneous-chain calls for compute-sanitizer (memcheck): every new Gaussian kernel of round 2 at edge shapes."""
import os, sys
sys.path.insert(0, "/root/repo")
import torch
import funsor_b200.gaussian as fg
def kargs(B, T, D, O, seed):
    g = torch.Generator().manual_seed(seed)
    F = 0.9 * torch.linalg.qr(torch.randn(B, D, D, generator=g))[0]
    H = torch.randn(B, D, O, generator=g) / D ** 0.5
    def tril(n):
        a = torch.randn(B, n, n, generator=g) * 0.2
        return torch.tril(a, -1) + torch.diag_embed(0.5 + torch.rand(B, n, generator=g))
    return [a.cuda() for a in [F, torch.randn(B, D, generator=g) * 0.1, tril(D), H, torch.randn(B, O, generator=g) * 0.1, tril(O), torch.randn(B, T, O, generator=g)]]
for (B, T, D, O) in [(2, 2, 32, 16), (1, 33, 32, 16), (3, 97, 20, 7), (2, 1025, 32, 16), (2, 70, 3, 2), (1, 300, 8, 5), (2, 4097, 17, 3)]:
    for tc in ("1", "0"):
        os.environ["FB2_KALMAN_TC"] = tc
        a = kargs(B, T, D, O, T)
        out = fg.kalman_chain(*a)
        w, P, s = fg.kalman_elements_homog(*a)
        out2 = fg.gaussian_chain_homog(w, P, s)
        torch.cuda.synchronize()
        d = max((x - y).abs().max().item() / max(y.abs().max().item(), 1.0) for x, y in zip(out, out2))
        print("B=%d T=%d D=%d O=%d tc=%s ok, fused vs sqrt-form path rel diff %.1e" % (B, T, D, O, tc, d), flush=True)
