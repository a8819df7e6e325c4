import os, sys
sys.path.insert(0, "/root/repo")
sys.path.insert(0, "/root/repo/tests")
import torch
import funsor_b200.gaussian as fg
os.environ["FB2_KALMAN_TC"] = "0"
def kargs(B, T, D, O, seed):
    g = torch.Generator().manual_seed(seed)
    F = 0.9 * torch.linalg.qr(torch.randn(B, D, D, generator=g))[0]
    H = torch.randn(B, D, O, generator=g) / D ** 0.5
    def tril(n):
        a = torch.randn(B, n, n, generator=g) * 0.2
        return torch.tril(a, -1) + torch.diag_embed(0.5 + torch.rand(B, n, generator=g))
    return [a.cuda() for a in [F, torch.randn(B, D, generator=g) * 0.1, tril(D), H, torch.randn(B, O, generator=g) * 0.1, tril(O), torch.randn(B, T, O, generator=g)]]
for D, O in ((3, 2), (4, 2), (8, 5), (16, 16), (21, 9), (32, 16)):
    bad = []
    for T in list(range(4, 40)) + [64, 65, 1000]:
        args = kargs(2, T, D, O, 300 * T + D)
        consts = fg.kalman_constants(*args[:6])
        elems = fg.kalman_eta_from_constants(consts, args[6])
        os.environ["FB2_GAUSS_LEVELS_WARP"] = "1"
        want = fg.gaussian_chain_info_homog(*elems)
        os.environ["FB2_GAUSS_LEVELS_WARP"] = "0"
        got = fg.gaussian_chain_info_homog(*elems)
        d = [(a - b).abs().max().item() for a, b in zip(got, want)]
        if max(d) > 0: bad.append((T, ["%.1e" % x for x in d]))
    print("D=%d O=%d: %d mismatching lengths; first: %s" % (D, O, len(bad), bad[:6]), flush=True)
