"""C4 (Kalman, D=32, O=16, B=64, T=100000): fused tile path for K / tiles-per-thread / side stream against the level-by-level path."""
import os, sys, time
sys.path.insert(0, "/root/repo")
import torch
import funsor_b200.gaussian as fg

B, T, D, O = 64, 100000, 32, 16
g = torch.Generator().manual_seed(0)
F = 0.9 * torch.linalg.qr(torch.randn(B, D, D, generator=g))[0]
H = torch.randn(B, D, O, generator=g) / D ** 0.5
def tril(n):
    a = torch.randn(B, n, n, generator=g) * 0.2
    return torch.tril(a, -1) + torch.diag_embed(0.5 + torch.rand(B, n, generator=g))
args = [a.cuda() for a in [F, torch.randn(B, D, generator=g) * 0.1, tril(D), H, torch.randn(B, O, generator=g) * 0.1, tril(O), torch.randn(B, T, O, generator=g)]]

def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

os.environ["FB2_KALMAN_FUSED"] = "0"
ref = fg.kalman_chain(*args)
print("level-by-level: %.3f ms" % timeit(lambda: fg.kalman_chain(*args, validate=False)))
os.environ["FB2_KALMAN_FUSED"] = "1"
consts = fg.kalman_constants(*args[:6])
print("constants kernel: %.3f ms" % timeit(lambda: fg.kalman_constants(*args[:6])))
for ntl in ("1", "2"):
    for side in ("0", "1"):
        for K in (3, 4, 5):
            os.environ["FB2_KALMAN_K"] = str(K)
            os.environ["FB2_KALMAN_NTL"] = ntl
            os.environ["FB2_KALMAN_SIDE"] = side
            got = fg.kalman_chain(*args)
            exact = all(torch.equal(a, b) for a, b in zip(got, fg.gaussian_chain_info_homog(*fg.kalman_eta_from_constants(consts, args[6]))))
            err = max((a - b).abs().max().item() / max(b.abs().max().item(), 1.0) for a, b in zip(got, ref))
            print("tiles/thread=%s side=%s K=%d total: %.3f ms   chain only: %.3f ms   bit-identical to the level path: %s   rel err vs torch-constants path %.2e" % (
                ntl, side, K, timeit(lambda: fg.kalman_chain(*args, validate=False)), timeit(lambda: fg.kalman_chain_fused(consts, args[6], validate=False)), exact, err))
