"""C4: FFMA tile kernel vs tensor-core tile kernel."""
import os, sys
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
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
res = {}
for tc in ("0", "1"):
    os.environ["FB2_KALMAN_TC"] = tc
    for K in (3, 4, 5):
        os.environ["FB2_KALMAN_K"] = str(K)
        out = fg.kalman_chain(*args)
        res[(tc, K)] = out
        err = max((a - b).abs().max().item() / max(b.abs().max().item(), 1.0) for a, b in zip(out, res[("0", K)]))
        print("tensor-core tiles=%s K=%d: %.3f ms  max rel diff vs FFMA tiles %.2e" % (tc, K, timeit(lambda: fg.kalman_chain(*args, validate=False)), err), flush=True)
