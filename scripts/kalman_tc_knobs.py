"""C4 with the tensor-core tile kernel: warps per CTA, tile depth K, upper levels on a side stream."""
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
os.environ["FB2_KALMAN_TC"] = "1"
consts = fg.kalman_constants(*args[:6])
for side in ("0", "1"):
    for nw in ("4", "6", "8", "12"):
        for K in (4, 5):
            os.environ.update({"FB2_KALMAN_K": str(K), "FB2_KALMAN_TC_WARPS": nw, "FB2_KALMAN_SIDE": side})
            print("side=%s warps=%s K=%d: total %.3f ms  chain only %.3f ms" % (side, nw, K, timeit(lambda: fg.kalman_chain(*args, validate=False)),
                  timeit(lambda: fg.kalman_chain_fused(consts, args[6], validate=False))), flush=True)
