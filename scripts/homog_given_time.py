"""Homogeneous Gaussian chain from given elements (the funsor rule's path), C4 size: level by level vs tile walk."""
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
elems = fg.kalman_eta_from_constants(fg.kalman_constants(*args[:6]), args[6])
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ref = None
for tiles, tc in (("0", "1"), ("1", "0"), ("1", "1")):
    os.environ.update({"FB2_GAUSS_HOMOG_TILES": tiles, "FB2_KALMAN_TC": tc})
    out = fg.gaussian_chain_info_homog(*elems)
    if ref is None: ref = out
    err = max((a - b).abs().max().item() / max(b.abs().max().item(), 1.0) for a, b in zip(out, ref))
    print("tiles=%s tensor-core=%s: %.3f ms   rel diff vs level-by-level %.2e" % (tiles, tc, timeit(lambda: fg.gaussian_chain_info_homog(*elems, validate=False)), err), flush=True)
