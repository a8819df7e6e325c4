"""gaussian_chain_homog(white_vec, prec_sqrt, scalar) at C4 size: the funsor rule's path for a Gaussian MarkovProduct with time-invariant prec_sqrt."""
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
w, P, s = fg.kalman_elements_homog(*args)
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ref = None
for eta_tc, tiles in (("0", "0"), ("0", "1"), ("1", "1")):
    os.environ.update({"FB2_GAUSS_ETA_TC": eta_tc, "FB2_GAUSS_HOMOG_TILES": tiles})
    out = fg.gaussian_chain_homog(w, P, s)
    if ref is None: ref = out
    err = max((a - b).abs().max().item() / max(b.abs().max().item(), 1.0) for a, b in zip(out, ref))
    print("eta on tensor cores=%s tile walk=%s: %.3f ms   rel diff vs first %.2e" % (eta_tc, tiles, timeit(lambda: fg.gaussian_chain_homog(w, P, s, validate=False)), err), flush=True)
