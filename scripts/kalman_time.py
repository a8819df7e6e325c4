"""Kalman / Gaussian chain timing: general (time-varying precision) scan at T=2048 and the homogeneous C4 shape."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import funsor_b200.gaussian as fg

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)


def timed(fn, reps=3):
    for _ in range(2):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def params(Bk, D, O):
    F = 0.9 * torch.linalg.qr(torch.randn(Bk, D, D, device=dev, generator=g))[0]
    H = torch.randn(Bk, D, O, device=dev, generator=g) / D ** 0.5
    eye = lambda n: torch.eye(n, device=dev).expand(Bk, n, n).contiguous()  # noqa: E731
    return F, torch.zeros(Bk, D, device=dev), eye(D), H, torch.zeros(Bk, O, device=dev), eye(O)


Bk, Tk, D, O = 64, 2048, 32, 16
P6 = params(Bk, D, O)
y = torch.randn(Bk, Tk, O, device=dev, generator=g)
w, P, s = fg.kalman_elements_sqrt(*P6, y)
Lam, eta, c = fg.sqrt_to_info(w, P, s)
ms, out = timed(lambda: fg.gaussian_sequential_sum_product(Lam, eta, c, validate=False))
print("general scan B=%d T=%d D=%d: %.3f ms, %.3e pair-steps/s, c[0]=%.4f" % (Bk, Tk, D, ms, Bk * Tk / ms * 1e3, float(out[2][0])))
if len(sys.argv) > 1 and sys.argv[1] == "short":
    sys.exit(0)
del Lam, eta, c, w, P, s
for (Bk, Tk, D, O) in [(64, 100000, 32, 16), (64, 100000, 16, 8), (256, 20000, 8, 4)]:
    P6 = params(Bk, D, O)
    y = torch.randn(Bk, Tk, O, device=dev, generator=g)
    ms, out = timed(lambda: fg.kalman_chain(*P6, y, validate=False), reps=2)
    print("kalman_chain (homogeneous) B=%d T=%d D=%d O=%d: %.1f ms, %.3e steps/s, c[0]=%.3f" % (Bk, Tk, D, O, ms, Bk * Tk / ms * 1e3, float(out[2][0])))
