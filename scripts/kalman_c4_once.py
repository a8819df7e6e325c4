import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import funsor_b200.gaussian as fg
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
Bk, Tk, D, O = 64, 100000, 32, 16
F = 0.9 * torch.linalg.qr(torch.randn(Bk, D, D, device=dev, generator=g))[0]
H = torch.randn(Bk, D, O, device=dev, generator=g) / D ** 0.5
eye = lambda n: torch.eye(n, device=dev).expand(Bk, n, n).contiguous()
y = torch.randn(Bk, Tk, O, device=dev, generator=g)
args = (F, torch.zeros(Bk, D, device=dev), eye(D), H, torch.zeros(Bk, O, device=dev), eye(O), y)
for _ in range(2):
    out = fg.kalman_chain(*args, validate=False)
torch.cuda.synchronize()
print("ok", float(out[2][0]))
