import sys
sys.path.insert(0, "/root/repo")
import torch
import funsor_b200.gaussian as fg
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
Bk, Tk, D, O = 64, 2048, 32, 16
F = 0.9 * torch.linalg.qr(torch.randn(Bk, D, D, device=dev, generator=g))[0]
H = torch.randn(Bk, D, O, device=dev, generator=g) / D ** 0.5
eye = lambda n: torch.eye(n, device=dev).expand(Bk, n, n).contiguous()
y = torch.randn(Bk, Tk, O, device=dev, generator=g)
w, P, s = fg.kalman_elements_sqrt(F, torch.zeros(Bk, D, device=dev), eye(D), H, torch.zeros(Bk, O, device=dev), eye(O), y)
Lam, eta, c = fg.sqrt_to_info(w, P, s)
for _ in range(2):
    out = fg.gaussian_sequential_sum_product(Lam, eta, c, validate=False); torch.cuda.synchronize()
print("ok", float(out[2][0]))
