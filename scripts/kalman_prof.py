import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import funsor_b200.gaussian as fg
device = "cuda"
g = torch.Generator(device=device).manual_seed(0)
Bk, Tk, D, O = 64, 100000, 32, 16
F = 0.9 * torch.linalg.qr(torch.randn(Bk, D, D, device=device, generator=g))[0]
H = torch.randn(Bk, D, O, device=device, generator=g) / D ** 0.5
eye = lambda n: torch.eye(n, device=device).expand(Bk, n, n).contiguous()
y = torch.randn(Bk, Tk, O, device=device, generator=g)
kargs = (F, torch.zeros(Bk, D, device=device), eye(D), H, torch.zeros(Bk, O, device=device), eye(O), y)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
fg.kalman_chain(*kargs, validate=False); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): out = fg.kalman_chain(*kargs, validate=False)
e1.record(); torch.cuda.synchronize()
print("kalman C4 ms", e0.elapsed_time(e1) / reps, float(out[2].sum()))
