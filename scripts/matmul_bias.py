"""Signed mean error of the log-semiring matmul against float64, as a function of the contraction length K."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import funsor_b200 as fb

dev = torch.device("cuda")
g = torch.Generator(device=dev); g.manual_seed(3)
for K in (64, 128, 256, 512, 1024, 2048):
    x = torch.randn(4, 256, K, device=dev, generator=g)
    y = torch.randn(4, K, 256, device=dev, generator=g)
    ref = torch.logsumexp(x.double().unsqueeze(-1) + y.double().unsqueeze(-3), -2)
    z = fb.semiring_matmul("logaddexp", "add", x, y)
    d = (z.double() - ref)
    print("K=%5d mean signed err %.3e  rms %.3e  max %.3e" % (K, d.mean().item(), d.pow(2).mean().sqrt().item(), d.abs().max().item()))
x = torch.randn(64, 1024, 1024, device=dev, generator=g)
y = torch.randn(64, 1024, 1024, device=dev, generator=g)
for _ in range(2):
    z = fb.semiring_matmul("logaddexp", "add", x, y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    z = fb.semiring_matmul("logaddexp", "add", x, y)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("64x1024^3: %.3f ms  %.1f useful TFLOP/s" % (ms, 2 * 64 * 1024**3 / ms / 1e9))
