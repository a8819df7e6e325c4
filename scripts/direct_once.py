import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import funsor_b200 as fb
dev = torch.device("cuda")
B, T, S = 16, 8192, 512
g = torch.Generator(device=dev); g.manual_seed(1)
A = torch.log_softmax(torch.randn(S, S, device=dev, generator=g), -1)
init = torch.log_softmax(torch.randn(B, S, device=dev, generator=g), -1)
obs = torch.randn(B, T, S, device=dev, generator=g)
for _ in range(2):
    z = fb.hmm_forward(init, A, obs)
torch.cuda.synchronize()
print("ok", z[0].item())
