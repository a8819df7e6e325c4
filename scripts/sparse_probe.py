import sys, os
sys.path.insert(0, "/root/repo")
import torch, numpy as np
import funsor_b200 as fb
B, S, T = 3, 160, 64
A = torch.full((S, S), -float("inf"))
for i in range(S):
    A[i, i] = np.log(0.6); A[i, (i + 1) % S] = np.log(0.4)
g = torch.Generator().manual_seed(16)
init = torch.full((B, S), -float("inf")); init[:, 0] = 0.0
obs = torch.randn(B, T, S, generator=g) * 3
print(fb.hmm_forward(init.cuda(), A.cuda(), obs.cuda()).cpu().numpy())
