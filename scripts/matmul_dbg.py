import sys, os
sys.path.insert(0, "/root/repo")
import torch, numpy as np
import funsor_b200 as fb
torch.set_printoptions(linewidth=200, precision=3)
def ref(x, y):
    return torch.logsumexp(x.double().unsqueeze(-1) + y.double().unsqueeze(-3), -2)
I=K=J=128
x = torch.zeros(1,1,I,K,device="cuda"); y = torch.zeros(1,1,K,J,device="cuda")
got = fb.semiring_matmul("logaddexp","add",x,y)[0,0]
print("zeros: expect", np.log(128), "got min/max", got.min().item(), got.max().item(), "finite frac", torch.isfinite(got).float().mean().item())
print(got[:4,:8]); print(got[60:64,60:68])
# structured: x[i,k] = log(i+1) if k==0 else -inf ; y[k,j] = log(j+1) if k==0 else -inf  -> out = log((i+1)(j+1))
x = torch.full((1,1,I,K), -1e30, device="cuda"); y = torch.full((1,1,K,J), -1e30, device="cuda")
x[0,0,:,5] = torch.log(torch.arange(1,I+1,device="cuda").float()); y[0,0,5,:] = torch.log(torch.arange(1,J+1,device="cuda").float())
got = fb.semiring_matmul("logaddexp","add",x,y)[0,0]
want = ref(x,y)[0,0]
print("rank1 err", (got.double()-want).abs().max().item()); print(got[:3,:6].exp()); print(want[:3,:6].exp())
